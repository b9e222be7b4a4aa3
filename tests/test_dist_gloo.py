"""N>1 host logic on CPU: two gloo ranks shard a batch, "compress" their shard with the CPU checker (the GPU is
not needed to test the plumbing), all_gather the sizes and rebuild the global stream index; the stitched stream
must equal the single-process result and decode back."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, n_items, tmp):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from zstd_jni_b200 import corpus, sharding
    from tests.oracle_util import oracle_compress
    s, e = sharding.shard_range(n_items, rank, world)
    frames = [oracle_compress(corpus.chunk(i)[:20000].tobytes(), 3) for i in range(s, e)]
    local = torch.tensor([len(f) for f in frames], dtype=torch.int64)
    sizes = sharding.gather_sizes(local, n_items)
    offs = sharding.global_offsets(sizes)
    lo, hi = sharding.rank_byte_range(offs, n_items, rank, world)
    blob = b"".join(frames)
    assert hi - lo == len(blob)
    np.save(os.path.join(tmp, f"part{rank}.npy"), np.frombuffer(blob, dtype=np.uint8))
    if rank == 0:
        np.save(os.path.join(tmp, "offsets.npy"), offs.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _plane_worker(rank, world, port, n_items, tmp):
    """The root holds the whole batch; scatter -> "compress" the shard (CPU checker) -> size all_gather + scan -> gatherv of the
    packed frames to the root -> the root scatters the frames' byte ranges back, every rank "decompresses", gather_fixed."""
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from zstd_jni_b200 import corpus, sharding
    from tests.oracle_util import oracle_compress, oracle_decompress
    item = 20000
    batch = None
    if rank == 0:
        batch = torch.from_numpy(np.concatenate([corpus.chunk(i)[:item] for i in range(n_items)]).copy())
    s, e = sharding.shard_range(n_items, rank, world)
    mine = torch.zeros((e - s) * item + 16, dtype=torch.uint8)
    assert sharding.scatter_chunks(batch, n_items, item, mine) == (s, e)
    frames = [oracle_compress(mine[k * item:(k + 1) * item].numpy().tobytes(), 3) for k in range(e - s)]
    local_sizes = torch.tensor([len(f) for f in frames], dtype=torch.int64)
    offs = sharding.global_offsets(sharding.gather_sizes(local_sizes, n_items))
    ranges = sharding.rank_byte_ranges(offs, n_items, world)
    packed = torch.from_numpy(np.frombuffer(b"".join(frames), dtype=np.uint8).copy()) if frames else torch.zeros(0, dtype=torch.uint8)
    stream = torch.zeros(int(offs[-1]), dtype=torch.uint8) if rank == 0 else None
    sharding.gatherv_bytes(packed, ranges, stream)
    # decompression direction: fixed-size outputs come back in order
    back_local = torch.from_numpy(np.frombuffer(b"".join(oracle_decompress(f, item) for f in frames), dtype=np.uint8).copy()) if frames else torch.zeros(0, dtype=torch.uint8)
    back = torch.zeros(n_items * item, dtype=torch.uint8) if rank == 0 else None
    sharding.gather_fixed(back_local, n_items, item, back)
    if rank == 0:
        np.save(os.path.join(tmp, "stream.npy"), stream.numpy()); np.save(os.path.join(tmp, "back.npy"), back.numpy()); np.save(os.path.join(tmp, "offs.npy"), offs.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_scatter_compress_gatherv(tmp_path):
    """SURVEY.md 8(e) items 1-3 on two gloo ranks: the stream gathered on the root is byte-identical to the single-process one."""
    n_items, world = 9, 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_plane_worker, args=(world, port, n_items, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, str(ROOT))
    from zstd_jni_b200 import corpus
    from tests.oracle_util import oracle_compress
    expect = [oracle_compress(corpus.chunk(i)[:20000].tobytes(), 3) for i in range(n_items)]
    assert np.load(tmp_path / "stream.npy").tobytes() == b"".join(expect)
    assert list(np.diff(np.load(tmp_path / "offs.npy"))) == [len(f) for f in expect]
    assert np.load(tmp_path / "back.npy").tobytes() == b"".join(corpus.chunk(i)[:20000].tobytes() for i in range(n_items))


def test_shard_range_is_a_partition():
    from zstd_jni_b200 import sharding
    for n in (0, 1, 7, 8, 8192, 10000):
        for w in (1, 2, 3, 8):
            r = [sharding.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            assert max(e - s for s, e in r) - min(e - s for s, e in r) <= 1


def test_two_rank_gloo_stream_index(tmp_path):
    n_items, world = 11, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_items, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, str(ROOT))
    from zstd_jni_b200 import corpus
    from tests.oracle_util import oracle_compress, oracle_decompress
    stream = np.concatenate([np.load(tmp_path / f"part{r}.npy") for r in range(world)]).tobytes()
    offs = np.load(tmp_path / "offsets.npy")
    expect = [oracle_compress(corpus.chunk(i)[:20000].tobytes(), 3) for i in range(n_items)]
    assert stream == b"".join(expect)
    assert list(np.diff(offs)) == [len(f) for f in expect]
    for i in (0, 5, 10):
        assert stream[offs[i]:offs[i + 1]] == expect[i]
    assert oracle_decompress(stream, n_items * 20000) == b"".join(corpus.chunk(i)[:20000].tobytes() for i in range(n_items))
