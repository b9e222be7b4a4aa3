"""The C-ABI library must load on a CPU-only machine and export every symbol include/zstdb200.h declares;
without a GPU its entry points must fail loudly (no CPU fallback) -- no compute calls are made here."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "zstdb200.h").read_text()
    return sorted(set(re.findall(r"ZSTDB200_API[^;]*?\b(ZSTD_\w+|zstdb200_\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    from zstd_jni_b200 import _native
    L = C.CDLL(str(_native.LIB_PATH))
    names = declared_symbols()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_host_side_helpers_work_without_gpu():
    from zstd_jni_b200 import _native, corpus
    from zstd_jni_b200.zstd import Zstd
    from tests.oracle_util import oracle_compress
    L = _native.lib()
    assert L.ZSTD_versionNumber() == 10507 and L.ZSTD_versionString() == b"1.5.7"
    for n in (0, 1, 255, 256, 131071, 131072, 1 << 20):
        assert Zstd.compressBound(n) == n + (n >> 8) + (((128 << 10) - n) >> 11 if n < (128 << 10) else 0)
    assert Zstd.isError((1 << 64) - 70) and not Zstd.isError(131072)
    assert Zstd.getErrorName((1 << 64) - 70) == "Destination buffer is too small"
    assert Zstd.getErrorName((1 << 64) - 20) == "Data corruption detected"
    assert Zstd.getErrorCode((1 << 64) - 72) == 72
    assert (Zstd.minCompressionLevel(), Zstd.maxCompressionLevel(), Zstd.defaultCompressionLevel()) == (-(1 << 17), 22, 3)
    data = corpus.chunk(1)[:30000].tobytes()
    z = oracle_compress(data, 3)
    assert Zstd.getFrameContentSize(z) == len(data)
    assert Zstd.findFrameCompressedSize(z + b"tail") == len(z)
    assert Zstd.getFrameContentSize(b"\x00" * 16) == -2


def test_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from zstd_jni_b200 import _native
    from zstd_jni_b200.zstd import Zstd, ZstdBatchContext, ZstdException
    L = _native.lib()
    assert L.zstdb200_device_count() == 0
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ZstdBatchContext(0)
    with pytest.raises(ZstdException) as ei:
        Zstd.compress(b"hello world, hello world, hello world", 3)
    assert ei.value.getErrorCode() == 1
    with pytest.raises(ZstdException):
        Zstd.decompress(bytes.fromhex("28b52ffd2000010000"), 0)


def test_stream_decoder_never_sizes_memory_from_the_frame_header():
    """A 17-byte frame promising 2^40 (or 2^64-1) bytes with one empty block: the streaming layer must answer with an error code --
    the reference streams in window-bounded memory -- instead of allocating from the header (no GPU needed: refused on the host)."""
    from zstd_jni_b200 import _native as N
    L = N.lib()
    for fcs in (1 << 40, (1 << 64) - 1):
        frame = bytes.fromhex("28b52ffd") + bytes([0xC0, 0x00]) + fcs.to_bytes(8, "little") + bytes([0x01, 0x00, 0x00])
        z = L.ZSTD_createDStream(); L.ZSTD_initDStream(z)
        src = C.create_string_buffer(frame, len(frame)); dst = C.create_string_buffer(1 << 16)
        ib = N.InBuffer(C.addressof(src), len(frame), 0); ob = N.OutBuffer(C.addressof(dst), len(dst), 0)
        r = L.ZSTD_decompressStream(z, C.byref(ob), C.byref(ib))
        assert N.is_error(r) and N.error_code(r) == 20, r        # corruption_detected
        L.ZSTD_freeDStream(z)
    # batch entry points refuse sizes whose sum wraps
    srcs = (C.c_void_p * 2)(C.addressof(src), C.addressof(src)); dsts = (C.c_void_p * 2)(C.addressof(dst), C.addressof(dst))
    ssz = (C.c_size_t * 2)(17, 17); caps = (C.c_size_t * 2)((1 << 64) - 1, 16); outs = (C.c_size_t * 2)()
    import torch
    if torch.cuda.is_available():
        ctx = L.zstdb200_create(0)
        r = L.zstdb200_decompress_batch(ctx, 2, srcs, ssz, dsts, caps, outs)
        assert N.is_error(r)
        L.zstdb200_free(ctx)


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    from zstd_jni_b200 import _native
    monkeypatch.setenv("ZSTDB200_LIBRARY", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_native, "_lib", None)
    with pytest.raises(_native.NativeLibraryMissing):
        _native.lib()
    monkeypatch.delenv("ZSTDB200_LIBRARY")
    monkeypatch.setattr(_native, "_lib", None)
    _native.lib()
