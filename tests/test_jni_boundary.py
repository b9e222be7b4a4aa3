"""The drop-in boundary itself (BASELINE.json configs[0]): the reference's UNMODIFIED JNI glue
(N/jni_zstd.c, N/jni_fast_zstd.c, ... compiled where they lie by `make -C oracle jni`) linked against libzstdb200.so, driven
without a JVM through a fake JNIEnv (oracle/jni_harness.c).  The Java_com_github_luben_zstd_* symbols are exactly what the Java
classes bind; every ZSTD_* call on the hot path inside them lands in the product library.

CPU: the library loads with all its symbols, host-side entry points answer like the reference, and the compute entry points fail
loudly (ZSTD_error_GENERIC) because there is no device -- the CPU libzstd linked behind it for the cold path is NOT reached.
GPU (-m gpu): Zstd.compressUnsafe / decompressUnsafe and ZstdCompressCtx / ZstdDecompressCtx byte-array calls round-trip one
128 KB block at level 3, byte-identical to the oracle.
"""
import ctypes as C
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
JNI_LIB = ROOT / "oracle" / "_ref" / "libzstd-jni-b200.so"
HARNESS = ROOT / "oracle" / "_ref" / "libjniharness.so"
P = "Java_com_github_luben_zstd_"
jl, ji, jb, vp = C.c_longlong, C.c_int, C.c_ubyte, C.c_void_p


@pytest.fixture(scope="module")
def jni():
    if not JNI_LIB.exists() or not HARNESS.exists():
        pytest.skip("oracle/_ref/libzstd-jni-b200.so not built (needs /root/reference: make -C oracle jni)")
    H = C.CDLL(str(HARNESS))
    H.jh_env.restype = vp
    H.jh_new_array.restype = vp; H.jh_new_array.argtypes = [ji]
    H.jh_array_data.restype = vp; H.jh_array_data.argtypes = [vp]
    H.jh_array_pinned.restype = ji; H.jh_array_pinned.argtypes = [vp]
    H.jh_free.argtypes = [vp]
    L = C.CDLL(str(JNI_LIB))

    def sig(name, res, *args):
        f = getattr(L, P + name)
        f.restype = res
        f.argtypes = [vp, vp, *args]          # JNIEnv*, jclass
        return f

    class J:
        env = H.jh_env()
        harness = H
        compressBound = sig("Zstd_compressBound", jl, jl)
        isError = sig("Zstd_isError", jb, jl)
        getErrorName = sig("Zstd_getErrorName", C.c_char_p, jl)           # the harness' NewStringUTF hands the C string through
        getErrorCode = sig("Zstd_getErrorCode", jl, jl)
        errDstSizeTooSmall = sig("Zstd_errDstSizeTooSmall", jl)
        errCorruptionDetected = sig("Zstd_errCorruptionDetected", jl)
        magicNumber = sig("Zstd_magicNumber", ji)
        windowLogMax = sig("Zstd_windowLogMax", ji)
        compressUnsafe = sig("Zstd_compressUnsafe", jl, jl, jl, jl, jl, ji, jb)
        decompressUnsafe = sig("Zstd_decompressUnsafe", jl, jl, jl, jl, jl)
        getFrameContentSize0 = sig("Zstd_getFrameContentSize0", jl, vp, ji, ji, jb)
        decompressedSize0 = sig("Zstd_decompressedSize0", jl, vp, ji, ji, jb)
        findFrameCompressedSize0 = sig("Zstd_findFrameCompressedSize0", jl, vp, ji, ji)
        getDictIdFromFrame = sig("Zstd_getDictIdFromFrame", jl, vp)
        cctxInit = sig("ZstdCompressCtx_init", jl)
        cctxFree = sig("ZstdCompressCtx_free", None, jl)
        cctxSetLevel = sig("ZstdCompressCtx_setLevel0", None, jl, ji)
        cctxSetChecksum = sig("ZstdCompressCtx_setChecksum0", None, jl, jb)
        cctxReset = sig("ZstdCompressCtx_reset0", jl, jl)
        cctxCompressByteArray = sig("ZstdCompressCtx_compressByteArray0", jl, jl, vp, ji, ji, vp, ji, ji)
        dctxInit = sig("ZstdDecompressCtx_init", jl)
        dctxFree = sig("ZstdDecompressCtx_free", None, jl)
        dctxDecompressByteArray = sig("ZstdDecompressCtx_decompressByteArray0", jl, jl, vp, ji, ji, vp, ji, ji)
        loadDictCompress = sig("Zstd_loadDictCompress", ji, jl, vp, ji)
        loadDictDecompress = sig("Zstd_loadDictDecompress", ji, jl, vp, ji)
        registerSequenceProducer = sig("Zstd_registerSequenceProducer", None, jl, jl, jl)
        setMagicless = sig("Zstd_setCompressionMagicless", ji, jl, jb)
        setHashLog = sig("Zstd_setCompressionHashLog", ji, jl, ji)

        @staticmethod
        def array(data: bytes = b"", size: int = None):
            n = len(data) if size is None else size
            a = H.jh_new_array(n)
            if data:
                C.memmove(H.jh_array_data(a), data, len(data))
            return a

        @staticmethod
        def bytes_of(a, n):
            return C.string_at(H.jh_array_data(a), n)

    return J


def _err(code):          # jlong -> libzstd error number (0 if none)
    u = code & ((1 << 64) - 1)
    return (1 << 64) - u if u > (1 << 64) - 120 else 0


def test_jni_glue_exports_and_links_against_the_product_library(jni):
    import subprocess
    syms = subprocess.run(["nm", "-D", "--defined-only", str(JNI_LIB)], capture_output=True, text=True, check=True).stdout
    names = [l.split()[-1] for l in syms.splitlines() if " T " in l and P in l]
    assert len(names) == 149, len(names)                              # the 149 Java_* entry points
    needed = subprocess.run(["readelf", "-d", str(JNI_LIB)], capture_output=True, text=True, check=True).stdout
    assert needed.index("libzstdb200.so") < needed.index("libzstd-cold.so")       # hot-path symbols resolve to the product library first


def test_jni_host_side_entry_points(jni):
    from tests.oracle_util import oracle_compress, oracle_compress_flags
    from zstd_jni_b200 import corpus
    J = jni
    assert J.compressBound(J.env, None, 131072) == 131584
    assert J.isError(J.env, None, -70) == 1 and J.isError(J.env, None, 1000) == 0
    assert J.getErrorName(J.env, None, -70) == b"Destination buffer is too small" and J.getErrorCode(J.env, None, -72) == 72
    assert J.errDstSizeTooSmall(J.env, None) == 70 and J.errCorruptionDetected(J.env, None) == 20
    assert J.magicNumber(J.env, None) == 0xFD2FB528 - (1 << 32) and J.windowLogMax(J.env, None) == 31
    data = corpus.chunk(1)[:30000].tobytes()
    z = oracle_compress_flags(data, 3, checksum=True)
    a = J.array(b"xx" + z + b"tail")
    try:
        assert J.getFrameContentSize0(J.env, None, a, 2, len(z), 0) == 30000
        assert J.decompressedSize0(J.env, None, a, 2, len(z), 0) == 30000
        assert J.getFrameContentSize0(J.env, None, a, 6, len(z) - 4, 1) == 30000            # magicless view of the same frame (N/jni_zstd.c:32-40)
        assert J.findFrameCompressedSize0(J.env, None, a, 2, len(z) + 4) == len(z)
        assert J.harness.jh_array_pinned(a) == 0                                            # every critical section was released
    finally:
        J.harness.jh_free(a)
    a = J.array(z)
    try:
        assert J.getDictIdFromFrame(J.env, None, a) == 0
    finally:
        J.harness.jh_free(a)
    c = J.cctxInit(J.env, None)
    assert c != 0
    J.cctxSetLevel(J.env, None, c, 5)
    J.cctxSetChecksum(J.env, None, c, 1)
    assert J.setMagicless(J.env, None, c, 1) == 1 and J.setHashLog(J.env, None, c, 12) == 12      # ZSTD_CCtx_setParameter returns the value set
    assert _err(J.setHashLog(J.env, None, c, 31)) == 42                                     # out of bounds like the reference
    assert J.cctxReset(J.env, None, c) == 0
    J.cctxFree(J.env, None, c)


def test_python_mirror_constants_equal_the_glue_s(jni):
    """J/Zstd.java's constant getters: the Python mirror must return what the reference's glue (compiled from the reference's headers) returns."""
    from zstd_jni_b200.zstd import Zstd
    L = C.CDLL(str(JNI_LIB))
    for name in list(Zstd._ERR) + ["magicNumber", "blockSizeMax", "windowLogMin", "windowLogMax", "chainLogMin", "chainLogMax", "hashLogMin", "hashLogMax",
                                   "searchLogMin", "searchLogMax"]:      # searchLengthMin / Max are declared in J/Zstd.java:1108-1109 but have no native in N/jni_zstd.c
        f = getattr(L, P + "Zstd_" + name)
        f.restype = jl if name.startswith("err") else ji
        f.argtypes = [vp, vp]
        assert f(jni.env, None) == getattr(Zstd, name)(), name


def test_jni_unbuilt_features_refuse_loudly(jni):
    """Dictionaries and foreign sequence producers are not built.  Their context-taking entry points are exported by libzstdb200
    (the contexts are this library's objects; a CPU libzstd behind must never see them) and refuse with parameter_unsupported."""
    import subprocess
    J = jni
    undefined = subprocess.run(["nm", "-D", "--undefined-only", str(JNI_LIB)], capture_output=True, text=True, check=True).stdout.split()
    from zstd_jni_b200 import _native as N
    ours = subprocess.run(["nm", "-D", "--defined-only", str(N.LIB_PATH)], capture_output=True, text=True, check=True).stdout.split()
    cold = sorted(x for x in undefined if x.startswith(("ZSTD_", "ZDICT_")) and x not in ours)
    # what is left to the CPU library takes no ZSTD_CCtx / ZSTD_DCtx: dictionary objects and the dictionary trainer
    assert cold == ["ZDICT_trainFromBuffer", "ZDICT_trainFromBuffer_legacy", "ZSTD_createCDict", "ZSTD_createCDict_byReference", "ZSTD_createDDict",
                    "ZSTD_createDDict_byReference", "ZSTD_freeCDict", "ZSTD_freeDDict", "ZSTD_getDictID_fromDict"], cold
    c, d = J.cctxInit(J.env, None), J.dctxInit(J.env, None)
    a = J.array(b"a dictionary of sorts " * 20)
    dst = J.array(size=1000)
    src = J.array(b"hello hello hello hello hello hello")
    try:
        assert _err(J.loadDictCompress(J.env, None, c, a, 440)) == 40 and _err(J.loadDictDecompress(J.env, None, d, a, 440)) == 40
        assert J.loadDictCompress(J.env, None, c, a, 0) == 0                              # clearing the dictionary is fine
        J.registerSequenceProducer(J.env, None, c, 0, 0x1234)                           # a foreign match finder would bypass the GPU parsers
        assert _err(J.cctxCompressByteArray(J.env, None, c, dst, 0, 1000, src, 0, 35)) == 40
        J.registerSequenceProducer(J.env, None, c, 0, 0)
        assert _err(J.cctxCompressByteArray(J.env, None, c, dst, 0, 1000, src, 0, 35)) in (0, 1)      # accepted again (GENERIC without a device)
    finally:
        for x in (a, dst, src):
            J.harness.jh_free(x)
        J.cctxFree(J.env, None, c)
        J.dctxFree(J.env, None, d)


def test_jni_compute_entry_points_fail_loudly_without_a_device(jni):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    J = jni
    src = C.create_string_buffer(b"hello world " * 1000)
    dst = C.create_string_buffer(20000)
    r = J.compressUnsafe(J.env, None, C.addressof(dst), 20000, C.addressof(src), 12000, 3, 0)
    assert _err(r) == 1                          # ZSTD_error_GENERIC from libzstdb200 -- the CPU libzstd behind it would have compressed
    frame = bytes.fromhex("28b52ffd2000010000")
    s2 = C.create_string_buffer(frame)
    assert _err(J.decompressUnsafe(J.env, None, C.addressof(dst), 20000, C.addressof(s2), len(frame))) == 1


@pytest.mark.gpu
def test_gpu_jni_round_trip_one_block(jni):
    """BASELINE.json configs[0]: Zstd.compress / decompress round trip of one 128 KB block at level 3 through the JNI entry points."""
    from tests.oracle_util import oracle_compress, oracle_compress_flags
    from zstd_jni_b200 import corpus
    J = jni
    data = corpus.chunk(0).tobytes()
    src = C.create_string_buffer(data, len(data))
    cap = J.compressBound(J.env, None, len(data))
    dst = C.create_string_buffer(cap)
    n = J.compressUnsafe(J.env, None, C.addressof(dst), cap, C.addressof(src), len(data), 3, 0)
    assert _err(n) == 0 and dst.raw[:n] == oracle_compress(data, 3)
    back = C.create_string_buffer(len(data))
    assert J.decompressUnsafe(J.env, None, C.addressof(back), len(data), C.addressof(dst), n) == len(data) and back.raw == data
    n2 = J.compressUnsafe(J.env, None, C.addressof(dst), cap, C.addressof(src), len(data), 3, 1)
    assert dst.raw[:n2] == oracle_compress_flags(data, 3, checksum=True)
    assert _err(J.compressUnsafe(J.env, None, C.addressof(dst), 100, C.addressof(src), len(data), 3, 0)) == 70
    # the context classes over byte arrays with offsets (N/jni_fast_zstd.c:615-646, 807-836)
    c, d = J.cctxInit(J.env, None), J.dctxInit(J.env, None)
    a_src, a_dst, a_back = J.array(b"pad" + data), J.array(size=cap + 10), J.array(size=len(data) + 5)
    try:
        J.cctxSetLevel(J.env, None, c, 1)
        n = J.cctxCompressByteArray(J.env, None, c, a_dst, 10, cap, a_src, 3, len(data))
        assert _err(n) == 0 and J.bytes_of(a_dst, 10 + n)[10:] == oracle_compress(data, 1)
        m = J.dctxDecompressByteArray(J.env, None, d, a_back, 5, len(data), a_dst, 10, n)
        assert m == len(data) and J.bytes_of(a_back, 5 + m)[5:] == data
        assert _err(J.dctxDecompressByteArray(J.env, None, d, a_back, 5, len(data) - 1, a_dst, 10, n)) == 70
        assert _err(J.cctxCompressByteArray(J.env, None, c, a_dst, 10, cap + 1, a_src, 3, len(data))) == 70     # the glue's own bounds check (:619-624)
        assert J.harness.jh_array_pinned(a_src) == 0 and J.harness.jh_array_pinned(a_dst) == 0
    finally:
        for a in (a_src, a_dst, a_back):
            J.harness.jh_free(a)
        J.cctxFree(J.env, None, c)
        J.dctxFree(J.env, None, d)
