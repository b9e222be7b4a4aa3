"""Python mirror of com.github.luben.zstd's hot-path API, over the C ABI of libzstdb200.so.

No JVM is available in the build image, so this module plays the role of the
reference's Java layer (J/ = src/main/java/com/github/luben/zstd/): same class and
method names, argument meaning and error behaviour, so that the parity tests read
like the reference's own Scala tests (src/test/scala/Zstd.scala).  Everything here is
plumbing: every byte is produced by the CUDA kernels behind the C ABI.

  Zstd.compress / decompress / compressBound / isError / getErrorName ...  J/Zstd.java
  ZstdCompressCtx / ZstdDecompressCtx                                      J/ZstdCompressCtx.java, J/ZstdDecompressCtx.java
  ZstdOutputStream / ZstdInputStream                                       J/ZstdOutputStreamNoFinalizer.java, J/ZstdInputStreamNoFinalizer.java
  ZstdException                                                            J/ZstdException.java
plus the batch entry points (Zstd.compressBatch / decompressBatch / ZstdBatchContext)
that a JNI maintainer would add next to them.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Sequence

import numpy as np

from . import _native as N

ZSTD_c_compressionLevel = 100
ZSTD_c_contentSizeFlag = 200
ZSTD_c_checksumFlag = 201
ZSTD_c_dictIDFlag = 202
ZSTD_c_format = 10
ZSTD_d_windowLogMax = 100
ZSTD_d_format = 1000
ZSTD_e_continue, ZSTD_e_flush, ZSTD_e_end = 0, 1, 2
BLOCK = 131072


class ZstdException(RuntimeError):
    """J/ZstdException.java:16-32 : carries the libzstd error code and its name."""

    def __init__(self, code: int, message: str | None = None):
        self.code = code
        super().__init__(message if message is not None else Zstd.getErrorName(code))

    def getErrorCode(self) -> int:
        return self.code


def _buf(b) -> tuple[C.c_void_p, int, object]:
    """(address, length, keepalive) for bytes / bytearray / numpy uint8."""
    if isinstance(b, np.ndarray):
        a = np.ascontiguousarray(b, dtype=np.uint8)
        return C.c_void_p(a.ctypes.data), a.size, a
    if isinstance(b, (bytes, bytearray, memoryview)):
        a = np.frombuffer(b, dtype=np.uint8)
        return C.c_void_p(a.ctypes.data if a.size else 0), a.size, a
    raise TypeError(f"unsupported buffer type {type(b)}")


class Zstd:
    """Static facade, J/Zstd.java."""

    # J/Zstd.java:929-951 (errNoError ... errDstBufferNull): the ZSTD_ErrorCode enum of N/zstd_errors.h
    _ERR = {"errNoError": 0, "errGeneric": 1, "errPrefixUnknown": 10, "errVersionUnsupported": 12, "errFrameParameterUnsupported": 14,
            "errFrameParameterWindowTooLarge": 16, "errCorruptionDetected": 20, "errChecksumWrong": 22, "errDictionaryCorrupted": 30,
            "errDictionaryWrong": 32, "errDictionaryCreationFailed": 34, "errParameterUnsupported": 40, "errParameterOutOfBound": 42,
            "errTableLogTooLarge": 44, "errMaxSymbolValueTooLarge": 46, "errMaxSymbolValueTooSmall": 48, "errStageWrong": 60, "errInitMissing": 62,
            "errMemoryAllocation": 64, "errWorkSpaceTooSmall": 66, "errDstSizeTooSmall": 70, "errSrcSizeWrong": 72, "errDstBufferNull": 74}

    @staticmethod
    def isError(code: int) -> bool:                      # J/Zstd.java:isError
        return bool(N.lib().ZSTD_isError(code & ((1 << 64) - 1)))

    @staticmethod
    def getErrorName(code: int) -> str:
        c = code if code > N.ERROR_MAX else ((1 << 64) - abs(code)) if code else 0
        return N.lib().ZSTD_getErrorName(c).decode()

    @staticmethod
    def getErrorCode(code: int) -> int:
        return N.lib().ZSTD_getErrorCode(code & ((1 << 64) - 1))

    @staticmethod
    def compressBound(srcSize: int) -> int:              # J/Zstd.java:compressBound -> N/jni_zstd.c
        return N.lib().ZSTD_compressBound(srcSize)

    @staticmethod
    def minCompressionLevel() -> int:
        return N.lib().ZSTD_minCLevel()

    @staticmethod
    def maxCompressionLevel() -> int:
        return N.lib().ZSTD_maxCLevel()

    @staticmethod
    def defaultCompressionLevel() -> int:
        return N.lib().ZSTD_defaultCLevel()

    @staticmethod
    def compress(src, level: int = 3) -> bytes:          # J/Zstd.java:1137-1145
        with ZstdCompressCtx() as ctx:
            ctx.setLevel(level)
            return ctx.compress(src)

    @staticmethod
    def compressInto(dst: bytearray, src, level: int = 3) -> int:   # compress(byte[] dst, byte[] src, int level) -> long
        with ZstdCompressCtx() as ctx:
            ctx.setLevel(level)
            return ctx.compressByteArray(dst, 0, len(dst), src, 0, len(src), raise_on_error=False)

    @staticmethod
    def decompress(src, originalSize: int) -> bytes:     # J/Zstd.java:1417-1424
        with ZstdDecompressCtx() as ctx:
            return ctx.decompress(src, originalSize)

    @staticmethod
    def decompressInto(dst: bytearray, src) -> int:      # decompress(byte[] dst, byte[] src) -> long (error code, not exception)
        with ZstdDecompressCtx() as ctx:
            return ctx.decompressByteArray(dst, 0, len(dst), src, 0, len(src), raise_on_error=False)

    @staticmethod
    def getFrameContentSize(src, magicless: bool = False) -> int:   # J/Zstd.java:729-739 ; -1 unknown, -2 error like the C API
        p, n, _k = _buf(src)
        if magicless:                                    # N/jni_zstd.c:32-40: header parse in the magicless format, 0 on any failure
            fh = N.FrameHeader()
            if N.lib().ZSTD_getFrameHeader_advanced(C.byref(fh), p, n, 1) != 0:
                return 0
            v = fh.frameContentSize
        else:
            v = N.lib().ZSTD_getFrameContentSize(p, n)
        return v if v < (1 << 63) else v - (1 << 64)

    @staticmethod
    def decompressedSize(src, magicless: bool = False) -> int:      # J/Zstd.java:754-764 (deprecated name); 0 when unknown / error
        v = Zstd.getFrameContentSize(src, magicless)
        return v if v >= 0 else 0

    @staticmethod
    def getDictIdFromFrame(src) -> int:                  # J/Zstd.java:getDictIdFromFrame -> N/jni_zstd.c:139 (0: none / undecodable)
        p, n, _k = _buf(src)
        return N.lib().ZSTD_getDictID_fromFrame(p, n)

    @staticmethod
    def getFrameHeader(src, magicless: bool = False) -> dict:
        """ZSTD_getFrameHeader_advanced as a dict (not in J/Zstd.java; what N/jni_zstd.c:32-40 looks at)."""
        p, n, _k = _buf(src)
        fh = N.FrameHeader()
        r = N.lib().ZSTD_getFrameHeader_advanced(C.byref(fh), p, n, 1 if magicless else 0)
        if N.is_error(r):
            raise ZstdException(N.error_code(r), N.lib().ZSTD_getErrorName(r).decode())
        if r:
            raise ZstdException(72, f"Src size is incorrect (header needs {r} bytes)")
        return {k: getattr(fh, k) for k, _t in N.FrameHeader._fields_ if not k.startswith("_")}

    @staticmethod
    def findFrameCompressedSize(src) -> int:
        p, n, _k = _buf(src)
        r = N.lib().ZSTD_findFrameCompressedSize(p, n)
        if N.is_error(r):
            raise ZstdException(N.error_code(r), N.lib().ZSTD_getErrorName(r).decode())
        return r

    # ---- constants of the format / parameter space (J/Zstd.java:929-951,1099-1110 -> N/jni_zstd.c:573-667: header macros and error enum)
    @staticmethod
    def magicNumber() -> int:
        return 0xFD2FB528 - (1 << 32)                    # a Java int

    @staticmethod
    def blockSizeMax() -> int:
        return BLOCK

    @staticmethod
    def windowLogMin() -> int:
        return 10

    @staticmethod
    def windowLogMax() -> int:
        return 31

    @staticmethod
    def chainLogMin() -> int:
        return 6

    @staticmethod
    def chainLogMax() -> int:
        return 30

    @staticmethod
    def hashLogMin() -> int:
        return 6

    @staticmethod
    def hashLogMax() -> int:
        return 30

    @staticmethod
    def searchLogMin() -> int:
        return 1

    @staticmethod
    def searchLogMax() -> int:
        return 30

    @staticmethod
    def searchLengthMin() -> int:                        # ZSTD_MINMATCH_MIN (declared in J/Zstd.java:1108 without a native in N/jni_zstd.c)
        return 3

    @staticmethod
    def searchLengthMax() -> int:
        return 7

    # ---- batch entry points (new surface)
    @staticmethod
    def compressBatch(chunks: Sequence, level: int = 3) -> List[bytes]:
        with ZstdBatchContext() as b:
            return b.compressBatch(chunks, level)

    @staticmethod
    def decompressBatch(frames: Sequence, originalSizes: Sequence[int]) -> List[bytes]:
        with ZstdBatchContext() as b:
            return b.decompressBatch(frames, originalSizes)


class _AutoClose:
    """J/AutoCloseBase.java : close() is idempotent, use-after-close raises."""
    _ptr = None

    def _live(self):
        if self._ptr is None:
            raise RuntimeError("Closed")          # IllegalStateException("Closed") in Java
        return self._ptr

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ZstdCompressCtx(_AutoClose):
    """J/ZstdCompressCtx.java (subset on the hot path)."""

    def __init__(self):
        self._ptr = N.lib().ZSTD_createCCtx()
        if not self._ptr:
            raise MemoryError("ZSTD_createCCtx failed")

    def close(self):
        if self._ptr is not None:
            N.lib().ZSTD_freeCCtx(self._ptr)
            self._ptr = None

    def _set(self, param: int, value: int):
        r = N.lib().ZSTD_CCtx_setParameter(self._live(), param, value)
        if N.is_error(r):
            raise ZstdException(N.error_code(r), N.lib().ZSTD_getErrorName(r).decode())
        return self

    def setLevel(self, level: int):                      # :69-75
        return self._set(ZSTD_c_compressionLevel, level)

    def setMultiFrame(self, flag: bool):
        """Extension (ZSTDB200_c_multiFrame): inputs > 128 KB become one independent frame per 128 KB instead of an error.
        The result is a legal zstd stream but not the reference's bytes."""
        return self._set(0xB200, int(flag))

    def setChecksum(self, flag: bool):
        return self._set(ZSTD_c_checksumFlag, int(flag))

    def setContentSize(self, flag: bool):
        return self._set(ZSTD_c_contentSizeFlag, int(flag))

    def setDictID(self, flag: bool):
        return self._set(ZSTD_c_dictIDFlag, int(flag))

    # explicit compression parameters (J/ZstdCompressCtx.java setWindowLog ... setStrategy -> N/jni_fast_zstd.c / N/jni_zstd.c
    # setCompression*): applied over the level's row like the reference; 0 restores the level's value
    def setWindowLog(self, v: int):
        return self._set(101, v)

    def setHashLog(self, v: int):
        return self._set(102, v)

    def setChainLog(self, v: int):
        return self._set(103, v)

    def setSearchLog(self, v: int):
        return self._set(104, v)

    def setMinMatch(self, v: int):
        return self._set(105, v)

    def setTargetLength(self, v: int):
        return self._set(106, v)

    def setStrategy(self, v: int):
        return self._set(107, v)

    def setMagicless(self, flag: bool):                  # :84-90 -> N/jni_zstd.c:362-363 (ZSTD_c_format)
        return self._set(ZSTD_c_format, int(flag))

    def getFrameProgression(self) -> dict:               # :477-480 -> N/jni_fast_zstd.c:373 (J/ZstdFrameProgression.java fields)
        fp = N.lib().ZSTD_getFrameProgression(self._live())
        return {k: getattr(fp, k) for k, _t in N.FrameProgression._fields_}

    def reset(self):
        N.lib().ZSTD_CCtx_reset(self._live(), 3)

    def compressByteArray(self, dst: bytearray, dstOffset: int, dstSize: int, src, srcOffset: int, srcSize: int, raise_on_error=True) -> int:
        """:691-710 ; bounds are checked the way N/jni_fast_zstd.c:615-624 does."""
        self._live()
        if dstOffset < 0 or dstSize < 0 or dstOffset + dstSize > len(dst):
            raise IndexError("dst range")
        sp, sn, _k = _buf(src)
        if srcOffset < 0 or srcSize < 0 or srcOffset + srcSize > sn:
            raise IndexError("src range")
        d = (C.c_char * len(dst)).from_buffer(dst)
        L = N.lib()
        L.ZSTD_CCtx_reset(self._ptr, 1)
        r = L.ZSTD_compress2(self._ptr, C.c_void_p(C.addressof(d) + dstOffset), dstSize, C.c_void_p((sp.value or 0) + srcOffset), srcSize)
        if N.is_error(r) and raise_on_error:
            raise ZstdException(N.error_code(r), L.ZSTD_getErrorName(r).decode())
        return r

    def compress(self, src) -> bytes:                    # :784-792
        _p, n, _k = _buf(src)
        dst = bytearray(max(Zstd.compressBound(n), 1))
        size = self.compressByteArray(dst, 0, len(dst), src, 0, n)
        return bytes(dst[:size])


class ZstdDecompressCtx(_AutoClose):
    """J/ZstdDecompressCtx.java (subset on the hot path)."""

    def __init__(self):
        self._ptr = N.lib().ZSTD_createDCtx()
        if not self._ptr:
            raise MemoryError("ZSTD_createDCtx failed")

    def close(self):
        if self._ptr is not None:
            N.lib().ZSTD_freeDCtx(self._ptr)
            self._ptr = None

    def setMagicless(self, flag: bool):                  # J/ZstdDecompressCtx.java:54-60 -> N/jni_zstd.c:413-414 (ZSTD_d_format)
        r = N.lib().ZSTD_DCtx_setParameter(self._live(), ZSTD_d_format, int(flag))
        if N.is_error(r):
            raise ZstdException(N.error_code(r), N.lib().ZSTD_getErrorName(r).decode())
        return self

    def reset(self):
        N.lib().ZSTD_DCtx_reset(self._live(), 3)

    def decompressByteArray(self, dst: bytearray, dstOffset: int, dstSize: int, src, srcOffset: int, srcSize: int, raise_on_error=True) -> int:
        self._live()
        if dstOffset < 0 or dstSize < 0 or dstOffset + dstSize > len(dst):
            raise IndexError("dst range")
        sp, sn, _k = _buf(src)
        if srcOffset < 0 or srcSize < 0 or srcOffset + srcSize > sn:
            raise IndexError("src range")
        L = N.lib()
        L.ZSTD_DCtx_reset(self._ptr, 1)
        if len(dst):
            d = (C.c_char * len(dst)).from_buffer(dst)
            dp = C.c_void_p(C.addressof(d) + dstOffset)
        else:
            dp = C.c_void_p(0)
        r = L.ZSTD_decompressDCtx(self._ptr, dp, dstSize, C.c_void_p((sp.value or 0) + srcOffset), srcSize)
        if N.is_error(r) and raise_on_error:
            raise ZstdException(N.error_code(r), L.ZSTD_getErrorName(r).decode())
        return r

    def decompress(self, src, originalSize: int) -> bytes:      # :381-394
        if originalSize < 0:
            raise ZstdException(72, "Src size is incorrect")
        dst = bytearray(originalSize)
        _p, n, _k = _buf(src)
        size = self.decompressByteArray(dst, 0, originalSize, src, 0, n)
        return bytes(dst[:size])


class ZstdBatchContext(_AutoClose):
    """Owner of a zstdb200_ctx: GPU workspaces + stream for batch calls (one per thread)."""

    def __init__(self, device: int = -1):
        L = N.lib()
        self._ptr = L.zstdb200_create(device)
        if not self._ptr:
            self._ptr = None
            raise RuntimeError("zstdb200_create failed (no CPU fallback): " + L.zstdb200_last_error().decode())

    def close(self):
        if self._ptr is not None:
            N.lib().zstdb200_free(self._ptr)
            self._ptr = None

    @property
    def handle(self):
        return self._live()

    def kernelLaunches(self) -> int:
        return N.lib().zstdb200_kernel_launches(self._live())

    def kernelTimes(self) -> dict:
        """{kernel: (average ms, launches)} since the previous call (needs setOption("timing", 1))."""
        buf = C.create_string_buffer(4096)
        N.lib().zstdb200_kernel_times(self._live(), buf, len(buf))
        out = {}
        for part in buf.value.decode().split(";"):
            if part:
                name, ms, cnt = part.split(":")
                out[name] = (float(ms), int(cnt))
        return out

    def setOption(self, name: str, value: int):
        if N.lib().zstdb200_set_option(self._live(), name.encode(), value) != 0:
            raise KeyError(name)

    def _raise(self, r: int):
        L = N.lib()
        msg = L.ZSTD_getErrorName(r).decode()
        extra = L.zstdb200_last_error().decode()
        raise ZstdException(N.error_code(r), msg + (f" [{extra}]" if extra and N.error_code(r) == 1 else ""))

    def compressChunks(self, src, chunkSize: int = BLOCK, level: int = 3):
        """One contiguous buffer -> (stream bytes (numpy uint8), frame sizes (numpy uint64))."""
        L = N.lib()
        sp, n, _k = _buf(src)
        nch = max(1, -(-n // chunkSize))
        cap = sum(max(18, L.ZSTD_compressBound(min(chunkSize, n - i * chunkSize) if n else 0)) for i in range(nch)) if nch < 64 else nch * max(18, L.ZSTD_compressBound(chunkSize))
        out = np.empty(cap, dtype=np.uint8)
        sizes = (C.c_size_t * nch)()
        total = C.c_size_t(0)
        r = L.zstdb200_compress_chunks(self._live(), level, sp, n, chunkSize, C.c_void_p(out.ctypes.data), cap, sizes, C.byref(total))
        if N.is_error(r):
            self._raise(r)
        return out[: total.value], np.ctypeslib.as_array(sizes).astype(np.uint64)

    def decompressFrames(self, stream, frameSizes: Sequence[int], originalSizes: Sequence[int]):
        """Packed frames -> (bytes (numpy uint8, items back to back), regenerated sizes)."""
        L = N.lib()
        sp, n, _k = _buf(stream)
        k = len(frameSizes)
        fs = (C.c_size_t * k)(*[int(x) for x in frameSizes])
        ds = (C.c_size_t * k)(*[int(x) for x in originalSizes])
        total = int(sum(int(x) for x in originalSizes))
        out = np.empty(max(total, 1), dtype=np.uint8)
        r = L.zstdb200_decompress_frames(self._live(), sp, fs, k, C.c_void_p(out.ctypes.data), total, ds)
        sizes = np.ctypeslib.as_array(ds).astype(np.uint64)
        if N.is_error(r):
            self._raise(r)
        return out[:total], sizes

    # ---- asynchronous forms (include/zstdb200.h): begin queues copy-in + kernels on work set `slot`, end collects.  The buffers are
    # numpy uint8 arrays owned by the caller (page-locked ones make the copies asynchronous) and must stay untouched in between.
    SLOTS = 4

    def compressChunksBegin(self, slot: int, src: np.ndarray, chunkSize: int = BLOCK, level: int = 3):
        r = N.lib().zstdb200_compress_chunks_begin(self._live(), slot, level, C.c_void_p(src.ctypes.data), src.size, chunkSize)
        if N.is_error(r):
            self._raise(r)

    def compressChunksEnd(self, slot: int, dst: np.ndarray, frameSizes=None) -> int:
        """Waits for slot's compression, copies the packed frames into dst; returns the stream size.  frameSizes: a ctypes size_t
        array that receives the per-frame sizes (optional)."""
        total = C.c_size_t(0)
        r = N.lib().zstdb200_compress_chunks_end(self._live(), slot, C.c_void_p(dst.ctypes.data), dst.size, frameSizes, C.byref(total))
        if N.is_error(r):
            self._raise(r)
        return total.value

    def decompressFramesBegin(self, slot: int, stream: np.ndarray, frameSizes, dst: np.ndarray, originalSizes):
        """frameSizes / originalSizes: ctypes size_t arrays of equal length (kept alive by the caller until End)."""
        r = N.lib().zstdb200_decompress_frames_begin(self._live(), slot, C.c_void_p(stream.ctypes.data), frameSizes, len(frameSizes),
                                                     C.c_void_p(dst.ctypes.data), dst.size, originalSizes)
        if N.is_error(r):
            self._raise(r)

    def decompressFramesEnd(self, slot: int, regenerated=None):
        r = N.lib().zstdb200_decompress_frames_end(self._live(), slot, regenerated)
        if N.is_error(r):
            self._raise(r)

    def compressBatch(self, chunks: Sequence, level: int = 3, raise_on_error: bool = True):
        L = N.lib()
        k = len(chunks)
        bufs = [_buf(c) for c in chunks]
        src = (C.c_void_p * k)(*[b[0] for b in bufs])
        ssz = (C.c_size_t * k)(*[b[1] for b in bufs])
        outs = [np.empty(max(18, L.ZSTD_compressBound(b[1])), dtype=np.uint8) for b in bufs]
        dst = (C.c_void_p * k)(*[o.ctypes.data for o in outs])
        dcap = (C.c_size_t * k)(*[o.size for o in outs])
        dsz = (C.c_size_t * k)()
        r = L.zstdb200_compress_batch(self._live(), level, k, src, ssz, dst, dcap, dsz)
        if N.is_error(r) and raise_on_error:
            self._raise(r)
        if raise_on_error:
            return [outs[i][: dsz[i]].tobytes() for i in range(k)]
        return [outs[i][: dsz[i]].tobytes() if not N.is_error(dsz[i]) else -N.error_code(dsz[i]) for i in range(k)]

    def decompressBatch(self, frames: Sequence, originalSizes: Sequence[int], raise_on_error: bool = True):
        L = N.lib()
        k = len(frames)
        bufs = [_buf(f) for f in frames]
        src = (C.c_void_p * k)(*[b[0] for b in bufs])
        ssz = (C.c_size_t * k)(*[b[1] for b in bufs])
        outs = [np.empty(max(int(s), 1), dtype=np.uint8) for s in originalSizes]
        dst = (C.c_void_p * k)(*[o.ctypes.data for o in outs])
        dcap = (C.c_size_t * k)(*[int(s) for s in originalSizes])
        dsz = (C.c_size_t * k)()
        r = L.zstdb200_decompress_batch(self._live(), k, src, ssz, dst, dcap, dsz)
        if N.is_error(r) and raise_on_error:
            self._raise(r)
        if raise_on_error:
            return [outs[i][: dsz[i]].tobytes() for i in range(k)]
        return [outs[i][: dsz[i]].tobytes() if not N.is_error(dsz[i]) else -N.error_code(dsz[i]) for i in range(k)]


    def generateSequences(self, blocks: Sequence, level: int = 3, raise_on_error: bool = True):
        """ZSTD_generateSequences for every block (<= 128 KB each): list of (n, 4) uint32 arrays with the columns of
        ZSTD_Sequence (offset, litLength, matchLength, rep); the last row of a block is its delimiter."""
        L = N.lib()
        k = len(blocks)
        bufs = [_buf(b) for b in blocks]
        src = (C.c_void_p * k)(*[b[0] for b in bufs])
        ssz = (C.c_size_t * k)(*[b[1] for b in bufs])
        outs = [np.zeros((b[1] // 3 + 2, 4), dtype=np.uint32) for b in bufs]         # ZSTD_sequenceBound
        dst = (C.c_void_p * k)(*[o.ctypes.data for o in outs])
        cap = (C.c_size_t * k)(*[o.shape[0] for o in outs])
        nb = (C.c_size_t * k)()
        r = L.zstdb200_generate_sequences(self._live(), level, k, src, ssz, dst, cap, nb)
        if N.is_error(r) and raise_on_error:
            self._raise(r)
        return [outs[i][: nb[i]] if not N.is_error(nb[i]) else -N.error_code(nb[i]) for i in range(k)]


class B200SequenceProducer:
    """A J/SequenceProducer.java implementation backed by the GPU match finder: the three methods return what
    Zstd.registerSequenceProducer / ZstdCompressCtx.registerSequenceProducer pass on to ZSTD_registerSequenceProducer
    (N/jni_zstd.c, N/jni_fast_zstd.c)."""

    def __init__(self, device: int = -1):
        self.device = device

    def getFunctionPointer(self) -> int:
        return C.cast(N.lib().zstdb200_sequenceProducer, C.c_void_p).value

    def createState(self) -> int:
        L = N.lib()
        p = L.zstdb200_createSequenceProducerState(self.device)
        if not p:
            raise RuntimeError("zstdb200_createSequenceProducerState failed (no CPU fallback): " + L.zstdb200_last_error().decode())
        return p

    def freeState(self, statePointer: int) -> None:
        N.lib().zstdb200_freeSequenceProducerState(statePointer)


for _name, _code in Zstd._ERR.items():
    setattr(Zstd, _name, staticmethod(lambda _c=_code: _c))


class ZstdOutputStream:
    """J/ZstdOutputStreamNoFinalizer.java:83-90,400-518 over ZSTD_compressStream2.

    GPU build semantics: every <=128 KB block is emitted as an independent frame
    (see INTEGRATION.md), so the bytes are a legal zstd stream but not the
    reference's single-frame stream."""

    def __init__(self, out, level: int = 3):
        self._out = out
        L = N.lib()
        self._z = L.ZSTD_createCStream()
        L.ZSTD_initCStream(self._z, level)
        self._dst = bytearray(L.ZSTD_CStreamOutSize())
        self._closed = False

    def _step(self, data: bytes, end_op: int):
        L = N.lib()
        src = np.frombuffer(data, dtype=np.uint8) if data else np.empty(0, dtype=np.uint8)
        ib = N.InBuffer(src.ctypes.data if src.size else 0, src.size, 0)
        d = (C.c_char * len(self._dst)).from_buffer(self._dst)
        while True:
            ob = N.OutBuffer(C.addressof(d), len(self._dst), 0)
            r = L.ZSTD_compressStream2(self._z, C.byref(ob), C.byref(ib), end_op)
            if N.is_error(r):
                raise ZstdException(N.error_code(r), L.ZSTD_getErrorName(r).decode())
            if ob.pos:
                self._out.write(bytes(self._dst[: ob.pos]))
            if end_op == ZSTD_e_continue:
                if ib.pos == ib.size and ob.pos < ob.size:
                    break
            elif r == 0:
                break

    def write(self, data):
        if self._closed:
            raise IOError("StreamClosed")
        self._step(bytes(data), ZSTD_e_continue)

    def flush(self):
        if self._closed:
            raise IOError("StreamClosed")
        self._step(b"", ZSTD_e_flush)
        if hasattr(self._out, "flush"):
            self._out.flush()

    def close(self):
        if self._closed:
            return
        self._step(b"", ZSTD_e_end)
        N.lib().ZSTD_freeCStream(self._z)
        self._closed = True

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class ZstdInputStream:
    """J/ZstdInputStreamNoFinalizer.java:148-226 over ZSTD_decompressStream."""

    def __init__(self, inp):
        self._in = inp
        L = N.lib()
        self._z = L.ZSTD_createDStream()
        L.ZSTD_initDStream(self._z)
        self._src = b""
        self._pos = 0
        self._eof = False
        self._closed = False

    def setLongMax(self, windowLogMax: int):              # J/ZstdInputStreamNoFinalizer.java:126-135 -> N/jni_zstd.c:403 (ZSTD_d_windowLogMax)
        if self._closed:
            raise IOError("Stream closed")
        r = N.lib().ZSTD_DCtx_setParameter(self._z, ZSTD_d_windowLogMax, windowLogMax)
        if N.is_error(r):
            raise ZstdException(N.error_code(r), N.lib().ZSTD_getErrorName(r).decode())
        return self

    def read(self, n: int = -1) -> bytes:
        if self._closed:
            raise IOError("Stream closed")
        L = N.lib()
        chunks = []
        want = n if n >= 0 else 1 << 62
        dst = bytearray(L.ZSTD_DStreamOutSize())
        d = (C.c_char * len(dst)).from_buffer(dst)
        while want > 0:
            if self._pos == len(self._src) and not self._eof:
                self._src = self._in.read(L.ZSTD_DStreamInSize())
                self._pos = 0
                if not self._src:
                    self._eof = True
            src = np.frombuffer(self._src, dtype=np.uint8) if self._src else np.empty(0, dtype=np.uint8)
            ib = N.InBuffer(src.ctypes.data if src.size else 0, src.size, self._pos)
            ob = N.OutBuffer(C.addressof(d), min(len(dst), want), 0)
            r = L.ZSTD_decompressStream(self._z, C.byref(ob), C.byref(ib))
            if N.is_error(r):
                raise ZstdException(N.error_code(r), L.ZSTD_getErrorName(r).decode())
            self._pos = ib.pos
            if ob.pos:
                chunks.append(bytes(dst[: ob.pos]))
                want -= ob.pos
            elif self._eof and self._pos == len(self._src):
                if r != 0:
                    raise IOError("Truncated source")    # J/ZstdInputStreamNoFinalizer.java:187-197
                break
        return b"".join(chunks)

    def close(self):
        if not self._closed:
            N.lib().ZSTD_freeDStream(self._z)
            self._closed = True

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class ByteBuffer:
    """The few pieces of java.nio.ByteBuffer the direct-buffer stream classes use (position / limit / remaining / flip / clear),
    over a numpy uint8 array -- page-locked when it comes from ``allocateDirect`` (torch pinned memory when torch is importable),
    which is what lets the copy engines take the bytes without a staging pass, like a registered DirectByteBuffer."""

    def __init__(self, array: np.ndarray):
        assert array.dtype == np.uint8 and array.ndim == 1 and array.flags["C_CONTIGUOUS"]
        self.array = array
        self._position = 0
        self._limit = array.size
        self._keep = None

    @staticmethod
    def allocateDirect(capacity: int) -> "ByteBuffer":
        try:
            import torch
            t = torch.empty(capacity, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
            b = ByteBuffer(t.numpy()); b._keep = t
            return b
        except ImportError:
            return ByteBuffer(np.empty(capacity, dtype=np.uint8))

    @staticmethod
    def wrap(data) -> "ByteBuffer":
        return ByteBuffer(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)

    def isDirect(self) -> bool: return True
    def capacity(self) -> int: return self.array.size
    def position(self, p: int | None = None):
        if p is None: return self._position
        assert 0 <= p <= self._limit; self._position = p; return self
    def limit(self, l: int | None = None):
        if l is None: return self._limit
        assert 0 <= l <= self.array.size; self._limit = l; self._position = min(self._position, l); return self
    def remaining(self) -> int: return self._limit - self._position
    def hasRemaining(self) -> bool: return self._position < self._limit
    def flip(self): self._limit = self._position; self._position = 0; return self
    def clear(self): self._position = 0; self._limit = self.array.size; return self
    def address(self) -> int: return self.array.ctypes.data


class ZstdDirectBufferCompressingStream:
    """J/ZstdDirectBufferCompressingStreamNoFinalizer.java:13-200 over ZSTD_compressStream / flushStream / endStream
    (N/jni_directbuffercompress_zstd.c): the whole remaining source goes to every native call, so the batch path behind
    ZSTD_compressStream2 sees hundreds of MiB at a time.  ``flushBuffer`` is the hook a subclass overrides to drain the target."""

    def __init__(self, target: ByteBuffer, level: int = 3):
        L = N.lib()
        self._target = target
        self._z = L.ZSTD_createCStream()
        self._level = level
        self._initialized = False
        self._closed = False

    @staticmethod
    def recommendedOutputBufferSize() -> int:
        return int(N.lib().ZSTD_CStreamOutSize())

    def flushBuffer(self, toFlush: ByteBuffer) -> ByteBuffer:
        return toFlush

    def _call(self, src: ByteBuffer | None, end_op: int) -> int:
        L = N.lib()
        t = self._target
        ib = N.InBuffer(src.address() + src.position() if src is not None else 0, src.remaining() if src is not None else 0, 0)
        ob = N.OutBuffer(t.address() + t.position(), t.remaining(), 0)
        r = L.ZSTD_compressStream2(self._z, C.byref(ob), C.byref(ib), end_op)
        if N.is_error(r):
            raise ZstdException(N.error_code(r), L.ZSTD_getErrorName(r).decode())
        t.position(t.position() + ob.pos)
        if src is not None:
            src.position(src.position() + ib.pos)
        return r

    def compress(self, source: ByteBuffer):
        if self._closed:
            raise IOError("Stream closed")
        if not self._initialized:
            N.lib().ZSTD_initCStream(self._z, self._level); self._initialized = True
        while source.hasRemaining():
            if not self._target.hasRemaining():
                self._target = self.flushBuffer(self._target)
                if not self._target.hasRemaining():
                    raise IOError("The target buffer has no more space, even after flushing, and there are still bytes to compress")
            self._call(source, ZSTD_e_continue)

    def _finish(self, end_op: int):
        if not self._initialized:
            return
        first = True
        while True:
            needed = self._call(None, end_op)
            self._target = self.flushBuffer(self._target)
            if needed > 0 and not self._target.hasRemaining() and not first:
                raise IOError("The target buffer has no more space, even after flushing, and there are still bytes to compress")
            first = False
            if needed == 0:
                break

    def flush(self):
        if self._closed:
            raise IOError("Already closed")
        self._finish(ZSTD_e_flush)

    def close(self):
        if self._closed:
            return
        try:
            self._finish(ZSTD_e_end)
        finally:
            N.lib().ZSTD_freeCStream(self._z)
            self._closed = True
            self._target = None

    def __enter__(self): return self
    def __exit__(self, *a): self.close()


class ZstdDirectBufferDecompressingStream:
    """J/ZstdDirectBufferDecompressingStreamNoFinalizer.java + J/BaseZstdBufferDecompressingStreamNoFinalizer.java:76-112 over
    ZSTD_decompressStream (N/jni_directbufferdecompress_zstd.c).  ``refill`` is the subclass hook that supplies more source."""

    def __init__(self, source: ByteBuffer):
        L = N.lib()
        self._source = source
        self._z = L.ZSTD_createDStream()
        L.ZSTD_initDStream(self._z)
        self._closed = False
        self._streamEnd = False
        self._finishedFrame = False

    @staticmethod
    def recommendedTargetBufferSize() -> int:
        return int(N.lib().ZSTD_DStreamOutSize())

    def refill(self, toRefill: ByteBuffer) -> ByteBuffer:
        return toRefill

    def hasRemaining(self) -> bool:
        return self._source is not None and not self._closed and not self._streamEnd and (self._source.hasRemaining() or not self._finishedFrame)

    def setLongMax(self, windowLogMax: int):
        if self._closed:
            raise IOError("Stream closed")
        r = N.lib().ZSTD_DCtx_setParameter(self._z, ZSTD_d_windowLogMax, windowLogMax)
        if N.is_error(r):
            raise ZstdException(N.error_code(r), N.lib().ZSTD_getErrorName(r).decode())
        return self

    def read(self, target: ByteBuffer) -> int:
        if self._closed:
            raise IOError("Stream closed")
        if self._streamEnd:
            return 0
        L = N.lib()
        s = self._source
        ib = N.InBuffer(s.address() + s.position(), s.remaining(), 0)
        ob = N.OutBuffer(target.address() + target.position(), target.remaining(), 0)
        remaining = L.ZSTD_decompressStream(self._z, C.byref(ob), C.byref(ib))
        if N.is_error(remaining):
            raise ZstdException(N.error_code(remaining), L.ZSTD_getErrorName(remaining).decode())
        s.position(s.position() + ib.pos)
        target.position(target.position() + ob.pos)
        if not s.hasRemaining():
            self._source = s = self.refill(s)
        self._finishedFrame = remaining == 0
        if self._finishedFrame:
            self._streamEnd = not s.hasRemaining()
        return ob.pos

    def close(self):
        if not self._closed:
            N.lib().ZSTD_freeDStream(self._z)
            self._closed = True
            self._source = None

    def __enter__(self): return self
    def __exit__(self, *a): self.close()
