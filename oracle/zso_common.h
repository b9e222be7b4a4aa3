/*
 * zso_common.h -- shared definitions for the CPU oracle ("zso" = zstd oracle).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped
 * product path: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build, link or call it.
 *
 * This is a from-scratch plain-C restatement of the Zstandard block format
 * (RFC 8878) and of the *encoder decisions* of libzstd 1.5.7 as vendored by
 * luben/zstd-jni under src/main/native/ (abbreviated N/ below).  Every
 * function cites the reference file:line whose behaviour it follows.
 * Parity is pinned two ways (see oracle/README.md):
 *   1. against the reference's own golden vectors (the .zst files under src/test/resources,
 *      copied digests only -- the oracle decoder must regenerate `xml`);
 *   2. against oracle/_ref/libzstd-oracle.so, the reference's own C sources
 *      compiled in place by oracle/Makefile (byte-identical frames).
 */
#ifndef ZSO_COMMON_H
#define ZSO_COMMON_H

#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* Error convention of N/common/error_private.h:49-54 : (size_t)-code.
 * Codes from N/zstd_errors.h:42-78. */
enum {
    ZSO_error_GENERIC = 1,
    ZSO_error_prefix_unknown = 10,
    ZSO_error_frameParameter_unsupported = 14,
    ZSO_error_frameParameter_windowTooLarge = 16,
    ZSO_error_corruption_detected = 20,
    ZSO_error_checksum_wrong = 22,
    ZSO_error_literals_headerWrong = 24,
    ZSO_error_dictionary_corrupted = 30,
    ZSO_error_dictionary_wrong = 32,
    ZSO_error_parameter_unsupported = 40,
    ZSO_error_tableLog_tooLarge = 44,
    ZSO_error_maxSymbolValue_tooLarge = 46,
    ZSO_error_maxSymbolValue_tooSmall = 48,
    ZSO_error_dstSize_tooSmall = 70,
    ZSO_error_srcSize_wrong = 72,
    ZSO_error_maxCode = 120
};
#define ZSO_ERROR(name) ((size_t)-(ptrdiff_t)ZSO_error_##name)
static inline int zso_isError(size_t c) { return c > (size_t)-(ptrdiff_t)ZSO_error_maxCode; }

/* format constants, N/common/zstd_internal.h:90-113 */
#define ZSO_BLOCKSIZE_MAX (1u << 17)
#define ZSO_MINMATCH 3
#define ZSO_MaxLL 35
#define ZSO_MaxML 52
#define ZSO_MaxOff 31
#define ZSO_DefaultMaxOff 28
#define ZSO_LLFSELog 9
#define ZSO_MLFSELog 9
#define ZSO_OffFSELog 8
#define ZSO_LitHufLog 11
#define ZSO_LONGNBSEQ 0x7F00
#define ZSO_HUF_TABLELOG_MAX 12

static inline unsigned zso_highbit32(uint32_t v) { return 31u - (unsigned)__builtin_clz(v); }
static inline uint16_t zso_rd16(const void* p) { uint16_t v; memcpy(&v, p, 2); return v; }
static inline uint32_t zso_rd24(const void* p) { const uint8_t* b = (const uint8_t*)p; return b[0] | (b[1] << 8) | ((uint32_t)b[2] << 16); }
static inline uint32_t zso_rd32(const void* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t zso_rd64(const void* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline void zso_wr16(void* p, uint16_t v) { memcpy(p, &v, 2); }
static inline void zso_wr24(void* p, uint32_t v) { uint8_t* b = (uint8_t*)p; b[0] = (uint8_t)v; b[1] = (uint8_t)(v >> 8); b[2] = (uint8_t)(v >> 16); }
static inline void zso_wr32(void* p, uint32_t v) { memcpy(p, &v, 4); }

/* extra-bit counts per code: N/common/zstd_internal.h:119-125 (LL), :136-144 (ML) */
extern const uint8_t zso_LL_bits[ZSO_MaxLL + 1];
extern const uint8_t zso_ML_bits[ZSO_MaxML + 1];
/* default distributions: N/common/zstd_internal.h:126-164 */
extern const int16_t zso_LL_defaultNorm[ZSO_MaxLL + 1];
extern const int16_t zso_ML_defaultNorm[ZSO_MaxML + 1];
extern const int16_t zso_OF_defaultNorm[ZSO_DefaultMaxOff + 1];
/* base values (derived from the bit counts; N/decompress/zstd_decompress_block.c uses LL_base/ML_base/OF_base) */
uint32_t zso_LL_base(unsigned code);
uint32_t zso_ML_base(unsigned code);
uint32_t zso_OF_base(unsigned code);

/* ---- public oracle API (oracle/zso.h mirrors this for ctypes users) ---- */
size_t zso_compressBound(size_t srcSize);
size_t zso_compress(void* dst, size_t dstCapacity, const void* src, size_t srcSize, int level);
size_t zso_compress_flags(void* dst, size_t dstCapacity, const void* src, size_t srcSize, int level, unsigned flags);   /* 1 = checksum, 2 = no content size */
size_t zso_decompress(void* dst, size_t dstCapacity, const void* src, size_t srcSize);
size_t zso_findFrameCompressedSize(const void* src, size_t srcSize);
unsigned long long zso_getFrameContentSize(const void* src, size_t srcSize);

#endif
