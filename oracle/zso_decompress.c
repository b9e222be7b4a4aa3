/*
 * zso_decompress.c -- CPU oracle: Zstandard frame / block decoder.
 * TEST INFRASTRUCTURE ONLY (see zso_common.h).
 *
 * Restates, scalar and unoptimised, what the reference does in
 *   N/decompress/zstd_decompress.c        (frame layer)
 *   N/decompress/zstd_decompress_block.c  (literals, sequences, execution)
 *   N/decompress/huf_decompress.c         (Huffman X1 tables + streams)
 *   N/common/entropy_common.c, N/common/fse_decompress.c (NCount, weights)
 * with N/ = /root/reference/src/main/native/.
 *
 * Scope: any number of concatenated frames, any number of blocks per frame,
 * no dictionary, optional skippable frames.  XXH64 content checksums are
 * verified (see zso_xxh64 below).
 */
#include "zso_common.h"

/* ---------------------------------------------------------------- tables */
const uint8_t zso_LL_bits[ZSO_MaxLL + 1] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12,
    13, 14, 15, 16 };
const uint8_t zso_ML_bits[ZSO_MaxML + 1] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11,
    12, 13, 14, 15, 16 };
const int16_t zso_LL_defaultNorm[ZSO_MaxLL + 1] = {
    4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1,
    2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1,
    -1, -1, -1, -1 };
const int16_t zso_ML_defaultNorm[ZSO_MaxML + 1] = {
    1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1,
    1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
    1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1,
    -1, -1, -1, -1, -1 };
const int16_t zso_OF_defaultNorm[ZSO_DefaultMaxOff + 1] = {
    1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1,
    1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1 };

/* base value of a length code = sum of the spans of all smaller codes */
uint32_t zso_LL_base(unsigned code) {
    uint32_t b = 0; unsigned c;
    for (c = 0; c < code; c++) b += 1u << zso_LL_bits[c];
    return b;
}
uint32_t zso_ML_base(unsigned code) {
    uint32_t b = ZSO_MINMATCH; unsigned c;
    for (c = 0; c < code; c++) b += 1u << zso_ML_bits[c];
    return b;
}
/* N/decompress/zstd_decompress_block.c OF_base: {0,1,1,5,0xD,...} = (1<<n)-3 for n>=2 */
uint32_t zso_OF_base(unsigned code) {
    if (code == 0) return 0;
    if (code == 1) return 1;
    return (1u << code) - 3;
}

/* ------------------------------------------- backward bit reader (abstract)
 * Semantics of BIT_initDStream / BIT_readBits / BIT_endOfDStream,
 * N/common/bitstream.h:254-299,362-367,449-452, restated position-wise:
 * the stream is a little-endian bit array; its last byte carries a 1-bit end
 * mark; bits are consumed from just below the mark downward.  `left` = number
 * of unread bits.  Reading past bit 0 sets `over` (the reference's
 * BIT_DStream_overflow) and yields zeros for the missing low bits. */
typedef struct { const uint8_t* p; int64_t left; int over; } zso_br;

static size_t br_init(zso_br* b, const uint8_t* p, size_t n) {
    b->p = p; b->left = 0; b->over = 0;
    if (n < 1) return ZSO_ERROR(srcSize_wrong);
    if (p[n - 1] == 0) return ZSO_ERROR(GENERIC);   /* end mark missing; callers map to corruption_detected */
    b->left = (int64_t)(n - 1) * 8 + zso_highbit32(p[n - 1]);
    return n;
}
static uint32_t br_peek_at(const zso_br* b, int64_t lo, unsigned n) {   /* bits [lo, lo+n), n<=32 */
    uint64_t v = 0; unsigned got = 0;
    if (n == 0) return 0;
    if (lo < 0) { unsigned miss = (unsigned)(-lo); if (miss >= n) return 0; return br_peek_at(b, 0, n - miss) << miss; }
    {   int64_t byte = lo >> 3; unsigned sh = (unsigned)(lo & 7);
        /* bytes beyond the stream end are never requested: lo+n <= left <= 8*size */
        while (got < n + sh) { v |= (uint64_t)b->p[byte++] << got; got += 8; if (got >= 64) break; }
        v >>= sh;
    }
    return (uint32_t)(v & ((n == 32) ? 0xFFFFFFFFu : ((1u << n) - 1)));
}
static uint32_t br_read(zso_br* b, unsigned n) {
    uint32_t v;
    b->left -= n;
    if (b->left < 0) b->over = 1;
    v = br_peek_at(b, b->left, n);
    return v;
}
static int br_finished(const zso_br* b) { return b->left == 0 && !b->over; }

/* ---------------------------------------------------------- FSE NCount parse
 * FSE_readNCount_body, N/common/entropy_common.c:42-187, restated on a bit
 * cursor (forward, LSB first).  *maxSV in: alphabet limit; out: last symbol. */
static uint32_t fwd_bits(const uint8_t* p, size_t size, size_t bitpos, unsigned n) {
    uint64_t v = 0; unsigned got = 0; size_t byte = bitpos >> 3; unsigned sh = (unsigned)(bitpos & 7);
    while (got < n + sh && got < 64) { v |= (uint64_t)(byte < size ? p[byte] : 0) << got; byte++; got += 8; }
    return (uint32_t)((v >> sh) & ((n >= 32) ? 0xFFFFFFFFu : ((1u << n) - 1)));
}

size_t zso_readNCount(int16_t* norm, unsigned* maxSV, unsigned* tableLog, const uint8_t* src, size_t srcSize) {
    size_t bitpos = 0;
    unsigned const maxSV1 = *maxSV + 1;
    unsigned charnum = 0;
    int nbBits, remaining, threshold, previous0 = 0;
    /* the reference zero-pads inputs shorter than 8 bytes (entropy_common.c:57-66)
     * and finally rejects if it consumed more than srcSize; fwd_bits() pads too. */
    if (srcSize == 0) return ZSO_ERROR(srcSize_wrong);
    memset(norm, 0, maxSV1 * sizeof(norm[0]));
    nbBits = (int)fwd_bits(src, srcSize, bitpos, 4) + 5; bitpos += 4;
    if (nbBits > 15) return ZSO_ERROR(tableLog_tooLarge);     /* FSE_TABLELOG_ABSOLUTE_MAX */
    *tableLog = (unsigned)nbBits;
    remaining = (1 << nbBits) + 1;
    threshold = 1 << nbBits;
    nbBits++;
    for (;;) {
        if (previous0) {
            /* 2-bit repeat codes; 0b11 means "3 more zeros and continue" (:82-104) */
            for (;;) {
                unsigned r = fwd_bits(src, srcSize, bitpos, 2); bitpos += 2;
                charnum += r;
                if (r != 3) break;
                if (bitpos > 8 * srcSize + 64) break;     /* runaway on padded zeros cannot happen (0b00 stops) */
            }
            if (charnum >= maxSV1) break;                 /* error reported after the loop (:114) */
        }
        {   int const max = (2 * threshold - 1) - remaining;
            int count;
            uint32_t const bs = fwd_bits(src, srcSize, bitpos, (unsigned)nbBits);
            if ((int)(bs & (uint32_t)(threshold - 1)) < max) {
                count = (int)(bs & (uint32_t)(threshold - 1));
                bitpos += (unsigned)nbBits - 1;
            } else {
                count = (int)(bs & (uint32_t)(2 * threshold - 1));
                if (count >= threshold) count -= max;
                bitpos += (unsigned)nbBits;
            }
            count--;
            if (count >= 0) remaining -= count; else remaining += count;
            norm[charnum++] = (int16_t)count;
            previous0 = !count;
            if (remaining < threshold) {
                if (remaining <= 1) break;
                nbBits = (int)zso_highbit32((uint32_t)remaining) + 1;
                threshold = 1 << (nbBits - 1);
            }
            if (charnum >= maxSV1) break;
        }
    }
    if (remaining != 1) return ZSO_ERROR(corruption_detected);
    if (charnum > maxSV1) return ZSO_ERROR(maxSymbolValue_tooSmall);
    *maxSV = charnum - 1;
    {   size_t const used = (bitpos + 7) >> 3;
        if (used > srcSize) return ZSO_ERROR(corruption_detected);   /* :64 and :182 */
        return used;
    }
}

/* ------------------------------------------------------- FSE decode tables
 * symbol spreading shared by FSE_buildDTable_internal (N/common/fse_decompress.c:58-159)
 * and ZSTD_buildFSETable_body (N/decompress/zstd_decompress_block.c:485-603).
 * Both branches of the reference (the "no low-prob" fast path and the generic
 * path) visit cells in the same order: position += step, skipping the
 * low-probability area at the top. */
static int fse_spread(uint8_t* tableSymbol, uint16_t* symbolNext, const int16_t* norm, unsigned maxSV, unsigned tableLog) {
    uint32_t const tableSize = 1u << tableLog, mask = tableSize - 1;
    uint32_t const step = (tableSize >> 1) + (tableSize >> 3) + 3;
    uint32_t high = tableSize - 1, pos = 0, s;
    for (s = 0; s <= maxSV; s++) {
        if (norm[s] == -1) { tableSymbol[high--] = (uint8_t)s; symbolNext[s] = 1; }
        else symbolNext[s] = (uint16_t)norm[s];
    }
    for (s = 0; s <= maxSV; s++) {
        int i;
        for (i = 0; i < norm[s]; i++) {
            tableSymbol[pos] = (uint8_t)s;
            pos = (pos + step) & mask;
            while (pos > high) pos = (pos + step) & mask;
        }
    }
    return pos == 0;
}

typedef struct { uint16_t nextState; uint8_t nbAddBits; uint8_t nbBits; uint32_t baseValue; } zso_seqSym;
typedef struct { zso_seqSym t[512]; unsigned log; } zso_seqTable;

/* kind: 0 = LL, 1 = OF, 2 = ML */
static uint32_t code_base(int kind, unsigned c) { return kind == 0 ? zso_LL_base(c) : kind == 1 ? zso_OF_base(c) : zso_ML_base(c); }
static uint8_t code_bits(int kind, unsigned c) { return kind == 0 ? zso_LL_bits[c] : kind == 1 ? (uint8_t)c : zso_ML_bits[c]; }

static void build_seq_table(zso_seqTable* dt, const int16_t* norm, unsigned maxSV, unsigned tableLog, int kind) {
    uint8_t sym[512]; uint16_t next[64]; uint32_t u, tableSize = 1u << tableLog;
    fse_spread(sym, next, norm, maxSV, tableLog);
    dt->log = tableLog;
    for (u = 0; u < tableSize; u++) {
        unsigned const s = sym[u];
        uint32_t const ns = next[s]++;
        uint8_t const nb = (uint8_t)(tableLog - zso_highbit32(ns));
        dt->t[u].nbBits = nb;
        dt->t[u].nextState = (uint16_t)((ns << nb) - tableSize);
        dt->t[u].nbAddBits = code_bits(kind, s);
        dt->t[u].baseValue = code_base(kind, s);
    }
}
static void build_seq_table_rle(zso_seqTable* dt, unsigned symbol, int kind) {  /* ZSTD_buildSeqTable_rle :463-477 */
    dt->log = 0;
    dt->t[0].nbBits = 0; dt->t[0].nextState = 0;
    dt->t[0].nbAddBits = code_bits(kind, symbol);
    dt->t[0].baseValue = code_base(kind, symbol);
}

/* --------------------------------------------------- Huffman weights / table
 * HUF_readStats_body, N/common/entropy_common.c:243-306 (with the 2-state FSE
 * decoder of N/common/fse_decompress.c:173-236 for compressed weights). */
typedef struct { uint8_t sym; uint8_t nbBits; } zso_hufD;
typedef struct { zso_hufD t[1 << ZSO_HUF_TABLELOG_MAX]; unsigned log; } zso_hufTable;

static size_t fse_decode_weights(uint8_t* dst, size_t dstCap, const uint8_t* src, size_t srcSize) {
    int16_t norm[256]; unsigned maxSV = 255, tableLog;
    uint8_t sym[64]; uint16_t next[256]; uint8_t dnb[64]; uint16_t dnew[64];
    size_t h = zso_readNCount(norm, &maxSV, &tableLog, src, srcSize);
    zso_br br; uint32_t s1, s2; size_t n = 0; uint32_t u;
    if (zso_isError(h)) return h;
    if (tableLog > 6) return ZSO_ERROR(tableLog_tooLarge);          /* fse_decompress.c:267 */
    if (!fse_spread(sym, next, norm, maxSV, tableLog)) return ZSO_ERROR(GENERIC);
    for (u = 0; u < (1u << tableLog); u++) {
        uint32_t const ns = next[sym[u]]++;
        dnb[u] = (uint8_t)(tableLog - zso_highbit32(ns));
        dnew[u] = (uint16_t)((ns << dnb[u]) - (1u << tableLog));
    }
    {   size_t e = br_init(&br, src + h, srcSize - h); if (zso_isError(e)) return e; }
    s1 = br_read(&br, tableLog); s2 = br_read(&br, tableLog);
    if (br.over) return ZSO_ERROR(corruption_detected);              /* :193 */
    /* alternate the two states until the stream overflows; then the other
     * state's pending symbol is the last one (:219-232) */
    for (;;) {
        if (n + 2 > dstCap) return ZSO_ERROR(dstSize_tooSmall);      /* op > omax-2 */
        dst[n++] = sym[s1]; s1 = dnew[s1] + br_read(&br, dnb[s1]);
        if (br.over) { dst[n++] = sym[s2]; break; }
        if (n + 2 > dstCap) return ZSO_ERROR(dstSize_tooSmall);
        dst[n++] = sym[s2]; s2 = dnew[s2] + br_read(&br, dnb[s2]);
        if (br.over) { dst[n++] = sym[s1]; break; }
    }
    return n;
}

/* returns bytes consumed; fills table.  HUF_readDTableX1_wksp N/decompress/huf_decompress.c:385-519
 * (the X1-vs-X2 choice :1794-1842 and the tableLog rescale :352-375 do not change decoded bytes). */
static size_t huf_read_table(zso_hufTable* ht, const uint8_t* src, size_t srcSize) {
    uint8_t w[256]; uint32_t rank[ZSO_HUF_TABLELOG_MAX + 2]; size_t iSize, oSize, n; uint32_t total = 0, tableLog;
    if (!srcSize) return ZSO_ERROR(srcSize_wrong);
    iSize = src[0];
    if (iSize >= 128) {
        oSize = iSize - 127; iSize = (oSize + 1) / 2;
        if (iSize + 1 > srcSize) return ZSO_ERROR(srcSize_wrong);
        if (oSize >= 256) return ZSO_ERROR(corruption_detected);
        for (n = 0; n < oSize; n += 2) { w[n] = src[1 + n / 2] >> 4; w[n + 1] = src[1 + n / 2] & 15; }
    } else {
        if (iSize + 1 > srcSize) return ZSO_ERROR(srcSize_wrong);
        oSize = fse_decode_weights(w, 255, src + 1, iSize);
        if (zso_isError(oSize)) return oSize;
    }
    memset(rank, 0, sizeof(rank));
    for (n = 0; n < oSize; n++) {
        if (w[n] > ZSO_HUF_TABLELOG_MAX) return ZSO_ERROR(corruption_detected);
        rank[w[n]]++; total += (1u << w[n]) >> 1;
    }
    if (total == 0) return ZSO_ERROR(corruption_detected);
    tableLog = zso_highbit32(total) + 1;
    if (tableLog > ZSO_HUF_TABLELOG_MAX) return ZSO_ERROR(corruption_detected);
    {   uint32_t const rest = (1u << tableLog) - total;
        uint32_t const last = zso_highbit32(rest) + 1;
        if ((1u << zso_highbit32(rest)) != rest) return ZSO_ERROR(corruption_detected);
        w[oSize] = (uint8_t)last; rank[last]++;
    }
    if (rank[1] < 2 || (rank[1] & 1)) return ZSO_ERROR(corruption_detected);
    /* fill: weights ascending, symbols ascending inside a weight (:455-517) */
    {   uint32_t start[ZSO_HUF_TABLELOG_MAX + 2]; uint32_t wgt, pos = 0, s;
        for (wgt = 1; wgt <= tableLog; wgt++) { start[wgt] = pos; pos += rank[wgt] << (wgt - 1); }
        for (s = 0; s <= oSize; s++) {
            uint32_t const ww = w[s], len = ww ? (1u << (ww - 1)) : 0; uint32_t u;
            for (u = 0; u < len; u++) { ht->t[start[ww] + u].sym = (uint8_t)s; ht->t[start[ww] + u].nbBits = (uint8_t)(tableLog + 1 - ww); }
            if (ww) start[ww] += len;
        }
    }
    ht->log = tableLog;
    return iSize + 1;
}

/* one stream: HUF_decompress1X1_usingDTable_internal_body :574-595 */
static size_t huf_decode_stream(uint8_t* dst, size_t n, const uint8_t* src, size_t srcSize, const zso_hufTable* ht) {
    zso_br br; size_t i; size_t e = br_init(&br, src, srcSize);
    if (zso_isError(e)) return ZSO_ERROR(corruption_detected);
    for (i = 0; i < n; i++) {
        /* look at the next `log` bits (zero-filled below bit 0), consume nbBits */
        uint32_t const idx = br_peek_at(&br, br.left - (int64_t)ht->log, ht->log);
        dst[i] = ht->t[idx].sym;
        br.left -= ht->t[idx].nbBits;
        if (br.left < 0) br.over = 1;
    }
    if (!br_finished(&br)) return ZSO_ERROR(corruption_detected);
    return n;
}
/* four streams: HUF_decompress4X1_usingDTable_internal_body :601-698 */
static size_t huf_decode_4x(uint8_t* dst, size_t n, const uint8_t* src, size_t srcSize, const zso_hufTable* ht) {
    size_t l1, l2, l3, l4, seg, k; const uint8_t* s[4]; size_t len[4], out[4];
    if (srcSize < 10) return ZSO_ERROR(corruption_detected);
    if (n < 6) return ZSO_ERROR(corruption_detected);
    l1 = zso_rd16(src); l2 = zso_rd16(src + 2); l3 = zso_rd16(src + 4);
    if (l1 + l2 + l3 + 6 > srcSize) return ZSO_ERROR(corruption_detected);
    l4 = srcSize - (l1 + l2 + l3 + 6);
    seg = (n + 3) / 4;
    if (3 * seg > n) return ZSO_ERROR(corruption_detected);
    s[0] = src + 6; s[1] = s[0] + l1; s[2] = s[1] + l2; s[3] = s[2] + l3;
    len[0] = l1; len[1] = l2; len[2] = l3; len[3] = l4;
    out[0] = out[1] = out[2] = seg; out[3] = n - 3 * seg;
    for (k = 0; k < 4; k++) {
        size_t r = huf_decode_stream(dst + k * seg, out[k], s[k], len[k], ht);
        if (zso_isError(r)) return r;
    }
    return n;
}

/* ----------------------------------------------------------- block decoder */
typedef struct {
    zso_seqTable LL, OF, ML;            /* persist for set_repeat */
    zso_hufTable huf;
    uint32_t rep[3];
    int litEntropy, fseEntropy;
    uint8_t lit[ZSO_BLOCKSIZE_MAX + 8];
    const uint8_t* litPtr; size_t litSize;
} zso_dctx;

/* ZSTD_decodeLiteralsBlock, N/decompress/zstd_decompress_block.c:134-340 */
static size_t decode_literals(zso_dctx* d, const uint8_t* src, size_t srcSize, size_t blockSizeMax, size_t dstCapacity) {
    unsigned type, lhl; size_t const expectedWrite = blockSizeMax < dstCapacity ? blockSizeMax : dstCapacity;
    if (srcSize < 2) return ZSO_ERROR(corruption_detected);   /* MIN_CBLOCK_SIZE */
    type = src[0] & 3; lhl = (src[0] >> 2) & 3;
    if (type == 3 && !d->litEntropy) return ZSO_ERROR(dictionary_corrupted);
    if (type >= 2) {
        size_t lhSize, litSize, litCSize; int single = 0; uint32_t lhc; size_t r;
        if (srcSize < 5) return ZSO_ERROR(corruption_detected);
        lhc = zso_rd32(src);
        switch (lhl) {
        case 0: case 1: default: single = !lhl; lhSize = 3; litSize = (lhc >> 4) & 0x3FF; litCSize = (lhc >> 14) & 0x3FF; break;
        case 2: lhSize = 4; litSize = (lhc >> 4) & 0x3FFF; litCSize = lhc >> 18; break;
        case 3: lhSize = 5; litSize = (lhc >> 4) & 0x3FFFF; litCSize = (lhc >> 22) + ((size_t)src[4] << 10); break;
        }
        if (litSize > blockSizeMax) return ZSO_ERROR(corruption_detected);
        if (!single && litSize < 6) return ZSO_ERROR(literals_headerWrong);
        if (litCSize + lhSize > srcSize) return ZSO_ERROR(corruption_detected);
        if (expectedWrite < litSize) return ZSO_ERROR(dstSize_tooSmall);
        {   const uint8_t* p = src + lhSize; size_t c = litCSize;
            if (type == 2) {
                size_t h;
                /* HUF_decompress4X_hufOnly_wksp :1924-1928 / HUF_decompress1X1_DCtx_wksp :1893-1903 */
                if (!single) { if (litSize == 0) return ZSO_ERROR(corruption_detected); if (c == 0) return ZSO_ERROR(corruption_detected); }
                h = huf_read_table(&d->huf, p, c);
                if (zso_isError(h)) return ZSO_ERROR(corruption_detected);
                if (h >= c) return ZSO_ERROR(corruption_detected);
                p += h; c -= h;
            }
            r = single ? huf_decode_stream(d->lit, litSize, p, c, &d->huf) : huf_decode_4x(d->lit, litSize, p, c, &d->huf);
            if (zso_isError(r)) return ZSO_ERROR(corruption_detected);
        }
        d->litPtr = d->lit; d->litSize = litSize; d->litEntropy = 1;
        return litCSize + lhSize;
    }
    {   size_t lhSize, litSize;
        switch (lhl) {
        case 0: case 2: default: lhSize = 1; litSize = src[0] >> 3; break;
        case 1: lhSize = 2; litSize = zso_rd16(src) >> 4; break;
        case 3: lhSize = 3; if (srcSize < 3) return ZSO_ERROR(corruption_detected); litSize = zso_rd24(src) >> 4; break;
        }
        if (type == 1) {   /* rle :298-335 */
            if (lhl == 1 && srcSize < 3) return ZSO_ERROR(corruption_detected);
            if (lhl == 3 && srcSize < 4) return ZSO_ERROR(corruption_detected);
        }
        if (litSize > blockSizeMax) return ZSO_ERROR(corruption_detected);
        if (expectedWrite < litSize) return ZSO_ERROR(dstSize_tooSmall);
        if (type == 0) {   /* raw :250-296 */
            if (litSize + lhSize > srcSize) return ZSO_ERROR(corruption_detected);
            d->litPtr = src + lhSize; d->litSize = litSize;
            return lhSize + litSize;
        }
        memset(d->lit, src[lhSize], litSize);
        d->litPtr = d->lit; d->litSize = litSize;
        return lhSize + 1;
    }
}

/* ZSTD_buildSeqTable, :647-693 */
static size_t build_seq_table_from_stream(zso_seqTable* dt, unsigned type, unsigned max, unsigned maxLog,
                                          const uint8_t* src, size_t srcSize, int kind, int flagRepeat) {
    switch (type) {
    case 1:
        if (!srcSize) return ZSO_ERROR(srcSize_wrong);
        if (src[0] > max) return ZSO_ERROR(corruption_detected);
        build_seq_table_rle(dt, src[0], kind);
        return 1;
    case 0: {
        const int16_t* dn = kind == 0 ? zso_LL_defaultNorm : kind == 1 ? zso_OF_defaultNorm : zso_ML_defaultNorm;
        build_seq_table(dt, dn, kind == 1 ? ZSO_DefaultMaxOff : max, kind == 1 ? 5 : 6, kind);
        return 0; }
    case 3:
        if (!flagRepeat) return ZSO_ERROR(corruption_detected);
        return 0;
    default: {
        int16_t norm[64]; unsigned tableLog; unsigned m = max;
        size_t h = zso_readNCount(norm, &m, &tableLog, src, srcSize);
        if (zso_isError(h)) return ZSO_ERROR(corruption_detected);
        if (tableLog > maxLog) return ZSO_ERROR(corruption_detected);
        build_seq_table(dt, norm, m, tableLog, kind);
        return h; }
    }
}

/* ZSTD_decompressBlock_internal :2066-2174 + decodeSeqHeaders :695-775 +
 * decompressSequences_body :1615-1690 (decodeSequence :1229-1346, execSequence :1001-1096) */
static size_t decode_block(zso_dctx* d, uint8_t* dstBase, size_t written, size_t dstCapacityLeft,
                           const uint8_t* src, size_t srcSize, size_t blockSizeMax) {
    const uint8_t* ip = src; size_t left = srcSize; int nbSeq;
    uint8_t* const ostart = dstBase + written; uint8_t* op = ostart; uint8_t* const oend = ostart + dstCapacityLeft;
    if (srcSize > blockSizeMax) return ZSO_ERROR(srcSize_wrong);
    {   size_t r = decode_literals(d, ip, left, blockSizeMax, dstCapacityLeft);
        if (zso_isError(r)) return r;
        ip += r; left -= r;
    }
    if (left < 1) return ZSO_ERROR(srcSize_wrong);
    nbSeq = *ip++; left--;
    if (nbSeq > 0x7F) {
        if (nbSeq == 0xFF) { if (left < 2) return ZSO_ERROR(srcSize_wrong); nbSeq = zso_rd16(ip) + ZSO_LONGNBSEQ; ip += 2; left -= 2; }
        else { if (left < 1) return ZSO_ERROR(srcSize_wrong); nbSeq = ((nbSeq - 0x80) << 8) + *ip++; left--; }
    }
    if (nbSeq == 0) {
        if (left != 0) return ZSO_ERROR(corruption_detected);
    } else {
        unsigned LLtype, OFtype, MLtype; size_t r;
        if (left < 1) return ZSO_ERROR(srcSize_wrong);
        if (*ip & 3) return ZSO_ERROR(corruption_detected);
        LLtype = *ip >> 6; OFtype = (*ip >> 4) & 3; MLtype = (*ip >> 2) & 3; ip++; left--;
        r = build_seq_table_from_stream(&d->LL, LLtype, ZSO_MaxLL, ZSO_LLFSELog, ip, left, 0, d->fseEntropy);
        if (zso_isError(r)) return ZSO_ERROR(corruption_detected); ip += r; left -= r;
        r = build_seq_table_from_stream(&d->OF, OFtype, ZSO_MaxOff, ZSO_OffFSELog, ip, left, 1, d->fseEntropy);
        if (zso_isError(r)) return ZSO_ERROR(corruption_detected); ip += r; left -= r;
        r = build_seq_table_from_stream(&d->ML, MLtype, ZSO_MaxML, ZSO_MLFSELog, ip, left, 2, d->fseEntropy);
        if (zso_isError(r)) return ZSO_ERROR(corruption_detected); ip += r; left -= r;
        if (dstCapacityLeft == 0) return ZSO_ERROR(dstSize_tooSmall);
    }
    {   const uint8_t* lit = d->litPtr; const uint8_t* const litEnd = lit + d->litSize;
        if (nbSeq) {
            zso_br br; uint32_t sLL, sOF, sML; uint32_t rep0 = d->rep[0], rep1 = d->rep[1], rep2 = d->rep[2]; int n;
            d->fseEntropy = 1;
            if (zso_isError(br_init(&br, ip, left))) return ZSO_ERROR(corruption_detected);
            sLL = br_read(&br, d->LL.log); sOF = br_read(&br, d->OF.log); sML = br_read(&br, d->ML.log);
            for (n = nbSeq; n; n--) {
                zso_seqSym const ll = d->LL.t[sLL], ml = d->ML.t[sML], of = d->OF.t[sOF];
                size_t litLength = ll.baseValue, matchLength = ml.baseValue, offset;
                if (of.nbAddBits > 1) {
                    offset = of.baseValue + br_read(&br, of.nbAddBits);
                    rep2 = rep1; rep1 = rep0; rep0 = (uint32_t)offset;
                } else {
                    uint32_t const ll0 = (ll.baseValue == 0);
                    if (of.nbAddBits == 0) {
                        /* repcode 1 (or 2 when litLength==0) */
                        uint32_t const r[3] = { rep0, rep1, rep2 };
                        offset = r[ll0];
                        rep1 = r[!ll0]; rep0 = (uint32_t)offset;
                    } else {
                        uint32_t const r[3] = { rep0, rep1, rep2 };
                        uint32_t const idx = of.baseValue + ll0 + br_read(&br, 1);
                        uint32_t temp = (idx == 3) ? r[0] - 1 : r[idx];
                        temp -= !temp;                     /* 0 is invalid: becomes 0xFFFFFFFF => caught below */
                        if (idx != 1) rep2 = rep1;
                        rep1 = rep0; rep0 = temp; offset = temp;
                    }
                }
                if (ml.nbAddBits) matchLength += br_read(&br, ml.nbAddBits);
                if (ll.nbAddBits) litLength += br_read(&br, ll.nbAddBits);
                if (n > 1) {
                    sLL = ll.nextState + br_read(&br, ll.nbBits);
                    sML = ml.nextState + br_read(&br, ml.nbBits);
                    sOF = of.nextState + br_read(&br, of.nbBits);
                }
                /* execute (ZSTD_execSequence / _End) */
                if (litLength + matchLength > (size_t)(oend - op)) return ZSO_ERROR(dstSize_tooSmall);
                if (litLength > (size_t)(litEnd - lit)) return ZSO_ERROR(corruption_detected);
                memcpy(op, lit, litLength); op += litLength; lit += litLength;
                if (offset > (size_t)(op - dstBase)) return ZSO_ERROR(corruption_detected);
                {   const uint8_t* m = op - offset; size_t k;
                    for (k = 0; k < matchLength; k++) op[k] = m[k];
                    op += matchLength;
                }
            }
            if (!br_finished(&br)) return ZSO_ERROR(corruption_detected);
            d->rep[0] = rep0; d->rep[1] = rep1; d->rep[2] = rep2;
        }
        {   size_t const last = (size_t)(litEnd - lit);
            if (last > (size_t)(oend - op)) return ZSO_ERROR(dstSize_tooSmall);
            memcpy(op, lit, last); op += last;
        }
    }
    return (size_t)(op - ostart);
}

/* ----------------------------------------------------------------- XXH64
 * Published xxHash64 algorithm (N/common/xxhash.h); only needed when the frame
 * header's checksum flag is set (zstd-jni default: off, N/jni_zstd.c:21). */
#define XP1 0x9E3779B185EBCA87ULL
#define XP2 0xC2B2AE3D27D4EB4FULL
#define XP3 0x165667B19E3779F9ULL
#define XP4 0x85EBCA77C2B2AE63ULL
#define XP5 0x27D4EB2F165667C5ULL
static uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static uint64_t xround(uint64_t acc, uint64_t in) { acc += in * XP2; acc = rotl64(acc, 31); return acc * XP1; }
static uint64_t xmerge(uint64_t acc, uint64_t v) { v = xround(0, v); acc ^= v; return acc * XP1 + XP4; }
uint64_t zso_xxh64(const void* data, size_t len, uint64_t seed) {
    const uint8_t* p = (const uint8_t*)data; const uint8_t* const end = p + len; uint64_t h;
    if (len >= 32) {
        uint64_t v1 = seed + XP1 + XP2, v2 = seed + XP2, v3 = seed, v4 = seed - XP1;
        do { v1 = xround(v1, zso_rd64(p)); v2 = xround(v2, zso_rd64(p + 8)); v3 = xround(v3, zso_rd64(p + 16)); v4 = xround(v4, zso_rd64(p + 24)); p += 32; } while (p + 32 <= end);
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = xmerge(h, v1); h = xmerge(h, v2); h = xmerge(h, v3); h = xmerge(h, v4);
    } else h = seed + XP5;
    h += (uint64_t)len;
    while (p + 8 <= end) { h ^= xround(0, zso_rd64(p)); h = rotl64(h, 27) * XP1 + XP4; p += 8; }
    if (p + 4 <= end) { h ^= (uint64_t)zso_rd32(p) * XP1; h = rotl64(h, 23) * XP2 + XP3; p += 4; }
    while (p < end) { h ^= (*p++) * XP5; h = rotl64(h, 11) * XP1; }
    h ^= h >> 33; h *= XP2; h ^= h >> 29; h *= XP3; h ^= h >> 32;
    return h;
}

/* ------------------------------------------------------------ frame layer */
typedef struct { size_t headerSize; uint64_t contentSize; uint64_t windowSize; size_t blockSizeMax; int checksum; int skippable; uint32_t skipLen; int hasContentSize; uint32_t dictID; } zso_fh;

/* ZSTD_getFrameHeader_advanced, N/decompress/zstd_decompress.c:447-551. returns 0, or wanted size, or error */
static size_t read_frame_header(zso_fh* fh, const uint8_t* src, size_t srcSize) {
    static const uint8_t did[4] = { 0, 1, 2, 4 }, fcs[4] = { 0, 2, 4, 8 };
    memset(fh, 0, sizeof(*fh));
    if (srcSize < 5) {
        /* partial magic check (:458-473) */
        uint8_t h[4] = { 0x28, 0xB5, 0x2F, 0xFD }; size_t k = srcSize < 4 ? srcSize : 4;
        if (srcSize > 0) {
            memcpy(h, src, k);
            if (zso_rd32(h) != 0xFD2FB528u) {
                uint8_t s[4] = { 0x50, 0x2A, 0x4D, 0x18 }; memcpy(s, src, k);
                if ((zso_rd32(s) & 0xFFFFFFF0u) != 0x184D2A50u) return ZSO_ERROR(prefix_unknown);
            }
        }
        return 5;
    }
    if (zso_rd32(src) != 0xFD2FB528u) {
        if ((zso_rd32(src) & 0xFFFFFFF0u) == 0x184D2A50u) {
            if (srcSize < 8) return 8;
            fh->skippable = 1; fh->skipLen = zso_rd32(src + 4); fh->headerSize = 8;
            return 0;
        }
        return ZSO_ERROR(prefix_unknown);
    }
    {   uint8_t const fhd = src[4]; unsigned const dictID = fhd & 3, single = (fhd >> 5) & 1, fcsID = fhd >> 6; size_t pos = 5;
        size_t const hs = 5 + !single + did[dictID] + fcs[fcsID] + (single && !fcsID);
        if (srcSize < hs) return hs;
        fh->headerSize = hs;
        if (fhd & 0x08) return ZSO_ERROR(frameParameter_unsupported);
        if (!single) {
            uint8_t const wl = src[pos++]; unsigned const windowLog = (wl >> 3) + 10;
            if (windowLog > 31) return ZSO_ERROR(frameParameter_windowTooLarge);
            fh->windowSize = 1ULL << windowLog; fh->windowSize += (fh->windowSize >> 3) * (wl & 7);
        }
        {   unsigned i; for (i = 0; i < did[dictID]; i++) fh->dictID |= (uint32_t)src[pos + i] << (8 * i); }     /* :508-516 */
        pos += did[dictID];
        fh->hasContentSize = 1;
        switch (fcsID) {
        case 0: if (single) fh->contentSize = src[pos]; else fh->hasContentSize = 0; break;
        case 1: fh->contentSize = (uint64_t)zso_rd16(src + pos) + 256; break;
        case 2: fh->contentSize = zso_rd32(src + pos); break;
        default: fh->contentSize = zso_rd64(src + pos); break;
        }
        if (single) fh->windowSize = fh->contentSize;
        fh->blockSizeMax = (size_t)(fh->windowSize < ZSO_BLOCKSIZE_MAX ? fh->windowSize : ZSO_BLOCKSIZE_MAX);
        fh->checksum = (fhd >> 2) & 1;
    }
    return 0;
}

/* ZSTD_findFrameSizeInfo :734-797 */
size_t zso_findFrameCompressedSize(const void* srcv, size_t srcSize) {
    const uint8_t* src = (const uint8_t*)srcv; zso_fh fh; size_t r = read_frame_header(&fh, src, srcSize); size_t pos;
    if (zso_isError(r)) return r;
    if (r > 0) return ZSO_ERROR(srcSize_wrong);
    if (fh.skippable) { if ((uint64_t)fh.skipLen + 8 > srcSize) return ZSO_ERROR(srcSize_wrong); return 8 + (size_t)fh.skipLen; }
    pos = fh.headerSize;
    for (;;) {
        uint32_t bh; unsigned type; size_t cSize;
        if (srcSize - pos < 3) return ZSO_ERROR(srcSize_wrong);
        bh = zso_rd24(src + pos); type = (bh >> 1) & 3; cSize = bh >> 3;
        if (type == 3) return ZSO_ERROR(corruption_detected);
        if (type == 1) cSize = 1;
        if (3 + cSize > srcSize - pos) return ZSO_ERROR(srcSize_wrong);
        pos += 3 + cSize;
        if (bh & 1) break;
    }
    if (fh.checksum) { if (srcSize - pos < 4) return ZSO_ERROR(srcSize_wrong); pos += 4; }
    return pos;
}

#define ZSO_CONTENTSIZE_UNKNOWN (0ULL - 1)
#define ZSO_CONTENTSIZE_ERROR (0ULL - 2)
unsigned long long zso_getFrameContentSize(const void* src, size_t srcSize) {   /* :629-648 */
    zso_fh fh; size_t r = read_frame_header(&fh, (const uint8_t*)src, srcSize);
    if (zso_isError(r) || r > 0) return ZSO_CONTENTSIZE_ERROR;
    if (fh.skippable) return 0;
    return fh.hasContentSize ? fh.contentSize : ZSO_CONTENTSIZE_UNKNOWN;
}

/* ZSTD_decompressFrame :953-1066 */
static size_t decode_frame(zso_dctx* d, uint8_t* dst, size_t dstCapacity, const uint8_t** srcPtr, size_t* srcSizePtr) {
    const uint8_t* ip = *srcPtr; size_t left = *srcSizePtr; zso_fh fh; size_t written = 0;
    if (left < 6 + 3) {   /* ZSTD_FRAMEHEADERSIZE_MIN(6) + blockHeader */
        return ZSO_ERROR(srcSize_wrong);
    }
    {   /* :972-979: header size from the descriptor byte alone; too short for header + block header => srcSize_wrong before validation */
        static const uint8_t didSz[4] = { 0, 1, 2, 4 }, fcsSz[4] = { 0, 2, 4, 8 };
        uint8_t const fhd = ip[4]; unsigned const single = (fhd >> 5) & 1;
        size_t const hs0 = 5 + !single + didSz[fhd & 3] + fcsSz[fhd >> 6] + (single && !(fhd >> 6));
        if (left < hs0 + 3) return ZSO_ERROR(srcSize_wrong); }
    {   size_t r = read_frame_header(&fh, ip, left);
        if (zso_isError(r)) return r;
        if (r > 0) return ZSO_ERROR(srcSize_wrong);
        if (left < fh.headerSize + 3) return ZSO_ERROR(srcSize_wrong);
        /* ZSTD_decodeFrameHeader :706-707: the frame names a dictionary, none is loaded */
        if (fh.dictID != 0) return ZSO_ERROR(dictionary_wrong);
        ip += fh.headerSize; left -= fh.headerSize;
    }
    d->rep[0] = 1; d->rep[1] = 4; d->rep[2] = 8; d->litEntropy = d->fseEntropy = 0;
    for (;;) {
        uint32_t bh; unsigned type; size_t cSize, decoded;
        if (left < 3) return ZSO_ERROR(srcSize_wrong);
        bh = zso_rd24(ip); type = (bh >> 1) & 3; cSize = bh >> 3;
        if (type == 3) return ZSO_ERROR(corruption_detected);
        if (type == 1) cSize = 1;
        ip += 3; left -= 3;
        if (cSize > left) return ZSO_ERROR(srcSize_wrong);
        switch (type) {
        case 2: decoded = decode_block(d, dst, written, dstCapacity - written, ip, cSize, fh.blockSizeMax); break;
        case 0: if (cSize > dstCapacity - written) return ZSO_ERROR(dstSize_tooSmall);
                if (cSize) memmove(dst + written, ip, cSize); decoded = cSize; break;
        default: { size_t const rl = bh >> 3;
                if (rl > dstCapacity - written) return ZSO_ERROR(dstSize_tooSmall);
                if (rl) memset(dst + written, *ip, rl); decoded = rl; } break;
        }
        if (zso_isError(decoded)) return decoded;
        written += decoded; ip += cSize; left -= cSize;
        if (bh & 1) break;
    }
    if (fh.hasContentSize && written != fh.contentSize) return ZSO_ERROR(corruption_detected);
    if (fh.checksum) {
        if (left < 4) return ZSO_ERROR(checksum_wrong);
        if ((uint32_t)zso_xxh64(dst, written, 0) != zso_rd32(ip)) return ZSO_ERROR(checksum_wrong);
        ip += 4; left -= 4;
    }
    *srcPtr = ip; *srcSizePtr = left;
    return written;
}

#include <stdlib.h>
/* ZSTD_decompressMultiFrame :1070-1169 */
size_t zso_decompress(void* dstv, size_t dstCapacity, const void* srcv, size_t srcSize) {
    uint8_t* dst = (uint8_t*)dstv; const uint8_t* src = (const uint8_t*)srcv; size_t total = 0; int more = 0;
    zso_dctx* d = (zso_dctx*)malloc(sizeof(zso_dctx));
    if (!d) return ZSO_ERROR(GENERIC);
    while (srcSize >= 4) {   /* ZSTD_startingInputLength */
        if (srcSize >= 8 && (zso_rd32(src) & 0xFFFFFFF0u) == 0x184D2A50u) {
            size_t const skip = 8 + (size_t)zso_rd32(src + 4);
            if (skip > srcSize) { free(d); return ZSO_ERROR(srcSize_wrong); }
            src += skip; srcSize -= skip; continue;
        }
        {   size_t const r = decode_frame(d, dst, dstCapacity, &src, &srcSize);
            if (zso_isError(r)) {
                free(d);
                if (more && r == ZSO_ERROR(prefix_unknown)) return ZSO_ERROR(srcSize_wrong);   /* :1148-1158 */
                return r;
            }
            dst += r; dstCapacity -= r; total += r; more = 1;
        }
    }
    free(d);
    if (srcSize) return ZSO_ERROR(srcSize_wrong);
    return total;
}
