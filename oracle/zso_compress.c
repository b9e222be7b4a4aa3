/*
 * zso_compress.c -- CPU oracle: one-shot Zstandard frame encoder for inputs
 * of at most one block (<= 128 KB), strategies ZSTD_fast and ZSTD_dfast
 * (levels 1-4 at these sizes), byte-identical to the reference's
 * ZSTD_compress2() on the same input and level.
 * TEST INFRASTRUCTURE ONLY (see zso_common.h).
 *
 * Restates the encoder decisions of libzstd 1.5.7 (N/ = src/main/native/):
 *   parameters      N/compress/clevels.h, N/compress/zstd_compress.c:1472-1609,7759-7782
 *   framing         N/compress/zstd_compress.c:4591-4743,5344-5381
 *   match finding   N/compress/zstd_double_fast.c:105-323, N/compress/zstd_fast.c:190-423
 *   literals        N/compress/zstd_compress_literals.c:129-235, N/compress/huf_compress.c
 *   sequences       N/compress/zstd_compress.c:2693-3042, N/compress/zstd_compress_sequences.c,
 *                   N/compress/fse_compress.c
 */
#include "zso_common.h"
#include <stdlib.h>

/* ------------------------------------------------------------ parameters */
typedef struct { unsigned windowLog, chainLog, hashLog, searchLog, minMatch, targetLength, strategy; } zso_cparams;
enum { ZSO_fast = 1, ZSO_dfast = 2, ZSO_greedy = 3, ZSO_lazy = 4, ZSO_lazy2 = 5, ZSO_btlazy2 = 6, ZSO_btopt = 7, ZSO_btultra = 8, ZSO_btultra2 = 9 };

/* rows 0..12 of the two tables that apply to srcSize <= 128 KB (N/compress/clevels.h:78-130);
 * higher levels use binary-tree strategies that this oracle does not restate. */
static const zso_cparams k_params_128k[13] = {
    {17,12,12,1,5,1,ZSO_fast}, {17,12,13,1,6,0,ZSO_fast}, {17,13,15,1,5,0,ZSO_fast}, {17,15,16,2,5,0,ZSO_dfast},
    {17,17,17,2,4,0,ZSO_dfast}, {17,16,17,3,4,2,ZSO_greedy}, {17,16,17,3,4,4,ZSO_lazy}, {17,16,17,3,4,8,ZSO_lazy2},
    {17,16,17,4,4,8,ZSO_lazy2}, {17,16,17,5,4,8,ZSO_lazy2}, {17,16,17,6,4,8,ZSO_lazy2}, {17,17,17,5,4,8,ZSO_btlazy2},
    {17,18,17,7,4,12,ZSO_btlazy2} };
static const zso_cparams k_params_16k[13] = {
    {14,12,13,1,5,1,ZSO_fast}, {14,14,15,1,5,0,ZSO_fast}, {14,14,15,1,4,0,ZSO_fast}, {14,14,15,2,4,0,ZSO_dfast},
    {14,14,14,4,4,2,ZSO_greedy}, {14,14,14,3,4,4,ZSO_lazy}, {14,14,14,4,4,8,ZSO_lazy2}, {14,14,14,6,4,8,ZSO_lazy2},
    {14,14,14,8,4,8,ZSO_lazy2}, {14,15,14,5,4,8,ZSO_btlazy2}, {14,15,14,9,4,8,ZSO_btlazy2}, {14,15,14,3,4,12,ZSO_btopt},
    {14,15,14,4,3,24,ZSO_btopt} };

/* ZSTD_getCParams_internal :7759-7782 + ZSTD_adjustCParams_internal :1472-1609, for a known srcSize <= 128 KB, no dict */
int zso_getCParams(zso_cparams* out, int level, size_t srcSize) {
    zso_cparams cp; int row = level;
    if (srcSize > ZSO_BLOCKSIZE_MAX) return -1;
    if (level == 0) row = 3;
    if (level < 0) row = 0;
    if (row > 12) return -1;
    cp = (srcSize <= 16 * 1024) ? k_params_16k[row] : k_params_128k[row];
    if (level < 0) { int l = level < -(1 << 17) ? -(1 << 17) : level; cp.targetLength = (unsigned)(-l); }
    {   uint32_t const tSize = (uint32_t)srcSize;
        uint32_t const srcLog = (tSize < 64) ? 6 : zso_highbit32(tSize - 1) + 1;
        if (cp.windowLog > srcLog) cp.windowLog = srcLog;
    }
    {   uint32_t const dictAndWindowLog = cp.windowLog;
        uint32_t const cycleLog = cp.chainLog - (cp.strategy >= ZSO_btlazy2);
        if (cp.hashLog > dictAndWindowLog + 1) cp.hashLog = dictAndWindowLog + 1;
        if (cycleLog > dictAndWindowLog) cp.chainLog -= (cycleLog - dictAndWindowLog);
    }
    if (cp.windowLog < 10) cp.windowLog = 10;
    /* row-matchfinder hashLog cap (:1596-1606) only bites for hashLog > 24+rowLog: never here */
    *out = cp;
    return 0;
}

/* ZSTD_compressBound, N/zstd.h:249 */
size_t zso_compressBound(size_t n) { return n + (n >> 8) + (n < (128u << 10) ? (((128u << 10) - n) >> 11) : 0); }

/* --------------------------------------------------------------- seqStore */
typedef struct { uint32_t litLength, offBase, matchLength; } zso_seq;   /* matchLength is the full length (>=3) */
typedef struct {
    zso_seq* seq; size_t nbSeq;
    uint8_t* lit; size_t litSize;
} zso_seqStore;

static void store_seq(zso_seqStore* ss, const uint8_t* literals, size_t litLength, uint32_t offBase, size_t matchLength) {
    /* ZSTD_storeSeq, N/compress/zstd_compress_internal.h:773-811 */
    memcpy(ss->lit + ss->litSize, literals, litLength); ss->litSize += litLength;
    ss->seq[ss->nbSeq].litLength = (uint32_t)litLength;
    ss->seq[ss->nbSeq].offBase = offBase;
    ss->seq[ss->nbSeq].matchLength = (uint32_t)matchLength;
    ss->nbSeq++;
}

/* hashes, N/compress/zstd_compress_internal.h:898-945 */
static size_t hash_ptr(const uint8_t* p, unsigned hBits, unsigned mls) {
    switch (mls) {
    default:
    case 4: return (size_t)((zso_rd32(p) * 2654435761U) >> (32 - hBits));
    case 5: return (size_t)(((zso_rd64(p) << 24) * 889523592379ULL) >> (64 - hBits));
    case 6: return (size_t)(((zso_rd64(p) << 16) * 227718039650203ULL) >> (64 - hBits));
    case 7: return (size_t)(((zso_rd64(p) << 8) * 58295818150454627ULL) >> (64 - hBits));
    case 8: return (size_t)((zso_rd64(p) * 0xCF1BBCDCB7A56463ULL) >> (64 - hBits));
    }
}
/* ZSTD_count :854-873: common prefix length of in[] and match[], in bounded by end */
static size_t count_match(const uint8_t* in, const uint8_t* match, const uint8_t* end) {
    const uint8_t* const s = in;
    while (in < end && *in == *match) { in++; match++; }
    return (size_t)(in - s);
}

/* ZSTD_compressBlock_doubleFast_noDict_generic, N/compress/zstd_double_fast.c:105-323.
 * Positions are expressed as indices into the table space of a fresh frame:
 * index = position + 2 (ZSTD_WINDOW_START_INDEX, zstd_compress_internal.h:266),
 * so prefixLowestIndex == 2 and zeroed table cells are "no candidate".
 * Returns the length of the trailing literal run. */
static size_t block_dfast(zso_seqStore* ss, uint32_t rep[3], const uint8_t* src, size_t srcSize,
                          uint32_t* hashLong, unsigned hBitsL, uint32_t* hashSmall, unsigned hBitsS, unsigned mls) {
    const uint8_t* const base = src - 2;
    const uint8_t* const istart = src;
    const uint8_t* const iend = src + srcSize;
    const uint8_t* const ilimit = iend - 8;
    const uint8_t* const prefixLowest = src;          /* base + prefixLowestIndex */
    uint32_t const prefixLowestIndex = 2;
    const uint8_t* anchor = istart;
    const uint8_t* ip = istart;
    uint32_t offset_1 = rep[0], offset_2 = rep[1], offsetSaved1 = 0, offsetSaved2 = 0;

    ip += (ip == prefixLowest);
    {   uint32_t const maxRep = (uint32_t)(ip - prefixLowest);       /* current - windowLow (:159-163) */
        if (offset_2 > maxRep) { offsetSaved2 = offset_2; offset_2 = 0; }
        if (offset_1 > maxRep) { offsetSaved1 = offset_1; offset_1 = 0; }
    }
    for (;;) {   /* one iteration per stored match */
        size_t step = 1; const uint8_t* nextStep = ip + 256; const uint8_t* ip1 = ip + step;
        size_t hl0, hl1 = 0, mLength; uint32_t idxl0, idxl1 = 0, curr = 0, offset = 0;
        const uint8_t* match;
        if (ip1 > ilimit) goto cleanup;
        hl0 = hash_ptr(ip, hBitsL, 8); idxl0 = hashLong[hl0];
        for (;;) {   /* one iteration per searched position */
            size_t const hs0 = hash_ptr(ip, hBitsS, mls);
            uint32_t const idxs0 = hashSmall[hs0];
            curr = (uint32_t)(ip - base);
            hashLong[hl0] = hashSmall[hs0] = curr;
            /* repcode at ip+1 (:190-195) */
            if ((offset_1 > 0) & (zso_rd32(ip + 1 - offset_1) == zso_rd32(ip + 1))) {
                mLength = count_match(ip + 1 + 4, ip + 1 + 4 - offset_1, iend) + 4;
                ip++;
                store_seq(ss, anchor, (size_t)(ip - anchor), 1 /* REPCODE1_TO_OFFBASE */, mLength);
                goto match_stored;
            }
            hl1 = hash_ptr(ip1, hBitsL, 8);
            /* long match at ip (:203-211); ZSTD_selectAddr keeps the candidate iff index >= lowLimit */
            if (idxl0 >= prefixLowestIndex && zso_rd64(base + idxl0) == zso_rd64(ip)) {
                match = base + idxl0;
                mLength = count_match(ip + 8, match + 8, iend) + 8;
                offset = (uint32_t)(ip - match);
                while (((ip > anchor) & (match > prefixLowest)) && (ip[-1] == match[-1])) { ip--; match--; mLength++; }
                goto match_found;
            }
            idxl1 = hashLong[hl1];
            /* short match at ip (:217-222) */
            if (idxs0 >= prefixLowestIndex && zso_rd32(base + idxs0) == zso_rd32(ip)) {
                match = base + idxs0;
                goto search_next_long;
            }
            if (ip1 >= nextStep) { step++; nextStep += 256; }
            ip = ip1; ip1 += step;
            hl0 = hl1; idxl0 = idxl1;
            if (ip1 > ilimit) goto cleanup;
        }
search_next_long:
        mLength = count_match(ip + 4, match + 4, iend) + 4;
        offset = (uint32_t)(ip - match);
        /* long match at ip1 (:260-269); note the strict > here */
        if ((idxl1 > prefixLowestIndex) && (zso_rd64(base + idxl1) == zso_rd64(ip1))) {
            const uint8_t* const matchl1 = base + idxl1;
            size_t const l1len = count_match(ip1 + 8, matchl1 + 8, iend) + 8;
            if (l1len > mLength) { ip = ip1; mLength = l1len; offset = (uint32_t)(ip - matchl1); match = matchl1; }
        }
        while (((ip > anchor) & (match > prefixLowest)) && (ip[-1] == match[-1])) { ip--; match--; mLength++; }
match_found:
        offset_2 = offset_1; offset_1 = offset;
        if (step < 4) hashLong[hl1] = (uint32_t)(ip1 - base);     /* :279-288 */
        store_seq(ss, anchor, (size_t)(ip - anchor), offset + 3 /* OFFSET_TO_OFFBASE */, mLength);
match_stored:
        ip += mLength; anchor = ip;
        if (ip <= ilimit) {
            uint32_t const indexToInsert = curr + 2;               /* :297-305 */
            hashLong[hash_ptr(base + indexToInsert, hBitsL, 8)] = indexToInsert;
            hashLong[hash_ptr(ip - 2, hBitsL, 8)] = (uint32_t)(ip - 2 - base);
            hashSmall[hash_ptr(base + indexToInsert, hBitsS, mls)] = indexToInsert;
            hashSmall[hash_ptr(ip - 1, hBitsS, mls)] = (uint32_t)(ip - 1 - base);
            while ((ip <= ilimit) && ((offset_2 > 0) & (zso_rd32(ip) == zso_rd32(ip - offset_2)))) {   /* :308-320 */
                size_t const rLength = count_match(ip + 4, ip + 4 - offset_2, iend) + 4;
                uint32_t const tmp = offset_2; offset_2 = offset_1; offset_1 = tmp;
                hashSmall[hash_ptr(ip, hBitsS, mls)] = (uint32_t)(ip - base);
                hashLong[hash_ptr(ip, hBitsL, 8)] = (uint32_t)(ip - base);
                store_seq(ss, anchor, 0, 1, rLength);
                ip += rLength; anchor = ip;
            }
        }
    }
cleanup:
    offsetSaved2 = ((offsetSaved1 != 0) && (offset_1 != 0)) ? offsetSaved1 : offsetSaved2;   /* :244 */
    rep[0] = offset_1 ? offset_1 : offsetSaved1;
    rep[1] = offset_2 ? offset_2 : offsetSaved2;
    return (size_t)(iend - anchor);
}

/* ----------------------------------------------------- forward bit writer
 * LSB-first append; closing adds a single 1 bit (BIT_closeCStream,
 * N/common/bitstream.h:226-242 and HUF_closeCStream, huf_compress.c:973-982). */
typedef struct { uint8_t* p; size_t cap; size_t nbytes; uint64_t acc; unsigned nb; int overflow; } zso_bw;
static void bw_init(zso_bw* w, uint8_t* dst, size_t cap) { w->p = dst; w->cap = cap; w->nbytes = 0; w->acc = 0; w->nb = 0; w->overflow = 0; }
static void bw_add(zso_bw* w, uint64_t value, unsigned nbBits) {
    if (!nbBits) return;
    value &= (nbBits >= 64) ? ~0ULL : ((1ULL << nbBits) - 1);
    w->acc |= value << w->nb; w->nb += nbBits;
    while (w->nb >= 8) {
        if (w->nbytes < w->cap) w->p[w->nbytes] = (uint8_t)w->acc; else w->overflow = 1;
        w->nbytes++; w->acc >>= 8; w->nb -= 8;
    }
}
static size_t bw_close(zso_bw* w) {   /* returns stream size in bytes, 0 if it did not fit */
    bw_add(w, 1, 1);
    /* the reference keeps 8 bytes of slack: endPtr = start + cap - 8 and a stream whose
     * whole-byte count reaches endPtr is reported as "does not fit" (bitstream.h:232-241) */
    if (w->cap <= 8 || w->nbytes >= w->cap - 8) return 0;
    if (w->nb) { w->p[w->nbytes] = (uint8_t)w->acc; return w->nbytes + 1; }
    return w->nbytes;
}

/* ------------------------------------------------------------ histogram
 * HIST_count_simple, N/compress/hist.c:39-74 (the parallel variant :76-148 computes the same values) */
static unsigned hist(unsigned* count, unsigned* maxSV, const uint8_t* src, size_t n) {
    unsigned m = *maxSV, s, largest = 0; size_t i;
    memset(count, 0, (m + 1) * sizeof(*count));
    if (n == 0) { *maxSV = 0; return 0; }
    for (i = 0; i < n; i++) count[src[i]]++;
    while (!count[m]) m--;
    *maxSV = m;
    for (s = 0; s <= m; s++) if (count[s] > largest) largest = count[s];
    return largest;
}

/* --------------------------------------------------------- FSE (encoder) */
/* FSE_optimalTableLog_internal, N/compress/fse_compress.c:348-369 */
static unsigned fse_optimalTableLog(unsigned maxTableLog, size_t srcSize, unsigned maxSV, unsigned minus) {
    uint32_t const maxBitsSrc = zso_highbit32((uint32_t)(srcSize - 1)) - minus;
    uint32_t tableLog = maxTableLog;
    uint32_t const minBitsSrc = zso_highbit32((uint32_t)srcSize) + 1;
    uint32_t const minBitsSymbols = zso_highbit32(maxSV) + 2;
    uint32_t const minBits = minBitsSrc < minBitsSymbols ? minBitsSrc : minBitsSymbols;
    if (tableLog == 0) tableLog = 11;
    if (maxBitsSrc < tableLog) tableLog = maxBitsSrc;
    if (minBits > tableLog) tableLog = minBits;
    if (tableLog < 5) tableLog = 5;
    if (tableLog > 12) tableLog = 12;
    return tableLog;
}

/* FSE_normalizeM2, :379-463 */
static size_t fse_normalizeM2(int16_t* norm, uint32_t tableLog, const unsigned* count, size_t total, uint32_t maxSV, int16_t lowProbCount) {
    int16_t const NOT_YET = -2; uint32_t s, distributed = 0, ToDistribute;
    uint32_t const lowThreshold = (uint32_t)(total >> tableLog);
    uint32_t lowOne = (uint32_t)((total * 3) >> (tableLog + 1));
    for (s = 0; s <= maxSV; s++) {
        if (count[s] == 0) { norm[s] = 0; continue; }
        if (count[s] <= lowThreshold) { norm[s] = lowProbCount; distributed++; total -= count[s]; continue; }
        if (count[s] <= lowOne) { norm[s] = 1; distributed++; total -= count[s]; continue; }
        norm[s] = NOT_YET;
    }
    ToDistribute = (1u << tableLog) - distributed;
    if (ToDistribute == 0) return 0;
    if ((total / ToDistribute) > lowOne) {
        lowOne = (uint32_t)((total * 3) / (ToDistribute * 2));
        for (s = 0; s <= maxSV; s++)
            if ((norm[s] == NOT_YET) && (count[s] <= lowOne)) { norm[s] = 1; distributed++; total -= count[s]; }
        ToDistribute = (1u << tableLog) - distributed;
    }
    if (distributed == maxSV + 1) {
        uint32_t maxV = 0, maxC = 0;
        for (s = 0; s <= maxSV; s++) if (count[s] > maxC) { maxV = s; maxC = count[s]; }
        norm[maxV] += (int16_t)ToDistribute;
        return 0;
    }
    if (total == 0) {
        for (s = 0; ToDistribute > 0; s = (s + 1) % (maxSV + 1))
            if (norm[s] > 0) { ToDistribute--; norm[s]++; }
        return 0;
    }
    {   uint64_t const vStepLog = 62 - tableLog;
        uint64_t const mid = (1ULL << (vStepLog - 1)) - 1;
        uint64_t const rStep = ((((uint64_t)1 << vStepLog) * ToDistribute) + mid) / (uint32_t)total;
        uint64_t tmpTotal = mid;
        for (s = 0; s <= maxSV; s++) {
            if (norm[s] == NOT_YET) {
                uint64_t const end = tmpTotal + (count[s] * rStep);
                uint32_t const sStart = (uint32_t)(tmpTotal >> vStepLog), sEnd = (uint32_t)(end >> vStepLog);
                uint32_t const weight = sEnd - sStart;
                if (weight < 1) return ZSO_ERROR(GENERIC);
                norm[s] = (int16_t)weight; tmpTotal = end;
            }
        }
    }
    return 0;
}

/* FSE_normalizeCount, :465-525 */
static size_t fse_normalizeCount(int16_t* norm, unsigned tableLog, const unsigned* count, size_t total, unsigned maxSV, unsigned useLowProbCount) {
    static uint32_t const rtbTable[] = { 0, 473195, 504333, 520860, 550000, 700000, 750000, 830000 };
    int16_t const lowProbCount = useLowProbCount ? -1 : 1;
    uint64_t const scale = 62 - tableLog;
    uint64_t const step = ((uint64_t)1 << 62) / (uint32_t)total;
    uint64_t const vStep = 1ULL << (scale - 20);
    int stillToDistribute = 1 << tableLog;
    unsigned s, largest = 0; int16_t largestP = 0;
    uint32_t const lowThreshold = (uint32_t)(total >> tableLog);
    if (tableLog < 5) return ZSO_ERROR(GENERIC);
    if (tableLog > 12) return ZSO_ERROR(tableLog_tooLarge);
    {   uint32_t const minBitsSrc = zso_highbit32((uint32_t)total) + 1, minBitsSymbols = zso_highbit32(maxSV) + 2;
        if (tableLog < (minBitsSrc < minBitsSymbols ? minBitsSrc : minBitsSymbols)) return ZSO_ERROR(GENERIC);
    }
    for (s = 0; s <= maxSV; s++) {
        if (count[s] == total) return 0;
        if (count[s] == 0) { norm[s] = 0; continue; }
        if (count[s] <= lowThreshold) { norm[s] = lowProbCount; stillToDistribute--; }
        else {
            int16_t proba = (int16_t)((count[s] * step) >> scale);
            if (proba < 8) {
                uint64_t const restToBeat = vStep * rtbTable[proba];
                proba += (count[s] * step) - ((uint64_t)proba << scale) > restToBeat;
            }
            if (proba > largestP) { largestP = proba; largest = s; }
            norm[s] = proba; stillToDistribute -= proba;
        }
    }
    if (-stillToDistribute >= (norm[largest] >> 1)) {
        size_t const e = fse_normalizeM2(norm, tableLog, count, total, maxSV, lowProbCount);
        if (zso_isError(e)) return e;
    } else norm[largest] += (int16_t)stillToDistribute;
    return tableLog;
}

/* FSE_writeNCount_generic, :233-327, on the bit writer (the reference's 16-bit
 * flushes only matter for buffer bounds; the produced bit sequence is the same) */
static size_t fse_writeNCount(uint8_t* dst, size_t cap, const int16_t* norm, unsigned maxSV, unsigned tableLog) {
    zso_bw w; int nbBits, remaining, threshold, previousIs0 = 0; unsigned symbol = 0; unsigned const alphabetSize = maxSV + 1;
    int const tableSize = 1 << tableLog;
    bw_init(&w, dst, cap);
    bw_add(&w, tableLog - 5, 4);
    remaining = tableSize + 1; threshold = tableSize; nbBits = (int)tableLog + 1;
    while ((symbol < alphabetSize) && (remaining > 1)) {
        if (previousIs0) {
            unsigned start = symbol;
            while ((symbol < alphabetSize) && !norm[symbol]) symbol++;
            if (symbol == alphabetSize) break;
            while (symbol >= start + 24) { start += 24; bw_add(&w, 0xFFFF, 16); }
            while (symbol >= start + 3) { start += 3; bw_add(&w, 3, 2); }
            bw_add(&w, symbol - start, 2);
        }
        {   int count = norm[symbol++];
            int const max = (2 * threshold - 1) - remaining;
            remaining -= count < 0 ? -count : count;
            count++;
            if (count >= threshold) count += max;
            bw_add(&w, (uint64_t)count, (unsigned)(nbBits - (count < max)));
            previousIs0 = (count == 1);
            if (remaining < 1) return ZSO_ERROR(GENERIC);
            while (remaining < threshold) { nbBits--; threshold >>= 1; }
        }
    }
    if (remaining != 1) return ZSO_ERROR(GENERIC);
    /* flush: (bitCount+7)/8 bytes, no end mark */
    if (w.nb) { if (w.nbytes < w.cap) w.p[w.nbytes] = (uint8_t)w.acc; else w.overflow = 1; w.nbytes++; }
    if (w.overflow) return ZSO_ERROR(dstSize_tooSmall);
    return w.nbytes;
}

/* FSE compression table, FSE_buildCTable_wksp :68-214 */
typedef struct { int deltaFindState; uint32_t deltaNbBits; } zso_symTT;
typedef struct { unsigned tableLog; uint16_t stateTable[1 << 9]; zso_symTT tt[64]; } zso_ctable;   /* tableLog <= 9 on this path */

static void fse_buildCTable(zso_ctable* ct, const int16_t* norm, unsigned maxSV, unsigned tableLog) {
    uint32_t const tableSize = 1u << tableLog, mask = tableSize - 1, step = (tableSize >> 1) + (tableSize >> 3) + 3;
    uint16_t cumul[66]; uint8_t tableSymbol[1 << 9]; uint32_t high = tableSize - 1, u, pos = 0, s;
    ct->tableLog = tableLog;
    cumul[0] = 0;
    for (u = 1; u <= maxSV + 1; u++) {
        if (norm[u - 1] == -1) { cumul[u] = cumul[u - 1] + 1; tableSymbol[high--] = (uint8_t)(u - 1); }
        else cumul[u] = cumul[u - 1] + (uint16_t)norm[u - 1];
    }
    cumul[maxSV + 1] = (uint16_t)(tableSize + 1);
    for (s = 0; s <= maxSV; s++) {
        int i;
        for (i = 0; i < norm[s]; i++) {
            tableSymbol[pos] = (uint8_t)s;
            pos = (pos + step) & mask;
            while (pos > high) pos = (pos + step) & mask;
        }
    }
    for (u = 0; u < tableSize; u++) { uint8_t const sy = tableSymbol[u]; ct->stateTable[cumul[sy]++] = (uint16_t)(tableSize + u); }
    {   unsigned total = 0;
        for (s = 0; s <= maxSV; s++) {
            switch (norm[s]) {
            case 0: ct->tt[s].deltaNbBits = ((tableLog + 1) << 16) - (1u << tableLog); ct->tt[s].deltaFindState = 0; break;
            case -1: case 1:
                ct->tt[s].deltaNbBits = (tableLog << 16) - (1u << tableLog);
                ct->tt[s].deltaFindState = (int)(total - 1); total++; break;
            default: {
                uint32_t const maxBitsOut = tableLog - zso_highbit32((uint32_t)norm[s] - 1);
                uint32_t const minStatePlus = (uint32_t)norm[s] << maxBitsOut;
                ct->tt[s].deltaNbBits = (maxBitsOut << 16) - minStatePlus;
                ct->tt[s].deltaFindState = (int)(total - (unsigned)norm[s]);
                total += (unsigned)norm[s]; }
            }
        }
    }
}
static void fse_buildCTable_rle(zso_ctable* ct, unsigned symbol) {   /* :528-548 */
    ct->tableLog = 0; ct->stateTable[0] = 0; ct->stateTable[1] = 0;
    ct->tt[symbol].deltaNbBits = 0; ct->tt[symbol].deltaFindState = 0;
}
/* FSE_initCState2 / FSE_encodeSymbol / FSE_flushCState, N/common/fse.h:428-467 */
static uint32_t fse_init_state2(const zso_ctable* ct, unsigned symbol) {
    zso_symTT const tt = ct->tt[symbol];
    uint32_t const nbBitsOut = (tt.deltaNbBits + (1 << 15)) >> 16;
    uint32_t const v = (nbBitsOut << 16) - tt.deltaNbBits;
    return ct->stateTable[(int)(v >> nbBitsOut) + tt.deltaFindState];
}
static uint32_t fse_encode(zso_bw* w, const zso_ctable* ct, uint32_t state, unsigned symbol) {
    zso_symTT const tt = ct->tt[symbol];
    uint32_t const nbBitsOut = (state + tt.deltaNbBits) >> 16;
    bw_add(w, state, nbBitsOut);
    return ct->stateTable[(int)(state >> nbBitsOut) + tt.deltaFindState];
}

/* ---------------------------------------------------------- Huffman encoder */
typedef struct { uint32_t count; uint16_t parent; uint8_t byte; uint8_t nbBits; } zso_node;

/* HUF_sort and helpers, N/compress/huf_compress.c:530-665.  Buckets 0..164 hold
 * one distinct count each; larger counts share log2 buckets which are sorted
 * with the reference's own (unstable) quicksort -- the permutation of
 * equal-count symbols it leaves behind decides code assignment, so it is
 * restated operation for operation. */
#define RANK_TABLE 192
#define LOG_BUCKETS_BEGIN 158                              /* (192-1) - 32 - 1 */
#define DISTINCT_CUTOFF (LOG_BUCKETS_BEGIN + 7)            /* + highbit32(158) = 165 */
static uint32_t huf_bucket(uint32_t count) { return count < DISTINCT_CUTOFF ? count : zso_highbit32(count) + LOG_BUCKETS_BEGIN; }
static void node_swap(zso_node* a, zso_node* b) { zso_node t = *a; *a = *b; *b = t; }
static void huf_insertion(zso_node* a, int low, int high) {
    int i, size = high - low + 1; a += low;
    for (i = 1; i < size; i++) { zso_node key = a[i]; int j = i - 1; while (j >= 0 && a[j].count < key.count) { a[j + 1] = a[j]; j--; } a[j + 1] = key; }
}
static int huf_partition(zso_node* a, int low, int high) {
    uint32_t const pivot = a[high].count; int i = low - 1, j;
    for (j = low; j < high; j++) if (a[j].count > pivot) { i++; node_swap(&a[i], &a[j]); }
    node_swap(&a[i + 1], &a[high]);
    return i + 1;
}
static void huf_quicksort(zso_node* a, int low, int high) {
    if (high - low < 8) { huf_insertion(a, low, high); return; }
    while (low < high) {
        int const idx = huf_partition(a, low, high);
        if (idx - low < high - idx) { huf_quicksort(a, low, idx - 1); low = idx + 1; }
        else { huf_quicksort(a, idx + 1, high); high = idx - 1; }
    }
}
static void huf_sort(zso_node* node, const unsigned* count, uint32_t maxSV) {
    struct { uint16_t base, curr; } rp[RANK_TABLE]; uint32_t n;
    memset(rp, 0, sizeof(rp));
    for (n = 0; n <= maxSV; n++) rp[huf_bucket(count[n])].base++;
    for (n = RANK_TABLE - 1; n > 0; n--) { rp[n - 1].base += rp[n].base; rp[n - 1].curr = rp[n - 1].base; }
    for (n = 0; n <= maxSV; n++) {
        uint32_t const c = count[n], r = huf_bucket(c) + 1, pos = rp[r].curr++;
        node[pos].count = c; node[pos].byte = (uint8_t)n;
    }
    for (n = DISTINCT_CUTOFF; n < RANK_TABLE - 1; n++) {
        int const bucketSize = rp[n].curr - rp[n].base;
        if (bucketSize > 1) huf_quicksort(node + rp[n].base, 0, bucketSize - 1);
    }
}

/* HUF_setMaxHeight, :376-498 */
static uint32_t huf_setMaxHeight(zso_node* node, uint32_t lastNonNull, uint32_t target) {
    uint32_t const largestBits = node[lastNonNull].nbBits;
    if (largestBits <= target) return largestBits;
    {   int totalCost = 0; uint32_t const baseCost = 1u << (largestBits - target); int n = (int)lastNonNull;
        while (node[n].nbBits > target) { totalCost += (int)(baseCost - (1u << (largestBits - node[n].nbBits))); node[n].nbBits = (uint8_t)target; n--; }
        while (node[n].nbBits == target) --n;
        totalCost >>= (largestBits - target);
        {   uint32_t const noSymbol = 0xF0F0F0F0; uint32_t rankLast[ZSO_HUF_TABLELOG_MAX + 2]; uint32_t currentNbBits = target; int pos;
            memset(rankLast, 0xF0, sizeof(rankLast));
            for (pos = n; pos >= 0; pos--) {
                if (node[pos].nbBits >= currentNbBits) continue;
                currentNbBits = node[pos].nbBits;
                rankLast[target - currentNbBits] = (uint32_t)pos;
            }
            while (totalCost > 0) {
                uint32_t nBitsToDecrease = zso_highbit32((uint32_t)totalCost) + 1;
                for (; nBitsToDecrease > 1; nBitsToDecrease--) {
                    uint32_t const highPos = rankLast[nBitsToDecrease], lowPos = rankLast[nBitsToDecrease - 1];
                    if (highPos == noSymbol) continue;
                    if (lowPos == noSymbol) break;
                    {   uint32_t const highTotal = node[highPos].count, lowTotal = 2 * node[lowPos].count;
                        if (highTotal <= lowTotal) break; }
                }
                while ((nBitsToDecrease <= ZSO_HUF_TABLELOG_MAX) && (rankLast[nBitsToDecrease] == noSymbol)) nBitsToDecrease++;
                totalCost -= 1 << (nBitsToDecrease - 1);
                node[rankLast[nBitsToDecrease]].nbBits++;
                if (rankLast[nBitsToDecrease - 1] == noSymbol) rankLast[nBitsToDecrease - 1] = rankLast[nBitsToDecrease];
                if (rankLast[nBitsToDecrease] == 0) rankLast[nBitsToDecrease] = noSymbol;
                else {
                    rankLast[nBitsToDecrease]--;
                    if (node[rankLast[nBitsToDecrease]].nbBits != target - nBitsToDecrease) rankLast[nBitsToDecrease] = noSymbol;
                }
            }
            while (totalCost < 0) {
                if (rankLast[1] == noSymbol) {
                    while (node[n].nbBits == target) n--;
                    node[n + 1].nbBits--; rankLast[1] = (uint32_t)(n + 1); totalCost++;
                    continue;
                }
                node[rankLast[1] + 1].nbBits--; rankLast[1]++; totalCost++;
            }
        }
    }
    return target;
}

typedef struct { uint8_t nbBits[256]; uint16_t code[256]; unsigned tableLog; unsigned maxSV; } zso_hufC;

/* HUF_buildCTable_wksp :755-791 = sort + HUF_buildTree :681-718 + setMaxHeight + HUF_buildCTableFromTree :730-753 */
static uint32_t huf_buildCTable(zso_hufC* ct, const unsigned* count, uint32_t maxSV, uint32_t maxNbBits) {
    zso_node table[2 * 256 + 2]; zso_node* const node0 = table; zso_node* const node = table + 1;
    int nonNull, lowS, lowN, nodeNb = 256, n, nodeRoot;
    memset(table, 0, sizeof(table));
    huf_sort(node, count, maxSV);
    nonNull = (int)maxSV;
    while (node[nonNull].count == 0) nonNull--;
    lowS = nonNull; nodeRoot = nodeNb + lowS - 1; lowN = nodeNb;
    node[nodeNb].count = node[lowS].count + node[lowS - 1].count;
    node[lowS].parent = node[lowS - 1].parent = (uint16_t)nodeNb;
    nodeNb++; lowS -= 2;
    for (n = nodeNb; n <= nodeRoot; n++) node[n].count = 1u << 30;
    node0[0].count = 1u << 31;
    while (nodeNb <= nodeRoot) {
        int const n1 = (node[lowS].count < node[lowN].count) ? lowS-- : lowN++;
        int const n2 = (node[lowS].count < node[lowN].count) ? lowS-- : lowN++;
        node[nodeNb].count = node[n1].count + node[n2].count;
        node[n1].parent = node[n2].parent = (uint16_t)nodeNb;
        nodeNb++;
    }
    node[nodeRoot].nbBits = 0;
    for (n = nodeRoot - 1; n >= 256; n--) node[n].nbBits = node[node[n].parent].nbBits + 1;
    for (n = 0; n <= nonNull; n++) node[n].nbBits = node[node[n].parent].nbBits + 1;
    maxNbBits = huf_setMaxHeight(node, (uint32_t)nonNull, maxNbBits);
    {   uint16_t nbPerRank[ZSO_HUF_TABLELOG_MAX + 1] = { 0 }, valPerRank[ZSO_HUF_TABLELOG_MAX + 1] = { 0 }; uint16_t min = 0;
        int const alphabetSize = (int)(maxSV + 1);
        for (n = 0; n <= nonNull; n++) nbPerRank[node[n].nbBits]++;
        for (n = (int)maxNbBits; n > 0; n--) { valPerRank[n] = min; min += nbPerRank[n]; min >>= 1; }
        for (n = 0; n < alphabetSize; n++) ct->nbBits[node[n].byte] = node[n].nbBits;
        for (n = 0; n < alphabetSize; n++) ct->code[n] = ct->nbBits[n] ? valPerRank[ct->nbBits[n]]++ : 0;
    }
    ct->tableLog = maxNbBits; ct->maxSV = maxSV;
    return maxNbBits;
}

/* FSE_compress_usingCTable_generic, N/compress/fse_compress.c:551-608 (2 interleaved states) */
static size_t fse_compress_2states(uint8_t* dst, size_t cap, const uint8_t* src, size_t srcSize, const zso_ctable* ct) {
    const uint8_t* ip = src + srcSize; zso_bw w; uint32_t s1, s2;
    if (srcSize <= 2) return 0;
    if (cap <= 8) return 0;
    bw_init(&w, dst, cap);
    if (srcSize & 1) { s1 = fse_init_state2(ct, *--ip); s2 = fse_init_state2(ct, *--ip); s1 = fse_encode(&w, ct, s1, *--ip); }
    else { s2 = fse_init_state2(ct, *--ip); s1 = fse_init_state2(ct, *--ip); }
    srcSize -= 2;
    if (srcSize & 2) { s2 = fse_encode(&w, ct, s2, *--ip); s1 = fse_encode(&w, ct, s1, *--ip); }
    while (ip > src) {
        s2 = fse_encode(&w, ct, s2, *--ip); s1 = fse_encode(&w, ct, s1, *--ip);
        s2 = fse_encode(&w, ct, s2, *--ip); s1 = fse_encode(&w, ct, s1, *--ip);
    }
    bw_add(&w, s2, ct->tableLog); bw_add(&w, s1, ct->tableLog);
    return bw_close(&w);
}

/* HUF_compressWeights :146-186 */
static size_t huf_compressWeights(uint8_t* dst, size_t cap, const uint8_t* weights, size_t wtSize) {
    unsigned count[ZSO_HUF_TABLELOG_MAX + 1]; int16_t norm[ZSO_HUF_TABLELOG_MAX + 1]; unsigned maxSV = ZSO_HUF_TABLELOG_MAX, tableLog; zso_ctable ct; size_t h, c;
    if (wtSize <= 1) return 0;
    {   unsigned const maxCount = hist(count, &maxSV, weights, wtSize);
        if (maxCount == wtSize) return 1;
        if (maxCount == 1) return 0; }
    tableLog = fse_optimalTableLog(6, wtSize, maxSV, 2);
    {   size_t e = fse_normalizeCount(norm, tableLog, count, wtSize, maxSV, 0); if (zso_isError(e)) return e; }
    h = fse_writeNCount(dst, cap, norm, maxSV, tableLog); if (zso_isError(h)) return h;
    fse_buildCTable(&ct, norm, maxSV, tableLog);
    c = fse_compress_2states(dst + h, cap - h, weights, wtSize, &ct);
    if (c == 0) return 0;
    return h + c;
}

/* HUF_writeCTable_wksp :248-289 */
static size_t huf_writeCTable(uint8_t* dst, size_t cap, const zso_hufC* ct) {
    uint8_t w[256]; unsigned n; unsigned const maxSV = ct->maxSV, huffLog = ct->tableLog; size_t hSize;
    for (n = 0; n < maxSV; n++) w[n] = ct->nbBits[n] ? (uint8_t)(huffLog + 1 - ct->nbBits[n]) : 0;
    if (cap < 1) return ZSO_ERROR(dstSize_tooSmall);
    hSize = huf_compressWeights(dst + 1, cap - 1, w, maxSV);
    if (zso_isError(hSize)) return hSize;
    if ((hSize > 1) & (hSize < maxSV / 2)) { dst[0] = (uint8_t)hSize; return hSize + 1; }
    if (maxSV > 128) return ZSO_ERROR(GENERIC);
    if (((maxSV + 1) / 2) + 1 > cap) return ZSO_ERROR(dstSize_tooSmall);
    dst[0] = (uint8_t)(128 + (maxSV - 1));
    w[maxSV] = 0;
    for (n = 0; n < maxSV; n += 2) dst[(n / 2) + 1] = (uint8_t)((w[n] << 4) + w[n + 1]);
    return ((maxSV + 1) / 2) + 1;
}

/* one Huffman stream: symbols appended last-to-first, then the end mark
 * (HUF_compress1X_usingCTable_internal_body :1055-1118; the unrolled two-container
 * loop emits exactly this bit sequence) */
static size_t huf_encode_1x(uint8_t* dst, size_t cap, const uint8_t* src, size_t n, const zso_hufC* ct) {
    zso_bw w; size_t i;
    if (cap <= 8) return 0;
    bw_init(&w, dst, cap);
    for (i = n; i > 0; i--) bw_add(&w, ct->code[src[i - 1]], ct->nbBits[src[i - 1]]);
    return bw_close(&w);
}
/* HUF_compress4X_usingCTable_internal :1167-1215 */
static size_t huf_encode_4x(uint8_t* dst, size_t cap, const uint8_t* src, size_t n, const zso_hufC* ct) {
    size_t const seg = (n + 3) / 4; uint8_t* op = dst + 6; uint8_t* const oend = dst + cap; const uint8_t* ip = src; int k;
    if (cap < 6 + 1 + 1 + 1 + 8) return 0;
    if (n < 12) return 0;
    for (k = 0; k < 4; k++) {
        size_t const len = (k < 3) ? seg : (size_t)(src + n - ip);
        size_t const c = huf_encode_1x(op, (size_t)(oend - op), ip, len, ct);
        if (c == 0 || c > 65535) return 0;
        if (k < 3) zso_wr16(dst + 2 * k, (uint16_t)c);
        op += c; ip += len;
    }
    return (size_t)(op - dst);
}

/* HUF_compress_internal :1332-1434 for a first block (no previous table):
 * returns 0 = not compressible, 1 = single symbol (dst[0] = it), else compressed size */
static size_t huf_compress(uint8_t* dst, size_t cap, const uint8_t* src, size_t n, int fourStreams, int suspectUncompressible) {
    unsigned count[256]; unsigned maxSV = 255; zso_hufC ct; unsigned huffLog; uint8_t* op = dst; size_t hSize, c;
    if (!n || !cap) return 0;
    if (n > ZSO_BLOCKSIZE_MAX) return ZSO_ERROR(srcSize_wrong);
    if (suspectUncompressible && n >= 4096 * 10) {   /* :1367-1379 */
        unsigned m1 = 255, m2 = 255; size_t largestTotal = 0;
        largestTotal += hist(count, &m1, src, 4096);
        largestTotal += hist(count, &m2, src + n - 4096, 4096);
        if (largestTotal <= ((2 * 4096) >> 7) + 4) return 0;
    }
    {   unsigned const largest = hist(count, &maxSV, src, n);
        if (largest == n) { *dst = src[0]; return 1; }
        if (largest <= (n >> 7) + 4) return 0; }
    huffLog = fse_optimalTableLog(ZSO_LitHufLog, n, maxSV, 1);   /* HUF_optimalTableLog cheap path :1284-1287 */
    huffLog = huf_buildCTable(&ct, count, maxSV, huffLog);
    hSize = huf_writeCTable(op, cap, &ct);
    if (zso_isError(hSize)) return hSize;
    if (hSize + 12ul >= n) return 0;
    op += hSize;
    c = fourStreams ? huf_encode_4x(op, (size_t)(dst + cap - op), src, n, &ct) : huf_encode_1x(op, (size_t)(dst + cap - op), src, n, &ct);
    if (c == 0) return 0;
    op += c;
    if ((size_t)(op - dst) >= n - 1) return 0;   /* HUF_compressCTable_internal :1237 */
    return (size_t)(op - dst);
}

/* ZSTD_noCompressLiterals :39-66 / ZSTD_compressRleLiteralsBlock :81-107 */
static size_t lit_raw(uint8_t* dst, size_t cap, const uint8_t* src, size_t n) {
    uint32_t const fl = 1 + (n > 31) + (n > 4095);
    if (n + fl > cap) return ZSO_ERROR(dstSize_tooSmall);
    switch (fl) {
    case 1: dst[0] = (uint8_t)(0 + (n << 3)); break;
    case 2: zso_wr16(dst, (uint16_t)(0 + (1 << 2) + (n << 4))); break;
    default: zso_wr32(dst, (uint32_t)(0 + (3 << 2) + (n << 4))); break;
    }
    memcpy(dst + fl, src, n);
    return n + fl;
}
static size_t lit_rle(uint8_t* dst, const uint8_t* src, size_t n) {
    uint32_t const fl = 1 + (n > 31) + (n > 4095);
    switch (fl) {
    case 1: dst[0] = (uint8_t)(1 + (n << 3)); break;
    case 2: zso_wr16(dst, (uint16_t)(1 + (1 << 2) + (n << 4))); break;
    default: zso_wr32(dst, (uint32_t)(1 + (3 << 2) + (n << 4))); break;
    }
    dst[fl] = src[0];
    return fl + 1;
}

/* ZSTD_compressLiterals, N/compress/zstd_compress_literals.c:129-235, first block */
static size_t compress_literals(uint8_t* dst, size_t cap, const uint8_t* src, size_t n, unsigned strategy, int disableLiteralCompression, int suspectUncompressible) {
    size_t const lhSize = 3 + (n >= 1024) + (n >= 16384); int const single = n < 256; size_t cLit;
    if (disableLiteralCompression) return lit_raw(dst, cap, src, n);
    {   int const shift = (9 - (int)strategy) < 3 ? (9 - (int)strategy) : 3;
        if (n < ((size_t)8 << shift)) return lit_raw(dst, cap, src, n); }
    if (cap < lhSize + 1) return ZSO_ERROR(dstSize_tooSmall);
    cLit = huf_compress(dst + lhSize, cap - lhSize, src, n, !single, suspectUncompressible);
    {   size_t const minGain = (n >> (strategy >= ZSO_btultra ? strategy - 1 : 6)) + 2;
        if (cLit == 0 || zso_isError(cLit) || cLit >= n - minGain) return lit_raw(dst, cap, src, n); }
    if (cLit == 1) {
        int same = 1; size_t i; for (i = 1; i < n; i++) if (src[i] != src[0]) { same = 0; break; }
        if (n >= 8 || same) return lit_rle(dst, src, n);
    }
    switch (lhSize) {
    case 3: zso_wr24(dst, (uint32_t)(2 + ((uint32_t)(!single) << 2) + ((uint32_t)n << 4) + ((uint32_t)cLit << 14))); break;
    case 4: zso_wr32(dst, (uint32_t)(2 + (2 << 2) + ((uint32_t)n << 4) + ((uint32_t)cLit << 18))); break;
    default: zso_wr32(dst, (uint32_t)(2 + (3 << 2) + ((uint32_t)n << 4) + ((uint32_t)cLit << 22))); dst[4] = (uint8_t)(cLit >> 10); break;
    }
    return lhSize + cLit;
}

/* ------------------------------------------------------ sequences section */
static unsigned ll_code(uint32_t ll) {   /* ZSTD_LLcode, zstd_compress_internal.h:584-596 */
    static const uint8_t t[64] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 16, 17, 17, 18, 18, 19, 19,
        20, 20, 20, 20, 21, 21, 21, 21, 22, 22, 22, 22, 22, 22, 22, 22, 23, 23, 23, 23, 23, 23, 23, 23,
        24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24, 24 };
    return ll > 63 ? zso_highbit32(ll) + 19 : t[ll];
}
static unsigned ml_code(uint32_t mlBase) {   /* ZSTD_MLcode, :601-613 */
    if (mlBase > 127) return zso_highbit32(mlBase) + 36;
    if (mlBase < 32) return mlBase;
    {   static const uint8_t t[96] = { 32, 32, 33, 33, 34, 34, 35, 35, 36, 36, 36, 36, 37, 37, 37, 37,
            38, 38, 38, 38, 38, 38, 38, 38, 39, 39, 39, 39, 39, 39, 39, 39,
            40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41, 41,
            42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42, 42 };
        return t[mlBase - 32]; }
}

/* -log2(x/256) * 256, zstd_compress_sequences.c:21-44 */
static unsigned const k_invProbLog256[256] = {
    0,    2048, 1792, 1642, 1536, 1453, 1386, 1329, 1280, 1236, 1197, 1162, 1130, 1100, 1073, 1047, 1024, 1001, 980,  960,  941,  923,  906,  889,
    874,  859,  844,  830,  817,  804,  791,  779,  768,  756,  745,  734,  724,  714,  704,  694,  685,  676,  667,  658,  650,  642,  633,  626,
    618,  610,  603,  595,  588,  581,  574,  567,  561,  554,  548,  542,  535,  529,  523,  517,  512,  506,  500,  495,  489,  484,  478,  473,
    468,  463,  458,  453,  448,  443,  438,  434,  429,  424,  420,  415,  411,  407,  402,  398,  394,  390,  386,  382,  377,  373,  370,  366,
    362,  358,  354,  350,  347,  343,  339,  336,  332,  329,  325,  322,  318,  315,  311,  308,  305,  302,  298,  295,  292,  289,  286,  282,
    279,  276,  273,  270,  267,  264,  261,  258,  256,  253,  250,  247,  244,  241,  239,  236,  233,  230,  228,  225,  222,  220,  217,  215,
    212,  209,  207,  204,  202,  199,  197,  194,  192,  190,  187,  185,  182,  180,  178,  175,  173,  171,  168,  166,  164,  162,  159,  157,
    155,  153,  151,  149,  146,  144,  142,  140,  138,  136,  134,  132,  130,  128,  126,  123,  121,  119,  117,  115,  114,  112,  110,  108,
    106,  104,  102,  100,  98,   96,   94,   93,   91,   89,   87,   85,   83,   82,   80,   78,   76,   74,   73,   71,   69,   67,   66,   64,
    62,   61,   59,   57,   55,   54,   52,   50,   49,   47,   46,   44,   42,   41,   39,   37,   36,   34,   33,   31,   30,   28,   26,   25,
    23,   22,   20,   19,   17,   16,   14,   13,   11,   10,   8,    7,    5,    4,    2,    1 };
static size_t fse_normalizeCount(int16_t* norm, unsigned tableLog, const unsigned* count, size_t total, unsigned maxSV, unsigned useLowProbCount);
static size_t fse_writeNCount(uint8_t* dst, size_t cap, const int16_t* norm, unsigned maxSV, unsigned tableLog);
static unsigned fse_optimalTableLog(unsigned maxTableLog, size_t srcSize, unsigned maxSV, unsigned minus);

/* ZSTD_selectEncodingType for a first block (no repeat mode), zstd_compress_sequences.c:156-234:
 * heuristics for strategy < lazy, estimated bit costs (:205-231 with ZSTD_crossEntropyCost :140-154,
 * ZSTD_NCountCost :71-79, ZSTD_entropyCost :85-99) from lazy on */
static unsigned select_encoding(const unsigned* count, unsigned max, size_t mostFrequent, size_t nbSeq, unsigned FSELog,
                                const int16_t* defaultNorm, unsigned defaultNormLog, int defaultAllowed, unsigned strategy) {
    if (mostFrequent == nbSeq) return (defaultAllowed && nbSeq <= 2) ? 0 : 1;
    if (strategy < ZSO_lazy) {
        if (defaultAllowed) {
            size_t const mult = 10 - strategy;
            size_t const dynamicFse_nbSeq_min = (((size_t)1 << defaultNormLog) * mult) >> 3;
            if ((nbSeq < dynamicFse_nbSeq_min) || (mostFrequent < (nbSeq >> (defaultNormLog - 1)))) return 0;
        }
    } else {
        size_t basicCost = (size_t)-1, compressedCost; unsigned s;
        if (defaultAllowed) {
            unsigned const shift = 8 - defaultNormLog; size_t cost = 0;
            for (s = 0; s <= max; ++s) {
                unsigned const normAcc = (defaultNorm[s] != -1) ? (unsigned)defaultNorm[s] : 1;
                cost += count[s] * k_invProbLog256[normAcc << shift];
            }
            basicCost = cost >> 8;
        }
        {   uint8_t wksp[512]; int16_t norm[64]; size_t NCountCost;
            unsigned const tableLog = fse_optimalTableLog(FSELog, nbSeq, max, 2);
            size_t const r = fse_normalizeCount(norm, tableLog, count, nbSeq, max, nbSeq >= 2048);
            NCountCost = zso_isError(r) ? r : fse_writeNCount(wksp, sizeof(wksp), norm, max, tableLog);
            {   unsigned cost = 0;
                for (s = 0; s <= max; ++s) {
                    unsigned norm256 = (unsigned)((256 * count[s]) / nbSeq);
                    if (count[s] != 0 && norm256 == 0) norm256 = 1;
                    cost += count[s] * k_invProbLog256[norm256];
                }
                compressedCost = (NCountCost << 3) + (cost >> 8); }
        }
        if (basicCost <= compressedCost) return 0;      /* repeatCost is an error value in a first block */
    }
    return 2;
}

/* ZSTD_buildCTable, :242-288 */
static size_t build_ctable(uint8_t* dst, size_t cap, zso_ctable* ct, unsigned FSELog, unsigned type, unsigned* count, unsigned max,
                           const uint8_t* codes, size_t nbSeq, const int16_t* defaultNorm, unsigned defaultNormLog, unsigned defaultMax) {
    switch (type) {
    case 1: fse_buildCTable_rle(ct, max); if (cap == 0) return ZSO_ERROR(dstSize_tooSmall); dst[0] = codes[0]; return 1;
    case 0: fse_buildCTable(ct, defaultNorm, defaultMax, defaultNormLog); return 0;
    default: {
        int16_t norm[64]; size_t nbSeq_1 = nbSeq; unsigned const tableLog = fse_optimalTableLog(FSELog, nbSeq, max, 2); size_t r;
        if (count[codes[nbSeq - 1]] > 1) { count[codes[nbSeq - 1]]--; nbSeq_1--; }
        r = fse_normalizeCount(norm, tableLog, count, nbSeq_1, max, nbSeq_1 >= 2048);
        if (zso_isError(r)) return r;
        r = fse_writeNCount(dst, cap, norm, max, tableLog);
        if (zso_isError(r)) return r;
        fse_buildCTable(ct, norm, max, tableLog);
        return r; }
    }
}

/* ZSTD_entropyCompressSeqStore_internal :2887-3003 (+ buildSequencesStatistics :2762-2880,
 * seqToCodes :2693-2719, ZSTD_encodeSequences_body zstd_compress_sequences.c:290-382) */
static size_t entropy_compress(uint8_t* dst, size_t cap, const zso_seqStore* ss, unsigned strategy, int disableLiteralCompression) {
    uint8_t* op = dst; uint8_t* const oend = dst + cap; size_t const nbSeq = ss->nbSeq; size_t lastCountSize = 0;
    uint8_t* llc = NULL; uint8_t* ofc; uint8_t* mlc; zso_ctable ctLL, ctOF, ctML; size_t result;
    {   int const suspect = (nbSeq == 0) || (ss->litSize / nbSeq >= 20);
        size_t const c = compress_literals(op, cap, ss->lit, ss->litSize, strategy, disableLiteralCompression, suspect);
        if (zso_isError(c)) return c;
        op += c; }
    if ((oend - op) < 3 + 1) return ZSO_ERROR(dstSize_tooSmall);
    if (nbSeq < 128) *op++ = (uint8_t)nbSeq;
    else if (nbSeq < ZSO_LONGNBSEQ) { op[0] = (uint8_t)((nbSeq >> 8) + 0x80); op[1] = (uint8_t)nbSeq; op += 2; }
    else { op[0] = 0xFF; zso_wr16(op + 1, (uint16_t)(nbSeq - ZSO_LONGNBSEQ)); op += 3; }
    if (nbSeq == 0) return (size_t)(op - dst);
    llc = (uint8_t*)malloc(3 * nbSeq); if (!llc) return ZSO_ERROR(GENERIC);
    ofc = llc + nbSeq; mlc = ofc + nbSeq;
    {   size_t u; for (u = 0; u < nbSeq; u++) {
            llc[u] = (uint8_t)ll_code(ss->seq[u].litLength);
            ofc[u] = (uint8_t)zso_highbit32(ss->seq[u].offBase);
            mlc[u] = (uint8_t)ml_code(ss->seq[u].matchLength - ZSO_MINMATCH);
        } }
    {   uint8_t* const seqHead = op++; unsigned count[64]; unsigned LLtype, OFtype, MLtype; size_t c;
        {   unsigned max = ZSO_MaxLL; size_t const mf = hist(count, &max, llc, nbSeq);
            LLtype = select_encoding(count, max, mf, nbSeq, ZSO_LLFSELog, zso_LL_defaultNorm, 6, 1, strategy);
            c = build_ctable(op, (size_t)(oend - op), &ctLL, ZSO_LLFSELog, LLtype, count, max, llc, nbSeq, zso_LL_defaultNorm, 6, ZSO_MaxLL);
            if (zso_isError(c)) { result = c; goto done; }
            if (LLtype == 2) lastCountSize = c;
            op += c; }
        {   unsigned max = ZSO_MaxOff; size_t const mf = hist(count, &max, ofc, nbSeq);
            OFtype = select_encoding(count, max, mf, nbSeq, ZSO_OffFSELog, zso_OF_defaultNorm, 5, max <= ZSO_DefaultMaxOff, strategy);
            c = build_ctable(op, (size_t)(oend - op), &ctOF, ZSO_OffFSELog, OFtype, count, max, ofc, nbSeq, zso_OF_defaultNorm, 5, ZSO_DefaultMaxOff);
            if (zso_isError(c)) { result = c; goto done; }
            if (OFtype == 2) lastCountSize = c;
            op += c; }
        {   unsigned max = ZSO_MaxML; size_t const mf = hist(count, &max, mlc, nbSeq);
            MLtype = select_encoding(count, max, mf, nbSeq, ZSO_MLFSELog, zso_ML_defaultNorm, 6, 1, strategy);
            c = build_ctable(op, (size_t)(oend - op), &ctML, ZSO_MLFSELog, MLtype, count, max, mlc, nbSeq, zso_ML_defaultNorm, 6, ZSO_MaxML);
            if (zso_isError(c)) { result = c; goto done; }
            if (MLtype == 2) lastCountSize = c;
            op += c; }
        *seqHead = (uint8_t)((LLtype << 6) + (OFtype << 4) + (MLtype << 2));
    }
    {   zso_bw w; uint32_t sML, sOF, sLL; size_t n; size_t streamSize;
        if ((size_t)(oend - op) <= 8) { result = ZSO_ERROR(dstSize_tooSmall); goto done; }
        bw_init(&w, op, (size_t)(oend - op));
        sML = fse_init_state2(&ctML, mlc[nbSeq - 1]);
        sOF = fse_init_state2(&ctOF, ofc[nbSeq - 1]);
        sLL = fse_init_state2(&ctLL, llc[nbSeq - 1]);
        bw_add(&w, ss->seq[nbSeq - 1].litLength, zso_LL_bits[llc[nbSeq - 1]]);
        bw_add(&w, ss->seq[nbSeq - 1].matchLength - ZSO_MINMATCH, zso_ML_bits[mlc[nbSeq - 1]]);
        bw_add(&w, ss->seq[nbSeq - 1].offBase, ofc[nbSeq - 1]);
        for (n = nbSeq - 2; n < nbSeq; n--) {
            sOF = fse_encode(&w, &ctOF, sOF, ofc[n]);
            sML = fse_encode(&w, &ctML, sML, mlc[n]);
            sLL = fse_encode(&w, &ctLL, sLL, llc[n]);
            bw_add(&w, ss->seq[n].litLength, zso_LL_bits[llc[n]]);
            bw_add(&w, ss->seq[n].matchLength - ZSO_MINMATCH, zso_ML_bits[mlc[n]]);
            bw_add(&w, ss->seq[n].offBase, ofc[n]);
        }
        bw_add(&w, sML, ctML.tableLog); bw_add(&w, sOF, ctOF.tableLog); bw_add(&w, sLL, ctLL.tableLog);
        streamSize = bw_close(&w);
        if (streamSize == 0) { result = ZSO_ERROR(dstSize_tooSmall); goto done; }
        op += streamSize;
        if (lastCountSize && (lastCountSize + streamSize) < 4) { result = 0; goto done; }   /* 1.3.4 guard :2992-2998 */
    }
    result = (size_t)(op - dst);
done:
    free(llc);
    return result;
}

/* ZSTD_compressBlock_fast_noDict_generic lives in zso_fast.c (level 1/2) */
size_t zso_block_fast(zso_seqStore* ss, uint32_t rep[3], const uint8_t* src, size_t srcSize,
                      uint32_t* hashTable, unsigned hlog, unsigned mls, unsigned targetLength);

/* greedy / lazy / lazy2 live in zso_lazy.c: row-based match finder (levels 5..10, srcSize > 16 KB) or hash chain (levels 4..8, <= 16 KB) */
size_t zso_block_lazy(void* ss, uint32_t rep[3], const uint8_t* src, size_t srcSize,
                      uint32_t* hashTable, uint8_t* tagTable, uint32_t* chainTable, unsigned hashLog, unsigned chainLog, unsigned searchLog,
                      unsigned minMatch, unsigned depth, int binaryTree);

/* ---------------------------------------------------------------- frame */
uint64_t zso_xxh64(const void* data, size_t len, uint64_t seed);
size_t zso_compress(void* dstv, size_t dstCapacity, const void* srcv, size_t srcSize, int level) { return zso_compress_flags(dstv, dstCapacity, srcv, srcSize, level, 0); }
/* flags: 1 = ZSTD_c_checksumFlag on, 2 = ZSTD_c_contentSizeFlag off */
size_t zso_compress_flags(void* dstv, size_t dstCapacity, const void* srcv, size_t srcSize, int level, unsigned flags) {
    int const checksum = (flags & 1) != 0, contentSize = !(flags & 2);
    uint32_t const sum = checksum ? (uint32_t)zso_xxh64(srcv, srcSize, 0) : 0; size_t const sumBytes = checksum ? 4 : 0;
    uint8_t* const dst = (uint8_t*)dstv; const uint8_t* const src = (const uint8_t*)srcv; zso_cparams cp; size_t pos = 0;
    if (zso_getCParams(&cp, level, srcSize)) return ZSO_ERROR(parameter_unsupported);
    /* supported block compressors: fast, dfast, greedy/lazy/lazy2 (row match finder when windowLog > 14,
     * ZSTD_resolveRowMatchFinderMode zstd_compress.c:238-245, else hash chain); the binary-tree finders are not restated */
    if (cp.strategy > ZSO_btlazy2) return ZSO_ERROR(parameter_unsupported);      /* btopt and up: optimal parser, not restated */
    if (dstCapacity < 18) return ZSO_ERROR(dstSize_tooSmall);   /* ZSTD_FRAMEHEADERSIZE_MAX :4716 */
    /* ZSTD_writeFrameHeader :4695-4743 : no dictID; the pledged size is known, so windowSize >= srcSize => singleSegment
     * whenever the content size is written; otherwise the window descriptor byte appears instead */
    {   uint32_t const fcsCode = contentSize ? (srcSize >= 256) + (srcSize >= 65536 + 256) : 0;
        zso_wr32(dst, 0xFD2FB528u); pos = 4;
        dst[pos++] = (uint8_t)((checksum ? 4 : 0) + (contentSize ? (1 << 5) : 0) + (fcsCode << 6));
        if (!contentSize) dst[pos++] = (uint8_t)((cp.windowLog - 10) << 3);
        else switch (fcsCode) {
        case 0: dst[pos++] = (uint8_t)srcSize; break;
        case 1: zso_wr16(dst + pos, (uint16_t)(srcSize - 256)); pos += 2; break;
        default: zso_wr32(dst + pos, (uint32_t)srcSize); pos += 4; break;
        } }
    if (srcSize == 0) {   /* ZSTD_writeEpilogue :5364-5372: one empty raw last block */
        if (dstCapacity - pos < 3 + sumBytes) return ZSO_ERROR(dstSize_tooSmall);
        zso_wr24(dst + pos, 1); if (checksum) zso_wr32(dst + pos + 3, sum); return pos + 3 + sumBytes;
    }
    {   size_t cSize = 0; uint8_t* const op = dst + pos; size_t const cap = dstCapacity - pos;
        if (cap < 3 + 2 + 1) return ZSO_ERROR(dstSize_tooSmall);   /* :4624-4626 */
        if (srcSize >= 7) {   /* ZSTD_buildSeqStore :3273-3280 */
            zso_seqStore ss; uint32_t rep[3] = { 1, 4, 8 }; size_t lastLL;
            uint32_t* hashLong = (uint32_t*)calloc((size_t)1 << cp.hashLog, 4);
            uint32_t* hashSmall = (uint32_t*)calloc((size_t)1 << cp.chainLog, 4);
            ss.seq = (zso_seq*)malloc(sizeof(zso_seq) * (srcSize / 3 + 1)); ss.nbSeq = 0;
            ss.lit = (uint8_t*)malloc(srcSize + 8); ss.litSize = 0;
            if (!hashLong || !hashSmall || !ss.seq || !ss.lit) { free(hashLong); free(hashSmall); free(ss.seq); free(ss.lit); return ZSO_ERROR(GENERIC); }
            if (cp.strategy >= ZSO_greedy) {
                int const bt = cp.strategy == ZSO_btlazy2;
                int const useRow = !bt && cp.windowLog > 14;
                uint8_t* const tagTable = useRow ? (uint8_t*)calloc((size_t)1 << cp.hashLog, 1) : NULL;
                if (useRow && !tagTable) { free(hashLong); free(hashSmall); free(ss.seq); free(ss.lit); return ZSO_ERROR(GENERIC); }
                lastLL = zso_block_lazy(&ss, rep, src, srcSize, hashLong, tagTable, hashSmall /* chain table, 1 << chainLog */, cp.hashLog, cp.chainLog,
                                        cp.searchLog, cp.minMatch, bt ? 2 : cp.strategy - ZSO_greedy, bt);
                free(tagTable);
            }
            else if (cp.strategy == ZSO_dfast) lastLL = block_dfast(&ss, rep, src, srcSize, hashLong, cp.hashLog, hashSmall, cp.chainLog, cp.minMatch);
            else lastLL = zso_block_fast(&ss, rep, src, srcSize, hashLong, cp.hashLog, cp.minMatch, cp.targetLength);
            memcpy(ss.lit + ss.litSize, src + srcSize - lastLL, lastLL); ss.litSize += lastLL;   /* ZSTD_storeLastLiterals */
            /* ZSTD_literalsCompressionIsDisabled (zstd_compress_internal.h:685-700): "auto" disables
             * Huffman for ZSTD_fast with targetLength > 0, i.e. the negative levels */
            cSize = entropy_compress(op + 3, cap - 3, &ss, cp.strategy, cp.strategy == ZSO_fast && cp.targetLength > 0);
            free(hashLong); free(hashSmall); free(ss.seq); free(ss.lit);
            /* ZSTD_entropyCompressSeqStore_wExtLitBuffer :3005-3042 */
            if (cSize == ZSO_ERROR(dstSize_tooSmall) && srcSize <= cap - 3) cSize = 0;
            if (zso_isError(cSize)) return cSize;
            if (cSize) { size_t const maxCSize = srcSize - ((srcSize >> 6) + 2); if (cSize >= maxCSize) cSize = 0; }
            /* first block is never turned into an RLE block (:4423-4434) */
        }
        if (cSize == 0) {   /* ZSTD_noCompressBlock */
            if (srcSize + 3 > cap) return ZSO_ERROR(dstSize_tooSmall);
            zso_wr24(op, (uint32_t)(1 + (0 << 1) + (srcSize << 3)));
            memcpy(op + 3, src, srcSize);
            cSize = srcSize;
        } else zso_wr24(op, (uint32_t)(1 + (2 << 1) + (cSize << 3)));
        if (checksum) {   /* ZSTD_writeEpilogue :5373-5378 */
            if (cap - (3 + cSize) < 4) return ZSO_ERROR(dstSize_tooSmall);
            zso_wr32(op + 3 + cSize, sum);
        }
        return pos + 3 + cSize + sumBytes;
    }
}
