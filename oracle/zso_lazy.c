/* zso_lazy.c -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * Plain-C restatement of the greedy / lazy / lazy2 block compressors with the row-based match finder, as libzstd
 * 1.5.7 runs them for a fresh frame without dictionary (levels 5..10 for 16 KB < srcSize <= 128 KB):
 *   ZSTD_compressBlock_lazy_generic      N/compress/zstd_lazy.c:1516-1779  (searchMethod = search_rowHash, noDict)
 *   ZSTD_RowFindBestMatch                N/compress/zstd_lazy.c:1141-1360
 *   ZSTD_row_update_internal(+Impl)      :885-943,  ZSTD_row_fillHashCache :837-857, ZSTD_row_nextCachedHash :865-878
 *   ZSTD_row_nextIndex :798-803, ZSTD_row_getMatchMask :1061-1121 (any of its SIMD/SWAR variants: same mask)
 * and with the hash-chain finder the reference uses when the window is <= 2^14 (inputs <= 16 KB, levels 4..8):
 *   ZSTD_HcFindBestMatch :667-733, ZSTD_insertAndFindFirstIndex_internal :632-657
 * and with the binary-tree finder of btlazy2 (levels 9..10 on inputs <= 16 KB, levels 11..12):
 *   ZSTD_BtFindBestMatch :399-408, ZSTD_updateDUBT :29-65, ZSTD_insertDUBT1 :74-163, ZSTD_DUBT_findBestMatch :243-395
 * Index convention as in the other parsers: index = position + 2, zeroed cells are "nothing".
 * The hash salt is 0: the salt is XORed before the shift, so it only permutes rows and tags and the emitted
 * sequences do not depend on it; tag rows start zeroed like a fresh context.
 */
#include <stdlib.h>
#include <string.h>
#include "zso_common.h"

typedef struct { uint32_t litLength, offBase, matchLength; } zso_seq;
typedef struct { zso_seq* seq; size_t nbSeq; uint8_t* lit; size_t litSize; } zso_seqStore;

#define ROW_TAG_BITS 8
#define ROW_CACHE 8
#define K_SEARCH_STRENGTH 8
#define K_LAZY_SKIPPING_STEP 8

typedef struct {
    uint32_t* hashTable; uint8_t* tagTable;
    uint32_t hashCache[ROW_CACHE];
    uint32_t rowHashLog, rowLog, searchLog, mls;
    uint32_t nextToUpdate; int lazySkipping;
    const uint8_t* base;
    int useRow; uint32_t* chainTable; uint32_t hashLog, chainLog;      /* hash-chain finder */
} row_state;

static void store_seq(zso_seqStore* ss, const uint8_t* literals, size_t litLength, uint32_t offBase, size_t matchLength) {
    memcpy(ss->lit + ss->litSize, literals, litLength); ss->litSize += litLength;
    ss->seq[ss->nbSeq].litLength = (uint32_t)litLength; ss->seq[ss->nbSeq].offBase = offBase; ss->seq[ss->nbSeq].matchLength = (uint32_t)matchLength;
    ss->nbSeq++;
}
static size_t count_match(const uint8_t* in, const uint8_t* match, const uint8_t* end) {
    const uint8_t* const s = in;
    while (in + 8 <= end) { uint64_t const d = zso_rd64(in) ^ zso_rd64(match); if (d) return (size_t)(in - s) + ((unsigned)__builtin_ctzll(d) >> 3); in += 8; match += 8; }
    while (in < end && *in == *match) { in++; match++; }
    return (size_t)(in - s);
}
/* ZSTD_hashPtrSalted with salt 0, zstd_compress_internal.h:898-962 */
static uint32_t row_hash(const uint8_t* p, uint32_t hBits, uint32_t mls) {
    switch (mls) {
    default:
    case 4: return (zso_rd32(p) * 2654435761U) >> (32 - hBits);
    case 5: return (uint32_t)(((zso_rd64(p) << 24) * 889523592379ULL) >> (64 - hBits));
    case 6: return (uint32_t)(((zso_rd64(p) << 16) * 227718039650203ULL) >> (64 - hBits));
    }
}
static uint32_t row_next_index(uint8_t* tagRow, uint32_t rowMask) {      /* :798-803 */
    uint32_t next = ((uint32_t)*tagRow - 1) & rowMask;
    next += (next == 0) ? rowMask : 0;
    *tagRow = (uint8_t)next;
    return next;
}
static void row_fill_cache(row_state* ms, uint32_t idx, const uint8_t* iLimit) {   /* :837-857 */
    uint32_t const maxElems = (ms->base + idx) > iLimit ? 0 : (uint32_t)(iLimit - (ms->base + idx) + 1);
    uint32_t const lim = idx + (ROW_CACHE < maxElems ? ROW_CACHE : maxElems);
    for (; idx < lim; ++idx) ms->hashCache[idx & (ROW_CACHE - 1)] = row_hash(ms->base + idx, ms->rowHashLog + ROW_TAG_BITS, ms->mls);
}
static uint32_t row_next_cached(row_state* ms, uint32_t idx) {        /* :865-878 */
    uint32_t const newHash = row_hash(ms->base + idx + ROW_CACHE, ms->rowHashLog + ROW_TAG_BITS, ms->mls);
    uint32_t const hash = ms->hashCache[idx & (ROW_CACHE - 1)];
    ms->hashCache[idx & (ROW_CACHE - 1)] = newHash;
    return hash;
}
static void row_update_impl(row_state* ms, uint32_t idx, uint32_t end) {   /* :885-908, useCache = 1 */
    uint32_t const rowMask = (1u << ms->rowLog) - 1;
    for (; idx < end; ++idx) {
        uint32_t const hash = row_next_cached(ms, idx);
        uint32_t const relRow = (hash >> ROW_TAG_BITS) << ms->rowLog;
        uint8_t* const tagRow = ms->tagTable + relRow;
        uint32_t const pos = row_next_index(tagRow, rowMask);
        tagRow[pos] = (uint8_t)hash;
        ms->hashTable[relRow + pos] = idx;
    }
}
static void row_update(row_state* ms, const uint8_t* ip) {               /* :916-943 */
    uint32_t idx = ms->nextToUpdate;
    uint32_t const target = (uint32_t)(ip - ms->base);
    if (target - idx > 384) {
        row_update_impl(ms, idx, idx + 96);
        idx = target - 32;
        row_fill_cache(ms, idx, ip + 1);
    }
    row_update_impl(ms, idx, target);
    ms->nextToUpdate = target;
}
/* ZSTD_RowFindBestMatch :1141-1283 (noDict).  Returns the best length (3 = nothing) and its offBase. */
static size_t row_find_best(row_state* ms, const uint8_t* ip, const uint8_t* iLimit, size_t* offBasePtr) {
    uint32_t const curr = (uint32_t)(ip - ms->base);
    uint32_t const lowLimit = 2;                         /* window.lowLimit of a fresh frame, input <= window */
    uint32_t const rowEntries = 1u << ms->rowLog, rowMask = rowEntries - 1;
    uint32_t const cappedSearchLog = ms->searchLog < ms->rowLog ? ms->searchLog : ms->rowLog;
    uint32_t nbAttempts = 1u << cappedSearchLog;
    size_t ml = 4 - 1;
    uint32_t hash;
    if (!ms->lazySkipping) { row_update(ms, ip); hash = row_next_cached(ms, curr); }
    else { hash = row_hash(ip, ms->rowHashLog + ROW_TAG_BITS, ms->mls); ms->nextToUpdate = curr; }
    {   uint32_t const relRow = (hash >> ROW_TAG_BITS) << ms->rowLog;
        uint32_t const tag = hash & 0xFF;
        uint32_t* const row = ms->hashTable + relRow;
        uint8_t* const tagRow = ms->tagTable + relRow;
        uint32_t const head = *tagRow & rowMask;
        uint32_t matchBuffer[64]; size_t numMatches = 0, currMatch;
        uint32_t k;
        /* the match mask rotated right by head, walked from bit 0: positions head, head+1, ... (mod rowEntries) */
        for (k = 0; k < rowEntries && nbAttempts > 0; k++) {
            uint32_t const matchPos = (head + k) & rowMask;
            uint32_t matchIndex;
            if (tagRow[matchPos] != (uint8_t)tag) continue;
            matchIndex = row[matchPos];
            if (matchPos == 0) continue;
            if (matchIndex < lowLimit) break;
            matchBuffer[numMatches++] = matchIndex;
            --nbAttempts;
        }
        {   uint32_t const pos = row_next_index(tagRow, rowMask);
            tagRow[pos] = (uint8_t)tag;
            row[pos] = ms->nextToUpdate++;
        }
        for (currMatch = 0; currMatch < numMatches; ++currMatch) {
            const uint8_t* const match = ms->base + matchBuffer[currMatch];
            size_t currentMl = 0;
            if (zso_rd32(match + ml - 3) == zso_rd32(ip + ml - 3)) currentMl = count_match(ip, match, iLimit);
            if (currentMl > ml) {
                ml = currentMl;
                *offBasePtr = (size_t)(curr - matchBuffer[currMatch]) + 3;     /* OFFSET_TO_OFFBASE */
                if (ip + currentMl == iLimit) break;
            }
        }
    }
    return ml;
}

/* ZSTD_HcFindBestMatch :667-733 with ZSTD_insertAndFindFirstIndex_internal :632-657 (noDict) */
static uint32_t hc_hash(const uint8_t* p, uint32_t hBits, uint32_t mls) { return row_hash(p, hBits, mls); }   /* ZSTD_hashPtr, same multipliers */
static size_t hc_find_best(row_state* ms, const uint8_t* ip, const uint8_t* iLimit, size_t* offBasePtr) {
    uint32_t const chainSize = 1u << ms->chainLog, chainMask = chainSize - 1;
    uint32_t const curr = (uint32_t)(ip - ms->base);
    uint32_t const lowLimit = 2;
    uint32_t const minChain = curr > chainSize ? curr - chainSize : 0;
    uint32_t nbAttempts = 1u << ms->searchLog;
    size_t ml = 4 - 1;
    uint32_t matchIndex;
    {   uint32_t idx = ms->nextToUpdate;
        while (idx < curr) {
            uint32_t const h = hc_hash(ms->base + idx, ms->hashLog, ms->mls);
            ms->chainTable[idx & chainMask] = ms->hashTable[h];
            ms->hashTable[h] = idx;
            idx++;
            if (ms->lazySkipping) break;
        }
        ms->nextToUpdate = curr;
        matchIndex = ms->hashTable[hc_hash(ip, ms->hashLog, ms->mls)]; }
    for (; (matchIndex >= lowLimit) & (nbAttempts > 0); nbAttempts--) {
        const uint8_t* const match = ms->base + matchIndex;
        size_t currentMl = 0;
        if (zso_rd32(match + ml - 3) == zso_rd32(ip + ml - 3)) currentMl = count_match(ip, match, iLimit);
        if (currentMl > ml) {
            ml = currentMl;
            *offBasePtr = (size_t)(curr - matchIndex) + 3;
            if (ip + currentMl == iLimit) break;
        }
        if (matchIndex <= minChain) break;
        matchIndex = ms->chainTable[matchIndex & chainMask];
    }
    return ml;
}
/* ---- binary tree of the "dual unsorted" kind (DUBT); bt = chainTable as pairs {smaller, larger}, btLog = chainLog - 1 */
#define DUBT_UNSORTED_MARK 1
static void dubt_update(row_state* ms, const uint8_t* ip) {            /* ZSTD_updateDUBT :29-65 */
    uint32_t const btMask = (1u << (ms->chainLog - 1)) - 1;
    uint32_t const target = (uint32_t)(ip - ms->base);
    uint32_t idx = ms->nextToUpdate;
    for (; idx < target; idx++) {
        uint32_t const h = hc_hash(ms->base + idx, ms->hashLog, ms->mls);
        uint32_t const matchIndex = ms->hashTable[h];
        uint32_t* const nextCandidatePtr = ms->chainTable + 2 * (idx & btMask);
        ms->hashTable[h] = idx;
        nextCandidatePtr[0] = matchIndex;
        nextCandidatePtr[1] = DUBT_UNSORTED_MARK;
    }
    ms->nextToUpdate = target;
}
static void dubt_insert1(row_state* ms, uint32_t curr, const uint8_t* iend, uint32_t nbCompares, uint32_t btLow) {   /* ZSTD_insertDUBT1 :74-163 */
    uint32_t* const bt = ms->chainTable;
    uint32_t const btMask = (1u << (ms->chainLog - 1)) - 1;
    size_t commonLengthSmaller = 0, commonLengthLarger = 0;
    const uint8_t* const ip = ms->base + curr;
    uint32_t* smallerPtr = bt + 2 * (curr & btMask);
    uint32_t* largerPtr = smallerPtr + 1;
    uint32_t matchIndex = *smallerPtr;
    uint32_t dummy32;
    uint32_t const windowLow = 2;
    for (; nbCompares && (matchIndex > windowLow); --nbCompares) {
        uint32_t* const nextPtr = bt + 2 * (matchIndex & btMask);
        size_t matchLength = commonLengthSmaller < commonLengthLarger ? commonLengthSmaller : commonLengthLarger;
        const uint8_t* const match = ms->base + matchIndex;
        matchLength += count_match(ip + matchLength, match + matchLength, iend);
        if (ip + matchLength == iend) break;
        if (match[matchLength] < ip[matchLength]) {
            *smallerPtr = matchIndex;
            commonLengthSmaller = matchLength;
            if (matchIndex <= btLow) { smallerPtr = &dummy32; break; }
            smallerPtr = nextPtr + 1;
            matchIndex = nextPtr[1];
        } else {
            *largerPtr = matchIndex;
            commonLengthLarger = matchLength;
            if (matchIndex <= btLow) { largerPtr = &dummy32; break; }
            largerPtr = nextPtr;
            matchIndex = nextPtr[0];
        }
    }
    *smallerPtr = *largerPtr = 0;
}
static size_t bt_find_best(row_state* ms, const uint8_t* ip, const uint8_t* iend, size_t* offBasePtr) {   /* :399-408 + :243-395 */
    uint32_t* const bt = ms->chainTable;
    uint32_t const btMask = (1u << (ms->chainLog - 1)) - 1;
    uint32_t const curr = (uint32_t)(ip - ms->base);
    uint32_t const windowLow = 2;
    uint32_t const btLow = (btMask >= curr) ? 0 : curr - btMask;
    uint32_t const unsortLimit = btLow > windowLow ? btLow : windowLow;
    uint32_t h, matchIndex, nbCompares = 1u << ms->searchLog, nbCandidates = nbCompares, previousCandidate = 0;
    uint32_t* nextCandidate; uint32_t* unsortedMark;
    if (ip < ms->base + ms->nextToUpdate) return 0;      /* skipped area */
    dubt_update(ms, ip);
    h = hc_hash(ip, ms->hashLog, ms->mls);
    matchIndex = ms->hashTable[h];
    nextCandidate = bt + 2 * (matchIndex & btMask); unsortedMark = nextCandidate + 1;
    while ((matchIndex > unsortLimit) && (*unsortedMark == DUBT_UNSORTED_MARK) && (nbCandidates > 1)) {
        *unsortedMark = previousCandidate;
        previousCandidate = matchIndex;
        matchIndex = *nextCandidate;
        nextCandidate = bt + 2 * (matchIndex & btMask); unsortedMark = nextCandidate + 1;
        nbCandidates--;
    }
    if ((matchIndex > unsortLimit) && (*unsortedMark == DUBT_UNSORTED_MARK)) *nextCandidate = *unsortedMark = 0;
    matchIndex = previousCandidate;
    while (matchIndex) {
        uint32_t* const nextCandidateIdxPtr = bt + 2 * (matchIndex & btMask) + 1;
        uint32_t const nextCandidateIdx = *nextCandidateIdxPtr;
        dubt_insert1(ms, matchIndex, iend, nbCandidates, unsortLimit);
        matchIndex = nextCandidateIdx;
        nbCandidates++;
    }
    {   size_t commonLengthSmaller = 0, commonLengthLarger = 0, bestLength = 0;
        uint32_t* smallerPtr = bt + 2 * (curr & btMask);
        uint32_t* largerPtr = smallerPtr + 1;
        uint32_t matchEndIdx = curr + 8 + 1;
        uint32_t dummy32;
        matchIndex = ms->hashTable[h];
        ms->hashTable[h] = curr;
        for (; nbCompares && (matchIndex > windowLow); --nbCompares) {
            uint32_t* const nextPtr = bt + 2 * (matchIndex & btMask);
            size_t matchLength = commonLengthSmaller < commonLengthLarger ? commonLengthSmaller : commonLengthLarger;
            const uint8_t* const match = ms->base + matchIndex;
            matchLength += count_match(ip + matchLength, match + matchLength, iend);
            if (matchLength > bestLength) {
                if (matchLength > matchEndIdx - matchIndex) matchEndIdx = matchIndex + (uint32_t)matchLength;
                if ((4 * (int)(matchLength - bestLength)) > (int)(zso_highbit32(curr - matchIndex + 1) - zso_highbit32((uint32_t)*offBasePtr))) {
                    bestLength = matchLength; *offBasePtr = (size_t)(curr - matchIndex) + 3; }
                if (ip + matchLength == iend) break;
            }
            if (match[matchLength] < ip[matchLength]) {
                *smallerPtr = matchIndex;
                commonLengthSmaller = matchLength;
                if (matchIndex <= btLow) { smallerPtr = &dummy32; break; }
                smallerPtr = nextPtr + 1;
                matchIndex = nextPtr[1];
            } else {
                *largerPtr = matchIndex;
                commonLengthLarger = matchLength;
                if (matchIndex <= btLow) { largerPtr = &dummy32; break; }
                largerPtr = nextPtr;
                matchIndex = nextPtr[0];
            }
        }
        *smallerPtr = *largerPtr = 0;
        ms->nextToUpdate = matchEndIdx - 8;
        return bestLength;
    }
}
static size_t find_best(row_state* ms, const uint8_t* ip, const uint8_t* iLimit, size_t* offBasePtr) {
    if (ms->useRow == 2) return bt_find_best(ms, ip, iLimit, offBasePtr);
    return ms->useRow ? row_find_best(ms, ip, iLimit, offBasePtr) : hc_find_best(ms, ip, iLimit, offBasePtr);
}

/* ZSTD_compressBlock_lazy_generic :1516-1779; depth 0 = greedy, 1 = lazy, 2 = lazy2.  Returns the trailing literal run. */
size_t zso_block_lazy(void* ssv, uint32_t rep[3], const uint8_t* src, size_t srcSize,
                      uint32_t* hashTable, uint8_t* tagTable, uint32_t* chainTable, unsigned hashLog, unsigned chainLog, unsigned searchLog,
                      unsigned minMatch, unsigned depth, int binaryTree) {     /* tagTable != NULL: row finder; else hash chain or binary tree */
    zso_seqStore* const ss = (zso_seqStore*)ssv;
    const uint8_t* const istart = src;
    const uint8_t* ip = istart;
    const uint8_t* anchor = istart;
    const uint8_t* const iend = istart + srcSize;
    int const useRow = tagTable != NULL;
    const uint8_t* const ilimit = useRow ? iend - 8 - ROW_CACHE : iend - 8;
    const uint8_t* const prefixLowest = src;
    uint32_t offset_1 = rep[0], offset_2 = rep[1], offsetSaved1 = 0, offsetSaved2 = 0;
    row_state ms;
    ms.hashTable = hashTable; ms.tagTable = tagTable; ms.base = src - 2;
    ms.mls = minMatch < 4 ? 4 : minMatch > 6 ? 6 : minMatch;
    ms.rowLog = searchLog < 4 ? 4 : searchLog > 6 ? 6 : searchLog;
    ms.searchLog = searchLog; ms.rowHashLog = hashLog - ms.rowLog;
    ms.nextToUpdate = 2; ms.lazySkipping = 0;
    ms.useRow = binaryTree ? 2 : useRow; ms.chainTable = chainTable; ms.hashLog = hashLog; ms.chainLog = chainLog;

    ip += 1;                                            /* dictAndPrefixLength == 0 */
    {   uint32_t const maxRep = (uint32_t)(ip - prefixLowest);
        if (offset_2 > maxRep) { offsetSaved2 = offset_2; offset_2 = 0; }
        if (offset_1 > maxRep) { offsetSaved1 = offset_1; offset_1 = 0; }
    }
    if (useRow) row_fill_cache(&ms, ms.nextToUpdate, ilimit);

    while (ip < ilimit) {
        size_t matchLength = 0;
        size_t offBase = 1;                             /* REPCODE1_TO_OFFBASE */
        const uint8_t* start = ip + 1;
        if ((offset_1 > 0) & (zso_rd32(ip + 1 - offset_1) == zso_rd32(ip + 1))) {
            matchLength = count_match(ip + 1 + 4, ip + 1 + 4 - offset_1, iend) + 4;
            if (depth == 0) goto _storeSequence;
        }
        {   size_t offbaseFound = 999999999;
            size_t const ml2 = find_best(&ms, ip, iend, &offbaseFound);
            if (ml2 > matchLength) { matchLength = ml2; start = ip; offBase = offbaseFound; }
        }
        if (matchLength < 4) {
            size_t const step = ((size_t)(ip - anchor) >> K_SEARCH_STRENGTH) + 1;
            ip += step;
            ms.lazySkipping = step > K_LAZY_SKIPPING_STEP;
            continue;
        }
        if (depth >= 1)
        while (ip < ilimit) {
            ip++;
            if ((offBase) && ((offset_1 > 0) & (zso_rd32(ip) == zso_rd32(ip - offset_1)))) {
                size_t const mlRep = count_match(ip + 4, ip + 4 - offset_1, iend) + 4;
                int const gain2 = (int)(mlRep * 3);
                int const gain1 = (int)(matchLength * 3 - zso_highbit32((uint32_t)offBase) + 1);
                if ((mlRep >= 4) && (gain2 > gain1)) { matchLength = mlRep; offBase = 1; start = ip; }
            }
            {   size_t ofbCandidate = 999999999;
                size_t const ml2 = find_best(&ms, ip, iend, &ofbCandidate);
                int const gain2 = (int)(ml2 * 4 - zso_highbit32((uint32_t)ofbCandidate));
                int const gain1 = (int)(matchLength * 4 - zso_highbit32((uint32_t)offBase) + 4);
                if ((ml2 >= 4) && (gain2 > gain1)) { matchLength = ml2; offBase = ofbCandidate; start = ip; continue; }
            }
            if ((depth == 2) && (ip < ilimit)) {
                ip++;
                if ((offBase) && ((offset_1 > 0) & (zso_rd32(ip) == zso_rd32(ip - offset_1)))) {
                    size_t const mlRep = count_match(ip + 4, ip + 4 - offset_1, iend) + 4;
                    int const gain2 = (int)(mlRep * 4);
                    int const gain1 = (int)(matchLength * 4 - zso_highbit32((uint32_t)offBase) + 1);
                    if ((mlRep >= 4) && (gain2 > gain1)) { matchLength = mlRep; offBase = 1; start = ip; }
                }
                {   size_t ofbCandidate = 999999999;
                    size_t const ml2 = find_best(&ms, ip, iend, &ofbCandidate);
                    int const gain2 = (int)(ml2 * 4 - zso_highbit32((uint32_t)ofbCandidate));
                    int const gain1 = (int)(matchLength * 4 - zso_highbit32((uint32_t)offBase) + 7);
                    if ((ml2 >= 4) && (gain2 > gain1)) { matchLength = ml2; offBase = ofbCandidate; start = ip; continue; }
                }
            }
            break;
        }
        if (offBase > 3) {                               /* OFFBASE_IS_OFFSET: catch up */
            size_t const off = offBase - 3;
            while (((start > anchor) & (start - off > prefixLowest)) && (start[-1] == (start - off)[-1])) { start--; matchLength++; }
            offset_2 = offset_1; offset_1 = (uint32_t)off;
        }
_storeSequence:
        store_seq(ss, anchor, (size_t)(start - anchor), (uint32_t)offBase, matchLength);
        anchor = ip = start + matchLength;
        if (ms.lazySkipping) { if (useRow) row_fill_cache(&ms, ms.nextToUpdate, ilimit); ms.lazySkipping = 0; }
        while (((ip <= ilimit) & (offset_2 > 0)) && (zso_rd32(ip) == zso_rd32(ip - offset_2))) {
            uint32_t tmp;
            matchLength = count_match(ip + 4, ip + 4 - offset_2, iend) + 4;
            tmp = offset_2; offset_2 = offset_1; offset_1 = tmp;
            store_seq(ss, anchor, 0, 1, matchLength);
            ip += matchLength; anchor = ip;
        }
    }
    offsetSaved2 = ((offsetSaved1 != 0) && (offset_1 != 0)) ? offsetSaved1 : offsetSaved2;
    rep[0] = offset_1 ? offset_1 : offsetSaved1;
    rep[1] = offset_2 ? offset_2 : offsetSaved2;
    return (size_t)(iend - anchor);
}
