/*
 * zso_fast.c -- CPU oracle: the single-table "fast" block parser (levels 1-2
 * and negative levels at <= 128 KB).  TEST INFRASTRUCTURE ONLY.
 *
 * Restates ZSTD_compressBlock_fast_noDict_generic,
 * N/compress/zstd_fast.c:190-423 (N/ = src/main/native/).  The reference
 * software-pipelines four positions (ip0..ip3); the decisions are restated in
 * the same order because the order of table writes vs. reads is observable.
 * Index convention as in zso_compress.c: index = position + 2, zero = empty.
 */
#include "zso_common.h"

typedef struct { uint32_t litLength, offBase, matchLength; } zso_seq;
typedef struct { zso_seq* seq; size_t nbSeq; uint8_t* lit; size_t litSize; } zso_seqStore;

static void store_seq(zso_seqStore* ss, const uint8_t* literals, size_t litLength, uint32_t offBase, size_t matchLength) {
    memcpy(ss->lit + ss->litSize, literals, litLength); ss->litSize += litLength;
    ss->seq[ss->nbSeq].litLength = (uint32_t)litLength;
    ss->seq[ss->nbSeq].offBase = offBase;
    ss->seq[ss->nbSeq].matchLength = (uint32_t)matchLength;
    ss->nbSeq++;
}
static size_t hash_ptr(const uint8_t* p, unsigned hBits, unsigned mls) {
    switch (mls) {
    default:
    case 4: return (size_t)((zso_rd32(p) * 2654435761U) >> (32 - hBits));
    case 5: return (size_t)(((zso_rd64(p) << 24) * 889523592379ULL) >> (64 - hBits));
    case 6: return (size_t)(((zso_rd64(p) << 16) * 227718039650203ULL) >> (64 - hBits));
    case 7: return (size_t)(((zso_rd64(p) << 8) * 58295818150454627ULL) >> (64 - hBits));
    }
}
static size_t count_match(const uint8_t* in, const uint8_t* match, const uint8_t* end) {
    const uint8_t* const s = in;
    while (in < end && *in == *match) { in++; match++; }
    return (size_t)(in - s);
}
/* ZSTD_match4Found_{cmov,branch} :102-141 : same predicate either way */
static int match4(const uint8_t* cur, const uint8_t* base, uint32_t idx, uint32_t low) {
    return idx >= low && zso_rd32(cur) == zso_rd32(base + idx);
}

size_t zso_block_fast(zso_seqStore* ss, uint32_t rep[3], const uint8_t* src, size_t srcSize,
                      uint32_t* hashTable, unsigned hlog, unsigned mls, unsigned targetLength) {
    size_t const stepSize = targetLength + !targetLength + 1;
    const uint8_t* const base = src - 2;
    uint32_t const prefixStartIndex = 2;
    const uint8_t* const prefixStart = src;
    const uint8_t* const iend = src + srcSize;
    const uint8_t* const ilimit = iend - 8;
    const uint8_t* anchor = src; const uint8_t* ip0 = src; const uint8_t* ip1; const uint8_t* ip2; const uint8_t* ip3;
    uint32_t current0 = 0, rep1 = rep[0], rep2 = rep[1], saved1 = 0, saved2 = 0;
    size_t hash0, hash1, step, mLength; uint32_t matchIdx, offcode; const uint8_t* match0; const uint8_t* nextStep;
    size_t const kStepIncr = 1 << (8 - 1);

    ip0 += (ip0 == prefixStart);
    {   uint32_t const maxRep = (uint32_t)(ip0 - prefixStart);
        if (rep2 > maxRep) { saved2 = rep2; rep2 = 0; }
        if (rep1 > maxRep) { saved1 = rep1; rep1 = 0; } }
start:
    step = stepSize; nextStep = ip0 + kStepIncr;
    ip1 = ip0 + 1; ip2 = ip0 + step; ip3 = ip2 + 1;
    if (ip3 >= ilimit) goto cleanup;
    hash0 = hash_ptr(ip0, hlog, mls); hash1 = hash_ptr(ip1, hlog, mls);
    matchIdx = hashTable[hash0];
    do {
        uint32_t const rval = zso_rd32(ip2 - rep1);
        current0 = (uint32_t)(ip0 - base);
        hashTable[hash0] = current0;
        if ((zso_rd32(ip2) == rval) & (rep1 > 0)) {            /* repcode at ip2 (:275-290) */
            ip0 = ip2; match0 = ip0 - rep1;
            mLength = ip0[-1] == match0[-1];
            ip0 -= mLength; match0 -= mLength;
            offcode = 1; mLength += 4;
            hashTable[hash1] = (uint32_t)(ip1 - base);
            goto match;
        }
        if (match4(ip0, base, matchIdx, prefixStartIndex)) {   /* :292-299 */
            hashTable[hash1] = (uint32_t)(ip1 - base);
            goto offset;
        }
        matchIdx = hashTable[hash1];
        hash0 = hash1; hash1 = hash_ptr(ip2, hlog, mls);
        ip0 = ip1; ip1 = ip2; ip2 = ip3;
        current0 = (uint32_t)(ip0 - base);
        hashTable[hash0] = current0;
        if (match4(ip0, base, matchIdx, prefixStartIndex)) {   /* :317-326 */
            if (step <= 4) hashTable[hash1] = (uint32_t)(ip1 - base);
            goto offset;
        }
        matchIdx = hashTable[hash1];
        hash0 = hash1; hash1 = hash_ptr(ip2, hlog, mls);
        ip0 = ip1; ip1 = ip2; ip2 = ip0 + step; ip3 = ip1 + step;
        if (ip2 >= nextStep) { step++; nextStep += kStepIncr; }
    } while (ip3 < ilimit);
cleanup:
    saved2 = ((saved1 != 0) && (rep1 != 0)) ? saved1 : saved2;
    rep[0] = rep1 ? rep1 : saved1;
    rep[1] = rep2 ? rep2 : saved2;
    return (size_t)(iend - anchor);
offset:
    match0 = base + matchIdx;
    rep2 = rep1; rep1 = (uint32_t)(ip0 - match0);
    offcode = rep1 + 3; mLength = 4;
    while (((ip0 > anchor) & (match0 > prefixStart)) && (ip0[-1] == match0[-1])) { ip0--; match0--; mLength++; }
match:
    mLength += count_match(ip0 + mLength, match0 + mLength, iend);
    store_seq(ss, anchor, (size_t)(ip0 - anchor), offcode, mLength);
    ip0 += mLength; anchor = ip0;
    if (ip0 <= ilimit) {
        hashTable[hash_ptr(base + current0 + 2, hlog, mls)] = current0 + 2;
        hashTable[hash_ptr(ip0 - 2, hlog, mls)] = (uint32_t)(ip0 - 2 - base);
        if (rep2 > 0) {
            while ((ip0 <= ilimit) && (zso_rd32(ip0) == zso_rd32(ip0 - rep2))) {
                size_t const rLength = count_match(ip0 + 4, ip0 + 4 - rep2, iend) + 4;
                { uint32_t const t = rep2; rep2 = rep1; rep1 = t; }
                hashTable[hash_ptr(ip0, hlog, mls)] = (uint32_t)(ip0 - base);
                ip0 += rLength;
                store_seq(ss, anchor, 0, 1, rLength);
                anchor = ip0;
            }
        }
    }
    goto start;
}
