// zb_decode.cuh -- warp-cooperative Zstandard decoder (one warp owns one item = one or
// more concatenated frames; blocks inside a frame are decoded in order by that warp).
//
// What the reference does for this path (N/ = luben/zstd-jni src/main/native/):
//   frame layer      N/decompress/zstd_decompress.c:447-551 (header), :953-1066 (frame), :1070-1169 (multi-frame)
//   literals         N/decompress/zstd_decompress_block.c:134-340, N/decompress/huf_decompress.c:385-698
//   sequence tables  N/decompress/zstd_decompress_block.c:485-603,647-775, N/common/entropy_common.c:42-306
//   decode + execute N/decompress/zstd_decompress_block.c:1001-1096,1229-1346,1615-1690
//
// GPU mapping (W = 32 lanes):
//   * header / table parsing is scalar work done by lane 0 (results broadcast);
//   * the three FSE decode tables are built by lanes 0..2 concurrently, tables live in shared memory;
//   * the 4 Huffman streams of a literals section are decoded by lanes 0..3 concurrently;
//   * sequences are decoded by lane 0 in batches of 32 into shared memory, then the whole warp
//     executes them in order: every lane copies one byte per step, overlapping matches
//     (offset < length) are copied from the already-written period so that all lanes stay independent.
#pragma once
#include "zb_common.cuh"

namespace zb {

constexpr u32 SEQ_BATCH = 32;

struct DecShared {
    u32 fse[3][512];        // [0]=LL [1]=OF [2]=ML ; nextState | nbBits<<16 | symbol<<24
    u16 huf[1u << HUF_TABLELOG_MAX];   // symbol | nbBits<<8
    u32 sLit[SEQ_BATCH], sMatch[SEQ_BATCH], sOff[SEQ_BATCH];
    i16 norm[3][64];
    u16 symNext[3][64];
    u16 cum[66];            // cumulative cell counts of the table being built (build_fse_dtable_warp)
    u8 spread[3][512];
    u8 weights[256];
    u32 wdt[64];            // FSE decode table of the Huffman-weight stream (tableLog <= 6)
    i16 wnorm[256];
    u16 wnext[256];
    u32 fseLog[3];
    u32 fseMode[3];         // scratch between parse and build
    u32 fseMax[3];
    u32 hufLog;
    u32 rep[3];
    u32 litEntropy, fseEntropy;
    u32 tmp[8];
};

// ------------------------------------------------------------------ NCount
// FSE_readNCount_body, N/common/entropy_common.c:42-187, on a forward bit cursor.
ZB_HD u32 fwd_bits(const u8* p, size_t size, size_t bitpos, u32 n) {
    size_t const byte = bitpos >> 3;
    if (byte >= size) return 0;
    u32 const sh = (u32)(bitpos & 7);
    size_t const availBytes = size - byte;
    u32 need = (sh + n + 7) >> 3;
    if (need > availBytes) need = (u32)availBytes;
    u64 w = load64_n(p + byte, need);
    if (need < 8) w &= (1ull << (need * 8)) - 1;
    return (u32)((w >> sh) & ((1ull << n) - 1));
}

ZB_HDN size_t read_ncount(i16* norm, u32* maxSV, u32* tableLog, const u8* src, size_t srcSize) {
    size_t bitpos = 0;
    u32 const maxSV1 = *maxSV + 1;
    u32 charnum = 0;
    int nbBits, remaining, threshold, previous0 = 0;
    if (srcSize == 0) return ERR(E_srcSize_wrong);
    for (u32 s = 0; s < maxSV1; s++) norm[s] = 0;
    nbBits = (int)fwd_bits(src, srcSize, bitpos, 4) + 5; bitpos += 4;
    if (nbBits > 15) return ERR(E_tableLog_tooLarge);
    *tableLog = (u32)nbBits;
    remaining = (1 << nbBits) + 1;
    threshold = 1 << nbBits;
    nbBits++;
    for (;;) {
        if (previous0) {
            for (;;) {
                u32 const r = fwd_bits(src, srcSize, bitpos, 2); bitpos += 2;
                charnum += r;
                if (r != 3) break;
            }
            if (charnum >= maxSV1) break;
        }
        {   int const max = (2 * threshold - 1) - remaining;
            int count;
            u32 const bs = fwd_bits(src, srcSize, bitpos, (u32)nbBits);
            if ((int)(bs & (u32)(threshold - 1)) < max) { count = (int)(bs & (u32)(threshold - 1)); bitpos += (u32)nbBits - 1; }
            else { count = (int)(bs & (u32)(2 * threshold - 1)); if (count >= threshold) count -= max; bitpos += (u32)nbBits; }
            count--;
            if (count >= 0) remaining -= count; else remaining += count;
            norm[charnum++] = (i16)count;
            previous0 = !count;
            if (remaining < threshold) {
                if (remaining <= 1) break;
                nbBits = (int)highbit32((u32)remaining) + 1;
                threshold = 1 << (nbBits - 1);
            }
            if (charnum >= maxSV1) break;
        }
    }
    if (remaining != 1) return ERR(E_corruption_detected);
    if (charnum > maxSV1) return ERR(E_maxSymbolValue_tooSmall);
    *maxSV = charnum - 1;
    size_t const used = (bitpos + 7) >> 3;
    if (used > srcSize) return ERR(E_corruption_detected);
    return used;
}

// ------------------------------------------------------- FSE decode tables
// Symbol spreading + state numbering shared by the sequence tables
// (ZSTD_buildFSETable_body, zstd_decompress_block.c:485-603) and the Huffman-weight
// table (FSE_buildDTable_internal, N/common/fse_decompress.c:58-159).
// Writes packed entries nextState | nbBits<<16 | symbol<<24.  Returns false when the
// distribution does not tile the table (reference: ERROR(GENERIC) / assert).
ZB_HDN bool build_fse_dtable(u32* table, const i16* norm, u32 maxSV, u32 tableLog, u16* symNext, u8* spread) {
    u32 const tableSize = 1u << tableLog, mask = tableSize - 1;
    u32 const step = (tableSize >> 1) + (tableSize >> 3) + 3;
    u32 high = tableSize - 1, pos = 0;
    for (u32 s = 0; s <= maxSV; s++) {
        if (norm[s] == -1) { spread[high--] = (u8)s; symNext[s] = 1; }
        else symNext[s] = (u16)norm[s];
    }
    for (u32 s = 0; s <= maxSV; s++) {
        int const n = norm[s];
        for (int i = 0; i < n; i++) {
            spread[pos] = (u8)s;
            pos = (pos + step) & mask;
            while (pos > high) pos = (pos + step) & mask;
        }
    }
    if (pos != 0) return false;
    for (u32 u = 0; u < tableSize; u++) {
        u32 const s = spread[u];
        u32 const ns = symNext[s]++;
        u32 const nb = tableLog - highbit32(ns);
        table[u] = (((ns << nb) - tableSize) & 0xFFFFu) | (nb << 16) | (s << 24);
    }
    return true;
}

// The same table built by all lanes of the warp (sequence tables: maxSV <= 52, tableLog <= 9).  Nothing about the result
// changes; the three serial loops of the reference become data-parallel steps:
//   * the spread visits positions (j * step) & mask for j = 0, 1, 2 ... and skips the cells above `high` (those belong to the
//     low-probability symbols), so the k-th visited cell takes the symbol whose cumulative count range holds k -- lanes take 32
//     consecutive j, rank themselves with a ballot and look the symbol up in the cumulative counts;
//   * the state numbering hands the cells of one symbol consecutive numbers in cell order -- lanes take 32 consecutive cells,
//     lanes with the same symbol rank themselves with match_any, the first of them advances the symbol's counter.
template <class C>
ZB_HDN bool build_fse_dtable_warp(const C& w, u32* table, const i16* norm, u32 maxSV, u32 tableLog, u16* symNext, u8* spread, u16* cum) {
    u32 const tableSize = 1u << tableLog, mask = tableSize - 1;
    u32 const step = (tableSize >> 1) + (tableSize >> 3) + 3;
    u32 const lane = (u32)w.lane, below = (1u << lane) - 1;
    // low-probability symbols from the top down, cumulative counts of the others
    u32 nLow = 0, total = 0;
    for (u32 s0 = 0; s0 <= maxSV; s0 += C::W) {
        u32 const sy = s0 + lane;
        int const n = sy <= maxSV ? (int)norm[sy] : 0;
        u32 const lowM = w.ballot(n == -1);
        if (n == -1) { spread[tableSize - 1 - nLow - popc32(lowM & below)] = (u8)sy; symNext[sy] = 1; }
        else if (sy <= maxSV) symNext[sy] = (u16)n;
        u32 const cnt = n > 0 ? (u32)n : 0;
        u32 const before = w.exscan(cnt);
        if (sy <= maxSV) cum[sy] = (u16)(total + before);
        total += w.sum(cnt); nLow += popc32(lowM);
    }
    u32 const high = tableSize - 1 - nLow;
    for (u32 sy = maxSV + 1 + lane; sy < 66; sy += C::W) cum[sy] = 0xFFFF;      // sentinels for the search
    w.sync();
    if (total != high + 1) return false;                   // the distribution does not tile the table (reference: pos != 0 after the spread)
    u32 kBase = 0;
    for (u32 j0 = 0; j0 < tableSize; j0 += C::W) {
        u32 const j = j0 + lane;
        u32 const pos = (j * step) & mask;
        bool const valid = j < tableSize && pos <= high;
        u32 const vm = w.ballot(valid);
        if (valid) {
            u32 const k = kBase + popc32(vm & below);
            u32 sy = 0;                                        // last symbol with cum[sy] <= k (cum is non-decreasing; 64 entries + sentinels)
            sy += cum[sy + 32] <= k ? 32u : 0u; sy += cum[sy + 16] <= k ? 16u : 0u; sy += cum[sy + 8] <= k ? 8u : 0u;
            sy += cum[sy + 4] <= k ? 4u : 0u; sy += cum[sy + 2] <= k ? 2u : 0u; sy += cum[sy + 1] <= k ? 1u : 0u;
            spread[pos] = (u8)sy;
        }
        kBase += popc32(vm);
    }
    w.sync();
    for (u32 u0 = 0; u0 < tableSize; u0 += C::W) {
        u32 const u = u0 + lane;
        bool const on = u < tableSize;
        u32 const sy = on ? spread[u] : 0xFFFFFFFFu - lane;
        u32 const same = w.match_any(sy);
        u32 const base = on ? symNext[sy] : 0;
        w.sync();
        if (on) {
            u32 const ns = base + popc32(same & below);
            u32 const nb = tableLog - highbit32(ns);
            table[u] = (((ns << nb) - tableSize) & 0xFFFFu) | (nb << 16) | (sy << 24);
            if (!(same & below)) symNext[sy] = (u16)(base + popc32(same));
        }
        w.sync();
    }
    return true;
}

// ---------------------------------------------------------- Huffman table
// HUF_readStats_body (entropy_common.c:243-306) + HUF_readDTableX1_wksp (huf_decompress.c:385-519).
// Lane 0 parses the weights; the table fill is spread over the warp.
// Returns header bytes consumed, or an error.  S.huf / S.hufLog are the result.
template <class C>
ZB_HDN size_t huf_read_table(const C& w, DecShared& S, const u8* src, size_t srcSize) {
    size_t result = 0;
    if (w.lane == 0) {
        u8* const wt = S.weights;
        size_t iSize, oSize = 0;
        result = 0;
        do {
            if (!srcSize) { result = ERR(E_srcSize_wrong); break; }
            iSize = src[0];
            if (iSize >= 128) {
                oSize = iSize - 127; iSize = (oSize + 1) / 2;
                if (iSize + 1 > srcSize) { result = ERR(E_srcSize_wrong); break; }
                for (size_t n = 0; n < oSize; n += 2) { u8 const b = src[1 + n / 2]; wt[n] = b >> 4; wt[n + 1] = b & 15; }
            } else {
                if (iSize + 1 > srcSize) { result = ERR(E_srcSize_wrong); break; }
                // FSE_decompress_wksp_body, fse_decompress.c:243-289, two interleaved states :173-236
                i16* const norm = S.wnorm; u32 maxSV = 255; u32 tableLog;
                size_t const h = read_ncount(norm, &maxSV, &tableLog, src + 1, iSize);
                if (isErr(h)) { result = h; break; }
                if (tableLog > 6) { result = ERR(E_tableLog_tooLarge); break; }
                u32* const dt = S.wdt;
                if (!build_fse_dtable(dt, norm, maxSV, tableLog, S.wnext, S.spread[0])) { result = ERR(E_GENERIC); break; }
                const u8* const bs = src + 1 + h; size_t const bsSize = iSize - h;
                if (bsSize < 1) { result = ERR(E_srcSize_wrong); break; }
                if (bs[bsSize - 1] == 0) { result = ERR(E_GENERIC); break; }
                i64 pos = (i64)(bsSize - 1) * 8 + highbit32(bs[bsSize - 1]);
                pos -= tableLog; u32 s1 = (u32)peek_bits(bs, pos, tableLog);
                pos -= tableLog; u32 s2 = (u32)peek_bits(bs, pos, tableLog);
                if (pos < 0) { result = ERR(E_corruption_detected); break; }
                size_t n = 0; bool bad = false;
                for (;;) {
                    if (n + 2 > 255) { bad = true; break; }
                    { u32 const e = dt[s1]; wt[n++] = (u8)(e >> 24); u32 const nb = (e >> 16) & 0xFF; pos -= nb; s1 = (e & 0xFFFF) + (u32)peek_bits(bs, pos, nb); }
                    if (pos < 0) { wt[n++] = (u8)(dt[s2] >> 24); break; }
                    if (n + 2 > 255) { bad = true; break; }
                    { u32 const e = dt[s2]; wt[n++] = (u8)(e >> 24); u32 const nb = (e >> 16) & 0xFF; pos -= nb; s2 = (e & 0xFFFF) + (u32)peek_bits(bs, pos, nb); }
                    if (pos < 0) { wt[n++] = (u8)(dt[s1] >> 24); break; }
                }
                if (bad) { result = ERR(E_dstSize_tooSmall); break; }
                oSize = n;
            }
            // rank statistics and implied last weight (:276-301)
            u32* const rank = S.tmp;   // ranks 1..12 packed into S.norm[1] region below
            u16* const rk = S.symNext[1];
            for (u32 r = 0; r <= HUF_TABLELOG_MAX + 1; r++) rk[r] = 0;
            u32 total = 0; bool bad = false;
            for (size_t n = 0; n < oSize; n++) { if (wt[n] > HUF_TABLELOG_MAX) { bad = true; break; } rk[wt[n]]++; total += (1u << wt[n]) >> 1; }
            if (bad || total == 0) { result = ERR(E_corruption_detected); break; }
            u32 const tableLog = highbit32(total) + 1;
            if (tableLog > HUF_TABLELOG_MAX) { result = ERR(E_corruption_detected); break; }
            u32 const rest = (1u << tableLog) - total;
            if ((1u << highbit32(rest)) != rest) { result = ERR(E_corruption_detected); break; }
            u32 const last = highbit32(rest) + 1;
            wt[oSize] = (u8)last; rk[last]++;
            if (rk[1] < 2 || (rk[1] & 1)) { result = ERR(E_corruption_detected); break; }
            // start cell of each symbol: weights ascending, symbols ascending inside a weight
            u16* const start = S.symNext[2];        // per-rank running start
            u32 posc = 0;
            for (u32 r = 1; r <= tableLog; r++) { start[r] = (u16)posc; posc += (u32)rk[r] << (r - 1); }
            u16* const symStart = reinterpret_cast<u16*>(S.spread[1]);   // 256 x u16 = 512 B
            for (size_t s = 0; s <= oSize; s++) {
                u32 const ww = wt[s];
                if (ww) { symStart[s] = start[ww]; start[ww] = (u16)(start[ww] + (1u << (ww - 1))); } else symStart[s] = 0;
            }
            S.hufLog = tableLog;
            rank[0] = (u32)oSize + 1;   // number of symbols
            result = iSize + 1;
        } while (0);
    }
    w.sync();
    result = w.bcast(result);
    if (isErr(result)) return result;
    {   // cooperative fill
        u32 const nSym = S.tmp[0], tableLog = S.hufLog;
        const u16* const symStart = reinterpret_cast<const u16*>(S.spread[1]);
        for (u32 s = 0; s < nSym; s++) {
            u32 const ww = S.weights[s];
            if (!ww) continue;
            u32 const len = 1u << (ww - 1), st = symStart[s];
            u16 const e = (u16)(s | ((tableLog + 1 - ww) << 8));
            for (u32 u = (u32)w.lane; u < len; u += C::W) S.huf[st + u] = e;
        }
    }
    w.sync();
    return result;
}

// one backward Huffman stream -> n symbols at dst (HUF_decompress1X1_usingDTable_internal_body :574-595)
// The symbol -> bit count -> next symbol chain is pure latency, so the loop is built to put as few instructions
// as possible on it: the unread bits sit left-aligned in a 64-bit register (the table index is a shift of its
// upper half), two symbols are decoded per refill check, and the two words that will slide in next are already in
// registers (their loads were issued two refills earlier).  Bits below the first byte read as zero, `pos` counts
// the unread bits and goes negative on overrun, like BackBits.
ZB_HDN bool huf_decode_stream(const u16* table, u32 log, const u8* src, size_t srcSize, u8* dst, size_t n) {
    if (srcSize < 1) return false;
    u8 const lastByte = src[srcSize - 1];
    if (lastByte == 0) return false;
    BackBits R;                                       // only used for its word addressing / masking helpers
    R.init(src, (int)(srcSize - 1) * 8 + (int)highbit32(lastByte));
    int pos = R.pos;
    // window: hi:lo of the reader with its consumed bits shifted out; then the two look-ahead words
    u64 win = (((u64)R.hi << 32) | R.lo) << R.c;
    int avail = 64 - (int)R.c;
    u32 n0 = R.nx & R.nmask; int wp = R.wp - 1; u32 n1 = R.raw(wp);
    u32 const sh = 32 - log;
#define ZB_HUF_REFILL() do { if (avail <= 32) { win |= (u64)n0 << (32 - avail); avail += 32; n0 = n1 & R.mask_of(wp); wp--; n1 = R.raw(wp); } } while (0)
#define ZB_HUF_SYM(outv) do { u16 const e_ = table[(u32)(win >> 32) >> sh]; u32 const nb_ = e_ >> 8; (outv) = e_ & 0xFF; win <<= nb_; avail -= (int)nb_; pos -= (int)nb_; } while (0)
    size_t i = 0;
    // head: single symbols until dst is 4-byte aligned, then four symbols per 32-bit store
    for (; i < n && ((reinterpret_cast<uintptr_t>(dst) + i) & 3); i++) { u32 s0; ZB_HUF_REFILL(); ZB_HUF_SYM(s0); dst[i] = (u8)s0; }
    for (; i + 4 <= n; i += 4) {
        u32 s0, s1, s2, s3;
        ZB_HUF_REFILL(); ZB_HUF_SYM(s0); ZB_HUF_SYM(s1);          // avail > 32 >= 2 * 11 bits
        ZB_HUF_REFILL(); ZB_HUF_SYM(s2); ZB_HUF_SYM(s3);
        *reinterpret_cast<u32*>(dst + i) = s0 | (s1 << 8) | (s2 << 16) | (s3 << 24);
    }
    for (; i < n; i++) { u32 s0; ZB_HUF_REFILL(); ZB_HUF_SYM(s0); dst[i] = (u8)s0; }
#undef ZB_HUF_REFILL
#undef ZB_HUF_SYM
    return pos == 0;
}

// ------------------------------------------------------- literals section
// ZSTD_decodeLiteralsBlock, zstd_decompress_block.c:134-340, in two steps: parse_literals() reads the header
// (and the Huffman table, into S.huf) and describes the streams; the symbols are decoded afterwards -- by
// lanes 0..3 of the same warp in the fused kernel, by one thread per stream in the batch pipeline.
struct LitInfo {
    u32 mode;            // 0 = raw (bytes at rawOff), 1 = rle (rleByte), 2 = Huffman streams
    u32 litSize;
    u32 nStreams;        // 1 or 4 when mode == 2
    u32 sOff[4], sLen[4];   // stream k: bytes [sOff, sOff+sLen) relative to the section start
    u32 oOff[4], oCnt[4];   // its symbols land at [oOff, oOff+oCnt) of the literal buffer
    u32 rawOff;
    u32 rleByte;
};

template <class C>
ZB_HDN size_t parse_literals(const C& w, DecShared& S, const u8* src, size_t srcSize, size_t blockSizeMax, size_t dstCapacity, LitInfo* li) {
    size_t const expectedWrite = blockSizeMax < dstCapacity ? blockSizeMax : dstCapacity;
    if (srcSize < 2) return ERR(E_corruption_detected);
    u32 const b0 = src[0], type = b0 & 3, lhl = (b0 >> 2) & 3;
    if (type == 3 && !S.litEntropy) return ERR(E_dictionary_corrupted);
    if (type >= 2) {
        if (srcSize < 5) return ERR(E_corruption_detected);
        u32 const lhc = load32(src);
        size_t lhSize, litSize, litCSize; bool single = false;
        switch (lhl) {
        case 0: case 1: default: single = !lhl; lhSize = 3; litSize = (lhc >> 4) & 0x3FF; litCSize = (lhc >> 14) & 0x3FF; break;
        case 2: lhSize = 4; litSize = (lhc >> 4) & 0x3FFF; litCSize = lhc >> 18; break;
        case 3: lhSize = 5; litSize = (lhc >> 4) & 0x3FFFF; litCSize = (lhc >> 22) + ((size_t)src[4] << 10); break;
        }
        if (litSize > blockSizeMax) return ERR(E_corruption_detected);
        if (!single && litSize < 6) return ERR(E_literals_headerWrong);
        if (litCSize + lhSize > srcSize) return ERR(E_corruption_detected);
        if (expectedWrite < litSize) return ERR(E_dstSize_tooSmall);
        const u8* p = src + lhSize; size_t c = litCSize;
        if (type == 2) {
            if (!single && (litSize == 0 || c == 0)) return ERR(E_corruption_detected);
            size_t const h = huf_read_table(w, S, p, c);
            if (isErr(h)) return ERR(E_corruption_detected);
            if (h >= c) return ERR(E_corruption_detected);
            p += h; c -= h;
        }
        // stream geometry (HUF_decompress4X1_usingDTable_internal_body :601-650)
        li->mode = 2; li->litSize = (u32)litSize;
        u32 const pOff = (u32)(p - src);
        if (single) { li->nStreams = 1; li->sOff[0] = pOff; li->sLen[0] = (u32)c; li->oOff[0] = 0; li->oCnt[0] = (u32)litSize; }
        else {
            li->nStreams = 4;
            if (c < 10 || litSize < 6) return ERR(E_corruption_detected);
            size_t const l1 = load16(p), l2 = load16(p + 2), l3 = load16(p + 4);
            if (l1 + l2 + l3 + 6 > c) return ERR(E_corruption_detected);
            size_t const seg = (litSize + 3) / 4;
            if (3 * seg > litSize) return ERR(E_corruption_detected);
            li->sOff[0] = pOff + 6; li->sOff[1] = li->sOff[0] + (u32)l1; li->sOff[2] = li->sOff[1] + (u32)l2; li->sOff[3] = li->sOff[2] + (u32)l3;
            li->sLen[0] = (u32)l1; li->sLen[1] = (u32)l2; li->sLen[2] = (u32)l3; li->sLen[3] = (u32)(c - (l1 + l2 + l3 + 6));
            for (int k = 0; k < 4; k++) { li->oOff[k] = (u32)(k * seg); li->oCnt[k] = (u32)(k < 3 ? seg : litSize - 3 * seg); }
        }
        if (w.lane == 0) S.litEntropy = 1;
        return litCSize + lhSize;
    }
    size_t lhSize, litSize;
    switch (lhl) {
    case 0: case 2: default: lhSize = 1; litSize = b0 >> 3; break;
    case 1: lhSize = 2; litSize = load16(src) >> 4; break;
    case 3: lhSize = 3; if (srcSize < 3) return ERR(E_corruption_detected); litSize = load24(src) >> 4; break;
    }
    if (type == 1) {
        if (lhl == 1 && srcSize < 3) return ERR(E_corruption_detected);
        if (lhl == 3 && srcSize < 4) return ERR(E_corruption_detected);
    }
    if (litSize > blockSizeMax) return ERR(E_corruption_detected);
    if (expectedWrite < litSize) return ERR(E_dstSize_tooSmall);
    li->litSize = (u32)litSize; li->nStreams = 0;
    if (type == 0) {
        if (litSize + lhSize > srcSize) return ERR(E_corruption_detected);
        li->mode = 0; li->rawOff = (u32)lhSize;
        return lhSize + litSize;
    }
    li->mode = 1; li->rleByte = src[lhSize];
    return lhSize + 1;
}

// fused variant: literals end up readable at *litPtr (in place for raw, `scratch` otherwise)
template <class C>
ZB_HDN size_t decode_literals(const C& w, DecShared& S, const u8* src, size_t srcSize, size_t blockSizeMax, size_t dstCapacity,
                              u8* scratch, const u8** litPtr, size_t* litSizeOut) {
    LitInfo li;
    bool const hadEntropy = S.litEntropy != 0;
    size_t const r = parse_literals(w, S, src, srcSize, blockSizeMax, dstCapacity, &li);
    if (isErr(r)) return r;
    *litSizeOut = li.litSize;
    if (li.mode == 0) { *litPtr = src + li.rawOff; return r; }
    if (li.mode == 1) {
        u8 const v = (u8)li.rleByte;
        for (size_t k = (size_t)w.lane; k < li.litSize; k += C::W) scratch[k] = v;
        w.sync();
        *litPtr = scratch; return r;
    }
    bool ok = true;   // lanes 0..3 decode one stream each (all of them on a 1-lane host context)
    for (int k = w.lane; k < (int)li.nStreams; k += C::W)
        ok = huf_decode_stream(S.huf, S.hufLog, src + li.sOff[k], li.sLen[k], scratch + li.oOff[k], li.oCnt[k]) && ok;
    bool const allOk = (w.ballot(!ok) == 0);
    w.sync();
    if (!allOk) { if (w.lane == 0 && !hadEntropy) S.litEntropy = 0; w.sync(); return ERR(E_corruption_detected); }
    *litPtr = scratch;
    return r;
}

// Sequences section header (ZSTD_decodeSeqHeaders :695-775) + table construction (ZSTD_buildSeqTable :647-693):
// lane 0 parses, lanes 0..2 build one table each into S.fse[] / S.fseLog[].  Returns bytes consumed.
template <class C>
ZB_HDN size_t parse_seq_section(const C& w, DecShared& S, const u8* ip, size_t left, size_t cap, int* nbSeqOut) {
    size_t hdr = 0; int nbSeq = 0;
    if (w.lane == 0) {
        do {
            const u8* p = ip; size_t l = left;
            if (l < 1) { hdr = ERR(E_srcSize_wrong); break; }
            nbSeq = *p++; l--;
            if (nbSeq > 0x7F) {
                if (nbSeq == 0xFF) { if (l < 2) { hdr = ERR(E_srcSize_wrong); break; } nbSeq = (int)load16(p) + (int)LONGNBSEQ; p += 2; l -= 2; }
                else { if (l < 1) { hdr = ERR(E_srcSize_wrong); break; } nbSeq = ((nbSeq - 0x80) << 8) + *p++; l--; }
            }
            if (nbSeq == 0) { if (l != 0) hdr = ERR(E_corruption_detected); else hdr = (size_t)(p - ip); break; }
            if (l < 1) { hdr = ERR(E_srcSize_wrong); break; }
            u32 const modes = *p++; l--;
            if (modes & 3) { hdr = ERR(E_corruption_detected); break; }
            u32 const type[3] = { modes >> 6, (modes >> 4) & 3, (modes >> 2) & 3 };
            u32 const maxSym[3] = { MaxLL, MaxOff, MaxML }, maxLog[3] = { LLFSELog, OffFSELog, MLFSELog };
            bool fail = false;
            for (int t = 0; t < 3 && !fail; t++) {
                S.fseMode[t] = type[t];
                switch (type[t]) {
                case 1:   // rle
                    if (!l) { hdr = ERR(E_corruption_detected); fail = true; break; }
                    if (*p > maxSym[t]) { hdr = ERR(E_corruption_detected); fail = true; break; }
                    S.fseMax[t] = *p; p++; l--; break;
                case 0: break;
                case 3: if (!S.fseEntropy) { hdr = ERR(E_corruption_detected); fail = true; } break;
                default: {
                    u32 m = maxSym[t], tl;
                    size_t const h = read_ncount(S.norm[t], &m, &tl, p, l);
                    if (isErr(h) || tl > maxLog[t]) { hdr = ERR(E_corruption_detected); fail = true; break; }
                    S.fseMax[t] = m; S.fseLog[t] = tl; p += h; l -= h; } break;
                }
            }
            if (fail) break;
            hdr = (size_t)(p - ip);
        } while (0);
    }
    w.sync();
    hdr = w.bcast(hdr); nbSeq = w.bcast(nbSeq);
    if (isErr(hdr)) return hdr;
    if (nbSeq) {
        if (cap == 0) return ERR(E_dstSize_tooSmall);
        // table construction, all lanes on one table after the other (ZSTD_buildSeqTable :647-693)
        for (int t = 0; t < 3; t++) {
            u32 const mode = S.fseMode[t];
            if (mode == 1) { if (w.lane == 0) { S.fse[t][0] = (S.fseMax[t] << 24); S.fseLog[t] = 0; } }
            else if (mode == 0) {
                const i16* dn = t == 0 ? ZB_T.LL_defaultNorm : t == 1 ? ZB_T.OF_defaultNorm : ZB_T.ML_defaultNorm;
                u32 const dmax = t == 0 ? MaxLL : t == 1 ? DefaultMaxOff : MaxML, dlog = t == 1 ? 5 : 6;
                for (u32 sy = (u32)w.lane; sy <= dmax; sy += C::W) S.norm[t][sy] = dn[sy];
                w.sync();
                build_fse_dtable_warp(w, S.fse[t], S.norm[t], dmax, dlog, S.symNext[t], S.spread[t], S.cum);
                if (w.lane == 0) S.fseLog[t] = dlog;
            } else if (mode == 2) {
                build_fse_dtable_warp(w, S.fse[t], S.norm[t], S.fseMax[t], S.fseLog[t], S.symNext[t], S.spread[t], S.cum);
            }
            w.sync();
        }
    }
    *nbSeqOut = nbSeq;
    return hdr;
}

// --------------------------------------------------- one compressed block
// ZSTD_decompressBlock_internal :2066-2174.  `frameStart` is the first output byte of
// the frame (offsets may reach back that far); writes at op, at most `cap` bytes.
template <class C>
ZB_HDN size_t decode_block(const C& w, DecShared& S, const u8* frameStart, u8* op0, size_t cap,
                           const u8* src, size_t srcSize, size_t blockSizeMax, u8* scratch) {
    if (srcSize > blockSizeMax) return ERR(E_srcSize_wrong);
    const u8* ip = src; size_t left = srcSize;
    const u8* lit = nullptr; size_t litSize = 0;
    {   size_t const r = decode_literals(w, S, ip, left, blockSizeMax, cap, scratch, &lit, &litSize);
        if (isErr(r)) return r;
        ip += r; left -= r;
    }
    int nbSeq = 0;
    {   size_t const hdr = parse_seq_section(w, S, ip, left, cap, &nbSeq);
        if (isErr(hdr)) return hdr;
        ip += hdr; left -= hdr; }

    u8* op = op0; u8* const oend = op0 + cap;   // like the reference, a block is bounded by the destination only
    const u8* const litEnd = lit + litSize;
    size_t err = 0;
    if (nbSeq) {
        S.fseEntropy = 1;
        if (left < 1 || ip[left - 1] == 0) return ERR(E_corruption_detected);
        // lane-0 decoder state (registers of lane 0 only)
        i64 pos = (i64)(left - 1) * 8 + highbit32(ip[left - 1]);
        u32 sLL = 0, sOF = 0, sML = 0, rep0 = S.rep[0], rep1 = S.rep[1], rep2 = S.rep[2];
        u32 const logLL = S.fseLog[0], logOF = S.fseLog[1], logML = S.fseLog[2];
        if (w.lane == 0) {
            pos -= logLL; sLL = (u32)peek_bits(ip, pos, logLL);
            pos -= logOF; sOF = (u32)peek_bits(ip, pos, logOF);
            pos -= logML; sML = (u32)peek_bits(ip, pos, logML);
        }
        int remaining = nbSeq;
        while (remaining > 0) {
            int const cnt = remaining < (int)SEQ_BATCH ? remaining : (int)SEQ_BATCH;
            if (w.lane == 0) {
                for (int k = 0; k < cnt; k++) {
                    u32 const eLL = S.fse[0][sLL], eOF = S.fse[1][sOF], eML = S.fse[2][sML];
                    u32 const llc = eLL >> 24, ofc = eOF >> 24, mlc = eML >> 24;
                    u32 const llBits = ZB_T.LL_bits[llc], mlBits = ZB_T.ML_bits[mlc], ofBits = ofc;
                    u32 litLength = ZB_T.LL_base[llc], matchLength = ZB_T.ML_base[mlc], offset;
                    if (ofBits > 1) {
                        pos -= ofBits;
                        offset = ((1u << ofBits) - 3) + (u32)peek_bits(ip, pos, ofBits);
                        rep2 = rep1; rep1 = rep0; rep0 = offset;
                    } else {
                        u32 const ll0 = (litLength == 0);
                        if (ofBits == 0) {
                            offset = ll0 ? rep1 : rep0;
                            rep1 = ll0 ? rep0 : rep1; rep0 = offset;
                        } else {
                            pos -= 1;
                            u32 const idx = 1 + ll0 + (u32)peek_bits(ip, pos, 1);
                            u32 temp = (idx == 3) ? rep0 - 1 : (idx == 1 ? rep1 : rep2);
                            temp -= !temp;
                            if (idx != 1) rep2 = rep1;
                            rep1 = rep0; rep0 = temp; offset = temp;
                        }
                    }
                    {   u32 const nb = mlBits + llBits;
                        pos -= nb;
                        u32 const x = (u32)peek_bits(ip, pos, nb);
                        matchLength += x >> llBits;
                        litLength += x & ((1u << llBits) - 1);
                    }
                    if (remaining - k > 1) {
                        u32 const nLL = (eLL >> 16) & 0xFF, nML = (eML >> 16) & 0xFF, nOF = (eOF >> 16) & 0xFF;
                        u32 const nb = nLL + nML + nOF;
                        pos -= nb;
                        u32 const y = (u32)peek_bits(ip, pos, nb);
                        sLL = (eLL & 0xFFFF) + (y >> (nML + nOF));
                        sML = (eML & 0xFFFF) + ((y >> nOF) & ((1u << nML) - 1));
                        sOF = (eOF & 0xFFFF) + (y & ((1u << nOF) - 1));
                    }
                    S.sLit[k] = litLength; S.sMatch[k] = matchLength; S.sOff[k] = offset;
                }
            }
            w.sync();
            // execute the batch, in order (ZSTD_execSequence :1001-1096 / _End :905-948)
            for (int k = 0; k < cnt; k++) {
                size_t const ll = S.sLit[k], ml = S.sMatch[k], off = S.sOff[k];
                if (ll + ml > (size_t)(oend - op)) { err = ERR(E_dstSize_tooSmall); break; }
                if (ll > (size_t)(litEnd - lit)) { err = ERR(E_corruption_detected); break; }
                for (size_t j = (size_t)w.lane; j < ll; j += C::W) op[j] = lit[j];
                op += ll; lit += ll;
                if (off > (size_t)(op - frameStart)) { err = ERR(E_corruption_detected); break; }
                w.sync();
                const u8* const m = op - off;
                if (off >= ml) { for (size_t j = (size_t)w.lane; j < ml; j += C::W) op[j] = m[j]; }
                else if (C::W == 1) { for (size_t j = 0; j < ml; j++) op[j] = m[j]; }
                else { for (size_t j = (size_t)w.lane; j < ml; j += C::W) op[j] = m[j % off]; }
                op += ml;
                w.sync();
            }
            if (err) break;
            remaining -= cnt;
            w.sync();
        }
        if (err) return err;
        bool bad = false;
        if (w.lane == 0) { bad = (pos != 0); S.rep[0] = rep0; S.rep[1] = rep1; S.rep[2] = rep2; }
        bad = w.bcast((u32)bad) != 0;
        w.sync();
        if (bad) return ERR(E_corruption_detected);
    }
    {   size_t const last = (size_t)(litEnd - lit);
        if (last > (size_t)(oend - op)) return ERR(E_dstSize_tooSmall);
        for (size_t j = (size_t)w.lane; j < last; j += C::W) op[j] = lit[j];
        op += last;
        w.sync();
    }
    return (size_t)(op - op0);
}

// ------------------------------------------------------------ frame layer
struct FrameHeader { u32 headerSize; u64 contentSize; u64 windowSize; u32 blockSizeMax; u32 checksum, skippable, skipLen, hasContentSize, dictID; };

// ZSTD_getFrameHeader_advanced :447-551.  0 = ok, >0 = bytes wanted, or error.
// magicless != 0: ZSTD_f_zstd1_magicless (N/zstd.h:1386) -- the frame starts at its descriptor byte, no magic number,
// hence no skippable frames either.
ZB_HDN size_t read_frame_header(FrameHeader* fh, const u8* src, size_t srcSize, u32 magicless = 0) {
    fh->headerSize = 0; fh->contentSize = 0; fh->windowSize = 0; fh->blockSizeMax = 0;
    fh->checksum = fh->skippable = fh->skipLen = 0; fh->hasContentSize = 0; fh->dictID = 0;
    u32 const mlen = magicless ? 0u : 4u;                 // ZSTD_startingInputLength - 1
    if (srcSize < mlen + 1) {
        if (srcSize > 0 && !magicless) {
            u32 const k = (u32)(srcSize < 4 ? srcSize : 4);
            u32 got = 0; for (u32 i = 0; i < k; i++) got |= (u32)src[i] << (8 * i);
            u32 const m = (k == 4) ? 0xFFFFFFFFu : ((1u << (8 * k)) - 1);
            if ((got & m) != (MAGIC & m)) {
                if ((got & m & 0xFFFFFFF0u) != (0x184D2A50u & m & 0xFFFFFFF0u)) return ERR(E_prefix_unknown);
            }
        }
        return mlen + 1;
    }
    if (!magicless) {
        u32 const magic = load32(src);
        if (magic != MAGIC) {
            if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {
                if (srcSize < 8) return 8;
                fh->skippable = 1; fh->skipLen = load32(src + 4); fh->headerSize = 8; fh->dictID = magic - 0x184D2A50u;
                return 0;
            }
            return ERR(E_prefix_unknown);
        }
    }
    u32 const fhd = src[mlen], dictID = fhd & 3, single = (fhd >> 5) & 1, fcsID = fhd >> 6;
    u32 const didSize = dictID == 3 ? 4 : dictID, fcsSize = fcsID == 0 ? 0 : (1u << fcsID);
    size_t const hs = mlen + 1 + !single + didSize + fcsSize + (single && !fcsID);
    if (srcSize < hs) return hs;
    fh->headerSize = (u32)hs;
    if (fhd & 0x08) return ERR(E_frameParameter_unsupported);
    size_t pos = mlen + 1;
    if (!single) {
        u32 const wl = src[pos++], windowLog = (wl >> 3) + 10;
        if (windowLog > 31) return ERR(E_frameParameter_windowTooLarge);
        fh->windowSize = 1ull << windowLog; fh->windowSize += (fh->windowSize >> 3) * (wl & 7);
    }
    for (u32 i = 0; i < didSize; i++) fh->dictID |= (u32)src[pos + i] << (8 * i);
    pos += didSize;
    fh->hasContentSize = 1;
    switch (fcsID) {
    case 0: if (single) fh->contentSize = src[pos]; else fh->hasContentSize = 0; break;
    case 1: fh->contentSize = (u64)load16(src + pos) + 256; break;
    case 2: fh->contentSize = load32(src + pos); break;
    default: fh->contentSize = load64_n(src + pos, 8); break;
    }
    if (single) fh->windowSize = fh->contentSize;
    fh->blockSizeMax = (u32)(fh->windowSize < BLOCKSIZE_MAX ? fh->windowSize : BLOCKSIZE_MAX);
    fh->checksum = (fhd >> 2) & 1;
    return 0;
}

// ZSTD_decompressMultiFrame :1070-1169 + ZSTD_decompressFrame :953-1066.
// Uniform across the warp; returns regenerated size or an error code.
template <class C>
ZB_HDN size_t decompress_item(const C& w, DecShared& S, const u8* src, size_t srcSize, u8* dst, size_t dstCapacity, u8* scratch, u32 magicless = 0) {
    size_t total = 0; bool more = false;
    while (srcSize >= (magicless ? 1u : 4u)) {
        if (!magicless && srcSize >= 8 && (load32(src) & 0xFFFFFFF0u) == 0x184D2A50u) {
            size_t const skip = 8 + (size_t)load32(src + 4);
            if (skip > srcSize) return ERR(E_srcSize_wrong);
            src += skip; srcSize -= skip; continue;
        }
        // ---- one frame
        if (srcSize < (magicless ? 2u : 6u) + 3) return ERR(E_srcSize_wrong);     // ZSTD_FRAMEHEADERSIZE_MIN(format) + block header
        {   // ZSTD_decompressFrame :972-979: the header size comes from the descriptor byte alone, and a frame too short for header + one
            // block header is srcSize_wrong before the header itself is validated (magic, reserved bit, window)
            u32 const fhd = src[magicless ? 0 : 4], did = fhd & 3, single = (fhd >> 5) & 1, fcs = fhd >> 6;
            size_t const hs = (magicless ? 1u : 5u) + !single + (did == 3 ? 4 : did) + (fcs == 0 ? 0 : (1u << fcs)) + (single && !fcs);
            if (srcSize < hs + 3) return ERR(E_srcSize_wrong); }
        FrameHeader fh;
        {   size_t const r = read_frame_header(&fh, src, srcSize, magicless);
            if (isErr(r)) return (more && r == ERR(E_prefix_unknown)) ? ERR(E_srcSize_wrong) : r;
            if (r > 0) return ERR(E_srcSize_wrong);
            if (srcSize < fh.headerSize + 3) return ERR(E_srcSize_wrong);
            // the frame names a dictionary and none can be loaded here (ZSTD_decodeFrameHeader, N/decompress/zstd_decompress.c:706-707)
            if (fh.dictID != 0) return ERR(E_dictionary_wrong);
        }
        const u8* ip = src + fh.headerSize; size_t left = srcSize - fh.headerSize;
        if (w.lane == 0) { S.rep[0] = 1; S.rep[1] = 4; S.rep[2] = 8; S.litEntropy = 0; S.fseEntropy = 0; }
        w.sync();
        size_t written = 0;
        for (;;) {
            if (left < 3) return ERR(E_srcSize_wrong);
            u32 const bh = load24(ip), type = (bh >> 1) & 3; size_t cSize = bh >> 3;
            if (type == 3) return ERR(E_corruption_detected);
            if (type == 1) cSize = 1;
            ip += 3; left -= 3;
            if (cSize > left) return ERR(E_srcSize_wrong);
            size_t decoded;
            if (type == 2) {
                decoded = decode_block(w, S, dst, dst + written, dstCapacity - written, ip, cSize, fh.blockSizeMax, scratch);
                if (isErr(decoded)) return decoded;
            } else if (type == 0) {
                if (cSize > dstCapacity - written) return ERR(E_dstSize_tooSmall);
                for (size_t j = (size_t)w.lane; j < cSize; j += C::W) dst[written + j] = ip[j];
                decoded = cSize; w.sync();
            } else {
                size_t const rl = bh >> 3;
                if (rl > dstCapacity - written) return ERR(E_dstSize_tooSmall);
                u8 const v = *ip;
                for (size_t j = (size_t)w.lane; j < rl; j += C::W) dst[written + j] = v;
                decoded = rl; w.sync();
            }
            written += decoded; ip += cSize; left -= cSize;
            if (bh & 1) break;
        }
        if (fh.hasContentSize && written != fh.contentSize) return ERR(E_corruption_detected);
        if (fh.checksum) {
            if (left < 4) return ERR(E_checksum_wrong);
            u32 calc = 0;
            if (w.lane == 0) calc = (u32)xxh64(dst, written);
            calc = w.bcast(calc);
            if (calc != load32(ip)) return ERR(E_checksum_wrong);
            ip += 4; left -= 4;
        }
        dst += written; dstCapacity -= written; total += written; more = true;
        src = ip; srcSize = left;
    }
    if (srcSize) return ERR(E_srcSize_wrong);
    return total;
}

}  // namespace zb
