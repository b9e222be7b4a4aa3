// zb_decode_fast.cuh -- staged ("transposed") decoder for batches of single-block frames.
//
// In a batch of thousands of frames the serial chains of the format (Huffman streams, the three coupled FSE
// states of the sequence stream) are the critical path.  The fused kernel (zb_decode.cuh) walks them with one
// lane of a warp while 31 lanes wait.  This pipeline gives every chain its own *thread* and runs all of them
// at once, then executes the sequences with whole warps:
//
//   A  dec_prepare   warp / frame    frame + block + section headers, Huffman and FSE decode tables -> HBM
//   B  dec_huf       thread / stream 4 x n threads, each decodes one Huffman stream into the literal buffer
//   C  dec_seq       thread / frame  sequence bitstream -> (litLength, matchLength, offset) packed in 8 bytes
//   D  dec_exec      warp / frame    32 sequences per step: positions by prefix sums, literal runs and
//                                    independent matches copied lane-per-sequence, dependent ones in waves
//
// An item is eligible when it is exactly one frame with one (last) block, carries its content size and no
// checksum; everything else (multi-block, multi-frame, skippable, checksummed, > MAXS sequences) is left to
// the fused kernel.  Results -- including the error code of corrupted input -- are identical to the fused
// path: the stage statuses are combined in the order the fused decoder would have met the errors.
//
// Reference behaviour reproduced: see zb_decode.cuh (N/decompress/zstd_decompress_block.c:134-340,695-775,
// 1001-1096,1229-1346,1615-1690; N/decompress/huf_decompress.c:574-698).
#pragma once
#include "zb_decode.cuh"

namespace zb {

constexpr u32 FAST_MAXS = 43776;            // > 131072 / MINMATCH: more sequences cannot fit a 128 KB block
constexpr u32 FAST_HUF_ENTRIES = 1u << HUF_TABLELOG_MAX;
constexpr u32 FAST_FSE_ENTRIES = 512 + 256 + 512;  // LL | OF | ML, packed u32 entries nextState | nbBits << 16 | symbol << 24
constexpr u32 FAST_FSE_OF = 512, FAST_FSE_ML = 768;

struct DecDesc {
    u32 mode;            // 0 = not eligible (fused kernel handles the item), 1 = compressed block, 2 = raw block, 3 = rle block
    u32 stA1, stB, stA2, stC, stD;   // positive error codes per stage (0 = fine)
    u32 blockOff, cSize; // block content inside the item
    u32 contentSize;
    u32 litMode, litSize, rawOff, rleByte, hufLog, nStreams;
    u32 sOff[4], sLen[4], oOff[4], oCnt[4];      // relative to the block content / the literal buffer
    u32 nbSeq, seqOff, seqLen;                   // sequence bitstream inside the block content
    u32 logLL, logOF, logML;
    u32 regen;
    u32 hufLitSize;      // litSize when the literals are Huffman coded, else 0 (sort key of the Huffman stage)
    u32 pad[2];
};

// ---------------------------------------------------------------------------------------- stage A
// All lanes of a warp; S is the warp's shared scratch; tables are copied out to hufOut / fseOut.
template <class C>
ZB_HDN void dec_prepare(const C& w, DecShared& S, const u8* src, size_t srcSize, size_t dstCapacity, DecDesc* d, u16* hufOut, u32* fseOut) {
    DecDesc L;   // built in registers/local, stored by lane 0 at the end
    L.mode = 0; L.stA1 = L.stB = L.stA2 = L.stC = L.stD = 0; L.regen = 0; L.nbSeq = 0; L.litMode = 0; L.litSize = 0; L.nStreams = 0;
    L.blockOff = L.cSize = L.contentSize = L.rawOff = L.rleByte = L.hufLog = L.seqOff = L.seqLen = L.logLL = L.logOF = L.logML = 0;
    for (int k = 0; k < 4; k++) { L.sOff[k] = L.sLen[k] = L.oOff[k] = L.oCnt[k] = 0; }
    L.pad[0] = L.pad[1] = 0; L.hufLitSize = 0;
    do {
        if (srcSize < 9) break;
        if (load32(src) != MAGIC) break;
        FrameHeader fh;
        size_t const r = read_frame_header(&fh, src, srcSize);
        if (r != 0 || fh.skippable || fh.checksum || !fh.hasContentSize || fh.dictID != 0) break;
        if (fh.contentSize > BLOCKSIZE_MAX) break;
        if (srcSize < fh.headerSize + 3) break;
        const u8* const bp = src + fh.headerSize;
        u32 const bh = load24(bp), type = (bh >> 1) & 3; size_t cSize = bh >> 3;
        if (!(bh & 1) || type == 3) break;
        if (type == 1) cSize = 1;
        if (fh.headerSize + 3 + cSize != srcSize) break;        // exactly one frame, one block
        L.blockOff = fh.headerSize + 3; L.cSize = (u32)cSize; L.contentSize = (u32)fh.contentSize;
        if (type == 0) { L.mode = 2; L.regen = (u32)cSize; break; }
        if (type == 1) { L.mode = 3; L.regen = bh >> 3; L.rleByte = bp[3]; break; }
        // compressed block: from here on the item is ours, errors are recorded instead of punting
        L.mode = 1;
        const u8* const blk = bp + 3;
        if (cSize > fh.blockSizeMax) { L.stA1 = E_srcSize_wrong; break; }
        if (w.lane == 0) { S.litEntropy = 0; S.fseEntropy = 0; }
        w.sync();
        LitInfo li;
        size_t const lr = parse_literals(w, S, blk, cSize, fh.blockSizeMax, dstCapacity, &li);
        if (isErr(lr)) { L.stA1 = (u32)(0 - lr); break; }
        L.litMode = li.mode; L.litSize = li.litSize; L.rawOff = li.rawOff; L.rleByte = li.rleByte; L.nStreams = li.nStreams; L.hufLog = S.hufLog;
        for (int k = 0; k < 4; k++) { L.sOff[k] = li.sOff[k]; L.sLen[k] = li.sLen[k]; L.oOff[k] = li.oOff[k]; L.oCnt[k] = li.oCnt[k]; }
        if (li.mode == 2) {
            L.hufLitSize = li.litSize;
            u32 const nE = 1u << S.hufLog;
            for (u32 i = (u32)w.lane; i < nE; i += C::W) hufOut[i] = S.huf[i];
        }
        int nbSeq = 0;
        size_t const hr = parse_seq_section(w, S, blk + lr, cSize - lr, dstCapacity, &nbSeq);
        if (isErr(hr)) { L.stA2 = (u32)(0 - hr); break; }
        if ((u32)nbSeq > FAST_MAXS) { L.mode = 0; break; }      // pathological: let the fused kernel deal with it
        L.nbSeq = (u32)nbSeq; L.seqOff = (u32)(lr + hr); L.seqLen = (u32)(cSize - lr - hr);
        if (nbSeq) {
            L.logLL = S.fseLog[0]; L.logOF = S.fseLog[1]; L.logML = S.fseLog[2];
            for (int t = 0; t < 3; t++) {
                u32 const nE = 1u << S.fseLog[t];
                u32* const out = fseOut + (t == 0 ? 0 : t == 1 ? FAST_FSE_OF : FAST_FSE_ML);
                // bits 10..14 of the copy carry the code's extra-bit count, so that stage C learns every bit count of
                // a sequence from the three state entries alone
                for (u32 i = (u32)w.lane; i < nE; i += C::W) {
                    u32 const e = S.fse[t][i], sym = e >> 24;
                    u32 const extra = t == 0 ? ZB_T.LL_bits[sym] : t == 1 ? sym : ZB_T.ML_bits[sym];
                    out[i] = e | (extra << 10);
                }
            }
        }
    } while (0);
    w.sync();
    if (w.lane == 0) *d = L;
    w.sync();
}

// ---------------------------------------------------------------------------------------- stage B
// one thread per (frame, stream)
ZB_HDN void dec_huf(DecDesc* d, int k, const u8* blk, const u16* huf, u8* lit) {
    if (d->mode != 1 || d->stA1 || d->litMode != 2 || k >= (int)d->nStreams) return;
    if (!huf_decode_stream(huf, d->hufLog, blk + d->sOff[k], d->sLen[k], lit + d->oOff[k], d->oCnt[k])) d->stB = E_corruption_detected;
}

// ---------------------------------------------------------------------------------------- stage C
// one thread per frame: ZSTD_decodeSequence :1229-1346 for every sequence, offsets resolved against the
// repcode history {1,4,8}; sequences are stored as litLength | matchLength << 18 | offset << 36 (each < 2^18
// for a block of at most 128 KB; larger values can only come from corrupt input and are clamped to 2^18-1 /
// 2^28-1, which the executor rejects exactly like the originals).
// `fse` = the frame's three tables (FAST_FSE_ENTRIES u32, in shared memory on the GPU), `ct` = code tables
// (shared memory copy on the GPU: per-lane indices would serialise in the constant cache).
// The decoder is a small state machine (begin / step / end) so that a kernel can keep every lane of a warp busy:
// a lane that finishes its frame picks up the next one while the other lanes keep stepping in lockstep.
struct SeqDecoder {
    DecDesc* d; const u32* tLL; const u32* tOF; const u32* tML; const CodeTables* ct; u64* seqOut;
    BackBits B; u32 sLL, sOF, sML, rep0, rep1, rep2, k, nbSeq;

    // false: nothing to decode for this item (not eligible, earlier error, no sequences, unusable stream)
    ZB_HD bool begin(DecDesc* d_, const u8* blk, const u32* fse, const CodeTables* ct_, u64* out) {
        d = d_;
        if (d->mode != 1 || d->stA1 || d->stA2 || d->nbSeq == 0) return false;
        const u8* const ip = blk + d->seqOff; size_t const left = d->seqLen;
        if (left < 1 || ip[left - 1] == 0) { d->stC = E_corruption_detected; return false; }
        tLL = fse; tOF = fse + FAST_FSE_OF; tML = fse + FAST_FSE_ML; ct = ct_; seqOut = out;
        B.init(ip, (int)(left - 1) * 8 + (int)highbit32(ip[left - 1]));
        rep0 = 1; rep1 = 4; rep2 = 8;
        sLL = B.take(d->logLL); sOF = B.take(d->logOF); sML = B.take(d->logML);
        k = 0; nbSeq = d->nbSeq;
        return true;
    }
    // one sequence; true while more remain
    ZB_HD bool step() {
        // entry: nextState (bits 0..9) | extra bits of the code (10..14) | nbBits (16..23) | code (24..31)
        u32 const eLL = tLL[sLL], eOF = tOF[sOF], eML = tML[sML];
        u32 const llc = eLL >> 24, ofc = eOF >> 24, mlc = eML >> 24;
        u32 const llBits = (eLL >> 10) & 31, mlBits = (eML >> 10) & 31, ofBits = ofc;
        u32 const nLL = (eLL >> 16) & 0xFF, nML = (eML >> 16) & 0xFF, nOF = (eOF >> 16) & 0xFF;
        bool const lastSeq = (k + 1 == nbSeq);
        // read order: offset bits, ML extra, LL extra, then LL / ML / OF state bits (ZSTD_decodeSequence :1229-1346)
        u32 const ofVal = B.take(ofBits);
        u32 const x = B.take(mlBits + llBits);
        u32 const y = lastSeq ? 0 : B.take(nLL + nML + nOF);
        u32 const matchLength = ct->ML_base[mlc] + (x >> llBits);
        u32 const litLength = ct->LL_base[llc] + (x & ((1u << llBits) - 1));
        u32 offset;
        if (ofBits > 1) { offset = ((1u << ofBits) - 3) + ofVal; rep2 = rep1; rep1 = rep0; rep0 = offset; }
        else {
            u32 const ll0 = (llc == 0);          // :1300 tests litLength base == 0, true for code 0 only
            if (ofBits == 0) { offset = ll0 ? rep1 : rep0; rep1 = ll0 ? rep0 : rep1; rep0 = offset; }
            else {
                u32 const idx = 1 + ll0 + ofVal;
                u32 temp = (idx == 3) ? rep0 - 1 : (idx == 1 ? rep1 : rep2);
                temp -= !temp;
                if (idx != 1) rep2 = rep1;
                rep1 = rep0; rep0 = temp; offset = temp;
            }
        }
        if (!lastSeq) {
            sLL = (eLL & 0x3FF) + (y >> (nML + nOF));
            sML = (eML & 0x3FF) + ((y >> nOF) & ((1u << nML) - 1));
            sOF = (eOF & 0x3FF) + (y & ((1u << nOF) - 1));
        }
        u64 const l = litLength > 0x3FFFF ? 0x3FFFF : litLength, m = matchLength > 0x3FFFF ? 0x3FFFF : matchLength;
        u64 const o = offset > 0xFFFFFFFu ? 0xFFFFFFFu : offset;
        seqOut[k] = l | (m << 18) | (o << 36);
        return ++k < nbSeq;
    }
    ZB_HD void end() { if (B.pos != 0) d->stC = E_corruption_detected; }
};

ZB_HDN void dec_seq(DecDesc* d, const u8* blk, const u32* fse, const CodeTables* ct, u64* seqOut) {
    SeqDecoder D;
    if (!D.begin(d, blk, fse, ct, seqOut)) return;
    while (D.step()) {}
    D.end();
}

// ---------------------------------------------------------------------------------------- stage D
// warp per frame.  Returns the item's final result (regenerated size or error code, libzstd convention).
template <class C>
ZB_HDN size_t dec_exec(const C& w, const DecDesc* dp, const u8* item, const u8* litBuf, const u64* seqs, u8* dst, size_t cap) {
    DecDesc const& d = *dp;
    if (d.mode == 2) {           // raw block
        if (d.cSize > cap) return ERR(E_dstSize_tooSmall);
        const u8* const s = item + d.blockOff;
        for (u32 j = (u32)w.lane; j < d.cSize; j += C::W) dst[j] = s[j];
        w.sync();
        return d.cSize == d.contentSize ? d.cSize : ERR(E_corruption_detected);
    }
    if (d.mode == 3) {           // rle block
        if (d.regen > cap) return ERR(E_dstSize_tooSmall);
        u8 const v = (u8)d.rleByte;
        for (u32 j = (u32)w.lane; j < d.regen; j += C::W) dst[j] = v;
        w.sync();
        return d.regen == d.contentSize ? d.regen : ERR(E_corruption_detected);
    }
    // error precedence of the fused decoder: literals header/table, Huffman streams, sequences header/tables, then execution
    if (d.stA1) return ERR((int)d.stA1);
    if (d.stB) return ERR((int)d.stB);
    if (d.stA2) return ERR((int)d.stA2);
    if (d.nbSeq && cap == 0) return ERR(E_dstSize_tooSmall);
    if (d.nbSeq && d.stC && d.seqLen < 1) return ERR((int)d.stC);            // unusable stream: nothing was decoded
    if (d.nbSeq && d.stC && (item + d.blockOff + d.seqOff)[d.seqLen - 1] == 0) return ERR((int)d.stC);
    const u8* const lit = d.litMode == 0 ? item + d.blockOff + d.rawOff : litBuf;
    bool const rle = d.litMode == 1; u8 const rleByte = (u8)d.rleByte;
    u32 const litSize = d.litSize, nbSeq = d.nbSeq;
    u32 op = 0, lp = 0;     // output / literal cursors (uniform)
    for (u32 base = 0; base < nbSeq; base += C::W) {
        u32 const i = base + (u32)w.lane;
        u64 const q = i < nbSeq ? seqs[i] : 0;
        u32 const ll = (u32)(q & 0x3FFFF), ml = (u32)((q >> 18) & 0x3FFFF), off = (u32)(q >> 36);
        u32 const preOut = w.exscan(ll + ml), preLit = w.exscan(ll);
        u32 const o = op + preOut, ls = lp + preLit, md = o + ll;
        // validity in sequence order (ZSTD_execSequenceEnd :919-932)
        u32 code = 0;
        if (i < nbSeq) {
            if ((size_t)ll + ml > cap - (size_t)(o < cap ? o : cap) || o > cap) code = E_dstSize_tooSmall;
            else if (ll > litSize - (ls < litSize ? ls : litSize) || ls > litSize) code = E_corruption_detected;
            else if (off > md) code = E_corruption_detected;
        }
        u32 const badMask = w.ballot(code != 0);
        u32 const good = badMask ? ((1u << ctz32(badMask)) - 1) : C::FULL;      // lanes before the first failure
        bool const mine = ((good >> w.lane) & 1) && i < nbSeq;
        // literals: short runs by their own lane, long ones by everybody
        if (mine && ll && ll <= 32) { if (rle) { for (u32 k = 0; k < ll; k++) dst[o + k] = rleByte; } else copy_fwd(dst + o, lit + ls, ll); }
        {   u32 big = w.ballot(mine && ll > 32);
            while (big) {
                int const b = (int)ctz32(big); big &= big - 1;
                u32 const L = w.shfl(ll, b), oo = w.shfl(o, b), lss = w.shfl(ls, b);
                for (u32 k = (u32)w.lane; k < L; k += C::W) dst[oo + k] = rle ? rleByte : lit[lss + k];
            } }
        w.sync();
        // matches: everything before the first unfinished match destination is final; a match whose source ends
        // there may run now, in its own lane (byte-serial, so overlapping copies replicate correctly)
        u32 pending = w.ballot(mine && ml > 0);
        while (pending) {
            int const f = (int)ctz32(pending);
            u32 const frontier = w.shfl(md, f), fml = w.shfl(ml, f);
            if (fml > 64) {          // long match at the frontier: all lanes, period trick for overlaps
                u32 const foff = w.shfl(off, f);
                u8* const t = dst + frontier; const u8* const m = t - foff;
                if (foff >= fml) { for (u32 k = (u32)w.lane; k < fml; k += C::W) t[k] = m[k]; }
                else if (C::W == 1) { for (u32 k = 0; k < fml; k++) t[k] = m[k]; }
                else { for (u32 k = (u32)w.lane; k < fml; k += C::W) t[k] = m[k % foff]; }
                pending &= pending - 1;
                w.sync();
                continue;
            }
            bool const ready = ((pending >> w.lane) & 1) && ((int)w.lane == f || (ml <= 64 && md - off + ml <= frontier));
            if (ready) copy_fwd(dst + md, dst + md - off, ml);
            pending &= ~w.ballot(ready);
            w.sync();
        }
        if (badMask) return ERR((int)w.shfl(code, (int)ctz32(badMask)));
        op += w.bcast(preOut + ll + ml, C::W - 1); lp += w.bcast(preLit + ll, C::W - 1);
    }
    if (nbSeq && d.stC) return ERR((int)d.stC);
    {   u32 const last = litSize - lp;
        if (last > cap - op) return ERR(E_dstSize_tooSmall);
        for (u32 k = (u32)w.lane; k < last; k += C::W) dst[op + k] = rle ? rleByte : lit[lp + k];
        op += last;
        w.sync(); }
    return op == d.contentSize ? op : ERR(E_corruption_detected);
}

}  // namespace zb
