// zb_hostsim.cpp -- TEST-ONLY host instantiation (1-lane warp context) of the codec
// templates in zstd_jni_b200/csrc/*.cuh.  It lets `pytest -m "not gpu"` exercise the
// very source the CUDA kernels are built from on a machine without a GPU.  It is never
// loaded by the product path (zstd_jni_b200/lib/libzstdb200.so has no CPU fallback).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../../zstd_jni_b200/csrc/zb_decode.cuh"
#include "../../zstd_jni_b200/csrc/zb_decode_fast.cuh"
#include "../../zstd_jni_b200/csrc/zb_encode.cuh"
#include "simt_emu.h"

extern "C" {

size_t zbh_compress_bound(size_t n) { return zb::compress_bound(n); }
#ifdef ZB_STATS
void zbh_parse_stats(unsigned long long* out, int reset) {
    memcpy(out, &zb::g_parseStats, sizeof zb::g_parseStats);
    if (reset) memset(&zb::g_parseStats, 0, sizeof zb::g_parseStats);
}
#endif

size_t zbh_compress_flags(void* dst, size_t dstCapacity, const void* src, size_t srcSize, int level, unsigned flags);
size_t zbh_compress(void* dst, size_t dstCapacity, const void* src, size_t srcSize, int level) { return zbh_compress_flags(dst, dstCapacity, src, srcSize, level, 0); }
size_t zbh_compress_flags(void* dst, size_t dstCapacity, const void* src, size_t srcSize, int level, unsigned flags) {
    using namespace zb;
    if (srcSize > BLOCKSIZE_MAX) return ERR(E_srcSize_wrong);
    WarpHost w;
    EncShared* S = (EncShared*)calloc(1, sizeof(EncShared));
    u8* wk = (u8*)calloc(1, enc_work_bytes() + 64);
    EncWork W = enc_work_carve(wk);
    size_t const bound = compress_bound(srcSize);
    u8* slot = (u8*)calloc(1, bound + 64);
    // inputs are read with aligned 8-byte loads: give the copy slack on both sides
    u8* in = (u8*)calloc(1, srcSize + 64);
    memcpy(in + 16, src, srcSize);
    size_t r = compress_frame(w, *S, W, slot, bound < 18 ? 18 : bound, in + 16, srcSize, level, flags);
    if (!isErr(r)) { if (r > dstCapacity) r = ERR(E_dstSize_tooSmall); else memcpy(dst, slot, r); }
    free(S); free(wk); free(slot); free(in);
    return r;
}

// explicit compression parameters (ZSTD_c_windowLog ... ZSTD_c_strategy; 0 = from the level): ov7 = {windowLog, chainLog, hashLog,
// searchLog, minMatch, targetLength, strategy}.  emu != 0: both stages on the 32-lane emulator.
size_t zbh_compress_params(void* dst, size_t dstCapacity, const void* src, size_t srcSize, int level, unsigned flags, const unsigned* ov7, int emu) {
    using namespace zb;
    if (srcSize > BLOCKSIZE_MAX) return ERR(E_srcSize_wrong);
    CParams ov = { ov7[0], ov7[1], ov7[2], ov7[3], ov7[4], ov7[5], ov7[6] };
    EncShared* S = (EncShared*)calloc(1, sizeof(EncShared));
    u8* wk = (u8*)calloc(1, enc_work_bytes() + 64);
    EncWork W = enc_work_carve(wk);
    size_t const bound = compress_bound(srcSize);
    u8* slot = (u8*)calloc(1, bound + 64);
    u8* in = (u8*)calloc(1, srcSize + 64);
    memcpy(in + 16, src, srcSize);
    size_t r;
    if (emu) {
        size_t results[32]; u32 nbSeqs[32], lastLLs[32];
        run_warp<32>([&](const WarpEmuT<32>& w) { results[w.lane] = parse_stage(w, W, in + 16, srcSize, level, &nbSeqs[w.lane], &lastLLs[w.lane], &ov); });
        r = results[0];
        for (int i = 1; i < 32; i++) if (results[i] != r || nbSeqs[i] != nbSeqs[0] || lastLLs[i] != lastLLs[0]) r = ERR(E_GENERIC);
        if (!isErr(r)) {
            run_warp<32>([&](const WarpEmuT<32>& w) { results[w.lane] = encode_stage(w, *S, W, slot, bound < 18 ? 18 : bound, in + 16, srcSize, level, nbSeqs[0], lastLLs[0], flags, &ov); });
            r = results[0];
            for (int i = 1; i < 32; i++) if (results[i] != r) r = ERR(E_GENERIC);
        }
    } else {
        WarpHost w;
        r = compress_frame(w, *S, W, slot, bound < 18 ? 18 : bound, in + 16, srcSize, level, flags, &ov);
    }
    if (!isErr(r)) { if (r > dstCapacity) r = ERR(E_dstSize_tooSmall); else memcpy(dst, slot, r); }
    free(S); free(wk); free(slot); free(in);
    return r;
}

size_t zbh_decompress_format(void* dst, size_t dstCapacity, const void* src, size_t srcSize, unsigned magicless);
size_t zbh_decompress(void* dst, size_t dstCapacity, const void* src, size_t srcSize) { return zbh_decompress_format(dst, dstCapacity, src, srcSize, 0); }
size_t zbh_decompress_format(void* dst, size_t dstCapacity, const void* src, size_t srcSize, unsigned magicless) {
    using namespace zb;
    WarpHost w;
    DecShared* S = (DecShared*)calloc(1, sizeof(DecShared));
    u8* scratch = (u8*)calloc(1, BLOCKSIZE_MAX + 64);
    u8* in = (u8*)calloc(1, srcSize + 64);
    memcpy(in + 16, src, srcSize);
    u8* out = (u8*)calloc(1, dstCapacity + 64);
    size_t const r = decompress_item(w, *S, in + 16, srcSize, out + 16, dstCapacity, scratch, magicless);
    if (!isErr(r)) memcpy(dst, out + 16, r);
    free(S); free(scratch); free(in); free(out);
    return r;
}

// ---- the same entry points on an emulated 32-lane warp (simt_emu.h): exercises the cooperative code paths
size_t zbe_compress(void* dst, size_t dstCapacity, const void* src, size_t srcSize, int level) {
    using namespace zb;
    if (srcSize > BLOCKSIZE_MAX) return ERR(E_srcSize_wrong);
    EncShared* S = (EncShared*)calloc(1, sizeof(EncShared));
    u8* wk = (u8*)calloc(1, enc_work_bytes() + 64);
    EncWork W = enc_work_carve(wk);
    size_t const bound = compress_bound(srcSize);
    u8* slot = (u8*)calloc(1, bound + 64);
    u8* in = (u8*)calloc(1, srcSize + 64);
    memcpy(in + 16, src, srcSize);
    // stage 1 on an 8-lane group (k_parse), stage 2 on a 32-lane warp (k_entropy), like the CUDA build
    size_t results[32]; u32 nbSeqs[32], lastLLs[32];
    const char* pl = getenv("ZB_EMU_PARSE_LANES");
    int const lanes = pl ? atoi(pl) : 32;
    if (lanes == 8) run_warp<8>([&](const WarpEmuT<8>& w) { results[w.lane] = parse_stage(w, W, in + 16, srcSize, level, &nbSeqs[w.lane], &lastLLs[w.lane]); });
    else run_warp<32>([&](const WarpEmuT<32>& w) { results[w.lane] = parse_stage(w, W, in + 16, srcSize, level, &nbSeqs[w.lane], &lastLLs[w.lane]); });
    size_t r = results[0];
    for (int i = 1; i < (lanes == 8 ? 8 : 32); i++) if (results[i] != r || nbSeqs[i] != nbSeqs[0] || lastLLs[i] != lastLLs[0]) r = ERR(E_GENERIC);
    if (!isErr(r)) {
        run_warp<32>([&](const WarpEmuT<32>& w) { results[w.lane] = encode_stage(w, *S, W, slot, bound < 18 ? 18 : bound, in + 16, srcSize, level, nbSeqs[0], lastLLs[0]); });
        r = results[0];
        for (int i = 1; i < 32; i++) if (results[i] != r) r = ERR(E_GENERIC);   // the return value must be warp-uniform
    }
    if (!isErr(r)) { if (r > dstCapacity) r = ERR(E_dstSize_tooSmall); else memcpy(dst, slot, r); }
    free(S); free(wk); free(slot); free(in);
    return r;
}

size_t zbe_decompress(void* dst, size_t dstCapacity, const void* src, size_t srcSize) {
    using namespace zb;
    DecShared* S = (DecShared*)calloc(1, sizeof(DecShared));
    u8* scratch = (u8*)calloc(1, BLOCKSIZE_MAX + 64);
    u8* in = (u8*)calloc(1, srcSize + 64);
    memcpy(in + 16, src, srcSize);
    u8* out = (u8*)calloc(1, dstCapacity + 64);
    size_t results[32];
    run_warp<32>([&](const WarpEmuT<32>& w) { results[w.lane] = decompress_item(w, *S, in + 16, srcSize, out + 16, dstCapacity, scratch); });
    size_t r = results[0];
    for (int i = 1; i < 32; i++) if (results[i] != r) r = ERR(E_GENERIC);
    if (!isErr(r)) memcpy(dst, out + 16, r);
    free(S); free(scratch); free(in); free(out);
    return r;
}

// ---- the staged batch decoder (zb_decode_fast.cuh) on one item; emu != 0 runs stages A and D on the 32-lane emulator
size_t zbp_decompress_at(void* dst, size_t dstCapacity, const void* src, size_t srcSize, int emu, int mis);
size_t zbp_decompress(void* dst, size_t dstCapacity, const void* src, size_t srcSize, int emu) { return zbp_decompress_at(dst, dstCapacity, src, srcSize, emu, 0); }
// mis = 0..15: the regenerated bytes start that far off a 16-byte boundary (a batch packs its outputs back to back)
size_t zbp_decompress_at(void* dst, size_t dstCapacity, const void* src, size_t srcSize, int emu, int mis) {
    using namespace zb;
    DecShared* S = (DecShared*)calloc(1, sizeof(DecShared));
    u8* in = (u8*)calloc(1, srcSize + 64);
    memcpy(in + 16, src, srcSize);
    u8* out = (u8*)aligned_alloc(16, (dstCapacity + 96) / 16 * 16); memset(out, 0, (dstCapacity + 96) / 16 * 16);
    u8* lit = (u8*)calloc(1, BLOCKSIZE_MAX + 64);
    u16* huf = (u16*)calloc(FAST_HUF_ENTRIES, 2);
    u32* fse = (u32*)calloc(FAST_FSE_ENTRIES, 4);
    ExecShared* X = (ExecShared*)aligned_alloc(16, (sizeof(ExecShared) + 15) / 16 * 16);
    u64* seqs = (u64*)calloc(FAST_MAXS + 8, 8);
    DecDesc d;
    size_t r;
    if (emu) run_warp<32>([&](const WarpEmuT<32>& w) { dec_prepare(w, *S, in + 16, srcSize, dstCapacity, &d, huf, fse); });
    else { WarpHost w; dec_prepare(w, *S, in + 16, srcSize, dstCapacity, &d, huf, fse); }
    if (d.mode == 0) {
        u8* scratch = (u8*)calloc(1, BLOCKSIZE_MAX + 64);
        WarpHost w; r = decompress_item(w, *S, in + 16, srcSize, out + 16 + mis, dstCapacity, scratch);
        free(scratch);
    } else {
        const u8* blk = in + 16 + d.blockOff;
        for (int k = 0; k < 4; k++) dec_huf(&d, k, blk, huf, lit);
        dec_seq(&d, blk, fse, &h_tables, seqs);
        if (getenv("ZB_DEBUG_STAGES")) fprintf(stderr, "mode %u stA1 %u stB %u stA2 %u stC %u nbSeq %u litMode %u litSize %u hufLog %u nStreams %u seqBits %u logs %u %u %u\n", d.mode, d.stA1, d.stB, d.stA2, d.stC, d.nbSeq, d.litMode, d.litSize, d.hufLog, d.nStreams, d.seqBits, d.logLL, d.logOF, d.logML);
        if (emu) {
            size_t results[32];
            run_warp<32>([&](const WarpEmuT<32>& w) { results[w.lane] = dec_exec(w, *X, &d, in + 16, lit, seqs, out + 16 + mis, dstCapacity); });
            r = results[0];
            for (int i = 1; i < 32; i++) if (results[i] != r) r = ERR(E_GENERIC);
        } else { WarpHost w; r = dec_exec(w, *X, &d, in + 16, lit, seqs, out + 16 + mis, dstCapacity); }
    }
    if (!isErr(r)) memcpy(dst, out + 16 + mis, r);
    free(S); free(in); free(out); free(lit); free(huf); free(fse); free(seqs); free(X);
    return r;
}

// ---- sequence export (ZSTD_Sequence records of one block): parse + export_sequences on 1 lane (emu = 0) or on the emulator
size_t zbh_generate_sequences(void* outSeqs, size_t outCapacity, const void* src, size_t srcSize, int level, int emu) {
    using namespace zb;
    if (srcSize > BLOCKSIZE_MAX) return ERR(E_srcSize_wrong);
    u8* wk = (u8*)calloc(1, enc_work_bytes() + 64);
    EncWork W = enc_work_carve(wk);
    u8* in = (u8*)calloc(1, srcSize + 64);
    memcpy(in + 16, src, srcSize);
    ZSeq* out = (ZSeq*)aligned_alloc(16, sizeof(ZSeq) * (MAX_SEQ + 2));
    size_t r; u32 n = 0;
    if (emu) {
        size_t results[32]; u32 nbSeqs[32], lastLLs[32], counts[32];
        run_warp<32>([&](const WarpEmuT<32>& w) { results[w.lane] = parse_stage(w, W, in + 16, srcSize, level, &nbSeqs[w.lane], &lastLLs[w.lane]); });
        r = results[0];
        if (!isErr(r)) {
            run_warp<32>([&](const WarpEmuT<32>& w) { counts[w.lane] = export_sequences(w, W.seq, nbSeqs[0], lastLLs[0], srcSize, out); });
            n = counts[0];
            for (int i = 1; i < 32; i++) if (counts[i] != n) r = ERR(E_GENERIC);
        }
    } else {
        WarpHost w; u32 nbSeq = 0, lastLL = 0;
        r = parse_stage(w, W, in + 16, srcSize, level, &nbSeq, &lastLL);
        if (!isErr(r)) n = export_sequences(w, W.seq, nbSeq, lastLL, srcSize, out);
    }
    if (!isErr(r)) { if (n > outCapacity) r = ERR(E_dstSize_tooSmall); else { memcpy(outSeqs, out, sizeof(ZSeq) * n); r = n; } }
    free(wk); free(in); free(out);
    return r;
}

size_t zbh_sizeof_dec_shared() { return sizeof(zb::DecShared); }
size_t zbh_sizeof_enc_shared() { return sizeof(zb::EncShared); }
}
