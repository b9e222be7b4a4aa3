// canary_fuzz.cpp -- TEST-ONLY memory-safety fuzz of the decoder source (zstd_jni_b200/csrc/*.cuh instantiated on the host).
// Frames of several levels / flag sets are corrupted (byte flips, truncation) or left intact and decoded through the fused decoder (1 lane and
// the 32-lane emulator) and the staged decoder into a buffer fenced by canaries: nothing may be written outside [dst, dst + capacity).
// Built and run briefly by tests/test_hostsim.py; for a deeper run with AddressSanitizer:
//   g++ -O1 -g -fsanitize=address -std=c++17 tests/hostsim/canary_fuzz.cpp -o /tmp/canary && ASAN_OPTIONS=detect_leaks=0 /tmp/canary <seed> <iterations>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <random>
#include "../../zstd_jni_b200/csrc/zb_decode.cuh"
#include "../../zstd_jni_b200/csrc/zb_decode_fast.cuh"
#include "../../zstd_jni_b200/csrc/zb_encode.cuh"
#include "simt_emu.h"
using namespace zb;
static std::vector<u8> compress(const std::vector<u8>& src, int level, u32 flags) {
    WarpHost w; EncShared* S = (EncShared*)calloc(1, sizeof(EncShared)); u8* wk = (u8*)calloc(1, enc_work_bytes() + 64); EncWork W = enc_work_carve(wk);
    size_t bound = compress_bound(src.size()); if (bound < 18) bound = 18;
    std::vector<u8> slot(bound + 64), in(src.size() + 64);
    memcpy(in.data() + 16, src.data(), src.size());
    size_t r = compress_frame(w, *S, W, slot.data(), bound, in.data() + 16, src.size(), level, flags);
    free(S); free(wk);
    if (isErr(r)) { printf("compress error\n"); exit(1); }
    slot.resize(r); return slot;
}
int main(int argc, char** argv) {
    unsigned seed = argc > 1 ? atoi(argv[1]) : 1; int iters = argc > 2 ? atoi(argv[2]) : 2000;
    std::mt19937 rng(seed);
    // inputs: synthetic mixes
    std::vector<std::vector<u8>> srcs;
    for (int k = 0; k < 6; k++) {
        size_t n = k < 3 ? 131072 : 1 + rng() % 50000; std::vector<u8> d(n);
        for (size_t i = 0; i < n; i++) {
            if (k % 3 == 0) d[i] = (u8)("the quick brown fox jumps over the lazy dog "[(i * 7 + (i >> 9)) % 44] ^ ((rng() % 37 == 0) ? 1 : 0));
            else if (k % 3 == 1) d[i] = (u8)((i % 64 < 48) ? (i * 13) : rng());
            else d[i] = (u8)(rng() % 4 + ((i >> 12) & 1) * 60);
        }
        srcs.push_back(d);
    }
    struct F { std::vector<u8> z; size_t cap; u32 ml; };
    std::vector<F> frames;
    for (auto& d : srcs) for (int lvl : {1, 3, 6, 9}) for (u32 fl : {0u, 1u, 2u, 4u}) frames.push_back({compress(d, lvl, fl), d.size(), (fl & 4) ? 1u : 0u});
    // a two-frame item
    { F a = frames[0]; a.z.insert(a.z.end(), frames[1].z.begin(), frames[1].z.end()); a.cap = frames[0].cap + frames[1].cap; frames.push_back(a); }
    DecShared* S = (DecShared*)calloc(1, sizeof(DecShared));
    u8* scratch = (u8*)calloc(1, BLOCKSIZE_MAX + 64);
    u8* lit = (u8*)calloc(1, BLOCKSIZE_MAX + 64); u16* huf = (u16*)calloc(FAST_HUF_ENTRIES, 2); u32* fse = (u32*)calloc(FAST_FSE_ENTRIES, 4); u64* seqs = (u64*)calloc(FAST_MAXS + 8, 8);
    ExecShared* X = (ExecShared*)aligned_alloc(16, (sizeof(ExecShared) + 15) / 16 * 16);
    long bad = 0, ok = 0, errs = 0;
    for (int it = 0; it < iters; it++) {
        F const& f = frames[rng() % frames.size()];
        std::vector<u8> b = f.z;
        int mode = rng() % 10;
        if (mode < 6) { int k = 1 + rng() % 3; for (int j = 0; j < k; j++) b[rng() % b.size()] ^= (u8)(1 + rng() % 255); }
        else if (mode < 8) b.resize(rng() % b.size());
        // mode 8,9: valid
        size_t cap = (rng() % 5 == 0) ? rng() % (f.cap + 1) : f.cap;
        u8* in = (u8*)malloc(b.size() + 64); memset(in, 0xEE, b.size() + 64); memcpy(in + 16, b.data(), b.size());
        size_t const PRE = 64, POST = 64;
        u8* out = (u8*)malloc(PRE + cap + POST);
        for (int variant = 0; variant < 3; variant++) {
            memset(out, 0xA5, PRE + cap + POST);
            size_t r;
            if (variant == 0) { WarpHost w; r = decompress_item(w, *S, in + 16, b.size(), out + PRE, cap, scratch, f.ml); }
            else if (variant == 1) {
                size_t results[32];
                run_warp<32>([&](const WarpEmuT<32>& w) { results[w.lane] = decompress_item(w, *S, in + 16, b.size(), out + PRE, cap, scratch, f.ml); });
                r = results[0];
            } else {
                if (f.ml) continue;
                DecDesc d; WarpHost w;
                dec_prepare(w, *S, in + 16, b.size(), cap, &d, huf, fse);
                if (d.mode == 0) continue;
                const u8* blk = in + 16 + d.blockOff;
                for (int k = 0; k < 4; k++) dec_huf(&d, k, blk, huf, lit);
                dec_seq(&d, blk, fse, &h_tables, seqs);
                size_t results[32];
                run_warp<32>([&](const WarpEmuT<32>& w2) { results[w2.lane] = dec_exec(w2, *X, &d, in + 16, lit, seqs, out + PRE, cap); });
                r = results[0];
            }
            for (size_t i = 0; i < PRE; i++) if (out[i] != 0xA5) { bad++; printf("seed %u it %d variant %d: write BEFORE dst at -%zu\n", seed, it, variant, PRE - i); break; }
            for (size_t i = 0; i < POST; i++) if (out[PRE + cap + i] != 0xA5) { bad++; printf("seed %u it %d variant %d: write AFTER dst+cap at +%zu (cap %zu, r %zx)\n", seed, it, variant, i, cap, r); break; }
            if (isErr(r)) errs++; else ok++;
        }
        free(in); free(out);
    }
    printf("seed %u: %d iterations, ok %ld, errors %ld, canary violations %ld\n", seed, iters, ok, errs, bad);
    return bad != 0;
}
