// zb_decode_chains.cuh -- device side of stages B + C of the staged decoder (zb_decode_fast.cuh): k_dec_chains.
//
// One persistent CTA per SM, four warps, each on its own scheduler:
//   warps 0 .. CH_FSE_WARPS-1        sequence chains: CH_FSE_LANES lanes per warp, one frame per lane (SeqChain)
//   warps CH_FSE_WARPS .. CH_WARPS-1 Huffman chains: 8 frames x 4 streams per warp (HufChain)
// Lanes are persistent: a lane (a group of 4 lanes for Huffman) that finishes its frame draws the next one from a
// longest-first work list while the other lanes keep stepping, so the kernel lasts (total steps / lanes) or as long
// as its longest chain, whichever is more -- not (waves x longest chain).
//
// Shared memory (~221 KB, the whole SM):
//   * per lane / frame slot the decode tables (3 FSE tables = 5 KB, one Huffman table = 4 KB), filled by bulk
//     asynchronous copies (cp.async.bulk global -> shared, completion on an mbarrier the owning lanes wait on);
//   * per lane a ring of CH_RING_GROUPS 16-byte cells of its bitstream: cp.async copies keep CH_RING_DEPTH cells
//     in flight below the reader, so the sequential, read-once compressed bytes come out of HBM hundreds of
//     cycles before the chain needs them and no chain ever waits for DRAM;
//   * one copy of the code tables (base values of the length codes).
#pragma once
#include "zb_decode_fast.cuh"

#if defined(__CUDACC__)
namespace zb {

constexpr int CH_FSE_WARPS = 2, CH_FSE_LANES = 14, CH_HUF_WARPS = 2;
constexpr int CH_WARPS = CH_FSE_WARPS + CH_HUF_WARPS;
constexpr int CH_RING_GROUPS = 8;        // 16-byte cells per lane ring
constexpr int CH_RING_DEPTH = 4;         // cells requested below the one being read
constexpr u32 CH_FSE_SLOT = FAST_FSE_ENTRIES * 4, CH_HUF_SLOT = FAST_HUF_ENTRIES * 2;
constexpr u32 CH_OFF_HUF = CH_FSE_WARPS * CH_FSE_LANES * CH_FSE_SLOT;
constexpr u32 CH_OFF_RING = CH_OFF_HUF + CH_HUF_WARPS * 8 * CH_HUF_SLOT;
constexpr u32 CH_OFF_CT = CH_OFF_RING + CH_WARPS * 32 * (16 + CH_RING_GROUPS * 16);
constexpr u32 CH_OFF_BAR = CH_OFF_CT + ((sizeof(CodeTables) + 15) / 16) * 16;
constexpr u32 CH_SMEM = CH_OFF_BAR + (CH_FSE_WARPS * CH_FSE_LANES + CH_HUF_WARPS * 8) * 8;

__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(u32 dst, const void* src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(u32 bar, u32 parity) {
    u32 ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(u32 dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
// wait for a table copy; gives up after ~2^22 polls (a copy that never lands must not hang the device: the caller
// reports corruption for the frame instead)
__device__ __forceinline__ bool mbar_wait(u32 bar, u32 parity) {
    for (u32 spin = 0; spin < (1u << 22); spin++) if (mbar_try_wait(bar, parity)) return true;
    return false;
}

// word source of a chain on the GPU: the lane's ring.  32 words (8 groups of 16 bytes) plus a mirror cell below them that
// repeats the top group, so that the four words k, k-1, k-2, k-3 of a window are always at one address and three immediate
// offsets below it -- no wrap-around arithmetic on the chain.
constexpr u32 CH_RING_LANE_BYTES = 16 + CH_RING_GROUPS * 16;       // mirror cell + ring
constexpr u32 CH_RING_WARP_BYTES = 32 * CH_RING_LANE_BYTES;
struct RingWords {
    StreamGeom g;
    u32 ring0;           // shared-memory address of the lane's ring word 0 (the mirror cell sits 16 bytes below)
    int gIssued;         // lowest 16-byte group requested so far
    __device__ __forceinline__ u32 lds(u32 addr) const { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }
    __device__ __forceinline__ u32 raw(int k) const { return lds(ring0 + (((u32)k & (CH_RING_GROUPS * 4 - 1)) << 2)); }
    __device__ __forceinline__ u32 word(int k) const { return k < g.kFirst ? 0u : g.fix(k, raw(k)); }
    __device__ __forceinline__ void fetch4(int k, u32& a, u32& b, u32& c, u32& d) const {
        if (k >= 8) {                                    // kFirst <= 3: no masks up here
            u32 const at = ring0 + (((u32)k & (CH_RING_GROUPS * 4 - 1)) << 2);
            asm volatile("ld.shared.u32 %0, [%4];\n\tld.shared.u32 %1, [%4+-4];\n\tld.shared.u32 %2, [%4+-8];\n\tld.shared.u32 %3, [%4+-12];"
                         : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(at) : "memory");
        } else { a = word(k); b = word(k - 1); c = word(k - 2); d = word(k - 3); }
    }
    __device__ __forceinline__ void fetch4_fast(int k, u32& a, u32& b, u32& c, u32& d) const {
        u32 const at = ring0 + (((u32)k & (CH_RING_GROUPS * 4 - 1)) << 2);
        asm volatile("ld.shared.u32 %0, [%4];\n\tld.shared.u32 %1, [%4+-4];\n\tld.shared.u32 %2, [%4+-8];\n\tld.shared.u32 %3, [%4+-12];"
                     : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(at) : "memory");
    }
    // advance() without a branch: the copies are predicated, a (possibly empty) group is committed every time -- empty groups only
    // make the real requests look older to wait_group, which keeps its guarantee
    __device__ __forceinline__ void advance_fast(int k) {
        bool const need = gIssued > (k >> 2) - CH_RING_DEPTH;
        gIssued -= need ? 1 : 0;
        u32 const m = (u32)gIssued & (CH_RING_GROUPS - 1);
        u32 const p1 = (need && gIssued >= 0) ? 1u : 0u, p2 = (p1 && m == CH_RING_GROUPS - 1) ? 1u : 0u;
        asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %3, 0;\n\tsetp.ne.u32 q, %4, 0;\n\t"
                     "@p cp.async.cg.shared.global [%0], [%2], 16;\n\t@q cp.async.cg.shared.global [%1], [%2], 16;\n\t"
                     "cp.async.commit_group;\n\t}"
                     :: "r"(ring0 + m * 16), "r"(ring0 - 16), "l"(g.W + 4 * gIssued), "r"(p1), "r"(p2) : "memory");
        cp_async_wait<CH_RING_DEPTH - 1>();
    }
    __device__ __forceinline__ void request_next() {    // one more group below the lowest requested one
        gIssued--;
        if (gIssued >= 0) {
            u32 const m = (u32)gIssued & (CH_RING_GROUPS - 1);
            cp_async16(ring0 + m * 16, g.W + 4 * gIssued);
            if (m == CH_RING_GROUPS - 1) cp_async16(ring0 - 16, g.W + 4 * gIssued);      // the top group also fills the mirror cell
        }
        cp_async_commit();
    }
    // the reader now stands at word k (at most one group below where it stood): keep CH_RING_DEPTH groups requested below
    // it, then make sure the group of k and the one below have landed (all but the CH_RING_DEPTH - 1 youngest requests)
    __device__ __forceinline__ void advance(int k) {
        if (gIssued > (k >> 2) - CH_RING_DEPTH) request_next();
        cp_async_wait<CH_RING_DEPTH - 1>();
    }
    __device__ __forceinline__ void start(const u8* ip, u32 bits) {
        g.set(ip);
        int const k = ((int)(g.floorBit + bits) - 1) >> 5;
        gIssued = (k >> 2) + 1;
        while (gIssued > (k >> 2) - CH_RING_DEPTH) request_next();
        cp_async_wait<CH_RING_DEPTH - 1>();
    }
};

struct ChainsArgs {
    const u8* srcBase; const u64* srcOff; u32 n; DecDesc* descs;
    const u32* fseBase; const u16* hufBase; u64* seqBase; u8* litBase; size_t litStride;
    u32* counterSeq; u32* counterHuf; const u32* orderSeq; const u32* orderHuf;
    u32 roles;           // experiments: bit 0 = run the sequence chains, bit 1 = run the Huffman chains (3 = both, the product setting)
};

__device__ __forceinline__ void chains_fse_warp(const ChainsArgs& a, unsigned char* smem, int warp, int lane) {
    if (lane >= CH_FSE_LANES) return;
    constexpr u32 MASK = (1u << CH_FSE_LANES) - 1;
    int const slot = warp * CH_FSE_LANES + lane;
    u32* const tab = reinterpret_cast<u32*>(smem + (size_t)slot * CH_FSE_SLOT);
    u32 const tabAddr = smem_u32(tab);
    const CodeTables* const ct = reinterpret_cast<const CodeTables*>(smem + CH_OFF_CT);
    u32 const bar = smem_u32(smem + CH_OFF_BAR + slot * 8);
    u32 parity = 0;
    RingWords ws; ws.ring0 = smem_u32(smem + CH_OFF_RING + warp * CH_RING_WARP_BYTES + lane * CH_RING_LANE_BYTES + 16); ws.gIssued = 0;
    SeqChain D; D.k = D.nbSeq = 0; D.top = 0; D.sLL = D.sOF = D.sML = 0; D.rep0 = D.rep1 = D.rep2 = 1; D.out = nullptr;
    DecDesc* d = nullptr;
    bool live = false, exhausted = false;
    for (;;) {
        // the common iteration: every lane in the middle of a frame -- one vote, then a straight line
        if (!__any_sync(MASK, !live || !D.plain())) { D.step_fast(ws, tab, tab + FAST_FSE_OF, tab + FAST_FSE_ML, ct); continue; }
        if (!live && !exhausted) {
            u32 item = atomicAdd(a.counterSeq, 1u);
            if (item >= a.n) exhausted = true;
            else {
                item = a.orderSeq[item];
                d = a.descs + item;
                u32 const nbSeq = d->nbSeq;
                if (nbSeq == 0) exhausted = true;            // longest first: only frames without sequences from here on
                else if (d->mode == 1 && !d->stA1 && !d->stA2 && !d->seqUnusable) {
                    u32 const logLL = d->logLL, logOF = d->logOF, logML = d->logML;
                    u32 const bLL = umax(16u, 4u << logLL), bOF = umax(16u, 4u << logOF), bML = umax(16u, 4u << logML);
                    const u32* const gt = a.fseBase + (size_t)item * FAST_FSE_ENTRIES;
                    fence_proxy_async();                     // the slot's previous tables were read through the generic proxy
                    mbar_expect_tx(bar, bLL + bOF + bML);
                    bulk_g2s(tabAddr, gt, bLL, bar);
                    bulk_g2s(tabAddr + FAST_FSE_OF * 4, gt + FAST_FSE_OF, bOF, bar);
                    bulk_g2s(tabAddr + FAST_FSE_ML * 4, gt + FAST_FSE_ML, bML, bar);
                    ws.start(a.srcBase + a.srcOff[item] + d->blockOff + d->seqOff, d->seqBits);
                    bool const landed = mbar_wait(bar, parity);
                    parity ^= 1;
                    if (!landed) { d->stC = E_GENERIC; exhausted = true; }
                    else {
                        D.begin(ws, ws.g.floorBit, d->seqBits, logLL, logOF, logML, nbSeq, a.seqBase + (size_t)item * FAST_MAXS);
                        live = true;
                    }
                }
            }
        }
        if (!__any_sync(MASK, live || !exhausted)) break;
        if (live) {
            D.step(ws, tab, tab + FAST_FSE_OF, tab + FAST_FSE_ML, ct);
            if (!D.more()) { if (!D.clean(ws.g.floorBit)) d->stC = E_corruption_detected; live = false; }
        }
    }
}

__device__ __forceinline__ void chains_huf_warp(const ChainsArgs& a, unsigned char* smem, int hwarp, int warp, int lane) {
    int const grp = lane >> 2, k = lane & 3;
    u32 const gmask = 0xFu << (lane & ~3);
    int const slot = hwarp * 8 + grp;
    u16* const tab = reinterpret_cast<u16*>(smem + CH_OFF_HUF + (size_t)slot * CH_HUF_SLOT);
    u32 const tabAddr = smem_u32(tab);
    u32 const bar = smem_u32(smem + CH_OFF_BAR + (CH_FSE_WARPS * CH_FSE_LANES + slot) * 8);
    u32 parity = 0;
    RingWords ws; ws.ring0 = smem_u32(smem + CH_OFF_RING + warp * CH_RING_WARP_BYTES + lane * CH_RING_LANE_BYTES + 16); ws.gIssued = 0;
    HufChain H; H.left = 0; H.op = nullptr; H.kNext = 0; H.hi = H.lo = 0; H.avail = 0; H.budget = 0;
    DecDesc* d = nullptr;
    u32 sh = 0;
    bool live = false, exhausted = false;
    for (;;) {
        if (!__any_sync(0xFFFFFFFFu, !live || !H.plain())) { H.step4_fast(ws, tab, sh); continue; }
        u32 const liveMask = __ballot_sync(0xFFFFFFFFu, live);
        u32 const busyMask = __ballot_sync(0xFFFFFFFFu, live || !exhausted);
        if (!busyMask) break;
        if ((liveMask & gmask) == 0 && !exhausted) {         // the group's four streams are done: next frame
            u32 item = 0xFFFFFFFFu;
            if (k == 0) { item = atomicAdd(a.counterHuf, 1u); item = item < a.n ? a.orderHuf[item] : 0xFFFFFFFFu; }
            item = __shfl_sync(gmask, item, lane & ~3);
            if (item == 0xFFFFFFFFu) exhausted = true;
            else {
                d = a.descs + item;
                if (d->hufLitSize == 0) exhausted = true;    // longest first: no Huffman-coded literals from here on
                else if (d->mode == 1 && !d->stA1 && d->litMode == 2) {
                    u32 const log = d->hufLog;
                    __syncwarp(gmask);                         // the siblings' last reads of the slot's old table come first
                    if (k == 0) {
                        fence_proxy_async();
                        u32 const bytes = umax(16u, 2u << log);
                        mbar_expect_tx(bar, bytes);
                        bulk_g2s(tabAddr, a.hufBase + (size_t)item * FAST_HUF_ENTRIES, bytes, bar);
                    }
                    u32 const bits = k < (int)d->nStreams ? d->sBits[k] : HUF_UNUSABLE;
                    const u8* const blk = a.srcBase + a.srcOff[item] + d->blockOff;
                    if (bits != HUF_UNUSABLE) ws.start(blk + d->sOff[k], bits);
                    bool const landed = mbar_wait(bar, parity);
                    parity ^= 1;
                    if (!landed) { d->stB = E_GENERIC; exhausted = true; }
                    else if (bits != HUF_UNUSABLE) {
                        H.begin(ws, ws.g.floorBit, bits, a.litBase + (size_t)item * a.litStride + d->oOff[k], d->oCnt[k]);
                        sh = 32 - log;
                        live = H.left != 0;
                        if (!live && !H.clean()) d->stB = E_corruption_detected;
                    }
                }
            }
        }
        if (live) {
            if (H.left >= 4 && H.aligned4()) H.step4(ws, tab, sh); else H.step1(ws, tab, sh);
            if (H.left == 0) { if (!H.clean()) d->stB = E_corruption_detected; live = false; }
        }
    }
}

__global__ void __launch_bounds__(CH_WARPS * 32, 1) k_dec_chains(ChainsArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {   const u32* const srcw = reinterpret_cast<const u32*>(&c_tables); u32* const dstw = reinterpret_cast<u32*>(smem + CH_OFF_CT);
        for (u32 j = threadIdx.x; j < sizeof(CodeTables) / 4; j += blockDim.x) dstw[j] = srcw[j]; }
    if (threadIdx.x == 0) {
        for (int j = 0; j < CH_FSE_WARPS * CH_FSE_LANES + CH_HUF_WARPS * 8; j++) mbar_init(smem_u32(smem + CH_OFF_BAR + j * 8), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp < CH_FSE_WARPS) { if (a.roles & 1) chains_fse_warp(a, smem, warp, lane); }
    else if (a.roles & 2) chains_huf_warp(a, smem, warp - CH_FSE_WARPS, warp, lane);
}

}  // namespace zb
#endif
