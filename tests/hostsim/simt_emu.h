// simt_emu.h -- TEST-ONLY 32-lane warp emulator for the codec templates.
//
// Each lane is a ucontext fiber; lanes run round-robin and switch only inside warp collectives
// (sync / bcast / shfl / ballot / match_any / reductions), which rendezvous all 32 lanes exactly like
// the *_sync intrinsics with a full mask.  Because a lane runs ahead until its next collective, data that
// another lane has not yet produced is really missing -- a forgotten __syncwarp() shows up here as a wrong
// result instead of being hidden by lock-step execution.
#pragma once
#include <ucontext.h>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

namespace zb {

struct WarpEmuShared {
    uint64_t slot[2][32];
    uint64_t arrived[2] = {0, 0};
    ucontext_t main_ctx;
    ucontext_t lane_ctx[32];
    bool done[32];
    int current = 0;
};

template <int LANES>
struct WarpEmuT {
    int lane = 0;
    static constexpr int W = LANES;
    static constexpr uint32_t FULL = (LANES >= 32) ? 0xFFFFFFFFu : ((1u << (LANES & 31)) - 1);
    WarpEmuShared* sh = nullptr;
    mutable uint64_t uses[2] = {0, 0};
    mutable unsigned seq = 0;

    void yield() const { swapcontext(&sh->lane_ctx[lane], &sh->main_ctx); }
    // all-lanes exchange of one 64-bit value; returns the buffer holding every lane's contribution
    const uint64_t* exchange(uint64_t v) const {
        int const b = (int)(seq++ & 1);
        sh->slot[b][lane] = v;
        sh->arrived[b]++;
        uses[b]++;
        while (sh->arrived[b] < (uint64_t)LANES * uses[b]) yield();
        return sh->slot[b];
    }
    void sync() const { (void)exchange(0); }
    template <class T> T shfl(T v, int src) const {
        uint64_t raw = 0; static_assert(sizeof(T) <= 8, "shfl payload"); memcpy(&raw, &v, sizeof(T));
        const uint64_t* all = exchange(raw);
        T out; memcpy(&out, &all[src & (LANES - 1)], sizeof(T)); return out;
    }
    template <class T> T bcast(T v, int src = 0) const { return shfl(v, src); }
    uint32_t ballot(bool p) const { const uint64_t* all = exchange(p ? 1 : 0); uint32_t m = 0; for (int i = 0; i < LANES; i++) m |= (uint32_t)(all[i] & 1) << i; return m; }
    uint32_t sum(uint32_t v) const { const uint64_t* all = exchange(v); uint32_t s = 0; for (int i = 0; i < LANES; i++) s += (uint32_t)all[i]; return s; }
    uint32_t max(uint32_t v) const { const uint64_t* all = exchange(v); uint32_t s = 0; for (int i = 0; i < LANES; i++) if ((uint32_t)all[i] > s) s = (uint32_t)all[i]; return s; }
    uint32_t match_any(uint32_t v) const { const uint64_t* all = exchange(v); uint32_t m = 0; for (int i = 0; i < LANES; i++) if ((uint32_t)all[i] == v) m |= 1u << i; return m; }
    void atomic_inc(uint32_t* p) const { ++*p; }
    void atomic_add(uint32_t* p, uint32_t v) const { *p += v; }
    uint32_t exscan(uint32_t v) const { const uint64_t* all = exchange(v); uint32_t s = 0; for (int i = 0; i < lane; i++) s += (uint32_t)all[i]; return s; }
    void atomic_or32(uint32_t* p, uint32_t v) const { *p |= v; }
    void atomic_or_byte(uint8_t* p, uint32_t v) const { *p = (uint8_t)(*p | v); }
};

typedef WarpEmuT<32> WarpEmu;
namespace emu_detail {
template <int LANES> struct Launch { std::function<void(const WarpEmuT<LANES>&)>* body; WarpEmuShared* sh; int lane; };
template <int LANES> inline void trampoline(unsigned lo, unsigned hi) {
    Launch<LANES>* l = reinterpret_cast<Launch<LANES>*>(((uintptr_t)hi << 32) | lo);
    WarpEmuT<LANES> w; w.lane = l->lane; w.sh = l->sh;
    (*l->body)(w);
    l->sh->done[l->lane] = true;
    swapcontext(&l->sh->lane_ctx[l->lane], &l->sh->main_ctx);
}
}  // namespace emu_detail

// run `body` once per lane of an emulated group of LANES lanes
template <int LANES = 32>
inline void run_warp(std::function<void(const WarpEmuT<LANES>&)> body) {
    WarpEmuShared* sh = new WarpEmuShared();
    size_t const stackSize = 1 << 20;
    std::vector<void*> stacks(LANES);
    std::vector<emu_detail::Launch<LANES>> launches(LANES);
    for (int i = 0; i < LANES; i++) {
        sh->done[i] = false;
        stacks[i] = malloc(stackSize);
        getcontext(&sh->lane_ctx[i]);
        sh->lane_ctx[i].uc_stack.ss_sp = stacks[i];
        sh->lane_ctx[i].uc_stack.ss_size = stackSize;
        sh->lane_ctx[i].uc_link = &sh->main_ctx;
        launches[i] = {&body, sh, i};
        uintptr_t const p = reinterpret_cast<uintptr_t>(&launches[i]);
        makecontext(&sh->lane_ctx[i], (void (*)())emu_detail::trampoline<LANES>, 2, (unsigned)(p & 0xFFFFFFFFu), (unsigned)(p >> 32));
    }
    for (;;) {
        bool any = false;
        for (int i = 0; i < LANES; i++) {
            if (sh->done[i]) continue;
            any = true;
            sh->current = i;
            swapcontext(&sh->main_ctx, &sh->lane_ctx[i]);
        }
        if (!any) break;
    }
    for (void* s : stacks) free(s);
    delete sh;
}

}  // namespace zb
