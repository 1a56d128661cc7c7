// tests/simt/simt.hpp — a 32-lane warp on the CPU: every lane is a fiber (ucontext), warp collectives are rendezvous points.
// TEST INFRASTRUCTURE.  It lets the `-m "not gpu"` suite execute the DEVICE SOURCE of a kernel (compiled by g++ with the
// stand-in headers of this directory) and compare its output with the oracle, so the lane-level logic of the shipped
// kernel — not a second restatement of it — is checked where there is no GPU.  What it models: lanes run one after the
// other between collectives; __shfl / __shfl_up / __ballot / __match_any / __reduce_or / __reduce_max / __syncwarp need ALL 32 lanes
// (the emulated sources only use them convergently) and fail loudly on a divergent call; shared memory is a bounds-checked
// byte array addressed by offsets; __ldg checks the registered source range.  What it does not model: timing, memory
// ordering between lanes inside a collective-free region (a missing __syncwarp is only caught when the lane order exposes it).
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#if !defined(__x86_64__)
#include <ucontext.h>
#endif

#include <functional>
#include <map>

namespace simt {

// ---- fibers: a minimal stack switch on x86-64 (callee-saved registers and the stack pointer; glibc's swapcontext costs a
// signal-mask system call per switch, and a warp primitive is 64 switches), ucontext elsewhere
#if defined(__x86_64__)
struct Ctx {
    void* sp;
};
extern "C" void simt_switch(Ctx* from, Ctx* to) __attribute__((visibility("hidden")));
asm(R"(
    .text
    .hidden simt_switch
    .globl simt_switch
    .type simt_switch,@function
simt_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq (%rsi), %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size simt_switch,.-simt_switch
)");
#else
struct Ctx {
    ucontext_t uc;
};
inline void simt_switch(Ctx* from, Ctx* to) { swapcontext(&from->uc, &to->uc); }
#endif

enum Op { OP_NONE = 0, OP_SHFL, OP_SHFL_UP, OP_SHFL_DOWN, OP_BALLOT, OP_MATCH_ANY, OP_REDUCE_OR, OP_REDUCE_MAX, OP_SYNC };

struct Warp {   // (a block of up to two warps: the parser / resolver pairs of the flag-LZ decode kernel)
    static constexpr int kLanes = 64;
    static constexpr size_t kStack = 256 * 1024;
    int nlanes = 32;
    // named barriers (PTX bar.sync / bar.arrive with a thread count): arrivals so far and the generation waiters sleep on
    static constexpr int kBars = 16;
    uint32_t bar_count[kBars] = {}, bar_gen[kBars] = {};
    int bar_wait_id[kLanes];       // -1: not waiting on a named barrier
    uint32_t bar_wait_gen[kLanes];
    // mbarriers (keyed by their shared-memory address): arrivals still expected in the current phase, transaction bytes still
    // in flight, completed phases; a lane sleeps in mbar_wait until the phase of the parity it names has completed
    struct MBar {
        uint32_t count = 1, pending = 1, phase = 0;
        int64_t tx = 0;
    };
    std::map<const void*, MBar> mbar;
    const void* mbar_wait_ptr[kLanes];   // nullptr: not waiting on an mbarrier
    uint32_t mbar_wait_parity[kLanes];
    Ctx sched, ctx[kLanes];
    char* stack[kLanes];
    bool finished[kLanes], waiting[kLanes];
    int cur = 0;
    int op[kLanes];
    uint64_t val[kLanes], res[kLanes];
    uint32_t aux[kLanes];
    std::function<void(int)> body;
    // memory
    uint8_t* smem = nullptr;
    size_t smem_size = 0;
    const uint8_t* g_lo = nullptr;   // readable global range for __ldg
    const uint8_t* g_hi = nullptr;
    const char* error = nullptr;
    uint64_t collectives = 0;
    // SIMT_SHUFFLE=<seed>: the lanes run in a new random order on every scheduler pass instead of 0, 1, 2, ... — a lane that
    // reads what another lane wrote earlier in program order WITHOUT a warp primitive in between then sees stale data some of the
    // time, so the comparison with the oracle doubles as a detector of missing __syncwarp / barrier orderings
    uint64_t shuffle = 0;
};

inline Warp*& current() {
    static thread_local Warp* w = nullptr;
    return w;
}

[[noreturn]] inline void fail(const char* what) {
    fprintf(stderr, "simt: %s (thread %d)\n", what, current() ? current()->cur : -1);
    abort();
}

inline void trampoline() {   // first activation of a fiber: the scheduler has set `cur`
    Warp* w = current();
    const int lane = w->cur;
    w->body(lane);
    w->finished[lane] = true;
    simt_switch(&w->ctx[lane], &w->sched);
    abort();   // a finished fiber is never resumed
}
inline void make_fiber(Warp& w, int l) {
#if defined(__x86_64__)
    void** sp = reinterpret_cast<void**>(reinterpret_cast<uintptr_t>(w.stack[l] + Warp::kStack) & ~uintptr_t(15));
    *--sp = nullptr;                                     // the "return address" of trampoline (never used): rsp % 16 == 8 at its entry
    *--sp = reinterpret_cast<void*>(&trampoline);        // simt_switch's `ret` lands here
    for (int i = 0; i < 6; i++) *--sp = nullptr;         // rbp, rbx, r12 - r15
    w.ctx[l].sp = sp;
#else
    getcontext(&w.ctx[l].uc);
    w.ctx[l].uc.uc_stack.ss_sp = w.stack[l];
    w.ctx[l].uc.uc_stack.ss_size = Warp::kStack;
    w.ctx[l].uc.uc_link = &w.sched.uc;
    makecontext(&w.ctx[l].uc, reinterpret_cast<void (*)()>(trampoline), 0);
#endif
}

// deposit an operand and yield until every lane has arrived
inline uint64_t collective(int op, uint64_t v, uint32_t aux) {
    Warp* w = current();
    const int l = w->cur;
    w->op[l] = op;
    w->val[l] = v;
    w->aux[l] = aux;
    w->waiting[l] = true;
    simt_switch(&w->ctx[l], &w->sched);
    return w->res[l];
}

inline void resolve(Warp* w, int base) {   // the 32 lanes from `base` all wait in a warp primitive
    const int op = w->op[base];
    for (int l = 0; l < 32; l++)
        if (w->op[base + l] != op) fail("divergent collective: lanes wait in different warp primitives");
    w->collectives++;
    const uint64_t* val = w->val + base;
    const uint32_t* aux = w->aux + base;
    for (int l = 0; l < 32; l++) {
        uint64_t r = 0;
        switch (op) {
            case OP_SHFL: r = val[aux[l] & 31]; break;
            case OP_SHFL_UP: r = l >= int(aux[l]) ? val[l - int(aux[l])] : val[l]; break;
            case OP_SHFL_DOWN: r = l + int(aux[l]) < 32 ? val[l + int(aux[l])] : val[l]; break;
            case OP_BALLOT:
                for (int j = 0; j < 32; j++) r |= uint64_t(val[j] != 0) << j;
                break;
            case OP_MATCH_ANY:
                for (int j = 0; j < 32; j++) r |= uint64_t(val[j] == val[l]) << j;
                break;
            case OP_REDUCE_OR:
                if (!((aux[l] >> l) & 1u)) fail("__reduce_or_sync: the calling lane is not in its own mask");
                for (int j = 0; j < 32; j++)
                    if ((aux[l] >> j) & 1u) {
                        if (aux[j] != aux[l]) fail("__reduce_or_sync: lanes of one group passed different masks");
                        r |= val[j];
                    }
                break;
            case OP_REDUCE_MAX:
                for (int j = 0; j < 32; j++)
                    if (val[j] > r) r = val[j];
                break;
            default: break;
        }
        w->res[base + l] = r;
    }
    for (int l = 0; l < 32; l++) w->waiting[base + l] = false;
}

// PTX bar.arrive / bar.sync a, 64: every calling THREAD counts; the 64th arrival releases the sleepers and resets the barrier
inline void bar_arrive_impl(Warp* w, uint32_t id) {
    if (id >= uint32_t(Warp::kBars)) fail("named barrier id out of range");
    if (++w->bar_count[id] == 64) {
        w->bar_count[id] = 0;
        w->bar_gen[id]++;
    } else if (w->bar_count[id] > 64) {
        fail("named barrier over-subscribed");
    }
}
inline void mbar_settle(Warp::MBar& b) {
    if (b.pending == 0 && b.tx == 0) {
        b.phase++;
        b.pending = b.count;
    }
}
inline void mbar_init_impl(const void* bar, uint32_t count) {
    Warp::MBar& b = current()->mbar[bar];
    b.count = b.pending = count;
    b.phase = 0;
    b.tx = 0;
}
inline void mbar_arrive_impl(const void* bar, int64_t expect_tx) {
    Warp::MBar& b = current()->mbar[bar];
    if (b.pending == 0) fail("mbarrier: more arrivals than its count");
    b.tx += expect_tx;
    b.pending--;
    mbar_settle(b);
}
inline void mbar_complete_tx(const void* bar, int64_t bytes) {
    Warp::MBar& b = current()->mbar[bar];
    b.tx -= bytes;
    mbar_settle(b);
}
inline bool mbar_test(const void* bar, uint32_t parity) { return (current()->mbar[bar].phase & 1u) != (parity & 1u); }
inline void mbar_wait_impl(const void* bar, uint32_t parity) {
    Warp* w = current();
    if (mbar_test(bar, parity)) return;
    const int l = w->cur;
    w->mbar_wait_ptr[l] = bar;
    w->mbar_wait_parity[l] = parity;
    simt_switch(&w->ctx[l], &w->sched);
}

inline void named_barrier(uint32_t id, bool wait) {
    Warp* w = current();
    const int l = w->cur;
    const uint32_t gen = w->bar_gen[id < uint32_t(Warp::kBars) ? id : 0];
    bar_arrive_impl(w, id);
    if (!wait) return;
    if (w->bar_gen[id] != gen) return;   // this arrival completed it
    w->bar_wait_id[l] = int(id);
    w->bar_wait_gen[l] = gen;
    simt_switch(&w->ctx[l], &w->sched);
}

// run `body(thread)` on `nwarps` x 32 lanes to completion (thread = warp * 32 + lane)
inline void run_block(Warp& w, int nwarps, std::function<void(int)> body) {
    current() = &w;
    w.body = std::move(body);
    w.nlanes = 32 * nwarps;
    if (w.nlanes > Warp::kLanes) fail("too many warps for the emulated block");
    for (int l = 0; l < w.nlanes; l++) {
        w.finished[l] = w.waiting[l] = false;
        w.bar_wait_id[l] = -1;
        w.mbar_wait_ptr[l] = nullptr;
        w.stack[l] = static_cast<char*>(malloc(Warp::kStack));
        make_fiber(w, l);
    }
    if (const char* e = getenv("SIMT_SHUFFLE")) w.shuffle = strtoull(e, nullptr, 10) * 0x9E3779B97F4A7C15ull + 1;
    int order[Warp::kLanes];
    for (int l = 0; l < w.nlanes; l++) order[l] = l;
    for (;;) {
        bool progress = false;
        int done = 0;
        if (w.shuffle) {
            for (int i = w.nlanes - 1; i > 0; i--) {
                w.shuffle ^= w.shuffle << 13;
                w.shuffle ^= w.shuffle >> 7;
                w.shuffle ^= w.shuffle << 17;
                const int j = int(w.shuffle % uint64_t(i + 1));
                const int t = order[i];
                order[i] = order[j];
                order[j] = t;
            }
        }
        for (int k = 0; k < w.nlanes; k++) {
            const int l = order[k];
            if (w.finished[l]) {
                done++;
                continue;
            }
            if (w.bar_wait_id[l] >= 0) {
                if (w.bar_gen[w.bar_wait_id[l]] == w.bar_wait_gen[l]) continue;   // still asleep in the named barrier
                w.bar_wait_id[l] = -1;
            }
            if (w.mbar_wait_ptr[l]) {
                if (!mbar_test(w.mbar_wait_ptr[l], w.mbar_wait_parity[l])) continue;   // still asleep in the mbarrier
                w.mbar_wait_ptr[l] = nullptr;
            }
            if (!w.waiting[l]) {
                w.cur = l;
                simt_switch(&w.sched, &w.ctx[l]);
                progress = true;
                if (w.finished[l]) done++;
            }
        }
        if (done == w.nlanes) break;
        for (int base = 0; base < w.nlanes; base += 32) {
            int waiting = 0, fin = 0;
            for (int l = base; l < base + 32; l++) {
                waiting += w.waiting[l] ? 1 : 0;
                fin += w.finished[l] ? 1 : 0;
            }
            if (waiting == 0) continue;
            if (waiting + fin == 32) {
                if (fin != 0) fail("a lane left the kernel while others of its warp wait in a warp primitive");
                resolve(&w, base);
                progress = true;
            }
        }
        if (!progress) fail("deadlock: every lane waits (a named barrier that never completes, or a divergent warp primitive)");
    }
    for (int l = 0; l < w.nlanes; l++) free(w.stack[l]);
    current() = nullptr;
}
inline void run_warp(Warp& w, std::function<void(int)> body) { run_block(w, 1, std::move(body)); }

}  // namespace simt

// ---- the CUDA names the emulated sources use -------------------------------------------------------------------------------
inline int simt_lane() { return simt::current()->cur & 31; }
inline int simt_warp() { return simt::current()->cur >> 5; }
template <class T>
inline T __shfl_sync(unsigned mask, T v, int src) {
    if (mask != 0xFFFFFFFFu) simt::fail("__shfl_sync with a partial mask is not modelled");
    return T(simt::collective(simt::OP_SHFL, uint64_t(v), uint32_t(src)));
}
template <class T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta) {
    if (mask != 0xFFFFFFFFu) simt::fail("__shfl_up_sync with a partial mask is not modelled");
    return T(simt::collective(simt::OP_SHFL_UP, uint64_t(v), delta));
}
template <class T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta) {
    if (mask != 0xFFFFFFFFu) simt::fail("__shfl_down_sync with a partial mask is not modelled");
    return T(simt::collective(simt::OP_SHFL_DOWN, uint64_t(v), delta));
}
inline unsigned __ballot_sync(unsigned mask, int pred);
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == 0xFFFFFFFFu; }
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline unsigned __ballot_sync(unsigned mask, int pred) {
    if (mask != 0xFFFFFFFFu) simt::fail("__ballot_sync with a partial mask is not modelled");
    return unsigned(simt::collective(simt::OP_BALLOT, uint64_t(pred != 0), 0));
}
inline unsigned __match_any_sync(unsigned mask, uint32_t v) {
    if (mask != 0xFFFFFFFFu) simt::fail("__match_any_sync with a partial mask is not modelled");
    return unsigned(simt::collective(simt::OP_MATCH_ANY, uint64_t(v), 0));
}
inline unsigned __reduce_or_sync(unsigned mask, unsigned v) { return unsigned(simt::collective(simt::OP_REDUCE_OR, uint64_t(v), mask)); }
inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
    if (mask != 0xFFFFFFFFu) simt::fail("__reduce_max_sync with a partial mask is not modelled");
    return unsigned(simt::collective(simt::OP_REDUCE_MAX, uint64_t(v), 0));
}
inline void __syncwarp() { simt::collective(simt::OP_SYNC, 0, 0); }
inline unsigned __brev(unsigned v) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz(unsigned(v)); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift) { return uint32_t(((uint64_t(hi) << 32) | lo) >> (shift & 31)); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
inline uint64_t min(uint64_t a, uint64_t b) { return a < b ? a : b; }
inline uint64_t max(uint64_t a, uint64_t b) { return a > b ? a : b; }
inline int64_t min(int64_t a, int64_t b) { return a < b ? a : b; }
inline int64_t max(int64_t a, int64_t b) { return a > b ? a : b; }
// CUDA's mixed overloads: (unsigned, int) compares as unsigned
inline uint32_t min(uint32_t a, int b) { return min(a, uint32_t(b)); }
inline uint32_t min(int a, uint32_t b) { return min(uint32_t(a), b); }
inline uint32_t max(uint32_t a, int b) { return max(a, uint32_t(b)); }
inline uint32_t max(int a, uint32_t b) { return max(uint32_t(a), b); }
template <class T>
inline T __ldg(const T* p) {
    const simt::Warp* w = simt::current();
    const uint8_t* b = reinterpret_cast<const uint8_t*>(p);
    if (b < w->g_lo || b + sizeof(T) > w->g_hi) simt::fail("__ldg outside the batch's source bytes");
    T v;
    memcpy(&v, p, sizeof(T));
    return v;
}
// shared memory: addresses are byte offsets into the warp's array
inline uint8_t* simt_smem(uint32_t addr, uint32_t bytes) {
    simt::Warp* w = simt::current();
    if (size_t(addr) + bytes > w->smem_size) simt::fail("shared-memory access outside the warp's tables");
    if (addr % bytes) simt::fail("misaligned shared-memory access");
    return w->smem + addr;
}
