// TEST INFRASTRUCTURE — never compiled into or loaded by the product.
//
// A small CPU emulator of the CUDA execution model for ONE kernel launch at a time, so that kernel source from
// neurocorrelation_b200/csrc/*.cuh can be compiled with g++ and exercised in the CPU test-suite (the container that runs
// `pytest -m "not gpu"` has no GPU).  It checks LOGIC — indexing, the use of barriers and warp collectives, capacities —
// not performance and not the memory model.
//
//   * every CUDA thread of a block is a fibre on one OS thread (own stack, a dozen-instruction x86-64 context switch: no
//     system call per switch, which is what makes whole parity scenarios affordable); blocks of a grid run one after the other;
//   * __syncthreads() parks a fibre until every live thread of the block has arrived;
//   * warp collectives (__syncwarp, __ballot_sync, __any_sync, __all_sync, __shfl_*_sync, __reduce_*_sync) park a lane until
//     every live lane named in the mask has arrived at a collective of the SAME kind (anything else aborts: on the GPU
//     it would be undefined behaviour);
//   * __shared__ becomes `static thread_local` (an OS thread runs one block at a time), atomics are real atomic operations;
//   * a fibre that leaves the kernel counts as arrived at every later barrier, as on the hardware;
//   * no progress by any fibre = deadlock (e.g. a barrier inside divergent code): aborts with a message.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#if !defined(__x86_64__)
#error "cuda_block_emu.h: the fibre switch is written for x86-64"
#endif
// emu_switch(&save_sp, to_sp): saves the callee-saved registers on the current stack, stores the stack pointer, switches
extern "C" void emu_switch(void** from_sp, void* to_sp);
asm(R"(
.text
.hidden emu_switch
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

#if defined(__SANITIZE_THREAD__)  // built by tools/emu_tsan.sh: ThreadSanitizer must be told about the fibres
extern "C" void* __tsan_get_current_fiber(void);
extern "C" void* __tsan_create_fiber(unsigned flags);
extern "C" void __tsan_destroy_fiber(void* fiber);
extern "C" void __tsan_switch_to_fiber(void* fiber, unsigned flags);
#define EMU_TSAN 1
#endif

namespace emu {

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

enum Wait { RUN = 0, AT_BLOCK = 1, AT_WARP = 2, DONE = 3 };
enum Kind { K_SYNCWARP = 1, K_BALLOT, K_ANY, K_ALL, K_SHFL, K_SHFL_UP, K_SHFL_DOWN, K_SHFL_XOR, K_RED_MAX, K_RED_MIN, K_RED_ADD };

struct Fibre {
    void* sp = nullptr;    // saved stack pointer while the fibre is parked
    void* tsan = nullptr;  // ThreadSanitizer's handle of this fibre (EMU_TSAN builds)
    char* stack = nullptr;
    Wait wait = RUN;
    dim3 tid;
    unsigned linear = 0;
    // the warp collective this lane is parked at
    int kind = 0;
    uint32_t mask = 0;
    uint64_t payload = 0;
    int aux = 0;
    uint64_t result = 0;
};

struct State {
    std::vector<Fibre> f;            // the threads of the block being run
    std::vector<char*> stacks;       // fibre stacks, kept across blocks and launches
    std::vector<unsigned> perm;      // resume order of the current block's threads
    void* schedSp = nullptr;
    void* schedTsan = nullptr;
    Fibre* cur = nullptr;
    dim3 bIdx, bDim, gDim;
    const std::function<void()>* body = nullptr;
    size_t stackBytes = 256 * 1024;
    unsigned long long collectives = 0, barriers = 0;
    const char* kernel = "?";           // text of the launch being run (diagnostics)
    unsigned onTheWay = 0, atBlock = 0;  // threads of the block that are running or parked at a warp collective / parked at the block barrier
};
inline State& S() { static thread_local State s; return s; }  // (one per OS thread: NC_EMU_THREADS runs blocks on several)
// dynamic shared memory of the block this OS thread is running, and the size the current launch asked for
inline std::vector<unsigned char>& dyn() { static thread_local std::vector<unsigned char> b; return b; }
inline size_t& dyn_bytes() { static size_t n = 0; return n; }

[[noreturn]] inline void die(const char* what) {
    fprintf(stderr, "cuda_block_emu: %s (kernel %s, block %u,%u thread %u)\n", what, S().kernel, S().bIdx.x, S().bIdx.y, S().cur ? S().cur->linear : 0u);
    abort();
}

inline void yield_to_scheduler() {
#if defined(EMU_TSAN)
    __tsan_switch_to_fiber(S().schedTsan, 0);
#endif
    emu_switch(&S().cur->sp, S().schedSp);
}

inline void fibre_entry() {
    (*S().body)();
    S().cur->wait = DONE;
    S().onTheWay--;
    yield_to_scheduler();
    die("resumed a finished fibre");
}

// ---- completion of barriers --------------------------------------------------------------------------------------------
inline void try_release_block() {
    State& s = S();
    if (s.onTheWay != 0 || s.atBlock == 0) return;  // somebody is still running / parked at a warp collective, or nobody waits
    for (auto& x : s.f)
        if (x.wait == AT_BLOCK) x.wait = RUN;
    s.onTheWay = s.atBlock;
    s.atBlock = 0;
    s.barriers++;
}

inline void try_release_warp(unsigned warp) {
    State& s = S();
    const unsigned lo = warp * 32u, hi = std::min<unsigned>(lo + 32u, (unsigned)s.f.size());
    // the collective is defined by the first parked lane; every live lane of its mask must be parked at the same one
    Fibre* lead = nullptr;
    for (unsigned i = lo; i < hi; i++)
        if (s.f[i].wait == AT_WARP) { lead = &s.f[i]; break; }
    if (!lead) return;
    const uint32_t mask = lead->mask;
    for (unsigned i = lo; i < hi; i++) {
        Fibre& x = s.f[i];
        const bool named = (mask >> (i - lo)) & 1u;
        if (x.wait == AT_WARP) {
            if (x.kind != lead->kind || x.mask != mask) {
                for (unsigned k = lo; k < hi; k++)
                    fprintf(stderr, "  lane %2u: state %d kind %d mask %08x payload %llx aux %d\n", k - lo, (int)s.f[k].wait, s.f[k].kind, s.f[k].mask, (unsigned long long)s.f[k].payload, s.f[k].aux);
                die("lanes of one warp are parked at different collectives / masks");
            }
            if (!named) die("a lane takes part in a collective whose mask does not name it");
        } else if (named && x.wait != DONE) {
            return;  // a named lane has not arrived yet (RUN, or parked at a block barrier = deadlock, caught by the scheduler)
        }
    }
    auto live = [&](unsigned l) { return lo + l < hi && ((mask >> l) & 1u) && s.f[lo + l].wait == AT_WARP; };
    uint32_t ballot = 0;
    bool any = false, all = true;
    uint64_t rmax = 0, rmin = ~0ull, radd = 0;
    for (unsigned l = 0; l < 32; l++)
        if (live(l)) {
            const uint64_t p = s.f[lo + l].payload;
            if (p) { ballot |= 1u << l; any = true; } else all = false;
            rmax = std::max(rmax, p); rmin = std::min(rmin, p); radd += p;
        }
    for (unsigned l = 0; l < 32; l++) {
        if (!live(l)) continue;
        Fibre& x = s.f[lo + l];
        int src = (int)l;
        switch (x.kind) {
            case K_SYNCWARP: x.result = 0; break;
            case K_BALLOT: x.result = ballot; break;
            case K_ANY: x.result = any; break;
            case K_ALL: x.result = all; break;
            case K_RED_MAX: x.result = rmax; break;
            case K_RED_MIN: x.result = rmin; break;
            case K_RED_ADD: x.result = radd; break;
            case K_SHFL: src = x.aux & 31; break;
            case K_SHFL_UP: src = (int)l - x.aux; if (src < 0) src = (int)l; break;
            case K_SHFL_DOWN: src = (int)l + x.aux; if (src > 31) src = (int)l; break;
            case K_SHFL_XOR: src = (int)l ^ x.aux; break;
            default: die("unknown collective");
        }
        if (x.kind >= K_SHFL && x.kind <= K_SHFL_XOR) {
            // reading from a lane that does not take part is undefined on the GPU; the emulator returns the lane's own value
            x.result = live((unsigned)src) ? s.f[lo + (unsigned)src].payload : x.payload;
        }
    }
    for (unsigned l = 0; l < 32; l++)
        if (live(l)) s.f[lo + l].wait = RUN;
    s.collectives++;
}

inline uint64_t warp_collective(int kind, uint32_t mask, uint64_t payload, int aux) {
    Fibre* me = S().cur;
    me->kind = kind; me->mask = mask; me->payload = payload; me->aux = aux;
    me->wait = AT_WARP;
    yield_to_scheduler();
    return me->result;
}

inline void block_barrier() {
    S().cur->wait = AT_BLOCK;
    S().onTheWay--; S().atBlock++;
    yield_to_scheduler();
}

// ---- launch -------------------------------------------------------------------------------------------------------------
inline void run_block(const std::function<void()>& body) {
    State& s = S();
    const unsigned T = s.bDim.x * s.bDim.y * s.bDim.z;
    s.f.assign(T, Fibre());
    while (s.stacks.size() < T) s.stacks.push_back((char*)malloc(s.stackBytes));
    s.body = &body;
    s.onTheWay = T; s.atBlock = 0;
    for (unsigned i = 0; i < T; i++) {
        Fibre& x = s.f[i];
        x.stack = s.stacks[i];
        x.wait = RUN; x.linear = i;
        x.tid = dim3(i % s.bDim.x, (i / s.bDim.x) % s.bDim.y, i / (s.bDim.x * s.bDim.y));
        // a fresh stack that emu_switch can "return" into: six zeroed callee-saved registers, then fibre_entry as the return
        // address, then a null return address for fibre_entry itself (which never returns) — rsp = 8 mod 16 at its entry
        uintptr_t top = ((uintptr_t)(x.stack + s.stackBytes)) & ~(uintptr_t)15;
        void** sp = (void**)top;
        *--sp = nullptr;
        *--sp = (void*)&fibre_entry;
        for (int r = 0; r < 6; r++) *--sp = nullptr;
        x.sp = (void*)sp;
    }
    // NC_EMU_ORDER: the order in which runnable threads are resumed — 0 ascending (default), 1 descending, 2 a fresh pseudo-random
    // permutation per pass.  Results must not depend on it: a difference means threads communicate through memory without
    // a barrier between them (on the GPU: a race, or reliance on lock-step execution of a warp).
    const int order = getenv("NC_EMU_ORDER") ? atoi(getenv("NC_EMU_ORDER")) : 0;  // (read per block: tests switch it within one process)
    static thread_local uint64_t rng = 0x2545F4914F6CDD1Dull;
    std::vector<unsigned>& perm = s.perm;
    perm.resize(T);
    for (unsigned i = 0; i < T; i++) perm[i] = order == 1 ? T - 1 - i : i;
    for (;;) {
        bool progressed = false, allDone = true;
        if (order == 2)
            for (unsigned i = T; i > 1; i--) { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; std::swap(perm[i - 1], perm[rng % i]); }
        for (unsigned pi = 0; pi < T; pi++) {
            const unsigned i = perm[pi];
            Fibre& x = s.f[i];
            if (x.wait != DONE) allDone = false;
            if (x.wait != RUN) continue;
            s.cur = &x;
#if defined(EMU_TSAN)
            if (!s.schedTsan) s.schedTsan = __tsan_get_current_fiber();
            if (!x.tsan) x.tsan = __tsan_create_fiber(0);
            __tsan_switch_to_fiber(x.tsan, 0);
#endif
            emu_switch(&s.schedSp, x.sp);
#if defined(EMU_TSAN)
            if (x.wait == DONE) { __tsan_destroy_fiber(x.tsan); x.tsan = nullptr; }
#endif
            s.cur = nullptr;
            progressed = true;
            if (x.wait == AT_WARP || x.wait == DONE) try_release_warp(i / 32u);
            if (x.wait == AT_BLOCK || x.wait == DONE) try_release_block();
        }
        if (allDone) break;
        if (!progressed) {
            // a lane leaving may complete a warp's collective, a warp completing may complete the block barrier: one more look
            for (unsigned w = 0; w * 32u < T; w++) try_release_warp(w);
            try_release_block();
            bool runnable = false;
            for (auto& x : s.f) runnable = runnable || x.wait == RUN;
            if (!runnable) die("deadlock: every live thread is parked (a barrier or collective in divergent code, or a spin-wait)");
        }
    }
}

struct Pool {  // worker threads that live as long as the process (their fibre stacks are reused from launch to launch)
    std::mutex m;
    std::condition_variable wake, done;
    std::vector<std::thread> workers;
    unsigned long long generation = 0;
    int running = 0;
    const std::function<void()>* body = nullptr;
    dim3 grid, block;
    const char* kernel = "?";
    unsigned long long nBlocks = 0;
    int order = 0;
    std::atomic<unsigned long long> next{0};
};
inline dim3 block_of(unsigned long long k, unsigned long long nBlocks, int order, dim3 grid) {
    unsigned long long b = order == 1 ? nBlocks - 1 - k : order == 2 ? (k * 0x9E3779B1ull + 7) % nBlocks : k;  // (order 2: a stride walk over the blocks)
    return dim3((unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((unsigned long long)grid.x * grid.y)));
}
inline void run_block(const std::function<void()>& body);
inline Pool& pool(int threads) {
    static Pool* p = new Pool();  // (never destroyed: the workers are detached daemons of the test process)
    while ((int)p->workers.size() < threads) {
        p->workers.emplace_back([] {
            Pool& q = *p;
            unsigned long long seen = 0;
            for (;;) {
                {
                    std::unique_lock<std::mutex> lk(q.m);
                    q.wake.wait(lk, [&] { return q.generation != seen; });
                    seen = q.generation;
                }
                State& w = S();
                w.gDim = q.grid; w.bDim = q.block; w.kernel = q.kernel;
                dyn().assign(dyn_bytes() + 64, 0xA5);
                for (;;) {
                    const unsigned long long k = q.next.fetch_add(1);
                    if (k >= q.nBlocks) break;
                    w.bIdx = block_of(k, q.nBlocks, q.order, q.grid);
                    run_block(*q.body);
                }
                {
                    std::unique_lock<std::mutex> lk(q.m);
                    if (--q.running == 0) q.done.notify_all();
                }
            }
        });
        p->workers.back().detach();
    }
    return *p;
}

// NC_EMU_THREADS=n (default 1): the blocks of a launch are run by n OS threads at once (each with its own fibres, shared
// memory and block index), atomics are real atomic operations.  Results must not depend on it: blocks may only meet through
// atomics.  (x86's memory model is stronger than the GPU's, so this finds logic that depends on block order, not missing fences.)
template <typename F>
inline void launch(dim3 grid, dim3 block, F&& kernel_call) {
    State& s = S();
    s.gDim = grid; s.bDim = block;
    const std::function<void()> body = kernel_call;
    const int order = getenv("NC_EMU_ORDER") ? atoi(getenv("NC_EMU_ORDER")) : 0;
    const int threads = getenv("NC_EMU_THREADS") ? atoi(getenv("NC_EMU_THREADS")) : 1;
    const unsigned long long nBlocks = (unsigned long long)grid.x * grid.y * grid.z;
    if (threads <= 1 || nBlocks < 2) {
        for (unsigned long long k = 0; k < nBlocks; k++) { s.bIdx = block_of(k, nBlocks, order, grid); run_block(body); }
        return;
    }
    Pool& p = pool(threads);
    {
        std::unique_lock<std::mutex> lk(p.m);
        p.body = &body; p.grid = grid; p.block = block; p.kernel = s.kernel; p.nBlocks = nBlocks; p.order = order;
        p.next.store(0);
        p.running = (int)p.workers.size();
        p.generation++;
    }
    p.wake.notify_all();
    std::unique_lock<std::mutex> lk(p.m);
    p.done.wait(lk, [&] { return p.running == 0; });
}

}  // namespace emu

// ---- the CUDA vocabulary the kernels use ------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __restrict__
#define threadIdx (emu::S().cur->tid)
#define blockIdx (emu::S().bIdx)
#define blockDim (emu::S().bDim)
#define gridDim (emu::S().gDim)
using emu::dim3;

inline void __syncthreads() { emu::block_barrier(); }
inline void __syncwarp(uint32_t mask = 0xffffffffu) { emu::warp_collective(emu::K_SYNCWARP, mask, 0, 0); }
inline uint32_t __ballot_sync(uint32_t mask, int pred) { return (uint32_t)emu::warp_collective(emu::K_BALLOT, mask, pred ? 1 : 0, 0); }
inline int __any_sync(uint32_t mask, int pred) { return (int)emu::warp_collective(emu::K_ANY, mask, pred ? 1 : 0, 0); }
inline int __all_sync(uint32_t mask, int pred) { return (int)emu::warp_collective(emu::K_ALL, mask, pred ? 1 : 0, 0); }
inline uint32_t __reduce_max_sync(uint32_t mask, uint32_t v) { return (uint32_t)emu::warp_collective(emu::K_RED_MAX, mask, v, 0); }
inline uint32_t __reduce_min_sync(uint32_t mask, uint32_t v) { return (uint32_t)emu::warp_collective(emu::K_RED_MIN, mask, v, 0); }
inline uint32_t __reduce_add_sync(uint32_t mask, uint32_t v) { return (uint32_t)emu::warp_collective(emu::K_RED_ADD, mask, v, 0); }

namespace emu {
template <typename T> inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, "shuffle payload"); memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
}  // namespace emu
template <typename T> inline T __shfl_sync(uint32_t mask, T v, int src) { return emu::from_bits<T>(emu::warp_collective(emu::K_SHFL, mask, emu::to_bits(v), src)); }
template <typename T> inline T __shfl_up_sync(uint32_t mask, T v, unsigned d) { return emu::from_bits<T>(emu::warp_collective(emu::K_SHFL_UP, mask, emu::to_bits(v), (int)d)); }
template <typename T> inline T __shfl_down_sync(uint32_t mask, T v, unsigned d) { return emu::from_bits<T>(emu::warp_collective(emu::K_SHFL_DOWN, mask, emu::to_bits(v), (int)d)); }
template <typename T> inline T __shfl_xor_sync(uint32_t mask, T v, int x) { return emu::from_bits<T>(emu::warp_collective(emu::K_SHFL_XOR, mask, emu::to_bits(v), x)); }

template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __ldcs(const T* p) { return *p; }
// atomics: real ones (blocks may run on several OS threads)
template <typename T> inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <typename T> inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <typename T> inline T atomicAnd(T* p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
template <typename T> inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <typename T> inline T atomicCAS(T* p, T cmp, T v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
template <typename T> inline T atomicMin(T* p, T v) {
    T o = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return o;
}
template <typename T> inline T atomicMax(T* p, T v) {
    T o = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return o;
}
inline int __popc(uint32_t x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
using std::max;
using std::min;
// round-to-nearest single operations (the host is built with -ffp-contract=off, so plain operators are the same thing)
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __int2float_rn(int x) { return (float)x; }
inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
