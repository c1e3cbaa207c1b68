// Fire exchange over NVLink peer memory: the one collective of the path, fused into the step.
//
// Every shard owns an exchange ARENA in its device memory — two gather buffers (window parity) with one block per shard
// (header unit {count, overflow} + fire records), one counter block per shard, and flag words — and maps every other
// shard's arena through CUDA IPC (all shards are processes of one node; NVSwitch gives every pair full bandwidth).
// After the neuron pass a shard STORES its block straight into slot `rank` of every peer's gather buffer, fences, and
// raises its flag in every peer (the window's sequence number); the consumer side waits on its own flag words and goes
// on to build the fire index.  The per-window counters (incl. the hidden rand() count that moves the stream on) travel the
// same way.  There is no host round trip and no library call in the step; the payload is exactly count + 1 units.
//
// Why two gather buffers are enough: a peer can push window w+1 only after it has finished window w, which needed this
// shard's block of window w and its counters — so it can be at most one window ahead, and never two.
#pragma once
#include <stdint.h>

#include "step_logic.cuh"

namespace ncx {
using namespace ncs;

#define NC_X_FLAGS_BYTES 4096      // flag words: [kind 0 fires / 1 window counters / 2 replay totals][NC_MAX_WORLD]
#define NC_X_CNT_BYTES 4096        // counter blocks: [0] the window's, [1] a replay's totals; each world x 10 x u64
#define NC_X_CNT_AREA 1024         // bytes per counter area (NC_MAX_WORLD x 10 x 8 = 640)
#define NC_X_TIMEOUT_CYCLES 6000000000ll  // ~3 s: a peer that died must not hang the GPU

struct PeerTab {  // this window's view of every shard's arena (device pointers valid on THIS device)
    FireRec* gather[NC_MAX_WORLD];            // the window's gather buffer of shard r
    unsigned long long* counters[NC_MAX_WORLD];
    uint32_t* flags[NC_MAX_WORLD];
};
struct XchgArgs {  // what the step's own kernels need to do the exchange themselves; world <= 1: no exchange
    PeerTab pt;
    uint32_t world, rank, blockUnits, seq;
    uint32_t* doneCtr;       // blocks of the neuron pass that have finished (the last one pushes)
    uint32_t* errWord;
};

#if defined(__CUDACC__)
#if defined(NC_BLOCK_EMU)  // (tests/native/cuda_runtime_emu.h compiles this header for the CPU: no PTX there, and one shard only)
__device__ __forceinline__ void st_flag(uint32_t* p, uint32_t v) { *(volatile uint32_t*)p = v; }
__device__ __forceinline__ uint32_t ld_flag(const uint32_t* p) { return *(const volatile uint32_t*)p; }
#else
__device__ __forceinline__ void st_flag(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_flag(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
#endif

// ---- the same, as the tail / head of the step's own kernels (no extra launches) ----
// Tail of the neuron pass: every block fences its fire records and takes a ticket; the LAST block stores the shard's block
// (header + count records) into slot `rank` of every shard's gather buffer and raises this shard's flag everywhere.
__device__ __forceinline__ void push_fires_tail(const View& v, const XchgArgs& xa) {
    __shared__ uint32_t sLast;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) sLast = (atomicAdd(xa.doneCtr, 1u) == gridDim.x - 1u) ? 1u : 0u;
    __syncthreads();
    if (!sLast) return;
    __threadfence();
    const uint32_t units = min(*(volatile uint32_t*)v.localHdr, v.fireCap) + 1u;
    const uint4* src = reinterpret_cast<const uint4*>(v.localHdr);
    for (uint32_t r = 0; r < xa.world; r++) {
        uint4* dst = reinterpret_cast<uint4*>(xa.pt.gather[r] + (uint64_t)xa.rank * xa.blockUnits);
        for (uint32_t i = threadIdx.x; i < units; i += blockDim.x) dst[i] = __ldcg(src + i);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        *xa.doneCtr = 0u;
        for (uint32_t r = 0; r < xa.world; r++) st_flag(xa.pt.flags[r] + xa.rank, xa.seq);
    }
}
// Head of a consumer kernel: wait until every shard's flag of `kind` has reached this window's sequence number.
__device__ __forceinline__ void wait_flags_head(const XchgArgs& xa, uint32_t kind) {
    if (threadIdx.x < xa.world) {
        const uint32_t* f = xa.pt.flags[xa.rank] + kind * NC_MAX_WORLD + threadIdx.x;
        const long long t0 = clock64();
        while ((int32_t)(ld_flag(f) - xa.seq) < 0) {
            __nanosleep(32);
            if (clock64() - t0 > NC_X_TIMEOUT_CYCLES) { atomicOr(xa.errWord, 4u); break; }
        }
    }
    __threadfence_system();
    __syncthreads();
}
// Tail of the end-of-window kernel (one warp): the window's counter block into slot `rank` of every shard's window area.
__device__ __forceinline__ void push_counters_tail(const unsigned long long* win, const XchgArgs& xa) {
    const uint32_t lane = threadIdx.x;
    __syncwarp();
    for (uint32_t r = 0; r < xa.world; r++)
        if (lane < 10u) xa.pt.counters[r][xa.rank * 10u + lane] = win[lane];
    __threadfence_system();
    __syncwarp();
    if (lane < xa.world) st_flag(xa.pt.flags[lane] + NC_MAX_WORLD + xa.rank, xa.seq);
}

// ---- stand-alone forms (a replay's totals; engines without the device-resident rand() stream) ----
// The window's counter block (10 x u64) -> slot `rank` of every shard's counter area, then the flag.  One warp.
__global__ void k_push_counters(const unsigned long long* win, PeerTab pt, uint32_t world, uint32_t rank, uint32_t seq, uint32_t kind) {
    const uint32_t lane = threadIdx.x;
    for (uint32_t r = 0; r < world; r++)
        if (lane < 10u) (pt.counters[r] + (kind == 2u ? NC_X_CNT_AREA / 8 : 0))[rank * 10u + lane] = win[lane];
    __threadfence_system();
    __syncwarp();
    if (lane < world) st_flag(pt.flags[lane] + kind * NC_MAX_WORLD + rank, seq);
}
// Wait until every shard's flag of `kind` has reached this window's sequence number (signed distance: the counter wraps).
// A peer that never arrives sets the shard's exchange-error word instead of hanging the GPU.
__global__ void k_wait_flags(const uint32_t* flags, uint32_t kind, uint32_t world, uint32_t seq, uint32_t* errWord) {
    const uint32_t lane = threadIdx.x;
    if (lane < world) {
        const long long t0 = clock64();
        while ((int32_t)(ld_flag(flags + kind * NC_MAX_WORLD + lane) - seq) < 0) {
            __nanosleep(64);
            if (clock64() - t0 > NC_X_TIMEOUT_CYCLES) { atomicOr(errWord, 4u); break; }
        }
    }
    __threadfence_system();
}
#endif

}  // namespace ncx
