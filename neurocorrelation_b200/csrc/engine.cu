// libneucor_b200.so — the device engine behind include/neucor_b200.h.
//
// One window (t0, t1] of the reference's event loop (NeuCor::run, /root/reference/src/NeuCor.cpp:583-617) is executed as two
// data-parallel passes over a post-synaptic-sorted CSR, neither of which visits idle synapses (DESIGN.md sections 3-4):
//
//   k_stage        keeps, per tile of 32 rows, the list of occupied slots (arrived, not yet cleared) current: per-word arrival
//                  bounds name the few index words worth looking at, the window's arrivals are merged into the persistent
//                  list, entries the last neuron pass cleared are dropped (rebuild from the busy / arrived bitmaps when needed).
//   k_neuron_pass  warps claim tiles; a tile's list goes into the warp's shared-memory pool in batches and every lane REPLAYS
//                  one neuron's in-window events (deliveries, +2 ms requeues, host / background events, the end-of-window
//                  sweep) in the canonical order with Neuron::run / fire semantics (NeuCor.cpp:619-714): ordered accumulation
//                  over its active slots in ascending presynaptic ID, passive decay, threshold / refractory check, AP waveform,
//                  activity.  Emits fire records, marks cleared slots, hands delivered / cleared slots to the synapse pass
//                  (flag list).  Never touches weights.  Over-long rows take a warp-per-row path with an exact prefix-sum
//                  form of the ordered accumulation.  On a sharded engine its last block pushes the shard's fire records
//                  into every peer's memory (peer_exchange.cuh).
//   k_index_build  fire records of all shards -> bitmask + per-neuron record lists.
//   k_syn_loads / k_syn_rows / k_syn_flagged   the synapse pass: only eventful slots, each owned by exactly one kernel — the
//                  out-synapses of fired neurons (Synapse::fire, NeuCor.cpp:727-738), the in-synapses of fired neurons
//                  (post-fire plasticity), the flagged slots (delivery: Synapse::run / synapticPlasticity, NeuCor.cpp:718-764;
//                  clear, NeuCor.cpp:697) — resolved in the canonical event order by resolve_slot() (step_logic.cuh).
//   k_bg_generate / k_bg_walk / k_merge_events / k_finish_step   libc's rand() stream on the device: background firing
//                  (NeuCor.cpp:604-607), its merge with the host's input-firer events, the hidden rand() calls (rand_stream.cuh).
//
// Arithmetic mirrors the reference's float/double typing operator by operator with explicit-rounding
// intrinsics (no FMA contraction; built with -fmad=false) and glibc-exact powf/exp (glibc_math.cuh).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/neucor_b200.h"
#include "step_logic.cuh"
#include "rand_stream.cuh"
#include "peer_exchange.cuh"

using namespace ncm;
using namespace ncs;

#define NC_WARPS_PER_BLOCK 4

// ------------------------------------------------------------------------------------------------
// Neuron pass
// ------------------------------------------------------------------------------------------------
struct EvPick {  // (time, rank<<32|k2) ordering; src: candidate index, or 0x80000000|host event index
    float t;
    unsigned long long code;
    uint32_t src;
};
struct P1Counters { unsigned long long fires, runs, visits, deliveries; };

// Neuron::run (NeuCor.cpp:619-641) executed by a whole warp: the neuron's state is replicated in every lane, the ordered
// accumulation over the row's active slots (charge_insynapses) is evaluated 32 slots at a time with the exact
// prefix-sum scheme of step_logic.cuh, slots that expire are marked by their own lane.
__device__ __forceinline__ void warp_neuron_run(const View& v, NeuronState& n, CandView& cv, uint32_t cnt, uint64_t rs, uint32_t q, float T,
                                                uint32_t rk1, uint32_t k2, uint32_t sentinel, uint32_t lane, P1Counters& ctr) {
    const uint32_t FULL = 0xffffffffu;
    float dT;
    if (!neuron_run_begin(n, T, dT)) return;
    ctr.runs++;
    float np = n.pot;
    if (cnt) {
        const double E = exp_glibc(mul64(0.3702, (double)dT));
        for (uint32_t base = 0; base < cnt; base += 32) {
            const uint32_t c = base + lane;
            bool act = false;
            double t = 0.0;
            if (c < cnt) {
                float a = cv.A(c);
                if (a > 0.0f) {                 // not cleared earlier in this window
                    float off = sub32(T, a);
                    if (off > 0.0f) {           // arrived
                        act = true;
                        t = chain_term(dT, cv.D(c), E);
                        if (2.0f < off) {       // NeuCor.cpp:697 — the slot becomes idle; leave the when-and-why for the synapse pass
                            cv.A(c) = -a;
                            uint64_t sidx = rs + cv.J(c);
                            v.ad[sidx] = make_float2(__uint_as_float(sentinel), T);
                        }
                    }
                }
            }
            uint32_t todo = __ballot_sync(FULL, act);
            ctr.visits += __popc(todo);
            while (todo) {
                const Binade b = binade_of(np);
                const bool neg = np < 0.0f;
                const bool mine = (todo >> lane) & 1u;
                double r = 0.0;
                bool flag = false;
                if (mine) flag = !chain_lane(b, neg, t, r);
                double pre = r;  // inclusive prefix sum over lanes; all values are multiples of u, the sums are exact
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    double y = __shfl_up_sync(FULL, pre, o);
                    if (lane >= (uint32_t)o) pre = add64(pre, y);
                }
                const double m = fabs((double)np);
                const double mi = add64(m, pre);
                if (mine && !(mi > b.lo && mi < b.hi)) flag = true;  // the running value would leave the open binade
                const uint32_t bad = __ballot_sync(FULL, flag) & todo;
                if (!bad) {
                    double tot = __shfl_sync(FULL, pre, 31);
                    double mm = add64(m, tot);
                    np = (float)(neg ? -mm : mm);
                    todo = 0u;
                } else {
                    const int f = __ffs(bad) - 1;
                    double acc = __shfl_sync(FULL, pre, f > 0 ? f - 1 : 0);
                    if (f == 0) acc = 0.0;
                    double mm = add64(m, acc);
                    np = (float)(neg ? -mm : mm);
                    double tf = __shfl_sync(FULL, t, f);
                    np = (float)add64((double)np, tf);  // this one slot exactly as the reference adds it
                    todo &= ~((2u << f) - 1u);
                }
            }
        }
    }
    if (neuron_run_finish(n, np, T, dT)) {
        ctr.fires++;
        if (lane == 0) emit_fire(v, q, T, rk1, k2);
    }
}

// ---- staging ------------------------------------------------------------------------------------------------
// Occupied slots (arrive != 0 && arrive <= t1) of one row, compacted in row order (= ascending presynaptic ID).
// The row is not scanned: the busy-slot index (one bit per slot, kept in step with `arrive` by the synapse pass) says which
// slots hold a spike at all, and only those are looked at.  Lane l takes the l-th 32-slot word of the row (one coalesced
// 128-byte load covers 1024 slots), gathers `arrive` of its set bits, and the arrived ones (the rest are still in flight)
// are ballot/prefix-compacted into the pool.
//   SPILL = true : entries go to the CandView (shared memory first, per-warp global spill area after) — warp-per-row path
//   SPILL = false: arrive times and slot indices (jbase + index in the row) go to pa/pj while they fit in `room`; the returned
//                  count tells the caller whether they did; the caller gathers depol for the whole batch in one go
// hasEv: some staged slot delivers (t0 < arrive) or re-queues its target (t0 < arrive + 2 <= t1) in this window.
template <bool SPILL>
__device__ __forceinline__ uint32_t stage_row(const View& v, const StepArgs& s, uint64_t rs, uint64_t re, CandView& cv,
                                              float* pa, uint32_t* pj, uint32_t jbase, uint32_t room, uint32_t lane, bool& hasEv) {
    const uint32_t FULL = 0xffffffffu;
    uint32_t cnt = 0;
    bool ev = false;
    const uint32_t t1b = __float_as_uint(s.t1), t0b = __float_as_uint(s.t0);
    const uint64_t w0 = rs >> 5, w1 = (re + 31) >> 5;  // the row's words of the busy index: [w0, w1)
    for (uint64_t wb = w0; wb < w1; wb += 32) {
        const uint64_t w = wb + lane;
        uint32_t word = 0u;
        if (w < w1) {
            word = __ldcg(v.busy + w);
            const uint32_t lo = (w == w0) ? (uint32_t)(rs & 31u) : 0u;                    // first bit that belongs to the row
            const uint32_t hn = (w == w1 - 1) ? (uint32_t)(re - (w << 5)) : 32u;          // bits [0, hn) belong to the row
            word &= (hn >= 32u ? FULL : ((1u << hn) - 1u)) & (FULL << lo);
        }
        if (!__any_sync(FULL, word != 0u)) continue;
        // which of this lane's busy slots have arrived: 0 < arrive <= t1 as ONE integer compare on the bit pattern
        // (arrive times are positive floats: (bits - 1) < bits(t1)); the others are still travelling
        const float2* ap = v.ad + (w << 5);
        uint32_t im = 0u;
        for (uint32_t m = word; m; m &= m - 1u) {
            const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
            const uint32_t ab = __float_as_uint(ap[b].x);
            if (ab - 1u < t1b) {
                im |= 1u << b;
                ev |= (ab > t0b) || (ab > s.reqLoB && ab <= s.reqHiB);
            }
        }
        __syncwarp();
        const uint32_t c = __popc(im);
        uint32_t inc = c;  // inclusive prefix sum over lanes = row order (lane = word, bits ascend within the word)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, inc, o);
            if (lane >= (uint32_t)o) inc += y;
        }
        uint32_t pos = cnt + inc - c;
        cnt += __shfl_sync(FULL, inc, 31);
        const uint32_t rel0 = (uint32_t)((w << 5) - rs);  // (32-bit wrap intended for the row's first, partial word)
        for (uint32_t m = im; m; m &= m - 1u) {
            const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
            const float a = ap[b].x;  // (second touch: L1 hit)
            const uint32_t jr = rel0 + b;
            if (SPILL) { cv.A(pos) = a; cv.D(pos) = v.ad[rs + jr].y; cv.J(pos) = jr; }
            else if (pos < room) { pa[pos] = a; pj[pos] = jbase + jr; }  // depol is gathered for the whole batch afterwards
            pos++;
        }
        __syncwarp();
    }
    hasEv = __any_sync(FULL, ev);
    return cnt;
}

// Hand-over to the synapse pass: the staged slots that delivered in this window (t0 < arrive; staging guarantees <= t1) or
// were cleared by one of its runs (arrive negated in the pool) go to the shard's flag list — one reservation per batch.
__device__ __forceinline__ void flag_reserve(const View& v, uint32_t mine, uint32_t lane, uint32_t& at) {
    const uint32_t FULL = 0xffffffffu;
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(FULL, inc, o);
        if (lane >= (uint32_t)o) inc += y;
    }
    const uint32_t tot = __shfl_sync(FULL, inc, 31);
    uint32_t base = 0u;
    if (tot) {
        if (lane == 0) base = atomicAdd(&v.flagCtl[0], tot);
        base = __shfl_sync(FULL, base, 0);
        if (base + tot > v.flagCap && lane == 0) v.flagCtl[1] = 1u;
    }
    at = base + inc - mine;
}

// ---- staging kernel ------------------------------------------------------------------------------------------
// Keeps, per tile of 32 rows, the list of the tile's occupied slots (spike arrived, not yet cleared) in row order —
// (arrive, depol) and the slot index relative to the tile's first slot, plus a count per row — for k_neuron_pass.
// Separate from the replay so that it can run at full occupancy (it is all memory latency).
//   * A list PERSISTS: while a slot is in it its (arrive, depol) cannot change, so a window only has to (1) find the slots
//     that ARRIVED in it — the per-word arrival bounds (wordNext) name the ~1 % of index words worth looking at — and
//     (2) drop the entries the previous neuron pass cleared (it negates their arrive in the list and sets the tile's dirty
//     bit).  The merged list goes to the tile's other region (ping-pong).  Nothing else of the synapse state is touched:
//     one 4-byte bound per 32 slots + the lists themselves.
//   * A tile without a valid list (first window, after a restore / state upload, after an overflow, after too many
//     arrivals at once) is REBUILT from the index: the warp streams the tile's busy / arrived words 32 at a time, lists the
//     set bits in shared memory so that the gathers of the (arrive, depol) records are spread evenly over the lanes, and
//     counts the entries per row with a ballot walk over the tile's row ends.
#define NC_STG_WARPS 8
#define NC_STG_NEW 128  // arrivals of one tile in one window that the incremental path takes (more: the tile is rebuilt)

struct StageTile {
    uint64_t tile, rowBase, myRe, tb, te, wBeg, wEnd;
    uint32_t nr;
};

// in-flight slots of index word w whose bound says one may have landed: marks the arrived ones, renews the bound; returns their bits
__device__ __forceinline__ uint32_t stage_check_word(const View& v, uint64_t w, uint32_t infl, bool whole, uint32_t t1b, unsigned long long& busySeen) {
    uint32_t im = 0u, nmin = 0x7f800000u;
    for (uint32_t m = infl; m; m &= m - 1u) {
        const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
        const uint32_t ab = __float_as_uint(v.ad[(w << 5) + b].x);
        if (ab <= t1b) im |= 1u << b; else nmin = min(nmin, ab);
    }
    busySeen += (unsigned long long)__popc(infl);  // (per-lane count, summed over the warp at the end)
    if (im) atomicOr(&v.arrived[w], im);
    if (whole) v.wordNext[w] = nmin;  // (a word that straddles two tiles keeps its old, lower bound: it is simply looked at every window)
    return im;
}
__device__ __forceinline__ uint32_t stage_tile_mask(const StageTile& t, uint64_t w, bool& whole) {
    uint32_t m = 0xffffffffu;
    whole = true;
    if (w == t.wBeg && (t.tb & 31u)) { m &= 0xffffffffu << (uint32_t)(t.tb & 31u); whole = false; }
    if (w == t.wEnd - 1 && (t.te & 31u)) { m &= (1u << (uint32_t)(t.te & 31u)) - 1u; whole = false; }
    return m;
}

// REBUILD: the tile's list from the index, into region 0.  Returns false when the entries do not fit the region.
__device__ bool stage_rebuild(const View& v, const StageTile& t, uint32_t t1b, uint16_t* list, uint32_t lane, uint32_t& myCnt, unsigned long long& busySeen) {
    const uint32_t FULL = 0xffffffffu, lt = (1u << lane) - 1u;
    const uint64_t reg = (t.tile * 2) * v.stCap;
    uint32_t used = 0, rcur = 0;
    myCnt = 0;
    uint3 next = make_uint3(0u, 0u, 0x7f800000u);
    if (t.wBeg + lane < t.wEnd) next = make_uint3(__ldcs(v.busy + t.wBeg + lane), __ldcs(v.arrived + t.wBeg + lane), __ldcs(v.wordNext + t.wBeg + lane));
    for (uint64_t wb = t.wBeg; wb < t.wEnd; wb += 32) {
        const uint64_t w = wb + lane;
        uint32_t word = next.x, arr = next.y;
        const uint32_t nx = next.z;
        next = make_uint3(0u, 0u, 0x7f800000u);  // the next 1024 slots' words are on their way while these are gathered
        if (w + 32 < t.wEnd) next = make_uint3(__ldcs(v.busy + w + 32), __ldcs(v.arrived + w + 32), __ldcs(v.wordNext + w + 32));
        bool whole;
        word &= stage_tile_mask(t, w, whole);
        arr &= word;
        const uint32_t infl = word & ~arr;
        if (infl && nx <= t1b) arr |= stage_check_word(v, w, infl, whole, t1b, busySeen);
        __syncwarp();
        word = arr;  // the slots whose spike has arrived: these are staged
        const uint32_t c = __popc(word);
        uint32_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, inc, o);
            if (lane >= (uint32_t)o) inc += y;
        }
        const uint32_t tot = __shfl_sync(FULL, inc, 31);
        if (!tot) continue;
        busySeen += c;
        uint32_t p = inc - c;
        for (uint32_t m = word; m; m &= m - 1u) list[p++] = (uint16_t)((lane << 5) | ((uint32_t)__ffs((int)m) - 1u));
        __syncwarp();
        for (uint32_t base = 0; base < tot; base += 32) {
            const uint32_t i = base + lane;
            const bool have = i < tot;
            const uint64_t slot = (wb << 5) + (have ? (uint32_t)list[i] : 0u);
            float2 ad = make_float2(0.0f, 0.0f);
            if (have) ad = v.ad[slot];
            const bool is = have && (__float_as_uint(ad.x) - 1u < t1b);
            const uint32_t m = __ballot_sync(FULL, is);
            if (!m) continue;
            const uint32_t n = (uint32_t)__popc(m);
            if (used + n > v.stCap) return false;
            if (is) {
                const uint64_t at = reg + used + (uint32_t)__popc(m & lt);
                v.stAD[at] = ad;
                v.stJ[at] = (uint32_t)(slot - t.tb);
            }
            used += n;
            uint32_t rem = m;  // per-row counts: the entries ascend in slot order, so do the rows
            while (rem) {
                const uint64_t re_r = __shfl_sync(FULL, t.myRe, rcur);
                const uint32_t inrow = __ballot_sync(FULL, is && slot < re_r) & rem;
                if (lane == rcur) myCnt += (uint32_t)__popc(inrow);
                rem &= ~inrow;
                if (rem) rcur++;
            }
        }
        __syncwarp();
    }
    return true;
}

// INCREMENTAL: returns 0 = done (list current), 1 = rebuild the tile instead (too many arrivals at once), 2 = the merged list does not fit.
// scratch: 2 KB per warp — newJ[128] u32, newAD[128] float2, hist[129..] u32 (arrivals-below histogram of the old entries).
__device__ int stage_incremental(const View& v, const StageTile& t, uint32_t state, uint32_t t1b, uint32_t* scratch, uint32_t lane, uint32_t& myCnt,
                                 unsigned long long& busySeen) {
    const uint32_t FULL = 0xffffffffu, lt = (1u << lane) - 1u;
    uint32_t* newJ = scratch;
    float2* newAD = reinterpret_cast<float2*>(scratch + NC_STG_NEW);
    uint32_t* hist = scratch + 3 * NC_STG_NEW;  // hist[b] = alive old entries with exactly b arrivals below them (b <= m < 128)
    uint32_t m = 0;  // arrivals of this window (warp-uniform)
    // (1) the slots that arrived in this window: only the index words whose arrival bound has been reached are looked at.
    //     Every lane takes four consecutive bounds per round (one 16-byte load: 4096 slots per warp and round trip).
    const uint64_t gBeg = t.wBeg & ~3ull;
    const uint4 none = make_uint4(0x7f800000u, 0x7f800000u, 0x7f800000u, 0x7f800000u);
    const uint4* nxp = reinterpret_cast<const uint4*>(v.wordNext);
    uint4 next = (gBeg + 4 * lane < t.wEnd) ? __ldcs(nxp + (gBeg >> 2) + lane) : none;
    for (uint64_t wb = gBeg; wb < t.wEnd; wb += 128) {
        const uint64_t w0 = wb + 4 * lane;
        const uint4 nx4 = next;
        next = (w0 + 128 < t.wEnd) ? __ldcs(nxp + ((wb + 128) >> 2) + lane) : none;
        const uint32_t nxs[4] = {nx4.x, nx4.y, nx4.z, nx4.w};
        uint32_t im[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint64_t w = w0 + q;
            if (nxs[q] <= t1b && w >= t.wBeg && w < t.wEnd) {
                bool whole;
                const uint32_t msk = stage_tile_mask(t, w, whole);
                const uint32_t infl = v.busy[w] & ~v.arrived[w] & msk;
                if (infl) im[q] = stage_check_word(v, w, infl, whole, t1b, busySeen);
                else if (whole) v.wordNext[w] = 0x7f800000u;  // nothing in flight any more: the bound was stale
            }
        }
        if (!__any_sync(FULL, (im[0] | im[1] | im[2] | im[3]) != 0u)) continue;
        const uint32_t c = __popc(im[0]) + __popc(im[1]) + __popc(im[2]) + __popc(im[3]);
        uint32_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, inc, o);
            if (lane >= (uint32_t)o) inc += y;
        }
        const uint32_t tot = __shfl_sync(FULL, inc, 31);
        if (m + tot > NC_STG_NEW - 1) return 1;  // (the arrived bits are set: the rebuild finds these slots through them)
        uint32_t p = m + inc - c;
#pragma unroll
        for (int q = 0; q < 4; q++)
            for (uint32_t mm = im[q]; mm; mm &= mm - 1u) {
                const uint64_t slot = ((w0 + q) << 5) + (uint32_t)__ffs((int)mm) - 1u;
                newJ[p] = (uint32_t)(slot - t.tb);
                newAD[p] = v.ad[slot];  // (second touch: L1)
                p++;
            }
        m += tot;
        __syncwarp();
    }
    const uint32_t oldCnt = lane < t.nr ? (v.stCnt[t.rowBase + lane] & 0x3fffffffu) : 0u;
    if (m == 0u && !(state & 4u)) { myCnt = oldCnt; return 0; }  // nothing arrived, nothing was cleared: the list stands
    uint32_t n = oldCnt;
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(FULL, n, o);
    // (2) merge: old entries that are still alive keep their order, the arrivals slot in by slot index; counts per row
    const uint32_t par = (state >> 1) & 1u;
    const uint64_t src = (t.tile * 2 + par) * v.stCap, dst = (t.tile * 2 + (par ^ 1u)) * v.stCap;
    for (uint32_t k = lane; k <= m; k += 32) hist[k] = 0u;
    uint32_t myNew = 0u, myAlive = 0u;
    __syncwarp();
    for (uint32_t k = 0; k < m; k++) {  // row of arrival k = number of rows that end at or before its slot
        const uint32_t r = (uint32_t)__popc(__ballot_sync(FULL, t.myRe <= t.tb + newJ[k]) & ((t.nr >= 32u) ? FULL : ((1u << t.nr) - 1u)));
        if (lane == r) myNew++;
    }
    uint32_t aliveBefore = 0u, rcur = 0u;
    for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t i = base + lane;
        const bool have = i < n;
        float2 ad = make_float2(0.0f, 0.0f);
        uint32_t j = 0xffffffffu;
        if (have) { ad = __ldcs(v.stAD + src + i); j = __ldcs(v.stJ + src + i); }
        const bool alive = have && ad.x > 0.0f;  // (a cleared entry carries its arrive negated)
        const uint32_t am = __ballot_sync(FULL, alive);
        uint32_t lo = 0u, hi = m;  // arrivals below this entry: lower bound in the (sorted) arrival list
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (newJ[mid] < j) lo = mid + 1u; else hi = mid; }
        if (alive) {
            const uint32_t out = aliveBefore + (uint32_t)__popc(am & lt) + lo;
            if (out < v.stCap) { v.stAD[dst + out] = ad; v.stJ[dst + out] = j; }
            atomicAdd(&hist[lo], 1u);
        }
        uint32_t rem = am;
        while (rem) {
            const uint64_t re_r = __shfl_sync(FULL, t.myRe, rcur);
            const uint32_t inrow = __ballot_sync(FULL, alive && t.tb + j < re_r) & rem;
            if (lane == rcur) myAlive += (uint32_t)__popc(inrow);
            rem &= ~inrow;
            if (rem) rcur++;
        }
        aliveBefore += (uint32_t)__popc(am);
    }
    if (aliveBefore + m > v.stCap) return 2;
    __syncwarp();
    // arrival k goes after the alive old entries that have at most k arrivals below them: prefix sum of the histogram
    uint32_t carry = 0u;
    for (uint32_t k0 = 0; k0 < m; k0 += 32) {
        const uint32_t k = k0 + lane;
        const uint32_t h = k < m ? hist[k] : 0u;
        uint32_t inc = h;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, inc, o);
            if (lane >= (uint32_t)o) inc += y;
        }
        if (k < m) {
            const uint32_t out = carry + inc + k;
            v.stAD[dst + out] = newAD[k];
            v.stJ[dst + out] = newJ[k];
        }
        carry += __shfl_sync(FULL, inc, 31);
    }
    myCnt = myAlive + myNew;
    if (lane == 0) v.tileState[t.tile] = 1u | ((par ^ 1u) << 1);
    return 0;
}

__global__ void __launch_bounds__(NC_STG_WARPS * 32, 8) k_stage(View v, StepArgs s) {
    __shared__ uint32_t sscr[NC_STG_WARPS][512];  // 2 KB per warp: the rebuild's slot list (u16[1024]) / the merge's arrivals
    const uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t* scratch = sscr[threadIdx.x >> 5];
    const uint32_t t1b = __float_as_uint(s.t1);
    const uint64_t nTiles = (v.nRows + 31) >> 5;
    unsigned long long busySeen = 0ull;  // slots of the synapse arrays this warp looked at (reported by nc_index_stats)
    for (;;) {
        uint32_t t32 = 0;
        if (lane == 0) t32 = atomicAdd(&v.tileCtr[1], 1u);
        StageTile t;
        t.tile = __shfl_sync(FULL, t32, 0);
        if (t.tile >= nTiles) break;
        t.rowBase = t.tile << 5;
        t.nr = (uint32_t)min((uint64_t)32, v.nRows - t.rowBase);
        t.myRe = v.rowptr[t.rowBase + min(lane, t.nr - 1u) + 1u];  // end of this lane's row (lanes >= nr repeat the last row)
        t.tb = v.rowptr[t.rowBase];
        t.te = __shfl_sync(FULL, t.myRe, 31);
        t.wBeg = t.tb >> 5; t.wEnd = (t.te + 31) >> 5;
        const uint32_t state = v.tileState[t.tile];
        uint32_t myCnt = 0u;
        int how = ((state & 1u) && !v.stageRebuild) ? stage_incremental(v, t, state, t1b, scratch, lane, myCnt, busySeen) : 1;
        __syncwarp();
        if (how == 1) {
            const bool ok = stage_rebuild(v, t, t1b, reinterpret_cast<uint16_t*>(scratch), lane, myCnt, busySeen);
            how = ok ? 0 : 2;
            if (ok && lane == 0) v.tileState[t.tile] = 1u;
        }
        if (how == 2) {  // does not fit the region: k_neuron_pass stages this tile itself, through the index
            myCnt = 0xffffffffu;
            if (lane == 0) v.tileState[t.tile] = 0u;
        } else {
            // bit 30 of every count: which of the tile's two regions holds the list (k_neuron_pass needs nothing but the counts)
            const uint32_t st2 = __shfl_sync(FULL, lane == 0 ? *(volatile uint32_t*)(v.tileState + t.tile) : 0u, 0);
            myCnt |= ((st2 >> 1) & 1u) << 30;
        }
        if (lane < t.nr) v.stCnt[t.rowBase + lane] = myCnt;
        __syncwarp();
    }
    for (int o = 16; o > 0; o >>= 1) busySeen += __shfl_xor_sync(FULL, busySeen, o);
    if (lane == 0 && busySeen) atomicAdd(&v.stats[8], busySeen);
}

// host events of neuron q: [evLo, evHi) in the (neuron, time)-sorted list (bit set by k_mark_events for rows that have any)
__device__ __forceinline__ void host_event_range(const View& v, const StepArgs& s, uint64_t row, uint32_t q, uint32_t& evLo, uint32_t& evHi) {
    evLo = 0; evHi = 0;
    if ((s.nEv || s.nEvDev) && ((v.evMask[row >> 5] >> (row & 31u)) & 1u)) {
        const uint32_t nEv = s.nEvDev ? *s.nEvDev : s.nEv;
        uint32_t lo = 0, hi = nEv;
        while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (s.ev[mid].neuron < q) lo = mid + 1; else hi = mid; }
        evLo = lo; hi = nEv;
        while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (s.ev[mid].neuron <= q) lo = mid + 1; else hi = mid; }
        evHi = lo;
        atomicAnd(&v.evMask[row >> 5], ~(1u << (row & 31u)));  // every row is visited once per window: it takes its mark down itself
    }
}
__device__ __forceinline__ bool in_subset(const StepArgs& s, uint32_t q) {  // nc_run_neurons: only the listed neurons are run
    uint32_t lo = 0, hi = s.nSubset;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (s.subset[mid] < q) lo = mid + 1; else hi = mid; }
    return lo < s.nSubset && s.subset[lo] == q;
}

// ---- warp-per-row path: rows whose occupied slots do not fit the warp's shared-memory pool ---------------------------
__device__ void warp_row(const View& v, const StepArgs& s, uint64_t row, CandView& cv, uint32_t lane, P1Counters& ctr) {
    const uint32_t q = (uint32_t)(v.row0 + row);
    const uint64_t rs = v.rowptr[row], re = v.rowptr[row + 1];
    bool hasEv;
    const uint32_t cnt = stage_row<true>(v, s, rs, re, cv, nullptr, nullptr, 0u, 0u, lane, hasEv);
    __syncwarp();
    // the row's host events: looked up by ONE lane and broadcast — host_event_range takes the row's mark down after reading it, so
    // 32 lanes calling it for the same row only agree as long as they run in lock-step (found by the CPU emulator, where they do not)
    uint32_t evLo = 0, evHi = 0;
    if (lane == 0) host_event_range(v, s, row, q, evLo, evHi);
    evLo = __shfl_sync(0xffffffffu, evLo, 0);
    evHi = __shfl_sync(0xffffffffu, evHi, 0);
    NeuronState n;
    float lfS0;
    {
        float2 pa = v.potAct[row];
        n.pot = pa.x; n.act = pa.y;
        n.lastRan = v.lastRan[row]; n.lastFire = v.lastFire[row]; n.actStart = v.actStart[row];
        n.firings = v.firings[row];
        n.sched = __uint_as_float(0x7fc00000u);
        for (uint32_t e = evLo; e < evHi; e++)
            if (s.ev[e].kind == 2u && (s.ev[e].index_or_flags & 1u)) n.sched = s.ev[e].time;
        lfS0 = n.lastFire;
        if (lane == 0) v.lfStart[row] = n.lastFire;
    }
    // ---- replay in-window events in canonical order ----
    if (hasEv || evHi > evLo || (s.sweep & NC_SWEEP_START)) {
        float curT = s.t0;
        unsigned long long curC = 0;
        bool first = true;  // events at exactly t0 are allowed for host events only
        for (;;) {
            EvPick best;
            best.t = INFINITY; best.code = ~0ull; best.src = 0xffffffffu;
            for (uint32_t c = lane; c < cnt; c += 32) {
                float a = fabsf(cv.A(c));
                if (a > s.t0) {  // delivery in this window (a <= t1 by staging)
                    unsigned long long code = (1ull << 32) | cv.J(c);  // in-row index: ascends with the presynaptic ID
                    if ((first || pick_less(curT, curC, a, code)) && pick_less(a, code, best.t, best.code)) {
                        best.t = a; best.code = code; best.src = c;
                    }
                }
                float tR = add32(a, 2.0f);  // Neuron::transfer's requeue (NeuCor.cpp:665)
                if (tR > s.t0 && tR <= s.t1) {
                    unsigned long long code = (2ull << 32);
                    if ((first || pick_less(curT, curC, tR, code)) && pick_less(tR, code, best.t, best.code)) {
                        best.t = tR; best.code = code; best.src = c;
                    }
                }
            }
            if ((s.sweep & NC_SWEEP_START) && lane == 0 && (first || pick_less(curT, curC, s.t0, 2ull << 32)) &&
                pick_less(s.t0, 2ull << 32, best.t, best.code)) {
                best.t = s.t0; best.code = 2ull << 32; best.src = 0xfffffffeu;  // runAll: queued at t0 (NeuCor.cpp:596)
            }
            for (uint32_t e = evLo + lane; e < evHi; e += 32) {
                nc_event ev = s.ev[e];
                unsigned long long code = ev.kind == 0u ? (unsigned long long)ev.index_or_flags : (2ull << 32);
                bool after = first ? (ev.time >= s.t0) : pick_less(curT, curC, ev.time, code);
                if (after && ev.time <= s.t1 && pick_less(ev.time, code, best.t, best.code)) {
                    best.t = ev.time; best.code = code; best.src = 0x80000000u | e;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                float ot = __shfl_xor_sync(0xffffffffu, best.t, o);
                unsigned long long oc = __shfl_xor_sync(0xffffffffu, best.code, o);
                uint32_t os = __shfl_xor_sync(0xffffffffu, best.src, o);
                if (pick_less(ot, oc, best.t, best.code)) { best.t = ot; best.code = oc; best.src = os; }
            }
            if (best.code == ~0ull) break;
            {
                uint32_t rank = (uint32_t)(best.code >> 32), k = (uint32_t)best.code;
                if (rank == 0) {  // InputFirer::run → Neuron::fire, no update, ignores refractory (NeuCor.cpp:326-331,643-645)
                    n.lastFire = best.t;
                    n.firings++;
                    ctr.fires++;
                    if (lane == 0) emit_fire(v, q, best.t, k, q);
                } else if (rank == 1) {  // Synapse::run → Neuron::transfer (NeuCor.cpp:718-726,663-666)
                    ctr.deliveries++;
                    warp_neuron_run(v, n, cv, cnt, rs, q, best.t, (1u << 30) | q, k, NC_SENT | (1u << 29) | cv.J(best.src), lane, ctr);
                } else {
                    warp_neuron_run(v, n, cv, cnt, rs, q, best.t, (2u << 30) | q, 0u, NC_SENT | (2u << 29), lane, ctr);
                }
            }
            __syncwarp();
            curT = best.t; curC = best.code; first = false;
        }
    }
    if (s.sweep & NC_SWEEP_END) warp_neuron_run(v, n, cv, cnt, rs, q, s.t1, (3u << 30) | q, 0u, NC_SENT | (3u << 29), lane, ctr);
    if (lane == 0) {
        v.potAct[row] = make_float2(n.pot, n.act);
        v.lastRan[row] = n.lastRan; v.lastFire[row] = n.lastFire; v.firings[row] = n.firings;
    }
    __syncwarp();
    // slots that delivered in this window or were cleared by it: over to the synapse pass
    for (uint32_t base = 0; base < cnt; base += 32) {
        const uint32_t c = base + lane;
        bool f = false;
        if (c < cnt) { const float a = cv.A(c); f = (a < 0.0f) || (a > s.t0); }
        uint32_t at;
        flag_reserve(v, f ? 1u : 0u, lane, at);
        if (f && at < v.flagCap) { FlagEnt fe; fe.slot = (uint32_t)(rs + cv.J(c)); fe.row = (uint32_t)row; fe.lfStart = lfS0; fe.inRow = cv.J(c); v.flagList[at] = fe; }
    }
    __syncwarp();
}

// ---- lane-per-row path: one lane replays one neuron; its occupied slots sit in a slice of the warp's pool ------------------
// All 32 lanes run this together and stay converged: the replay proceeds in ROUNDS, in every round each lane executes the
// next event of its own neuron (or idles once it has none left), and the loops over the staged slots run to the largest
// count of the batch.  The end-of-window sweep run is simply every lane's last event, so lanes with few events sweep while
// others are still delivering.  The pass over a lane's slots adds the active slots' contributions for the event being
// executed — the reference's own loop, one slot after the other in ascending presynaptic ID with a double -> float rounding
// per addition (NeuCor.cpp:688-700); the neuron's next event is then picked among the few slots that carry one.
struct LanePick { float t; unsigned long long code; uint32_t src; };

__device__ __forceinline__ void pick_from_slot(const View& v, const StepArgs& s, uint32_t rowOff, const uint32_t* J, uint32_t c, float a,
                                               bool first, float curT, unsigned long long curC, LanePick& nx) {
    if (a > s.t0) {  // delivery in this window (a <= t1 by staging)
        // equal-time deliveries to one neuron are ordered by presynaptic ID = by the slot's index within the row (rows ascend in it)
        const unsigned long long code = (1ull << 32) | (J[c] - rowOff);
        if ((first || pick_less(curT, curC, a, code)) && pick_less(a, code, nx.t, nx.code)) { nx.t = a; nx.code = code; nx.src = c; }
    }
    const float tR = add32(a, 2.0f);  // Neuron::transfer's requeue (NeuCor.cpp:665)
    if (tR > s.t0 && tR <= s.t1) {
        const unsigned long long code = (2ull << 32);
        if ((first || pick_less(curT, curC, tR, code)) && pick_less(tR, code, nx.t, nx.code)) { nx.t = tR; nx.code = code; nx.src = c; }
    }
}
__device__ __forceinline__ void pick_from_host(const StepArgs& s, uint32_t evLo, uint32_t evHi, bool first, float curT, unsigned long long curC,
                                               LanePick& nx) {
    if (s.sweep & NC_SWEEP_START)  // runAll: queued at t0 (NeuCor.cpp:596)
        if ((first || pick_less(curT, curC, s.t0, 2ull << 32)) && pick_less(s.t0, 2ull << 32, nx.t, nx.code)) { nx.t = s.t0; nx.code = 2ull << 32; nx.src = 0xfffffffeu; }
    for (uint32_t e = evLo; e < evHi; e++) {
        const nc_event ev = s.ev[e];
        const unsigned long long code = ev.kind == 0u ? (unsigned long long)ev.index_or_flags : (2ull << 32);
        const bool after = first ? (ev.time >= s.t0) : pick_less(curT, curC, ev.time, code);
        if (after && ev.time <= s.t1 && pick_less(ev.time, code, nx.t, nx.code)) { nx.t = ev.time; nx.code = code; nx.src = 0x80000000u | e; }
    }
}

// J holds slot indices relative to `tb`, the first slot of the tile's first row.
__device__ void lanes_replay(const View& v, const StepArgs& s, bool valid, uint64_t row, uint64_t tb, float* A, const float* D, const uint32_t* J,
                             uint32_t cnt, bool hasEv, P1Counters& ctr) {
    const uint32_t FULL = 0xffffffffu;
    const unsigned long long NONE = ~0ull, SWEEP = 3ull << 32;
    if (!valid) { cnt = 0; hasEv = false; row = 0; }
    const uint32_t q = (uint32_t)(v.row0 + row);
    const uint32_t rowOff = valid ? (uint32_t)(v.rowptr[row] - tb) : 0u;  // J - rowOff = index within the row
    uint32_t evLo = 0, evHi = 0;
    NeuronState n;
    n.pot = 0.f; n.act = 0.f; n.lastRan = 0.f; n.lastFire = 0.f; n.actStart = 0.f; n.firings = 0u; n.sched = __uint_as_float(0x7fc00000u);
    float lfS0 = 0.0f;
    if (valid) {
        host_event_range(v, s, row, q, evLo, evHi);
        float2 pa = v.potAct[row];
        n.pot = pa.x; n.act = pa.y;
        n.lastRan = v.lastRan[row]; n.lastFire = v.lastFire[row]; n.actStart = v.actStart[row];
        n.firings = v.firings[row];
        for (uint32_t e = evLo; e < evHi; e++)
            if (s.ev[e].kind == 2u && (s.ev[e].index_or_flags & 1u)) n.sched = s.ev[e].time;
        lfS0 = n.lastFire;
        v.lfStart[row] = n.lastFire;
    }
    const uint32_t maxCnt = __reduce_max_sync(FULL, cnt);
    const bool sweepEnd = (s.sweep & NC_SWEEP_END) != 0;
    float actT = 0.0f; uint32_t actF = 0u; bool ran = false;
    // ---- first event of every lane ----
    // which of this lane's staged slots carry an event in this window (delivery or requeue): usually one or two of dozens, so
    // they are remembered as a bit mask (first 64 slots; beyond that the tail is scanned) and only those are looked at when
    // the neuron's next event is picked
    unsigned long long em = 0ull;
    bool evTail = false;
    uint32_t nFlag = 0u;  // this lane's slots that deliver in the window or get cleared by it (handed to the synapse pass at the end)
    if (__any_sync(FULL, hasEv))
        for (uint32_t c = 0; c < maxCnt; c++)
            if (hasEv && c < cnt) {
                const float a = A[c];
                const float tR = add32(a, 2.0f);
                nFlag += (a > s.t0) ? 1u : 0u;
                if ((a > s.t0) || (tR > s.t0 && tR <= s.t1)) {
                    if (c < 64u) em |= 1ull << c; else evTail = true;
                }
            }
    auto pick_slots = [&](bool first, float curT, unsigned long long curC, LanePick& out) {
        unsigned long long m = em;
        while (m) {
            const uint32_t c = (uint32_t)__ffsll((long long)m) - 1u;
            m &= m - 1ull;
            pick_from_slot(v, s, rowOff, J, c, fabsf(A[c]), first, curT, curC, out);
        }
        if (evTail)
            for (uint32_t c = 64u; c < cnt; c++) pick_from_slot(v, s, rowOff, J, c, fabsf(A[c]), first, curT, curC, out);
    };
    LanePick nx;
    nx.t = INFINITY; nx.code = NONE; nx.src = 0xffffffffu;
    pick_slots(true, s.t0, 0ull, nx);
    if (valid) pick_from_host(s, evLo, evHi, true, s.t0, 0ull, nx);
    bool swept = !sweepEnd || !valid;
    if (nx.code == NONE && !swept) { nx.t = s.t1; nx.code = SWEEP; swept = true; }
    // ---- rounds ----
    while (__any_sync(FULL, nx.code != NONE)) {
        const LanePick cur = nx;
        nx.t = INFINITY; nx.code = NONE; nx.src = 0xffffffffu;
        const bool act = cur.code != NONE;
        const uint32_t rank = (uint32_t)(cur.code >> 32), k = (uint32_t)cur.code;
        const float T = cur.t;
        bool running = false;
        float dT = 0.0f;
        uint32_t rk1 = 0u, k2 = 0u, sentinel = 0u;
        if (act) {
            if (rank == 0u) {  // InputFirer::run → Neuron::fire, no update, ignores refractory (NeuCor.cpp:326-331,643-645)
                n.lastFire = T;
                n.firings++;
                ctr.fires++;
                emit_fire(v, q, T, k, q);
            } else {
                if (rank == 1u) {  // Synapse::run → Neuron::transfer (NeuCor.cpp:718-726,663-666)
                    ctr.deliveries++;
                    rk1 = (1u << 30) | q; k2 = k; sentinel = NC_SENT | (1u << 29) | (J[cur.src] - rowOff);
                } else {           // rank 2: queued Neuron::run; rank 3: end-of-window sweep
                    rk1 = (rank << 30) | q; k2 = 0u; sentinel = NC_SENT | (rank << 29);
                }
                running = neuron_run_begin(n, T, dT);  // false when no time has passed (NeuCor.cpp:626)
                if (running) ctr.runs++;
            }
        }
        const bool pickMore = act && rank != 3u;  // the sweep is a lane's last event
        float np = n.pot;
        double E = 0.0;
        if (running && cnt) E = exp_glibc(mul64(0.3702, (double)dT));
        if (__any_sync(FULL, running && cnt))
            for (uint32_t c = 0; c < maxCnt; c++) {
                if (c < cnt && running) {
                    const float araw = A[c];
                    if (araw > 0.0f) {                  // not cleared earlier in this window
                        const float off = sub32(T, araw);
                        if (off > 0.0f) {               // arrived
                            ctr.visits++;
                            np = (float)add64((double)np, chain_term(dT, D[c], E));
                            if (2.0f < off) {           // NeuCor.cpp:697 — the slot becomes idle; leave the when-and-why for the synapse pass
                                A[c] = -araw;
                                nFlag++;
                                const uint64_t sidx = tb + J[c];
                                v.ad[sidx] = make_float2(__uint_as_float(sentinel), T);
                            }
                        }
                    }
                }
            }
        if (pickMore) {
            pick_slots(false, cur.t, cur.code, nx);
            pick_from_host(s, evLo, evHi, false, cur.t, cur.code, nx);
        }
        if (running) {
            const bool fired = neuron_run_finish(n, np, T, dT, false);
            actT = T; actF = n.firings; ran = true;
            if (fired) { ctr.fires++; emit_fire(v, q, T, rk1, k2); }
        }
        if (act && nx.code == NONE && !swept) { nx.t = s.t1; nx.code = SWEEP; swept = true; }
    }
    if (valid) {
        if (ran) n.act = neuron_activity(actF, actT, n.actStart);
        v.potAct[row] = make_float2(n.pot, n.act);
        v.lastRan[row] = n.lastRan; v.lastFire[row] = n.lastFire; v.firings[row] = n.firings;
    }
    // slots that delivered in this window (t0 < arrive) or were cleared by it (negated in the pool): over to the synapse pass
    if (__any_sync(FULL, nFlag != 0u)) {
        uint32_t at;
        flag_reserve(v, nFlag, threadIdx.x & 31u, at);
        if (nFlag)
            for (uint32_t c = 0; c < cnt; c++) {
                const float a = A[c];
                if ((a < 0.0f) || (a > s.t0)) {
                    if (at < v.flagCap) { FlagEnt fe; fe.slot = (uint32_t)(tb + J[c]); fe.row = (uint32_t)row; fe.lfStart = lfS0; fe.inRow = J[c] - rowOff; v.flagList[at] = fe; }
                    at++;
                }
            }
    }
}

// Neuron pass.  A warp takes tiles of 32 consecutive rows.  It stages the occupied slots of as many rows as fit its
// shared-memory pool (cooperative, coalesced row scans), then every lane replays ONE of those neurons on its own
// (lane-per-row: the per-neuron math — ordered accumulation, powf/exp, threshold, AP — is not replicated across lanes);
// a row that alone exceeds the pool takes the warp-per-row path with the global spill area.
// Two register budgets of the same code: MINB = 6 (80 registers; pool of 1024 staged slots per warp) for busy networks,
// MINB = 8 (64 registers, a few spills in the replay; pool of 512) when rows hold few occupied slots and the pass is bound
// by the latency of the row scan — a third more resident warps.  The host picks per window from the last window's counters;
// the pool size only changes how rows are batched, never a result.
template <int MINB>
__global__ void __launch_bounds__(NC_WARPS_PER_BLOCK * 32, MINB) k_neuron_pass(View v, StepArgs s, ncx::XchgArgs xa) {
    extern __shared__ unsigned char smem[];
    math_tables_to_shared();
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    const uint32_t cap = s.candCap;
    // pool of staged slots: arrival time and depolarisation factor (touched on every visit) in shared memory, the slot index
    // (needed only where a slot delivers or is cleared) in an L2-resident per-warp scratch — 8 instead of 12 bytes of shared
    // memory per staged slot buys two more resident blocks per SM
    float* sA = reinterpret_cast<float*>(smem) + (size_t)wib * 3 * cap;
    float* sD = sA + cap;
    uint32_t* sJ = reinterpret_cast<uint32_t*>(sD + cap);
    const uint64_t gw = (uint64_t)blockIdx.x * NC_WARPS_PER_BLOCK + wib;
    CandView cv;
    cv.a = sA; cv.d = sD; cv.j = sJ; cv.cap = cap;
    cv.sa = v.spillA + gw * v.spillPerWarp; cv.sd = v.spillD + gw * v.spillPerWarp; cv.sj = v.spillJ + gw * v.spillPerWarp;
    P1Counters ctrW = {0, 0, 0, 0};  // warp-uniform counts of the warp-per-row path
    P1Counters ctrL = {0, 0, 0, 0};  // this lane's counts of the lane-per-row path
    const uint64_t nTiles = (v.nRows + 31) >> 5;

    uint32_t t32 = 0;  // tiles are claimed dynamically (their cost varies with the activity of their neurons), one tile ahead:
    if (lane == 0) t32 = atomicAdd(&v.tileCtr[0], 1u);  // the counter's round trip overlaps with the work on the current tile
    for (;;) {
        const uint64_t tile = __shfl_sync(0xffffffffu, t32, 0);
        if (tile >= nTiles) break;
        if (lane == 0) t32 = atomicAdd(&v.tileCtr[0], 1u);
        const uint64_t rowBase = tile << 5;
        const uint32_t nr = (uint32_t)min((uint64_t)32, v.nRows - rowBase);
        const uint64_t tb = v.rowptr[rowBase];  // staged slot indices are relative to the tile's first slot (rows < 2^27 slots)
        uint32_t r = 0;
        const uint32_t stcRaw = (lane < nr) ? v.stCnt[rowBase + lane] : 0u;
        const uint32_t stc = stcRaw & 0x3fffffffu;
        if (__shfl_sync(0xffffffffu, stcRaw, 0) != 0xffffffffu) {
            // ---- the tile was staged by k_stage: copy batches of rows from its region into the pool and replay, lane = row ----
            const bool inSub = lane < nr && (!s.subset || in_subset(s, (uint32_t)(v.row0 + rowBase + lane)));
            uint32_t inc = stc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= (uint32_t)o) inc += y;
            }
            const uint32_t offL = inc - stc;  // this row's first entry within the region
            const uint64_t reg = (tile * 2 + ((__shfl_sync(0xffffffffu, stcRaw, 0) >> 30) & 1u)) * v.stCap;  // the region that holds the tile's current list
            bool anyDead = false;
            while (r < nr) {
                const uint32_t base = __shfl_sync(0xffffffffu, offL, r);
                const bool fits = lane >= r && lane < nr && (inc - base <= cap);
                const uint32_t fm = __ballot_sync(0xffffffffu, fits) >> r;
                const uint32_t nb = min(fm == 0xffffffffu ? 32u : (uint32_t)__ffs((int)~fm) - 1u, nr - r);  // rows r .. r+nb-1 fit the pool together
                if (nb == 0u) {  // a row whose occupied slots alone exceed the pool: staged and replayed by the whole warp, through the index;
                    // the tile's list does not learn what that run cleared, so it is rebuilt by the next staging pass
                    if (__shfl_sync(0xffffffffu, inSub ? 1u : 0u, r)) { warp_row(v, s, rowBase + r, cv, lane, ctrW); if (lane == 0) v.tileState[tile] = 0u; }
                    r++;
                    continue;
                }
                const uint32_t used = __shfl_sync(0xffffffffu, inc, r + nb - 1u) - base;
                const float2* src = v.stAD + reg + base;
                const uint32_t* srcJ = v.stJ + reg + base;
                for (uint32_t i = lane; i < used; i += 32) { const float2 x = __ldcs(src + i); sA[i] = x.x; sD[i] = x.y; sJ[i] = __ldcs(srcJ + i); }
                __syncwarp();
                const bool mine = inSub && lane >= r && lane < r + nb;
                const uint32_t o = mine ? offL - base : 0u;
                lanes_replay(v, s, mine, rowBase + lane, tb, sA + o, sD + o, sJ + o, mine ? stc : 0u, mine && stc != 0u, ctrL);
                __syncwarp();
                // the entries this replay cleared (arrive negated in the pool) are marked in the persistent list: the next staging pass drops them
                for (uint32_t i = lane; i < used; i += 32) {
                    const float a = sA[i];
                    if (a < 0.0f) { v.stAD[reg + base + i].x = a; anyDead = true; }
                }
                __syncwarp();
                r += nb;
            }
            if (__any_sync(0xffffffffu, anyDead) && lane == 0) atomicOr(&v.tileState[tile], 4u);
        }
        while (r < nr) {  // (tiles that did not fit their staging region are staged here, through the busy-slot index, row by row)
            // ---- batch: stage rows r, r+1, ... while their occupied slots fit the pool; the i-th staged row goes to lane i ----
            uint32_t used = 0, nb = 0, myRow = 0, myOff = 0, myCnt = 0;
            bool myEv = false, heavy = false;
            while (r < nr) {
                const uint64_t row = rowBase + r;
                if (s.subset && !in_subset(s, (uint32_t)(v.row0 + row))) { r++; continue; }
                const uint64_t rs = v.rowptr[row], re = v.rowptr[row + 1];
                bool ev;
                const uint32_t c = stage_row<false>(v, s, rs, re, cv, sA + used, sJ + used, (uint32_t)(rs - tb), cap - used, lane, ev);
                if (c > cap - used) { heavy = (nb == 0); break; }  // does not fit: close the batch (an over-long row goes alone)
                if (lane == nb) { myRow = r; myOff = used; myCnt = c; myEv = ev; }
                used += c; nb++; r++;
            }
            __syncwarp();
            if (heavy) { warp_row(v, s, rowBase + r, cv, lane, ctrW); r++; continue; }
            // depolarisation factors of all staged slots of the batch: one round of independent gathers instead of a dependent
            // load per 128-slot group during staging
            for (uint32_t i = lane; i < used; i += 32) sD[i] = v.ad[tb + sJ[i]].y;
            __syncwarp();
            lanes_replay(v, s, lane < nb, rowBase + myRow, tb, sA + myOff, sD + myOff, sJ + myOff, myCnt, myEv, ctrL);
            __syncwarp();
        }
    }
    unsigned long long c4[4] = {ctrL.fires, ctrL.deliveries, ctrL.runs, ctrL.visits};
#pragma unroll
    for (int i = 0; i < 4; i++)
        for (int o = 16; o > 0; o >>= 1) c4[i] += __shfl_xor_sync(0xffffffffu, c4[i], o);
    if (lane == 0) {  // the warp-per-row path counted the same (warp-uniform) events in every lane
        c4[0] += ctrW.fires; c4[1] += ctrW.deliveries; c4[2] += ctrW.runs; c4[3] += ctrW.visits;
        if (c4[0]) atomicAdd(&v.stats[0], c4[0]);
        if (c4[1]) atomicAdd(&v.stats[1], c4[1]);
        if (c4[2]) atomicAdd(&v.stats[6], c4[2]);
        if (c4[3]) atomicAdd(&v.stats[7], c4[3]);
    }
    if (xa.world > 1u) ncx::push_fires_tail(v, xa);  // the fire exchange: this shard's records straight into every peer's memory
}

// ------------------------------------------------------------------------------------------------
// Fire index (bitmask + per-neuron record lists) over the gathered records of all shards
// ------------------------------------------------------------------------------------------------
// Block b of the gathered buffer starts with its header unit {count, overflow}; record i of block b is unit
// b*gStride + 1 + i, and that unit index is what head[] / next[] hold.
__device__ __forceinline__ uint32_t block_count(const View& v, const StepArgs& s, uint32_t b) {
    const uint32_t* hdr = reinterpret_cast<const uint32_t*>(v.gRecs + (uint64_t)b * s.gStride);
    return min(hdr[0], s.gStride - 1u);
}
__global__ void k_index_build(View v, StepArgs s, ncx::XchgArgs xa) {
    if (xa.world > 1u) ncx::wait_flags_head(xa, 0u);  // every shard's records of this window have landed in this shard's gather buffer
    const uint32_t b = blockIdx.y;
    const uint32_t n = block_count(v, s, b);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t idx = b * s.gStride + 1u + i;
        const uint32_t nrn = v.gRecs[idx].neuron;
        v.next[idx] = atomicExch(&v.head[nrn], (int32_t)idx);
        atomicOr(&v.mask[nrn >> 5], 1u << (nrn & 31u));
    }
}
__global__ void k_index_reset(View v, StepArgs s) {
    const uint32_t b = blockIdx.y;
    const uint32_t n = block_count(v, s, b);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t nrn = v.gRecs[b * s.gStride + 1u + i].neuron;
        v.head[nrn] = -1;
        v.mask[nrn >> 5] = 0u;
    }
}
// End of a window: publish (or, for replay, accumulate) the shard's counters and its exchange header into `out`
// (10 x u64: the 8 nc_step_stats counters, fire count, overflow flag), then clear both for the next window.
__global__ void k_finish_step(View v, unsigned long long* out, int accumulate, unsigned long long* win, ncx::XchgArgs xa, ncr::RandTables rtab,
                              uint32_t* randState) {
    const uint32_t i = threadIdx.x;
    if (i < 8) {
        unsigned long long x = v.stats[i];
        out[i] = accumulate ? out[i] + x : x;
        win[i] = x;  // this window's own counters (the rand() stream moves on by win[5], summed over the shards)
        v.stats[i] = 0ull;
    } else if (i == 9) {  // running totals for nc_index_stats: busy slots visited, flag-list entries
        v.stats[10] += v.stats[8]; v.stats[8] = 0ull;
        v.stats[11] += min(v.flagCtl[0], v.flagCap);
    } else if (i == 8) {
        out[8] = v.localHdr[0];
        const unsigned long long ovf = (unsigned long long)v.localHdr[1] | ((unsigned long long)v.flagCtl[1] << 1);  // bit 0: fire records, bit 1: flag list
        out[9] = accumulate ? (out[9] | ovf) : ovf;
    }
    __syncwarp();
    if (i == 8) { v.localHdr[0] = 0u; v.localHdr[1] = 0u; v.tileCtr[0] = 0u; v.tileCtr[1] = 0u; v.flagCtl[0] = 0u; v.flagCtl[1] = 0u; }
    if (xa.world > 1u) {  // the window's counters to every shard (the hidden rand() count moves every shard's stream on)
        if (i == 8) { win[8] = out[8]; win[9] = out[9]; }
        ncx::push_counters_tail(win, xa);
    } else if (randState) {  // single shard: the hidden rand() calls of this window move the device-resident stream on right here
        __shared__ uint32_t st[32];
        __syncwarp();
        if (i < NC_RS_K) st[i] = randState[i];
        __syncwarp();
        ncr::rs_jump(rtab, st, win[5], i);
        if (i < NC_RS_K) randState[i] = st[i];
    }
}

// ------------------------------------------------------------------------------------------------
// Synapse pass
// ------------------------------------------------------------------------------------------------
// Only the synapses something happened to are touched.  Three sources name them, and every eventful
// slot is owned by exactly one:
//   k_syn_loads    for every neuron that fired (any shard): its out-synapses in this shard (CSC index) — Synapse::fire.
//                  Skips slots whose row fired (k_syn_rows owns them) and slots on the flag list (k_syn_flagged owns them).
//   k_syn_rows     for every neuron of this shard that fired: all its in-synapses (post-fire plasticity, plus whatever
//                  else happened to them in the window).
//   k_syn_flagged  the slots the neuron pass flagged (delivery in the window / cleared by one of its runs), unless their
//                  row fired.
// k_syn_loads runs first: it decides "flagged" from `arrive` as the neuron pass left it, which only it may have changed for
// the slots it owns.  Each slot is resolved by resolve_slot() — all of the window's operations on it in canonical order.
__device__ __forceinline__ bool fired_g(const View& v, uint32_t n) { return (__ldg(v.mask + (n >> 5)) >> (n & 31u)) & 1u; }

__device__ __forceinline__ void flush_syn_counters(const View& v, uint32_t* cnt) {
#pragma unroll
    for (int i = 0; i < 5; i++) {
        uint32_t x = cnt[i];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        cnt[i] = x;
    }
    if ((threadIdx.x & 31u) == 0u) {
        if (cnt[0]) atomicAdd(&v.stats[2], (unsigned long long)cnt[0]);
        if (cnt[1]) atomicAdd(&v.stats[3], (unsigned long long)cnt[1]);
        if (cnt[2]) atomicAdd(&v.stats[4], (unsigned long long)cnt[2]);
        if (cnt[3]) atomicAdd(&v.stats[5], (unsigned long long)cnt[3]);
    }
}

#define NC_SYN_THREADS 128

// Work item = (fire record, 128-entry chunk of its list): one thread per synapse, so that a window with a few hundred fires
// still fills the machine.  The record at the head of its neuron's list stands for the neuron (a neuron can fire more than
// once in a window: input firers ignore the refractory period).
__global__ void __launch_bounds__(NC_SYN_THREADS) k_syn_loads(View v, StepArgs s) {
    const uint32_t b = blockIdx.y;
    const uint64_t items = (uint64_t)block_count(v, s, b) * v.cprLoads;
    uint32_t cnt[5] = {0, 0, 0, 0, 0};  // loads accepted, dropped, plasticity calls, hidden rand, deliveries
    for (uint64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const uint32_t idx = b * s.gStride + 1u + (uint32_t)(item / v.cprLoads);
        const uint32_t p = v.gRecs[idx].neuron;
        if (v.head[p] != (int32_t)idx) continue;
        const uint64_t e1 = v.cscPtr[p + 1];
        {
            const uint64_t e = v.cscPtr[p] + (item % v.cprLoads) * NC_SYN_THREADS + threadIdx.x;
            if (e >= e1) continue;
            const SlotRow sr = v.cscEnt[e];
            const uint32_t q = (uint32_t)(v.row0 + sr.row);
            if (fired_g(v, q)) continue;  // k_syn_rows
            const uint32_t ab = __float_as_uint(v.ad[sr.slot].x);
            const float a = __uint_as_float(ab);
            if ((ab & NC_SENT) || (ab != 0u && a > s.t0 && a <= s.t1)) continue;  // k_syn_flagged
            // only loads can apply: neither the row context nor the plasticity tables are needed
            resolve_slot(v, s, sr.slot, 0u, q, v.rec[sr.slot], ab, true, false, 0.0f, cnt);
        }
    }
    flush_syn_counters(v, cnt);
}
__global__ void __launch_bounds__(NC_SYN_THREADS) k_syn_rows(View v, StepArgs s) {
    math_tables_to_shared();
    const uint32_t b = blockIdx.y;
    const uint64_t items = (uint64_t)block_count(v, s, b) * v.cprRows;
    uint32_t cnt[5] = {0, 0, 0, 0, 0};
    for (uint64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const uint32_t idx = b * s.gStride + 1u + (uint32_t)(item / v.cprRows);
        const uint32_t q = v.gRecs[idx].neuron;
        if (q < v.row0 || q >= v.row0 + v.nRows || v.head[q] != (int32_t)idx) continue;
        const uint64_t row = q - v.row0;
        const uint64_t rs = v.rowptr[row], re = v.rowptr[row + 1];
        const float lfS = v.lfStart[row];
        {
            const uint64_t j = rs + (item % v.cprRows) * NC_SYN_THREADS + threadIdx.x;
            if (j >= re) continue;
            const SynRec r0 = v.rec[j];
            resolve_slot(v, s, j, (uint32_t)(j - rs), q, r0, __float_as_uint(v.ad[j].x), fired_g(v, r0.pre & 0x7fffffffu), true, lfS, cnt);
        }
    }
    flush_syn_counters(v, cnt);
}
__global__ void __launch_bounds__(256) k_syn_flagged(View v, StepArgs s) {
    math_tables_to_shared();
    const uint32_t n = min(v.flagCtl[0], v.flagCap);
    uint32_t cnt[5] = {0, 0, 0, 0, 0};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const FlagEnt fe = v.flagList[i];
        const uint32_t q = (uint32_t)(v.row0 + fe.row);
        if (fired_g(v, q)) continue;  // k_syn_rows
        const SynRec r0 = v.rec[fe.slot];
        resolve_slot(v, s, fe.slot, fe.inRow, q, r0, __float_as_uint(v.ad[fe.slot].x), fired_g(v, r0.pre & 0x7fffffffu), false, fe.lfStart, cnt);
    }
    flush_syn_counters(v, cnt);
}

// Marks (set = 1) or unmarks the rows of this shard that have host events in this window.
__global__ void k_mark_events(View v, const nc_event* ev, uint32_t nEv, const uint32_t* nEvDev, int set) {
    if (nEvDev) nEv = *nEvDev;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nEv; i += gridDim.x * blockDim.x) {
        uint64_t n = ev[i].neuron;
        if (n < v.row0 || n >= v.row0 + v.nRows) continue;
        uint64_t row = n - v.row0;
        if (set) atomicOr(&v.evMask[row >> 5], 1u << (row & 31u));
        else v.evMask[row >> 5] = 0u;
    }
}
// The window's event list when the background events live on the device: the host's (input-firer) events, sorted by neuron,
// merged with the background events of the run that fall into the window — sorted by neuron, input events before background
// events of the same neuron, each group in its own order (what the host's stable sort of [inputs..., background...] gives).
// One block; both lists are short.  bgCtl[0] = number of background events of the run; outCount <- merged length.
__global__ void __launch_bounds__(1024) k_merge_events(View v, const nc_event* host, uint32_t nHost, const nc_event* bg, const uint32_t* bgCtl, float t0, float t1,
                                                      int strict, nc_event* out, uint32_t outCap, uint32_t* outCount) {
    const uint32_t SM = 2048;  // background events kept in shared memory (neuron; ~0u when outside the window); longer lists read global memory
    __shared__ uint32_t sN[SM];
    __shared__ uint32_t sIn;
    const uint32_t nBg = bgCtl[0];
    if (threadIdx.x == 0) sIn = 0u;
    auto inWin = [&](const nc_event& e) { return (strict ? e.time > t0 : e.time >= t0) && e.time <= t1; };
    for (uint32_t k = threadIdx.x; k < min(nBg, SM); k += blockDim.x) { const nc_event e = bg[k]; sN[k] = inWin(e) ? e.neuron : 0xffffffffu; }
    __syncthreads();
    auto key = [&](uint32_t k) -> uint32_t { if (k < SM) return sN[k]; const nc_event e = bg[k]; return inWin(e) ? e.neuron : 0xffffffffu; };
    // background event j -> position = (in-window background events before it) + (host events with neuron <= its neuron)
    for (uint32_t j = threadIdx.x; j < nBg; j += blockDim.x) {
        const uint32_t nj = key(j);
        if (nj == 0xffffffffu) continue;
        uint32_t before = 0;
        for (uint32_t k = 0; k < j; k++) before += key(k) != 0xffffffffu ? 1u : 0u;
        uint32_t lo = 0, hi = nHost;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (host[mid].neuron <= nj) lo = mid + 1; else hi = mid; }
        if (before + lo < outCap) out[before + lo] = bg[j];
        atomicAdd(&sIn, 1u);
    }
    // host event i -> position = i + (in-window background events with a smaller neuron; the list is sorted by neuron)
    for (uint32_t i = threadIdx.x; i < nHost; i += blockDim.x) {
        const nc_event e = host[i];
        uint32_t below = 0;
        for (uint32_t k = 0; k < nBg; k++) { const uint32_t nk = key(k); below += (nk != 0xffffffffu && nk < e.neuron) ? 1u : 0u; }
        if (i + below < outCap) out[i + below] = e;
    }
    __syncthreads();
    const uint32_t n = min(nHost + sIn, outCap);
    if (threadIdx.x == 0) *outCount = n;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {  // mark the rows of this shard that have events in the window
        const uint64_t q = out[i].neuron;
        if (q >= v.row0 && q < v.row0 + v.nRows) atomicOr(&v.evMask[(q - v.row0) >> 5], 1u << ((q - v.row0) & 31u));
    }
}

// ------------------------------------------------------------------------------------------------
// Small utility kernels
// ------------------------------------------------------------------------------------------------
__global__ void k_init_neurons(View v) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.nRows) return;
    v.potAct[i] = make_float2(-70.0f, 0.0f);  // NeuCor.cpp:389,394
    v.lastRan[i] = 0.0f;
    v.lastFire[i] = __uint_as_float(0x7fc00000u);  // NAN, NeuCor.cpp:392
    v.lfStart[i] = __uint_as_float(0x7fc00000u);
    v.actStart[i] = 0.0f;
    v.firings[i] = 0u;
}
__global__ void k_init_synapses(View v, const uint32_t* pre, const float* weight, const float* length, const unsigned char* inh) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.S) return;
    v.ad[i] = make_float2(0.0f, 0.0f); v.lastStart[i] = 0.0f;
    SynRec r;
    r.pre = pre[i] | (inh[i] ? 0x80000000u : 0u);
    r.weight = weight[i];
    r.lastArr = __uint_as_float(0xff800000u);  // -INFINITY, NeuCor.cpp:469
    r.delay = mul32(length[i], 2.0f);           // length * AP_speed, NeuCor.cpp:485,733
    v.rec[i] = r;
}
__global__ void k_fill_i32(int32_t* p, uint64_t n, int32_t val) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = val;
}
// Out-synapse index (CSC over global presynaptic IDs) of the shard's rows, built once at upload: count, (host) prefix sum, fill.
// The order of a neuron's entries is arbitrary — every slot is resolved on its own.
__global__ void k_csc_count(View v, uint32_t* cnt) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < v.S; j += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd(&cnt[v.rec[j].pre & 0x7fffffffu], 1u);
}
__global__ void k_csc_fill(View v, const uint64_t* ptr, uint32_t* cursor, SlotRow* ent) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t gw = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nW = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t row = gw; row < v.nRows; row += nW) {
        const uint64_t rs = v.rowptr[row], re = v.rowptr[row + 1];
        for (uint64_t j = rs + lane; j < re; j += 32) {
            const uint32_t p = v.rec[j].pre & 0x7fffffffu;
            SlotRow sr; sr.slot = (uint32_t)j; sr.row = (uint32_t)row;
            ent[ptr[p] + atomicAdd(&cursor[p], 1u)] = sr;
        }
    }
}
// Position-weighted 64-bit checksums of the shard's state (the six fields the parity fixtures pin, tests/helpers.py):
// sum over i of bits(x[i]) * (i + 1) * 0x9E3779B97F4A7C15 mod 2^64 for pot, act, lastFire, weight, arrive (+ depol of the
// busy slots), lastArr.  Integer sums commute, so the result does not depend on the thread layout.
__global__ void k_state_signature(View v, unsigned long long* out) {
    const unsigned long long C = 0x9E3779B97F4A7C15ull;
    unsigned long long a[6] = {0, 0, 0, 0, 0, 0};
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nT = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = tid; i < v.nRows; i += nT) {
        const float2 pa = v.potAct[i];
        const unsigned long long k = (i + 1) * C;
        a[0] += (unsigned long long)__float_as_uint(pa.x) * k;
        a[1] += (unsigned long long)__float_as_uint(pa.y) * k;
        a[2] += (unsigned long long)__float_as_uint(v.lastFire[i]) * k;
    }
    for (uint64_t j = tid; j < v.S; j += nT) {
        const unsigned long long k = (j + 1) * C;
        const uint32_t ab = __float_as_uint(v.ad[j].x);
        a[3] += (unsigned long long)__float_as_uint(v.rec[j].weight) * k;
        a[4] += (unsigned long long)ab * k;
        if (v.ad[j].x != 0.0f) a[4] += (unsigned long long)__float_as_uint(v.ad[j].y) * k;
        a[5] += (unsigned long long)__float_as_uint(v.rec[j].lastArr) * k;
    }
#pragma unroll
    for (int f = 0; f < 6; f++) {
        unsigned long long x = a[f];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31u) == 0u && x) atomicAdd(&out[f], x);
    }
}
__global__ void k_extract_ad(View v, int which, float* out) {  // which: 0 arrive, 1 depol, 2 weight, 3 lastArr
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < v.S; j += (uint64_t)gridDim.x * blockDim.x)
        out[j] = which == 0 ? v.ad[j].x : which == 1 ? v.ad[j].y : which == 2 ? v.rec[j].weight : v.rec[j].lastArr;
}
// state injection (checkpoint resume): which 0 arrive, 1 depol, 2 weight, 3 lastArr
__global__ void k_inject_syn(View v, int which, const float* in) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < v.S; j += (uint64_t)gridDim.x * blockDim.x) {
        const float x = in[j];
        if (which == 0) v.ad[j].x = x; else if (which == 1) v.ad[j].y = x; else if (which == 2) v.rec[j].weight = x; else v.rec[j].lastArr = x;
    }
}
// the network as uploaded, back from the device records: which 0 pre, 1 length (delay / 2, exact), 2 inhibitory flag
__global__ void k_extract_net(View v, int which, uint32_t* out) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < v.S; j += (uint64_t)gridDim.x * blockDim.x) {
        const SynRec r = v.rec[j];
        out[j] = which == 0 ? (r.pre & 0x7fffffffu) : which == 1 ? __float_as_uint(mul32(r.delay, 0.5f)) : (r.pre >> 31);
    }
}
// busy-slot index from `arrive` (after a restore, or for state loaded from a file)
__global__ void k_rebuild_busy(View v, uint64_t words) {
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t m = 0u;
        const uint64_t j0 = w << 5;
        for (uint32_t b = 0; b < 32u && j0 + b < v.S; b++) m |= (v.ad[j0 + b].x != 0.0f ? 1u : 0u) << b;
        v.busy[w] = m;
        v.arrived[w] = 0u;   // unknown: the staging kernel looks at every busy slot of the word once (bound 0) and sorts them out
        v.wordNext[w] = 0u;
    }
}
__global__ void k_reset_activities(View v, float now) {  // Neuron::resetActivity, NeuCor.cpp:460
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.nRows) return;
    v.firings[i] = 0u; v.actStart[i] = now;
    float2 pa = v.potAct[i]; pa.y = 0.0f; v.potAct[i] = pa;
}
// VoltageDetector::getVoltage's sequential float sum (NeuCor.cpp:360-365) — one thread, exact order.
// One warp.  The sum itself is the reference's: sequential, in list order, one float rounding per addition (NeuCor.cpp:360-365) —
// every lane carries the same running sum; what is parallel is the memory side: 32 potentials are gathered per round trip
// (the next 32 already in flight) and handed round by shuffles, instead of one dependent load pair per neuron on one thread.
__global__ void k_detector_mean(View v, const uint32_t* near, uint32_t n, float* out) {
    const uint32_t lane = threadIdx.x & 31u;
    float avg = 0.0f;
    float nxt = lane < n ? v.potAct[near[lane] - v.row0].x : 0.0f;
    for (uint32_t base = 0; base < n; base += 32u) {
        const float cur = nxt;
        const uint32_t ahead = base + 32u + lane;
        nxt = ahead < n ? v.potAct[near[ahead] - v.row0].x : 0.0f;
        const uint32_t cnt = min(32u, n - base);  // (the same in every lane: the shuffles below stay converged)
        for (uint32_t j = 0; j < cnt; j++) avg = add32(avg, __shfl_sync(0xffffffffu, cur, (int)j));
    }
    if (lane == 0) *out = div32(avg, (float)n);
}
// ---- NeuCor_Renderer's "Statistics" panel on the device (/root/reference/src/NeuCor_Renderer.cpp:1733-1876) -----------------------
// span index = floor(spans * (x - range_min) / range) with the reference's float typing (int * float, float / float, floor);
// values outside [0, spans) are counted as below / above.  which: 0 = Neuron::activity() of every neuron (:1749-1761),
// 1 = Synapse::getWeight() of every synapse (:1801-1815).  out[0..spans) bins, out[spans] below, out[spans+1] above.
__global__ void k_render_histogram(View v, int which, uint32_t spans, float rmin, float range, unsigned int* out) {
    const uint64_t n = which ? v.S : v.nRows;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float x = which ? v.rec[i].weight : v.potAct[i].y;
        const float f = floorf(div32(mul32((float)(int)spans, sub32(x, rmin)), range));
        uint32_t slot;
        if (!(f >= 0.0f)) slot = spans;                  // negative, or NaN (the reference's (int)NaN is INT_MIN)
        else if (f >= (float)spans) slot = spans + 1u;
        else slot = (uint32_t)f;
        atomicAdd(&out[slot], 1u);
    }
}
// Raster frame by the GUI's rule: neurons with now - lastFire < runSpeed (Renderer.cpp:1856-1862); IDs in no particular order.
__global__ void k_render_raster(View v, float now, float runSpeed, uint32_t cap, uint32_t* ids, unsigned int* count) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < v.nRows; i += (uint64_t)gridDim.x * blockDim.x) {
        if (sub32(now, v.lastFire[i]) < runSpeed) {
            const uint32_t at = atomicAdd(count, 1u);
            if (at < cap) ids[at] = (uint32_t)(v.row0 + i);
        }
    }
}
// Synapse::getPrePot / getPostPot (NeuCor.cpp:547-567)
__device__ __forceinline__ float render_behaviour(float valf) {  // AP_RENDER_BEHAVIOUR, NeuCor.cpp:547-550
    float val = (float)fmin(fmax((double)valf, 0.0), 0.7);
    if ((double)val < 0.5) {
        float x = (float)div64((double)val, 5.0);
        float p = (x >= 1.17549435e-38f) ? powf_pos(x, 3.0f) : 0.0f;
        return (float)mul64(mul64(8.0, 1000.0), (double)p);
    }
    float x = (float)sub64(3.5, (double)mul32(5.0f, val));
    return (float)mul64(8.0, (double)mul32(x, x));
}
__global__ void k_synapse_pots(View v, float now, float* prePot, float* postPot) {
    math_tables_to_shared();
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.S) return;
    float a = v.ad[i].x, w = v.rec[i].weight, dl = v.rec[i].delay;
    float pre = 0.0f, post = 0.0f;
    if (a != 0.0f) {
        pre = mul32(render_behaviour(div32(sub32(now, v.lastStart[i]), dl)), w);
        if (now < a) post = mul32(render_behaviour(div32(sub32(a, now), dl)), w);
    }
    if (prePot) prePot[i] = pre;
    if (postPot) postPot[i] = post;
}

// ------------------------------------------------------------------------------------------------
// Host side of the C ABI
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

struct TapeStep { float t0, t1; int sweep; uint64_t evOff; uint32_t nEv; uint32_t units; uint32_t fires; bool bgDraw, bgActive, bgStrict; ncr::BgArgs bg;
                  int32_t randIdx; };  // randIdx >= 0: the application moved libc's stream before this window (entry of the taped rand() states)

// NCCL is bound at run time (dlopen) so that single-GPU users need no NCCL at all; only the five entry points below are used.
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, nc_comm_id, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

struct nc_engine {
    nc_config cfg;
    std::string err;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    bool uploaded = false;
    View v;
    nc_event* dEv = nullptr; uint32_t evCap = 0;
    nc_event* hEvPinned = nullptr; uint32_t hEvCap = 0;
    // per-window result block: 10 x u64 per shard (8 counters, fire count, overflow flag)
    unsigned long long* dSig = nullptr; unsigned long long* hSig = nullptr;  // nc_state_signature
    // device-resident rand() stream and background firing (rand_stream.cuh)
    ncr::RandTables rtab = {nullptr, nullptr};
    uint32_t *dRandState = nullptr, *dRandNext = nullptr;  // 31-word stream state (oldest value first) and the buffer the next one is written to
    uint32_t* hRand = nullptr;               // pinned: [0..31) state after the last window, [32..36) background control block
    bool randOn = false, hRandFresh = false;
    uint32_t* dDraws = nullptr; uint64_t drawsCap = 0; uint32_t* dHitMask = nullptr;
    nc_event *dBgTmp = nullptr, *dBgEv = nullptr; uint32_t bgCap = 0;
    uint32_t *dCand = nullptr, *dCandRaw = nullptr; uint32_t candListCap = 0;
    uint32_t* dBgCtl = nullptr;              // [0] events of this shard, [1] hits (network), [2] overflow, [3] draws consumed
    bool bgActive = false, bgFirstWindow = false;
    nc_event* dEvMerged = nullptr; uint32_t mergedCap = 0; uint32_t* dMergedCount = nullptr;
    bool bgPendingTape = false; ncr::BgArgs bgPendingArgs = {};
    bool pendingBgActive = false, pendingBgStrict = false;  // of the window in flight (for the tape)
    unsigned long long* dWin = nullptr; unsigned long long* dWinAll = nullptr;  // the last window's own counters (replay accumulates dOut)
    uint32_t* snapRand = nullptr; bool snapRandOn = false;
    unsigned long long* dOut = nullptr;     // this shard's block
    unsigned long long* dOutAll = nullptr;  // world blocks (world > 1)
    unsigned long long* hOut = nullptr;     // pinned, world blocks
    uint32_t* hHdrAll = nullptr;            // pinned, world x 4: gathered exchange headers
    FireRec* dGather = nullptr;             // world blocks of (fireCap + 1) units (world > 1)
    // exchange transport (world > 1): NCCL communicator or caller-provided all-gather
    void* comm = nullptr;
    nc_allgather_fn xchgFn = nullptr; void* xchgCtx = nullptr;
    uint32_t xchgUnits = 1u + 1024u;        // units per shard moved by the fire exchange; grows on demand
    // peer exchange over NVLink (peer_exchange.cuh): this shard's arena and every shard's arena as mapped here
    bool p2p = false;
    char* arena = nullptr; char* peerBase[NC_MAX_WORLD] = {nullptr};
    size_t offCnt = 0, offG[2] = {0, 0};
    uint32_t xseq = 0; uint32_t* dPushCtr = nullptr;
    ncx::XchgArgs xa = {};                  // the exchange arguments of the window in flight (world = 0: none)
    uint32_t lastCounts[NC_MAX_WORLD] = {0}; uint32_t lastStride = 0;
    bool pending = false;                   // nc_step_launch issued, nc_step_collect outstanding
    StepArgs pendingArgs;
    float lr = 1.0f, preF = 0.13f, postF = 0.30f, preD = 0.75f, postD = 0.65f;
    float minDelay = INFINITY;
    uint32_t candCap = 704, candCapSparse = 512, candCapBig = 1024, grid1 = 0, grid1s = 0, grid1b = 0, gridStage = 0;
    int forceVariant = 0;                   // 0 auto, 1 big, 2 dense, 3 sparse
    double lastSlotsPerRun = 1e9;           // occupied slots visited per neuron run in the last window (picks the variant)
    size_t smem1 = 0, smem1s = 0, smem1b = 0;
    uint64_t* dCscPtr = nullptr; SlotRow* dCscEnt = nullptr;  // out-synapse index (CSC over global presynaptic IDs)
    uint64_t launches = 0;
    void* dScratch = nullptr; size_t scratchBytes = 0;  // read-back / render entry points
    cudaEvent_t* tick = nullptr;            // per-kernel timing of a replay: recorded between the staging kernel and the neuron pass
    float breakdown[4] = {0, 0, 0, 0};      // last per-kernel replay: k_stage, k_neuron_pass, fire exchange, synapse kernels [ms, summed]
    // tape
    bool taping = false; std::vector<TapeStep> tape; nc_event* dTape = nullptr; uint64_t tapeCap = 0, tapeUsed = 0; uint32_t tapeMaxSteps = 0;
    // rand() states the host handed over while taping (nc_rand_set_state: the application drew from libc between two run() calls):
    // a replay puts them back at the same windows, so that it sees the stream the live run saw
    std::vector<uint32_t> tapeRand; uint32_t* dTapeRand = nullptr; bool tapeRandPending = false; uint32_t tapeRandPend[32] = {0};
    // snapshot
    struct Snap { float2* ad; SynRec* rec; float *lastStart, *lastRan, *lastFire, *lfStart, *actStart; float2* potAct; uint32_t *firings, *busy, *arrived, *wordNext; bool valid; } snap = {};
    uint64_t busyWords = 0, nTiles = 0;
    int smCount = 148;
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            e->err = std::string(#call) + ": " + cudaGetErrorString(_e);                           \
            return NC_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

static int fail(nc_engine* e, int code, const std::string& msg) { e->err = msg; return code; }

extern "C" const char* nc_global_error(void) { return g_err.c_str(); }
extern "C" const char* nc_last_error(const nc_engine* e) { return e ? e->err.c_str() : g_err.c_str(); }
extern "C" uint64_t nc_launch_count(const nc_engine* e) { return e->launches; }

extern "C" int nc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int nc_create(const nc_config* cfg, nc_engine** out) {
    if (!cfg || !out) { g_err = "nc_create: null argument"; return NC_ERR_INVALID; }
    int n = nc_device_count();
    if (n <= 0) { g_err = "nc_create: no usable CUDA device (this library has no CPU path)"; return NC_ERR_NO_DEVICE; }
    if (cfg->device < 0 || cfg->device >= n) { g_err = "nc_create: device ordinal out of range"; return NC_ERR_INVALID; }
    if (cfg->world < 1 || cfg->world > NC_MAX_WORLD || cfg->rank < 0 || cfg->rank >= cfg->world) { g_err = "nc_create: bad rank/world"; return NC_ERR_INVALID; }
    nc_engine* e = new nc_engine();
    e->cfg = *cfg;
    memset(&e->v, 0, sizeof(View));
    cudaError_t ce = cudaSetDevice(cfg->device);
    if (ce == cudaSuccess && !getenv("NC_KEEP_L2_FETCH")) {
        // both passes gather single 8-/16-byte records scattered over tens of GB: ask L2 to fetch 32-byte sectors, not
        // 64/128-byte groups, from HBM (a hint; measured effect in profiles/README.md)
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
        cudaGetLastError();
    }
    if (ce == cudaSuccess) {
        if (cfg->stream) { e->stream = (cudaStream_t)cfg->stream; }
        else { ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking); e->ownStream = true; }
    }
    if (ce == cudaSuccess) ce = cudaMemcpyToSymbol(d_POWF_LOG2_TAB, NC_POWF_LOG2_TAB, sizeof(NC_POWF_LOG2_TAB));
    if (ce == cudaSuccess) ce = cudaMemcpyToSymbol(d_EXP2F_TAB, NC_EXP2F_TAB, sizeof(NC_EXP2F_TAB));
    if (ce == cudaSuccess) ce = cudaMemcpyToSymbol(d_EXP_TAB, NC_EXP_TAB, sizeof(NC_EXP_TAB));
    if (ce == cudaSuccess) ce = cudaMallocHost(&e->hOut, (size_t)cfg->world * 10 * sizeof(unsigned long long));
    if (ce == cudaSuccess) ce = cudaMallocHost(&e->hHdrAll, (size_t)cfg->world * 4 * sizeof(uint32_t));
    if (ce == cudaSuccess) ce = cudaMalloc(&e->dSig, 6 * sizeof(unsigned long long));
    if (ce == cudaSuccess) ce = cudaMallocHost(&e->hSig, 6 * sizeof(unsigned long long));
    if (ce == cudaSuccess) {
        std::vector<uint32_t> T, J;
        ncr::build_rand_tables(T, J);
        uint32_t *dT = nullptr, *dJ = nullptr;
        ce = cudaMalloc(&dT, T.size() * 4);
        if (ce == cudaSuccess) ce = cudaMalloc(&dJ, J.size() * 4);
        if (ce == cudaSuccess) ce = cudaMemcpy(dT, T.data(), T.size() * 4, cudaMemcpyHostToDevice);
        if (ce == cudaSuccess) ce = cudaMemcpy(dJ, J.data(), J.size() * 4, cudaMemcpyHostToDevice);
        e->rtab.T = dT; e->rtab.J = dJ;
    }
    if (ce == cudaSuccess) ce = cudaMalloc(&e->dRandState, 32 * 4);
    if (ce == cudaSuccess) ce = cudaMalloc(&e->dRandNext, 32 * 4);
    if (ce == cudaSuccess) ce = cudaMalloc(&e->snapRand, 32 * 4);
    if (ce == cudaSuccess) ce = cudaMallocHost(&e->hRand, 40 * 4);
    if (ce == cudaSuccess) ce = cudaMalloc(&e->dBgCtl, 8 * 4);  // [0..4) control block of the walk, [4] candidate counter of the generator
    if (ce == cudaSuccess) ce = cudaMemset(e->dBgCtl, 0, 8 * 4);
    if (ce == cudaSuccess) ce = cudaMalloc(&e->dMergedCount, 4);
    if (ce == cudaSuccess) ce = cudaMalloc(&e->dWin, 10 * sizeof(unsigned long long));
    if (ce == cudaSuccess) ce = cudaMemset(e->dWin, 0, 10 * sizeof(unsigned long long));
    if (ce == cudaSuccess) ce = cudaMalloc(&e->dWinAll, (size_t)cfg->world * 10 * sizeof(unsigned long long));
    if (ce == cudaSuccess) ce = cudaMalloc(&e->dOut, 10 * sizeof(unsigned long long));
    if (ce == cudaSuccess) ce = cudaMemset(e->dOut, 0, 10 * sizeof(unsigned long long));
    if (ce == cudaSuccess) ce = cudaMalloc(&e->dOutAll, (size_t)cfg->world * 10 * sizeof(unsigned long long));
    if (ce == cudaSuccess) ce = cudaMalloc(&e->v.stats, 12 * sizeof(unsigned long long));
    if (ce == cudaSuccess) ce = cudaMemset(e->v.stats, 0, 12 * sizeof(unsigned long long));
    if (ce == cudaSuccess) ce = cudaMalloc(&e->v.tileCtr, 2 * sizeof(uint32_t));
    if (ce == cudaSuccess) ce = cudaMemset(e->v.tileCtr, 0, 2 * sizeof(uint32_t));
    cudaDeviceProp prop;
    if (ce == cudaSuccess) ce = cudaGetDeviceProperties(&prop, cfg->device);
    if (ce != cudaSuccess) { g_err = std::string("nc_create: ") + cudaGetErrorString(ce); delete e; return NC_ERR_CUDA; }
    e->smCount = prop.multiProcessorCount;
    if (cfg->cand_smem) e->candCap = cfg->cand_smem;
    *out = e;
    return NC_OK;
}

static void free_all(nc_engine* e) {
    View& v = e->v;
    cudaFree((void*)v.rowptr); cudaFree(v.rec); cudaFree(v.ad);
    cudaFree(v.lastStart); cudaFree(v.potAct); cudaFree(v.lastRan); cudaFree(v.lastFire); cudaFree(v.lfStart);
    cudaFree(v.actStart); cudaFree(v.firings); cudaFree(v.localHdr); cudaFree(v.head); cudaFree(v.next); cudaFree(v.mask); cudaFree(v.evMask); cudaFree(v.busy); cudaFree(v.arrived); cudaFree(v.wordNext); cudaFree(v.stCnt); cudaFree(v.stAD); cudaFree(v.stJ); cudaFree(v.tileState); cudaFree(v.flagList); cudaFree(v.flagCtl); cudaFree(e->dCscPtr); cudaFree(e->dCscEnt);
    cudaFree(v.spillA); cudaFree(v.spillD); cudaFree(v.spillJ);
    cudaFree(e->dGather);
    cudaFree(e->dEv); cudaFree(e->dTape); cudaFree(e->dTapeRand);
    auto& s = e->snap;
    cudaFree(s.ad); cudaFree(s.rec); cudaFree(s.lastStart); cudaFree(s.lastRan);
    cudaFree(s.lastFire); cudaFree(s.lfStart); cudaFree(s.actStart); cudaFree(s.potAct); cudaFree(s.firings); cudaFree(s.busy); cudaFree(s.arrived); cudaFree(s.wordNext);
}

extern "C" void nc_destroy(nc_engine* e) {
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    cudaStreamSynchronize(e->stream);
    free_all(e);
    cudaFree(e->v.stats); cudaFree(e->v.tileCtr); cudaFree(e->dOut); cudaFree(e->dOutAll);
    cudaFreeHost(e->hOut); cudaFreeHost(e->hHdrAll); cudaFreeHost(e->hEvPinned); cudaFree(e->dSig); cudaFreeHost(e->hSig);
    cudaFree((void*)e->rtab.T); cudaFree((void*)e->rtab.J); cudaFree(e->dRandState); cudaFree(e->dRandNext); cudaFree(e->snapRand); cudaFreeHost(e->hRand);
    cudaFree(e->dDraws); cudaFree(e->dHitMask); cudaFree(e->dBgTmp); cudaFree(e->dBgEv); cudaFree(e->dCand); cudaFree(e->dCandRaw); cudaFree(e->dBgCtl); cudaFree(e->dEvMerged); cudaFree(e->dMergedCount);
    cudaFree(e->dWin); cudaFree(e->dWinAll);
    for (int r = 0; r < e->cfg.world; r++) if (e->p2p && r != e->cfg.rank && e->peerBase[r]) cudaIpcCloseMemHandle(e->peerBase[r]);
    cudaFree(e->arena); cudaFree(e->dPushCtr); cudaFree(e->dScratch);
    if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
    if (e->ownStream) cudaStreamDestroy(e->stream);
    delete e;
}

extern "C" int nc_set_plasticity(nc_engine* e, float lr, float preF, float postF, float preD, float postD) {
    auto okbase = [](float x) { return x > 0.0f && x < INFINITY && x >= 1.17549435e-38f; };
    if (!okbase(preD) || !okbase(postD)) return fail(e, NC_ERR_INVALID, "nc_set_plasticity: trace decays must be positive normal floats");
    e->lr = lr; e->preF = preF; e->postF = postF; e->preD = preD; e->postD = postD;
    return NC_OK;
}

// Validation of a device-resident CSR (nc_upload_network_device): row lengths, ascending presynaptic IDs, ranges,
// smallest delay. out[0] = error bits, out[1] = max row length, out[2] = min delay (float bits; positive floats order as uints)
__global__ void k_validate_csr(uint64_t nGlobal, uint64_t nRows, const uint64_t* rowptr, const uint32_t* pre, const float* length,
                               unsigned int* out) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t gw = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nW = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    unsigned int err = 0, maxRow = 0, minD = 0x7f800000u;
    for (uint64_t row = gw; row < nRows; row += nW) {
        uint64_t rs = rowptr[row], re = rowptr[row + 1];
        if (re < rs) { err |= 1u; continue; }
        if (re - rs >= (1ull << 27)) { err |= 2u; continue; }
        maxRow = max(maxRow, (unsigned int)(re - rs));
        for (uint64_t j = rs + lane; j < re; j += 32) {
            uint32_t p = pre[j];
            if (p >= nGlobal) err |= 4u;
            if (j > rs && p <= pre[j - 1]) err |= 8u;
            float d = __fmul_rn(length[j], 2.0f);
            if (!(d > 0.0f)) err |= 16u; else minD = min(minD, __float_as_uint(d));
        }
    }
    if (err) atomicOr(&out[0], err);
    atomicMax(&out[1], maxRow);
    atomicMin(&out[2], minD);
}

// common tail of the two upload entry points: allocate state, copy the CSR (host->device or device->device), initialise
static int upload_common(nc_engine* e, uint64_t nGlobal, uint64_t row0, uint64_t nRows, uint64_t S, uint64_t maxRow, float minDelay,
                         const uint64_t* rowptr, const uint32_t* pre, const float* weight, const float* length, const uint8_t* inh,
                         cudaMemcpyKind kind) {
    View& v = e->v;
    v.nGlobal = nGlobal; v.row0 = row0; v.nRows = nRows; v.S = S;
    const uint64_t S1 = std::max<uint64_t>(S, 1), N1 = std::max<uint64_t>(nRows, 1), G1 = std::max<uint64_t>(nGlobal, 1);
    CK(cudaMalloc((void**)&v.rowptr, (nRows + 1) * 8));
    const uint64_t SP = ((S1 + 127) / 128 + 1) * 128;  // 16-byte row loads may touch the rest of the last 128-slot group
    CK(cudaMalloc(&v.rec, S1 * sizeof(SynRec))); CK(cudaMalloc(&v.ad, SP * 8));
    CK(cudaMemsetAsync(v.ad, 0, SP * 8, e->stream));
    // event index: busy-slot bitmap (1 bit per slot), flag list (neuron pass -> synapse pass), out-synapse index (CSC)
    const uint64_t busyWords = SP / 32 + 64;
    e->busyWords = busyWords;
    CK(cudaMalloc(&v.busy, busyWords * 4));
    CK(cudaMemsetAsync(v.busy, 0, busyWords * 4, e->stream));
    CK(cudaMalloc(&v.arrived, busyWords * 4));
    CK(cudaMemsetAsync(v.arrived, 0, busyWords * 4, e->stream));
    CK(cudaMalloc(&v.wordNext, busyWords * 4));
    k_fill_i32<<<(unsigned)((busyWords + 255) / 256), 256, 0, e->stream>>>(reinterpret_cast<int32_t*>(v.wordNext), busyWords, 0x7f800000); e->launches++;  // +inf: nothing in flight
    v.flagCap = e->cfg.flag_capacity ? e->cfg.flag_capacity : (uint32_t)std::min<uint64_t>(std::max<uint64_t>(S1 / 8 + (1u << 16), 1u << 20), S1 + 1024);
    CK(cudaMalloc(&v.flagList, (uint64_t)v.flagCap * sizeof(FlagEnt)));
    CK(cudaMalloc(&v.flagCtl, 2 * sizeof(uint32_t)));
    CK(cudaMemsetAsync(v.flagCtl, 0, 2 * sizeof(uint32_t), e->stream));
    CK(cudaMalloc(&v.lastStart, S1 * 4));
    CK(cudaMalloc(&v.potAct, N1 * 8)); CK(cudaMalloc(&v.lastRan, N1 * 4)); CK(cudaMalloc(&v.lastFire, N1 * 4));
    CK(cudaMalloc(&v.lfStart, N1 * 4)); CK(cudaMalloc(&v.actStart, N1 * 4)); CK(cudaMalloc(&v.firings, N1 * 4));
    v.fireCap = e->cfg.fire_capacity ? e->cfg.fire_capacity : (uint32_t)std::min<uint64_t>(4 * nRows + 1024, (1u << 28) / (uint32_t)e->cfg.world);
    const uint64_t blockUnits = (uint64_t)v.fireCap + 1;  // header unit + records
    e->xchgUnits = (uint32_t)std::min<uint64_t>(e->xchgUnits, blockUnits);  // never gather past a shard's block
    CK(cudaMalloc(&v.localHdr, blockUnits * sizeof(FireRec)));
    v.localRecs = reinterpret_cast<FireRec*>(v.localHdr) + 1;
    CK(cudaMemsetAsync(v.localHdr, 0, 16, e->stream));
    v.gRecs = reinterpret_cast<FireRec*>(v.localHdr);  // (world > 1: the gathered blocks — peer-exchange arena, or ensure_gather's buffer)
    CK(cudaMalloc(&v.head, G1 * 4)); CK(cudaMalloc(&v.next, (uint64_t)e->cfg.world * blockUnits * 4));
    CK(cudaMalloc(&v.mask, ((G1 + 31) / 32) * 4));
    CK(cudaMemsetAsync(v.mask, 0, ((G1 + 31) / 32) * 4, e->stream));
    CK(cudaMalloc(&v.evMask, ((N1 + 31) / 32) * 4));
    CK(cudaMemsetAsync(v.evMask, 0, ((N1 + 31) / 32) * 4, e->stream));
    const uint32_t* dPre = pre; const float* dW = weight; const float* dLen = length; const unsigned char* dInh = inh;
    uint32_t* tmpPre = nullptr; float *tmpW = nullptr, *tmpLen = nullptr; unsigned char* tmpInh = nullptr;
    CK(cudaMemcpyAsync((void*)v.rowptr, rowptr, (nRows + 1) * 8, kind, e->stream));
    if (S) {
        if (kind == cudaMemcpyHostToDevice) {  // (device-resident arrays are read where they lie)
            CK(cudaMalloc(&tmpPre, S1 * 4)); CK(cudaMalloc(&tmpW, S1 * 4)); CK(cudaMalloc(&tmpLen, S1 * 4)); CK(cudaMalloc(&tmpInh, S1));
            CK(cudaMemcpyAsync(tmpPre, pre, S * 4, kind, e->stream));
            CK(cudaMemcpyAsync(tmpW, weight, S * 4, kind, e->stream));
            CK(cudaMemcpyAsync(tmpLen, length, S * 4, kind, e->stream));
            CK(cudaMemcpyAsync(tmpInh, inh, S, kind, e->stream));
            dPre = tmpPre; dW = tmpW; dLen = tmpLen; dInh = tmpInh;
        }
        k_init_synapses<<<(unsigned)((S + 255) / 256), 256, 0, e->stream>>>(v, dPre, dW, dLen, dInh);
        e->launches++;
    }
    if (nRows) { k_init_neurons<<<(unsigned)((nRows + 255) / 256), 256, 0, e->stream>>>(v); e->launches++; }
    k_fill_i32<<<(unsigned)((G1 + 255) / 256), 256, 0, e->stream>>>(v.head, G1, -1); e->launches++;
    CK(cudaGetLastError());
    {   // out-synapse index
        if (S >= (1ull << 32)) return fail(e, NC_ERR_INVALID, "nc_upload_network: at most 2^32-1 synapses per shard");
        uint32_t* dCnt = nullptr;
        CK(cudaMalloc(&dCnt, G1 * 4));
        CK(cudaMemsetAsync(dCnt, 0, G1 * 4, e->stream));
        CK(cudaMalloc(&e->dCscPtr, (G1 + 1) * 8));
        CK(cudaMalloc(&e->dCscEnt, S1 * sizeof(SlotRow)));
        if (S) { k_csc_count<<<e->smCount * 16, 256, 0, e->stream>>>(v, dCnt); e->launches++; }
        std::vector<uint32_t> hCnt(G1);
        std::vector<uint64_t> hPtr(G1 + 1);
        CK(cudaMemcpyAsync(hCnt.data(), dCnt, G1 * 4, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        hPtr[0] = 0;
        uint32_t maxOut = 0;
        for (uint64_t i = 0; i < G1; i++) { hPtr[i + 1] = hPtr[i] + hCnt[i]; maxOut = std::max(maxOut, hCnt[i]); }
        v.cprLoads = std::max<uint32_t>(1u, (maxOut + NC_SYN_THREADS - 1) / NC_SYN_THREADS);
        v.cprRows = std::max<uint32_t>(1u, (uint32_t)((maxRow + NC_SYN_THREADS - 1) / NC_SYN_THREADS));
        CK(cudaMemcpyAsync(e->dCscPtr, hPtr.data(), (G1 + 1) * 8, cudaMemcpyHostToDevice, e->stream));
        CK(cudaMemsetAsync(dCnt, 0, G1 * 4, e->stream));
        if (S) { k_csc_fill<<<e->smCount * 16, 256, 0, e->stream>>>(v, e->dCscPtr, dCnt, e->dCscEnt); e->launches++; }
        CK(cudaStreamSynchronize(e->stream));
        cudaFree(dCnt);
        v.cscPtr = e->dCscPtr; v.cscEnt = e->dCscEnt;
        CK(cudaGetLastError());
    }
    // launch geometry: persistent grids sized to the SM count x resident blocks per SM
    // the warp's shared-memory pool of staged slots: shared by the rows of a lane-per-row batch, so not tied to the row length
    // Three builds of the same neuron pass, picked per window from the last window's occupancy (fill_args): a 1024-slot pool at
    // 4 blocks per SM for busy networks (a tile of 32 rows then mostly fits ONE batch), 704 slots at 6 blocks, 512 at 8 blocks.
    // nc_config.cand_smem pins all three to one size (tests).
    if (e->cfg.cand_smem) { e->candCap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(e->candCap, 32), 1024); e->candCapBig = e->candCapSparse = e->candCap; }
    else { e->candCap = 704; e->candCapSparse = 512; e->candCapBig = 1024; }
    e->smem1 = (size_t)NC_WARPS_PER_BLOCK * 3 * e->candCap * 4;
    e->smem1s = (size_t)NC_WARPS_PER_BLOCK * 3 * e->candCapSparse * 4;
    e->smem1b = (size_t)NC_WARPS_PER_BLOCK * 3 * e->candCapBig * 4;
    CK(cudaFuncSetAttribute(k_neuron_pass<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem1b));
    CK(cudaFuncSetAttribute(k_neuron_pass<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem1));
    CK(cudaFuncSetAttribute(k_neuron_pass<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem1s));
    int occ1 = 1, occ1s = 1, occ1b = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1b, k_neuron_pass<4>, NC_WARPS_PER_BLOCK * 32, e->smem1b));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, k_neuron_pass<6>, NC_WARPS_PER_BLOCK * 32, e->smem1));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1s, k_neuron_pass<8>, NC_WARPS_PER_BLOCK * 32, e->smem1s));
    uint64_t needBlocks = ((nRows + 31) / 32 + NC_WARPS_PER_BLOCK - 1) / NC_WARPS_PER_BLOCK;  // one tile of 32 rows per warp at a time
    e->grid1b = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(needBlocks, (uint64_t)e->smCount * std::max(occ1b, 1)));
    e->grid1 = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(needBlocks, (uint64_t)e->smCount * std::max(occ1, 1)));
    e->grid1s = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(needBlocks, (uint64_t)e->smCount * std::max(occ1s, 1)));
    {
        const char* f = getenv("NC_NEURON_VARIANT");  // tuning/tests: "big", "dense" or "sparse" pins the variant
        e->forceVariant = f ? (f[0] == 's' ? 3 : f[0] == 'b' ? 1 : 2) : 0;
    }
    {   // staging regions: one per tile of 32 rows, large enough for every slot of the largest tile up to 8192 entries
        // (beyond that a tile that is this busy takes the in-kernel staging path); NC_STAGE_CAP overrides (tests)
        uint64_t cap = std::min<uint64_t>(std::max<uint64_t>((32 * maxRow + 31) / 32 * 32, 64), 8192);
        if (const char* f = getenv("NC_STAGE_CAP")) cap = std::max<uint64_t>(strtoull(f, nullptr, 10), 1);
        v.stCap = (uint32_t)cap;
        const uint64_t nTiles = (N1 + 31) / 32;
        CK(cudaMalloc(&v.stCnt, nTiles * 32 * 4));
        CK(cudaMalloc(&v.stAD, nTiles * 2 * cap * 8));  // two regions per tile (ping-pong)
        CK(cudaMalloc(&v.stJ, nTiles * 2 * cap * 4));
        CK(cudaMalloc(&v.tileState, nTiles * 4));
        CK(cudaMemsetAsync(v.tileState, 0, nTiles * 4, e->stream));  // no tile has a list yet: the first staging pass builds them
        e->nTiles = nTiles;
        { const char* f = getenv("NC_STAGE_MODE"); v.stageRebuild = (f && f[0] == 'r') ? 1u : 0u; }
        e->gridStage = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((nTiles + NC_STG_WARPS - 1) / NC_STG_WARPS, (uint64_t)e->smCount * 8));
    }
    // scratch sized for whichever variant needs more: spill beyond the smaller pool, one pool of slot indices per resident warp
    v.spillPerWarp = (uint32_t)(maxRow > e->candCapSparse ? maxRow - e->candCapSparse : 0);
    const uint64_t maxWarps = (uint64_t)std::max(std::max(e->grid1, e->grid1s), e->grid1b) * NC_WARPS_PER_BLOCK;
    uint64_t spillN = std::max<uint64_t>(1, maxWarps * v.spillPerWarp);
    CK(cudaMalloc(&v.spillA, spillN * 4)); CK(cudaMalloc(&v.spillD, spillN * 4)); CK(cudaMalloc(&v.spillJ, spillN * 4));
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(tmpPre); cudaFree(tmpW); cudaFree(tmpLen); cudaFree(tmpInh);
    e->minDelay = minDelay;
    e->uploaded = true;
    return NC_OK;
}

static int upload_precheck(nc_engine* e, uint64_t nGlobal, uint64_t row0, uint64_t nRows, const void* rowptr) {
    if (e->uploaded) return fail(e, NC_ERR_STATE, "nc_upload_network: network already uploaded");
    if (!rowptr) return fail(e, NC_ERR_INVALID, "nc_upload_network: rowptr is null");
    if (nGlobal >= (1ull << 30)) return fail(e, NC_ERR_INVALID, "nc_upload_network: at most 2^30-1 neurons");
    if (row0 + nRows > nGlobal) return fail(e, NC_ERR_INVALID, "nc_upload_network: row range exceeds neuron count");
    return NC_OK;
}

extern "C" int nc_upload_network(nc_engine* e, uint64_t nGlobal, uint64_t row0, uint64_t nRows, const uint64_t* rowptr,
                                 const uint32_t* pre, const float* weight, const float* length, const uint8_t* inh) {
    int rc = upload_precheck(e, nGlobal, row0, nRows, rowptr);
    if (rc) return rc;
    if (rowptr[0] != 0) return fail(e, NC_ERR_INVALID, "nc_upload_network: rowptr[0] must be 0");
    uint64_t S = rowptr[nRows], maxRow = 0;
    if (S && (!pre || !weight || !length || !inh)) return fail(e, NC_ERR_INVALID, "nc_upload_network: null synapse array");
    float minDelay = INFINITY;
    for (uint64_t r = 0; r < nRows; r++) {
        if (rowptr[r + 1] < rowptr[r]) return fail(e, NC_ERR_INVALID, "nc_upload_network: rowptr not monotone");
        uint64_t len = rowptr[r + 1] - rowptr[r];
        maxRow = std::max(maxRow, len);
        if (len >= (1ull << 27)) return fail(e, NC_ERR_INVALID, "nc_upload_network: row longer than 2^27-1");
        for (uint64_t j = rowptr[r]; j < rowptr[r + 1]; j++) {
            if (pre[j] >= nGlobal) return fail(e, NC_ERR_INVALID, "nc_upload_network: presynaptic ID out of range");
            if (j > rowptr[r] && pre[j] <= pre[j - 1]) return fail(e, NC_ERR_INVALID, "nc_upload_network: row not strictly ascending in presynaptic ID");
            float d = length[j] * 2.0f;
            if (!(d > 0.0f)) return fail(e, NC_ERR_INVALID, "nc_upload_network: synapse length must be positive");
            minDelay = std::min(minDelay, d);
        }
    }
    cudaSetDevice(e->cfg.device);
    return upload_common(e, nGlobal, row0, nRows, S, maxRow, minDelay, rowptr, pre, weight, length, inh, cudaMemcpyHostToDevice);
}

// Same as nc_upload_network with all five arrays already resident on this engine's device (e.g. built there by the
// caller); they are copied, so the caller may free them afterwards.
extern "C" int nc_upload_network_device(nc_engine* e, uint64_t nGlobal, uint64_t row0, uint64_t nRows, const uint64_t* dRowptr,
                                        const uint32_t* dPre, const float* dWeight, const float* dLength, const uint8_t* dInh) {
    int rc = upload_precheck(e, nGlobal, row0, nRows, dRowptr);
    if (rc) return rc;
    cudaSetDevice(e->cfg.device);
    uint64_t ends[1], first[1];
    CK(cudaMemcpy(first, dRowptr, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ends, dRowptr + nRows, 8, cudaMemcpyDeviceToHost));
    if (first[0] != 0) return fail(e, NC_ERR_INVALID, "nc_upload_network: rowptr[0] must be 0");
    uint64_t S = ends[0];
    if (S && (!dPre || !dWeight || !dLength || !dInh)) return fail(e, NC_ERR_INVALID, "nc_upload_network: null synapse array");
    unsigned int* dOut; unsigned int hOut[3] = {0u, 0u, 0x7f800000u};
    CK(cudaMalloc(&dOut, 12));
    CK(cudaMemcpy(dOut, hOut, 12, cudaMemcpyHostToDevice));
    k_validate_csr<<<e->smCount * 8, 256>>>(nGlobal, nRows, dRowptr, dPre, dLength, dOut); e->launches++;
    CK(cudaMemcpy(hOut, dOut, 12, cudaMemcpyDeviceToHost));
    cudaFree(dOut);
    if (hOut[0] & 1u) return fail(e, NC_ERR_INVALID, "nc_upload_network: rowptr not monotone");
    if (hOut[0] & 2u) return fail(e, NC_ERR_INVALID, "nc_upload_network: row longer than 2^27-1");
    if (hOut[0] & 4u) return fail(e, NC_ERR_INVALID, "nc_upload_network: presynaptic ID out of range");
    if (hOut[0] & 8u) return fail(e, NC_ERR_INVALID, "nc_upload_network: row not strictly ascending in presynaptic ID");
    if (hOut[0] & 16u) return fail(e, NC_ERR_INVALID, "nc_upload_network: synapse length must be positive");
    float minDelay;
    memcpy(&minDelay, &hOut[2], 4);
    return upload_common(e, nGlobal, row0, nRows, S, hOut[1], minDelay, dRowptr, dPre, dWeight, dLength, dInh, cudaMemcpyDeviceToDevice);
}

extern "C" int nc_min_delay(const nc_engine* e, float* out) {
    if (!e->uploaded) return NC_ERR_STATE;
    *out = e->minDelay;
    return NC_OK;
}

static int check_window(nc_engine* e, float t0, float t1) {
    if (!e->uploaded) return fail(e, NC_ERR_STATE, "step: no network uploaded");
    if (!(t1 > t0)) return fail(e, NC_ERR_INVALID, "step: window must have t1 > t0");
    if (!(t1 - t0 < e->minDelay)) return fail(e, NC_ERR_INVALID, "step: window t1-t0 must be shorter than the smallest synaptic delay (split it)");
    if (!(t1 - t0 < 2.0f)) return fail(e, NC_ERR_INVALID, "step: window must be shorter than the 2 ms spike duration");
    return NC_OK;
}

// Largest positive-float bit pattern b (1 <= b <= hi) for which pred(float(b)) holds, 0 if none; pred must be monotone
// (true for small values, false for large ones).  Evaluated with the same fp32 operations the kernels would use.
template <typename P>
static uint32_t last_true_bits(uint32_t hi, P pred) {
    auto f = [](uint32_t b) { float x; memcpy(&x, &b, 4); return x; };
    if (hi == 0 || !pred(f(1u))) return 0u;
    uint32_t lo = 1u;  // pred(lo) holds
    while (lo < hi) {
        uint32_t mid = lo + (hi - lo + 1u) / 2u;
        if (pred(f(mid))) lo = mid; else hi = mid - 1u;
    }
    return lo;
}
struct WaitCounters {  // head of k_rand_advance on a sharded engine: every shard's counter block of the window has arrived
    ncx::XchgArgs xa;
    __device__ void operator()() const { if (xa.world > 1u) ncx::wait_flags_head(xa, 1u); }
};
static void fill_args(nc_engine* e, StepArgs& a, float t0, float t1, int sweep, const nc_event* dEv, uint32_t nEv) {
    memset(&a, 0, sizeof(a));
    {
        uint32_t t1b; memcpy(&t1b, &t1, 4);
        const volatile float two = 2.0f;
        a.clrB = last_true_bits(t1b, [&](float x) { volatile float d = t1 - x; return d > two; });
        a.reqHiB = last_true_bits(t1b, [&](float x) { volatile float r = x + two; return r <= t1; });
        a.reqLoB = last_true_bits(t1b, [&](float x) { volatile float r = x + two; return r <= t0; });
    }
    a.t0 = t0; a.t1 = t1; a.sweep = sweep;
    a.lr = e->lr; a.preFactor = e->preF; a.postFactor = e->postF; a.preDecay = e->preD; a.postDecay = e->postD;
    a.ev = dEv; a.nEv = nEv; a.world = (uint32_t)e->cfg.world;
    {   // sparse variant while a tile of 32 rows needs a small fraction of its 512-slot pool (measured: 14 % faster in the quiet
        // regime with ~2 occupied slots per row, 4 % slower at 8.5 per row where the replay starts to matter)
        const double perTile = e->lastSlotsPerRun * 32.0;  // occupied slots of a tile of 32 rows, from the last window
        a.variant = e->forceVariant ? (uint32_t)e->forceVariant - 1u : perTile * 4.0 < (double)e->candCapSparse ? 2u : perTile > 0.85 * (double)e->candCap ? 0u : 1u;
        a.candCap = a.variant == 0u ? e->candCapBig : a.variant == 1u ? e->candCap : e->candCapSparse;
    }
    a.gStride = e->cfg.world == 1 ? e->v.fireCap + 1u : e->xchgUnits;
}

// one block per (expected) fire record, capped; the kernels stride over the records they find on the device
static void launch_synapse_pass(nc_engine* e, const StepArgs& a, uint32_t expectFires) {
    const uint64_t cap = (uint64_t)e->smCount * 32u, f = std::max<uint32_t>(expectFires, 1u);
    dim3 gl((uint32_t)std::min<uint64_t>(f * e->v.cprLoads, cap), a.world), gr((uint32_t)std::min<uint64_t>(f * e->v.cprRows, cap), a.world);
    k_syn_loads<<<gl, NC_SYN_THREADS, 0, e->stream>>>(e->v, a);
    k_syn_rows<<<gr, NC_SYN_THREADS, 0, e->stream>>>(e->v, a);
    k_syn_flagged<<<e->smCount * 8, 256, 0, e->stream>>>(e->v, a);
    e->launches += 3;
}
static void launch_neuron_pass(nc_engine* e, const StepArgs& a) {
    k_stage<<<e->gridStage, NC_STG_WARPS * 32, 0, e->stream>>>(e->v, a);
    e->launches++;
    if (e->tick) cudaEventRecord(*e->tick, e->stream);
    if (a.variant == 2u) k_neuron_pass<8><<<e->grid1s, NC_WARPS_PER_BLOCK * 32, e->smem1s, e->stream>>>(e->v, a, e->xa);
    else if (a.variant == 0u) k_neuron_pass<4><<<e->grid1b, NC_WARPS_PER_BLOCK * 32, e->smem1b, e->stream>>>(e->v, a, e->xa);
    else k_neuron_pass<6><<<e->grid1, NC_WARPS_PER_BLOCK * 32, e->smem1, e->stream>>>(e->v, a, e->xa);
    e->launches++;
}
static void mark_events(nc_engine* e, const StepArgs& a, int set) {
    if (!a.nEv && !a.nEvDev) return;
    const unsigned blocks = a.nEvDev ? 32u : (a.nEv + 255u) / 256u;
    k_mark_events<<<blocks, 256, 0, e->stream>>>(e->v, a.ev, a.nEv, a.nEvDev, set);
    e->launches++;
}
static int launch_pass1(nc_engine* e, const StepArgs& a) {
    if (!a.nEvDev) mark_events(e, a, 1);  // (merged lists are marked by k_merge_events; rows un-mark themselves in the neuron pass)
    launch_neuron_pass(e, a);
    if (a.subset) mark_events(e, a, 0);   // (nc_run_neurons skips rows: none of them may keep a mark)
    CK(cudaGetLastError());
    return NC_OK;
}
// the background events of the current run() live on the device: merge those of this window into the host's event list
static int merge_background(nc_engine* e, StepArgs& a, bool strict) {
    const uint32_t need = a.nEv + e->bgCap;
    if (need > e->mergedCap) { cudaFree(e->dEvMerged); e->mergedCap = need * 2 + 1024; CK(cudaMalloc(&e->dEvMerged, (size_t)e->mergedCap * sizeof(nc_event))); }
    k_merge_events<<<1, 1024, 0, e->stream>>>(e->v, a.ev, a.nEv, e->dBgEv, e->dBgCtl, a.t0, a.t1, strict ? 1 : 0, e->dEvMerged, e->mergedCap, e->dMergedCount);
    e->launches++;
    a.ev = e->dEvMerged; a.nEvDev = e->dMergedCount;
    return NC_OK;
}
static int launch_background(nc_engine* e, const ncr::BgArgs& a) {
    const unsigned blocks = (unsigned)((a.nDraws + NC_RS_CHUNK - 1) / NC_RS_CHUNK);
    ncr::k_bg_generate<<<blocks, NC_RS_ROWS, 0, e->stream>>>(e->rtab, e->dRandState, a, e->dDraws, e->dCandRaw, e->candListCap, e->dBgCtl + 4);
    ncr::k_bg_walk<<<1, 1024, 0, e->stream>>>(e->dRandState, a, e->dDraws, e->dCandRaw, e->dBgCtl + 4, e->dCand, e->candListCap, e->dBgTmp, e->dBgEv, e->bgCap, e->dBgCtl, e->dRandNext);
    e->launches += 2;
    std::swap(e->dRandState, e->dRandNext);
    CK(cudaGetLastError());
    return NC_OK;
}
// after a window's counters are final: the hidden rand() calls of its plasticity move the stream ahead (NeuCor.cpp:752)
static int rand_after_window(nc_engine* e, const unsigned long long* counters, bool toHost) {
    if (!e->randOn) return NC_OK;
    if (e->cfg.world > 1) {  // (a single shard's stream was moved on by k_finish_step)
        WaitCounters wc; wc.xa = e->xa;
        ncr::k_rand_advance<<<1, 32, 0, e->stream>>>(e->rtab, e->dRandState, counters, (uint32_t)e->cfg.world, wc);
        e->launches++;
    }
    if (toHost) {
        CK(cudaMemcpyAsync(e->hRand, e->dRandState, 31 * 4, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaMemcpyAsync(e->hRand + 32, e->dBgCtl, 4 * 4, cudaMemcpyDeviceToHost, e->stream));
        e->hRandFresh = true;
    }
    CK(cudaGetLastError());
    return NC_OK;
}
static int rand_after_window(nc_engine* e);  // replay form: gathers the window's counters over the shards first
// Index build over the gathered blocks (counts are read from the block headers on the device), synapse pass, index reset,
// and the end-of-window kernel that publishes the counters + exchange header into dOut and clears them.
static int launch_pass2(nc_engine* e, const StepArgs& a, uint32_t expectMax, int accumulate) {
    const uint32_t gx = std::min<uint32_t>(std::max<uint32_t>((expectMax + 255u) / 256u, 1u), 1024u);
    dim3 g(gx, a.world);
    k_index_build<<<g, 256, 0, e->stream>>>(e->v, a, e->xa);
    launch_synapse_pass(e, a, expectMax);
    k_index_reset<<<g, 256, 0, e->stream>>>(e->v, a);
    k_finish_step<<<1, 32, 0, e->stream>>>(e->v, e->dOut, accumulate, e->dWin, e->xa, e->rtab, (e->randOn && e->cfg.world == 1) ? e->dRandState : nullptr);
    e->launches += 3;
    CK(cudaGetLastError());
    return NC_OK;
}

static int upload_events(nc_engine* e, const nc_event* events, uint32_t nEv, const nc_event** dOut) {
    *dOut = nullptr;
    if (!nEv) return NC_OK;
    for (uint32_t i = 1; i < nEv; i++)
        if (events[i].neuron < events[i - 1].neuron) return fail(e, NC_ERR_INVALID, "step: events must be sorted by neuron");
    if (nEv > e->hEvCap) { cudaFreeHost(e->hEvPinned); e->hEvCap = nEv * 2 + 1024; CK(cudaMallocHost(&e->hEvPinned, (size_t)e->hEvCap * sizeof(nc_event))); }
    memcpy(e->hEvPinned, events, (size_t)nEv * sizeof(nc_event));
    if (e->taping) {
        if (e->tapeUsed + nEv > e->tapeCap) return fail(e, NC_ERR_CAPACITY, "tape: event capacity exceeded");
        nc_event* dst = e->dTape + e->tapeUsed;
        CK(cudaMemcpyAsync(dst, e->hEvPinned, (size_t)nEv * sizeof(nc_event), cudaMemcpyHostToDevice, e->stream));
        *dOut = dst;
        return NC_OK;
    }
    if (nEv > e->evCap) { cudaFree(e->dEv); e->evCap = nEv * 2 + 1024; CK(cudaMalloc(&e->dEv, (size_t)e->evCap * sizeof(nc_event))); }
    CK(cudaMemcpyAsync(e->dEv, e->hEvPinned, (size_t)nEv * sizeof(nc_event), cudaMemcpyHostToDevice, e->stream));
    *dOut = e->dEv;
    return NC_OK;
}

// ---- exchange transport (world > 1) ----
static bool nccl_load(std::string& err) {
    if (g_nccl.lib) return true;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { err = std::string("NCCL not loadable: ") + dlerror(); return false; }
    NcclApi a;
    a.lib = lib;
    *(void**)&a.GetUniqueId = dlsym(lib, "ncclGetUniqueId");
    *(void**)&a.CommInitRank = dlsym(lib, "ncclCommInitRank");
    *(void**)&a.AllGather = dlsym(lib, "ncclAllGather");
    *(void**)&a.CommDestroy = dlsym(lib, "ncclCommDestroy");
    *(void**)&a.GetErrorString = dlsym(lib, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.AllGather || !a.CommDestroy || !a.GetErrorString) { err = "NCCL: missing symbols"; return false; }
    g_nccl = a;
    return true;
}
// Peer exchange set-up (collective over the job's ranks, uses the fresh NCCL communicator as its out-of-band channel):
// allocate this shard's arena, all-gather the CUDA IPC handles, map every peer's arena.  Any rank that cannot do it makes
// all ranks keep the NCCL all-gather path (the decision is itself all-gathered).  NC_NO_P2P=1 skips it.
static int nccl_gather_bytes(nc_engine* e, const void* hostSrc, void* hostDst, size_t bytes) {
    const int W = e->cfg.world;
    char* d = nullptr;
    CK(cudaMalloc(&d, (size_t)(W + 1) * bytes));
    CK(cudaMemcpy(d + (size_t)W * bytes, hostSrc, bytes, cudaMemcpyHostToDevice));
    int rc = g_nccl.AllGather(d + (size_t)W * bytes, d, bytes, /*ncclChar*/ 0, e->comm, e->stream);
    if (rc) { cudaFree(d); return fail(e, NC_ERR_CUDA, std::string("ncclAllGather: ") + g_nccl.GetErrorString(rc)); }
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(hostDst, d, (size_t)W * bytes, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return NC_OK;
}
static int p2p_setup(nc_engine* e) {
    const int W = e->cfg.world;
    if (W < 2 || !e->uploaded) return NC_OK;
    const uint64_t blockUnits = (uint64_t)e->v.fireCap + 1;
    const size_t gBytes = (((size_t)W * blockUnits * sizeof(FireRec)) + 4095) / 4096 * 4096;
    e->offCnt = NC_X_FLAGS_BYTES; e->offG[0] = NC_X_FLAGS_BYTES + NC_X_CNT_BYTES; e->offG[1] = e->offG[0] + gBytes;
    int ok = getenv("NC_NO_P2P") ? 0 : 1;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (ok && cudaMalloc(&e->arena, e->offG[1] + gBytes) != cudaSuccess) { cudaGetLastError(); e->arena = nullptr; ok = 0; }
    if (ok && cudaMemset(e->arena, 0, e->offG[1] + gBytes) != cudaSuccess) ok = 0;
    if (ok && cudaIpcGetMemHandle(&mine, e->arena) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    std::vector<cudaIpcMemHandle_t> all(W);
    int rc = nccl_gather_bytes(e, &mine, all.data(), sizeof(mine));
    if (rc) return rc;
    std::vector<int> oks(W);
    rc = nccl_gather_bytes(e, &ok, oks.data(), sizeof(int));
    if (rc) return rc;
    for (int r = 0; r < W; r++) ok = ok && oks[r];
    if (ok) {
        for (int r = 0; r < W && ok; r++) {
            if (r == e->cfg.rank) { e->peerBase[r] = e->arena; continue; }
            void* p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; }
            e->peerBase[r] = (char*)p;
        }
    }
    rc = nccl_gather_bytes(e, &ok, oks.data(), sizeof(int));  // everybody mapped everybody (also: nobody pushes before all are ready)
    if (rc) return rc;
    for (int r = 0; r < W; r++) ok = ok && oks[r];
    if (!ok) {
        for (int r = 0; r < W; r++) if (r != e->cfg.rank && e->peerBase[r]) cudaIpcCloseMemHandle(e->peerBase[r]);
        cudaFree(e->arena); e->arena = nullptr;
        return NC_OK;  // the NCCL all-gather path stays
    }
    CK(cudaMalloc(&e->dPushCtr, 4));
    CK(cudaMemset(e->dPushCtr, 0, 4));
    e->p2p = true;
    return NC_OK;
}
static ncx::PeerTab peer_tab(const nc_engine* e, uint32_t parity) {
    ncx::PeerTab t;
    memset(&t, 0, sizeof(t));
    for (int r = 0; r < e->cfg.world; r++) {
        t.gather[r] = reinterpret_cast<FireRec*>(e->peerBase[r] + e->offG[parity]);
        t.counters[r] = reinterpret_cast<unsigned long long*>(e->peerBase[r] + e->offCnt);
        t.flags[r] = reinterpret_cast<uint32_t*>(e->peerBase[r]);
    }
    return t;
}
// A sharded window begins: its sequence number, gather buffer (window parity) and the arguments the step's own kernels need
// to push / await the fire records and the counters (peer_exchange.cuh).  Without the peer mapping: no-op (all-gather path).
static void begin_exchange(nc_engine* e, StepArgs& a) {
    memset(&e->xa, 0, sizeof(e->xa));
    if (!e->p2p) return;
    const uint32_t seq = ++e->xseq, parity = seq & 1u;
    e->xa.pt = peer_tab(e, parity);
    e->xa.world = (uint32_t)e->cfg.world; e->xa.rank = (uint32_t)e->cfg.rank; e->xa.blockUnits = e->v.fireCap + 1u; e->xa.seq = seq;
    e->xa.doneCtr = e->dPushCtr; e->xa.errWord = e->v.flagCtl + 1;
    e->v.gRecs = reinterpret_cast<const FireRec*>(e->arena + e->offG[parity]);
    a.gStride = e->xa.blockUnits;
    e->lastStride = e->xa.blockUnits;
}
// a replay's accumulated counters -> every shard (their own area and flag kind), wait for everybody's
static int p2p_exchange_totals(nc_engine* e, const unsigned long long* totals) {
    const uint32_t seq = ++e->xseq;
    const ncx::PeerTab pt = peer_tab(e, seq & 1u);
    ncx::k_push_counters<<<1, 32, 0, e->stream>>>(totals, pt, (uint32_t)e->cfg.world, (uint32_t)e->cfg.rank, seq, 2u);
    ncx::k_wait_flags<<<1, 32, 0, e->stream>>>(reinterpret_cast<const uint32_t*>(e->arena), 2u, (uint32_t)e->cfg.world, seq, e->v.flagCtl + 1);
    e->launches += 2;
    CK(cudaGetLastError());
    return NC_OK;
}
extern "C" int nc_comm_unique_id(nc_comm_id* out) {
    if (!out) { g_err = "nc_comm_unique_id: null argument"; return NC_ERR_INVALID; }
    if (!nccl_load(g_err)) return NC_ERR_NO_DEVICE;
    int rc = g_nccl.GetUniqueId(out);
    if (rc) { g_err = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(rc); return NC_ERR_CUDA; }
    return NC_OK;
}
extern "C" int nc_comm_init(nc_engine* e, const nc_comm_id* id) {
    if (!id) return fail(e, NC_ERR_INVALID, "nc_comm_init: null id");
    if (e->comm) return fail(e, NC_ERR_STATE, "nc_comm_init: communicator already initialised");
    if (!nccl_load(e->err)) return NC_ERR_NO_DEVICE;
    cudaSetDevice(e->cfg.device);
    int rc = g_nccl.CommInitRank(&e->comm, e->cfg.world, *id, e->cfg.rank);
    if (rc) { e->comm = nullptr; return fail(e, NC_ERR_CUDA, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(rc)); }
    return p2p_setup(e);
}
extern "C" int nc_set_exchange(nc_engine* e, nc_allgather_fn fn, void* ctx) {
    e->xchgFn = fn; e->xchgCtx = ctx;
    return NC_OK;
}
// all-gather of `bytes` per shard, stream-ordered with NCCL; with a caller-provided transport the stream is drained first
static int exchange(nc_engine* e, const void* send, void* recv, size_t bytes) {
    if (e->comm) {
        int rc = g_nccl.AllGather(send, recv, bytes, /*ncclChar*/ 0, e->comm, e->stream);
        if (rc) return fail(e, NC_ERR_CUDA, std::string("ncclAllGather: ") + g_nccl.GetErrorString(rc));
        return NC_OK;
    }
    if (e->xchgFn) {
        CK(cudaStreamSynchronize(e->stream));
        if (e->xchgFn(e->xchgCtx, send, recv, (uint64_t)bytes)) return fail(e, NC_ERR_CUDA, "exchange: the caller's all-gather failed");
        return NC_OK;
    }
    return fail(e, NC_ERR_STATE, "exchange: world > 1 needs nc_comm_init or nc_set_exchange");
}

static void sum_out(nc_engine* e, uint64_t* hidden, nc_step_stats* st) {
    unsigned long long t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < e->cfg.world; b++)
        for (int i = 0; i < 8; i++) t[i] += e->hOut[b * 10 + i];
    if (t[6]) e->lastSlotsPerRun = (double)t[7] / (double)t[6];
    if (hidden) *hidden = t[5];
    if (st) {
        st->fires = t[0]; st->deliveries = t[1]; st->loads_accepted = t[2]; st->loads_dropped = t[3];
        st->plasticity_calls = t[4]; st->hidden_rand_calls = t[5]; st->neuron_runs = t[6]; st->active_visits = t[7];
    }
}
// Per-window result blocks (own, or all shards' after an all-gather): enqueue the read-back ...
static int enqueue_counters(nc_engine* e, bool totals = false) {
    const int W = e->cfg.world;
    if (W == 1) {
        CK(cudaMemcpyAsync(e->hOut, e->dOut, 10 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    } else if (e->p2p && totals) {  // end of a replay: the accumulated counters of every shard
        int rc = p2p_exchange_totals(e, e->dOut);
        if (rc) return rc;
        CK(cudaMemcpyAsync(e->hOut, e->arena + e->offCnt + NC_X_CNT_AREA, (size_t)W * 10 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    } else if (e->p2p) {  // a live window: k_finish_step pushed the counters; they are awaited by k_rand_advance or here
        if (!e->randOn) {
            ncx::k_wait_flags<<<1, 32, 0, e->stream>>>(reinterpret_cast<const uint32_t*>(e->arena), 1u, (uint32_t)W, e->xa.seq, e->v.flagCtl + 1);
            e->launches++;
        }
        CK(cudaMemcpyAsync(e->hOut, e->arena + e->offCnt, (size_t)W * 10 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    } else {
        int rc = exchange(e, e->dOut, e->dOutAll, 10 * sizeof(unsigned long long));
        if (rc) return rc;
        CK(cudaMemcpyAsync(e->hOut, e->dOutAll, (size_t)W * 10 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    }
    return NC_OK;
}
// ... and wait for it: reports the network-wide counters.
static int wait_counters(nc_engine* e, uint64_t* hidden, nc_step_stats* st) {
    const int W = e->cfg.world;
    CK(cudaStreamSynchronize(e->stream));
    if (e->randOn && e->bgActive && e->hRandFresh && e->hRand[32 + 2])
        return fail(e, NC_ERR_CAPACITY, "background firing: more hits than the device generator made room for");
    sum_out(e, hidden, st);
    for (int b = 0; b < W; b++) {
        if (e->hOut[b * 10 + 9] & 8ull) return fail(e, NC_ERR_CUDA, "step: the peer exchange timed out (a shard of the job stopped stepping)");
        if (e->hOut[b * 10 + 9]) return fail(e, NC_ERR_CAPACITY, (e->hOut[b * 10 + 9] & 1ull) ? "step: fire-record capacity exceeded (raise nc_config.fire_capacity)"
                                                                                              : "step: flag-list capacity exceeded (raise nc_config.flag_capacity)");
    }
    return NC_OK;
}
static int finish_counters(nc_engine* e, uint64_t* hidden, nc_step_stats* st) {
    int rc = enqueue_counters(e, true);
    if (rc) return rc;
    return wait_counters(e, hidden, st);
}

// The fire exchange of a live window (world > 1): all-gather of the first xchgUnits units of every shard's block, then a
// look at the gathered headers; when a shard fired more than that the size is raised and the all-gather repeated.
static int ensure_gather(nc_engine* e) {  // gather buffer of the all-gather path (NCCL without peer mapping, caller-provided transport)
    if (e->dGather) return NC_OK;
    CK(cudaMalloc(&e->dGather, (uint64_t)e->cfg.world * ((uint64_t)e->v.fireCap + 1) * sizeof(FireRec)));
    e->v.gRecs = e->dGather;
    return NC_OK;
}
static int exchange_fires(nc_engine* e, StepArgs& a, uint32_t* maxCount) {
    const int W = e->cfg.world;
    { int rc0 = ensure_gather(e); if (rc0) return rc0; }
    for (;;) {
        int rc = exchange(e, e->v.localHdr, e->dGather, (size_t)e->xchgUnits * sizeof(FireRec));
        if (rc) return rc;
        CK(cudaMemcpy2DAsync(e->hHdrAll, 16, e->dGather, (size_t)e->xchgUnits * sizeof(FireRec), 16, W, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        uint32_t mx = 0;
        for (int b = 0; b < W; b++) {
            if (e->hHdrAll[4 * b + 1]) return fail(e, NC_ERR_CAPACITY, "step: fire-record capacity exceeded (raise nc_config.fire_capacity)");
            mx = std::max(mx, e->hHdrAll[4 * b]);
        }
        if (mx + 1u <= e->xchgUnits) {
            for (int b = 0; b < W; b++) e->lastCounts[b] = e->hHdrAll[4 * b];
            e->lastStride = e->xchgUnits;
            a.gStride = e->xchgUnits;
            *maxCount = mx;
            return NC_OK;
        }
        uint64_t want = 1;
        while (want < 2ull * mx) want <<= 1;
        e->xchgUnits = (uint32_t)std::min<uint64_t>(want + 1, (uint64_t)e->v.fireCap + 1);
    }
}

// exchange (world > 1), index build, synapse pass, end-of-window kernel, counter read-back enqueued
static int step_second_half(nc_engine* e) {
    StepArgs& a = e->pendingArgs;
    uint32_t expect = std::max<uint32_t>(e->lastCounts[0], 256u);
    if (e->cfg.world > 1) {
        if (e->p2p) expect = (uint32_t)std::min<uint64_t>((uint64_t)expect * e->cfg.world, 1u << 24);  // (own count of the last window x shards: a launch-size hint)
        else { int rc = exchange_fires(e, a, &expect); if (rc) return rc; }
    }
    if (e->taping) {
        int32_t randIdx = -1;
        if (e->tapeRandPending) { randIdx = (int32_t)(e->tapeRand.size() / 32); e->tapeRand.insert(e->tapeRand.end(), e->tapeRandPend, e->tapeRandPend + 32); e->tapeRandPending = false; }
        TapeStep ts = {a.t0, a.t1, a.sweep, e->tapeUsed, a.nEv, a.gStride, expect, e->bgPendingTape, e->pendingBgActive, e->pendingBgStrict, e->bgPendingArgs, randIdx};
        e->tape.push_back(ts);
        e->tapeUsed += a.nEv;
        e->bgPendingTape = false;
    }
    int rc = launch_pass2(e, a, expect, 0);
    if (rc) return rc;
    if (e->p2p) {  // (the stream advance awaits the shards' counters; the read-back follows it)
        rc = rand_after_window(e, reinterpret_cast<const unsigned long long*>(e->arena + e->offCnt), true);
        if (rc) return rc;
        return enqueue_counters(e);
    }
    rc = enqueue_counters(e);
    if (rc) return rc;
    return rand_after_window(e, e->cfg.world == 1 ? e->dOut : e->dOutAll, true);
}
static int rand_after_window(nc_engine* e) {
    if (!e->randOn) return NC_OK;
    if (e->cfg.world > 1 && !e->p2p) {
        int rc = exchange(e, e->dWin, e->dWinAll, 10 * sizeof(unsigned long long));
        if (rc) return rc;
    }
    const unsigned long long* all = e->cfg.world == 1 ? e->dWin : e->p2p ? reinterpret_cast<const unsigned long long*>(e->arena + e->offCnt) : e->dWinAll;
    return rand_after_window(e, all, false);
}
extern "C" int nc_step_launch(nc_engine* e, float t0, float t1, int sweep, const nc_event* events, uint32_t nEv) {
    if (e->pending) return fail(e, NC_ERR_STATE, "nc_step_launch: the previous window has not been collected");
    int rc = check_window(e, t0, t1);
    if (rc) return rc;
    cudaSetDevice(e->cfg.device);
    const nc_event* dEv = nullptr;
    rc = upload_events(e, events, nEv, &dEv);
    if (rc) return rc;
    if (e->taping && e->tape.size() >= e->tapeMaxSteps) return fail(e, NC_ERR_CAPACITY, "tape: step capacity exceeded");
    StepArgs a;
    fill_args(e, a, t0, t1, sweep, dEv, nEv);
    if (e->cfg.world > 1) begin_exchange(e, a);
    e->pendingBgActive = e->bgActive; e->pendingBgStrict = !e->bgFirstWindow;
    if (e->bgActive) {
        rc = merge_background(e, a, !e->bgFirstWindow);
        if (rc) return rc;
        e->bgFirstWindow = false;
    }
    rc = launch_pass1(e, a);
    if (rc) return rc;
    e->pendingArgs = a;
    e->pending = true;
    // a single shard needs nothing from the host between the passes: enqueue the rest of the window right away
    if (e->cfg.world == 1) {
        rc = step_second_half(e);
        if (rc) e->pending = false;
        return rc;
    }
    return NC_OK;
}
extern "C" int nc_step_collect(nc_engine* e, uint64_t* hidden, nc_step_stats* st) {
    if (!e->pending) return fail(e, NC_ERR_STATE, "nc_step_collect: no window in flight");
    cudaSetDevice(e->cfg.device);
    e->pending = false;
    if (e->cfg.world > 1) {  // the fire exchange looks at the gathered headers on the host (blocks until the neuron pass is done)
        int rc = step_second_half(e);
        if (rc) return rc;
    }
    int rc = wait_counters(e, hidden, st);
    if (e->cfg.world == 1) { e->lastCounts[0] = (uint32_t)std::min<unsigned long long>(e->hOut[8], e->v.fireCap); e->lastStride = e->v.fireCap + 1u; }
    else if (e->p2p) for (int b = 0; b < e->cfg.world; b++) e->lastCounts[b] = (uint32_t)std::min<unsigned long long>(e->hOut[b * 10 + 8], e->v.fireCap);
    return rc;
}
extern "C" int nc_step(nc_engine* e, float t0, float t1, int sweep, const nc_event* events, uint32_t nEv, uint64_t* hidden,
                       nc_step_stats* st) {
    int rc = nc_step_launch(e, t0, t1, sweep, events, nEv);
    if (rc) return rc;
    return nc_step_collect(e, hidden, st);
}

static int scratch(nc_engine* e, size_t bytes, void** out);
extern "C" int nc_run_neurons(nc_engine* e, float now, const uint32_t* ids, uint32_t nIds, uint64_t* hidden, nc_step_stats* st) {
    if (e->cfg.world != 1) return fail(e, NC_ERR_STATE, "nc_run_neurons: single-shard engines only");
    if (!e->uploaded) return fail(e, NC_ERR_STATE, "nc_run_neurons: no network uploaded");
    cudaSetDevice(e->cfg.device);
    uint32_t* dIds = nullptr;
    if (ids) {
        if (!nIds) return NC_OK;
        for (uint32_t i = 0; i < nIds; i++) {
            if (ids[i] < e->v.row0 || ids[i] >= e->v.row0 + e->v.nRows) return fail(e, NC_ERR_INVALID, "nc_run_neurons: neuron outside this shard");
            if (i && ids[i] <= ids[i - 1]) return fail(e, NC_ERR_INVALID, "nc_run_neurons: ids must be strictly ascending");
        }
        void* d = nullptr;  // (the engine's persistent read-back scratch: no allocation per call, nothing to leak on an error path)
        int rcs = scratch(e, (size_t)nIds * 4, &d);
        if (rcs) return rcs;
        dIds = (uint32_t*)d;
        CK(cudaMemcpyAsync(dIds, ids, (size_t)nIds * 4, cudaMemcpyHostToDevice, e->stream));
    }
    StepArgs a;
    fill_args(e, a, now, now, NC_SWEEP_END, nullptr, 0);
    a.subset = dIds; a.nSubset = nIds;
    int rc = launch_pass1(e, a);
    if (!rc) rc = launch_pass2(e, a, std::max<uint32_t>(e->lastCounts[0], 256u), 0);
    if (!rc) rc = finish_counters(e, hidden, st);
    if (!rc) { e->lastCounts[0] = (uint32_t)std::min<unsigned long long>(e->hOut[8], e->v.fireCap); e->lastStride = e->v.fireCap + 1u; }
    return rc;
}

extern "C" int nc_read_neurons(nc_engine* e, float* potAct, float* lastFire, float* lastRan) {
    if (!e->uploaded) return fail(e, NC_ERR_STATE, "read: no network");
    cudaSetDevice(e->cfg.device);
    CK(cudaStreamSynchronize(e->stream));
    if (potAct) CK(cudaMemcpy(potAct, e->v.potAct, e->v.nRows * 8, cudaMemcpyDeviceToHost));
    if (lastFire) CK(cudaMemcpy(lastFire, e->v.lastFire, e->v.nRows * 4, cudaMemcpyDeviceToHost));
    if (lastRan) CK(cudaMemcpy(lastRan, e->v.lastRan, e->v.nRows * 4, cudaMemcpyDeviceToHost));
    return NC_OK;
}
extern "C" int nc_read_synapses(nc_engine* e, float* weight, float* arrive, float* depol, float* lastArr, float* lastStart) {
    if (!e->uploaded) return fail(e, NC_ERR_STATE, "read: no network");
    cudaSetDevice(e->cfg.device);
    CK(cudaStreamSynchronize(e->stream));
    uint64_t b = e->v.S * 4;
    if ((arrive || depol || weight || lastArr) && e->v.S) {  // these live in interleaved records on the device: de-interleave through a temporary
        void* d = nullptr;
        int rcs = scratch(e, b, &d);
        if (rcs) return rcs;
        float* tmp = (float*)d;
        float* dsts[4] = {arrive, depol, weight, lastArr};
        for (int which = 0; which < 4; which++) {
            float* dst = dsts[which];
            if (!dst) continue;
            k_extract_ad<<<(unsigned)std::min<uint64_t>((e->v.S + 255) / 256, 1u << 20), 256, 0, e->stream>>>(e->v, which, tmp); e->launches++;
            CK(cudaMemcpyAsync(dst, tmp, b, cudaMemcpyDeviceToHost, e->stream));
            CK(cudaStreamSynchronize(e->stream));
        }
    }
    if (lastStart) CK(cudaMemcpy(lastStart, e->v.lastStart, b, cudaMemcpyDeviceToHost));
    return NC_OK;
}
static int scratch(nc_engine* e, size_t bytes, void** out);
extern "C" int nc_read_neuron_counters(nc_engine* e, float* actStart, uint32_t* firings) {
    if (!e->uploaded) return fail(e, NC_ERR_STATE, "read: no network");
    cudaSetDevice(e->cfg.device);
    CK(cudaStreamSynchronize(e->stream));
    if (actStart) CK(cudaMemcpy(actStart, e->v.actStart, e->v.nRows * 4, cudaMemcpyDeviceToHost));
    if (firings) CK(cudaMemcpy(firings, e->v.firings, e->v.nRows * 4, cudaMemcpyDeviceToHost));
    return NC_OK;
}
extern "C" int nc_read_network(nc_engine* e, uint64_t* rowptr, uint32_t* pre, float* length, uint8_t* inhibitory) {
    if (!e->uploaded) return fail(e, NC_ERR_STATE, "read: no network");
    cudaSetDevice(e->cfg.device);
    CK(cudaStreamSynchronize(e->stream));
    if (rowptr) CK(cudaMemcpy(rowptr, e->v.rowptr, (e->v.nRows + 1) * 8, cudaMemcpyDeviceToHost));
    const uint64_t S = e->v.S;
    if (!S) return NC_OK;
    void* d = nullptr;
    int rc = scratch(e, S * 4, &d);
    if (rc) return rc;
    const unsigned blocks = (unsigned)std::min<uint64_t>((S + 255) / 256, 1u << 20);
    std::vector<uint32_t> tmp;
    for (int which = 0; which < 3; which++) {
        void* dst = which == 0 ? (void*)pre : which == 1 ? (void*)length : (void*)inhibitory;
        if (!dst) continue;
        k_extract_net<<<blocks, 256, 0, e->stream>>>(e->v, which, (uint32_t*)d); e->launches++;
        if (which < 2) { CK(cudaMemcpyAsync(dst, d, S * 4, cudaMemcpyDeviceToHost, e->stream)); CK(cudaStreamSynchronize(e->stream)); }
        else {
            tmp.resize(S);
            CK(cudaMemcpyAsync(tmp.data(), d, S * 4, cudaMemcpyDeviceToHost, e->stream)); CK(cudaStreamSynchronize(e->stream));
            for (uint64_t j = 0; j < S; j++) inhibitory[j] = (uint8_t)tmp[j];
        }
    }
    return NC_OK;
}
// Checkpoint resume: the inverse of nc_read_neurons / nc_read_neuron_counters / nc_read_synapses (NULL = keep what is there).
extern "C" int nc_write_neurons(nc_engine* e, const float* potAct, const float* lastFire, const float* lastRan, const float* actStart, const uint32_t* firings) {
    if (!e->uploaded) return fail(e, NC_ERR_STATE, "write: no network");
    if (e->pending) return fail(e, NC_ERR_STATE, "write: a window is in flight");
    cudaSetDevice(e->cfg.device);
    CK(cudaStreamSynchronize(e->stream));
    const uint64_t n = e->v.nRows;
    if (potAct) CK(cudaMemcpy(e->v.potAct, potAct, n * 8, cudaMemcpyHostToDevice));
    if (lastFire) CK(cudaMemcpy(e->v.lastFire, lastFire, n * 4, cudaMemcpyHostToDevice));
    if (lastRan) CK(cudaMemcpy(e->v.lastRan, lastRan, n * 4, cudaMemcpyHostToDevice));
    if (actStart) CK(cudaMemcpy(e->v.actStart, actStart, n * 4, cudaMemcpyHostToDevice));
    if (firings) CK(cudaMemcpy(e->v.firings, firings, n * 4, cudaMemcpyHostToDevice));
    return NC_OK;
}
extern "C" int nc_write_synapses(nc_engine* e, const float* weight, const float* arrive, const float* depol, const float* lastArr, const float* lastStart) {
    if (!e->uploaded) return fail(e, NC_ERR_STATE, "write: no network");
    if (e->pending) return fail(e, NC_ERR_STATE, "write: a window is in flight");
    cudaSetDevice(e->cfg.device);
    const uint64_t S = e->v.S;
    if (!S) return NC_OK;
    void* d = nullptr;
    int rc = scratch(e, S * 4, &d);
    if (rc) return rc;
    const unsigned blocks = (unsigned)std::min<uint64_t>((S + 255) / 256, 1u << 20);
    const float* srcs[4] = {arrive, depol, weight, lastArr};
    for (int which = 0; which < 4; which++) {
        if (!srcs[which]) continue;
        CK(cudaMemcpyAsync(d, srcs[which], S * 4, cudaMemcpyHostToDevice, e->stream));
        k_inject_syn<<<blocks, 256, 0, e->stream>>>(e->v, which, (const float*)d); e->launches++;
        CK(cudaStreamSynchronize(e->stream));
    }
    if (lastStart) CK(cudaMemcpy(e->v.lastStart, lastStart, S * 4, cudaMemcpyHostToDevice));
    if (arrive) {  // the event index follows `arrive`: every busy slot is looked at once by the next staging pass, which sorts out the arrived ones
        k_rebuild_busy<<<(unsigned)std::min<uint64_t>((e->busyWords + 255) / 256, 1u << 20), 256, 0, e->stream>>>(e->v, e->busyWords); e->launches++;
        CK(cudaMemsetAsync(e->v.tileState, 0, e->nTiles * 4, e->stream));
        CK(cudaStreamSynchronize(e->stream));
    }
    return NC_OK;
}
extern "C" int nc_read_fires(nc_engine* e, uint32_t capacity, uint32_t* neuron, float* time, uint32_t* count) {
    if (!e->uploaded) return fail(e, NC_ERR_STATE, "read: no network");
    cudaSetDevice(e->cfg.device);
    CK(cudaStreamSynchronize(e->stream));
    uint32_t total = 0;
    std::vector<FireRec> tmp;
    for (uint32_t b = 0; b < (uint32_t)e->cfg.world; b++) {
        uint32_t c = e->lastCounts[b];
        if (!c) continue;
        if (!capacity) { total += c; continue; }  // count only
        tmp.resize(c);
        CK(cudaMemcpy(tmp.data(), e->v.gRecs + (uint64_t)b * e->lastStride + 1, (uint64_t)c * sizeof(FireRec), cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < c; i++, total++)
            if (total < capacity) { if (neuron) neuron[total] = tmp[i].neuron; if (time) time[total] = tmp[i].time; }
    }
    *count = total;
    return NC_OK;
}
// persistent device scratch of the read-back entry points (grown on demand, freed with the engine)
static int scratch(nc_engine* e, size_t bytes, void** out) {
    if (bytes > e->scratchBytes) {
        cudaFree(e->dScratch); e->dScratch = nullptr; e->scratchBytes = 0;
        CK(cudaMalloc(&e->dScratch, bytes + bytes / 4 + 256));
        e->scratchBytes = bytes + bytes / 4 + 256;
    }
    *out = e->dScratch;
    return NC_OK;
}
// Per-frame synapse potentials written straight into caller-provided DEVICE (or host-mapped / graphics-interop) buffers:
// what NeuCor_Renderer::updateView gathers per synapse every frame (Renderer.cpp:655-699), without a staging copy.
extern "C" int nc_synapse_pots_device(nc_engine* e, float now, float* dPrePot, float* dPostPot) {
    if (!e->uploaded) return fail(e, NC_ERR_STATE, "read: no network");
    cudaSetDevice(e->cfg.device);
    if (!e->v.S) return NC_OK;
    k_synapse_pots<<<(unsigned)((e->v.S + 255) / 256), 256, 0, e->stream>>>(e->v, now, dPrePot, dPostPot); e->launches++;
    CK(cudaGetLastError());
    return NC_OK;
}
extern "C" int nc_read_synapse_pots(nc_engine* e, float now, float* prePot, float* postPot) {
    if (!e->uploaded) return fail(e, NC_ERR_STATE, "read: no network");
    cudaSetDevice(e->cfg.device);
    const uint64_t S = e->v.S;
    if (!S) return NC_OK;
    void* d = nullptr;
    int rc = scratch(e, S * 8, &d);
    if (rc) return rc;
    float *dPre = (float*)d, *dPost = dPre + S;
    rc = nc_synapse_pots_device(e, now, dPre, dPost);
    if (rc) return rc;
    if (prePot) CK(cudaMemcpyAsync(prePot, dPre, S * 4, cudaMemcpyDeviceToHost, e->stream));
    if (postPot) CK(cudaMemcpyAsync(postPot, dPost, S * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return NC_OK;
}
static int render_histogram(nc_engine* e, int which, uint32_t spans, float rmin, float rmax, uint32_t* bins, uint32_t* below, uint32_t* above) {
    if (!e->uploaded) return fail(e, NC_ERR_STATE, "render: no network");
    if (spans == 0 || spans > (1u << 20) || !bins) return fail(e, NC_ERR_INVALID, "render: bad histogram arguments");
    cudaSetDevice(e->cfg.device);
    for (uint32_t i = 0; i < spans; i++) bins[i] = 0;
    if (below) *below = 0;
    if (above) *above = 0;
    const float range = rmax - rmin;
    if (!(0.0f < range)) return NC_OK;  // the reference leaves the distribution empty (Renderer.cpp:1748,1799)
    void* d = nullptr;
    int rc = scratch(e, (size_t)(spans + 2) * 4, &d);
    if (rc) return rc;
    CK(cudaMemsetAsync(d, 0, (size_t)(spans + 2) * 4, e->stream));
    const uint64_t n = which ? e->v.S : e->v.nRows;
    const unsigned blocks = (unsigned)std::min<uint64_t>(std::max<uint64_t>((n + 255) / 256, 1), (uint64_t)e->smCount * 16);
    k_render_histogram<<<blocks, 256, 0, e->stream>>>(e->v, which, spans, rmin, range, (unsigned int*)d); e->launches++;
    std::vector<uint32_t> h(spans + 2);
    CK(cudaMemcpyAsync(h.data(), d, (size_t)(spans + 2) * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    for (uint32_t i = 0; i < spans; i++) bins[i] = h[i];
    if (below) *below = h[spans];
    if (above) *above = h[spans + 1];
    return NC_OK;
}
extern "C" int nc_render_activity_histogram(nc_engine* e, uint32_t spans, float rmin, float rmax, uint32_t* bins, uint32_t* below, uint32_t* above) {
    return render_histogram(e, 0, spans, rmin, rmax, bins, below, above);
}
extern "C" int nc_render_weight_histogram(nc_engine* e, uint32_t spans, float rmin, float rmax, uint32_t* bins, uint32_t* below, uint32_t* above) {
    return render_histogram(e, 1, spans, rmin, rmax, bins, below, above);
}
extern "C" int nc_render_raster(nc_engine* e, float now, float runSpeed, uint32_t capacity, uint32_t* ids, uint32_t* count) {
    if (!e->uploaded) return fail(e, NC_ERR_STATE, "render: no network");
    if (!count) return fail(e, NC_ERR_INVALID, "render: null count");
    cudaSetDevice(e->cfg.device);
    void* d = nullptr;
    int rc = scratch(e, (size_t)capacity * 4 + 16, &d);
    if (rc) return rc;
    unsigned int* dCount = (unsigned int*)d;
    uint32_t* dIds = (uint32_t*)d + 4;
    CK(cudaMemsetAsync(dCount, 0, 4, e->stream));
    const unsigned blocks = (unsigned)std::min<uint64_t>(std::max<uint64_t>((e->v.nRows + 255) / 256, 1), (uint64_t)e->smCount * 16);
    k_render_raster<<<blocks, 256, 0, e->stream>>>(e->v, now, runSpeed, capacity, dIds, dCount); e->launches++;
    uint32_t c = 0;
    CK(cudaMemcpyAsync(&c, dCount, 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    *count = c;
    const uint32_t k = std::min(c, capacity);
    if (ids && k) {
        CK(cudaMemcpy(ids, dIds, (size_t)k * 4, cudaMemcpyDeviceToHost));
        std::sort(ids, ids + k);  // ascending ID, the order the GUI's loop over brain->neurons produces
    }
    return NC_OK;
}
extern "C" int nc_state_signature(nc_engine* e, uint64_t* out6) {
    if (!e->uploaded) return fail(e, NC_ERR_STATE, "signature: no network");
    if (!out6) return fail(e, NC_ERR_INVALID, "signature: null output");
    cudaSetDevice(e->cfg.device);
    CK(cudaMemsetAsync(e->dSig, 0, 6 * sizeof(unsigned long long), e->stream));
    const uint64_t work = std::max<uint64_t>(e->v.S, e->v.nRows);
    const unsigned blocks = (unsigned)std::min<uint64_t>(std::max<uint64_t>((work + 255) / 256, 1), (uint64_t)e->smCount * 16);
    k_state_signature<<<blocks, 256, 0, e->stream>>>(e->v, e->dSig); e->launches++;
    CK(cudaMemcpyAsync(e->hSig, e->dSig, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    for (int i = 0; i < 6; i++) out6[i] = e->hSig[i];
    return NC_OK;
}
extern "C" int nc_reset_activities(nc_engine* e, float now) {
    if (!e->uploaded) return fail(e, NC_ERR_STATE, "reset: no network");
    cudaSetDevice(e->cfg.device);
    if (e->v.nRows) { k_reset_activities<<<(unsigned)((e->v.nRows + 255) / 256), 256, 0, e->stream>>>(e->v, now); e->launches++; }
    CK(cudaStreamSynchronize(e->stream));
    return NC_OK;
}
extern "C" int nc_detector_mean(nc_engine* e, const uint32_t* near, uint32_t n, float* out) {
    if (!e->uploaded) return fail(e, NC_ERR_STATE, "detector: no network");
    cudaSetDevice(e->cfg.device);
    if (!n) { *out = NAN; return NC_OK; }  // 0/0 as in the reference
    for (uint32_t i = 0; i < n; i++)
        if (near[i] < e->v.row0 || near[i] >= e->v.row0 + e->v.nRows) return fail(e, NC_ERR_INVALID, "detector: neuron outside this shard");
    void* d = nullptr;
    int rc = scratch(e, (size_t)n * 4 + 16, &d);
    if (rc) return rc;
    float* dOut = (float*)d; uint32_t* dNear = (uint32_t*)d + 4;
    CK(cudaMemcpyAsync(dNear, near, (size_t)n * 4, cudaMemcpyHostToDevice, e->stream));
    k_detector_mean<<<1, 32, 0, e->stream>>>(e->v, dNear, n, dOut); e->launches++;
    CK(cudaMemcpyAsync(out, dOut, 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return NC_OK;
}

// ---- tape / snapshot / replay ----
extern "C" int nc_tape_begin(nc_engine* e, uint32_t maxSteps, uint64_t maxEvents) {
    if (!e->uploaded) return fail(e, NC_ERR_STATE, "tape: no network");
    cudaSetDevice(e->cfg.device);
    cudaFree(e->dTape); e->dTape = nullptr;
    e->tapeCap = std::max<uint64_t>(maxEvents, 1);
    CK(cudaMalloc(&e->dTape, e->tapeCap * sizeof(nc_event)));
    e->tape.clear(); e->tapeUsed = 0; e->tapeMaxSteps = maxSteps; e->taping = true;
    e->tapeRand.clear(); e->tapeRandPending = false; cudaFree(e->dTapeRand); e->dTapeRand = nullptr;
    return NC_OK;
}
extern "C" int nc_tape_end(nc_engine* e) {
    e->taping = false;
    if (!e->tapeRand.empty()) {
        cudaSetDevice(e->cfg.device);
        CK(cudaMalloc(&e->dTapeRand, e->tapeRand.size() * 4));
        CK(cudaMemcpy(e->dTapeRand, e->tapeRand.data(), e->tapeRand.size() * 4, cudaMemcpyHostToDevice));
    }
    return NC_OK;
}

template <typename T>
static cudaError_t snap_copy(T*& dst, const T* src, uint64_t n, cudaStream_t st, bool toSnap) {
    cudaError_t ce = cudaSuccess;
    if (toSnap && !dst) ce = cudaMalloc(&dst, std::max<uint64_t>(n, 1) * sizeof(T));
    if (ce != cudaSuccess) return ce;
    if (toSnap) return cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToDevice, st);
    return cudaMemcpyAsync((void*)src, dst, n * sizeof(T), cudaMemcpyDeviceToDevice, st);
}
static int snap_all(nc_engine* e, bool toSnap) {
    View& v = e->v; auto& s = e->snap;
    CK(snap_copy(s.ad, v.ad, v.S, e->stream, toSnap));
    CK(snap_copy(s.rec, v.rec, v.S, e->stream, toSnap));
    CK(snap_copy(s.lastStart, v.lastStart, v.S, e->stream, toSnap)); CK(snap_copy(s.lastRan, v.lastRan, v.nRows, e->stream, toSnap));
    CK(snap_copy(s.lastFire, v.lastFire, v.nRows, e->stream, toSnap)); CK(snap_copy(s.lfStart, v.lfStart, v.nRows, e->stream, toSnap));
    CK(snap_copy(s.actStart, v.actStart, v.nRows, e->stream, toSnap)); CK(snap_copy(s.potAct, v.potAct, v.nRows, e->stream, toSnap));
    CK(snap_copy(s.firings, v.firings, v.nRows, e->stream, toSnap));
    if (toSnap) { CK(cudaMemcpyAsync(e->snapRand, e->dRandState, 31 * 4, cudaMemcpyDeviceToDevice, e->stream)); e->snapRandOn = e->randOn; }
    else {
        CK(cudaMemcpyAsync(e->dRandState, e->snapRand, 31 * 4, cudaMemcpyDeviceToDevice, e->stream)); e->randOn = e->snapRandOn; e->hRandFresh = false;
        CK(cudaMemsetAsync(v.tileState, 0, e->nTiles * 4, e->stream));  // the staged lists belong to the state that was replaced: rebuilt by the next window
    }
    CK(snap_copy(s.busy, v.busy, e->busyWords, e->stream, toSnap));
    CK(snap_copy(s.arrived, v.arrived, e->busyWords, e->stream, toSnap));
    CK(snap_copy(s.wordNext, v.wordNext, e->busyWords, e->stream, toSnap));
    CK(cudaStreamSynchronize(e->stream));
    return NC_OK;
}
extern "C" int nc_snapshot(nc_engine* e) {
    if (!e->uploaded) return fail(e, NC_ERR_STATE, "snapshot: no network");
    cudaSetDevice(e->cfg.device);
    int rc = snap_all(e, true);
    if (!rc) e->snap.valid = true;
    return rc;
}
extern "C" int nc_restore(nc_engine* e) {
    if (!e->snap.valid) return fail(e, NC_ERR_STATE, "restore: no snapshot");
    cudaSetDevice(e->cfg.device);
    return snap_all(e, false);
}

// Replays taped windows back to back with no host<->device traffic and no host synchronisation inside the timed region:
// the index kernels read the fire counts from the block headers on the device, and for world > 1 the fire exchange is an
// in-stream NCCL all-gather of the size the live run settled on for that window.
extern "C" int nc_tape_replay(nc_engine* e, uint32_t first, uint32_t count, float* msTotal, float* msP1, float* msP2, float* msXchg,
                              uint64_t* hidden, nc_step_stats* st) {
    if ((uint64_t)first + count > e->tape.size()) return fail(e, NC_ERR_INVALID, "replay: step range outside the tape");
    if (e->cfg.world > 1 && !e->comm) return fail(e, NC_ERR_STATE, "replay: world > 1 needs the NCCL communicator (nc_comm_init)");
    cudaSetDevice(e->cfg.device);
    CK(cudaMemsetAsync(e->v.stats, 0, 8 * sizeof(unsigned long long), e->stream));
    CK(cudaMemsetAsync(e->dOut, 0, 10 * sizeof(unsigned long long), e->stream));
    const bool perKernel = msP1 || msP2 || msXchg;
    std::vector<cudaEvent_t> evs;
    std::vector<cudaEvent_t> evStage;
    if (perKernel) { evs.resize((size_t)count * 5); for (auto& x : evs) CK(cudaEventCreate(&x)); evStage.resize(count); for (auto& x : evStage) CK(cudaEventCreate(&x)); }
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const uint32_t gx = (uint32_t)std::min<uint64_t>(std::max<uint64_t>((e->v.nRows / 64 + 255) / 256, 1), 1024);
    CK(cudaEventRecord(e0, e->stream));
    for (uint32_t k = 0; k < count; k++) {
        const TapeStep& ts = e->tape[first + k];
        StepArgs a;
        fill_args(e, a, ts.t0, ts.t1, ts.sweep, e->dTape + ts.evOff, ts.nEv);
        a.gStride = ts.units;
        if (e->cfg.world > 1) begin_exchange(e, a);
        if (ts.randIdx >= 0) CK(cudaMemcpyAsync(e->dRandState, e->dTapeRand + (size_t)ts.randIdx * 32, 31 * 4, cudaMemcpyDeviceToDevice, e->stream));
        if (ts.bgDraw) { int rc = launch_background(e, ts.bg); if (rc) return rc; }
        if (ts.bgActive) { int rc = merge_background(e, a, ts.bgStrict); if (rc) return rc; }
        if (!a.nEvDev) mark_events(e, a, 1);
        if (perKernel) { CK(cudaEventRecord(evs[5 * k], e->stream)); e->tick = &evStage[k]; }
        launch_neuron_pass(e, a);
        e->tick = nullptr;
        if (perKernel) CK(cudaEventRecord(evs[5 * k + 1], e->stream));
        if (e->cfg.world > 1 && !e->p2p) {
            int rc = ensure_gather(e);
            if (!rc) rc = exchange(e, e->v.localHdr, e->dGather, (size_t)ts.units * sizeof(FireRec));
            if (rc) return rc;
        }
        if (perKernel) CK(cudaEventRecord(evs[5 * k + 2], e->stream));
        dim3 g(gx, a.world);
        k_index_build<<<g, 256, 0, e->stream>>>(e->v, a, e->xa);
        if (perKernel) CK(cudaEventRecord(evs[5 * k + 3], e->stream));
        launch_synapse_pass(e, a, std::max<uint32_t>(ts.fires, 64u));
        if (perKernel) CK(cudaEventRecord(evs[5 * k + 4], e->stream));
        k_index_reset<<<g, 256, 0, e->stream>>>(e->v, a);
        k_finish_step<<<1, 32, 0, e->stream>>>(e->v, e->dOut, 1, e->dWin, e->xa, e->rtab, (e->randOn && e->cfg.world == 1) ? e->dRandState : nullptr);
        { int rc = rand_after_window(e); if (rc) return rc; }
        e->launches += 4;
    }
    CK(cudaEventRecord(e1, e->stream));
    CK(cudaGetLastError());
    int rc = finish_counters(e, hidden, st);
    CK(cudaMemsetAsync(e->dOut, 0, 10 * sizeof(unsigned long long), e->stream));
    if (msTotal) CK(cudaEventElapsedTime(msTotal, e0, e1));
    if (perKernel) {
        float s1 = 0, s2 = 0, sx = 0, sS = 0, x;
        for (uint32_t k = 0; k < count; k++) {
            CK(cudaEventElapsedTime(&x, evs[5 * k], evStage[k])); sS += x;
            CK(cudaEventElapsedTime(&x, evs[5 * k], evs[5 * k + 1])); s1 += x;
            CK(cudaEventElapsedTime(&x, evs[5 * k + 1], evs[5 * k + 3])); sx += x;  // end of the neuron pass -> fire index of all shards built
            CK(cudaEventElapsedTime(&x, evs[5 * k + 3], evs[5 * k + 4])); s2 += x;
        }
        if (msP1) *msP1 = s1;
        if (msP2) *msP2 = s2;
        if (msXchg) *msXchg = sx;
        e->breakdown[0] = sS; e->breakdown[1] = s1 - sS; e->breakdown[2] = sx; e->breakdown[3] = s2;
        for (auto& x2 : evs) cudaEventDestroy(x2);
        for (auto& x2 : evStage) cudaEventDestroy(x2);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return rc;
}

extern "C" int nc_rand_set_state(nc_engine* e, const uint32_t* x31) {
    if (!x31) return fail(e, NC_ERR_INVALID, "nc_rand_set_state: null state");
    cudaSetDevice(e->cfg.device);
    memcpy(e->hRand, x31, 31 * 4);
    if (e->taping) { memcpy(e->tapeRandPend, x31, 31 * 4); e->tapeRandPending = true; }
    CK(cudaMemcpyAsync(e->dRandState, e->hRand, 31 * 4, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));  // (hRand is reused by the read-back of the next window)
    e->randOn = true; e->hRandFresh = true;
    return NC_OK;
}
extern "C" int nc_rand_get_state(nc_engine* e, uint32_t* x31) {
    if (!x31) return fail(e, NC_ERR_INVALID, "nc_rand_get_state: null output");
    if (!e->randOn) return fail(e, NC_ERR_STATE, "nc_rand_get_state: no stream state has been set");
    cudaSetDevice(e->cfg.device);
    if (!e->hRandFresh) CK(cudaMemcpyAsync(e->hRand, e->dRandState, 31 * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->hRandFresh = true;
    memcpy(x31, e->hRand, 31 * 4);
    return NC_OK;
}
extern "C" int nc_background_draw(nc_engine* e, float t0, float runSpeed, uint32_t period, uint64_t nNeurons) {
    if (!e->uploaded) return fail(e, NC_ERR_STATE, "nc_background_draw: no network uploaded");
    if (!e->randOn) return fail(e, NC_ERR_STATE, "nc_background_draw: set the stream state first (nc_rand_set_state)");
    if (e->pending) return fail(e, NC_ERR_STATE, "nc_background_draw: a window is in flight");
    if (period == 0 || nNeurons == 0 || nNeurons >= (1ull << 31)) return fail(e, NC_ERR_INVALID, "nc_background_draw: bad period / neuron count");
    cudaSetDevice(e->cfg.device);
    // one test draw per neuron + two more per hit; room for four times the expected number of hits (never more than one per neuron)
    const uint64_t room = std::min<uint64_t>(nNeurons, 4 * (nNeurons / period) + 256);
    ncr::BgArgs a = {};
    a.t0 = t0; a.runSpeed = runSpeed; a.period = period; a.nNeurons = nNeurons; a.nDraws = nNeurons + 2 * room + 64;
    a.row0 = e->v.row0; a.nRows = e->v.nRows;
    if (a.nDraws > e->drawsCap) {
        cudaFree(e->dDraws);
        e->drawsCap = a.nDraws + a.nDraws / 8;
        CK(cudaMalloc(&e->dDraws, e->drawsCap * 4));
    }
    if (a.nDraws >= (1ull << 32)) return fail(e, NC_ERR_INVALID, "nc_background_draw: too many neurons for 32-bit draw positions");
    const uint32_t wantCap = (uint32_t)std::min<uint64_t>(room + 64, 1u << 26);
    if (wantCap > e->bgCap) {
        cudaFree(e->dBgTmp); cudaFree(e->dBgEv); cudaFree(e->dCand); cudaFree(e->dCandRaw);
        e->bgCap = wantCap;
        e->candListCap = (uint32_t)std::min<uint64_t>(2ull * wantCap + 1024, 1u << 27);
        CK(cudaMalloc(&e->dCand, (size_t)e->candListCap * 4));
        CK(cudaMalloc(&e->dCandRaw, (size_t)e->candListCap * 4));
        CK(cudaMalloc(&e->dBgTmp, (size_t)e->bgCap * sizeof(nc_event)));
        CK(cudaMalloc(&e->dBgEv, (size_t)e->bgCap * sizeof(nc_event)));
    }
    int rc = launch_background(e, a);
    if (rc) return rc;
    e->bgActive = true; e->bgFirstWindow = true; e->hRandFresh = false;
    if (e->taping) { e->bgPendingTape = true; e->bgPendingArgs = a; }
    return NC_OK;
}
extern "C" int nc_background_clear(nc_engine* e) { e->bgActive = false; return NC_OK; }
extern "C" int nc_background_read(nc_engine* e, uint32_t capacity, nc_event* out, uint32_t* count, uint32_t* hits) {
    if (!e->bgActive) return fail(e, NC_ERR_STATE, "nc_background_read: no background draw is active");
    cudaSetDevice(e->cfg.device);
    uint32_t ctl[4];
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(ctl, e->dBgCtl, 16, cudaMemcpyDeviceToHost));
    if (ctl[2]) return fail(e, NC_ERR_CAPACITY, "background firing: more hits than the device generator made room for");
    if (count) *count = ctl[0];
    if (hits) *hits = ctl[1];
    if (out && capacity) CK(cudaMemcpy(out, e->dBgEv, (size_t)std::min(capacity, ctl[0]) * sizeof(nc_event), cudaMemcpyDeviceToHost));
    return NC_OK;
}
extern "C" int nc_index_stats(nc_engine* e, uint64_t* out2) {
    if (!out2) return fail(e, NC_ERR_INVALID, "nc_index_stats: null output");
    cudaSetDevice(e->cfg.device);
    CK(cudaStreamSynchronize(e->stream));
    unsigned long long h[2];
    CK(cudaMemcpy(h, e->v.stats + 10, 16, cudaMemcpyDeviceToHost));
    CK(cudaMemset(e->v.stats + 10, 0, 16));
    out2[0] = h[0]; out2[1] = h[1];
    return NC_OK;
}
extern "C" int nc_replay_breakdown(const nc_engine* e, float* out4) {
    if (!out4) return NC_ERR_INVALID;
    for (int i = 0; i < 4; i++) out4[i] = e->breakdown[i];
    return NC_OK;
}

// ------------------------------------------------------------------------------------------------
// Self-test entry points: evaluate the device math replicas on caller-provided arrays so that tests can
// compare them bit-for-bit with the host libm the reference uses (SURVEY.md hard part #1).
// ------------------------------------------------------------------------------------------------
__global__ void k_selftest_powf(const float* x, const float* y, float* out, uint64_t n) {
    math_tables_to_shared();
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = powf_pos(x[i], y[i]);
}
__global__ void k_selftest_exp(const double* x, double* out, uint64_t n) {
    math_tables_to_shared();
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = exp_glibc(x[i]);
}
extern "C" int nc_selftest_powf(nc_engine* e, const float* x, const float* y, float* out, uint64_t n) {
    cudaSetDevice(e->cfg.device);
    float *dx, *dy, *dout;
    CK(cudaMalloc(&dx, n * 4)); CK(cudaMalloc(&dy, n * 4)); CK(cudaMalloc(&dout, n * 4));
    CK(cudaMemcpyAsync(dx, x, n * 4, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(dy, y, n * 4, cudaMemcpyHostToDevice, e->stream));
    k_selftest_powf<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(dx, dy, dout, n); e->launches++;
    CK(cudaMemcpyAsync(out, dout, n * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(dx); cudaFree(dy); cudaFree(dout);
    return NC_OK;
}
extern "C" int nc_selftest_exp(nc_engine* e, const double* x, double* out, uint64_t n) {
    cudaSetDevice(e->cfg.device);
    double *dx, *dout;
    CK(cudaMalloc(&dx, n * 8)); CK(cudaMalloc(&dout, n * 8));
    CK(cudaMemcpyAsync(dx, x, n * 8, cudaMemcpyHostToDevice, e->stream));
    k_selftest_exp<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(dx, dout, n); e->launches++;
    CK(cudaMemcpyAsync(out, dout, n * 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(dx); cudaFree(dout);
    return NC_OK;
}
