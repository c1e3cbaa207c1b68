// Per-neuron and per-synapse step logic shared by the CUDA kernels (engine.cu) and — compiled for the
// host by tests/native/model_twopass.cpp — by the CPU test model that lets the algorithm be checked
// against the oracle without a GPU.  Nothing here is a fallback: the product library only ever runs
// these functions inside kernels.
//
// Reference semantics restated (file:line under /root/reference/src):
//   chain_* / neuron_run_begin / neuron_run_finish   Neuron::run NeuCor.cpp:619-641 (charge_insynapses :688-700,
//                  charge_passive :677-680, charge_thresholdCheck :682-686, Neuron::fire :643-645, AP :706-714, activity :640)
//   plasticity     Synapse::synapticPlasticity NeuCor.cpp:740-764, Neuron::getTrace :671-675
//   resolve_slot   Synapse::fire :727-738, the slot clear at :697, Synapse::run :718-726
// Every operator keeps the reference's float/double typing; no contraction (explicit-rounding helpers).
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/neucor_b200.h"
#include "glibc_math.cuh"

#if !defined(__CUDACC__)
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#endif

namespace ncs {
using namespace ncm;

struct FireRec {       // 16 B; canonical key of the event that caused the fire
    uint32_t neuron;   // global ID
    float time;
    uint32_t rk1;      // rank << 30 | k1
    uint32_t k2;
};
// Canonical event order: (time, rank, k1, k2)
//   rank 0 input firer  (k1 = firer index, k2 = neuron)      InputFirer::run
//   rank 1 delivery     (k1 = target,      k2 = the synapse's index within the target's row: ascends with the parent ID)   Synapse::run
//   rank 2 neuron event (k1 = neuron,      k2 = 0)           Neuron::run queued by transfer/scheduleFire
//   rank 3 sweep        (k1 = neuron,      k2 = 0)           end-of-window run of every neuron, ascending ID
// Within one event the operations on a synapse are ordered load/clear (0) < post-fire plasticity (1) < delivery (2).

#define NC_SENT 0x80000000u  // arrive[] bit-pattern marker: slot cleared during this window's neuron pass
                             // bits 30..29 = rank of the clearing event, bits 28..0 = delivering slot (rank 1);
                             // the clear time is parked in depol[] (dead once the slot is idle)
#define NC_MAX_WORLD 8

struct Key {
    float t;
    uint32_t rk1, k2, local;
};
NC_HD bool key_less(const Key& a, const Key& b) {
    if (a.t != b.t) return a.t < b.t;
    if (a.rk1 != b.rk1) return a.rk1 < b.rk1;
    if (a.k2 != b.k2) return a.k2 < b.k2;
    return a.local < b.local;
}
NC_HD bool key_le3(float t, uint32_t rk1, uint32_t k2, const Key& b) {  // (t, rk1, k2) <= b's, ignoring `local`
    if (t != b.t) return t < b.t;
    if (rk1 != b.rk1) return rk1 < b.rk1;
    return k2 <= b.k2;
}
NC_HD bool pick_less(float t1, unsigned long long c1, float t2, unsigned long long c2) {
    return t1 < t2 || (t1 == t2 && c1 < c2);
}

struct SlotRow {       // 8 B: a synapse slot of this shard and the row (target neuron - row0) it belongs to
    uint32_t slot, row;
};

struct SynRec {        // 16 B: what resolving a synapse needs besides its (arrive, depol) record — one sector per visit
    uint32_t pre;      // presynaptic neuron (global ID); bit 31 = Synapse::inhibitory
    float weight;      // Synapse::weight
    float lastArr;     // Synapse::lastSpikeArrival (-inf initially)
    float delay;       // length * AP_speed, computed once in fp32 exactly as NeuCor.cpp:733
};
struct FlagEnt {       // 16 B: a slot the neuron pass hands to the synapse pass (delivered in the window / cleared by one of its runs)
    uint32_t slot, row;
    float lfStart;     // the row neuron's lastFire at the start of the window
    uint32_t inRow;    // the slot's index within its row
};

struct View {
    uint64_t nGlobal, row0, nRows, S;
    const uint64_t* rowptr;
    SynRec* rec;
    float2* ad;  // per slot (arrive, depol): Synapse::AP_fireTime (0 = idle) and AP_depolFac of the spike in flight — one 8-byte record, one sector per visit
    float* lastStart;  // Synapse::lastSpikeStart (renderer only)
    float2* potAct;
    float *lastRan, *lastFire, *lfStart, *actStart;
    uint32_t* firings;
    // exchange
    uint32_t* localHdr;  // [0] = fire count of this shard, [1] = overflow flag
    FireRec* localRecs;
    uint32_t fireCap;
    // gathered fire index.  A shard's exchange block is a run of 16-byte units: unit 0 is the header
    // {count, overflow, 0, 0}, units 1..count are the records; localHdr/localRecs are this shard's own block.
    const FireRec* gRecs;  // `world` blocks of `gStride` units each (world 1: the local block itself)
    int32_t* head;         // per global neuron: unit index of its first record or -1
    int32_t* next;         // per unit
    uint32_t* mask;        // one bit per global neuron: fired in this window
    uint32_t* evMask;      // one bit per row of this shard: has host events in this window
    // event index: which slots can matter to a window, so that neither pass visits idle synapses
    uint32_t* busy;        // one bit per slot: arrive != 0 (a spike is in flight or being integrated); NULL in the CPU test double
    uint32_t* arrived;     // one bit per slot, subset of busy: the spike has arrived (it was staged by an earlier window)
    uint32_t* wordNext;    // per 32-slot word: bit pattern of a lower bound on the earliest arrival among its slots still in flight (+inf: none)
    FlagEnt* flagList;     // written by the neuron pass: slots that delivered in this window or were cleared by one of its runs
    uint32_t* flagCtl;     // [0] = entries in flagList, [1] = overflow flag
    uint32_t flagCap;
    // staged rows (k_stage -> k_neuron_pass): per tile of 32 rows a region of stCap entries holding the tile's occupied slots
    // (arrive <= t1) in row order — (arrive, depol) and the slot index relative to the tile's first slot — and per row its count
    // The lists PERSIST across windows: a slot's (arrive, depol) never changes while it is in the list, so the staging kernel
    // only merges the (few) slots that arrived in this window and drops the ones the last neuron pass cleared (it negates their
    // arrive in the list).  Two regions per tile (ping-pong); tileState says which one is current.
    uint32_t* stCnt;       // per row; 0xffffffff on every row of a tile whose occupied slots did not fit its region
    float2* stAD;          // [tile][2][stCap]
    uint32_t* stJ;
    uint32_t stCap;
    uint32_t stageRebuild; // 1: every window rebuilds every list from the index (NC_STAGE_MODE=rebuild; measurements)
    uint32_t* tileState;   // per tile: bit 0 = the list is valid, bit 1 = which region holds it, bit 2 = the last neuron pass cleared entries
    uint32_t cprLoads, cprRows;  // 128-entry chunks that cover the longest out-list / the longest row (work split of the synapse kernels)
    const uint64_t* cscPtr;  // per GLOBAL presynaptic neuron: its out-synapses that land in this shard are cscEnt[cscPtr[p] .. cscPtr[p+1])
    const SlotRow* cscEnt;
    // spill area for rows with more occupied slots than fit in shared memory
    float* spillA;
    float* spillD;
    uint32_t* spillJ;
    uint32_t spillPerWarp;
    unsigned long long* stats;  // 8 counters (nc_step_stats order)
    uint32_t* tileCtr;          // [0] neuron pass, [1] synapse pass: next unclaimed tile of the window (dynamic scheduling)
};

struct StepArgs {
    float t0, t1;
    int sweep;  // NC_SWEEP_END | NC_SWEEP_START
    float lr, preFactor, postFactor, preDecay, postDecay;
    const nc_event* ev;
    uint32_t nEv;
    const uint32_t* nEvDev;  // when set: the number of events is read from the device (host events merged with the device-generated background events)
    const uint32_t* subset;  // nc_run_neurons: ascending IDs that take part in the sweep (NULL = all)
    uint32_t nSubset;
    uint32_t candCap;
    uint32_t variant;  // which build of the neuron pass runs this window (0: 1024-slot pool, 1: 704, 2: 512)
    uint32_t gStride;  // units per gathered block (header included)
    uint32_t world;
    // float comparisons of the row scan as integer compares on the bit patterns of (positive) arrival times, fixed per window:
    //   t1 - a > 2  (old enough to be cleared)  <=>  bits(a) <= clrB;     t0 < a + 2 <= t1 (requeue)  <=>  reqLoB < bits(a) <= reqHiB
    uint32_t clrB, reqLoB, reqHiB;
};

struct NeuronState {
    float pot, act, lastRan, lastFire, actStart, sched;
    uint32_t firings;
};

struct CandView {  // occupied slots of the current row: shared memory first, spill area after
    float* a;      // arrive time; negated once cleared in this window
    float* d;      // depolarisation factor
    uint32_t* j;   // slot index within the row
    float* sa;
    float* sd;
    uint32_t* sj;
    uint32_t cap;
    NC_HDM float& A(uint32_t c) { return c < cap ? a[c] : sa[c - cap]; }
    NC_HDM float& D(uint32_t c) { return c < cap ? d[c] : sd[c - cap]; }
    NC_HDM uint32_t& J(uint32_t c) { return c < cap ? j[c] : sj[c - cap]; }
};

NC_HD void emit_fire(const View& v, uint32_t q, float T, uint32_t rk1, uint32_t k2) {
#if defined(__CUDA_ARCH__)
    uint32_t idx = atomicAdd(&v.localHdr[0], 1u);
#else
    uint32_t idx = v.localHdr[0]++;
#endif
    if (idx < v.fireCap) {
        FireRec r;
        r.neuron = q; r.time = T; r.rk1 = rk1; r.k2 = k2;
        v.localRecs[idx] = r;
    } else {
        v.localHdr[1] = 1u;
    }
}

// ---- ordered accumulation over the active slots (charge_insynapses, NeuCor.cpp:688-700) -------------------------
// Reference semantics, one slot after the other in ascending presynaptic ID:
//     newPot = (float)((double)newPot + (double)(deltaT*depolFac) * 0.9943 * exp(0.3702*deltaT))
// i.e. every addition is rounded to double and then to float, so the result depends on the order.  The kernels
// evaluate a whole group of 32 slots at once without changing a single bit:
//   while the running value stays strictly inside one binade [2^e, 2^(e+1)) its float grid is u = 2^(e-23) and its
//   double grid g = 2^(e-52); the running value is a multiple of u, so   fl64(np + t) = np + RN_g(t)   and
//   fl32(np + t') = np + RN_u(t')  — both roundings act on the TERM alone (ties-to-even included for the first one
//   because np is an even multiple of g).  RN_u(RN_g(t)) is therefore an order-independent per-slot quantity r_i, the
//   partial sums  np + r_1 + ... + r_i  are exact in double, and a warp prefix sum reproduces the serial chain.
//   The three cases where this does not hold are detected and that one slot is added serially in plain double
//   arithmetic before the group continues: (1) an exact tie of the float rounding (then the parity of the running value
//   decides), (2) a partial sum that leaves the open binade (the grids change), (3) |t| >= 2^(e-1) or np not a
//   normal number.
NC_HD double chain_term(float dT, float depol, double E) { return mul64(mul64((double)mul32(dT, depol), 0.9943), E); }

struct Binade {
    double lo, hi;   // 2^e, 2^(e+1)
    double C1, C2;   // 1.5*2^e (ulp = g), 1.5*2^(e+29) (ulp = u)
    double halfu;    // u/2
    double tmax;     // 2^(e-1)
    int ok;
};
NC_HD Binade binade_of(float np) {
    Binade b;
    uint32_t ex = (as_u32(np) >> 23) & 0xffu;
    b.ok = (ex != 0u && ex != 255u);
    unsigned long long e = (unsigned long long)ex + (1023ull - 127ull);  // biased double exponent of 2^e
    b.lo = as_f64(e << 52);
    b.hi = as_f64((e + 1ull) << 52);
    b.C1 = as_f64((e << 52) | 0x0008000000000000ULL);
    b.C2 = as_f64(((e + 29ull) << 52) | 0x0008000000000000ULL);
    b.halfu = as_f64((e - 24ull) << 52);
    b.tmax = as_f64((e - 1ull) << 52);
    return b;
}
// r = this slot's contribution to the MAGNITUDE of the running value, rounded as the reference's two roundings would;
// returns false when the slot must be added serially (cases 1 and 3 above; case 2 is checked on the prefix sums).
NC_HD bool chain_lane(const Binade& b, bool neg, double t, double& r) {
    if (!b.ok || !(fabs(t) < b.tmax)) { r = 0.0; return false; }
    double te = neg ? -t : t;
    double tp = sub64(add64(te, b.C1), b.C1);
    r = sub64(add64(tp, b.C2), b.C2);
    return fabs(sub64(tp, r)) != b.halfu;
}

// Neuron::run, part 1: deltaT bookkeeping. Returns false when no time has passed (NeuCor.cpp:626).
NC_HD bool neuron_run_begin(NeuronState& n, float T, float& dT) {
    dT = sub32(T, n.lastRan);
    n.lastRan = T;
    return dT != 0.0f;
}
// Neuron::run, part 3 (after charge_insynapses produced `np`): passive decay, threshold check, AP waveform, activity.
// Returns true when the neuron fired; the caller emits the fire record.
NC_HD float neuron_activity(uint32_t firings, float T, float actStart) {  // NeuCor.cpp:640
    return (float)div64((double)firings, div64((double)sub32(T, actStart), 10.0));
}
// withAct = false: the caller evaluates neuron_activity() once, for the last run of the window (activity is only ever read
// back, never fed into the dynamics, so intermediate values are dead).
NC_HD bool neuron_run_finish(NeuronState& n, float np, float T, float dT, bool withAct = true) {
    const float baselevel = -70.0f, threshold = -55.0f, recharge = 0.5f, AP_cutoff = 2.0f;
    // charge_passive (NeuCor.cpp:677-680)
    np = add32(mul32(sub32(np, baselevel), powf_pos(recharge, dT)), baselevel);
    n.pot = np;
    // charge_thresholdCheck (NeuCor.cpp:682-686; the vesicle term is always true)
    bool fired = false;
    float lf = n.lastFire;
    if ((threshold < n.pot || n.sched == T) && (lf != lf || AP_cutoff < sub32(T, lf))) {
        n.lastFire = T;  // Neuron::fire, NeuCor.cpp:644-645
        n.firings++;
        fired = true;
    }
    // AP (NeuCor.cpp:706-714): analytic double-Gaussian waveform while within the cutoff; powf(x, 2.0) is x*x at -O3
    lf = n.lastFire;
    if (!(lf != lf || AP_cutoff < sub32(T, lf))) {
        const double D1 = (2.0 * (double)0.3f) * (double)0.3f, D2 = (2.0 * (double)0.6f) * (double)0.6f;
        float t = sub32(T, lf);
        float x1 = sub32(t, 1.0f);
        float x2 = sub32(sub32(t, 1.0f), 1.16f);
        double e1 = exp_glibc(div64((double)(-mul32(x1, x1)), D1));
        double e2 = exp_glibc(div64((double)(-mul32(x2, x2)), D2));
        double wave = mul64(100.0, sub64(e1, mul64(e2, (double)0.2f)));
        double tail = mul64((double)sub32(threshold, baselevel), fmax(sub64(add64(1.0, (double)lf), (double)T), 0.0));
        n.pot = (float)add64(add64(wave, (double)baselevel), tail);
    }
    // activity (NeuCor.cpp:640)
    if (withAct) n.act = neuron_activity(n.firings, T, n.actStart);
    return fired;
}

// Synapse::synapticPlasticity at time T with the target's lastFire `lfq`.
NC_HD float plasticity(const StepArgs& s, float w, bool inh, float T, float lastArr, float lfq, uint32_t& hidden) {
    float traceS = powf_pos(s.preDecay, sub32(T, lastArr));
    float traceT = powf_pos(s.postDecay, sub32(T, lfq));
    if (traceT != traceT) traceT = 0.0f;
    if (traceT == 1.0f) traceT = 0.0f;
    if (traceS == 1.0f) traceS = 0.0f;
    if (w == 0.0f && !inh) hidden++;  // the rand() hidden in NeuCor.cpp:752's short-circuit
    float change = sub32(mul32(s.preFactor, traceS), mul32(s.postFactor, traceT));
    w = add32(w, mul32(change, s.lr));
    if (!inh) w = (float)fmax(fmin((double)w, 1.0), 0.0);
    else w = (float)fmax(fmin((double)w, 0.0), -1.0);
    return w;
}

// All operations of one window on synapse j = (p -> q), applied in canonical order:
//   L  one per fire of p   Synapse::fire (slot busy -> dropped)
//   C  slot cleared by one of q's runs (recorded by the neuron pass)
//   P  one per fire of q   synapticPlasticity from Neuron::fire
//   D  delivery when t0 < arrive <= t1: lastSpikeArrival = now; synapticPlasticity
// cnt: [0] loads accepted, [1] loads dropped, [2] plasticity calls, [3] hidden rand, [4] deliveries
NC_HD void resolve_slot(const View& v, const StepArgs& s, uint64_t j, uint32_t inRow, uint32_t q, const SynRec r0,
                        uint32_t abits, bool pFired, bool qFired, float lfStart, uint32_t* cnt) {
    const uint32_t p = r0.pre & 0x7fffffffu;
    const bool inh = (r0.pre >> 31) != 0u;
    float a = as_f32(abits);
    const bool cleared = (abits & NC_SENT) != 0u;
    Key kc;
    kc.local = 0; kc.t = 0.0f; kc.rk1 = 0; kc.k2 = 0;
    if (cleared) {
        uint32_t rank = (abits >> 29) & 3u;
        kc.rk1 = (rank << 30) | q;
        if (rank == 3u) { kc.t = s.t1; kc.k2 = 0u; }
        else if (rank == 2u) { kc.t = v.ad[j].y; kc.k2 = 0u; }
        else { kc.t = v.ad[j].y; kc.k2 = abits & 0x1fffffffu; }  // the delivering slot's index within the row
    }
    const bool hasD = !cleared && a != 0.0f && a > s.t0 && a <= s.t1;
    Key kd;
    kd.t = a; kd.rk1 = (1u << 30) | q; kd.k2 = inRow; kd.local = 2;
    bool busy = cleared || a != 0.0f;
    float w = r0.weight, lastArr = r0.lastArr;
    float newArrive = cleared ? 0.0f : a, newDepol = 0.0f, newStart = 0.0f;
    bool loaded = false, wChanged = false, laChanged = false;
    Key cur;
    cur.t = 0.0f; cur.rk1 = 0; cur.k2 = 0; cur.local = 0;
    bool have = false;
    for (;;) {
        Key best;
        best.t = 0.0f; best.rk1 = 0; best.k2 = 0; best.local = 0;
        int kind = -1;  // 0 L, 1 C, 2 P, 3 D
        float bestT = 0.0f;
        if (pFired)
            for (int32_t f = v.head[p]; f >= 0; f = v.next[f]) {
                FireRec r = v.gRecs[f];
                Key k; k.t = r.time; k.rk1 = r.rk1; k.k2 = r.k2; k.local = 0;
                if ((!have || key_less(cur, k)) && (kind < 0 || key_less(k, best))) { best = k; kind = 0; bestT = r.time; }
            }
        if (cleared && (!have || key_less(cur, kc)) && (kind < 0 || key_less(kc, best))) { best = kc; kind = 1; }
        if (qFired)
            for (int32_t f = v.head[q]; f >= 0; f = v.next[f]) {
                FireRec r = v.gRecs[f];
                Key k; k.t = r.time; k.rk1 = r.rk1; k.k2 = r.k2; k.local = 1;
                if ((!have || key_less(cur, k)) && (kind < 0 || key_less(k, best))) { best = k; kind = 2; bestT = r.time; }
            }
        if (hasD && (!have || key_less(cur, kd)) && (kind < 0 || key_less(kd, best))) { best = kd; kind = 3; }
        if (kind < 0) break;
        if (kind == 0) {
            if (busy) cnt[1]++;
            else {
                float d = (float)mul64((double)0.2f, 52.0);  // AP_depolFac *= 52.0 (float *= double)
                newDepol = mul32(d, w);                      // AP_depolFac *= weight
                newArrive = add32(r0.delay, bestT);          // length*AP_speed + now
                newStart = bestT;
                busy = true; loaded = true;
                cnt[0]++;
            }
        } else if (kind == 1) {
            busy = false;
            newArrive = 0.0f;
        } else {
            float T = (kind == 2) ? bestT : a;
            if (kind == 3) { lastArr = T; laChanged = true; cnt[4]++; }
            // the target's lastFire as of this operation: latest fire of q whose event key is <= this one's
            float lfq = lfStart;
            if (qFired) {
                Key lk; lk.t = 0.0f; lk.rk1 = 0; lk.k2 = 0; lk.local = 0;
                bool lhave = false;
                for (int32_t f = v.head[q]; f >= 0; f = v.next[f]) {
                    FireRec r = v.gRecs[f];
                    if (!key_le3(r.time, r.rk1, r.k2, best)) continue;
                    Key k; k.t = r.time; k.rk1 = r.rk1; k.k2 = r.k2; k.local = 0;
                    if (!lhave || key_less(lk, k)) { lk = k; lhave = true; lfq = r.time; }
                }
            }
            w = plasticity(s, w, inh, T, lastArr, lfq, cnt[3]);
            wChanged = true;
            cnt[2]++;
        }
        cur = best; have = true;
    }
    if (cleared || loaded) {
        v.ad[j].x = newArrive;
        if (v.busy) {  // keep the event index in step (integer atomics only)
            const uint32_t bit = 1u << (uint32_t)(j & 31u);
#if defined(__CUDA_ARCH__)
            if ((newArrive != 0.0f) != (cleared || a != 0.0f)) { if (newArrive != 0.0f) atomicOr(&v.busy[j >> 5], bit); else atomicAnd(&v.busy[j >> 5], ~bit); }
            if (cleared) atomicAnd(&v.arrived[j >> 5], ~bit);                      // an idle or re-loaded slot has not arrived
            if (loaded) atomicMin(&v.wordNext[j >> 5], as_u32(newArrive));         // positive floats order as unsigned integers
#else
            if ((newArrive != 0.0f) != (cleared || a != 0.0f)) { if (newArrive != 0.0f) v.busy[j >> 5] |= bit; else v.busy[j >> 5] &= ~bit; }
#endif
        }
    }
    if (loaded) { v.ad[j].y = newDepol; v.lastStart[j] = newStart; }
    if (wChanged) v.rec[j].weight = w;
    if (laChanged) v.rec[j].lastArr = lastArr;
}

}  // namespace ncs
