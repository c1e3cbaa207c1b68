// TEST DOUBLE — a single-threaded CPU model of the device engine behind include/neucor_b200.h.
//
// It exists so that the two-pass algorithm and the host class can be checked against the oracle in
// CI without a GPU.  It is built only by tests/ into tests/native/_build/ and is
// never part of, linked into, or loaded by the product (libneucor_b200.so has no CPU path).  The
// per-neuron / per-synapse logic is the SAME source the kernels compile (csrc/step_logic.cuh); only
// the warp-cooperative scaffolding of engine.cu (ballot staging, shuffle min-reduction, grid loops) is
// restated serially here.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../neurocorrelation_b200/csrc/step_logic.cuh"

using namespace ncm;
using namespace ncs;

struct nc_engine {
    std::string err;
    View v;
    std::vector<uint64_t> rowptr;
    std::vector<SynRec> rec;
    std::vector<uint32_t> firings, hdr, mask, spillJ;
    std::vector<float2> ad;
    std::vector<float> lastStart, lastRan, lastFire, lfStart, actStart, spillA, spillD;
    std::vector<float2> potAct;
    std::vector<FireRec> units, gather;  // exchange block (header unit + records) and the gathered blocks of all shards
    std::vector<int32_t> head, next;
    int rank = 0, world = 1;
    nc_allgather_fn xchgFn = nullptr; void* xchgCtx = nullptr;
    uint32_t xchgUnits = 1u + 4u;  // deliberately tiny so that the grow-and-repeat path of the exchange is exercised
    uint32_t counts[NC_MAX_WORLD] = {0}; uint32_t stride = 0;
    unsigned long long stats[8];
    float lr = 1.0f, preF = 0.13f, postF = 0.30f, preD = 0.75f, postD = 0.65f, minDelay = INFINITY;
    uint32_t candCap = 8;  // deliberately tiny so that the spill path is exercised
    bool uploaded = false;
    // serial model of the device-resident rand() stream (csrc/rand_stream.cuh): x[n] = x[n-31] + x[n-3], rand() = x[n] >> 1
    uint32_t rstate[31]; bool randOn = false;
    std::vector<nc_event> bg; bool bgActive = false, bgFirst = false;
    uint32_t rs_next() { uint32_t x = rstate[0] + rstate[28]; memmove(rstate, rstate + 1, 30 * 4); rstate[30] = x; return x; }
};
static std::string g_err;
static int fail(nc_engine* e, int code, const char* m) { e->err = m; return code; }

extern "C" {
const char* nc_global_error(void) { return g_err.c_str(); }
const char* nc_last_error(const nc_engine* e) { return e ? e->err.c_str() : g_err.c_str(); }
int nc_device_count(void) { return 1; }
uint64_t nc_launch_count(const nc_engine*) { return 0; }
int nc_index_stats(nc_engine*, uint64_t* out2) { out2[0] = out2[1] = 0; return NC_OK; }
int nc_replay_breakdown(const nc_engine*, float* out4) { for (int i = 0; i < 4; i++) out4[i] = 0.0f; return NC_OK; }
int nc_create(const nc_config* cfg, nc_engine** out) {
    nc_engine* e = new nc_engine();
    memset(&e->v, 0, sizeof(View));
    memset(e->stats, 0, sizeof(e->stats));
    if (cfg->cand_smem) e->candCap = cfg->cand_smem;
    e->rank = cfg->rank; e->world = cfg->world;
    *out = e;
    return NC_OK;
}
void nc_destroy(nc_engine* e) { delete e; }
int nc_set_plasticity(nc_engine* e, float lr, float a, float b, float c, float d) { e->lr = lr; e->preF = a; e->postF = b; e->preD = c; e->postD = d; return NC_OK; }
int nc_min_delay(const nc_engine* e, float* out) { *out = e->minDelay; return NC_OK; }

int nc_upload_network(nc_engine* e, uint64_t nGlobal, uint64_t row0, uint64_t nRows, const uint64_t* rowptr, const uint32_t* pre,
                      const float* weight, const float* length, const uint8_t* inh) {
    uint64_t S = rowptr[nRows], maxRow = 0;
    e->rowptr.assign(rowptr, rowptr + nRows + 1);
    e->rec.resize(S);
    e->ad.assign(S, make_float2(0.0f, 0.0f)); e->lastStart.assign(S, 0.0f);
    for (uint64_t i = 0; i < S; i++) {
        SynRec r; r.pre = pre[i] | (inh[i] ? 0x80000000u : 0u); r.weight = weight[i]; r.lastArr = -INFINITY; r.delay = length[i] * 2.0f;
        e->rec[i] = r; e->minDelay = std::min(e->minDelay, r.delay);
    }
    for (uint64_t r = 0; r < nRows; r++) maxRow = std::max(maxRow, rowptr[r + 1] - rowptr[r]);
    e->potAct.assign(nRows, make_float2(-70.0f, 0.0f)); e->lastRan.assign(nRows, 0.0f); e->lastFire.assign(nRows, NAN);
    e->lfStart.assign(nRows, NAN); e->actStart.assign(nRows, 0.0f); e->firings.assign(nRows, 0u);
    uint32_t cap = (uint32_t)(4 * nRows + 1024);
    e->units.assign((size_t)cap + 1, FireRec{0, 0, 0, 0}); e->gather.assign((size_t)e->world * (cap + 1), FireRec{0, 0, 0, 0});
    e->head.assign(nGlobal, -1); e->next.assign((size_t)e->world * (cap + 1), -1); e->mask.assign((nGlobal + 31) / 32 + 1, 0u);
    e->spillA.resize(maxRow + 1); e->spillD.resize(maxRow + 1); e->spillJ.resize(maxRow + 1);
    View& v = e->v;
    v.nGlobal = nGlobal; v.row0 = row0; v.nRows = nRows; v.S = S; v.rowptr = e->rowptr.data(); v.rec = e->rec.data();
    v.ad = e->ad.data(); v.lastStart = e->lastStart.data();
    v.potAct = e->potAct.data(); v.lastRan = e->lastRan.data(); v.lastFire = e->lastFire.data(); v.lfStart = e->lfStart.data();
    v.actStart = e->actStart.data(); v.firings = e->firings.data(); v.localHdr = reinterpret_cast<uint32_t*>(e->units.data()); v.localRecs = e->units.data() + 1; v.fireCap = cap;
    v.gRecs = e->world > 1 ? e->gather.data() : e->units.data(); v.head = e->head.data(); v.next = e->next.data(); v.mask = e->mask.data();
    v.spillA = e->spillA.data(); v.spillD = e->spillD.data(); v.spillJ = e->spillJ.data(); v.spillPerWarp = (uint32_t)maxRow; v.stats = e->stats;
    e->uploaded = true;
    return NC_OK;
}

int nc_upload_network_device(nc_engine* e, uint64_t, uint64_t, uint64_t, const uint64_t*, const uint32_t*, const float*, const float*, const uint8_t*) {
    return fail(e, NC_ERR_INVALID, "the CPU test double has no device memory");
}

// serial emulation of engine.cu's warp_neuron_run: the same 32-slot groups, the same per-lane math (chain_lane),
// the same first-flagged-lane handling — lanes become loop iterations
static void model_neuron_run(const View& v, NeuronState& n, CandView& cv, uint32_t cnt, uint64_t rs, uint32_t q, float T, uint32_t rk1,
                             uint32_t k2, uint32_t sentinel, unsigned long long& nFires, unsigned long long& nRuns, unsigned long long& nVisits) {
    float dT;
    if (!neuron_run_begin(n, T, dT)) return;
    nRuns++;
    float np = n.pot;
    if (cnt) {
        const double E = exp_glibc(mul64(0.3702, (double)dT));
        for (uint32_t base = 0; base < cnt; base += 32) {
            double t[32]; uint32_t todo = 0;
            for (uint32_t lane = 0; lane < 32; lane++) {
                uint32_t c = base + lane; t[lane] = 0.0;
                if (c >= cnt) continue;
                float a = cv.A(c);
                if (!(a > 0.0f)) continue;
                float off = sub32(T, a);
                if (!(off > 0.0f)) continue;
                todo |= 1u << lane;
                t[lane] = chain_term(dT, cv.D(c), E);
                if (2.0f < off) { cv.A(c) = -a; uint64_t sidx = rs + cv.J(c); v.ad[sidx] = make_float2(as_f32(sentinel), T); }
            }
            nVisits += __builtin_popcount(todo);
            while (todo) {
                const Binade b = binade_of(np);
                const bool neg = np < 0.0f;
                double r[32], pre[32]; bool flag[32];
                for (uint32_t lane = 0; lane < 32; lane++) { r[lane] = 0.0; flag[lane] = false; if ((todo >> lane) & 1u) flag[lane] = !chain_lane(b, neg, t[lane], r[lane]); }
                double run = 0.0;
                for (uint32_t lane = 0; lane < 32; lane++) { run = run + r[lane]; pre[lane] = run; }
                const double m = fabs((double)np);
                uint32_t bad = 0;
                for (uint32_t lane = 0; lane < 32; lane++) {
                    double mi = m + pre[lane];
                    if (((todo >> lane) & 1u) && !(mi > b.lo && mi < b.hi)) flag[lane] = true;
                    if (flag[lane] && ((todo >> lane) & 1u)) bad |= 1u << lane;
                }
                if (!bad) { double mm = m + pre[31]; np = (float)(neg ? -mm : mm); todo = 0; }
                else {
                    int f = __builtin_ffs(bad) - 1;
                    double acc = f > 0 ? pre[f - 1] : 0.0;
                    double mm = m + acc;
                    np = (float)(neg ? -mm : mm);
                    np = (float)((double)np + t[f]);
                    todo &= ~((2u << f) - 1u);
                }
            }
        }
    }
    if (neuron_run_finish(n, np, T, dT)) { nFires++; emit_fire(v, q, T, rk1, k2); }
}

// serial restatement of k_neuron_pass
static void model_pass1(nc_engine* e, const StepArgs& s) {
    View& v = e->v;
    std::vector<float> sa(s.candCap), sd(s.candCap);
    std::vector<uint32_t> sj(s.candCap);
    CandView cv;
    cv.a = sa.data(); cv.d = sd.data(); cv.j = sj.data(); cv.cap = s.candCap; cv.sa = v.spillA; cv.sd = v.spillD; cv.sj = v.spillJ;
    unsigned long long nFires = 0, nRuns = 0, nVisits = 0, nDeliv = 0;
    for (uint64_t row = 0; row < v.nRows; row++) {
        const uint32_t q = (uint32_t)(v.row0 + row);
        if (s.subset && !std::binary_search(s.subset, s.subset + s.nSubset, q)) continue;
        const uint64_t rs = v.rowptr[row], re = v.rowptr[row + 1];
        uint32_t cnt = 0;
        for (uint64_t j = rs; j < re; j++) {
            float a = v.ad[j].x;
            if (a != 0.0f && a <= s.t1) { cv.A(cnt) = a; cv.D(cnt) = v.ad[j].y; cv.J(cnt) = (uint32_t)(j - rs); cnt++; }
        }
        uint32_t evLo = (uint32_t)(std::lower_bound(s.ev, s.ev + s.nEv, q, [](const nc_event& x, uint32_t k) { return x.neuron < k; }) - s.ev);
        uint32_t evHi = (uint32_t)(std::upper_bound(s.ev, s.ev + s.nEv, q, [](uint32_t k, const nc_event& x) { return k < x.neuron; }) - s.ev);
        NeuronState n;
        n.pot = v.potAct[row].x; n.act = v.potAct[row].y; n.lastRan = v.lastRan[row]; n.lastFire = v.lastFire[row];
        n.actStart = v.actStart[row]; n.firings = v.firings[row]; n.sched = NAN;
        for (uint32_t k = evLo; k < evHi; k++) if (s.ev[k].kind == 2u && (s.ev[k].index_or_flags & 1u)) n.sched = s.ev[k].time;
        v.lfStart[row] = n.lastFire;
        if (cnt || evHi > evLo || (s.sweep & NC_SWEEP_START)) {
            float curT = s.t0; unsigned long long curC = 0; bool first = true;
            for (;;) {
                float bt = INFINITY; unsigned long long bc = ~0ull; uint32_t bsrc = 0xffffffffu;
                for (uint32_t c = 0; c < cnt; c++) {
                    float a = fabsf(cv.A(c));
                    if (a > s.t0) {
                        unsigned long long code = (1ull << 32) | cv.J(c);
                        if ((first || pick_less(curT, curC, a, code)) && pick_less(a, code, bt, bc)) { bt = a; bc = code; bsrc = c; }
                    }
                    float tR = add32(a, 2.0f);
                    if (tR > s.t0 && tR <= s.t1) {
                        unsigned long long code = (2ull << 32);
                        if ((first || pick_less(curT, curC, tR, code)) && pick_less(tR, code, bt, bc)) { bt = tR; bc = code; bsrc = c; }
                    }
                }
                if ((s.sweep & NC_SWEEP_START) && (first || pick_less(curT, curC, s.t0, 2ull << 32)) && pick_less(s.t0, 2ull << 32, bt, bc)) { bt = s.t0; bc = 2ull << 32; bsrc = 0xfffffffeu; }
                for (uint32_t k = evLo; k < evHi; k++) {
                    nc_event ev = s.ev[k];
                    unsigned long long code = ev.kind == 0u ? (unsigned long long)ev.index_or_flags : (2ull << 32);
                    bool after = first ? (ev.time >= s.t0) : pick_less(curT, curC, ev.time, code);
                    if (after && ev.time <= s.t1 && pick_less(ev.time, code, bt, bc)) { bt = ev.time; bc = code; bsrc = 0x80000000u | k; }
                }
                if (bc == ~0ull) break;
                uint32_t rank = (uint32_t)(bc >> 32), k = (uint32_t)bc;
                if (rank == 0) { n.lastFire = bt; n.firings++; nFires++; emit_fire(v, q, bt, k, q); }
                else if (rank == 1) { nDeliv++; model_neuron_run(v, n, cv, cnt, rs, q, bt, (1u << 30) | q, k, NC_SENT | (1u << 29) | cv.J(bsrc), nFires, nRuns, nVisits); }
                else model_neuron_run(v, n, cv, cnt, rs, q, bt, (2u << 30) | q, 0u, NC_SENT | (2u << 29), nFires, nRuns, nVisits);
                curT = bt; curC = bc; first = false;
            }
        }
        if (s.sweep & NC_SWEEP_END) model_neuron_run(v, n, cv, cnt, rs, q, s.t1, (3u << 30) | q, 0u, NC_SENT | (3u << 29), nFires, nRuns, nVisits);
        v.potAct[row] = make_float2(n.pot, n.act); v.lastRan[row] = n.lastRan; v.lastFire[row] = n.lastFire; v.firings[row] = n.firings;
    }
    v.stats[0] += nFires; v.stats[1] += nDeliv; v.stats[6] += nRuns; v.stats[7] += nVisits;
}
// serial restatement of k_index_build / k_synapse_pass / k_index_reset / k_finish_step
static void model_pass2(nc_engine* e, const StepArgs& s) {
    View& v = e->v;
    if (s.world == 1) {
        // records were appended in row order here; on the device the order is arbitrary — shuffle to make sure nothing depends on it
        uint32_t count = v.localHdr[0];
        for (uint32_t i = 0; i + 1 < count; i += 2) std::swap(v.localRecs[i], v.localRecs[i + 1]);
    }
    for (uint32_t b = 0; b < s.world; b++) {
        uint32_t count = std::min(reinterpret_cast<const uint32_t*>(v.gRecs + (uint64_t)b * s.gStride)[0], s.gStride - 1u);
        e->counts[b] = count;
        for (uint32_t i = 0; i < count; i++) {
            uint32_t idx = b * s.gStride + 1u + i, nrn = v.gRecs[idx].neuron;
            v.next[idx] = v.head[nrn]; v.head[nrn] = (int32_t)idx; v.mask[nrn >> 5] |= 1u << (nrn & 31u);
        }
    }
    e->stride = s.gStride;
    uint32_t cnt[5] = {0, 0, 0, 0, 0};
    for (uint64_t row = 0; row < v.nRows; row++) {
        const uint32_t q = (uint32_t)(v.row0 + row);
        const uint64_t rs = v.rowptr[row], re = v.rowptr[row + 1];
        const bool qFired = (v.mask[q >> 5] >> (q & 31u)) & 1u;
        const float lfS = v.lfStart[row];
        for (uint64_t j = rs; j < re; j++) {
            const SynRec r0 = v.rec[j];
            uint32_t p = r0.pre & 0x7fffffffu, ab = as_u32(v.ad[j].x);
            bool pFired = (v.mask[p >> 5] >> (p & 31u)) & 1u;
            float a = as_f32(ab);
            bool eventful = qFired || pFired || (ab & NC_SENT) || (ab != 0u && a > s.t0 && a <= s.t1);
            if (eventful) resolve_slot(v, s, j, (uint32_t)(j - rs), q, r0, ab, pFired, qFired, lfS, cnt);
        }
    }
    v.stats[2] += cnt[0]; v.stats[3] += cnt[1]; v.stats[4] += cnt[2]; v.stats[5] += cnt[3];
    for (uint32_t b = 0; b < s.world; b++)
        for (uint32_t i = 0; i < e->counts[b]; i++) { uint32_t nrn = v.gRecs[b * s.gStride + 1u + i].neuron; v.head[nrn] = -1; v.mask[nrn >> 5] = 0u; }
    v.localHdr[0] = 0; v.localHdr[1] = 0;
}
static void fill(nc_engine* e, StepArgs& a, float t0, float t1, int sweep, const nc_event* ev, uint32_t nEv) {
    memset(&a, 0, sizeof(a));
    a.t0 = t0; a.t1 = t1; a.sweep = sweep; a.lr = e->lr; a.preFactor = e->preF; a.postFactor = e->postF; a.preDecay = e->preD; a.postDecay = e->postD;
    a.ev = ev; a.nEv = nEv; a.candCap = e->candCap; a.gStride = e->v.fireCap + 1u; a.world = (uint32_t)e->world;
}
static int finish(nc_engine* e, uint64_t* hidden, nc_step_stats* st) {
    unsigned long long t[8];
    memcpy(t, e->stats, sizeof(t));
    if (e->world > 1) {  // network-wide counters: all-gather the shards' blocks and add them up
        std::vector<unsigned long long> all((size_t)e->world * 8);
        if (!e->xchgFn || e->xchgFn(e->xchgCtx, e->stats, all.data(), sizeof(t))) return fail(e, NC_ERR_STATE, "exchange failed");
        memset(t, 0, sizeof(t));
        for (int b = 0; b < e->world; b++) for (int i = 0; i < 8; i++) t[i] += all[(size_t)b * 8 + i];
    }
    if (hidden) *hidden = t[5];
    if (st) { st->fires = t[0]; st->deliveries = t[1]; st->loads_accepted = t[2]; st->loads_dropped = t[3];
              st->plasticity_calls = t[4]; st->hidden_rand_calls = t[5]; st->neuron_runs = t[6]; st->active_visits = t[7]; }
    memset(e->stats, 0, sizeof(e->stats));
    return NC_OK;
}
// the live fire exchange of engine.cu (exchange_fires): all-gather the first xchgUnits units, look at the headers, grow and repeat
static int exchange_fires(nc_engine* e, StepArgs& a) {
    if (!e->xchgFn) return fail(e, NC_ERR_STATE, "exchange: world > 1 needs nc_set_exchange");
    for (;;) {
        if (e->xchgFn(e->xchgCtx, e->units.data(), e->gather.data(), (uint64_t)e->xchgUnits * sizeof(FireRec))) return fail(e, NC_ERR_STATE, "exchange failed");
        uint32_t mx = 0;
        for (int b = 0; b < e->world; b++) {
            const uint32_t* hdr = reinterpret_cast<const uint32_t*>(e->gather.data() + (size_t)b * e->xchgUnits);
            if (hdr[1]) return fail(e, NC_ERR_CAPACITY, "fire capacity (gathered header)");
            mx = std::max(mx, hdr[0]);
        }
        if (mx + 1u <= e->xchgUnits) { a.gStride = e->xchgUnits; return NC_OK; }
        uint64_t want = 1;
        while (want < 2ull * mx) want <<= 1;
        e->xchgUnits = (uint32_t)std::min<uint64_t>(want + 1, (uint64_t)e->v.fireCap + 1);
    }
}
int nc_set_exchange(nc_engine* e, nc_allgather_fn fn, void* ctx) { e->xchgFn = fn; e->xchgCtx = ctx; return NC_OK; }
int nc_comm_unique_id(nc_comm_id*) { g_err = "the CPU test double has no NCCL"; return NC_ERR_NO_DEVICE; }
int nc_comm_init(nc_engine* e, const nc_comm_id*) { return fail(e, NC_ERR_NO_DEVICE, "the CPU test double has no NCCL"); }
int nc_rand_set_state(nc_engine* e, const uint32_t* x31) { memcpy(e->rstate, x31, 31 * 4); e->randOn = true; return NC_OK; }
int nc_rand_get_state(nc_engine* e, uint32_t* x31) { if (!e->randOn) return fail(e, NC_ERR_STATE, "no stream state"); memcpy(x31, e->rstate, 31 * 4); return NC_OK; }
// NeuCor::run's background-firing loop (NeuCor.cpp:604-607), serially, on the model's copy of the stream
int nc_background_draw(nc_engine* e, float t0, float runSpeed, uint32_t period, uint64_t nNeurons) {
    if (!e->randOn) return fail(e, NC_ERR_STATE, "nc_background_draw: set the stream state first");
    e->bg.clear();
    for (uint64_t i = 0; i < nNeurons; i++) {
        if ((e->rs_next() >> 1) % period != 0u) continue;
        uint32_t n = (uint32_t)((uint64_t)(e->rs_next() >> 1) % nNeurons);
        float u = (float)(int)(e->rs_next() >> 1) / 2147483648.0f;
        float time = t0 + u * runSpeed;
        if (n >= e->v.row0 && n < e->v.row0 + e->v.nRows) e->bg.push_back(nc_event{n, time, 2u, 0u});
    }
    std::stable_sort(e->bg.begin(), e->bg.end(), [](const nc_event& a, const nc_event& b) { return a.neuron < b.neuron; });
    for (size_t i = 0; i < e->bg.size(); i++)
        if (i + 1 == e->bg.size() || e->bg[i + 1].neuron != e->bg[i].neuron) e->bg[i].index_or_flags = 1u;
    e->bgActive = true; e->bgFirst = true;
    return NC_OK;
}
int nc_background_clear(nc_engine* e) { e->bgActive = false; return NC_OK; }
int nc_background_read(nc_engine* e, uint32_t capacity, nc_event* out, uint32_t* count, uint32_t* hits) {
    if (count) *count = (uint32_t)e->bg.size();
    if (hits) *hits = 0;
    for (uint32_t i = 0; out && i < capacity && i < e->bg.size(); i++) out[i] = e->bg[i];
    return NC_OK;
}
int nc_step(nc_engine* e, float t0, float t1, int sweep, const nc_event* ev, uint32_t nEv, uint64_t* hidden, nc_step_stats* st) {
    if (!e->uploaded) return fail(e, NC_ERR_STATE, "step: no network uploaded");
    if (!(t1 > t0)) return fail(e, NC_ERR_INVALID, "step: window must have t1 > t0");
    if (!(t1 - t0 < e->minDelay) || !(t1 - t0 < 2.0f)) return fail(e, NC_ERR_INVALID, "step: window too long");
    std::vector<nc_event> merged;
    if (e->bgActive) {  // the window's share of the run's background events, after the host's events of the same neuron
        merged.assign(ev, ev + nEv);
        for (auto& b : e->bg)
            if ((e->bgFirst ? b.time >= t0 : b.time > t0) && b.time <= t1) merged.push_back(b);
        std::stable_sort(merged.begin(), merged.end(), [](const nc_event& a, const nc_event& b) { return a.neuron < b.neuron; });
        e->bgFirst = false;
        ev = merged.data(); nEv = (uint32_t)merged.size();
    }
    StepArgs a; fill(e, a, t0, t1, sweep, ev, nEv);
    model_pass1(e, a);
    if (e->v.localHdr[1]) return fail(e, NC_ERR_CAPACITY, "fire capacity");
    if (e->world > 1) { int rc = exchange_fires(e, a); if (rc) return rc; }
    model_pass2(e, a);
    uint64_t h = 0;
    int rc = finish(e, &h, st);
    if (hidden) *hidden = h;
    if (rc == NC_OK && e->randOn) for (uint64_t k = 0; k < h; k++) (void)e->rs_next();  // the hidden rand() calls of the window (NeuCor.cpp:752)
    return rc;
}
// the two-halves form of nc_step: the test double does the work in the first half and hands the result over in the second
static uint64_t g_pendingHidden; static nc_step_stats g_pendingStats; static int g_pendingRc = NC_ERR_STATE;
int nc_step_launch(nc_engine* e, float t0, float t1, int sweep, const nc_event* ev, uint32_t nEv) {
    g_pendingRc = nc_step(e, t0, t1, sweep, ev, nEv, &g_pendingHidden, &g_pendingStats);
    return g_pendingRc;
}
int nc_step_collect(nc_engine* e, uint64_t* hidden, nc_step_stats* st) {
    if (g_pendingRc != NC_OK) return fail(e, NC_ERR_STATE, "nc_step_collect: no window in flight");
    if (hidden) *hidden = g_pendingHidden;
    if (st) *st = g_pendingStats;
    g_pendingRc = NC_ERR_STATE;
    return NC_OK;
}
int nc_run_neurons(nc_engine* e, float now, const uint32_t* ids, uint32_t n, uint64_t* hidden, nc_step_stats* st) {
    if (ids && !n) return NC_OK;
    StepArgs a; fill(e, a, now, now, NC_SWEEP_END, nullptr, 0);
    a.subset = ids; a.nSubset = n;
    if (e->world > 1) return fail(e, NC_ERR_STATE, "nc_run_neurons: single-shard engines only");
    model_pass1(e, a);
    model_pass2(e, a);
    return finish(e, hidden, st);
}
int nc_read_neurons(nc_engine* e, float* potAct, float* lastFire, float* lastRan) {
    uint64_t N = e->v.nRows;
    if (potAct) memcpy(potAct, e->potAct.data(), N * 8);
    if (lastFire) memcpy(lastFire, e->lastFire.data(), N * 4);
    if (lastRan) memcpy(lastRan, e->lastRan.data(), N * 4);
    return NC_OK;
}
int nc_read_synapses(nc_engine* e, float* w, float* a, float* d, float* la, float* ls) {
    uint64_t b = e->v.S * 4;
    for (uint64_t j = 0; j < e->v.S; j++) { if (w) w[j] = e->rec[j].weight; if (la) la[j] = e->rec[j].lastArr; }
    for (uint64_t j = 0; j < e->v.S; j++) { if (a) a[j] = e->ad[j].x; if (d) d[j] = e->ad[j].y; }
    if (ls) memcpy(ls, e->lastStart.data(), b);
    return NC_OK;
}
int nc_read_neuron_counters(nc_engine* e, float* actStart, uint32_t* firings) {
    if (actStart) memcpy(actStart, e->actStart.data(), e->v.nRows * 4);
    if (firings) memcpy(firings, e->firings.data(), e->v.nRows * 4);
    return NC_OK;
}
int nc_read_network(nc_engine* e, uint64_t* rowptr, uint32_t* pre, float* length, uint8_t* inh) {
    if (rowptr) memcpy(rowptr, e->rowptr.data(), (e->v.nRows + 1) * 8);
    for (uint64_t j = 0; j < e->v.S; j++) {
        if (pre) pre[j] = e->rec[j].pre & 0x7fffffffu;
        if (length) length[j] = e->rec[j].delay * 0.5f;
        if (inh) inh[j] = (uint8_t)(e->rec[j].pre >> 31);
    }
    return NC_OK;
}
int nc_write_neurons(nc_engine* e, const float* potAct, const float* lastFire, const float* lastRan, const float* actStart, const uint32_t* firings) {
    uint64_t n = e->v.nRows;
    if (potAct) memcpy(e->potAct.data(), potAct, n * 8);
    if (lastFire) memcpy(e->lastFire.data(), lastFire, n * 4);
    if (lastRan) memcpy(e->lastRan.data(), lastRan, n * 4);
    if (actStart) memcpy(e->actStart.data(), actStart, n * 4);
    if (firings) memcpy(e->firings.data(), firings, n * 4);
    return NC_OK;
}
int nc_write_synapses(nc_engine* e, const float* w, const float* a, const float* d, const float* la, const float* ls) {
    for (uint64_t j = 0; j < e->v.S; j++) {
        if (w) e->rec[j].weight = w[j];
        if (a) e->ad[j].x = a[j];
        if (d) e->ad[j].y = d[j];
        if (la) e->rec[j].lastArr = la[j];
        if (ls) e->lastStart[j] = ls[j];
    }
    return NC_OK;
}
int nc_read_fires(nc_engine* e, uint32_t cap, uint32_t* neuron, float* time, uint32_t* count) {
    uint32_t total = 0;
    for (int b = 0; b < e->world; b++)
        for (uint32_t i = 0; i < e->counts[b]; i++, total++)
            if (total < cap) { const FireRec& r = e->v.gRecs[(size_t)b * e->stride + 1u + i]; if (neuron) neuron[total] = r.neuron; if (time) time[total] = r.time; }
    *count = total;
    return NC_OK;
}
// Synapse::getPrePot / getPostPot (NeuCor.cpp:547-567), the same float/double typing as k_synapse_pots in csrc/engine.cu
static float model_render_behaviour(float valf) {  // AP_RENDER_BEHAVIOUR, NeuCor.cpp:547-550
    float val = (float)fmin(fmax((double)valf, 0.0), 0.7);
    if ((double)val < 0.5) {
        float x = (float)div64((double)val, 5.0);
        float p = (x >= 1.17549435e-38f) ? powf_pos(x, 3.0f) : 0.0f;
        return (float)mul64(mul64(8.0, 1000.0), (double)p);
    }
    float x = (float)sub64(3.5, (double)mul32(5.0f, val));
    return (float)mul64(8.0, (double)mul32(x, x));
}
int nc_read_synapse_pots(nc_engine* e, float now, float* pre, float* post) {
    for (uint64_t i = 0; i < e->v.S; i++) {
        const float a = e->ad[i].x, w = e->rec[i].weight, dl = e->rec[i].delay;
        float p = 0.0f, q = 0.0f;
        if (a != 0.0f) {
            p = mul32(model_render_behaviour(div32(sub32(now, e->lastStart[i]), dl)), w);
            if (now < a) q = mul32(model_render_behaviour(div32(sub32(a, now), dl)), w);
        }
        if (pre) pre[i] = p;
        if (post) post[i] = q;
    }
    return NC_OK;
}
int nc_synapse_pots_device(nc_engine*, float, float*, float*) { return NC_OK; }
static int model_histogram(nc_engine* e, int which, uint32_t spans, float rmin, float rmax, uint32_t* bins, uint32_t* below, uint32_t* above) {
    for (uint32_t i = 0; i < spans; i++) bins[i] = 0;
    uint32_t lo = 0, hi = 0;
    const float range = rmax - rmin;
    const uint64_t n = which ? e->v.S : e->v.nRows;
    for (uint64_t i = 0; 0.0f < range && i < n; i++) {
        const float x = which ? e->rec[i].weight : e->potAct[i].y;
        const float f = floorf(((float)(int)spans * (x - rmin)) / range);
        if (!(f >= 0.0f)) lo++; else if (f >= (float)spans) hi++; else bins[(uint32_t)f]++;
    }
    if (below) *below = lo;
    if (above) *above = hi;
    return NC_OK;
}
int nc_render_activity_histogram(nc_engine* e, uint32_t spans, float a, float b, uint32_t* bins, uint32_t* lo, uint32_t* hi) { return model_histogram(e, 0, spans, a, b, bins, lo, hi); }
int nc_render_weight_histogram(nc_engine* e, uint32_t spans, float a, float b, uint32_t* bins, uint32_t* lo, uint32_t* hi) { return model_histogram(e, 1, spans, a, b, bins, lo, hi); }
int nc_render_raster(nc_engine* e, float now, float runSpeed, uint32_t cap, uint32_t* ids, uint32_t* count) {
    uint32_t c = 0;
    for (uint64_t i = 0; i < e->v.nRows; i++)
        if (now - e->lastFire[i] < runSpeed) { if (c < cap && ids) ids[c] = (uint32_t)(e->v.row0 + i); c++; }
    *count = c;
    return NC_OK;
}
int nc_state_signature(nc_engine* e, uint64_t* out) {
    const unsigned long long C = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < 6; i++) out[i] = 0;
    for (uint64_t i = 0; i < e->v.nRows; i++) {
        out[0] += (unsigned long long)as_u32(e->potAct[i].x) * ((i + 1) * C);
        out[1] += (unsigned long long)as_u32(e->potAct[i].y) * ((i + 1) * C);
        out[2] += (unsigned long long)as_u32(e->lastFire[i]) * ((i + 1) * C);
    }
    for (uint64_t j = 0; j < e->v.S; j++) {
        out[3] += (unsigned long long)as_u32(e->rec[j].weight) * ((j + 1) * C);
        out[4] += (unsigned long long)as_u32(e->ad[j].x) * ((j + 1) * C);
        if (e->ad[j].x != 0.0f) out[4] += (unsigned long long)as_u32(e->ad[j].y) * ((j + 1) * C);
        out[5] += (unsigned long long)as_u32(e->rec[j].lastArr) * ((j + 1) * C);
    }
    return NC_OK;
}
int nc_reset_activities(nc_engine* e, float now) {
    for (uint64_t i = 0; i < e->v.nRows; i++) { e->firings[i] = 0; e->actStart[i] = now; e->potAct[i].y = 0.0f; }
    return NC_OK;
}
int nc_detector_mean(nc_engine* e, const uint32_t* near, uint32_t n, float* out) {
    float avg = 0.0f;
    for (uint32_t i = 0; i < n; i++) avg = avg + e->potAct[near[i] - e->v.row0].x;
    *out = avg / (float)n;
    return NC_OK;
}
}  // extern "C"
