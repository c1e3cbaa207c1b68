/* TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's hot path (the oracle "port").
 *
 * A serial, event-driven restatement of NeuCor::run() and everything it dispatches to
 * (/root/reference/src/NeuCor.cpp:583-764, :326-345), on flat arrays instead of the reference's
 * deque<Neuron>/vector<Synapse>/map containers.  It is NOT the product and is never linked into
 * it: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * oracle/libneucor_oracle.so.  It is deliberately a *different decomposition* from the CUDA
 * engine (one global time-ordered queue, processed one event at a time, like the reference),
 * so agreement between the two is evidence and not a tautology.
 *
 * Parity pin: tests/test_oracle_pinned.py checks this file bit-for-bit (every neuron and synapse
 * field, every step) against oracle/_ref (the reference's own NeuCor.cpp compiled unmodified, and
 * its tie-canonicalised build) and against the committed golden fixtures in tests/golden/ that
 * were generated from oracle/_ref by tests/golden/make_golden.py.  The reference ships no tests
 * or golden vectors for this path, so the pin is "outputs of the reference itself
 * run here".
 *
 * Differences from the reference, all of them outside its observable arithmetic:
 *  - equal-time queue events pop in the canonical order (time, rank, a, b) of SURVEY.md App. C
 *    (InputFirer (0,index,0) < Synapse (1,target,parent) < Neuron (2,id,0)) instead of libstdc++
 *    heap order; this matches oracle/_ref/libneucor_ref_canon.so always and the unmodified
 *    reference up to its tie horizon H;
 *  - getSynapse()'s linear search (NeuCor.cpp:244-250) is replaced by direct CSR indexing, which
 *    changes cost (O(K) instead of O(K_in*K_out) per neuron update) but no value;
 *  - Neuron::scheduledFireTime is never initialised by the reference (NeuCor.h:241); here it
 *    starts as NaN ("never equal"), SURVEY.md H10.
 * libm/libc calls (powf, exp, rand) go to the same glibc the reference uses.
 * Compile: gcc -O2 -ffp-contract=off -fno-fast-math (no FMA contraction).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
    float time;
    uint32_t rank; /* 0 input firer, 1 synapse delivery, 2 neuron */
    uint32_t a, b; /* canonical tie keys */
    uint64_t idx;  /* input index / synapse CSR index / neuron id */
} Event;

typedef struct {
    /* network, post-sorted CSR (rows = target neuron, in-row ascending presynaptic ID) */
    uint64_t N, S;
    uint64_t *rowptr;
    uint32_t *pre, *post;
    float *len, *weight, *arrive, *depol, *lastArr, *lastStart;
    uint8_t *inh;
    /* out-adjacency: CSR indices of each neuron's outgoing synapses */
    uint64_t *outptr, *outidx;
    /* neurons */
    float *pot, *act, *lastRan, *lastFire, *sched, *actStart;
    uint32_t *firings;
    /* input firers (NeuCor.h:170-180) */
    uint32_t G;
    uint64_t *nearptr;
    uint32_t *near;
    float *inLastFire;
    uint8_t *inEnabled;
    float *rates;
    /* scheduler */
    Event *heap;
    uint64_t heapN, heapCap;
    float currentTime;
    /* public fields of NeuCor (NeuCor.cpp:18-29) */
    float runSpeed, learningRate, preDecay, postDecay, preFactor, postFactor;
    int runAll;
    /* statistics (not part of the reference) */
    uint64_t nFires, nDeliveries, nLoadsAccepted, nLoadsDropped, nPlast, nHiddenRand, nRuns, nActiveVisits;
    /* optional fire log of the current step: (neuron, time) */
    uint32_t *fireLogN;
    float *fireLogT;
    uint64_t fireLogCount, fireLogCap;
} Oracle;

/* ---- canonical min-heap -------------------------------------------------------------------- */
static int ev_less(const Event *x, const Event *y) {
    if (x->time != y->time) return x->time < y->time;
    if (x->rank != y->rank) return x->rank < y->rank;
    if (x->a != y->a) return x->a < y->a;
    return x->b < y->b;
}
static void heap_push(Oracle *o, Event e) {
    if (o->heapN == o->heapCap) {
        o->heapCap = o->heapCap ? o->heapCap * 2 : 1024;
        o->heap = (Event *)realloc(o->heap, o->heapCap * sizeof(Event));
    }
    uint64_t i = o->heapN++;
    while (i > 0) {
        uint64_t p = (i - 1) / 2;
        if (!ev_less(&e, &o->heap[p])) break;
        o->heap[i] = o->heap[p];
        i = p;
    }
    o->heap[i] = e;
}
static void heap_pop(Oracle *o) {
    Event e = o->heap[--o->heapN];
    uint64_t i = 0, n = o->heapN;
    for (;;) {
        uint64_t c = 2 * i + 1;
        if (c >= n) break;
        if (c + 1 < n && ev_less(&o->heap[c + 1], &o->heap[c])) c++;
        if (!ev_less(&o->heap[c], &e)) break;
        o->heap[i] = o->heap[c];
        i = c;
    }
    if (n) o->heap[i] = e;
}
/* NeuCor::queueSimulation, NeuCor.cpp:227-229: absolute time = currentTime + delay (fp32) */
static void queue_neuron(Oracle *o, uint32_t id, float delay) {
    Event e = {o->currentTime + delay, 2, id, 0, id};
    heap_push(o, e);
}

/* ---- Synapse ------------------------------------------------------------------------------- */
/* Synapse::synapticPlasticity, NeuCor.cpp:740-764 */
static void plasticity(Oracle *o, uint64_t s) {
    float now = o->currentTime;
    float traceS = powf(o->preDecay, now - o->lastArr[s]);
    /* Neuron::getTrace, NeuCor.cpp:671-675 */
    float traceT = powf(o->postDecay, now - o->lastFire[o->post[s]]);
    if (!(traceT == traceT)) traceT = 0;
    if (traceT == 1) traceT = 0;
    if (traceS == 1) traceS = 0;
    o->nPlast++;
    /* NeuCor.cpp:752: `weight == 0 && !inhibitory && rand()%120 == 0 && false` — by short-circuit
       evaluation rand() IS called when the first two terms hold; the branch is never taken. */
    if (o->weight[s] == 0 && !o->inh[s]) {
        (void)rand();
        o->nHiddenRand++;
    }
    float weightChange = o->preFactor * traceS - o->postFactor * traceT;
    float w = o->weight[s];
    w += weightChange * o->learningRate;
    if (!o->inh[s]) w = (float)fmax(fmin(w, 1.0), 0.0);
    else w = (float)fmax(fmin(w, 0.0), -1.0);
    o->weight[s] = w;
}
/* Synapse::fire, NeuCor.cpp:727-738 (one spike in flight per synapse) */
static void synapse_load(Oracle *o, uint64_t s) {
    if (o->arrive[s] != 0) { o->nLoadsDropped++; return; }
    float d = 0.2f;            /* AP_depolFac handed over by the neuron (NeuCor.cpp:384,648) */
    d *= 52.0;                 /* float *= double, rounded to float */
    d *= o->weight[s];
    o->depol[s] = d;
    float a = o->len[s] * 2.0f; /* length * AP_speed (NeuCor.cpp:485,733) */
    Event e = {o->currentTime + a, 1, o->post[s], o->pre[s], s};
    heap_push(o, e);
    a += o->currentTime;
    o->arrive[s] = a;
    o->lastStart[s] = o->currentTime;
    o->nLoadsAccepted++;
}

/* ---- Neuron -------------------------------------------------------------------------------- */
/* Neuron::fire, NeuCor.cpp:643-656 */
static void neuron_fire(Oracle *o, uint32_t q) {
    o->lastFire[q] = o->currentTime;
    o->firings[q]++;
    o->nFires++;
    if (o->fireLogN && o->fireLogCount < o->fireLogCap) {
        o->fireLogN[o->fireLogCount] = q;
        o->fireLogT[o->fireLogCount] = o->currentTime;
    }
    o->fireLogCount++;
    for (uint64_t k = o->outptr[q]; k < o->outptr[q + 1]; k++) synapse_load(o, o->outidx[k]);
    for (uint64_t s = o->rowptr[q]; s < o->rowptr[q + 1]; s++) plasticity(o, s);
}
/* Neuron::run, NeuCor.cpp:619-641 */
static void neuron_run(Oracle *o, uint32_t q) {
    const float baselevel = -70.0f, threshold = -55.0f, recharge = 0.5f;
    const float AP_h = 100.0f, AP_depolW = 0.3f, AP_polW = 0.6f, AP_deltaPol = 1.16f, AP_depolFac = 0.2f,
                AP_deltaStart = 1.0f, AP_cutoff = 2.0f;
    float currentT = o->currentTime;
    float deltaT = currentT - o->lastRan[q];
    o->lastRan[q] = currentT;
    if (deltaT == 0) return;
    o->nRuns++;

    /* charge_insynapses, NeuCor.cpp:688-700 — ascending presynaptic ID (std::map order) */
    float newPot = o->pot[q];
    for (uint64_t s = o->rowptr[q]; s < o->rowptr[q + 1]; s++) {
        float a = o->arrive[s];
        float timeOffset = currentT - a;
        if (timeOffset <= 0.0 || a == 0) continue;
        newPot += deltaT * o->depol[s] * 0.9943 * exp(0.3702 * deltaT);
        o->nActiveVisits++;
        if (AP_cutoff < timeOffset) o->arrive[s] = 0;
    }
    /* charge_passive, NeuCor.cpp:677-680 */
    newPot = (newPot - baselevel) * powf(recharge, deltaT) + baselevel;
    o->pot[q] = newPot;
    /* charge_thresholdCheck, NeuCor.cpp:682-686 (vesicles term is always true) */
    float lf = o->lastFire[q];
    if ((threshold < o->pot[q] || o->sched[q] == currentT) && (lf != lf || AP_cutoff < currentT - lf)) neuron_fire(o, q);
    /* AP, NeuCor.cpp:706-714 — powf(x, 2.0) is x*x in the reference's -O3 build */
    lf = o->lastFire[q];
    if (!(lf != lf || AP_cutoff < currentT - lf)) {
        float x1 = (currentT - lf) - AP_deltaStart;
        float x2 = (currentT - lf) - AP_deltaStart - AP_deltaPol;
        float currentAP = AP_h * (exp(-(x1 * x1) / (2.0 * AP_depolW * AP_depolW)) -
                                  exp(-(x2 * x2) / (2.0 * AP_polW * AP_polW)) * AP_depolFac) +
                          baselevel + (threshold - baselevel) * fmax(1.0 + lf - currentT, 0.0);
        o->pot[q] = currentAP;
    }
    /* activity, NeuCor.cpp:640 */
    o->act[q] = o->firings[q] / ((currentT - o->actStart[q]) / 10.0);
}

/* ---- NeuCor::run, NeuCor.cpp:583-617 ------------------------------------------------------- */
static void run_queue(Oracle *o) {
    if (o->runSpeed <= 0.0f) return;
    if (o->runAll)
        for (uint64_t q = 0; q < o->N; q++) queue_neuron(o, (uint32_t)q, 0.0f);
    /* InputFirer::schedule, NeuCor.cpp:333-345 */
    for (uint32_t g = 0; g < o->G; g++) {
        float frequency = o->rates ? o->rates[g] : 0.0f;
        if (frequency == 0 || !o->inEnabled[g]) continue;
        if (o->inLastFire[g] != o->inLastFire[g]) o->inLastFire[g] = 0;
        float currentT = o->currentTime, deltaT = o->runSpeed;
        for (float fireTime = o->inLastFire[g] + 1000.0 / frequency; fireTime < currentT + deltaT;
             fireTime += 1000.0 / frequency) {
            if (currentT < fireTime) {
                Event e = {o->currentTime + (fireTime - currentT), 0, g, 0, g};
                heap_push(o, e);
                o->inLastFire[g] = fireTime;
            }
        }
    }
    /* background firing, NeuCor.cpp:604-607 (argument evaluation: rand()%N first, then randomUnit()) */
    for (uint64_t i = 0; i < o->N; i++) {
        int period = (int)(600.0f / o->runSpeed);
        if (period < 1) period = 1;
        if (rand() % period == 0) {
            uint32_t n = (uint32_t)((uint64_t)rand() % o->N);
            float t = ((float)rand() / (float)RAND_MAX) * o->runSpeed;
            o->sched[n] = o->currentTime + t; /* Neuron::scheduleFire, NeuCor.cpp:658-661 */
            queue_neuron(o, n, t);
        }
    }
    float targetTime = o->currentTime + o->runSpeed;
    while (o->heapN) {
        Event e = o->heap[0];
        o->currentTime = e.time;
        if (targetTime < o->currentTime) break;
        heap_pop(o);
        if (e.rank == 0) { /* InputFirer::run, NeuCor.cpp:326-331 */
            uint32_t g = (uint32_t)e.idx;
            if (o->inEnabled[g])
                for (uint64_t k = o->nearptr[g]; k < o->nearptr[g + 1]; k++) neuron_fire(o, o->near[k]);
        } else if (e.rank == 1) { /* Synapse::run, NeuCor.cpp:718-726 */
            uint64_t s = e.idx;
            if (o->arrive[s] < o->currentTime) continue;
            uint32_t q = o->post[s];
            neuron_run(o, q);          /* Neuron::transfer, NeuCor.cpp:663-666 */
            queue_neuron(o, q, 2.0f);
            o->lastArr[s] = o->currentTime;
            o->nDeliveries++;
            plasticity(o, s);
        } else {
            neuron_run(o, (uint32_t)e.idx);
        }
    }
    o->currentTime = targetTime;
}

/* ---- C API --------------------------------------------------------------------------------- */
Oracle *orc_create(uint64_t N, uint64_t S, const uint64_t *rowptr, const uint32_t *pre, const float *weight,
                   const float *len, const uint8_t *flag) {
    Oracle *o = (Oracle *)calloc(1, sizeof(Oracle));
    o->N = N; o->S = S;
    o->rowptr = (uint64_t *)malloc((N + 1) * sizeof(uint64_t));
    memcpy(o->rowptr, rowptr, (N + 1) * sizeof(uint64_t));
    uint64_t S1 = S ? S : 1, N1 = N ? N : 1;
    o->pre = (uint32_t *)malloc(S1 * 4); o->post = (uint32_t *)malloc(S1 * 4);
    o->len = (float *)malloc(S1 * 4); o->weight = (float *)malloc(S1 * 4);
    o->arrive = (float *)calloc(S1, 4); o->depol = (float *)calloc(S1, 4);
    o->lastArr = (float *)malloc(S1 * 4); o->lastStart = (float *)calloc(S1, 4);
    o->inh = (uint8_t *)malloc(S1);
    memcpy(o->pre, pre, S * 4); memcpy(o->len, len, S * 4); memcpy(o->weight, weight, S * 4);
    for (uint64_t s = 0; s < S; s++) { o->inh[s] = flag[s] != 0; o->lastArr[s] = -INFINITY; }
    for (uint64_t q = 0; q < N; q++)
        for (uint64_t s = rowptr[q]; s < rowptr[q + 1]; s++) o->post[s] = (uint32_t)q;
    o->outptr = (uint64_t *)calloc(N + 2, sizeof(uint64_t));
    o->outidx = (uint64_t *)malloc(S1 * sizeof(uint64_t));
    for (uint64_t s = 0; s < S; s++) o->outptr[pre[s] + 2]++;
    for (uint64_t q = 0; q < N; q++) o->outptr[q + 2] += o->outptr[q + 1];
    for (uint64_t s = 0; s < S; s++) o->outidx[o->outptr[pre[s] + 1]++] = s;
    o->pot = (float *)malloc(N1 * 4); o->act = (float *)calloc(N1, 4); o->lastRan = (float *)calloc(N1, 4);
    o->lastFire = (float *)malloc(N1 * 4); o->sched = (float *)malloc(N1 * 4); o->actStart = (float *)calloc(N1, 4);
    o->firings = (uint32_t *)calloc(N1, 4);
    for (uint64_t q = 0; q < N; q++) { o->pot[q] = -70.0f; o->lastFire[q] = NAN; o->sched[q] = NAN; }
    o->runSpeed = 1.0f; o->learningRate = 1.0f;
    o->preDecay = 0.75f; o->postDecay = 0.65f; o->preFactor = 0.13f; o->postFactor = 0.30f;
    o->nearptr = (uint64_t *)calloc(1, sizeof(uint64_t));
    return o;
}
void orc_destroy(Oracle *o) {
    free(o->rowptr); free(o->pre); free(o->post); free(o->len); free(o->weight); free(o->arrive); free(o->depol);
    free(o->lastArr); free(o->lastStart); free(o->inh); free(o->outptr); free(o->outidx); free(o->pot); free(o->act);
    free(o->lastRan); free(o->lastFire); free(o->sched); free(o->actStart); free(o->firings); free(o->nearptr);
    free(o->near); free(o->inLastFire); free(o->inEnabled); free(o->rates); free(o->heap); free(o->fireLogN);
    free(o->fireLogT); free(o);
}
/* input firers with precomputed `near` lists (ascending neuron ID, NeuCor.cpp:319-323) */
void orc_set_inputs(Oracle *o, uint32_t G, const uint64_t *nearptr, const uint32_t *near, const float *lastFire,
                    const float *rates) {
    free(o->nearptr); free(o->near); free(o->inLastFire); free(o->inEnabled); free(o->rates);
    o->G = G;
    o->nearptr = (uint64_t *)malloc((G + 1) * sizeof(uint64_t));
    memcpy(o->nearptr, nearptr, (G + 1) * sizeof(uint64_t));
    uint64_t M = nearptr[G];
    o->near = (uint32_t *)malloc((M ? M : 1) * 4);
    memcpy(o->near, near, M * 4);
    o->inLastFire = (float *)malloc((G ? G : 1) * 4);
    o->inEnabled = (uint8_t *)malloc(G ? G : 1);
    o->rates = (float *)malloc((G ? G : 1) * 4);
    for (uint32_t g = 0; g < G; g++) { o->inLastFire[g] = lastFire ? lastFire[g] : 0.0f; o->inEnabled[g] = 1; o->rates[g] = rates[g]; }
}
void orc_set_rate(Oracle *o, uint32_t g, float v) { o->rates[g] = v; }
void orc_set_rates(Oracle *o, const float *v) { memcpy(o->rates, v, o->G * 4); }
void orc_set_input_enabled(Oracle *o, uint32_t g, int en) { o->inEnabled[g] = en != 0; }
void orc_add_input_offset(Oracle *o, uint32_t g, float t) { o->inLastFire[g] += t; }
void orc_set_params(Oracle *o, float runSpeed, float learningRate, int runAll) {
    o->runSpeed = runSpeed; o->learningRate = learningRate; o->runAll = runAll;
}
void orc_set_factors(Oracle *o, float pre, float post) { o->preFactor = pre; o->postFactor = post; }
void orc_set_decays(Oracle *o, float pre, float post) { o->preDecay = pre; o->postDecay = post; }
float orc_time(Oracle *o) { return o->currentTime; }
void orc_enable_fire_log(Oracle *o, uint64_t cap) {
    o->fireLogCap = cap;
    o->fireLogN = (uint32_t *)malloc(cap * 4);
    o->fireLogT = (float *)malloc(cap * 4);
}
uint64_t orc_fire_log(Oracle *o, uint32_t *n, float *t, uint64_t cap) {
    uint64_t c = o->fireLogCount < o->fireLogCap ? o->fireLogCount : o->fireLogCap;
    if (c > cap) c = cap;
    memcpy(n, o->fireLogN, c * 4); memcpy(t, o->fireLogT, c * 4);
    return o->fireLogCount;
}
/* One step. sweep != 0: afterwards run every neuron at the end time in ascending ID and return the
   mean potential — VoltageDetector::getVoltage with a full-radius detector, NeuCor.cpp:359-366. */
float orc_step(Oracle *o, int sweep) {
    o->fireLogCount = 0;
    run_queue(o);
    if (!sweep) return 0.0f;
    float avgV = 0;
    for (uint64_t q = 0; q < o->N; q++) {
        neuron_run(o, (uint32_t)q);
        avgV += o->pot[q];
    }
    return avgV / o->N;
}
double orc_run_timed(Oracle *o, int steps, int sweep) {
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < steps; i++) orc_step(o, sweep);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
/* Neuron::resetActivity for all neurons, NeuCor.cpp:233-235,460 */
void orc_reset_activities(Oracle *o) {
    for (uint64_t q = 0; q < o->N; q++) { o->firings[q] = 0; o->actStart[q] = o->currentTime; o->act[q] = 0.0f; }
}
void orc_read_neurons(Oracle *o, float *pot, float *act, float *lastFire, float *lastRan) {
    if (pot) memcpy(pot, o->pot, o->N * 4);
    if (act) memcpy(act, o->act, o->N * 4);
    if (lastFire) memcpy(lastFire, o->lastFire, o->N * 4);
    if (lastRan) memcpy(lastRan, o->lastRan, o->N * 4);
}
void orc_read_synapses(Oracle *o, float *weight, float *arrive, float *depol, float *lastArr, float *lastStart) {
    if (weight) memcpy(weight, o->weight, o->S * 4);
    if (arrive) memcpy(arrive, o->arrive, o->S * 4);
    if (depol) memcpy(depol, o->depol, o->S * 4);
    if (lastArr) memcpy(lastArr, o->lastArr, o->S * 4);
    if (lastStart) memcpy(lastStart, o->lastStart, o->S * 4);
}
void orc_read_input_lastfire(Oracle *o, float *lf) { memcpy(lf, o->inLastFire, o->G * 4); }
/* out[0..7] = fires, deliveries, loads accepted, loads dropped, plasticity calls, hidden rand() calls,
   effective neuron runs, active in-synapse visits */
void orc_stats(Oracle *o, uint64_t *out) {
    out[0] = o->nFires; out[1] = o->nDeliveries; out[2] = o->nLoadsAccepted; out[3] = o->nLoadsDropped;
    out[4] = o->nPlast; out[5] = o->nHiddenRand; out[6] = o->nRuns; out[7] = o->nActiveVisits;
}
static uint64_t fnv(uint64_t h, const void *p, size_t n) {
    const unsigned char *c = (const unsigned char *)p;
    for (size_t i = 0; i < n; i++) { h ^= c[i]; h *= 1099511628211ull; }
    return h;
}
/* same definition as ref_state_hash in oracle/ref_harness.cpp */
/* The six position-weighted checksums of tests/helpers.state_signature (and of the engine's nc_state_signature):
   sum of bits(x[i]) * (i + 1) * 0x9E3779B97F4A7C15 mod 2^64 for pot, act, lastFire, weight, arrive (+ depol of busy slots), lastArr. */
void orc_state_signature(Oracle *o, uint64_t *out) {
    const uint64_t C = 0x9E3779B97F4A7C15ull;
    uint32_t u;
    for (int i = 0; i < 6; i++) out[i] = 0;
    for (uint64_t q = 0; q < o->N; q++) {
        const uint64_t k = (q + 1) * C;
        memcpy(&u, &o->pot[q], 4); out[0] += (uint64_t)u * k;
        memcpy(&u, &o->act[q], 4); out[1] += (uint64_t)u * k;
        memcpy(&u, &o->lastFire[q], 4); out[2] += (uint64_t)u * k;
    }
    for (uint64_t s = 0; s < o->S; s++) {
        const uint64_t k = (s + 1) * C;
        memcpy(&u, &o->weight[s], 4); out[3] += (uint64_t)u * k;
        memcpy(&u, &o->arrive[s], 4); out[4] += (uint64_t)u * k;
        if (o->arrive[s] != 0) { memcpy(&u, &o->depol[s], 4); out[4] += (uint64_t)u * k; }
        memcpy(&u, &o->lastArr[s], 4); out[5] += (uint64_t)u * k;
    }
}
void orc_state_hash(Oracle *o, uint64_t *out) {
    uint64_t h0 = 1469598103934665603ull, h1 = h0, h2 = h0, h3 = h0, h4 = h0, h5 = h0;
    for (uint64_t q = 0; q < o->N; q++) {
        h0 = fnv(h0, &o->pot[q], 4);
        h1 = fnv(h1, &o->act[q], 4);
        float lf = o->lastFire[q];
        if (lf != lf) lf = -1.0f;
        h2 = fnv(h2, &lf, 4);
    }
    for (uint64_t s = 0; s < o->S; s++) {
        h3 = fnv(h3, &o->weight[s], 4);
        h4 = fnv(h4, &o->arrive[s], 4);
        if (o->arrive[s] != 0) h4 = fnv(h4, &o->depol[s], 4);
        h5 = fnv(h5, &o->lastArr[s], 4);
    }
    out[0] = h0; out[1] = h1; out[2] = h2; out[3] = h3; out[4] = h4; out[5] = h5;
}
