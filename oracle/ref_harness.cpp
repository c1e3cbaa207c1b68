// TEST INFRASTRUCTURE ONLY — headless C-ABI harness around the UNMODIFIED reference core.
//
// This translation unit is compiled together with /root/reference/src/NeuCor.cpp (taken from
// where it lies, never copied into this repo) by oracle/build_ref.sh into
// oracle/_ref/libneucor_ref.so (and, with the tie-canonicalising comparator patch generated at
// build time, oracle/_ref/libneucor_ref_canon.so).  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load those libraries.
//
// Access to the reference's protected/private state is legal C++: NeuCor and Synapse both declare
// `friend class NeuCor_Renderer` (NeuCor.h:98, NeuCor.h:270).  The real renderer needs OpenGL and
// cannot be compiled here, so this TU supplies its own class of that name and uses the friendship
// to read `neurons`, `potAct`, `inputHandler`, and Synapse::{pN,tN,weight,length,inhibitory,
// AP_fireTime,AP_depolFac,lastSpikeArrival,lastSpikeStart}.
//
// Driver ("sweep mode"): run() followed by getDetectorVoltage(0) on a detector
// whose radius covers every neuron, which runs all neurons at the step's end time in ascending ID
// (NeuCor.cpp:359-366) — the same pattern the GUI uses (main.cpp:151, NeuCor_Renderer.cpp:1274).
#include "NeuCor.h"

#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

// Deterministic heap for the reference: every C++ allocation made inside this library is zero-filled.
// The reference reads two members it never initialises — Neuron::scheduledFireTime (NeuCor.h:241, compared at
// NeuCor.cpp:683) and Synapse::inhibitory after a copy (NeuCor.cpp:487-524, read at :760) — so on a recycled
// heap its results depend on whatever bytes the allocator hands back (e.g. a stale 1.0f makes a neuron fire
// spuriously at t = 1 ms).  Zero-filled memory is what the reference sees on a pristine heap; it makes
// oracle/_ref reproducible across processes without touching the reference's source.  (-Wl,-Bsymbolic keeps
// these definitions private to this library.)
void* operator new(std::size_t n) {
    void* p = calloc(1, n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void* operator new[](std::size_t n) { return operator new(n); }
void operator delete(void* p) noexcept { free(p); }
void operator delete[](void* p) noexcept { free(p); }
void operator delete(void* p, std::size_t) noexcept { free(p); }
void operator delete[](void* p, std::size_t) noexcept { free(p); }

class NeuCor_Renderer {
public:
    static std::deque<Neuron>& neurons(NeuCor* b) { return b->neurons; }
    static std::vector<float>& potAct(NeuCor* b) { return b->potAct; }
    static std::vector<coord3>& positions(NeuCor* b) { return b->positions; }
    static std::vector<InputFirer>& inputs(NeuCor* b) { return b->inputHandler; }
    static std::vector<VoltageDetector>& detectors(NeuCor* b) { return b->voltageDetectors; }
    static Synapse* syn(NeuCor* b, std::size_t from, std::size_t to) { return b->getSynapse(from, to); }
    static void resetActivities(NeuCor* b) { b->resetActivities(); }

    static std::size_t pN(const Synapse& s) { return s.pN; }
    static std::size_t tN(const Synapse& s) { return s.tN; }
    static float weight(const Synapse& s) { return s.weight; }
    static float length(const Synapse& s) { return s.length; }
    static unsigned char flagByte(const Synapse& s) {
        unsigned char b;
        std::memcpy(&b, &s.inhibitory, 1);  // read the (possibly uninitialised) flag as a raw byte
        return b;
    }
    static void setFlag(Synapse& s, bool v) { s.inhibitory = v; }
    static float arrive(const Synapse& s) { return s.AP_fireTime; }
    static float depol(const Synapse& s) { return s.AP_depolFac; }
    static float lastArr(const Synapse& s) { return s.lastSpikeArrival; }
    static float lastStart(const Synapse& s) { return s.lastSpikeStart; }
    static float prePot(const Synapse& s) { return s.getPrePot(); }
    static float postPot(const Synapse& s) { return s.getPostPot(); }
};
typedef NeuCor_Renderer R;

struct RefHandle {
    NeuCor* brain;
    std::vector<float> rates;   // the caller-owned array NeuCor keeps a pointer to (NeuCor.cpp:47)
    bool sweep;
    uint64_t S;
};

static uint64_t fnv(uint64_t h, const void* p, size_t n) {
    const unsigned char* c = (const unsigned char*)p;
    for (size_t i = 0; i < n; i++) { h ^= c[i]; h *= 1099511628211ull; }
    return h;
}

extern "C" {

void ref_srand(unsigned seed) { srand(seed); }
int ref_rand(void) { return rand(); }

void* ref_create(int n_neurons) {
    RefHandle* h = new RefHandle();
    h->brain = new NeuCor(n_neurons);
    h->sweep = false;
    h->S = 0;
    return h;
}
void ref_destroy(void* hv) {
    RefHandle* h = (RefHandle*)hv;
    delete h->brain;
    delete h;
}
void ref_create_neuron(void* hv, float x, float y, float z) {
    coord3 c{x, y, z};
    ((RefHandle*)hv)->brain->createNeuron(c);
}
void ref_create_synapse(void* hv, uint64_t to, uint64_t from, float w) {
    ((RefHandle*)hv)->brain->createSynapse(to, from, w);
}
void ref_make_connections(void* hv) { ((RefHandle*)hv)->brain->makeConnections(); }

// positions/radii may be NULL (random positions, radius 1.0: NeuCor.cpp:53-56)
void ref_set_inputs(void* hv, const float* rates, unsigned n, const float* pos_xyz, const float* radius) {
    RefHandle* h = (RefHandle*)hv;
    h->rates.assign(rates, rates + n);
    if (pos_xyz) {
        std::vector<coord3> p(n);
        std::vector<float> r(radius, radius + n);
        for (unsigned i = 0; i < n; i++) p[i] = coord3{pos_xyz[3 * i], pos_xyz[3 * i + 1], pos_xyz[3 * i + 2]};
        h->brain->setInputRateArray(h->rates.data(), n, p.data(), r.data());
    } else {
        h->brain->setInputRateArray(h->rates.data(), n);
    }
}
void ref_set_rate(void* hv, unsigned i, float v) { ((RefHandle*)hv)->rates.at(i) = v; }
float ref_get_rate(void* hv, unsigned i) { return ((RefHandle*)hv)->rates.at(i); }
void ref_add_input_offset(void* hv, unsigned i, float t) { ((RefHandle*)hv)->brain->addInputOffset(i, t); }
void ref_set_input_enabled(void* hv, unsigned i, int en) { R::inputs(((RefHandle*)hv)->brain).at(i).enabled = en != 0; }

// Detector 0 with a radius covering everything: its `near` list is all neurons in ascending ID.
void ref_enable_sweep(void* hv) {
    RefHandle* h = (RefHandle*)hv;
    coord3 c{0, 0, 0};
    float r = 1e9f;
    h->brain->setDetectors(1, &c, &r);
    h->sweep = true;
}
void ref_set_params(void* hv, float runSpeed, float learningRate, int runAll) {
    RefHandle* h = (RefHandle*)hv;
    h->brain->runSpeed = runSpeed;
    h->brain->learningRate = learningRate;
    h->brain->runAll = runAll != 0;
}
void ref_set_factors(void* hv, float pre, float post) {
    ((RefHandle*)hv)->brain->presynapticFactor = pre;
    ((RefHandle*)hv)->brain->postsynapticFactor = post;
}
float ref_time(void* hv) { return ((RefHandle*)hv)->brain->getTime(); }

// "Normalised flags" (SURVEY S5): public API only — setWeight(getWeight()) re-derives the flag.
void ref_normalise_flags(void* hv) {
    NeuCor* b = ((RefHandle*)hv)->brain;
    for (auto& n : R::neurons(b))
        for (auto& s : n.outSynapses) s.setWeight(s.getWeight());
}

// One step: run() [+ full sweep through detector 0]. Returns the detector voltage (or 0).
float ref_step(void* hv) {
    RefHandle* h = (RefHandle*)hv;
    h->brain->run();
    if (h->sweep) return h->brain->getDetectorVoltage(0);
    return 0.0f;
}
// Timed stepping loop (cpu_baseline): seconds of wall-clock for `steps` steps.
double ref_run_timed(void* hv, int steps) {
    RefHandle* h = (RefHandle*)hv;
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < steps; i++) {
        h->brain->run();
        if (h->sweep) h->brain->getDetectorVoltage(0);
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
float ref_detector_voltage(void* hv, unsigned id) { return ((RefHandle*)hv)->brain->getDetectorVoltage(id); }
void ref_add_detector(void* hv, float x, float y, float z, float radius) {
    coord3 c{x, y, z};
    ((RefHandle*)hv)->brain->setDetectors(1, &c, &radius);
}
void ref_reset_activities(void* hv) { R::resetActivities(((RefHandle*)hv)->brain); }

void ref_counts(void* hv, uint64_t* N, uint64_t* S) {
    RefHandle* h = (RefHandle*)hv;
    NeuCor* b = h->brain;
    uint64_t s = 0;
    for (auto& n : R::neurons(b)) s += n.inSynapses.size();
    h->S = s;
    *N = R::neurons(b).size();
    *S = s;
}

// Post-sorted CSR export: rows by target ID, in-row ascending presynaptic ID — the iteration order
// of Neuron::inSynapses (std::map, NeuCor.h:212) that charge_insynapses uses (NeuCor.cpp:690).
void ref_export_network(void* hv, uint64_t* rowptr, uint32_t* pre, float* weight, float* length, unsigned char* flag) {
    NeuCor* b = ((RefHandle*)hv)->brain;
    uint64_t k = 0, q = 0;
    for (auto& n : R::neurons(b)) {
        rowptr[q++] = k;
        for (auto& e : n.inSynapses) {
            Synapse* s = R::syn(b, e.first, e.second);
            pre[k] = (uint32_t)e.first;
            weight[k] = R::weight(*s);
            length[k] = R::length(*s);
            flag[k] = R::flagByte(*s);
            k++;
        }
    }
    rowptr[q] = k;
}
void ref_export_positions(void* hv, float* xyz) {
    NeuCor* b = ((RefHandle*)hv)->brain;
    auto& p = R::positions(b);
    for (size_t i = 0; i < p.size(); i++) { xyz[3 * i] = p[i].x; xyz[3 * i + 1] = p[i].y; xyz[3 * i + 2] = p[i].z; }
}
unsigned ref_input_count(void* hv) { return (unsigned)R::inputs(((RefHandle*)hv)->brain).size(); }
uint64_t ref_input_near_count(void* hv, unsigned i) { return R::inputs(((RefHandle*)hv)->brain).at(i).near.size(); }
void ref_export_input(void* hv, unsigned i, uint32_t* near, float* lastFire, float* pos_xyz, float* radius) {
    InputFirer& f = R::inputs(((RefHandle*)hv)->brain).at(i);
    for (size_t k = 0; k < f.near.size(); k++) near[k] = (uint32_t)f.near[k];
    *lastFire = f.lastFire;
    pos_xyz[0] = f.a.x; pos_xyz[1] = f.a.y; pos_xyz[2] = f.a.z;
    *radius = f.radius;
}

// Neuron state, ID order. potAct is the renderer-visible interleaved pair array (NeuCor.h:105).
void ref_read_neurons(void* hv, float* pot, float* act, float* lastFire, float* lastRan) {
    NeuCor* b = ((RefHandle*)hv)->brain;
    auto& pa = R::potAct(b);
    size_t i = 0;
    for (auto& n : R::neurons(b)) {
        if (pot) pot[i] = pa[2 * i];
        if (act) act[i] = pa[2 * i + 1];
        if (lastFire) lastFire[i] = n.lastFire;
        if (lastRan) lastRan[i] = n.lastRan;
        i++;
    }
}
// Synapse state in the CSR order of ref_export_network.
void ref_read_synapses(void* hv, float* weight, float* arrive, float* depol, float* lastArr, float* lastStart) {
    NeuCor* b = ((RefHandle*)hv)->brain;
    uint64_t k = 0;
    for (auto& n : R::neurons(b))
        for (auto& e : n.inSynapses) {
            Synapse* s = R::syn(b, e.first, e.second);
            if (weight) weight[k] = R::weight(*s);
            if (arrive) arrive[k] = R::arrive(*s);
            if (depol) depol[k] = R::depol(*s);
            if (lastArr) lastArr[k] = R::lastArr(*s);
            if (lastStart) lastStart[k] = R::lastStart(*s);
            k++;
        }
}
// Renderer-facing per-synapse values (NeuCor.cpp:551-567), CSR order.
void ref_read_synapse_pots(void* hv, float* prePot, float* postPot) {
    NeuCor* b = ((RefHandle*)hv)->brain;
    uint64_t k = 0;
    for (auto& n : R::neurons(b))
        for (auto& e : n.inSynapses) {
            Synapse* s = R::syn(b, e.first, e.second);
            // lastSpikeStart is uninitialised until the first fire; both getters guard on AP_fireTime != 0
            prePot[k] = R::prePot(*s);
            postPot[k] = R::postPot(*s);
            k++;
        }
}

// FNV-1a hashes of the state arrays: out[0]=pot, [1]=act, [2]=lastFire, [3]=weight, [4]=arrive, [5]=lastArr.
// `arrive` is hashed together with depol only while the slot is busy (depol of an idle slot is stale).
void ref_state_hash(void* hv, uint64_t* out) {
    NeuCor* b = ((RefHandle*)hv)->brain;
    auto& pa = R::potAct(b);
    uint64_t h0 = 1469598103934665603ull, h1 = h0, h2 = h0, h3 = h0, h4 = h0, h5 = h0;
    size_t i = 0;
    for (auto& n : R::neurons(b)) {
        h0 = fnv(h0, &pa[2 * i], 4);
        h1 = fnv(h1, &pa[2 * i + 1], 4);
        float lf = n.lastFire;
        if (lf != lf) lf = -1.0f;  // canonical NaN
        h2 = fnv(h2, &lf, 4);
        i++;
    }
    for (auto& n : R::neurons(b))
        for (auto& e : n.inSynapses) {
            Synapse* s = R::syn(b, e.first, e.second);
            float w = R::weight(*s), a = R::arrive(*s), la = R::lastArr(*s);
            h3 = fnv(h3, &w, 4);
            h4 = fnv(h4, &a, 4);
            if (a != 0) { float d = R::depol(*s); h4 = fnv(h4, &d, 4); }
            h5 = fnv(h5, &la, 4);
        }
    out[0] = h0; out[1] = h1; out[2] = h2; out[3] = h3; out[4] = h4; out[5] = h5;
}

}  // extern "C"
