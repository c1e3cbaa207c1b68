// Flat C wrapper around the host-side NeuCor class so that Python (ctypes) can drive it the way
// main.cpp drives the reference's class.  Errors (the C++ exceptions the reference would throw) are
// caught here and reported through nch_last_error(); every function returns 0 on success.
#include <stdlib.h>
#include <string.h>

#include <exception>
#include <string>
#include <vector>

#include "../../include/neucor_b200.h"
#include "NeuCor.h"

namespace {
thread_local std::string g_err;
struct Handle {
    NeuCor* brain;
    std::vector<float> rates;  // the caller-owned array NeuCor keeps a pointer to (NeuCor.cpp:47)
    uint64_t walk = 0x9E3779B97F4A7C15ull;  // private generator of nch_random_walk_rates (xorshift64)
};
template <typename F>
int guard(F f) {
    try { f(); return 0; }
    catch (const std::exception& e) { g_err = e.what(); return -1; }
    catch (...) { g_err = "unknown exception"; return -1; }
}
}  // namespace

extern "C" {

const char* nch_last_error(void) { return g_err.c_str(); }
void nch_srand(unsigned seed) { srand(seed); }
int nch_rand(void) { return rand(); }

void* nch_create(int n_neurons, int device) {
    Handle* h = nullptr;
    if (guard([&] { h = new Handle(); h->brain = new NeuCor(n_neurons); h->brain->deviceOrdinal = device; })) return nullptr;
    return h;
}
void nch_destroy(void* hv) {
    Handle* h = (Handle*)hv;
    if (!h) return;
    delete h->brain;
    delete h;
}
#define B (((Handle*)hv)->brain)
int nch_create_neuron(void* hv, float x, float y, float z) { return guard([&] { B->createNeuron(coord3{x, y, z}); }); }
int nch_create_synapse(void* hv, uint64_t to, uint64_t from, float w) { return guard([&] { B->createSynapse(to, from, w); }); }
int nch_make_connections(void* hv) { return guard([&] { B->makeConnections(); }); }
int nch_import_network(void* hv, uint64_t n, const uint64_t* rowptr, const uint32_t* pre, const float* weight, const float* length,
                       const uint8_t* flag, const float* xyz) {
    return guard([&] { B->importNetwork(n, rowptr, pre, weight, length, flag, xyz); });
}
int nch_import_network_device(void* hv, uint64_t n, uint64_t S, const void* rowptr, const void* pre, const void* weight, const void* length, const void* flag) {
    return guard([&] { B->importNetworkDevice(n, S, (const uint64_t*)rowptr, (const uint32_t*)pre, (const float*)weight, (const float*)length, (const uint8_t*)flag); });
}
int nch_import_shard_device(void* hv, uint64_t n, uint64_t S, const void* rowptr, const void* pre, const void* weight, const void* length, const void* flag, float globalMinDelay) {
    return guard([&] {
        B->importShardDevice(n, S, (const uint64_t*)rowptr, (const uint32_t*)pre, (const float*)weight, (const float*)length, (const uint8_t*)flag);
        B->globalMinDelay = globalMinDelay;
    });
}
// positions of a network that was imported without them (device-resident imports): what the firers' / detectors' `near` lists are built from
int nch_set_positions(void* hv, const float* xyz, uint64_t n) { return guard([&] { B->setPositions(xyz, n); }); }
// One frame of main.cpp's input random walk (main.cpp:100-105) over the handle's rate array: rate += (u - 0.5) * 2 with u uniform
// in [0, 1], clamped to [0, max]; with `paired`, every odd input shares the rate of the even one before it (inputs[1] = inputs[0]:
// the "correlated" groups).  use_libc != 0: u = randomUnit() from libc's rand(), exactly as the reference's driver draws it (the
// brain then sees that the application moved the stream, as with main.cpp); 0: a private generator, libc's stream untouched.
int nch_random_walk_rates(void* hv, float max_rate, int paired, int use_libc, uint64_t* n_rand) {
    return guard([&] {
        Handle* h = (Handle*)hv;
        std::vector<float>& r = h->rates;
        uint64_t draws = 0;
        for (float& input : r) {
            float u;
            if (use_libc) { u = static_cast<float>(rand()) / static_cast<float>(RAND_MAX); draws++; }
            else { h->walk ^= h->walk << 13; h->walk ^= h->walk >> 7; h->walk ^= h->walk << 17; u = (float)(h->walk >> 40) / 16777216.0f; }
            input += (u - 0.5f) * 2.0f;
            input = input < 0.0f ? 0.0f : input > max_rate ? max_rate : input;
        }
        if (paired)
            for (size_t i = 1; i < r.size(); i += 2) r[i] = r[i - 1];
        if (n_rand) *n_rand = draws;
    });
}
int nch_set_shard(void* hv, int rank, int world) { return guard([&] { B->setShard(rank, world); }); }
int nch_set_comm_id(void* hv, const void* id128) { return guard([&] { B->setCommId(id128); }); }
int nch_set_exchange(void* hv, int (*fn)(void*, const void*, void*, uint64_t), void* ctx) { return guard([&] { B->setExchange(fn, ctx); }); }
void nch_shard_info(void* hv, uint64_t* row0, uint64_t* rows, uint64_t* synapses) { *row0 = B->shardRow0(); *rows = B->shardRows(); *synapses = B->shardSynapses(); }
// (off: the caller wants no per-step state on the host at all — neither the swept mean nor the potAct / lastFire mirrors)
int nch_set_sweep_mean(void* hv, int on) { return guard([&] { B->sweepReturnsMean = on != 0; B->mirrorAfterRun = on != 0; }); }
int nch_set_inputs(void* hv, const float* rates, unsigned n, const float* pos_xyz, const float* radius) {
    return guard([&] {
        Handle* h = (Handle*)hv;
        h->rates.assign(rates, rates + n);
        if (pos_xyz) {
            std::vector<coord3> p(n);
            std::vector<float> r(radius, radius + n);
            for (unsigned i = 0; i < n; i++) p[i] = coord3{pos_xyz[3 * i], pos_xyz[3 * i + 1], pos_xyz[3 * i + 2]};
            h->brain->setInputRateArray(h->rates.data(), n, p.data(), r.data());
        } else {
            h->brain->setInputRateArray(h->rates.data(), n);
        }
    });
}
int nch_set_rate(void* hv, unsigned i, float v) { return guard([&] { ((Handle*)hv)->rates.at(i) = v; }); }
int nch_set_input_near(void* hv, unsigned i, const uint32_t* ids, uint64_t n) { return guard([&] { B->setInputNear(i, ids, n); }); }
int nch_set_input_lastfire(void* hv, unsigned i, float t) { return guard([&] { B->setInputLastFire(i, t); }); }
int nch_add_input_offset(void* hv, unsigned i, float t) { return guard([&] { B->addInputOffset(i, t); }); }
int nch_set_input_enabled(void* hv, unsigned i, int en) { return guard([&] { B->setInputEnabled(i, en != 0); }); }
int nch_add_detector(void* hv, float x, float y, float z, float radius) {
    return guard([&] { coord3 c{x, y, z}; B->setDetectors(1, &c, &radius); });
}
int nch_detector_voltage(void* hv, unsigned id, float* out) { return guard([&] { *out = B->getDetectorVoltage(id); }); }
int nch_set_params(void* hv, float runSpeed, float learningRate, int runAll) {
    return guard([&] { B->runSpeed = runSpeed; B->learningRate = learningRate; B->runAll = runAll != 0; });
}
int nch_set_factors(void* hv, float pre, float post) { return guard([&] { B->presynapticFactor = pre; B->postsynapticFactor = post; }); }
int nch_set_candidate_smem(void* hv, unsigned n) { return guard([&] { B->candidateSmem = n; }); }
float nch_time(void* hv) { return B->getTime(); }
int nch_finalize(void* hv) { return guard([&] { B->finalize(); }); }
int nch_run(void* hv) { return guard([&] { B->run(); }); }
int nch_run_swept(void* hv, float* mean) { return guard([&] { float m = B->runSwept(); if (mean) *mean = m; }); }
int nch_reset_activities(void* hv) { return guard([&] { B->resetActivities(); }); }
uint64_t nch_neuron_count(void* hv) { return B->getNeuronCount(); }
uint64_t nch_synapse_count(void* hv) { return B->synapseCount(); }
int nch_export_network(void* hv, uint64_t* rowptr, uint32_t* pre, float* length, uint8_t* flag, float* xyz) {
    return guard([&] {
        B->finalize();
        memcpy(rowptr, B->csrRowptr().data(), B->csrRowptr().size() * 8);
        memcpy(pre, B->csrPre().data(), B->csrPre().size() * 4);
        memcpy(length, B->csrLength().data(), B->csrLength().size() * 4);
        memcpy(flag, B->csrFlags().data(), B->csrFlags().size());
        if (xyz) for (size_t i = 0; i < B->positions.size(); i++) { xyz[3 * i] = B->positions[i].x; xyz[3 * i + 1] = B->positions[i].y; xyz[3 * i + 2] = B->positions[i].z; }
    });
}
int nch_read_neurons(void* hv, float* pot, float* act, float* lastFire, float* lastRan) { return guard([&] { B->readNeurons(pot, act, lastFire, lastRan); }); }
int nch_read_synapses(void* hv, float* w, float* arrive, float* depol, float* lastArr, float* lastStart) {
    return guard([&] { B->readSynapses(w, arrive, depol, lastArr, lastStart); });
}
int nch_save_checkpoint(void* hv, const char* path) { return guard([&] { B->saveCheckpoint(path); }); }
int nch_load_checkpoint(void* hv, const char* path) {
    return guard([&] {
        Handle* h = (Handle*)hv;
        h->brain->loadCheckpoint(path, &h->rates);
        h->brain->attachInputRates(h->rates.data(), (unsigned)h->rates.size());
    });
}
int nch_state_signature(void* hv, uint64_t* out6) { return guard([&] { B->stateSignature(out6); }); }
int nch_record_fires(void* hv, int on) { return guard([&] { B->recordFires = on != 0; }); }
uint64_t nch_last_fires_count(void* hv) { return B->lastFiresNeuron().size(); }
int nch_last_fires(void* hv, uint32_t* neuron, float* time) {
    return guard([&] {
        auto& n = B->lastFiresNeuron(); auto& t = B->lastFiresTime();
        if (n.size()) { memcpy(neuron, n.data(), n.size() * 4); memcpy(time, t.data(), t.size() * 4); }
    });
}
unsigned nch_input_count(void* hv) { return (unsigned)B->inputNear().size(); }
uint64_t nch_input_near_count(void* hv, unsigned i) { return B->inputNear().at(i).size(); }
int nch_input_near(void* hv, unsigned i, uint32_t* out) {
    return guard([&] { auto v = B->inputNear().at(i); memcpy(out, v.data(), v.size() * 4); });
}
float nch_input_lastfire(void* hv, unsigned i) { return B->inputLastFire().at(i); }
// out[0..7]: fires, deliveries, loads accepted, loads dropped, plasticity calls, hidden rand, neuron runs, active visits
void nch_stats(void* hv, int total, uint64_t* out) {
    NeuCor::StepStats s = total ? B->totalStats() : B->lastStats();
    out[0] = s.fires; out[1] = s.deliveries; out[2] = s.loadsAccepted; out[3] = s.loadsDropped; out[4] = s.plasticityCalls;
    out[5] = s.hiddenRand; out[6] = s.neuronRuns; out[7] = s.activeVisits;
}
void nch_traffic(void* hv, uint64_t* h2d, uint64_t* d2h) { *h2d = B->h2dBytes(); *d2h = B->d2hBytes(); }
void* nch_engine(void* hv) { return B->engine(); }
int nch_snapshot_counts(void* hv, uint64_t* neurons, uint64_t* synapses, uint64_t* inputs) {
    return guard([&] {
        *neurons = B->getNeuronSnapshots().size();
        *synapses = B->getSynapseSnapshots().size();
        *inputs = B->getInputSnapshots().size();
    });
}
int nch_synapse_snapshots(void* hv, uint64_t cap, uint64_t* from, uint64_t* to, float* weight, float* prePot, float* postPot, uint8_t* inh) {
    return guard([&] {
        auto v = B->getSynapseSnapshots();
        for (size_t i = 0; i < v.size() && i < cap; i++) {
            from[i] = v[i].fromID; to[i] = v[i].toID; weight[i] = v[i].weight; prePot[i] = v[i].prePotential; postPot[i] = v[i].postPotential; inh[i] = v[i].inhibitory;
        }
    });
}
#undef B
}  // extern "C"
