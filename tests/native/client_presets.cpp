// Boundary test client: code written against the reference's NeuCor class surface (NeuCor.h:36-138) the way main.cpp and
// NeuCor_Renderer use it, compiled TWICE from this one file —
//   -I/root/reference/src  + /root/reference/src/NeuCor.cpp        (the reference itself; CPU container only), and
//   -Ineurocorrelation_b200/host + libneucor_host.so               (the drop-in: same class name, same members)
// — and run in lock-step: both binaries must print the same lines.
//
// It restates, headless, what the three non-interactive presets of main.cpp set up (STANDARD :80-124, FEW_NEURONS :162-189,
// ONE_INPUT :191-210) and what the renderer does per frame: scale runSpeed by the frame time (Renderer.cpp:632-637; a fixed
// 1/64 s frame here, so 4 ms/s becomes 0.0625 ms per run()), run(), read a detector (Renderer.cpp:1272-1274 — with a
// full-radius detector this runs every neuron in ascending ID, which is what makes the result independent of the
// reference's heap order), and read state: public snapshots plus the members the renderer reads through
// friendship (positions, potAct, inputHandler[i].{a, radius, enabled, lastFire}, voltageDetectors, resetActivities).
//
// The one reference-only block: the reference leaves Synapse::inhibitory uninitialised after std::vector reallocations
// (NeuCor.cpp:487-524); its build normalises the flags through the public Synapse::setWeight so that both
// sides start from the same network (the drop-in's constructor sets flag = weight < 0, the intent of NeuCor.cpp:473-475).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "NeuCor.h"

class NeuCor_Renderer {  // the renderer's friend access (NeuCor.h:98), headless
public:
    static std::vector<coord3>& positions(NeuCor& b) { return b.positions; }
    static std::vector<float>& potAct(NeuCor& b) { return b.potAct; }
    static size_t inputCount(NeuCor& b) { return b.inputHandler.size(); }
    static float inputLastFire(NeuCor& b, size_t i) { return b.inputHandler[i].lastFire; }
    static float inputRadius(NeuCor& b, size_t i) { return b.inputHandler[i].radius; }
    static coord3 inputPos(NeuCor& b, size_t i) { return b.inputHandler[i].a; }
    static void setInputEnabled(NeuCor& b, size_t i, bool en) { b.inputHandler[i].enabled = en; }
    static size_t detectorCount(NeuCor& b) { return b.voltageDetectors.size(); }
    static void resetActivities(NeuCor& b) { b.resetActivities(); }
    // the object graph the renderer walks (NeuCor.h:117-125; Renderer.cpp:655-699, 1433-1478, 1749-1862)
    static decltype(auto) neurons(NeuCor& b) { return (b.neurons); }
    static Neuron* getNeuron(NeuCor& b, size_t id) { return b.getNeuron(id); }
    static Synapse* getSynapse(NeuCor& b, std::pair<std::size_t, std::size_t> id) { return b.getSynapse(id); }
    static size_t pN(const Synapse& s) { return s.pN; }
    static size_t tN(const Synapse& s) { return s.tN; }
    static float prePot(const Synapse& s) { return s.getPrePot(); }
    static float postPot(const Synapse& s) { return s.getPostPot(); }
#ifdef CLIENT_REFERENCE_BUILD
    static void normaliseFlags(NeuCor& b) {
        for (auto& n : b.neurons)
            for (auto& s : n.outSynapses) s.setWeight(s.getWeight());
    }
#else
    static void normaliseFlags(NeuCor&) {}
#endif
};
typedef NeuCor_Renderer R;

static uint64_t fnv(uint64_t h, const void* p, size_t n) {
    const unsigned char* c = (const unsigned char*)p;
    for (size_t i = 0; i < n; i++) { h ^= c[i]; h *= 1099511628211ull; }
    return h;
}
static float randomUnit() { return static_cast<float>(rand()) / static_cast<float>(RAND_MAX); }

static void report(NeuCor& brain, int step, float volt) {
    // the raw vectors the renderer uploads every frame (Renderer.cpp:773-779) are read FIRST, with nothing but run() and the
    // detector read before them: they must already be what the snapshots say
    uint64_t hp = 1469598103934665603ull;
    {
        std::vector<float>& pa = R::potAct(brain);
        for (size_t i = 0; i + 1 < pa.size(); i += 2) { hp = fnv(hp, &pa[i], 4); hp = fnv(hp, &pa[i + 1], 4); }
    }
    uint64_t h = 1469598103934665603ull;
    for (auto& n : brain.getNeuronSnapshots()) { h = fnv(h, &n.potential, 4); h = fnv(h, &n.activity, 4); }
    uint64_t hw = 1469598103934665603ull;
    size_t nInh = 0;
    for (auto& s : brain.getSynapseSnapshots()) { hw = fnv(hw, &s.fromID, sizeof(s.fromID)); hw = fnv(hw, &s.toID, sizeof(s.toID)); hw = fnv(hw, &s.weight, 4); nInh += s.inhibitory ? 1 : 0; }
    uint64_t hi = 1469598103934665603ull;
    for (size_t i = 0; i < R::inputCount(brain); i++) { float lf = R::inputLastFire(brain, i); hi = fnv(hi, &lf, 4); }
    // one renderer frame over the object graph: per synapse both end potentials, weight and the end points' positions /
    // activities through getNeuron (Renderer.cpp:655-699); the raster rule and the activity list (Renderer.cpp:1749-1862);
    // one neuron's in-synapses through getSynapse(pair) (Renderer.cpp:1433-1434)
    uint64_t ho = 1469598103934665603ull;
    size_t nObj = 0, nLive = 0, nRaster = 0;
    for (auto& neu : R::neurons(brain)) {
        const float act = neu.activity(), pot = neu.potential();
        ho = fnv(ho, &act, 4); ho = fnv(ho, &pot, 4);
        if (brain.getTime() - neu.lastFire < 0.0625f) nRaster++;
        for (auto& syn : neu.outSynapses) {
            const size_t from = R::pN(syn), to = R::tN(syn);
            const float w = syn.getWeight(), a = R::prePot(syn), b = R::postPot(syn);
            const coord3 pa = R::getNeuron(brain, from)->position(), pb = R::getNeuron(brain, to)->position();
            const float la = logf(R::getNeuron(brain, to)->activity() + 1.f);
            ho = fnv(ho, &from, sizeof(from)); ho = fnv(ho, &to, sizeof(to)); ho = fnv(ho, &w, 4); ho = fnv(ho, &a, 4); ho = fnv(ho, &b, 4);
            ho = fnv(ho, &pa, sizeof(pa)); ho = fnv(ho, &pb, sizeof(pb)); ho = fnv(ho, &la, 4);
            nObj++;
            nLive += (a != 0.0f || b != 0.0f) ? 1 : 0;
        }
    }
    size_t nIn = 0;
    if (R::neurons(brain).size() > 7)
        for (auto& synM : R::getNeuron(brain, 7)->inSynapses) {
            Synapse* syn = R::getSynapse(brain, synM);
            const size_t from = R::pN(*syn);
            const float w = syn->getWeight();
            ho = fnv(ho, &from, sizeof(from)); ho = fnv(ho, &w, 4);
            nIn++;
        }
    uint32_t vb, tb;
    float t = brain.getTime();
    memcpy(&vb, &volt, 4); memcpy(&tb, &t, 4);
    printf("step %d time %08x volt %08x neurons %016llx potAct %s synapses %016llx inh %zu inputs %016llx objects %016llx (%zu synapses, %zu carrying a spike, %zu in-synapses of neuron 7, raster %zu)\n",
           step, tb, vb, (unsigned long long)h, hp == h ? "same" : "DIFFERENT", (unsigned long long)hw, nInh, (unsigned long long)hi, (unsigned long long)ho, nObj, nLive, nIn, nRaster);
}

int main(int argc, char** argv) {
    const char* preset = argc > 1 ? argv[1] : "standard";
    const unsigned seed = argc > 2 ? (unsigned)atoi(argv[2]) : 1u;
    const int steps = argc > 3 ? atoi(argv[3]) : 400;
    const float frame = 1.0f / 64.0f;  // seconds per frame
    srand(seed);
    if (!strcmp(preset, "standard") || !strcmp(preset, "one_input")) {
        const bool standard = !strcmp(preset, "standard");
        NeuCor brain(750);
        R::normaliseFlags(brain);
        brain.runAll = false;
        brain.runSpeed = 4.0f;
        std::vector<float> inputs, radius;
        std::vector<coord3> pos;
        if (standard) {
            for (int i = 0; i < 3; i++) inputs.push_back(randomUnit() * 75.0f);
            radius = {0.8f, 0.8f, 0.8f};
            pos = {{cosf(0.0f) * 2.0f, sinf(0.0f) * 2.0f, 0.0f}, {cosf(2.0944f) * 2.0f, sinf(2.0944f) * 2.0f, 0.0f}, {cosf(4.1888f) * 2.0f, sinf(4.1888f) * 2.0f, 0.0f}};
        } else {
            inputs = {35.0f}; radius = {0.8f}; pos = {{2.0f, 0.0f, 0.0f}};
        }
        brain.setInputRateArray(inputs.data(), (unsigned)inputs.size(), pos.data(), radius.data());
        coord3 centre{0.0f, 0.0f, 0.0f};
        float everything = 1e9f;
        brain.setDetectors(1, &centre, &everything);
        printf("%s seed %u: %zu neurons, %zu synapses, %zu inputs, %zu detectors\n", preset, seed, brain.getNeuronCount(), brain.getSynapseSnapshots().size(),
               brain.getInputSnapshots().size(), R::detectorCount(brain));
        srand(777);
        for (int k = 0; k < steps; k++) {
            if (standard) {  // onFrame, main.cpp:100-105
                for (float& input : inputs) {
                    input += (randomUnit() - 0.5f) * 2.0f;
                    input = std::min(std::max(input, 0.0f), 75.0f);
                }
                inputs[1] = inputs[0];
            }
            if (k == steps / 2) {  // what the GUI's buttons do: toggle an input (Renderer.cpp:2036), reset the activities (:1619)
                R::setInputEnabled(brain, 0, false);
                R::resetActivities(brain);
                brain.learningRate = 0.5f;
            }
            if (k == steps / 2 + 40) R::setInputEnabled(brain, 0, true);
            const float staticRunSpeed = brain.runSpeed;  // realRunspeed, Renderer.cpp:632-637
            brain.runSpeed *= frame;
            brain.run();
            brain.runSpeed = staticRunSpeed;
            const float volt = brain.getDetectorVoltage(0);
            if (k % 50 == 49 || k == steps - 1) report(brain, k, volt);
        }
    } else if (!strcmp(preset, "few_neurons")) {
        NeuCor brain(0);
        brain.runAll = true;
        brain.runSpeed = 0.02f;
        std::vector<coord3> np = {{0, 0, 0}, {0.3f, 0.3f, 0}, {0.3f, -0.3f, 0}};
        for (const coord3& p : np) brain.createNeuron(p);
        brain.createSynapse(1, 0, 0.5f);
        brain.createSynapse(2, 0, 0.5f);
        R::normaliseFlags(brain);
        std::vector<float> inputs = {50.0f, 50.0f, 50.0f}, radius = {0.1f, 0.1f, 0.1f};
        brain.setInputRateArray(inputs.data(), 3, np.data(), radius.data());
        brain.addInputOffset(1, 2.0f);
        brain.addInputOffset(2, -2.0f);
        for (int k = 0; k < steps; k++) brain.run();
        // the essay's known answer (section 2.5.1): w(0->1) rises to 1, w(0->2) falls to 0.  runAll pops equal-time neuron events in
        // heap order in the reference and in ascending ID here, so only the outcome is compared, not every bit.
        auto syn = brain.getSynapseSnapshots();
        printf("few_neurons: %zu synapses, w(0->%zu) %s 0.9, w(0->%zu) %s 0.1\n", syn.size(), syn[0].toID, syn[0].weight > 0.9f ? ">" : "<=", syn[1].toID,
               syn[1].weight < 0.1f ? "<" : ">=");
    } else {
        fprintf(stderr, "unknown preset %s\n", preset);
        return 2;
    }
    return 0;
}
