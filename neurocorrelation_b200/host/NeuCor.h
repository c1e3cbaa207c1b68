// Host-side drop-in for the reference's simulation-core class (/root/reference/src/NeuCor.h:36-138).
//
// Same class name, same public members with the same meaning and error behaviour
// (std::out_of_range on bad IDs, run() a no-op for runSpeed <= 0) so that main.cpp-style code
// builds against it unchanged; underneath, the network lives on a B200 as a post-synaptic-sorted
// CSR and run() drives the CUDA engine through the C ABI in include/neucor_b200.h.  There is no
// CPU execution path: constructing the device engine fails loudly without a GPU.
//
// What stays on the host (exactly as the reference does it, same libc rand() stream):
//   * network construction — NeuCor(int), createNeuron, createSynapse, makeConnections
//     (NeuCor.cpp:17-42,154-195,418-444,463-486);
//   * InputFirer::schedule and the background-firing draws of every run() (NeuCor.cpp:333-345,604-607);
//   * the hidden rand() calls of synapticPlasticity (NeuCor.cpp:752): the device reports their number per
//     window and run() advances rand() by it.
// Extensions (not in the reference): importNetwork (exported CSR incl. the flag byte),
// runSwept() (run() fused with a full-radius detector read — the oracle's sweep mode), raw state readers.
#ifndef NEUCOR_B200_HOST_NEUCOR_H
#define NEUCOR_B200_HOST_NEUCOR_H

#include <math.h>
#include <stdint.h>

#include <cstddef>
#include <deque>
#include <map>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

struct nc_engine;
struct nc_event;

struct coord3 {
    float x, y, z;
    float getDist(coord3 c) const {  // NeuCor.h:16-18 (powf(d, 2) is d*d in the reference's -O3 build)
        float dx = x - c.x, dy = y - c.y, dz = z - c.z;
        float s = dx * dx;
        s = s + dy * dy;
        s = s + dz * dz;
        return sqrtf(s);
    }
    void setNAN() { x = NAN; y = NAN; z = NAN; }
};

class NeuCor;
class Neuron;

// ---- object view: what NeuCor_Renderer walks through friendship (reference NeuCor.h:117-125, 192-294) ----------------
// The network itself lives on the device as a CSR; these objects are a host MIRROR of it with the reference's names and
// access levels, so that renderer code such as `for (auto& neu : brain->neurons) for (auto& syn : neu.outSynapses)
// ... syn.getPrePot() ... brain->getNeuron(syn.tN)->position()` (Renderer.cpp:655-699, 852-853, 1092, 1433-1478, 1749-1862)
// compiles and reads current values.  The mirror is brought up to date lazily — on the first access through
// `neurons` / getNeuron / getSynapse after a run() — and is read-only: writes to it do not reach the simulation.
class Synapse {
public:
    float getWeight() const { return weight; }
protected:
    friend class NeuCor;
    friend class Neuron;
    friend class NeuCor_Renderer;
    float getPrePot() const { return prePot_; }    // Synapse::getPrePot / getPostPot (NeuCor.cpp:547-567), evaluated on the device
    float getPostPot() const { return postPot_; }  // at the brain's current time when the mirror is refreshed
    std::size_t pN = 0;                            // parent neuron ID
    std::size_t tN = 0;                            // target neuron ID
    float lastSpikeStart = 0.0f;
    float lastSpikeArrival = 0.0f;
    float AP_fireTime = 0.0f;
    bool inhibitory = false;
private:
    float length = 0.0f;
    float weight = 0.0f;
    float prePot_ = 0.0f, postPot_ = 0.0f;
    uint64_t slot_ = 0;                            // position in the post-sorted CSR
};

class Neuron {
public:
    std::vector<Synapse> outSynapses;               // creation order, as the reference owns them (NeuCor.h:211)
    std::map<std::size_t, std::size_t> inSynapses;  // {from neuron, this neuron} (NeuCor.h:212, NeuCor.cpp:467)
    inline coord3 position() const;
    inline float potential() const;
    inline float activity() const;
    std::size_t getID() const { return ownID; }
    float lastFire = NAN;                           // NeuCor.h:226 (NAN until the first fire, NeuCor.cpp:391)
private:
    friend class NeuCor;
    NeuCor* parentNet = nullptr;
    std::size_t ownID = 0;
};

// Stands in for `std::deque<Neuron> neurons` (NeuCor.h:117) with the same reading interface; begin() / at() / [] bring the
// mirror up to date first.
class NeuronList {
public:
    typedef std::deque<Neuron>::iterator iterator;
    inline iterator begin();
    iterator end() { return d_.end(); }
    std::size_t size() const { return d_.size(); }
    bool empty() const { return d_.empty(); }
    inline Neuron& at(std::size_t i);
    inline Neuron& operator[](std::size_t i);
private:
    friend class NeuCor;
    std::deque<Neuron> d_;
    NeuCor* owner_ = nullptr;
};

class NeuCor {
public:
    struct NeuronSnapshot { std::size_t id; coord3 position; float potential; float activity; };
    struct SynapseSnapshot {
        std::size_t fromID, toID;
        coord3 from, to;
        float weight, prePotential, postPotential;
        bool inhibitory;
    };
    struct InputSnapshot { std::size_t id; coord3 position; float radius; float frequency; bool enabled; };

    NeuCor(int n_neurons);
    ~NeuCor();
    NeuCor(const NeuCor&) = delete;
    NeuCor& operator=(const NeuCor&) = delete;

    void run();
    float runSpeed;
    bool runAll;
    float getTime() const;
    float learningRate;
    float presynapticTraceDecay, postsynapticTraceDecay;
    float presynapticFactor, postsynapticFactor;

    void setInputRateArray(float inputs[], unsigned inputCount, coord3 inputPositions[] = nullptr, float inputRadius[] = nullptr);
    void addInputOffset(unsigned inputID, float t);

    void setDetectors(unsigned detectorNumber, coord3 detectorPositions[] = nullptr, float detectorRadius[] = nullptr);
    float getDetectorVoltage(unsigned ID);
    std::vector<float> getDetectorVoltages();

    void createNeuron(coord3 position);
    void createSynapse(std::size_t toID, std::size_t fromID, float weight);
    void makeConnections();
    std::size_t getNeuronCount() const;
    std::vector<NeuronSnapshot> getNeuronSnapshots() const;
    std::vector<SynapseSnapshot> getSynapseSnapshots() const;
    std::vector<InputSnapshot> getInputSnapshots() const;

    // ---- what NeuCor_Renderer reads through friendship in the reference (NeuCor.h:98-106) ----
    std::vector<coord3> positions;  // neuron positions in ID order
    std::vector<float> potAct;      // (potential, activity) pairs in ID order; refreshed by syncState()
    void resetActivities();
    void setInputEnabled(unsigned inputID, bool enabled);  // InputFirer::enabled (Renderer.cpp:2036)
    // the object view (see class Synapse / Neuron above): kept for networks built through createNeuron / createSynapse /
    // NeuCor(n) and for imported host networks of up to objectViewLimit synapses; empty otherwise (device-resident imports, shards)
    NeuronList neurons;                                    // NeuCor.h:117
    Neuron* getNeuron(std::size_t ID);                     // NeuCor.h:118-119: std::out_of_range on a bad ID
    const Neuron* getNeuron(std::size_t ID) const;
    Synapse* getSynapse(std::size_t fromID, std::size_t toID);  // NeuCor.h:120-123 (NeuCor.cpp:244-263: (from, to), nullptr when absent)
    const Synapse* getSynapse(std::size_t fromID, std::size_t toID) const;
    Synapse* getSynapse(std::pair<std::size_t, std::size_t> ID);
    const Synapse* getSynapse(std::pair<std::size_t, std::size_t> ID) const;
    void refreshObjects();                                 // device -> object view, if anything ran since the last refresh
    std::size_t objectViewLimit = (std::size_t)1 << 20;
    // run() ends with syncState() (positions / potAct / lastFire mirrors current, as NeuCor_Renderer expects every frame,
    // Renderer.cpp:773-779): 8 bytes per neuron device -> host per run().  Callers stepping large networks switch it off.
    bool mirrorAfterRun = true;

    // ---- extensions ----
    // On-disk network + state file (host/checkpoint.cpp): the post-sorted CSR incl. the flag bytes, positions, the complete
    // dynamic state, input firers / detectors with their `near` lists and libc's rand() position.  loadCheckpoint works on an
    // empty NeuCor(0) before its first run(); the caller owns the rate array as with setInputRateArray (NeuCor.cpp:46-48):
    // the saved rates are handed back and attachInputRates() points the brain at the caller's array.
    void saveCheckpoint(const char* path);
    void loadCheckpoint(const char* path, std::vector<float>* inputRates = nullptr);
    void attachInputRates(float inputs[], unsigned inputCount);
    // Replaces the network by an exported one (post-sorted CSR incl. the reference's flag byte). Only before the first run().
    void importNetwork(std::size_t n, const uint64_t* rowptr, const uint32_t* pre, const float* weight, const float* length,
                       const uint8_t* inhibitory, const float* positions_xyz /* may be null */);
    // Same with the CSR already resident on the device `deviceOrdinal` (device pointers; copied at finalize()).
    void importNetworkDevice(std::size_t n, uint64_t synapses, const uint64_t* d_rowptr, const uint32_t* d_pre, const float* d_weight,
                             const float* d_length, const uint8_t* d_inhibitory);
    // Multi-GPU: this process simulates the neuron-ID range [N*rank/world, N*(rank+1)/world) and the
    // synapses incoming to it; every process makes the same calls in the same order with the same libc rand() state.
    // Call before the first run().  The fire exchange runs inside run(): over an NCCL communicator created from `commId`
    // (a 128-byte ncclUniqueId made by rank 0, see nc_comm_unique_id) or through a caller-provided all-gather.
    void setShard(int rank, int world);
    void setCommId(const void* commId128);
    void setExchange(int (*allgather)(void* ctx, const void* send, void* recv, uint64_t bytes), void* ctx);
    float globalMinDelay = 0.0f;    // device-resident shards only: smallest 2*length over ALL shards (bounds the window)
    std::size_t shardRow0() const { return row0_; }
    std::size_t shardRows() const { return engine_ ? nRows_ : positions.size(); }
    std::size_t shardSynapses() const { return engine_ ? sLocal_ : synapseCount(); }
    // The shard's rows of a network of `n` neurons, already resident on the device (local rowptr starting at 0).
    void importShardDevice(std::size_t n, uint64_t localSynapses, const uint64_t* d_rowptr, const uint32_t* d_pre, const float* d_weight,
                           const float* d_length, const uint8_t* d_inhibitory);
    void setPositions(const float* xyz, std::size_t n);  // positions of a network imported without them (n = neuron count); before setInputRateArray / setDetectors
    void setInputNear(unsigned inputID, const uint32_t* ids, std::size_t n);  // overrides an input's `near` list
    void setInputLastFire(unsigned inputID, float t);
    float runSwept();               // run() + "run every neuron at the new time, ascending ID"; returns the mean potential
    bool sweepReturnsMean = true;   // false: runSwept() skips the device->host read of all potentials and returns 0
    void syncState();               // device -> potAct / lastFire mirrors
    // network in CSR order (valid after finalize) and raw state readers
    void finalize();                // builds the CSR and uploads it (implicit on first run)
    bool finalized() const { return engine_ != nullptr; }
    std::size_t synapseCount() const;
    const std::vector<uint64_t>& csrRowptr() const { return rowptr_; }
    const std::vector<uint32_t>& csrPre() const { return pre_; }
    const std::vector<float>& csrLength() const { return length_; }
    const std::vector<uint8_t>& csrFlags() const { return flag_; }
    void readNeurons(float* pot, float* act, float* lastFire, float* lastRan);
    void readSynapses(float* weight, float* arrive, float* depol, float* lastArrival, float* lastStart);
    void stateSignature(uint64_t out6[6]);  // six per-field checksums of this shard's state, computed on the device
    bool recordFires = false;       // keep (neuron, time) of every fire of the last run() — the raster source (Renderer.cpp:1856-1862)
    const std::vector<uint32_t>& lastFiresNeuron() const { return firesNeuron_; }
    const std::vector<float>& lastFiresTime() const { return firesTime_; }
    std::vector<float> inputLastFire() const;
    std::vector<std::vector<uint32_t>> inputNear() const;
    struct StepStats { uint64_t fires, deliveries, loadsAccepted, loadsDropped, plasticityCalls, hiddenRand, neuronRuns, activeVisits; };
    StepStats lastStats() const { return lastStats_; }
    StepStats totalStats() const { return totalStats_; }
    uint64_t h2dBytes() const { return h2dBytes_; }
    uint64_t d2hBytes() const { return d2hBytes_; }
    nc_engine* engine() { return engine_; }
    int deviceOrdinal = 0;          // CUDA device used by finalize()
    unsigned candidateSmem = 0;     // nc_config.cand_smem override (0 = default)

private:
    friend class NeuCor_Renderer;   // as in the reference (NeuCor.h:98): the renderer reads inputHandler, voltageDetectors, ... directly
    static bool checkpointPeekRand(uint32_t x31[31]);
    static bool checkpointPokeRand(const uint32_t x31[31]);
    struct InputFirer {
        coord3 a; float radius; bool enabled; float lastFire; std::vector<uint32_t> near;
    };
    struct VoltageDetector { coord3 a; float radius; std::vector<uint32_t> near; };
    struct OutSyn { uint32_t to; float weight, length; uint8_t flag; };

    void nearList(coord3 c, float radius, std::vector<uint32_t>& out);
    static uint64_t packKey(long x, long y, long z);
    uint64_t cellKey(float x, float y, float z) const;
    std::unordered_map<uint64_t, std::vector<uint32_t>> gridCells_;  // unit-cell grid over `positions` for nearList
    float gridLo_[3] = {0, 0, 0};
    std::size_t gridN_ = 0;
    bool gridOk_ = false;
    void scheduleInput(unsigned i, float deltaT, float frequency, std::vector<nc_event>& ev);
    void window(float t0, float t1, int flags, std::vector<nc_event>& ev);
    float stepInternal(bool sweep);
    void check(int rc, const char* what);

    bool viewDirty_ = true;
    void buildImportedView();
    static void viewSet(Synapse& s, std::size_t from, std::size_t to, float weight, float length, bool inhibitory, uint64_t slot);
    float currentTime = 0.0f;
    unsigned totalGenNeurons = 0;
    float* inputArray = nullptr;
    unsigned inputArraySize = 0;
    std::vector<InputFirer> inputHandler;
    std::vector<VoltageDetector> voltageDetectors;
    std::vector<std::vector<OutSyn>> out_;   // per presynaptic neuron, creation order (the reference's outSynapses)
    // finalized CSR
    std::vector<uint64_t> rowptr_;
    std::vector<uint32_t> pre_;
    std::vector<float> weight_, length_;
    std::vector<uint8_t> flag_;
    bool imported_ = false;
    struct { const uint64_t* rowptr; const uint32_t* pre; const float *weight, *length; const uint8_t* inh; uint64_t S; bool set; } dev_ = {};
    nc_engine* engine_ = nullptr;
    int rank_ = 0, world_ = 1;
    std::size_t row0_ = 0, nRows_ = 0, sLocal_ = 0;
    bool shardImport_ = false;
    char commId_[128]; bool haveCommId_ = false;
    int (*xchgFn_)(void*, const void*, void*, uint64_t) = nullptr; void* xchgCtx_ = nullptr;
    float minDelay_ = 0.0f;
    float preDecayLatched_ = 0.75f, postDecayLatched_ = 0.65f;
    std::vector<uint32_t> firesNeuron_;
    std::vector<float> firesTime_;
    std::vector<float> lastFireMirror_;
    std::vector<nc_event> events_, winEvents_;
    std::vector<float> schedScratch_;
    struct RandStream;              // look-ahead view of libc's rand() stream (NeuCor.cpp)
    RandStream* rs_ = nullptr;
    uint64_t lastHidden_ = 0;
    bool randThread_ = false;
    int deviceRand_ = -1;           // -1 undecided, 1 rand() stream on the device, 0 on the host (NC_HOST_RAND=1 or not glibc's TYPE_3 generator)
    bool devRandThisRun_ = false;
    uint32_t randMirror_[31];       // libc's state as this class last left it (detects draws by the application in between)
    bool randMirrorValid_ = false;
    StepStats lastStats_ = {}, totalStats_ = {};
    uint64_t h2dBytes_ = 0, d2hBytes_ = 0;
};

// ---- object view: inline members that need the complete NeuCor ----
inline coord3 Neuron::position() const { return parentNet->positions[ownID]; }
inline float Neuron::potential() const { return parentNet->potAct[2 * ownID]; }
inline float Neuron::activity() const { return parentNet->potAct[2 * ownID + 1]; }
inline NeuronList::iterator NeuronList::begin() { if (owner_) owner_->refreshObjects(); return d_.begin(); }
inline Neuron& NeuronList::at(std::size_t i) { if (owner_) owner_->refreshObjects(); return d_.at(i); }
inline Neuron& NeuronList::operator[](std::size_t i) { if (owner_) owner_->refreshObjects(); return d_[i]; }

#endif
