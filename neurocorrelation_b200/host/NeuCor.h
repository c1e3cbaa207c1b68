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
#include <string>
#include <unordered_map>
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

    // ---- extensions ----
    // On-disk network + state file (host/checkpoint.cpp): the post-sorted CSR incl. the flag bytes, positions, the complete
    // dynamic state, input firers / detectors with their `near` lists and libc's rand() position.  loadCheckpoint works on an
    // empty NeuCor(0) before its first run(); the caller owns the rate array as with setInputRateArray (NeuCor.cpp:46-48):
    // the saved rates are handed back and attachInputRates() points the brain at the caller's array.
    void saveCheckpoint(const char* path);
    void loadCheckpoint(const char* path, std::vector<float>* inputRates = nullptr);
    void attachInputRates(float inputs[], unsigned inputCount);
    struct Synapse {                // one row entry of the exported network
        uint32_t from, to;
        float weight, length;
        uint8_t inhibitory;
    };
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

#endif
