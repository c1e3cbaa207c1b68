// On-disk network + state file of the drop-in NeuCor class.
//
// The reference has no serialisation of any kind.  One file holds what the oracle harness exports
// from the reference process — the post-sorted CSR with the raw inhibitory flag bytes, positions, input firers with their
// `near` lists — plus the complete dynamic state of the engine and libc's rand() position, so it serves three purposes:
// exchanging networks with the reference harness, building large networks once and re-using them, and checkpoint / resume
// (a resumed run continues bit for bit, background firing included).
//
// Layout (little endian): magic "NCB200\2\0", header (counts, time, parameters), shard header, then the arrays in the order
// written below, each as raw elements.  Arrays are streamed through a bounded buffer, never held twice.
//
// Sharded runs (setShard(rank, world)): every process writes ITS OWN file — the rows [row0, row0 + rows) of the CSR with a
// rowptr that starts at 0, the state of those rows and synapses, and everything that is identical on all shards (positions
// of the whole network, firers, detectors, rand() position, the network-wide smallest delay that bounds the window).
// A shard file is loaded by a process that has been given the same (rank, world); version-1 files (whole networks only,
// no shard header) are still read.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/neucor_b200.h"
#include "NeuCor.h"

namespace {
const char MAGIC[8] = {'N', 'C', 'B', '2', '0', '0', 2, 0};  // byte 6 = format version
struct Header {
    char magic[8];
    uint64_t neurons, synapses, inputs, detectors;
    float time, runSpeed, learningRate, preDecay, postDecay, preFactor, postFactor;
    uint32_t runAll, hasRand, hasPositions, reserved;
};
struct ShardHeader {  // version >= 2
    uint32_t rank, world;
    uint64_t row0, rows;
    float minDelay;  // smallest synaptic delay of the WHOLE network (every shard must split windows identically)
    uint32_t reserved;
};
struct File {
    FILE* f;
    std::string path;
    File(const char* p, const char* mode) : f(fopen(p, mode)), path(p) {
        if (!f) throw std::runtime_error("NeuCor checkpoint: cannot open " + path);
    }
    ~File() { if (f) fclose(f); }
    void put(const void* p, size_t n) { if (n && fwrite(p, 1, n, f) != n) throw std::runtime_error("NeuCor checkpoint: write failed: " + path); }
    void get(void* p, size_t n) { if (n && fread(p, 1, n, f) != n) throw std::runtime_error("NeuCor checkpoint: file truncated: " + path); }
    template <typename T> void putv(const std::vector<T>& v) { put(v.data(), v.size() * sizeof(T)); }
    template <typename T> void getv(std::vector<T>& v, size_t n) { v.resize(n); get(v.data(), n * sizeof(T)); }
};
}  // namespace

void NeuCor::saveCheckpoint(const char* path) {
    finalize();
    const size_t N = positions.size();   // neurons of the whole network
    const size_t R = nRows_;             // rows of this shard (= N for world 1)
    const uint64_t S = sLocal_;
    File out(path, "wb");
    Header h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, MAGIC, 8);
    h.neurons = N; h.synapses = S; h.inputs = inputHandler.size(); h.detectors = voltageDetectors.size();
    h.time = currentTime; h.runSpeed = runSpeed; h.learningRate = learningRate;
    h.preDecay = preDecayLatched_; h.postDecay = postDecayLatched_; h.preFactor = presynapticFactor; h.postFactor = postsynapticFactor;
    h.runAll = runAll ? 1u : 0u;
    uint32_t rnd[31];
    h.hasRand = checkpointPeekRand(rnd) ? 1u : 0u;
    h.hasPositions = 1u;
    out.put(&h, sizeof(h));
    ShardHeader sh;
    memset(&sh, 0, sizeof(sh));
    sh.rank = (uint32_t)rank_; sh.world = (uint32_t)world_; sh.row0 = row0_; sh.rows = R; sh.minDelay = minDelay_;
    out.put(&sh, sizeof(sh));
    // ---- network (as it lives on the device: also valid for networks imported from device memory) ----
    {
        std::vector<uint64_t> rowptr(R + 1);
        std::vector<uint32_t> pre(S);
        std::vector<float> length(S);
        std::vector<uint8_t> flag(S);
        check(nc_read_network(engine_, rowptr.data(), pre.data(), length.data(), flag.data()), "nc_read_network");
        out.putv(rowptr); out.putv(pre); out.putv(length); out.putv(flag);
    }
    {
        std::vector<float> xyz(3 * N);
        for (size_t i = 0; i < N; i++) { xyz[3 * i] = positions[i].x; xyz[3 * i + 1] = positions[i].y; xyz[3 * i + 2] = positions[i].z; }
        out.putv(xyz);
    }
    // ---- dynamic state ----
    {
        std::vector<float> a(S);
        float* dst[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        for (int k = 0; k < 5; k++) {  // weight, arrive, depol, lastArr, lastStart — one array at a time
            for (int j = 0; j < 5; j++) dst[j] = j == k ? a.data() : nullptr;
            check(nc_read_synapses(engine_, dst[0], dst[1], dst[2], dst[3], dst[4]), "nc_read_synapses");
            out.putv(a);
        }
    }
    {
        std::vector<float> potAct2(2 * R), lastFire(R), lastRan(R), actStart(R);
        std::vector<uint32_t> firings(R);
        check(nc_read_neurons(engine_, potAct2.data(), lastFire.data(), lastRan.data()), "nc_read_neurons");
        check(nc_read_neuron_counters(engine_, actStart.data(), firings.data()), "nc_read_neuron_counters");
        out.putv(potAct2); out.putv(lastFire); out.putv(lastRan); out.putv(actStart); out.putv(firings);
    }
    // ---- input firers, detectors ----
    for (size_t i = 0; i < inputHandler.size(); i++) {
        const InputFirer& f = inputHandler[i];
        const float rate = inputArray != nullptr && i < inputArraySize ? inputArray[i] : 0.0f;
        const float rec[7] = {f.a.x, f.a.y, f.a.z, f.radius, f.lastFire, rate, f.enabled ? 1.0f : 0.0f};
        const uint64_t n = f.near.size();
        out.put(rec, sizeof(rec)); out.put(&n, 8); out.putv(f.near);
    }
    for (const VoltageDetector& d : voltageDetectors) {
        const float rec[4] = {d.a.x, d.a.y, d.a.z, d.radius};
        const uint64_t n = d.near.size();
        out.put(rec, sizeof(rec)); out.put(&n, 8); out.putv(d.near);
    }
    if (h.hasRand) out.put(rnd, sizeof(rnd));
}

void NeuCor::loadCheckpoint(const char* path, std::vector<float>* inputRates) {
    if (engine_ || imported_ || !positions.empty()) throw std::logic_error("NeuCor::loadCheckpoint: only into an empty NeuCor(0)");
    File in(path, "rb");
    Header h;
    in.get(&h, sizeof(h));
    if (memcmp(h.magic, MAGIC, 6) != 0 || h.magic[7] != 0 || (h.magic[6] != 1 && h.magic[6] != 2))
        throw std::runtime_error("NeuCor::loadCheckpoint: not a NeuCor checkpoint: " + in.path);
    const size_t N = h.neurons;
    const uint64_t S = h.synapses;
    ShardHeader sh;
    memset(&sh, 0, sizeof(sh));
    sh.world = 1; sh.rows = N;
    if (h.magic[6] >= 2) in.get(&sh, sizeof(sh));
    if ((int)sh.world != world_ || (int)sh.rank != rank_)
        throw std::logic_error("NeuCor::loadCheckpoint: the file holds shard " + std::to_string(sh.rank) + " of " + std::to_string(sh.world) +
                               ", this process is shard " + std::to_string(rank_) + " of " + std::to_string(world_) + " (setShard first)");
    const size_t R = sh.rows;
    {
        std::vector<uint64_t> rowptr, local;
        std::vector<uint32_t> pre;
        std::vector<float> length, weight0(S, 0.0f), xyz;
        std::vector<uint8_t> flag;
        in.getv(local, R + 1); in.getv(pre, S); in.getv(length, S); in.getv(flag, S); in.getv(xyz, 3 * N);
        if (world_ == 1) rowptr.swap(local);
        else {  // the shard's rows inside an otherwise empty CSR of the whole network: finalize() slices exactly them out again
            if (sh.row0 + R > N || local[R] != S) throw std::runtime_error("NeuCor::loadCheckpoint: inconsistent shard header: " + in.path);
            rowptr.assign(N + 1, 0);
            for (size_t i = 0; i <= N; i++) rowptr[i] = i <= sh.row0 ? 0 : i <= sh.row0 + R ? local[i - sh.row0] : S;
            globalMinDelay = sh.minDelay;
        }
        importNetwork(N, rowptr.data(), pre.data(), weight0.data(), length.data(), flag.data(), xyz.data());
    }
    runSpeed = h.runSpeed; learningRate = h.learningRate; presynapticTraceDecay = h.preDecay; postsynapticTraceDecay = h.postDecay;
    presynapticFactor = h.preFactor; postsynapticFactor = h.postFactor; runAll = h.runAll != 0;
    finalize();
    if (row0_ != sh.row0 && world_ > 1) throw std::runtime_error("NeuCor::loadCheckpoint: the shard's row range does not match this rank's: " + in.path);
    {
        std::vector<float> a;
        const float* src[5];
        for (int k = 0; k < 5; k++) {
            in.getv(a, S);
            for (int j = 0; j < 5; j++) src[j] = j == k ? a.data() : nullptr;
            if (k == 0) weight_.assign(a.begin(), a.end());
            check(nc_write_synapses(engine_, src[0], src[1], src[2], src[3], src[4]), "nc_write_synapses");
        }
    }
    {
        std::vector<float> potAct2, lastFire, lastRan, actStart;
        std::vector<uint32_t> firings;
        in.getv(potAct2, 2 * R); in.getv(lastFire, R); in.getv(lastRan, R); in.getv(actStart, R); in.getv(firings, R);
        check(nc_write_neurons(engine_, potAct2.data(), lastFire.data(), lastRan.data(), actStart.data(), firings.data()), "nc_write_neurons");
        std::copy(potAct2.begin(), potAct2.end(), potAct.begin() + 2 * row0_);  // (the mirror spans the whole network)
    }
    currentTime = h.time;
    std::vector<float> rates(h.inputs, 0.0f);
    inputHandler.clear();
    for (uint64_t i = 0; i < h.inputs; i++) {
        float rec[7];
        uint64_t n = 0;
        in.get(rec, sizeof(rec)); in.get(&n, 8);
        InputFirer f;
        f.a = coord3{rec[0], rec[1], rec[2]}; f.radius = rec[3]; f.lastFire = rec[4]; f.enabled = rec[6] != 0.0f;
        rates[i] = rec[5];
        in.getv(f.near, n);
        inputHandler.push_back(std::move(f));
    }
    voltageDetectors.clear();
    for (uint64_t i = 0; i < h.detectors; i++) {
        float rec[4];
        uint64_t n = 0;
        in.get(rec, sizeof(rec)); in.get(&n, 8);
        VoltageDetector d;
        d.a = coord3{rec[0], rec[1], rec[2]}; d.radius = rec[3];
        in.getv(d.near, n);
        voltageDetectors.push_back(std::move(d));
    }
    if (h.hasRand) {
        uint32_t rnd[31];
        in.get(rnd, sizeof(rnd));
        checkpointPokeRand(rnd);
    }
    // the caller owns the rate array (NeuCor.cpp:46-48): hand the saved rates back and let it call setInputRateArray's
    // pointer-only form (attachInputRates) once its array is in place
    if (inputRates) *inputRates = rates;
    inputArray = nullptr; inputArraySize = 0;
}

void NeuCor::attachInputRates(float inputs[], unsigned inputCount) {
    inputArray = inputs;
    inputArraySize = inputCount;
}
