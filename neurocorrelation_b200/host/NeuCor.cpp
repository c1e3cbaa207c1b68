// Host side of the drop-in NeuCor class: network construction and per-run scheduling exactly as the
// reference does them on the CPU (/root/reference/src/NeuCor.cpp:17-366, 583-617), with the hot loop
// (NeuCor.cpp:609-616 and everything it dispatches to) handed to the CUDA engine through the C ABI.
// Compile with -ffp-contract=off: every float expression below is typed as in the reference.
#include "NeuCor.h"

#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cassert>
#include <cmath>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <stdexcept>
#include <unordered_map>

#include "../../include/neucor_b200.h"

namespace {
inline float randomUnit() {  // NeuCor.cpp:12-14
    return static_cast<float>(rand()) / static_cast<float>(RAND_MAX);
}

// Bulk access to libc's rand() stream.  run() must draw once per neuron per call (NeuCor.cpp:604-607) from the very
// generator the application seeds with srand() and shares with the core, and libc's rand() costs ~8 ns a call (lock +
// PLT).  glibc's generator is the TYPE_3 additive feedback generator r[i] = r[i-31] + r[i-3], output r[i] >> 1; its state
// array is reachable through the public setstate() API, so a window borrows the live state, advances it in place with
// the same recurrence, and hands it back — every draw is exactly the value rand() would have returned, and rand()
// continues from the right position afterwards.  A self-test at first use compares the two; on any mismatch (other
// libc, other generator type) the window silently degrades to calling rand().
class RandWindow {
public:
    RandWindow() {
        if (!checked_) { checked_ = true; usable_ = selfTest(); }
        if (usable_) open();
    }
    ~RandWindow() { close(); }
    inline int next() {
        if (!words_) return rand();
        int32_t* st = words_ + 1;
        st[f_] = (int32_t)((uint32_t)st[f_] + (uint32_t)st[r_]);
        int out = (int)(((uint32_t)st[f_]) >> 1);
        if (++f_ == 31) f_ = 0;
        if (++r_ == 31) r_ = 0;
        return out;
    }
    void skip(uint64_t n) { for (uint64_t k = 0; k < n; k++) (void)next(); }
    bool bulk() const { return words_ != nullptr; }

private:
    void open() {
        static int32_t parking[34];
        static bool parked = false;
        if (!parked) {  // a valid TYPE_3 state to park the global generator on while we hold the real one
            parking[0] = 3;
            for (int i = 1; i < 34; i++) parking[i] = (int32_t)((uint32_t)i * 1103515245u + 12345u);
            parked = true;
        }
        parking[0] = 3;  // rear = 0, type 3
        char* cur = setstate(reinterpret_cast<char*>(parking));  // saves the live positions into cur[0] and returns it
        if (!cur) return;
        int32_t* w = reinterpret_cast<int32_t*>(cur);
        if (w[0] % 5 != 3) { setstate(cur); return; }  // not the TYPE_3 generator: leave it alone
        words_ = w;
        r_ = w[0] / 5;
        f_ = (r_ + 3) % 31;
    }
    void close() {
        if (!words_) return;
        words_[0] = r_ * 5 + 3;
        setstate(reinterpret_cast<char*>(words_));
        words_ = nullptr;
    }
    static bool selfTest() {
        // clone the live state, replay 64 draws on the clone with the recurrence, compare with rand(), then rewind
        static int32_t parking2[34];
        parking2[0] = 3;
        for (int i = 1; i < 34; i++) parking2[i] = (int32_t)((uint32_t)i * 69069u + 1u);
        char* cur = setstate(reinterpret_cast<char*>(parking2));
        if (!cur) return false;
        int32_t* w = reinterpret_cast<int32_t*>(cur);
        int32_t saved[34];
        for (int i = 0; i < 32; i++) saved[i] = w[i];
        bool ok = (w[0] % 5 == 3);
        setstate(cur);
        if (!ok) return false;
        int32_t clone[32];
        for (int i = 0; i < 32; i++) clone[i] = saved[i];
        int r = clone[0] / 5, f = (r + 3) % 31;
        for (int k = 0; k < 64 && ok; k++) {
            int32_t* st = clone + 1;
            st[f] = (int32_t)((uint32_t)st[f] + (uint32_t)st[r]);
            int expect = (int)(((uint32_t)st[f]) >> 1);
            if (++f == 31) f = 0;
            if (++r == 31) r = 0;
            ok = (rand() == expect);
        }
        // rewind the live generator to where it was
        cur = setstate(reinterpret_cast<char*>(parking2));
        w = reinterpret_cast<int32_t*>(cur);
        for (int i = 0; i < 32; i++) w[i] = saved[i];
        setstate(cur);
        return ok;
    }
    int32_t* words_ = nullptr;
    int r_ = 0, f_ = 0;
    static bool checked_, usable_;
};
bool RandWindow::checked_ = false;
bool RandWindow::usable_ = false;

// libc's generator state as the device-resident stream wants it: the 31 most recent raw values, oldest first.
// (glibc TYPE_3 only; the state array is reached through the public setstate() API, as above.)
int32_t g_parking3[34];
bool libcPeek(uint32_t h[31]) {
    g_parking3[0] = 3;
    for (int i = 1; i < 34; i++) g_parking3[i] = (int32_t)((uint32_t)i * 2654435761u + 7u);
    char* cur = setstate(reinterpret_cast<char*>(g_parking3));
    if (!cur) return false;
    int32_t* w = reinterpret_cast<int32_t*>(cur);
    const bool ok = (w[0] % 5 == 3);
    if (ok) {
        const int r = w[0] / 5, f = (r + 3) % 31;
        for (int i = 0; i < 31; i++) h[i] = (uint32_t)w[1 + (f + i) % 31];
    }
    setstate(cur);
    return ok;
}
bool libcPoke(const uint32_t h[31]) {
    g_parking3[0] = 3;
    char* cur = setstate(reinterpret_cast<char*>(g_parking3));
    if (!cur) return false;
    int32_t* w = reinterpret_cast<int32_t*>(cur);
    if (w[0] % 5 != 3) { setstate(cur); return false; }
    for (int i = 0; i < 31; i++) w[1 + (3 + i) % 31] = (int32_t)h[i];  // oldest value at ring position 3, rear index 0
    w[0] = 3;
    setstate(cur);
    return true;
}
}  // namespace

// ---- RandStream: look-ahead view of libc's rand() stream -------------------------------------------------------
// run() draws once per neuron per call from libc's generator (NeuCor.cpp:604-607), and where that stream continues
// depends on a count only the device knows at the end of the window (the hidden rand() calls, NeuCor.cpp:752).  The
// VALUES of the stream do not depend on that count, only the position does — so the raw stream and the positions whose
// draw is a background hit (draw % period == 0) are generated ahead, while the device is busy, and the per-window work on
// the critical path shrinks to walking the (rare) hits.  glibc's TYPE_3 generator is x[n] = x[n-31] + x[n-3] (mod 2^32),
// rand() = x[n] >> 1; its 31-word state is borrowed through setstate() for the duration of run() and handed back at the
// position actually consumed, so the application's own rand() calls continue exactly where the reference's would.  If the
// application drew from (or re-seeded) the generator in between, the look-ahead no longer matches and is rebuilt.
struct NeuCor::RandStream {
    // x[k + 31] = raw value of draw k (k counted from the buffer's origin); x[0..31) = the 31 values before draw 0.
    // A plain array (no zero-fill on growth), compacted at every hand-back so that it stays cache-resident.
    uint32_t* x = nullptr;
    std::size_t len = 0, cap = 0;  // values held / allocated (incl. the 31 history values)
    std::size_t pos = 0;           // next unconsumed draw
    std::vector<uint64_t> hits;    // ascending draw numbers whose rand() value is divisible by `period`
    std::size_t hitHead = 0;       // hits[0..hitHead) are known to lie before pos
    std::size_t scanned = 0;       // draws [0, scanned) have been tested
    int period = 0;
    uint64_t magic = 0;
    bool attached = false, usable = true, valid = false;
    int32_t* words = nullptr;      // libc's state block while attached
    // Optional generator thread (large networks): keeps the look-ahead `want` draws ahead of the consumer, in chunks, under
    // one mutex that every method below takes; the consumer generates for itself whatever is still missing when it asks.
    std::mutex mu;
    std::condition_variable cv;
    std::thread worker;
    bool workerOn = false, stop = false;
    std::size_t want = 0;          // draws [0, want) are wanted ahead of time

    ~RandStream() {
        if (workerOn) {
            { std::lock_guard<std::mutex> g(mu); stop = true; }
            cv.notify_all();
            worker.join();
        }
        delete[] x;
    }
    std::size_t ready() const { return period > 1 ? std::min(len >= 31 ? len - 31 : 0, scanned) : (len >= 31 ? len - 31 : 0); }
    void run_worker() {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [&] { return stop || (valid && ready() < want); });
            if (stop) return;
            generate_l(std::min(want, ready() + (std::size_t)(1u << 15)));
            lk.unlock();  // let the consumer in between chunks
            std::this_thread::yield();
            lk.lock();
        }
    }
    void reserve(std::size_t n) {
        if (n <= cap) return;
        std::size_t c = std::max<std::size_t>(n + n / 2, 1u << 16);
        uint32_t* y = new uint32_t[c];
        if (len) memcpy(y, x, len * sizeof(uint32_t));
        delete[] x;
        x = y; cap = c;
    }
    bool attach() {
        std::lock_guard<std::mutex> g(mu);
        static int32_t parking[34];
        parking[0] = 3;
        for (int i = 1; i < 34; i++) parking[i] = (int32_t)((uint32_t)i * 1103515245u + 12345u);
        { RandWindow probe; if (!probe.bulk()) { usable = false; return false; } }  // self-test of the recurrence against rand()
        char* cur = setstate(reinterpret_cast<char*>(parking));  // saves the live positions into cur[0] and returns it
        if (!cur) { usable = false; return false; }
        int32_t* w = reinterpret_cast<int32_t*>(cur);
        if (w[0] % 5 != 3) { setstate(cur); usable = false; return false; }
        words = w;
        const int r = w[0] / 5, f = (r + 3) % 31;
        uint32_t h[31];  // ring order from the oldest value: x[n-31] .. x[n-1]
        for (int i = 0; i < 31; i++) h[i] = (uint32_t)w[1 + (f + i) % 31];
        bool same = valid && len >= pos + 31;
        for (int i = 0; i < 31 && same; i++) same = x[pos + i] == h[i];
        if (!same) {
            reserve(31);
            memcpy(x, h, sizeof(h));
            len = 31; pos = 0; hits.clear(); hitHead = 0; scanned = 0; valid = true; want = 0;
        }
        attached = true;
        return true;
    }
    void detach() {
        std::lock_guard<std::mutex> g(mu);
        if (!attached) return;
        // libc continues at `pos`: oldest value at ring position 3, rear index 0
        for (int i = 0; i < 31; i++) words[1 + (3 + i) % 31] = (int32_t)x[pos + i];
        words[0] = 3;
        setstate(reinterpret_cast<char*>(words));
        words = nullptr; attached = false;
        if (pos > (1u << 16)) {  // drop the consumed part
            memmove(x, x + pos, (len - pos) * sizeof(uint32_t));
            len -= pos;
            std::size_t k = 0;
            for (std::size_t i = hitHead; i < hits.size(); i++) if (hits[i] >= pos) hits[k++] = hits[i] - pos;
            hits.resize(k); hitHead = 0;
            scanned = scanned > pos ? scanned - pos : 0;
            want = want > pos ? want - pos : 0;
            pos = 0;
        }
    }
    void setPeriod(int p) {
        std::lock_guard<std::mutex> g(mu);
        if (p == period) return;
        period = p; magic = ~0ull / (uint64_t)p + 1ull;
        hits.clear(); hitHead = 0; scanned = pos;
    }
    // draws [0, upto) generated and tested for hits
    void generate_l(std::size_t upto) {  // (mu held)
        const uint64_t M = magic;
        // rand() = x >> 1 is a multiple of P only if its low ctz(P) bits are zero — a one-instruction screen that lets all but
        // 1 in 2^ctz(P) draws through; then (v * ceil(2^64/P)) wraps below ceil(2^64/P) exactly for the multiples of P
        const int tz = period > 1 ? __builtin_ctz((unsigned)period) : 0;
        const uint32_t low = ((1u << (tz > 8 ? 8 : tz)) - 1u) << 1;
        if (len < upto + 31) {
            reserve(upto + 31);
            uint32_t* y = x;
            std::size_t n = len;
            if (period > 1 && scanned + 31 == len) {
                // generate and test in one pass while the value is in a register (dependency distance 3: ~1 ns a draw)
                for (; n < upto + 31; n++) {
                    const uint32_t v = y[n - 31] + y[n - 3];
                    y[n] = v;
                    if (__builtin_expect((v & low) == 0u, 0) && (uint64_t)(v >> 1) * M < M) hits.push_back(n - 31);
                }
                scanned = upto;
            } else {
                for (; n < upto + 31; n++) y[n] = y[n - 31] + y[n - 3];
            }
            len = upto + 31;
        }
        if (period > 1 && scanned < upto) {  // draws generated earlier (another period, or ahead of the scan)
            const uint32_t* d = x + 31;
            for (std::size_t k = scanned; k < upto; k++)
                if ((d[k] & low) == 0u && (uint64_t)(d[k] >> 1) * M < M) hits.push_back(k);
            scanned = upto;
        }
    }
    // Asks for `ahead` draws beyond the current position.  With the generator thread this only posts the request.
    void prefetch(std::size_t ahead, bool threaded) {
        std::unique_lock<std::mutex> lk(mu);
        want = std::max(want, pos + ahead);
        if (!threaded) { generate_l(want); return; }
        if (!workerOn) { workerOn = true; worker = std::thread([this] { run_worker(); }); }
        lk.unlock();
        cv.notify_one();
    }
    int32_t value(std::size_t k) {
        std::lock_guard<std::mutex> g(mu);
        if (len < k + 32) generate_l(k + 1 + 4096);
        return (int32_t)(x[k + 31] >> 1);
    }
    void advance(uint64_t n) {
        std::lock_guard<std::mutex> g(mu);
        pos += n;
        if (len < pos + 31) generate_l(pos);
    }
    // first hit at or after draw `from` and before `limit`, or `limit`
    std::size_t nextHit(std::size_t from, std::size_t limit) {
        std::lock_guard<std::mutex> g(mu);
        if (scanned < limit) generate_l(limit);
        while (hitHead < hits.size() && hits[hitHead] < from) hitHead++;
        return (hitHead < hits.size() && hits[hitHead] < limit) ? (std::size_t)hits[hitHead] : limit;
    }
};

NeuCor::NeuCor(int n_neurons) {  // NeuCor.cpp:17-42
    runSpeed = 1.0;
    runAll = false;
    learningRate = 1.0;
    presynapticTraceDecay = 0.75;
    postsynapticTraceDecay = 0.65;
    presynapticFactor = 0.13;
    postsynapticFactor = 0.30;
    neurons.owner_ = this;

    totalGenNeurons = n_neurons;
    for (int n = 0; n < n_neurons; n++) {
        coord3 d;
        d.setNAN();
        createNeuron(d);
    }
    totalGenNeurons = 0;
    if (n_neurons > 0) makeConnections();
}

NeuCor::~NeuCor() {
    if (engine_) nc_destroy(engine_);
    delete rs_;
}

void NeuCor::check(int rc, const char* what) {
    if (rc == NC_OK) return;
    std::string msg = std::string(what) + ": " + (engine_ ? nc_last_error(engine_) : nc_global_error());
    throw std::runtime_error(msg);
}

float NeuCor::getTime() const { return currentTime; }
std::size_t NeuCor::getNeuronCount() const { return positions.size(); }
std::size_t NeuCor::synapseCount() const {
    if (dev_.set) return dev_.S;
    if (engine_ || imported_) return pre_.size();
    std::size_t s = 0;
    for (auto& o : out_) s += o.size();
    return s;
}

void NeuCor::createNeuron(coord3 position) {  // NeuCor.cpp:154-186 (SPAWN_SPHERE branch)
    if (engine_ || imported_) throw std::logic_error("NeuCor::createNeuron: the network is frozen once it is on the device");
    float spawnSize = 2.0;
    if (position.x != position.x) {
        if (totalGenNeurons != 0) spawnSize = powf(totalGenNeurons / (1.3333 * 3.1459 * 8), 0.33333) * 2.0;
        do {
            position.x = (randomUnit() - 0.5f) * spawnSize;
            position.y = (randomUnit() - 0.5f) * spawnSize;
            position.z = (randomUnit() - 0.5f) * spawnSize;
        } while ((double)position.x * (double)position.x + (double)position.y * (double)position.y +
                     (double)position.z * (double)position.z >
                 (spawnSize / 2.0) * (spawnSize / 2.0));
    }
    positions.push_back(position);
    potAct.push_back(-70.0f);  // Neuron ctor: setPotential(baselevel), setActivity(0)  NeuCor.cpp:389,394
    potAct.push_back(0.0f);
    out_.emplace_back();
    neurons.d_.emplace_back();
    neurons.d_.back().parentNet = this;
    neurons.d_.back().ownID = positions.size() - 1;
}

// ---- object view (NeuCor.h:117-125) ----------------------------------------------------------------------------
static void viewAddSynapse(std::deque<Neuron>& d, std::size_t from, std::size_t to, float weight, float length, bool inhibitory, uint64_t slot,
                           void (*set)(Synapse&, std::size_t, std::size_t, float, float, bool, uint64_t)) {
    d[from].outSynapses.emplace_back();
    set(d[from].outSynapses.back(), from, to, weight, length, inhibitory, slot);
    d[to].inSynapses.emplace(from, to);
}
void NeuCor::viewSet(Synapse& s, std::size_t from, std::size_t to, float weight, float length, bool inhibitory, uint64_t slot) {
    s.pN = from; s.tN = to; s.weight = weight; s.length = length; s.inhibitory = inhibitory; s.slot_ = slot;
}
void NeuCor::buildImportedView() {  // an imported network has no creation order: every neuron's out-synapses by ascending target
    neurons.d_.clear();
    const std::size_t N = positions.size();
    if (dev_.set || world_ > 1 || pre_.size() > objectViewLimit) return;
    neurons.d_.resize(N);
    for (std::size_t i = 0; i < N; i++) { neurons.d_[i].parentNet = this; neurons.d_[i].ownID = i; }
    for (std::size_t q = 0; q < N; q++)
        for (uint64_t k = rowptr_[q]; k < rowptr_[q + 1]; k++)
            viewAddSynapse(neurons.d_, pre_[k], q, weight_[k], length_[k], flag_[k] != 0, k, &NeuCor::viewSet);
    viewDirty_ = true;
}
Neuron* NeuCor::getNeuron(std::size_t ID) { return &neurons.at(ID); }
const Neuron* NeuCor::getNeuron(std::size_t ID) const { return &const_cast<NeuCor*>(this)->neurons.at(ID); }
Synapse* NeuCor::getSynapse(std::size_t fromID, std::size_t toID) {  // NeuCor.cpp:244-250
    Neuron* n = getNeuron(fromID);
    for (auto& s : n->outSynapses)
        if (s.tN == toID) return &s;
    return nullptr;
}
const Synapse* NeuCor::getSynapse(std::size_t fromID, std::size_t toID) const { return const_cast<NeuCor*>(this)->getSynapse(fromID, toID); }
Synapse* NeuCor::getSynapse(std::pair<std::size_t, std::size_t> ID) { return getSynapse(ID.first, ID.second); }
const Synapse* NeuCor::getSynapse(std::pair<std::size_t, std::size_t> ID) const { return getSynapse(ID.first, ID.second); }

// Device -> object view.  Nothing happens unless something ran since the last refresh.  Sharded engines keep no view.
void NeuCor::refreshObjects() {
    if (!viewDirty_ || !engine_ || world_ > 1 || neurons.d_.empty()) return;
    viewDirty_ = false;
    syncState();
    const std::size_t N = positions.size(), S = pre_.size();
    d2hBytes_ += N * 12 + S * 24;
    for (std::size_t i = 0; i < N; i++) neurons.d_[i].lastFire = lastFireMirror_[i];
    if (!S) return;
    std::vector<float> w(S), arrive(S), lastArr(S), lastStart(S), prePot(S), postPot(S);
    check(nc_read_synapses(engine_, w.data(), arrive.data(), nullptr, lastArr.data(), lastStart.data()), "nc_read_synapses");
    check(nc_read_synapse_pots(engine_, currentTime, prePot.data(), postPot.data()), "nc_read_synapse_pots");
    for (auto& n : neurons.d_)
        for (auto& sy : n.outSynapses) {
            const uint64_t k = sy.slot_;
            sy.weight = w[k]; sy.AP_fireTime = arrive[k]; sy.lastSpikeArrival = lastArr[k]; sy.lastSpikeStart = lastStart[k];
            sy.prePot_ = prePot[k]; sy.postPot_ = postPot[k];
        }
}

void NeuCor::createSynapse(std::size_t toID, std::size_t fromID, float weight) {  // NeuCor.cpp:187-195
    if (engine_ || imported_) throw std::logic_error("NeuCor::createSynapse: the network is frozen once it is on the device");
    auto& outs = out_.at(fromID);
    for (auto& t : outs)
        if (t.to == toID) return;  // silently ignores duplicates
    coord3 n1 = positions.at(toID);
    coord3 n2 = positions.at(fromID);
    // the Synapse ctor still draws its random weight and sign before setWeight overrides them (NeuCor.cpp:471-473)
    (void)randomUnit();
    (void)randomUnit();
    OutSyn s;
    s.to = (uint32_t)toID;
    s.weight = weight;
    s.flag = weight < 0.0;
    s.length = n2.getDist(n1);
    outs.push_back(s);
    viewAddSynapse(neurons.d_, fromID, toID, s.weight, s.length, s.flag != 0, 0, &NeuCor::viewSet);
}

void NeuCor::makeConnections() {  // NeuCor.cpp:89-93 → Neuron::makeConnections NeuCor.cpp:418-444
    if (engine_ || imported_) throw std::logic_error("NeuCor::makeConnections: the network is frozen once it is on the device");
    const std::size_t N = positions.size();
    // Uniform grid with unit cells: candidates come from the 27 surrounding cells and are visited in ascending ID,
    // which is the order the reference's O(N^2) loop meets them in (and so the order rand() is consumed in).
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (auto& p : positions) {
        lo[0] = std::min(lo[0], p.x); lo[1] = std::min(lo[1], p.y); lo[2] = std::min(lo[2], p.z);
        hi[0] = std::max(hi[0], p.x); hi[1] = std::max(hi[1], p.y); hi[2] = std::max(hi[2], p.z);
    }
    auto cellOf = [&](float v, int a) { return (long)std::floor((double)v - (double)lo[a]); };
    long dim[3];
    for (int a = 0; a < 3; a++) dim[a] = N ? cellOf(hi[a], a) + 1 : 1;
    std::unordered_map<long, std::vector<uint32_t>> grid;
    auto keyOf = [&](long cx, long cy, long cz) { return (cx * dim[1] + cy) * dim[2] + cz; };
    for (std::size_t i = 0; i < N; i++) grid[keyOf(cellOf(positions[i].x, 0), cellOf(positions[i].y, 1), cellOf(positions[i].z, 2))].push_back((uint32_t)i);
    std::vector<uint32_t> cand;
    for (std::size_t n = 0; n < N; n++) {
        coord3 nPos = positions[n];
        long cx = cellOf(nPos.x, 0), cy = cellOf(nPos.y, 1), cz = cellOf(nPos.z, 2);
        cand.clear();
        for (long dx = -1; dx <= 1; dx++)
            for (long dy = -1; dy <= 1; dy++)
                for (long dz = -1; dz <= 1; dz++) {
                    long x = cx + dx, y = cy + dy, z = cz + dz;
                    if (x < 0 || y < 0 || z < 0 || x >= dim[0] || y >= dim[1] || z >= dim[2]) continue;
                    auto it = grid.find(keyOf(x, y, z));
                    if (it != grid.end()) cand.insert(cand.end(), it->second.begin(), it->second.end());
                }
        std::sort(cand.begin(), cand.end());
        auto& outs = out_[n];
        for (uint32_t i : cand) {
            float distance = nPos.getDist(positions[i]);
            if (distance < 1.0 && i != n) {
                bool allowed = true;
                for (auto& o : outs)
                    if (o.to == i) { allowed = false; break; }
                if (!allowed) continue;
                // Synapse ctor, NeuCor.cpp:463-486
                OutSyn s;
                s.to = i;
                s.weight = randomUnit() * 0.8f + 0.2f;
                if (randomUnit() < 0.2f) s.weight = -s.weight;
                s.flag = s.weight < 0.0;
                s.length = nPos.getDist(positions[i]);
                outs.push_back(s);
                viewAddSynapse(neurons.d_, n, i, s.weight, s.length, s.flag != 0, 0, &NeuCor::viewSet);
            }
        }
    }
}

void NeuCor::importNetwork(std::size_t n, const uint64_t* rowptr, const uint32_t* pre, const float* weight, const float* length,
                           const uint8_t* inhibitory, const float* xyz) {
    if (engine_) throw std::logic_error("NeuCor::importNetwork: the network is already on the device");
    positions.resize(n);
    potAct.assign(2 * n, 0.0f);
    for (std::size_t i = 0; i < n; i++) {
        if (xyz) positions[i] = coord3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        else positions[i].setNAN();
        potAct[2 * i] = -70.0f;
    }
    out_.clear();
    gridN_ = 0;
    rowptr_.assign(rowptr, rowptr + n + 1);
    uint64_t S = rowptr[n];
    pre_.assign(pre, pre + S);
    weight_.assign(weight, weight + S);
    length_.assign(length, length + S);
    flag_.assign(inhibitory, inhibitory + S);
    imported_ = true;
    neurons.owner_ = this;
    buildImportedView();
}

void NeuCor::importNetworkDevice(std::size_t n, uint64_t synapses, const uint64_t* d_rowptr, const uint32_t* d_pre, const float* d_weight,
                                 const float* d_length, const uint8_t* d_inhibitory) {
    if (engine_) throw std::logic_error("NeuCor::importNetworkDevice: the network is already on the device");
    positions.resize(n);
    potAct.assign(2 * n, 0.0f);
    for (std::size_t i = 0; i < n; i++) { positions[i].setNAN(); potAct[2 * i] = -70.0f; }
    out_.clear();
    dev_.rowptr = d_rowptr; dev_.pre = d_pre; dev_.weight = d_weight; dev_.length = d_length; dev_.inh = d_inhibitory; dev_.S = synapses; dev_.set = true;
    imported_ = true;
    neurons.d_.clear();  // (no object view of a device-resident network)
}

void NeuCor::setPositions(const float* xyz, std::size_t n) {
    if (n != positions.size()) throw std::out_of_range("NeuCor::setPositions: one position per neuron");
    for (std::size_t i = 0; i < n; i++) positions[i] = coord3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    gridN_ = 0;  // the near-list grid is rebuilt from the new positions
}

void NeuCor::setShard(int rank, int world) {
    if (engine_) throw std::logic_error("NeuCor::setShard: the network is already on the device");
    if (world < 1 || rank < 0 || rank >= world) throw std::out_of_range("NeuCor::setShard: bad rank/world");
    rank_ = rank; world_ = world;
}
void NeuCor::setCommId(const void* commId128) { memcpy(commId_, commId128, 128); haveCommId_ = true; }
void NeuCor::setExchange(int (*allgather)(void*, const void*, void*, uint64_t), void* ctx) { xchgFn_ = allgather; xchgCtx_ = ctx; }

void NeuCor::importShardDevice(std::size_t n, uint64_t localSynapses, const uint64_t* d_rowptr, const uint32_t* d_pre, const float* d_weight,
                               const float* d_length, const uint8_t* d_inhibitory) {
    importNetworkDevice(n, localSynapses, d_rowptr, d_pre, d_weight, d_length, d_inhibitory);
    shardImport_ = true;
}

void NeuCor::finalize() {
    if (engine_) return;
    const std::size_t N = positions.size();
    row0_ = (std::size_t)((uint64_t)N * (uint64_t)rank_ / (uint64_t)world_);
    nRows_ = (std::size_t)((uint64_t)N * (uint64_t)(rank_ + 1) / (uint64_t)world_) - row0_;
    if (!imported_) {  // rows = target, in-row ascending presynaptic ID: Neuron::inSynapses' std::map order (NeuCor.h:212)
        rowptr_.assign(N + 1, 0);
        for (auto& outs : out_)
            for (auto& s : outs) rowptr_[s.to + 1]++;
        for (std::size_t i = 0; i < N; i++) rowptr_[i + 1] += rowptr_[i];
        uint64_t S = rowptr_[N];
        pre_.resize(S); weight_.resize(S); length_.resize(S); flag_.resize(S);
        std::vector<uint64_t> fill(rowptr_.begin(), rowptr_.end() - 1);
        for (std::size_t p = 0; p < N; p++) {  // ascending p => each row ends up ascending in presynaptic ID
            std::size_t j = 0;
            for (auto& s : out_[p]) {
                uint64_t k = fill[s.to]++;
                pre_[k] = (uint32_t)p; weight_[k] = s.weight; length_[k] = s.length; flag_[k] = s.flag;
                if (p < neurons.d_.size() && j < neurons.d_[p].outSynapses.size()) neurons.d_[p].outSynapses[j].slot_ = k;  // (object view: same order as out_)
                j++;
            }
        }
    }
    viewDirty_ = true;
    nc_config cfg = {};
    cfg.device = deviceOrdinal;
    cfg.rank = rank_;
    cfg.world = world_;
    cfg.cand_smem = candidateSmem;
    nc_engine* e = nullptr;
    int rc = nc_create(&cfg, &e);
    if (rc != NC_OK) throw std::runtime_error(std::string("NeuCor: cannot create the CUDA engine: ") + nc_global_error());
    engine_ = e;
    if (dev_.set) {
        if (world_ > 1 && !shardImport_) throw std::logic_error("NeuCor: a sharded run takes its device-resident rows through importShardDevice");
        check(nc_upload_network_device(engine_, N, row0_, nRows_, dev_.rowptr, dev_.pre, dev_.weight, dev_.length, dev_.inh), "nc_upload_network_device");
        sLocal_ = dev_.S;
    } else if (world_ == 1) {
        check(nc_upload_network(engine_, N, 0, N, rowptr_.data(), pre_.data(), weight_.data(), length_.data(), flag_.data()), "nc_upload_network");
        sLocal_ = pre_.size();
        h2dBytes_ += (N + 1) * 8 + pre_.size() * 13;
    } else {  // this shard's rows of the host CSR, re-based to start at 0
        const uint64_t lo = rowptr_[row0_], hi = rowptr_[row0_ + nRows_];
        std::vector<uint64_t> rp(nRows_ + 1);
        for (std::size_t r = 0; r <= nRows_; r++) rp[r] = rowptr_[row0_ + r] - lo;
        check(nc_upload_network(engine_, N, row0_, nRows_, rp.data(), pre_.data() + lo, weight_.data() + lo, length_.data() + lo, flag_.data() + lo), "nc_upload_network");
        sLocal_ = hi - lo;
        h2dBytes_ += (nRows_ + 1) * 8 + sLocal_ * 13;
    }
    if (world_ > 1) {
        if (haveCommId_) check(nc_comm_init(engine_, reinterpret_cast<const nc_comm_id*>(commId_)), "nc_comm_init");
        else if (xchgFn_) check(nc_set_exchange(engine_, xchgFn_, xchgCtx_), "nc_set_exchange");
        else throw std::logic_error("NeuCor: a sharded run needs setCommId or setExchange before the first run()");
    }
    // the smallest delay of the WHOLE network bounds the window (every shard must split windows identically)
    minDelay_ = INFINITY;
    if (sLocal_) check(nc_min_delay(engine_, &minDelay_), "nc_min_delay");
    if (world_ > 1) {
        if (dev_.set) {
            if (!(globalMinDelay > 0.0f)) throw std::logic_error("NeuCor: set globalMinDelay (smallest 2*length over all shards) for device-resident shards");
            minDelay_ = globalMinDelay;
        } else if (globalMinDelay > 0.0f) {  // a shard restored from its checkpoint file: only its own rows are on this host
            minDelay_ = globalMinDelay;
        } else {
            for (float l : length_) minDelay_ = std::min(minDelay_, l * 2.0f);
        }
    }
    lastFireMirror_.assign(N, NAN);
    // the trace decays are per-object copies taken when neurons / synapses are constructed (NeuCor.cpp:369,464;
    // NeuCor.h:236,289): later writes to the NeuCor members do nothing in the reference, so they are latched here
    preDecayLatched_ = presynapticTraceDecay;
    postDecayLatched_ = postsynapticTraceDecay;
}

// ---- inputs and detectors ---------------------------------------------------------------------------------
// Neurons within `radius` of `c`, ascending ID — what InputFirer / VoltageDetector collect with a loop over all neurons
// (NeuCor.cpp:319-323, 352-356).  Large networks with many firers (C2-C4: N/250 of them) go through a uniform grid of
// unit cells instead of N distance evaluations per firer; the distance test itself is the same float expression.
void NeuCor::nearList(coord3 c, float radius, std::vector<uint32_t>& out) {
    const std::size_t N = positions.size();
    out.clear();
    bool finite = std::isfinite(c.x) && std::isfinite(c.y) && std::isfinite(c.z) && std::isfinite(radius);
    if (N < 4096 || !finite) {
        for (std::size_t n = 0; n < N; n++)
            if (positions[n].getDist(c) < radius) out.push_back((uint32_t)n);
        return;
    }
    if (gridN_ != N) {  // (re)build: neurons are only ever appended
        // Only neurons with finite coordinates go into the grid.  An imported network may carry no positions (NaN): the distance
        // to such a neuron is NaN (or inf), never < a finite radius, so it is near nothing — as in the reference — and costs nothing.
        gridCells_.clear();
        gridLo_[0] = gridLo_[1] = gridLo_[2] = INFINITY;
        auto fin = [](const coord3& p) { return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z); };
        for (auto& p : positions)
            if (fin(p)) { gridLo_[0] = std::min(gridLo_[0], p.x); gridLo_[1] = std::min(gridLo_[1], p.y); gridLo_[2] = std::min(gridLo_[2], p.z); }
        gridOk_ = std::isfinite(gridLo_[0]);  // false: no neuron has a position
        if (gridOk_)
            for (std::size_t i = 0; i < N; i++)
                if (fin(positions[i])) gridCells_[cellKey(positions[i].x, positions[i].y, positions[i].z)].push_back((uint32_t)i);
        gridN_ = N;
    }
    if (!gridOk_) return;
    const double r = (double)radius + 1e-3;  // cells that can hold a neuron within the radius (with slack for float rounding)
    const long x0 = (long)std::floor((double)c.x - r - gridLo_[0]), x1 = (long)std::floor((double)c.x + r - gridLo_[0]);
    const long y0 = (long)std::floor((double)c.y - r - gridLo_[1]), y1 = (long)std::floor((double)c.y + r - gridLo_[1]);
    const long z0 = (long)std::floor((double)c.z - r - gridLo_[2]), z1 = (long)std::floor((double)c.z + r - gridLo_[2]);
    for (long x = x0; x <= x1; x++)
        for (long y = y0; y <= y1; y++)
            for (long z = z0; z <= z1; z++) {
                auto it = gridCells_.find(packKey(x, y, z));
                if (it == gridCells_.end()) continue;
                for (uint32_t n : it->second)
                    if (positions[n].getDist(c) < radius) out.push_back(n);
            }
    std::sort(out.begin(), out.end());
}
uint64_t NeuCor::packKey(long x, long y, long z) {
    return ((uint64_t)(x + (1L << 20)) << 42) ^ ((uint64_t)(y + (1L << 20)) << 21) ^ (uint64_t)(z + (1L << 20));
}
uint64_t NeuCor::cellKey(float x, float y, float z) const {
    return packKey((long)std::floor((double)x - gridLo_[0]), (long)std::floor((double)y - gridLo_[1]), (long)std::floor((double)z - gridLo_[2]));
}

void NeuCor::setInputRateArray(float inputs[], unsigned inputCount, coord3 inputPositions[], float inputRadius[]) {  // NeuCor.cpp:46-64
    inputArray = inputs;
    inputArraySize = inputCount;
    int change = (int)inputArraySize - (int)inputHandler.size();
    if (0 < change) {
        for (int i = 0; i < change; i++) {
            InputFirer f;  // InputFirer ctor, NeuCor.cpp:305-324
            f.lastFire = 0.0f;
            f.enabled = true;
            if (inputPositions != NULL) { f.a = inputPositions[i]; f.radius = inputRadius[i]; }
            else { f.radius = 1.0f; f.a.setNAN(); }
            if (!(f.a.x == f.a.x)) {
                f.a.x = (randomUnit() - 0.5f) * 5.f; f.a.y = (randomUnit() - 0.5f) * 5.f; f.a.z = (randomUnit() - 0.5f) * 5.f;
            }
            nearList(f.a, f.radius, f.near);
            inputHandler.push_back(std::move(f));
        }
    } else if (change < 0) {
        for (int i = 0; i < -change; i++) inputHandler.pop_back();
    }
}
void NeuCor::addInputOffset(unsigned inputID, float t) { inputHandler.at(inputID).lastFire += t; }
void NeuCor::setInputEnabled(unsigned inputID, bool en) { inputHandler.at(inputID).enabled = en; }
void NeuCor::setInputNear(unsigned inputID, const uint32_t* ids, std::size_t n) { inputHandler.at(inputID).near.assign(ids, ids + n); }
void NeuCor::setInputLastFire(unsigned inputID, float t) { inputHandler.at(inputID).lastFire = t; }
std::vector<float> NeuCor::inputLastFire() const {
    std::vector<float> r;
    for (auto& f : inputHandler) r.push_back(f.lastFire);
    return r;
}
std::vector<std::vector<uint32_t>> NeuCor::inputNear() const {
    std::vector<std::vector<uint32_t>> r;
    for (auto& f : inputHandler) r.push_back(f.near);
    return r;
}

void NeuCor::setDetectors(unsigned detectorNumber, coord3 detectorPositions[], float detectorRadius[]) {  // NeuCor.cpp:70-77,347-357
    for (unsigned i = 0; i < detectorNumber; i++) {
        VoltageDetector d;
        if (detectorPositions != NULL) { d.a = detectorPositions[i]; d.radius = detectorRadius[i]; }
        else { d.radius = 1.0f; d.a.setNAN(); }
        if (!(d.a.x == d.a.x)) {
            d.a.x = (randomUnit() - 0.5f) * 3.f; d.a.y = (randomUnit() - 0.5f) * 3.f; d.a.z = (randomUnit() - 0.5f) * 3.f;
        }
        nearList(d.a, d.radius, d.near);
        voltageDetectors.push_back(std::move(d));
    }
}

float NeuCor::getDetectorVoltage(unsigned ID) {  // VoltageDetector::getVoltage, NeuCor.cpp:359-366
    VoltageDetector& d = voltageDetectors.at(ID);
    finalize();
    if (world_ > 1) throw std::logic_error("NeuCor::getDetectorVoltage: not available on a sharded network (use runSwept)");
    if (d.near.empty()) return 0.0f / (float)d.near.size();  // nothing within the radius: nothing is run, avgV/0 = NaN (NeuCor.cpp:360-365)
    uint64_t hidden = 0;
    nc_step_stats st;
    check(nc_run_neurons(engine_, currentTime, d.near.data(), (uint32_t)d.near.size(), &hidden, &st), "nc_run_neurons");
    if (hidden) { RandWindow rw; rw.skip(hidden); }
    float out = 0.0f;
    check(nc_detector_mean(engine_, d.near.data(), (uint32_t)d.near.size(), &out), "nc_detector_mean");
    viewDirty_ = true;  // (the read ran its neurons: the mirrors follow, as the reference's live objects do)
    if (mirrorAfterRun) { syncState(); d2hBytes_ += nRows_ * 12; }
    return out;
}
std::vector<float> NeuCor::getDetectorVoltages() {
    std::vector<float> voltages;
    for (unsigned i = 0; i < voltageDetectors.size(); i++) voltages.push_back(getDetectorVoltage(i));
    return voltages;
}

// ---- the step ---------------------------------------------------------------------------------------------
// InputFirer::schedule, NeuCor.cpp:333-345 — emits one event per (fire time, near neuron)
void NeuCor::scheduleInput(unsigned i, float deltaT, float frequency, std::vector<nc_event>& ev) {
    InputFirer& f = inputHandler[i];
    if (frequency == 0 || !f.enabled) return;
    if (f.lastFire != f.lastFire) f.lastFire = 0;
    float currentT = currentTime;
    for (float fireTime = f.lastFire + 1000.0 / frequency; fireTime < currentT + deltaT; fireTime += 1000.0 / frequency) {
        if (currentT < fireTime) {
            float stime = currentTime + (fireTime - currentT);  // queueSimulation, NeuCor.cpp:227-229
            for (uint32_t n : f.near) ev.push_back(nc_event{n, stime, 0u, i});
            f.lastFire = fireTime;
        }
    }
}

void NeuCor::window(float t0, float t1, int flags, std::vector<nc_event>& ev) {
    const bool devRand = devRandThisRun_;
    uint64_t hidden = 0;
    nc_step_stats st;
    if (world_ > 1) {  // every process schedules the whole network's events (same rand() stream); a shard takes those of its rows
        std::size_t k = 0;
        for (auto& x : ev)
            if (x.neuron >= row0_ && x.neuron < row0_ + nRows_) ev[k++] = x;
        ev.resize(k);
    }
    check(nc_step_launch(engine_, t0, t1, flags, ev.data(), (uint32_t)ev.size()), "nc_step");
    // while the device runs: extend the look-ahead of the rand() stream to cover this window's hidden calls (a guess from
    // the last window) and the next run()'s per-neuron draws
    if (rs_ && rs_->attached) rs_->prefetch((std::size_t)(2 * lastHidden_ + positions.size() + positions.size() / 64 + 8192), randThread_);
    check(nc_step_collect(engine_, &hidden, &st), "nc_step");
    h2dBytes_ += ev.size() * sizeof(nc_event);
    d2hBytes_ += 16 + 8 * sizeof(uint64_t);
    // the rand() calls hidden in synapticPlasticity's short-circuit (NeuCor.cpp:752): only their number matters
    lastHidden_ = hidden;
    if (devRand) d2hBytes_ += 31 * 4 + 16;  // the stream state and the background control block come back with the counters
    if (hidden && !devRand) {  // (the device-resident stream has moved on by itself)
        if (rs_ && rs_->attached) rs_->advance(hidden);
        else { RandWindow rw; rw.skip(hidden); }
    }
    if (recordFires) {  // (neuron, time) of every Neuron::fire of this window, all shards — the GUI's raster source (Renderer.cpp:1856-1862)
        uint32_t n = 0;
        const std::size_t at = firesNeuron_.size();
        check(nc_read_fires(engine_, 0, nullptr, nullptr, &n), "nc_read_fires");
        firesNeuron_.resize(at + n); firesTime_.resize(at + n);
        if (n) check(nc_read_fires(engine_, n, firesNeuron_.data() + at, firesTime_.data() + at, &n), "nc_read_fires");
        d2hBytes_ += (uint64_t)n * 8;
    }
    lastStats_.fires += st.fires; lastStats_.deliveries += st.deliveries; lastStats_.loadsAccepted += st.loads_accepted;
    lastStats_.loadsDropped += st.loads_dropped; lastStats_.plasticityCalls += st.plasticity_calls; lastStats_.hiddenRand += st.hidden_rand_calls;
    lastStats_.neuronRuns += st.neuron_runs; lastStats_.activeVisits += st.active_visits;
}

float NeuCor::stepInternal(bool sweep) {  // NeuCor::run, NeuCor.cpp:583-617
    assert(0 <= runSpeed);
    if (runSpeed <= 0.0f) return 0.0f;
    finalize();
    check(nc_set_plasticity(engine_, learningRate, presynapticFactor, postsynapticFactor, preDecayLatched_, postDecayLatched_), "nc_set_plasticity");
    lastStats_ = StepStats{};
    const std::size_t N = positions.size();
    // borrow libc's generator for the duration of this call (handed back, at the position consumed, on every exit path)
    if (!rs_) {
        rs_ = new RandStream();
        // NC_RAND_THREAD=1 moves the look-ahead generation to a helper thread.  Off by default: on the B200 box it did not beat
        // generating under the device's shadow in the calling thread (C3 e2e 3.94 vs 3.72 ms per step), see profiles/README.md.
        const char* env = getenv("NC_RAND_THREAD");
        randThread_ = env ? atoi(env) != 0 : false;
    }
    // Device-resident rand() stream: the per-neuron background draws and the hidden calls of the
    // plasticity never touch the host; libc's state is handed to the device when the application has drawn from it since
    // the last run() and put back afterwards.  NC_HOST_RAND=1 keeps the draws on the host (the look-ahead stream below).
    const int bgPeriod = std::max(1, static_cast<int>(600.0f / runSpeed));
    devRandThisRun_ = false;
    if (deviceRand_ < 0) { const char* env = getenv("NC_HOST_RAND"); deviceRand_ = (env && atoi(env) != 0) ? 0 : 1; }
    if (deviceRand_ == 1 && bgPeriod >= 2 && N > 0) {
        { RandWindow probe; if (!probe.bulk()) deviceRand_ = 0; }  // self-test of the recurrence against rand()
        uint32_t h[31];
        if (deviceRand_ == 1 && libcPeek(h)) {
            if (!randMirrorValid_ || memcmp(h, randMirror_, sizeof(h)) != 0) {
                check(nc_rand_set_state(engine_, h), "nc_rand_set_state");
                h2dBytes_ += sizeof(h);
            }
            check(nc_background_draw(engine_, currentTime, runSpeed, (uint32_t)bgPeriod, (uint64_t)N), "nc_background_draw");
            devRandThisRun_ = true;
        }
    }
    if (!devRandThisRun_) { check(nc_background_clear(engine_), "nc_background_clear"); randMirrorValid_ = false; }
    struct Borrow {
        RandStream* r;
        ~Borrow() { if (r) r->detach(); }
    } borrow{(!devRandThisRun_ && rs_->usable && bgPeriod > 1 && rs_->attach()) ? rs_ : nullptr};
    events_.clear();
    firesNeuron_.clear(); firesTime_.clear();
    for (unsigned i = 0; i < inputHandler.size(); i++) {
        const float inputFrequency = inputArray != nullptr && i < inputArraySize ? inputArray[i] : 0.0f;
        scheduleInput(i, runSpeed, inputFrequency, events_);
    }
    // background firing, NeuCor.cpp:604-607
    const std::size_t bgBegin = events_.size();
    if (!devRandThisRun_) {
        const int backgroundFirePeriod = std::max(1, static_cast<int>(600.0f / runSpeed));
        if (rs_ && rs_->attached && backgroundFirePeriod > 1) {
            // neuron i tests draw number p + i + 2 * (hits before it); a hit consumes the next two draws (NeuCor.cpp:606)
            RandStream& rs = *rs_;
            rs.setPeriod(backgroundFirePeriod);
            std::size_t p = rs.pos, i = 0;
            while (i < N) {
                const std::size_t j = rs.nextHit(p, p + (N - i));
                if (j >= p + (N - i)) { p += N - i; break; }
                i += j - p;
                uint32_t n = (uint32_t)(rs.value(j + 1) % N);
                float t = (static_cast<float>(rs.value(j + 2)) / static_cast<float>(RAND_MAX)) * runSpeed;
                events_.push_back(nc_event{n, currentTime + t, 2u, 0u});  // Neuron::scheduleFire, NeuCor.cpp:658-661
                p = j + 3;
                i += 1;
            }
            rs.advance(p - rs.pos);
        } else {
            RandWindow rw;  // the same draws rand() would return, without its per-call cost
            for (std::size_t i = 0; i < N; ++i) {
                if (rw.next() % backgroundFirePeriod == 0) {
                    uint32_t n = (uint32_t)(rw.next() % N);
                    float t = (static_cast<float>(rw.next()) / static_cast<float>(RAND_MAX)) * runSpeed;
                    events_.push_back(nc_event{n, currentTime + t, 2u, 0u});  // Neuron::scheduleFire, NeuCor.cpp:658-661
                }
            }
        }
        // scheduledFireTime keeps the LAST value drawn for a neuron
        for (std::size_t k = events_.size(); k-- > bgBegin;) {
            bool later = false;
            for (std::size_t m = k + 1; m < events_.size() && !later; m++) later = events_[m].neuron == events_[k].neuron;
            if (!later) events_[k].index_or_flags = 1u;
        }
    }
    std::stable_sort(events_.begin(), events_.end(), [](const nc_event& a, const nc_event& b) { return a.neuron < b.neuron; });

    const float t0 = currentTime;
    const float targetTime = currentTime + runSpeed;
    const float limit = std::min(minDelay_, 2.0f);
    int startFlag = runAll ? NC_SWEEP_START : 0;
    if (targetTime - t0 < limit) {
        window(t0, targetTime, startFlag | (sweep ? NC_SWEEP_END : 0), events_);
    } else {
        // The window must be shorter than the smallest synaptic delay: split it. Boundaries are invisible to the
        // semantics (no neuron is run at them); only the last sub-window carries the end sweep, the first one that is
        // executed carries the start sweep and the events at exactly t0.  Boundaries are forced to advance by at least
        // one float ulp; when even one ulp is no shorter than the limit (simulated time beyond ~2^23 * minDelay ms) the
        // float clock can no longer order a spike's departure and arrival and the engine refuses to go on.
        int pieces = (int)std::ceil((double)(targetTime - t0) / ((double)limit * 0.5)) + 1;
        float a = t0;
        bool first = true;
        for (int k = 1; k <= pieces && a < targetTime; k++) {
            float b = (k == pieces) ? targetTime : (float)((double)t0 + ((double)targetTime - (double)t0) * k / pieces);
            if (!(b > a)) b = std::nextafterf(a, INFINITY);
            if (b > targetTime) b = targetTime;
            if (!(b - a < limit))
                throw std::runtime_error("NeuCor::run: simulated time too large — one float ulp of the clock is no shorter than the smallest synaptic delay");
            const bool last = !(b < targetTime);
            winEvents_.clear();
            for (auto& e : events_)
                if ((first ? e.time >= a : e.time > a) && e.time <= b) winEvents_.push_back(e);
            window(a, b, (first ? startFlag : 0) | ((sweep && last) ? NC_SWEEP_END : 0), winEvents_);
            first = false;
            a = b;
        }
    }
    if (devRandThisRun_) {  // libc's generator continues where the reference's would
        uint32_t h[31];
        check(nc_rand_get_state(engine_, h), "nc_rand_get_state");
        libcPoke(h);
        memcpy(randMirror_, h, sizeof(h));
        randMirrorValid_ = true;
    }
    currentTime = targetTime;
    viewDirty_ = true;
    totalStats_.fires += lastStats_.fires; totalStats_.deliveries += lastStats_.deliveries; totalStats_.loadsAccepted += lastStats_.loadsAccepted;
    totalStats_.loadsDropped += lastStats_.loadsDropped; totalStats_.plasticityCalls += lastStats_.plasticityCalls; totalStats_.hiddenRand += lastStats_.hiddenRand;
    totalStats_.neuronRuns += lastStats_.neuronRuns; totalStats_.activeVisits += lastStats_.activeVisits;
    return 0.0f;
}

void NeuCor::run() {
    stepInternal(false);
    if (mirrorAfterRun && engine_ && !(runSpeed <= 0.0f)) { syncState(); d2hBytes_ += nRows_ * 12; }  // positions / potAct / lastFire as the renderer reads them
}

float NeuCor::runSwept() {
    stepInternal(true);
    if (!sweepReturnsMean || world_ > 1) {
        if (mirrorAfterRun && engine_ && !(runSpeed <= 0.0f)) { syncState(); d2hBytes_ += nRows_ * 12; }
        return 0.0f;
    }
    // mean potential of all neurons, summed in ID order in float — VoltageDetector::getVoltage, NeuCor.cpp:360-365
    syncState();
    d2hBytes_ += potAct.size() * 4;
    float avgV = 0;
    const std::size_t N = positions.size();
    for (std::size_t i = 0; i < N; i++) avgV += potAct[2 * i];
    return avgV / N;
}

// ---- state read-back --------------------------------------------------------------------------------------
void NeuCor::syncState() {
    if (!engine_) return;
    check(nc_read_neurons(engine_, potAct.data() + 2 * row0_, lastFireMirror_.data() + row0_, nullptr), "nc_read_neurons");
}
void NeuCor::readNeurons(float* pot, float* act, float* lastFire, float* lastRan) {
    finalize();
    // this shard's rows (all of them for world 1), in ID order starting at shardRow0()
    check(nc_read_neurons(engine_, potAct.data() + 2 * row0_, lastFire, lastRan), "nc_read_neurons");
    for (std::size_t i = 0; i < nRows_; i++) {
        if (pot) pot[i] = potAct[2 * (row0_ + i)];
        if (act) act[i] = potAct[2 * (row0_ + i) + 1];
    }
}
void NeuCor::readSynapses(float* weight, float* arrive, float* depol, float* lastArrival, float* lastStart) {
    finalize();
    check(nc_read_synapses(engine_, weight, arrive, depol, lastArrival, lastStart), "nc_read_synapses");
}
bool NeuCor::checkpointPeekRand(uint32_t x31[31]) {
    { RandWindow probe; if (!probe.bulk()) return false; }  // (the probe borrows libc's state while it lives)
    return libcPeek(x31);
}
bool NeuCor::checkpointPokeRand(const uint32_t x31[31]) { return libcPoke(x31); }
void NeuCor::stateSignature(uint64_t out6[6]) {
    finalize();
    check(nc_state_signature(engine_, out6), "nc_state_signature");
}
void NeuCor::resetActivities() {  // NeuCor.cpp:233-235
    finalize();
    check(nc_reset_activities(engine_, currentTime), "nc_reset_activities");
    viewDirty_ = true;
}

std::vector<NeuCor::NeuronSnapshot> NeuCor::getNeuronSnapshots() const {  // NeuCor.cpp:99-113
    const_cast<NeuCor*>(this)->syncState();
    std::vector<NeuronSnapshot> snapshots;
    snapshots.reserve(positions.size());
    for (std::size_t i = 0; i < positions.size(); i++) snapshots.push_back({i, positions[i], potAct[2 * i], potAct[2 * i + 1]});
    return snapshots;
}
std::vector<NeuCor::SynapseSnapshot> NeuCor::getSynapseSnapshots() const {  // NeuCor.cpp:115-134
    NeuCor* self = const_cast<NeuCor*>(this);
    self->finalize();
    if (dev_.set) throw std::logic_error("NeuCor::getSynapseSnapshots: not available for a network imported from device memory");
    const std::size_t S = pre_.size(), N = positions.size();
    std::vector<float> w(S), prePot(S), postPot(S);
    self->check(nc_read_synapses(engine_, w.data(), nullptr, nullptr, nullptr, nullptr), "nc_read_synapses");
    self->check(nc_read_synapse_pots(engine_, currentTime, prePot.data(), postPot.data()), "nc_read_synapse_pots");
    // the reference iterates neurons in ID order and each neuron's outSynapses in creation order; an imported network
    // has no creation order, so it is listed by (from, to)
    std::vector<SynapseSnapshot> snapshots;
    snapshots.reserve(S);
    std::vector<std::vector<std::pair<uint32_t, uint64_t>>> byFrom(N);
    for (std::size_t q = 0; q < N; q++)
        for (uint64_t k = rowptr_[q]; k < rowptr_[q + 1]; k++) byFrom[pre_[k]].push_back({(uint32_t)q, k});
    for (std::size_t p = 0; p < N; p++) {
        auto emit = [&](uint32_t to, uint64_t k) {
            snapshots.push_back({p, to, positions[p], positions[to], w[k], prePot[k], postPot[k], w[k] < 0.0f});
        };
        if (!imported_ && p < out_.size()) {
            for (auto& o : out_[p])
                for (auto& e : byFrom[p])
                    if (e.first == o.to) { emit(e.first, e.second); break; }
        } else {
            for (auto& e : byFrom[p]) emit(e.first, e.second);
        }
    }
    return snapshots;
}
std::vector<NeuCor::InputSnapshot> NeuCor::getInputSnapshots() const {  // NeuCor.cpp:136-152
    std::vector<InputSnapshot> snapshots;
    snapshots.reserve(inputHandler.size());
    for (std::size_t i = 0; i < inputHandler.size(); ++i) {
        const InputFirer& input = inputHandler.at(i);
        snapshots.push_back({i, input.a, input.radius, inputArray != nullptr && i < inputArraySize ? inputArray[i] : 0.0f, input.enabled});
    }
    return snapshots;
}
