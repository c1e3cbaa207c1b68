/* neucor_b200.h — C ABI of the B200-native simulation core for NeuroCorrelation's per-step
 * spiking-network update (libneucor_b200.so).
 *
 * The reference has no FFI/plugin layer: its boundary is the C++ class `NeuCor`
 * (/root/reference/src/NeuCor.h:36-138) exported by the static library `neurocorrelation_core`
 * (CMakeLists.txt:13-25).  The drop-in keeps that class on the host
 * (neurocorrelation_b200/host/NeuCor.h, same public members) and drives the GPU through the
 * entry points below — plain pointers and sizes, `int` status returns (0 = ok, negative = error,
 * text via nc_last_error), no exceptions and no torch types across the boundary, one calling
 * thread per handle, device memory owned by the handle, host arrays borrowed for the duration of
 * a call.  There is NO CPU fallback: every entry point fails with NC_ERR_NO_DEVICE when no CUDA
 * device is usable.
 *
 * Each entry point names the reference code it replaces (file:line under /root/reference/src).
 */
#ifndef NEUCOR_B200_H
#define NEUCOR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NC_OK 0
#define NC_ERR_INVALID (-1)    /* bad argument / precondition (message says which) */
#define NC_ERR_NO_DEVICE (-2)  /* no usable CUDA device — there is no CPU path */
#define NC_ERR_CUDA (-3)       /* a CUDA call failed */
#define NC_ERR_CAPACITY (-4)   /* a fixed-capacity device buffer overflowed (fires per step) */
#define NC_ERR_STATE (-5)      /* call order (e.g. step before upload) */

/* nc_step `sweep` flags */
#define NC_SWEEP_END 1    /* after the window, run every neuron at t1 in ascending ID (detector read / renderer) */
#define NC_SWEEP_START 2  /* before the window, run every neuron at t0 — NeuCor::runAll (NeuCor.cpp:595-597),
                             in ascending ID instead of the reference's heap order */

typedef struct nc_engine nc_engine;

typedef struct nc_config {
    int32_t device;            /* CUDA device ordinal */
    int32_t rank, world;       /* neuron-range shard of this engine (world = 1: whole network) */
    uint32_t fire_capacity;    /* max fire records of this shard per step (0 = default: 4*n_rows + 1024) */
    uint32_t cand_smem;        /* staged-slot pool per warp in shared memory, shared by the rows of a batch (0 = default 1024) */
    void* stream;              /* cudaStream_t to launch on (NULL = the engine creates its own) */
    uint32_t flag_capacity;    /* max slots per step that deliver or are cleared (0 = default: S/8 + 65536, at least 2^20) */
    uint32_t reserved;
} nc_config;

/* Host-generated event of one step (NeuCor::run's scheduling phase, NeuCor.cpp:599-607).
 * kind 0: input firer `index` fires `neuron` at `time` (InputFirer::run, NeuCor.cpp:326-331)
 * kind 2: neuron is run at `time` (Neuron::scheduleFire's queue entry, NeuCor.cpp:658-661);
 *         flags bit 0: `time` is the neuron's scheduledFireTime (the last one drawn this step). */
typedef struct nc_event {
    uint32_t neuron;
    float time;
    uint32_t kind;
    uint32_t index_or_flags;
} nc_event;

typedef struct nc_step_stats {
    uint64_t fires;             /* Neuron::fire calls (NeuCor.cpp:643) */
    uint64_t deliveries;        /* Synapse::run past its guard (NeuCor.cpp:719-721) */
    uint64_t loads_accepted;    /* Synapse::fire that took the slot (NeuCor.cpp:728-737) */
    uint64_t loads_dropped;     /* Synapse::fire returning early on a busy slot (NeuCor.cpp:728) */
    uint64_t plasticity_calls;  /* Synapse::synapticPlasticity (NeuCor.cpp:740) */
    uint64_t hidden_rand_calls; /* rand() evaluated by the short-circuit at NeuCor.cpp:752 */
    uint64_t neuron_runs;       /* Neuron::run with deltaT != 0 (NeuCor.cpp:626) */
    uint64_t active_visits;     /* in-synapse contributions added (NeuCor.cpp:695) */
} nc_step_stats;

/* Library-level: text of the last error of a failed nc_create (no handle exists then). */
const char* nc_global_error(void);
/* Number of usable CUDA devices (0 when there is none; never throws). */
int nc_device_count(void);

/* Replaces NeuCor::NeuCor's device-independent setup (NeuCor.cpp:17-30). */
int nc_create(const nc_config* cfg, nc_engine** out);
void nc_destroy(nc_engine* e);
const char* nc_last_error(const nc_engine* e);

/* Network upload — replaces the deque<Neuron>/vector<Synapse>/map containers (NeuCor.h:116,
 * 211-212) by a post-synaptic-sorted CSR: rows = target neurons [row0, row0+n_rows) owned by this
 * engine, in-row ascending presynaptic ID (the iteration order of Neuron::inSynapses that
 * charge_insynapses uses, NeuCor.cpp:690).  rowptr has n_rows+1 entries starting at 0.
 * `length` is Synapse::length (delay = length * 2.0f, NeuCor.cpp:485,733); `inhibitory` the flag
 * byte that selects the weight clamp (NeuCor.cpp:760-761).  Initial state is the reference's:
 * potential -70, lastFire NaN, idle slots, lastSpikeArrival -inf (NeuCor.cpp:376-394,469,484). */
int nc_upload_network(nc_engine* e, uint64_t n_global, uint64_t row0, uint64_t n_rows, const uint64_t* rowptr,
                      const uint32_t* pre, const float* weight, const float* length, const uint8_t* inhibitory);
/* Same, with the five arrays already resident on this engine's device (they are copied; the caller keeps ownership). */
int nc_upload_network_device(nc_engine* e, uint64_t n_global, uint64_t row0, uint64_t n_rows, const uint64_t* d_rowptr,
                             const uint32_t* d_pre, const float* d_weight, const float* d_length, const uint8_t* d_inhibitory);
/* Smallest synaptic delay (2*length) of the uploaded shard; a step window must be shorter. */
int nc_min_delay(const nc_engine* e, float* out);

/* NeuCor::learningRate / presynapticFactor / postsynapticFactor are read live on every plasticity
 * call (NeuCor.cpp:757-758); the trace decays are the per-object copies (NeuCor.cpp:369,464). */
int nc_set_plasticity(nc_engine* e, float learning_rate, float pre_factor, float post_factor, float pre_decay,
                      float post_decay);

/* One window (t0, t1] of NeuCor::run()'s hot loop (NeuCor.cpp:609-616) on the device: neuron pass
 * (Neuron::run/fire/transfer, NeuCor.cpp:619-714), fire-event exchange, synapse pass
 * (Synapse::fire/run/synapticPlasticity, NeuCor.cpp:718-764).  `events` are the host-scheduled
 * input-firer and background events with t0 <= time <= t1, sorted by (neuron, time).
 * sweep != 0: afterwards every neuron is run at t1 in ascending ID — what the reference's
 * VoltageDetector::getVoltage does to its `near` neurons (NeuCor.cpp:359-366) and what
 * runAll / the renderer rely on.  Requires t1 - t0 < nc_min_delay().
 * hidden_rand_calls: number of rand() evaluations the reference would have made inside
 * synapticPlasticity during this window (NeuCor.cpp:752); the host must advance rand() by it. */
int nc_step(nc_engine* e, float t0, float t1, int sweep, const nc_event* events, uint32_t n_events,
            uint64_t* hidden_rand_calls, nc_step_stats* stats_or_null);

/* The same window in two halves, so that the host can work while the device runs: nc_step_launch enqueues everything
 * (for world > 1 it blocks briefly inside the fire exchange) and returns; nc_step_collect waits and reports. */
int nc_step_launch(nc_engine* e, float t0, float t1, int sweep, const nc_event* events, uint32_t n_events);
int nc_step_collect(nc_engine* e, uint64_t* hidden_rand_calls, nc_step_stats* stats_or_null);

/* The side effect of VoltageDetector::getVoltage (NeuCor.cpp:361-362): run the listed neurons
 * (ascending IDs of this shard; NULL = all) at time `now`, including any fires this causes. */
int nc_run_neurons(nc_engine* e, float now, const uint32_t* ids, uint32_t n_ids, uint64_t* hidden_rand_calls,
                   nc_step_stats* stats_or_null);

/* Renderer-/snapshot-visible state (NeuCor.h:104-105, NeuCor.cpp:99-134, Renderer.cpp:655-699).
 * Any pointer may be NULL. potAct is the interleaved (potential, activity) array of this shard. */
int nc_read_neurons(nc_engine* e, float* potAct, float* lastFire, float* lastRan);
int nc_read_synapses(nc_engine* e, float* weight, float* arrive, float* depol, float* lastArrival, float* lastStart);
/* Neuron::activityStartTime / firings (NeuCor.h:239-240) — with nc_read_neurons and nc_read_synapses the complete dynamic state. */
int nc_read_neuron_counters(nc_engine* e, float* actStart, uint32_t* firings);
/* The shard's network as uploaded (rows of the CSR), read back from the device records. Any pointer may be NULL. */
int nc_read_network(nc_engine* e, uint64_t* rowptr, uint32_t* pre, float* length, uint8_t* inhibitory);
/* Checkpoint resume (the reference has no serialisation at all): the inverse of the readers above.
 * NULL keeps what is there.  After nc_write_synapses the event index is rebuilt from `arrive`. */
int nc_write_neurons(nc_engine* e, const float* potAct, const float* lastFire, const float* lastRan, const float* actStart, const uint32_t* firings);
int nc_write_synapses(nc_engine* e, const float* weight, const float* arrive, const float* depol, const float* lastArrival, const float* lastStart);
/* Fire events of the last step, all shards: (neuron, time); returns the total count in *count. */
int nc_read_fires(nc_engine* e, uint32_t capacity, uint32_t* neuron, float* time, uint32_t* count);
/* Synapse::getPrePot / getPostPot at time `now` for every synapse of the shard (NeuCor.cpp:547-567). */
int nc_read_synapse_pots(nc_engine* e, float now, float* prePot, float* postPot);
/* The same potentials written into caller-provided DEVICE (or host-mapped / graphics-interop) buffers of S floats each,
 * in-stream, no staging copy — the renderer's per-frame gather (Renderer.cpp:655-699). Either pointer may be NULL. */
int nc_synapse_pots_device(nc_engine* e, float now, float* d_prePot, float* d_postPot);
/* NeuCor_Renderer's "Statistics" panel reduced on the device (Renderer.cpp:1733-1876): the neuron-activity and the
 * synapse-weight distribution (span index = floor(spans * (x - range_min) / (range_max - range_min)), values outside the
 * range counted in *below / *above; an empty range leaves all bins 0 as the reference does), and one frame of the raster
 * plot by the GUI's own rule `now - lastFire < runSpeed` (ascending neuron IDs; *count may exceed capacity). */
int nc_render_activity_histogram(nc_engine* e, uint32_t spans, float range_min, float range_max, uint32_t* bins, uint32_t* below, uint32_t* above);
int nc_render_weight_histogram(nc_engine* e, uint32_t spans, float range_min, float range_max, uint32_t* bins, uint32_t* below, uint32_t* above);
int nc_render_raster(nc_engine* e, float now, float run_speed, uint32_t capacity, uint32_t* ids, uint32_t* count);
/* Six position-weighted 64-bit checksums of the shard's state, computed on the device: potential, activity, lastFire,
 * weight, arrive (+ depol of busy slots), lastSpikeArrival — sum of bits(x[i])*(i+1)*0x9E3779B97F4A7C15 mod 2^64 each.
 * What the parity fixtures pin per step (the reference harness computes the same words from NeuCor's members), and a cheap
 * way to validate a checkpoint. */
int nc_state_signature(nc_engine* e, uint64_t* out6);
/* NeuCor::resetActivities (NeuCor.cpp:233-235,460). */
int nc_reset_activities(nc_engine* e, float now);
/* VoltageDetector::getVoltage's averaging (NeuCor.cpp:360-365) over `near` (ascending IDs of this
 * shard), summed sequentially in float on the device; the neurons must already be at `now`. */
int nc_detector_mean(nc_engine* e, const uint32_t* near, uint32_t n_near, float* out);

/* Device-resident stepping for measurement: record the event lists of live steps on the device
 * ("tape"), snapshot/restore the full state, and replay the taped steps back to back with no
 * host<->device traffic inside the timed region.  Replay is bit-identical to the live run. */
int nc_tape_begin(nc_engine* e, uint32_t max_steps, uint64_t max_events);
int nc_tape_end(nc_engine* e);
int nc_snapshot(nc_engine* e);
int nc_restore(nc_engine* e);
/* Replays taped steps [first, first+count); ms_total = CUDA-event time of the whole region,
 * ms_pass1 / ms_pass2 / ms_exchange = summed per-launch times of the neuron pass, the synapse pass and the
 * fire exchange (may be NULL; asking for them adds event records between the launches). */
int nc_tape_replay(nc_engine* e, uint32_t first, uint32_t count, float* ms_total, float* ms_pass1, float* ms_pass2,
                   float* ms_exchange, uint64_t* hidden_rand_calls, nc_step_stats* stats_or_null);
/* Per-kernel sums [ms] of the last nc_tape_replay that asked for per-kernel times:
 * out4 = { k_stage, k_neuron_pass, fire exchange, synapse kernels (loads + rows + flagged) }. */
int nc_replay_breakdown(const nc_engine* e, float* out4);
/* Device-resident replica of libc's rand() stream.  glibc's default generator is
 * x[n] = x[n-31] + x[n-3] (mod 2^32), rand() = x[n] >> 1; `x31` are its 31 most recent raw values, oldest first.  Once a
 * state is set the engine keeps the stream's position itself: nc_background_draw consumes the draws of NeuCor::run's
 * background-firing loop (NeuCor.cpp:604-607), every window moves the stream on by its hidden rand() calls
 * (NeuCor.cpp:752, summed over the shards), and nc_rand_get_state returns where the stream stands so that the host can
 * put libc's generator there (the application's own rand() calls go on where the reference's would). */
int nc_rand_set_state(nc_engine* e, const uint32_t* x31);
int nc_rand_get_state(nc_engine* e, uint32_t* x31);
/* The background-firing draws of one NeuCor::run() starting at t0 (NeuCor.cpp:604-607), on the device: one test draw per
 * neuron of the WHOLE network (`period` = max(1, int(600 / runSpeed))), two more per hit; the resulting events (kind 2,
 * the last one of a neuron flagged as its scheduledFireTime) of this shard's neurons are merged into the event lists of
 * the nc_step windows that follow, until the next draw or nc_background_clear.  Every shard makes the same call. */
int nc_background_draw(nc_engine* e, float t0, float run_speed, uint32_t period, uint64_t n_neurons);
int nc_background_clear(nc_engine* e);
/* The events of the active background draw that belong to this shard (sorted by neuron) and the number of hits of the whole network. */
int nc_background_read(nc_engine* e, uint32_t capacity, nc_event* out, uint32_t* count, uint32_t* hits);
/* Work done through the event index since the last call: out2 = { busy slots the staging kernel looked at,
 * flag-list entries (slots that delivered or were cleared) }.  For the traffic model of bench.py. */
int nc_index_stats(nc_engine* e, uint64_t* out2);
/* Number of kernel launches issued by this engine since creation. */
uint64_t nc_launch_count(const nc_engine* e);

/* Multi-GPU: the network is partitioned by neuron-ID range, one engine (process, GPU) per
 * shard, each owning the synapses incoming to its neurons.  The only data that crosses shards is the window's
 * fire records — there is nothing like it in the single-threaded reference; it replaces the shared
 * address space through which Neuron::fire reaches its outSynapses (NeuCor.cpp:646-649).  Inside nc_step
 * (and nc_tape_replay) the shards' record blocks are all-gathered between the neuron pass and the synapse
 * pass, and the per-window counters (incl. the hidden rand() count) are summed over shards, so every
 * shard's nc_step reports network-wide numbers.  Transport: an NCCL communicator owned by the engine
 * (in-stream ncclAllGather over NVLink; libnccl is bound at run time) or a caller-provided all-gather. */
typedef struct nc_comm_id { char internal[128]; } nc_comm_id; /* = ncclUniqueId */
/* Rank 0 creates the id, the caller distributes it to all ranks (any out-of-band channel). */
int nc_comm_unique_id(nc_comm_id* out);
/* Collective over all ranks of the job: creates this engine's communicator (rank/world from nc_config). */
int nc_comm_init(nc_engine* e, const nc_comm_id* id);
/* Alternative transport: fn(ctx, send, recv, bytes) must all-gather `bytes` bytes from every rank's `send`
 * into `recv` (rank-major) and return 0; the engine drains its stream before calling it.  Pointers are in the
 * engine's memory space (device memory for the CUDA engine). */
typedef int (*nc_allgather_fn)(void* ctx, const void* send, void* recv, uint64_t bytes);
int nc_set_exchange(nc_engine* e, nc_allgather_fn fn, void* ctx);

/* Self-test: the device replicas of glibc's powf / exp (the libm calls at NeuCor.cpp:672,678,695,710-711,741)
 * evaluated on host arrays, for bit-for-bit comparison against the host libm. x > 0 normal for powf. */
int nc_selftest_powf(nc_engine* e, const float* x, const float* y, float* out, uint64_t n);
int nc_selftest_exp(nc_engine* e, const double* x, double* out, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif /* NEUCOR_B200_H */
