// TEST INFRASTRUCTURE: the kernels of neurocorrelation_b200/csrc/rand_stream.cuh (device-resident rand() stream, background
// firing draws) compiled for the CPU through cuda_block_emu.h, behind two C entry points that size their buffers the way
// nc_background_draw / k_rand_advance's callers in csrc/engine.cu do.  Used by tests/test_rand_stream.py against libc itself.
#include "cuda_block_emu.h"
#define NC_BLOCK_EMU 1
#include "../../neurocorrelation_b200/csrc/rand_stream.cuh"

namespace {
struct Tables {
    std::vector<uint32_t> T, J;
    ncr::RandTables tb;
    Tables() { ncr::build_rand_tables(T, J); tb.T = T.data(); tb.J = J.data(); }
};
Tables& tables() { static Tables t; return t; }
}  // namespace

// One run()'s background draws (NeuCor.cpp:604-607) of the shard [row0, row0+nRows) of a network of nNeurons.
// ctl4 = {events written, hits (whole network), overflow flags, draws consumed}; returns 0, or 1 when outCap is too small.
extern "C" int emu_background_draw(const uint32_t* state31, float t0, float runSpeed, uint32_t period, uint64_t nNeurons, uint64_t row0, uint64_t nRows,
                                   nc_event* out, uint32_t outCap, uint32_t* ctl4, uint32_t* newState31, uint32_t candCapOverride) {
    const uint64_t room = std::min<uint64_t>(nNeurons, 4 * (nNeurons / period) + 256);  // as nc_background_draw
    ncr::BgArgs a = {};
    a.t0 = t0; a.runSpeed = runSpeed; a.period = period; a.nNeurons = nNeurons; a.nDraws = nNeurons + 2 * room + 64;
    a.row0 = row0; a.nRows = nRows;
    const uint32_t bgCap = (uint32_t)std::min<uint64_t>(room + 64, 1u << 26);
    const uint32_t candCap = candCapOverride ? candCapOverride : (uint32_t)std::min<uint64_t>(2ull * bgCap + 1024, 1u << 27);
    std::vector<uint32_t> draws(a.nDraws + a.nDraws / 8), candRaw(candCap), cand(candCap), state(state31, state31 + 31);
    std::vector<nc_event> tmp(bgCap), ev(bgCap);
    uint32_t bgCtl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const unsigned blocks = (unsigned)((a.nDraws + NC_RS_CHUNK - 1) / NC_RS_CHUNK);
    const ncr::RandTables tb = tables().tb;
    emu::launch(dim3(blocks), dim3(NC_RS_ROWS), [&] { ncr::k_bg_generate(tb, state.data(), a, draws.data(), candRaw.data(), candCap, bgCtl + 4); });
    emu::launch(dim3(1), dim3(1024), [&] {
        ncr::k_bg_walk(state.data(), a, draws.data(), candRaw.data(), bgCtl + 4, cand.data(), candCap, tmp.data(), ev.data(), bgCap, bgCtl, newState31);
    });
    memcpy(ctl4, bgCtl, 16);
    const uint32_t n = std::min(bgCtl[0], outCap);
    memcpy(out, ev.data(), (size_t)n * sizeof(nc_event));
    return bgCtl[0] > outCap ? 1 : 0;
}

// The stream moved ahead by the hidden rand() calls of a window (k_rand_advance; `counters`: world blocks of 10 x u64).
extern "C" void emu_rand_advance(uint32_t* state31, const unsigned long long* counters, uint32_t world) {
    const ncr::RandTables tb = tables().tb;
    auto nowait = [] {};
    emu::launch(dim3(1), dim3(32), [&] { ncr::k_rand_advance(tb, state31, counters, world, nowait); });
}
