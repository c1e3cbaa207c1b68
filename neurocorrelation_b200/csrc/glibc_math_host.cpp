// Host build of glibc_math.cuh for CPU-side verification of the device math replicas against the
// host libm (tests/test_libm_replica.py).  Compile: g++ -O2 -mfma -ffp-contract=off.
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "glibc_math.cuh"

extern "C" {
float nc_mathhost_powf(float x, float y) { return ncm::powf_pos(x, y); }
double nc_mathhost_exp(double x) { return ncm::exp_glibc(x); }

// Compares powf_pos(base, y) with libm powf for every float y whose bit pattern lies in
// [lo_bits, hi_bits] with the given stride. Returns the number of mismatches (first one in *bad_y).
uint64_t nc_mathhost_check_powf_range(float base, uint32_t lo_bits, uint32_t hi_bits, uint32_t stride, float* bad_y) {
    uint64_t bad = 0;
    for (uint64_t b = lo_bits; b <= hi_bits; b += stride) {
        uint32_t u = (uint32_t)b;
        float y; memcpy(&y, &u, 4);
        float a = powf(base, y), m = ncm::powf_pos(base, y);
        if (memcmp(&a, &m, 4) != 0 && !(a != a && m != m)) { if (!bad && bad_y) *bad_y = y; bad++; }
    }
    return bad;
}
// exp(k * (double)dT) for every float dT in the bit range — the charge_insynapses argument (k = 0.3702).
uint64_t nc_mathhost_check_exp_scaled_range(double k, uint32_t lo_bits, uint32_t hi_bits, uint32_t stride, float* bad_x) {
    uint64_t bad = 0;
    for (uint64_t b = lo_bits; b <= hi_bits; b += stride) {
        uint32_t u = (uint32_t)b;
        float x; memcpy(&x, &u, 4);
        double a = exp(k * (double)x), m = ncm::exp_glibc(k * (double)x);
        if (memcmp(&a, &m, 8) != 0) { if (!bad && bad_x) *bad_x = x; bad++; }
    }
    return bad;
}
// The AP waveform's arguments (NeuCor.cpp:710-711): -(x*x)/denom for every float t in the bit range,
// x = t - off1 - off2 in float.
uint64_t nc_mathhost_check_exp_gauss_range(float off1, float off2, double denom, uint32_t lo_bits, uint32_t hi_bits, uint32_t stride) {
    uint64_t bad = 0;
    for (uint64_t b = lo_bits; b <= hi_bits; b += stride) {
        uint32_t u = (uint32_t)b;
        float t; memcpy(&t, &u, 4);
        float x = t - off1 - off2;
        double arg = -(x * x) / denom;
        double a = exp(arg), m = ncm::exp_glibc(arg);
        if (memcmp(&a, &m, 8) != 0) bad++;
    }
    return bad;
}
// Uniformly random doubles in [lo, hi] (splitmix64).
uint64_t nc_mathhost_check_exp_random(double lo, double hi, uint64_t n, uint64_t seed, double* bad_x) {
    uint64_t bad = 0, s = seed;
    for (uint64_t i = 0; i < n; i++) {
        s += 0x9E3779B97F4A7C15ULL;
        uint64_t z = s; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z ^= z >> 31;
        double x = lo + (hi - lo) * ((z >> 11) * 0x1.0p-53);
        double a = exp(x), m = ncm::exp_glibc(x);
        if (memcmp(&a, &m, 8) != 0 && !(a != a && m != m)) { if (!bad && bad_x) *bad_x = x; bad++; }
    }
    return bad;
}
uint64_t nc_mathhost_check_powf_random(float base_lo, float base_hi, float y_lo, float y_hi, uint64_t n, uint64_t seed) {
    uint64_t bad = 0, s = seed;
    for (uint64_t i = 0; i < n; i++) {
        s += 0x9E3779B97F4A7C15ULL;
        uint64_t z = s; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z ^= z >> 31;
        float x = base_lo + (base_hi - base_lo) * (float)((z >> 40) * 0x1.0p-24);
        float y = y_lo + (y_hi - y_lo) * (float)(((z >> 8) & 0xffffff) * 0x1.0p-24);
        float a = powf(x, y), m = ncm::powf_pos(x, y);
        if (memcmp(&a, &m, 4) != 0 && !(a != a && m != m)) bad++;
    }
    return bad;
}
}
