// Bit-exact replicas of glibc 2.39's powf() and exp() for device (and host-test) code.
//
// Why: the reference's arithmetic goes through libm — Neuron::charge_passive powf(0.5, dT)
// (NeuCor.cpp:678), Neuron::getTrace / Synapse::synapticPlasticity powf(decay, dt)
// (NeuCor.cpp:672,741), charge_insynapses exp(0.3702*dT) (NeuCor.cpp:695) and the AP waveform's two
// exp() (NeuCor.cpp:710-711).  glibc's powf is not correctly rounded, so CUDA's powf/exp would
// flip a threshold decision within tens of simulated ms.  These are restatements
// of the algorithms glibc ships (ARM optimized-routines: sysdeps/ieee754/flt-32/e_powf.c,
// sysdeps/ieee754/dbl-64/e_exp.c) with the FMA contraction pattern of the x86-64 `_fma` IFUNC
// variants that run on every FMA-capable host (verified against the libm.so.6 disassembly and
// on >10^8 samples per function, tests/test_libm_replica.py).  Tables: glibc_tables.h (extracted
// from libm.so.6 by tools/extract_libm_tables.py).
//
// Every floating-point operation is spelled with an explicit-rounding intrinsic so that nvcc can
// neither contract nor reassociate anything (the .cu files are also built with -fmad=false).
#pragma once
#include <stdint.h>
#include "glibc_tables.h"

#if defined(__CUDACC__)
#define NC_HD __host__ __device__ __forceinline__
#define NC_HDM __host__ __device__ __forceinline__
#else
#define NC_HD static inline
#define NC_HDM inline
#endif

namespace ncm {

#if defined(__CUDA_ARCH__)
NC_HD double as_f64(unsigned long long u) { return __longlong_as_double((long long)u); }
NC_HD unsigned long long as_u64(double d) { return (unsigned long long)__double_as_longlong(d); }
NC_HD float as_f32(uint32_t u) { return __uint_as_float(u); }
NC_HD uint32_t as_u32(float f) { return __float_as_uint(f); }
NC_HD double mul64(double a, double b) { return __dmul_rn(a, b); }
NC_HD double add64(double a, double b) { return __dadd_rn(a, b); }
NC_HD double sub64(double a, double b) { return __dsub_rn(a, b); }
NC_HD double fma64(double a, double b, double c) { return __fma_rn(a, b, c); }
NC_HD float mul32(float a, float b) { return __fmul_rn(a, b); }
NC_HD float add32(float a, float b) { return __fadd_rn(a, b); }
NC_HD float sub32(float a, float b) { return __fsub_rn(a, b); }
NC_HD double div64(double a, double b) { return __ddiv_rn(a, b); }
NC_HD float div32(float a, float b) { return __fdiv_rn(a, b); }
#else
NC_HD double as_f64(unsigned long long u) { double d; __builtin_memcpy(&d, &u, 8); return d; }
NC_HD unsigned long long as_u64(double d) { unsigned long long u; __builtin_memcpy(&u, &d, 8); return u; }
NC_HD float as_f32(uint32_t u) { float f; __builtin_memcpy(&f, &u, 4); return f; }
NC_HD uint32_t as_u32(float f) { uint32_t u; __builtin_memcpy(&u, &f, 4); return u; }
// host build: compile with -ffp-contract=off so that only the spelled-out fma()s fuse
NC_HD double mul64(double a, double b) { return a * b; }
NC_HD double add64(double a, double b) { return a + b; }
NC_HD double sub64(double a, double b) { return a - b; }
NC_HD double fma64(double a, double b, double c) { return __builtin_fma(a, b, c); }
NC_HD float mul32(float a, float b) { return a * b; }
NC_HD float add32(float a, float b) { return a + b; }
NC_HD float sub32(float a, float b) { return a - b; }
NC_HD double div64(double a, double b) { return a / b; }
NC_HD float div32(float a, float b) { return a / b; }
#endif

// Tables (2.5 KB in total) are uploaded to __constant__ memory once and copied to shared memory at the start of every
// kernel that evaluates powf/exp (math_tables_to_shared): the lanes of a warp index them with DIFFERENT values, which the
// constant cache would serialise.
#if defined(__CUDACC__)
__device__ __constant__ unsigned long long d_POWF_LOG2_TAB[32];
__device__ __constant__ unsigned long long d_EXP2F_TAB[32];
__device__ __constant__ unsigned long long d_EXP_TAB[256];
__shared__ unsigned long long s_POWF_LOG2_TAB[32];
__shared__ unsigned long long s_EXP2F_TAB[32];
__shared__ unsigned long long s_EXP_TAB[256];
// every thread of the block must call this before the first powf_pos / exp_glibc
__device__ __forceinline__ void math_tables_to_shared() {
    for (unsigned i = threadIdx.x; i < 256u; i += blockDim.x) {
        s_EXP_TAB[i] = d_EXP_TAB[i];
        if (i < 32u) { s_POWF_LOG2_TAB[i] = d_POWF_LOG2_TAB[i]; s_EXP2F_TAB[i] = d_EXP2F_TAB[i]; }
    }
    __syncthreads();
}
#endif

NC_HD unsigned long long tab_powf_log2(int i) {
#if defined(__CUDA_ARCH__)
    return s_POWF_LOG2_TAB[i];
#else
    return NC_POWF_LOG2_TAB[i];
#endif
}
NC_HD unsigned long long tab_exp2f(int i) {
#if defined(__CUDA_ARCH__)
    return s_EXP2F_TAB[i];
#else
    return NC_EXP2F_TAB[i];
#endif
}
NC_HD unsigned long long tab_exp(int i) {
#if defined(__CUDA_ARCH__)
    return s_EXP_TAB[i];
#else
    return NC_EXP_TAB[i];
#endif
}

// ---- powf -----------------------------------------------------------------------------------
// e_powf.c log2_inline: log2(x) for positive normal x given as its bit pattern.
NC_HD double powf_log2_inline(uint32_t ix) {
    const double A0 = as_f64(0x3fd27616c9496e0bULL), A1 = as_f64(0xbfd71969a075c67aULL),
                 A2 = as_f64(0x3fdec70a6ca7baddULL), A3 = as_f64(0xbfe7154748bef6c8ULL),
                 A4 = as_f64(0x3ff71547652ab82bULL);
    uint32_t tmp = ix - 0x3f330000u;
    int i = (int)((tmp >> (23 - 4)) & 15u);
    uint32_t top = tmp & 0xff800000u;
    uint32_t iz = ix - top;
    int k = (int32_t)top >> 23;
    double invc = as_f64(tab_powf_log2(2 * i)), logc = as_f64(tab_powf_log2(2 * i + 1));
    double z = (double)as_f32(iz);
    double r = fma64(z, invc, -1.0);
    double y0 = add64(logc, (double)k);
    double r2 = mul64(r, r);
    double y = fma64(A0, r, A1);
    double p = fma64(A2, r, A3);
    double r4 = mul64(r2, r2);
    double q = fma64(A4, r, y0);
    q = fma64(p, r2, q);
    y = fma64(y, r4, q);
    return y;
}
// e_powf.c exp2_inline with sign_bias = 0.
NC_HD float powf_exp2_inline(double xd) {
    const double SHIFT = as_f64(0x42e8000000000000ULL);  // 0x1.8p52 / 32
    const double C0 = as_f64(0x3fac6af84b912394ULL), C1 = as_f64(0x3fcebfce50fac4f3ULL),
                 C2 = as_f64(0x3fe62e42ff0c52d6ULL);
    double kd = add64(xd, SHIFT);
    unsigned long long ki = as_u64(kd);
    kd = sub64(kd, SHIFT);
    double r = sub64(xd, kd);
    unsigned long long t = tab_exp2f((int)(ki & 31u));
    t += ki << (52 - 5);
    double s = as_f64(t);
    double z = fma64(C0, r, C1);
    double r2 = mul64(r, r);
    double y = fma64(C2, r, 1.0);
    y = fma64(z, r2, y);
    y = mul64(y, s);
    return (float)y;
}
// powf(x, y) for x a positive, finite, normal float (the engine's bases are the reference's
// constants 0.5 / 0.75 / 0.65 or user-set decays, validated on upload) and ANY y:
// y = 0 -> 1; y = +-inf -> 0 / inf / 1 by |x| vs 1; y NaN -> NaN (1 if x == 1) — e_powf.c:158-173.
NC_HD float powf_pos(float x, float y) {
    uint32_t ix = as_u32(x), iy = as_u32(y);
    if (2u * iy - 1u >= 2u * 0x7f800000u - 1u) {  // zeroinfnan(iy)
        if (2u * iy == 0u) return 1.0f;
        if (ix == 0x3f800000u) return 1.0f;
        if (2u * iy > 2u * 0x7f800000u) return add32(x, y);  // NaN
        if ((ix < 0x3f800000u) == !(iy & 0x80000000u)) return 0.0f;
        return mul32(y, y);
    }
    double logx = powf_log2_inline(ix);
    double ylogx = mul64((double)y, logx);
    if (((as_u64(ylogx) >> 47) & 0xffffu) >= (0x405f800000000000ULL >> 47)) {  // |y*log2(x)| >= 126
        if (ylogx > as_f64(0x405fffffffd1d571ULL)) return as_f32(0x7f800000u);     // overflow -> +inf
        if (ylogx <= -150.0) return 0.0f;                                           // underflow -> +0
    }
    return powf_exp2_inline(ylogx);
}

// ---- exp (double) ---------------------------------------------------------------------------
NC_HD double exp_specialcase(double tmp, unsigned long long sbits, unsigned long long ki) {
    double scale, y;
    if ((ki & 0x80000000ULL) == 0) {
        sbits -= 1009ULL << 52;
        scale = as_f64(sbits);
        y = mul64(as_f64(0x7f00000000000000ULL), fma64(scale, tmp, scale));  // 0x1p1009 * (scale + scale*tmp)
        return y;
    }
    sbits += 1022ULL << 52;
    scale = as_f64(sbits);
    y = fma64(scale, tmp, scale);
    if (y < 1.0) {
        double hi, lo;
        lo = fma64(scale, tmp, sub64(scale, y));
        hi = add64(1.0, y);
        lo = add64(add64(sub64(1.0, hi), y), lo);
        y = sub64(add64(hi, lo), 1.0);
        if (y == 0.0) y = 0.0;
    }
    return mul64(as_f64(0x0010000000000000ULL), y);  // 0x1p-1022 * y
}
// e_exp.c __exp (N = 128, polynomial order 5), FMA-contracted as in the x86-64 __exp_fma build.
NC_HD double exp_glibc(double x) {
    const double InvLn2N = as_f64(0x40671547652b82feULL), Shift = as_f64(0x4338000000000000ULL),
                 NegLn2hiN = as_f64(0xbf762e42fefa0000ULL), NegLn2loN = as_f64(0xbd0cf79abc9e3b3aULL),
                 C2 = as_f64(0x3fdffffffffffdbdULL), C3 = as_f64(0x3fc555555555543cULL),
                 C4 = as_f64(0x3fa55555cf172b91ULL), C5 = as_f64(0x3f81111167a4d017ULL);
    uint32_t abstop = (uint32_t)(as_u64(x) >> 52) & 0x7ffu;
    if (abstop - 0x3c9u >= 0x408u - 0x3c9u) {
        if (abstop - 0x3c9u >= 0x80000000u) return add64(1.0, x);  // |x| < 2^-54
        if (abstop >= 0x409u) {                                   // |x| >= 1024
            if (as_u64(x) == 0xfff0000000000000ULL) return 0.0;
            if (abstop >= 0x7ffu) return add64(1.0, x);
            if (as_u64(x) >> 63) return 0.0;
            return as_f64(0x7ff0000000000000ULL);
        }
        abstop = 0;  // large |x|: handled in exp_specialcase
    }
    double kd = fma64(InvLn2N, x, Shift);
    unsigned long long ki = as_u64(kd);
    kd = sub64(kd, Shift);
    double r = fma64(kd, NegLn2hiN, x);
    r = fma64(kd, NegLn2loN, r);
    unsigned idx = 2u * (unsigned)(ki & 127u);
    unsigned long long top = ki << (52 - 7);
    double tail = as_f64(tab_exp(idx));
    unsigned long long sbits = tab_exp(idx + 1) + top;
    double r2 = mul64(r, r);
    double p23 = fma64(r, C3, C2);
    double p45 = fma64(r, C5, C4);
    double r4 = mul64(r2, r2);
    double tmp = fma64(r2, p23, add64(tail, r));
    tmp = fma64(r4, p45, tmp);
    if (abstop == 0) return exp_specialcase(tmp, sbits, ki);
    double scale = as_f64(sbits);
    return fma64(scale, tmp, scale);
}

}  // namespace ncm
