// TEST INFRASTRUCTURE — never compiled into or loaded by the product.
//
// The slice of the CUDA runtime API and of the device vocabulary that csrc/engine.cu uses, restated for the CPU on top of
// cuda_block_emu.h, so that the engine's SOURCE (kernels and host side of the C ABI alike) can be compiled with g++ and
// driven through the parity scenarios in the CPU test-suite (tests/test_engine_emulated.py).  "Device memory" is host
// memory, streams are synchronous, a launch runs its blocks one after the other as fibres.  One device, world = 1 only
// (no IPC, no NCCL).  tests/emu_build.py turns `kernel<<<grid, block, smem, stream>>>(args)` into EMU_LAUNCH(...) and the
// `extern __shared__` declaration into a pointer to the emulator's dynamic shared memory; nothing else of the source changes.
#pragma once
#include <chrono>
#include <string>
#include <utility>

#include "cuda_block_emu.h"

// ---- vector types ------------------------------------------------------------------------------------------------------------
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct uint4 { unsigned x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }
static inline uint3 make_uint3(unsigned x, unsigned y, unsigned z) { uint3 r; r.x = x; r.y = y; r.z = z; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
#define NC_EMU_HAS_FLOAT2 1

// ---- more device vocabulary ----------------------------------------------------------------------------------------------------
#define __constant__
template <typename T> inline T __ldcg(const T* p) { return *p; }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __threadfence_system() {}
inline void __nanosleep(unsigned) {}
inline long long clock64() { static thread_local long long c = 0; return c += 64; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __fma_rn(double a, double b, double c) { return __builtin_fma(a, b, c); }
inline double __longlong_as_double(long long x) { double d; memcpy(&d, &x, 8); return d; }
inline long long __double_as_longlong(double d) { long long x; memcpy(&x, &d, 8); return x; }
// CUDA's mixed-width min/max overloads
inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
inline unsigned int min(unsigned int a, unsigned int b) { return a < b ? a : b; }
inline unsigned int max(unsigned int a, unsigned int b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned long long min(unsigned long long a, unsigned int b) { return a < b ? a : b; }
inline unsigned long long min(unsigned int a, unsigned long long b) { return a < b ? a : b; }
inline unsigned long min(unsigned long a, unsigned int b) { return a < b ? a : b; }
inline unsigned long min(unsigned int a, unsigned long b) { return a < b ? a : b; }
inline unsigned long max(unsigned long a, unsigned int b) { return a > b ? a : b; }
inline unsigned long max(unsigned int a, unsigned long b) { return a > b ? a : b; }
inline unsigned int min(unsigned int a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
inline unsigned int min(int a, unsigned int b) { return (unsigned)a < b ? (unsigned)a : b; }
inline unsigned int max(unsigned int a, int b) { return a > (unsigned)b ? a : (unsigned)b; }
inline unsigned int max(int a, unsigned int b) { return (unsigned)a > b ? (unsigned)a : b; }
inline float min(float a, float b) { return fminf(a, b); }
inline float max(float a, float b) { return fmaxf(a, b); }
inline double min(double a, double b) { return fmin(a, b); }
inline double max(double a, double b) { return fmax(a, b); }
// atomics on the signed / mixed argument types the engine uses (the templates in cuda_block_emu.h need both arguments of one type)
inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAnd(unsigned* p, unsigned v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicMin(unsigned* p, unsigned v) { return atomicMin<unsigned>(p, v); }
inline unsigned atomicMax(unsigned* p, unsigned v) { return atomicMax<unsigned>(p, v); }
inline unsigned atomicCAS(unsigned* p, unsigned cmp, unsigned v) { return atomicCAS<unsigned>(p, cmp, v); }

// ---- dynamic shared memory ----------------------------------------------------------------------------------------------------
namespace emu {
inline unsigned char* dyn_smem() { return dyn().data(); }
struct Limits { int sms = 1; int blocksPerSm = 2; };  // one "SM" with two resident blocks: persistent grids still have several blocks (NC_EMU_SMS overrides)
inline Limits limits() {  // (read per call: tests switch NC_EMU_SMS between engines)
    Limits l;
    if (const char* s = getenv("NC_EMU_SMS")) l.sms = std::max(1, atoi(s));
    return l;
}
inline unsigned long long& launches() { static unsigned long long n = 0; return n; }
}  // namespace emu

// ---- runtime API ---------------------------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
typedef int cudaMemcpyKind;
enum { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
struct EmuStream { int dummy; };
typedef EmuStream* cudaStream_t;
struct EmuEvent { std::chrono::steady_clock::time_point t; };
typedef EmuEvent* cudaEvent_t;
enum { cudaStreamNonBlocking = 1 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaLimitMaxL2FetchGranularity = 5 };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
struct cudaIpcMemHandle_t { char reserved[64]; };
struct cudaDeviceProp { int multiProcessorCount; char name[64]; size_t sharedMemPerBlockOptin; size_t totalGlobalMem; int major, minor; };

inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : e == cudaErrorMemoryAllocation ? "out of memory (emulator)" : "not supported by the CPU emulator"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceSetLimit(int, size_t) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof(*p));
    p->multiProcessorCount = emu::limits().sms;
    strcpy(p->name, "CPU block emulator");
    p->sharedMemPerBlockOptin = 227 * 1024; p->totalGlobalMem = (size_t)8 << 30; p->major = 10; p->minor = 0;
    return cudaSuccess;
}
template <typename T> inline cudaError_t cudaMalloc(T** p, size_t n) {
    // a little slack and a fill pattern: reads of never-written device memory show up as garbage, not as zeros
    *p = (T*)malloc(n ? n + 64 : 64);
    if (!*p) return cudaErrorMemoryAllocation;
    memset((void*)*p, getenv("NC_EMU_ZERO") ? 0 : 0xA5, n ? n + 64 : 64);
    return cudaSuccess;
}
template <typename T> inline cudaError_t cudaMallocHost(T** p, size_t n) { *p = (T*)malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaFree(const void* p) { free(const_cast<void*>(p)); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t = nullptr) {
    for (size_t r = 0; r < height; r++) memmove((char*)d + r * dpitch, (const char*)s + r * spitch, width);
    return cudaSuccess;
}
inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { if (n) memset(d, v, n); return cudaSuccess; }
#define cudaMemcpyToSymbol(sym, src, n) (memcpy((void*)&(sym), (src), (n)), cudaSuccess)
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new EmuStream(); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new EmuEvent(); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
template <typename F> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = emu::limits().blocksPerSm; return cudaSuccess; }
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorNotSupported; }
inline cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaErrorNotSupported; }

// kernel<<<grid, block, smem, stream>>>(args)  ->  EMU_LAUNCH((grid), (block), (smem), kernel(args))
// NC_EMU_PROFILE=1: wall time, launches and fibres per kernel on stderr at exit (where the emulator itself spends its time)
namespace emu {
struct Prof {
    struct Row { double s = 0; unsigned long long launches = 0, fibres = 0; };
    std::vector<std::pair<std::string, Row>> rows;
    bool on = getenv("NC_EMU_PROFILE") != nullptr;
    Row& row(const char* call) {
        std::string k(call);
        k = k.substr(0, k.find('('));
        for (auto& r : rows) if (r.first == k) return r.second;
        rows.push_back({k, Row()});
        return rows.back().second;
    }
    ~Prof() {
        if (!on) return;
        for (auto& r : rows) fprintf(stderr, "[emu] %-28s %8.3f s %8llu launches %10llu fibres\n", r.first.c_str(), r.second.s, r.second.launches, r.second.fibres);
    }
};
inline Prof& prof() { static Prof p; return p; }
}  // namespace emu
#define EMU_LAUNCH(grid, block, smem, call)                                                          \
    do {                                                                                             \
        emu::dyn_bytes() = (size_t)(smem);                                                           \
        emu::dyn().assign((size_t)(smem) + 64, 0xA5);                                                \
        emu::launches()++;                                                                           \
        const dim3 g_ = dim3(grid), b_ = dim3(block);                                                \
        emu::S().kernel = #call;                                                                     \
        const auto t_ = std::chrono::steady_clock::now();                                            \
        emu::launch(g_, b_, [&] { call; });                                                          \
        if (emu::prof().on) {                                                                        \
            auto& r_ = emu::prof().row(#call);                                                       \
            r_.s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_).count();    \
            r_.launches++;                                                                           \
            r_.fibres += (unsigned long long)g_.x * g_.y * g_.z * b_.x * b_.y * b_.z;                \
        }                                                                                            \
    } while (0)
