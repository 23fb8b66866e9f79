// cuda_shim.h — TEST INFRASTRUCTURE ONLY.  Lets g++ compile the device headers (csrc/device/*.cuh) and an emitted
// model translation unit for the HOST, so the per-pair device logic (event walk, closed forms, ODE solvers, likelihood)
// can be exercised against the oracle in this GPU-less authoring container before GPU minutes are spent on it.
// It is a debugging double of the kernels, not a product path: nothing under pharmsol_b200/ includes or links it, the
// shipped library has no CPU route (tests/test_abi_host.py::test_compute_fails_loudly_without_device), and no parity or
// performance claim rests on it — the `-m gpu` tests are the parity tests.
// One "thread" runs at a time: threadIdx / blockIdx are plain globals, warp collectives degenerate to one lane.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define PSI_HOST_SIM 1
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __constant__ const
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)
#define __CUDACC__ 1

struct hs_dim3 { unsigned x = 1, y = 1, z = 1; };
extern hs_dim3 threadIdx, blockIdx, blockDim, gridDim;
struct double2 { double x, y; };

template <class T> inline T __ldg(const T* p) { return *p; }
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
inline int __double2hiint(double d) { return (int)(__double_as_longlong(d) >> 32); }
inline int __double2loint(double d) { return (int)(unsigned int)(__double_as_longlong(d) & 0xffffffffLL); }
inline double __hiloint2double(int hi, int lo) { return __longlong_as_double((long long)(((unsigned long long)(unsigned int)hi << 32) | (unsigned int)lo)); }
inline float __powf(float a, float b) { return std::pow(a, b); }
inline float __logf(float a) { return std::log(a); }
inline void __sincosf(float a, float* s, float* c) { *s = std::sin(a); *c = std::cos(a); }
inline float rsqrtf(float a) { return 1.0f / std::sqrt(a); }
inline void sincos(double a, double* s, double* c) { *s = std::sin(a); *c = std::cos(a); }
inline double rsqrt(double a) { return 1.0 / std::sqrt(a); }
inline unsigned int __umulhi(unsigned int a, unsigned int b) { return (unsigned int)(((unsigned long long)a * b) >> 32); }
inline unsigned __activemask() { return 1u; }       // one lane: flush_counters takes its per-thread path
template <class T> inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
template <class T> inline T __shfl_up_sync(unsigned, T v, int) { return v; }
inline void __syncthreads() {}
inline int __syncthreads_or(int p) { return p; }
inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; if (v < o) *p = v; return o; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p += v; return o; }
inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { const unsigned int o = *p; *p += v; return o; }
inline int atomicAdd(int* p, int v) { const int o = *p; *p += v; return o; }
using std::fmax; using std::fmin; using std::fabs; using std::sqrt; using std::exp; using std::log; using std::pow; using std::fma;
using std::erfc; using std::cbrt; using std::atan2; using std::ceil; using std::floor; using std::isfinite;
inline float fminf(float a, float b) { return std::fmin(a, b); }
inline float fmaxf(float a, float b) { return std::fmax(a, b); }
inline float sqrtf(float a) { return std::sqrt(a); }
