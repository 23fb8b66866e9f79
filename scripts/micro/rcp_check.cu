// Accuracy of the MUFU.RCP64H seed and of the two refinements built on it (psi_common.cuh rcp_nr), measured on the device:
// max relative error over log-uniform random doubles, in units of 2^-53.   nvcc -arch=sm_100a -o rcp_check rcp_check.cu
#include <cstdio>
#include <cstdint>
#include <cmath>
__device__ __forceinline__ double seed(double x) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }
__device__ __forceinline__ double newton2(double x) { double r = seed(x); r = fma(fma(-x, r, 1.0), r, r); r = fma(fma(-x, r, 1.0), r, r); return r; }
__device__ __forceinline__ double cubic(double x) { double r = seed(x); const double e = fma(-x, r, 1.0); return fma(fma(e, e, e), r, r); }
__device__ unsigned long long splitmix(unsigned long long& s) { unsigned long long z = (s += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
__global__ void check(double* out, int per_thread) {
    unsigned long long s = 0x1234567ull + 7919ull * (blockIdx.x * blockDim.x + threadIdx.x);
    double m0 = 0, m1 = 0, m2 = 0;
    for (int i = 0; i < per_thread; ++i) {
        const unsigned long long b = splitmix(s);
        const double mant = 1.0 + (double)(b >> 12) * (1.0 / 4503599627370496.0);
        const int ex = (int)(b & 0xfff) % 600 - 300;
        const double x = ldexp(mant, ex) * ((b & 0x800) ? -1.0 : 1.0);
        const double exact = 1.0 / x;
        m0 = fmax(m0, fabs(seed(x) - exact) / fabs(exact));
        m1 = fmax(m1, fabs(newton2(x) - exact) / fabs(exact));
        m2 = fmax(m2, fabs(cubic(x) - exact) / fabs(exact));
    }
    atomicMax((unsigned long long*)&out[0], __double_as_longlong(m0));
    atomicMax((unsigned long long*)&out[1], __double_as_longlong(m1));
    atomicMax((unsigned long long*)&out[2], __double_as_longlong(m2));
}
int main() {
    double* d; cudaMalloc(&d, 24); cudaMemset(d, 0, 24);
    check<<<592, 256>>>(d, 2000);
    double h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    const double u = ldexp(1.0, -53);
    printf("{\"samples\": %lld, \"seed_rel_err\": %.3e, \"seed_bits\": %.2f, \"newton2_ulp53\": %.3f, \"cubic_ulp53\": %.3f}\n", 592ll * 256 * 2000, h[0], -log2(h[0]), h[1] / u, h[2] / u);
    return cudaGetLastError() != cudaSuccess;
}
