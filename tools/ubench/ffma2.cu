// Microbenchmark: issue rate of FFMA / FADD vs the packed FFMA2 / FADD2 (fma.rn.f32x2, add.rn.f32x2) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
template <int MODE>
__global__ void k(float* out, long long* clk, int iters) {
    float a[16]; u64 p[16];
    for (int i = 0; i < 16; ++i) { a[i] = threadIdx.x * 0.001f + i; p[i] = (u64)__float_as_uint(a[i]) | ((u64)__float_as_uint(a[i] + 0.5f) << 32); }
    const float m = 1.0001f, c = 0.5f;
    const u64 m2 = (u64)__float_as_uint(m) | ((u64)__float_as_uint(m) << 32), c2 = (u64)__float_as_uint(c) | ((u64)__float_as_uint(c) << 32);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], m, c);
            if (MODE == 1) a[i] = a[i] + a[(i + 1) & 15];
            if (MODE == 2) p[i] = fma2(p[i], m2, c2);
            if (MODE == 3) p[i] = add2(p[i], p[(i + 1) & 15]);
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 16; ++i) s += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* clk; cudaMalloc(&out, 1 << 20); cudaMalloc(&clk, 1024);
    const int iters = 2048;
    const char* names[4] = {"FFMA ", "FADD ", "FFMA2", "FADD2"};
    for (int threads : {128, 256, 512, 1024}) {
        for (int mode = 0; mode < 4; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<1, threads>>>(out, clk, iters);
                if (mode == 1) k<1><<<1, threads>>>(out, clk, iters);
                if (mode == 2) k<2><<<1, threads>>>(out, clk, iters);
                if (mode == 3) k<3><<<1, threads>>>(out, clk, iters);
                cudaDeviceSynchronize();
            }
            long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
            const double per = (double)c / (iters * 16.0) / (threads / 128.0);
            printf("%s threads=%4d: %.2f clk per warp-instruction per SMSP\n", names[mode], threads, per);
        }
    }
    return 0;
}
