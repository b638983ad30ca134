// fp64 pipe micro-benchmarks that shaped the EKF kernel (sm_100a): dependent-chain latency, issue cost of a
// warp instruction as a function of active lanes, sincos / division / shuffle cost.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_micro fp64_micro.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ long long clk() { return clock64(); }

template <int CH>
__global__ void k_fma(int iters, int active, double* sink, long long* cyc) {
    double a[CH];
    for (int i = 0; i < CH; ++i) a[i] = 1.0 + 1e-3 * (threadIdx.x + i);
    const double m = 1.0 - 1e-9, c = 1e-9;
    if ((threadIdx.x & 31) >= active) return;
    long long t0 = clk();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) a[i] = fma(a[i], m, c);
    }
    long long t1 = clk();
    double t = 0;
    for (int i = 0; i < CH; ++i) t += a[i];
    if (t == 123.456) sink[0] = t;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_sincos(int iters, double x0, double* sink, long long* cyc) {
    double x = x0 + 1e-3 * threadIdx.x, acc = 0;
    long long t0 = clk();
    for (int it = 0; it < iters; ++it) {
        double s, c;
        sincos(x, &s, &c);
        x = x0 + 1e-6 * (s + c);  // dependent
        acc += s;
    }
    long long t1 = clk();
    if (acc == 123.456) sink[0] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_div(int iters, double* sink, long long* cyc) {
    double x = 1.5 + 1e-3 * threadIdx.x;
    long long t0 = clk();
    for (int it = 0; it < iters; ++it) x = 1.0 / x + 0.5;
    long long t1 = clk();
    if (x == 123.456) sink[0] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__device__ __forceinline__ double rcp_nr(double s) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
    double e = fma(-s, r, 1.0);
    r = fma(r, e, r);
    e = fma(-s, r, 1.0);
    r = fma(r, e, r);
    return r;
}
__global__ void k_rcp(int iters, double* sink, long long* cyc, double* err) {
    double x = 1.5 + 1e-3 * threadIdx.x;
    double worst = 0;
    long long t0 = clk();
    for (int it = 0; it < iters; ++it) x = rcp_nr(x) + 0.5;
    long long t1 = clk();
    for (int i = 0; i < 1000; ++i) {
        double s = 0.001 + 0.37 * i + threadIdx.x * 1e-4;
        double r = rcp_nr(s), e = fabs(r * s - 1.0);
        if (e > worst) worst = e;
    }
    if (x == 123.456) sink[0] = x;
    if (threadIdx.x == 0) { cyc[blockIdx.x] = t1 - t0; err[0] = worst; }
}

__global__ void k_shfl(int iters, double* sink, long long* cyc) {
    double x = 1.5 + 1e-3 * threadIdx.x;
    long long t0 = clk();
    for (int it = 0; it < iters; ++it) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
    long long t1 = clk();
    if (x == 123.456) sink[0] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// several warps per block on one SM: does fp64 issue scale over the 4 sub-partitions?
__global__ void k_fma_warps(int iters, double* sink, long long* cyc) {
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + 1e-3 * (threadIdx.x + i);
    const double m = 1.0 - 1e-9, c = 1e-9;
    __syncthreads();
    long long t0 = clk();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
    }
    __syncthreads();
    long long t1 = clk();
    double t = 0;
    for (int i = 0; i < 8; ++i) t += a[i];
    if (t == 123.456) sink[0] = t;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    double* sink; long long* cyc; double* err;
    cudaMalloc(&sink, 64); cudaMalloc(&cyc, 1024); cudaMalloc(&err, 8);
    long long h; double he;
    const int it = 20000;
#define RUN(label, launch, per)                                   \
    launch; cudaDeviceSynchronize();                              \
    launch; cudaDeviceSynchronize();                              \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);               \
    printf("%-44s %8.2f cycles\n", label, double(h) / (per));
    RUN("dfma dependent latency (1 chain)", (k_fma<1><<<1, 32>>>(it, 32, sink, cyc)), it)
    RUN("dfma 2 chains, per instr", (k_fma<2><<<1, 32>>>(it, 32, sink, cyc)), it * 2.0)
    RUN("dfma 4 chains, per instr", (k_fma<4><<<1, 32>>>(it, 32, sink, cyc)), it * 4.0)
    RUN("dfma 8 chains, per instr, 32 lanes", (k_fma<8><<<1, 32>>>(it, 32, sink, cyc)), it * 8.0)
    RUN("dfma 8 chains, per instr, 16 lanes", (k_fma<8><<<1, 32>>>(it, 16, sink, cyc)), it * 8.0)
    RUN("dfma 8 chains, per instr, 8 lanes", (k_fma<8><<<1, 32>>>(it, 8, sink, cyc)), it * 8.0)
    RUN("dfma 8 chains, per instr, 4 lanes", (k_fma<8><<<1, 32>>>(it, 4, sink, cyc)), it * 8.0)
    RUN("dfma 16 chains, per instr, 32 lanes", (k_fma<16><<<1, 32>>>(it, 32, sink, cyc)), it * 16.0)
    RUN("dfma 8 chains x 4 warps (1 SM), per warp-instr", (k_fma_warps<<<1, 128>>>(it, sink, cyc)), it * 8.0)
    RUN("dfma 8 chains x 8 warps (1 SM), per warp-instr", (k_fma_warps<<<1, 256>>>(it, sink, cyc)), it * 8.0)
    RUN("dfma 8 chains x 16 warps (1 SM), per warp-instr", (k_fma_warps<<<1, 512>>>(it, sink, cyc)), it * 8.0)
    RUN("sincos(x~1) dependent", (k_sincos<<<1, 32>>>(2000, 1.0, sink, cyc)), 2000)
    RUN("sincos(x~30) dependent", (k_sincos<<<1, 32>>>(2000, 30.0, sink, cyc)), 2000)
    RUN("sincos(x~6e4) dependent", (k_sincos<<<1, 32>>>(2000, 6.0e4, sink, cyc)), 2000)
    RUN("sincos(x~6e5, slow path) dependent", (k_sincos<<<1, 32>>>(2000, 6.0e5, sink, cyc)), 2000)
    RUN("1/x + add dependent", (k_div<<<1, 32>>>(it, sink, cyc)), it)
    RUN("rcp.approx + 2 NR + add dependent", (k_rcp<<<1, 32>>>(it, sink, cyc, err)), it)
    cudaMemcpy(&he, err, 8, cudaMemcpyDeviceToHost);
    printf("rcp_nr worst |r*s-1| = %.3e\n", he);
    RUN("shfl_xor(double) + add dependent", (k_shfl<<<1, 32>>>(it, sink, cyc)), it)
    return 0;
}
