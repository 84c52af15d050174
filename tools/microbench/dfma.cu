// Micro-benchmark: fp64 FMA latency and per-SM throughput on the GPU at hand (B200: sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma dfma.cu && ./dfma
// Prints cycles per DFMA for a single warp with 1..8 independent dependency chains (latency / ILP) and the
// aggregate DFMA rate per SM per clock with 1..16 resident warps per SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double* out, long long* cycles, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x*1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x*blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int ILP> void run(int warps, int blocks, const char* label) {
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double)*blocks*warps*32); cudaMalloc(&cyc, sizeof(long long)*blocks);
    const int iters = 4096;
    chain<ILP><<<blocks, warps*32>>>(out, cyc, iters, 0.999999, 1e-7);
    chain<ILP><<<blocks, warps*32>>>(out, cyc, iters, 0.999999, 1e-7);
    long long h[1024];
    cudaMemcpy(h, cyc, sizeof(long long)*blocks, cudaMemcpyDeviceToHost);
    double c = (double) h[0];
    printf("%s ILP=%d warps/SM=%2d : %.2f cycles per DFMA per warp, %.1f DFMA lanes/clk/SM\n", label, ILP, warps,
           c/(iters*ILP), 32.0*warps*iters*ILP/c);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<1>(1, 1, "latency   ");
    run<2>(1, 1, "latency   ");
    run<4>(1, 1, "latency   ");
    run<8>(1, 1, "latency   ");
    run<1>(4, 148, "throughput"); run<1>(8, 148, "throughput"); run<1>(16, 148, "throughput"); run<1>(32, 148, "throughput");
    run<4>(4, 148, "throughput"); run<4>(8, 148, "throughput"); run<4>(16, 148, "throughput");
    run<8>(8, 148, "throughput");
    return 0;
}
