// Micro-benchmark, second round: the one-pass water step's byte pattern (see tile_stream.cu) with the atom outputs written
//   MODE 0  by per-lane 8-byte stores (three per atom and array, rows of 24 bytes)            - what the kernels do
//   MODE 1  through shared memory and ONE bulk store per array and tile (cp.async.bulk.global.shared::cta, 2304 bytes)
//   MODE 2  not at all (reads + state writes only), MODE 3 no reads of forces / coordinates (writes only) - the two halves
//   MODE 4  every lane writes the 72 contiguous bytes of ITS body's three atoms: four 16-byte stores + one 8-byte store per array
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tile_stream2 tile_stream2.cu && ./tile_stream2
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NP = 27;
__device__ __forceinline__ unsigned smemAddr(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }
template <int MODE>
__global__ void __launch_bounds__(32, 8) stream(double* state, size_t ld, const double* d, size_t as, const double* f, double* pos, double* vel,
                                               int numTiles) {
    __shared__ alignas(128) double sOut[2][2][288];
    const int lane = threadIdx.x;
    int it = 0;
    for (int t = blockIdx.x; t < numTiles; t += gridDim.x, it++) {
        const size_t b = (size_t) t*32 + lane;
        double* s = state + b;
        double v[18];
#pragma unroll
        for (int k = 0; k < 18; k++) v[k] = s[k*ld];
        double a[9], g[9];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const size_t at = (size_t) t*96 + j*32 + lane;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                a[3*j + c] = MODE == 3 ? 1.0 : d[c*as + at];
                g[3*j + c] = MODE == 3 ? 2.0 : f[3*at + c];
            }
        }
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 18; k++) acc += v[k];
#pragma unroll
        for (int k = 0; k < 14; k++) s[k*ld] = v[k] + 1e-300*acc;
        if (MODE == 0 || MODE == 3) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const size_t at = (size_t) t*96 + j*32 + lane;
#pragma unroll
                for (int c = 0; c < 3; c++) { pos[3*at + c] = a[3*j + c] + g[3*j + c]; vel[3*at + c] = a[3*j + c] - g[3*j + c]; }
            }
        }
        else if (MODE == 4) {
            double* const out[2] = {pos + (size_t) t*288 + 9*lane, vel + (size_t) t*288 + 9*lane};
#pragma unroll
            for (int w = 0; w < 2; w++) {
                double x[9];
#pragma unroll
                for (int k = 0; k < 9; k++) x[k] = w ? a[k] - g[k] : a[k] + g[k];
                if (lane & 1) {                                   // 72*lane is 8 mod 16: one word, then four aligned pairs
                    out[w][0] = x[0];
#pragma unroll
                    for (int k = 0; k < 4; k++) *reinterpret_cast<double2*>(out[w] + 1 + 2*k) = make_double2(x[1 + 2*k], x[2 + 2*k]);
                }
                else {
#pragma unroll
                    for (int k = 0; k < 4; k++) *reinterpret_cast<double2*>(out[w] + 2*k) = make_double2(x[2*k], x[2*k + 1]);
                    out[w][8] = x[8];
                }
            }
        }
        else if (MODE == 1) {
            const int buf = it & 1;
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");      // the stores that read this buffer two tiles ago
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const int at = j*32 + lane;
#pragma unroll
                for (int c = 0; c < 3; c++) { sOut[buf][0][3*at + c] = a[3*j + c] + g[3*j + c]; sOut[buf][1][3*at + c] = a[3*j + c] - g[3*j + c]; }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 2304;" :: "l"(pos + (size_t) t*288), "r"(smemAddr(sOut[buf][0])) : "memory");
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 2304;" :: "l"(vel + (size_t) t*288), "r"(smemAddr(sOut[buf][1])) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        else {
            double x = 0.0;
#pragma unroll
            for (int k = 0; k < 9; k++) x += a[k] + g[k];
            if (x == 1.2345e300) pos[lane] = x;
        }
    }
    if (MODE == 1 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int MODE> void run(const char* name, double bytes, double* state, size_t ld, const double* d, size_t as, const double* f, double* pos, double* vel, int numTiles) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int ctas = 8; ctas <= 16; ctas += 8) {
        float best = 1e9f;
        for (int rep = 0; rep < 12; rep++) {
            cudaEventRecord(e0);
            stream<MODE><<<148*ctas, 32>>>(state, ld, d, as, f, pos, vel, numTiles);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep >= 2 && ms < best) best = ms;
        }
        printf("%-44s %2d one-warp CTAs per SM: %6.1f us, %5.0f GB/s (%.0f MB)\n", name, ctas, best*1e3, bytes/best/1e6, bytes/1e6);
    }
}

int main() {
    const int nB = 1000000, numTiles = (nB + 31)/32;
    const size_t ld = (size_t) numTiles*32, as = (size_t) numTiles*96;
    double *state, *d, *f, *pos, *vel;
    cudaMalloc(&state, ld*NP*8); cudaMalloc(&d, as*3*8); cudaMalloc(&f, as*3*8); cudaMalloc(&pos, as*3*8); cudaMalloc(&vel, as*3*8);
    cudaMemset(state, 0, ld*NP*8); cudaMemset(d, 0, as*3*8); cudaMemset(f, 0, as*3*8);
    const double st = (double) numTiles*32*(18 + 14)*8, in = (double) numTiles*96*6*8, out = (double) numTiles*96*6*8;
    run<0>("per-lane stores", st + in + out, state, ld, d, as, f, pos, vel, numTiles);
    run<1>("bulk stores from shared memory", st + in + out, state, ld, d, as, f, pos, vel, numTiles);
    run<2>("no atom outputs", st + in, state, ld, d, as, f, pos, vel, numTiles);
    run<3>("no atom inputs", st + out, state, ld, d, as, f, pos, vel, numTiles);
    run<4>("72 contiguous bytes per lane, 16-byte stores", st + in + out, state, ld, d, as, f, pos, vel, numTiles);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
