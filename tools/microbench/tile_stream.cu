// Micro-benchmark: does the body-state LAYOUT limit the one-pass kernel's DRAM rate?  Moves exactly the bytes of one fused
// water step (per 32-body tile: 18 state planes in, 14 out; per atom: 3 coordinates + 3 force components in, 3 + 3 out) with no
// arithmetic, one warp per CTA, 8 CTAs per SM, tiles handed out round-robin - once with the state as 27 separate planes
// (plane stride = all bodies; what librbk uses) and once tile-blocked ([tile][plane][32]: a tile's planes are contiguous).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tile_stream tile_stream.cu && ./tile_stream
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NP = 27;
template <bool BLOCKED>
__global__ void __launch_bounds__(32, 8) stream(double* state, size_t ld, const double* d, size_t as, const double* f, double* pos, double* vel,
                                               int numTiles) {
    const int lane = threadIdx.x;
    for (int t = blockIdx.x; t < numTiles; t += gridDim.x) {
        const size_t b = (size_t) t*32 + lane;
        double* s = BLOCKED ? state + (size_t) t*NP*32 + lane : state + b;
        const size_t st = BLOCKED ? 32 : ld;
        double v[18];
#pragma unroll
        for (int k = 0; k < 18; k++) v[k] = s[k*st];
        double a[9], g[9];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const size_t at = (size_t) t*96 + j*32 + lane;
#pragma unroll
            for (int c = 0; c < 3; c++) { a[3*j + c] = d[c*as + at]; g[3*j + c] = f[3*at + c]; }
        }
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 18; k++) acc += v[k];
#pragma unroll
        for (int k = 0; k < 14; k++) s[k*st] = v[k] + 1e-300*acc;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const size_t at = (size_t) t*96 + j*32 + lane;
#pragma unroll
            for (int c = 0; c < 3; c++) { pos[3*at + c] = a[3*j + c] + g[3*j + c]; vel[3*at + c] = a[3*j + c] - g[3*j + c]; }
        }
    }
}

int main() {
    const int nB = 1000000, numTiles = (nB + 31)/32;
    const size_t ld = (size_t) numTiles*32, as = (size_t) numTiles*96;
    double *state, *d, *f, *pos, *vel;
    cudaMalloc(&state, ld*NP*8); cudaMalloc(&d, as*3*8); cudaMalloc(&f, as*3*8); cudaMalloc(&pos, as*3*8); cudaMalloc(&vel, as*3*8);
    cudaMemset(state, 0, ld*NP*8); cudaMemset(d, 0, as*3*8); cudaMemset(f, 0, as*3*8);
    const double bytes = (double) numTiles*32*(18 + 14)*8 + (double) numTiles*96*(3 + 3 + 3 + 3)*8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int blocked = 0; blocked < 2; blocked++)
        for (int ctas = 8; ctas <= 16; ctas += 8) {
            float best = 1e9f;
            for (int rep = 0; rep < 12; rep++) {
                cudaEventRecord(e0);
                if (blocked) stream<true><<<148*ctas, 32>>>(state, ld, d, as, f, pos, vel, numTiles);
                else stream<false><<<148*ctas, 32>>>(state, ld, d, as, f, pos, vel, numTiles);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep >= 2 && ms < best) best = ms;
            }
            printf("%s state, %2d one-warp CTAs per SM: %.1f us, %.0f GB/s (%.0f MB)\n", blocked ? "tile-blocked" : "planar      ", ctas, best*1e3,
                   bytes/best/1e6, bytes/1e6);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
