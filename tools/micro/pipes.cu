// Microbenchmark: do warp shuffles share the shared-memory data pipe with LDS?  (design input for k_asm_tiles v2)
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int NL, int NS, int ND, int MODE>
__global__ void __launch_bounds__(512) k(double *out, int iters, const int *perm)
{
    __shared__ double s[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = i * 0.5;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int idx;
    if (MODE == 0) idx = threadIdx.x;                 // conflict-free, 32 distinct doubles per warp
    else if (MODE == 1) idx = (threadIdx.x & ~31) + (lane & 15);   // two half-warps read the same 16 doubles
    else if (MODE == 2) idx = (threadIdx.x & ~31) + (lane & 3);    // 4 distinct addresses per warp
    else idx = perm[threadIdx.x];                      // random slots
    double acc = 0.0, v = lane * 1.0, f = 1.0000001, g = 0.5;
    int src = (lane * 7 + 3) & 31;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NL; ++j) acc += s[(idx + j * 32 + it) & 4095];
#pragma unroll
        for (int j = 0; j < NS; ++j) { v = __shfl_sync(0xffffffffu, v, (src + j) & 31); }
#pragma unroll
        for (int j = 0; j < ND; ++j) { g = fma(g, f, 1e-9); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + v + g;
}

template <int NL, int NS, int ND, int MODE>
int run(const char *name, double *out, const int *perm)
{
    const int iters = 2000, grid = 148 * 2, thr = 512;
    k<NL, NS, ND, MODE><<<grid, thr>>>(out, 10, perm);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<NL, NS, ND, MODE><<<grid, thr>>>(out, iters, perm);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // per SM: 2 CTAs * 16 warps = 32 warps; warp-instructions per SM per iteration
    const double cyc = ms * 1e-3 * 1.965e9;
    const double per_it = cyc / iters;   // cycles per iteration per SM (all 32 warps)
    printf("%-34s NL=%d NS=%d ND=%d : %.3f ms, %.1f cycles/iter/SM  (per warp-instr: LDS %.2f SHFL %.2f DFMA %.2f)\n", name, NL, NS, ND, ms, per_it,
           NL ? per_it / (32.0 * NL) : 0.0, NS ? per_it / (32.0 * NS) : 0.0, ND ? per_it / (32.0 * ND) : 0.0);
    return 0;
}

int main()
{
    double *out; int *perm;
    CK(cudaMalloc(&out, 148 * 2 * 512 * 8));
    CK(cudaMalloc(&perm, 512 * 4));
    int h[512];
    unsigned s = 12345;
    for (int i = 0; i < 512; ++i) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) & 255; }
    CK(cudaMemcpy(perm, h, sizeof(h), cudaMemcpyHostToDevice));
    run<8, 0, 0, 0>("LDS.64 conflict-free", out, perm);
    run<8, 0, 0, 1>("LDS.64 halves read same 16", out, perm);
    run<8, 0, 0, 2>("LDS.64 4 distinct", out, perm);
    run<8, 0, 0, 3>("LDS.64 random of 256 slots", out, perm);
    run<0, 8, 0, 0>("SHFL only", out, perm);
    run<8, 8, 0, 0>("LDS.64 + SHFL", out, perm);
    run<8, 16, 0, 0>("LDS.64 + 2x SHFL", out, perm);
    run<0, 0, 16, 0>("DFMA only (dependent chain)", out, perm);
    run<8, 0, 16, 0>("LDS.64 + DFMA", out, perm);
    run<0, 8, 16, 0>("SHFL + DFMA", out, perm);
    run<8, 8, 16, 0>("LDS + SHFL + DFMA", out, perm);
    return 0;
}
