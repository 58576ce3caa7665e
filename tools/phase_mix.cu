// phase_mix.cu -- does the FP64 pipe stay busy when resident warps alternate between a latency-bound DFMA phase
// (the evaluation phase of K_ne: dependent chains, small ILP) and a DMMA phase (the Gram phase: 3 accumulator chains)?
// Every warp runs: [ne_dfma DFMAs in ILP independent chains] then [n_dmma DMMA m8n8k4 in 3 chains], repeated; warps are
// desynchronised by a per-warp phase offset.  Output: achieved TFLOP/s (both instruction kinds) vs warps per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/phase_mix tools/phase_mix.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int ILP>
__global__ void __launch_bounds__(128) k_mix(int iters, int ne_dfma, int n_dmma3, double* out, double seed)
{
    double f[ILP], c[3][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) f[i] = seed + i + threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < 3; ++i) { c[i][0] = seed; c[i][1] = seed + i; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    // desynchronise: warp w starts with (w % 5) / 5 of a DFMA phase
    for (int k = 0; k < (warp % 5) * ne_dfma / (5 * ILP); ++k) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) f[i] = fma(f[i], 1.0000001, 1e-9);
    }
    for (int it = 0; it < iters; ++it) {
        for (int k = 0; k < ne_dfma / ILP; ++k) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) f[i] = fma(f[i], 1.0000001, 1e-9);
        }
        for (int k = 0; k < n_dmma3; ++k) {
#pragma unroll
            for (int i = 0; i < 3; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += f[i];
    if (s == 12345.678) out[0] = s;
}

int main()
{
    int sms = 0;
    CK(cudaSetDevice(0));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    double* out;
    CK(cudaMalloc(&out, 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 400;
    printf("{\"unit\": \"TFLOP/s\", \"dfma_per_phase\": 128, \"dmma_per_phase\": 48, \"rows\": [\n");
    const int ne_list[3] = {128, 0, 128}, nd_list[3] = {16, 16, 0};
    bool first = true;
    for (int cfg = 0; cfg < 3; ++cfg)
        for (int ctas = 1; ctas <= 5; ++ctas) {
            const int ne = ne_list[cfg], nd = nd_list[cfg];
            double res[3];
            for (int v = 0; v < 3; ++v) {
                float best = 1e30f, ms;
                for (int rep = 0; rep < 4; ++rep) {
                    CK(cudaEventRecord(e0));
                    if (v == 0) k_mix<1><<<sms * ctas, 128>>>(iters, ne, nd, out, 1.0);
                    if (v == 1) k_mix<2><<<sms * ctas, 128>>>(iters, ne, nd, out, 1.0);
                    if (v == 2) k_mix<4><<<sms * ctas, 128>>>(iters, ne, nd, out, 1.0);
                    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (rep > 0 && ms < best) best = ms;
                }
                const double flops = (double)sms * ctas * 4 * iters * (ne * 64.0 + nd * 3 * 512.0);
                res[v] = flops / (best * 1e-3) / 1e12;
            }
            printf("%s  {\"dfma\": %d, \"dmma\": %d, \"warps_per_sm\": %d, \"ilp1\": %.2f, \"ilp2\": %.2f, \"ilp4\": %.2f}", first ? "" : ",\n",
                   ne, nd * 3, ctas * 4, res[0], res[1], res[2]);
            first = false;
        }
    printf("\n]}\n");
    return 0;
}
