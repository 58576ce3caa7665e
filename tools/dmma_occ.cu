// dmma_occ.cu -- FP64 tensor-core throughput versus resident warps per SM (how many warps does K_ne need in its
// Gram phase to saturate the FP64 pipe?).  6 independent accumulator chains per warp, like the kernel.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int CH>
__global__ void k884(int iters, double* out, double seed)
{
    double c[CH][2];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c[i][0] = seed; c[i][1] = seed + i; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

template <int CH>
__global__ void k1688(int iters, double* out, double seed)
{
    double c[CH][4];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c[i][0] = seed; c[i][1] = seed + i; c[i][2] = seed; c[i][3] = seed - i; }
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 1e-9 * threadIdx.x, b1 = b0 * 2;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

template <int CH>
__global__ void k16816(int iters, double* out, double seed)
{
    double c[CH][4];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c[i][0] = seed; c[i][1] = seed + i; c[i][2] = seed; c[i][3] = seed - i; }
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = 1e-9 * threadIdx.x * (i + 1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

// DFMA with CH independent chains per thread
template <int CH>
__global__ void kdfma(int iters, double* out, double seed)
{
    double a[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) a[i] = seed + threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) a[i] = fma(a[i], 1.0000001, 1e-9);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += a[i];
    if (s == 12345.678) out[0] = s;
}

int main()
{
    int sms = 0;
    CK(cudaSetDevice(0));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    double* out; CK(cudaMalloc(&out, 8));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 20000;
    printf("{\"sm_count\": %d, \"unit\": \"TFLOP/s\", \"chains_per_warp\": 6, \"rows\": [\n", sms);
    const int warps[] = {4, 8, 12, 16, 24, 32};
    for (int wi = 0; wi < 6; ++wi) {
        const int w = warps[wi];
        float ms; double r[5];
        for (int kind = 0; kind < 5; ++kind) {
            double best = 1e30;
            for (int rep = 0; rep < 4; ++rep) {
                CK(cudaEventRecord(e0));
                if (kind == 0) k884<6><<<sms, w * 32>>>(iters, out, 1.0);
                if (kind == 1) k1688<3><<<sms, w * 32>>>(iters, out, 1.0);
                if (kind == 2) k16816<3><<<sms, w * 32>>>(iters, out, 1.0);
                if (kind == 3) kdfma<6><<<sms, w * 32>>>(iters, out, 1.0);
                if (kind == 4) k884<2><<<sms, w * 32>>>(iters, out, 1.0);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
                CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
            }
            const double flop_per_warp_iter = kind == 0 ? 6 * 512.0 : kind == 1 ? 3 * 2048.0 : kind == 2 ? 3 * 4096.0 : kind == 3 ? 6 * 64.0 : 2 * 512.0;
            r[kind] = (double)sms * w * iters * flop_per_warp_iter / (best * 1e-3) / 1e12;
        }
        printf("  {\"warps_per_sm\": %d, \"dmma_m8n8k4\": %.2f, \"dmma_m16n8k8\": %.2f, \"dmma_m16n8k16\": %.2f, \"dfma\": %.2f, \"dmma_m8n8k4_2chains\": %.2f}%s\n",
               w, r[0], r[1], r[2], r[3], r[4], wi < 5 ? "," : "");
    }
    printf("]}\n");
    return 0;
}
