// fp64_peak.cu -- measures the FP64 roofs the normal-equation kernel is judged against (DESIGN.md "Roofline"):
//   DFMA   : dependent-chain-free double FMA throughput on the CUDA cores
//   DMMA   : mma.sync m8n8k4 / m16n8k8 f64 throughput on the tensor cores
//   LDS    : shared-memory read bandwidth with 8-byte and 16-byte loads
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(256) k_dfma(int iters, double* out, double seed)
{
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + threadIdx.x * 1e-9 + i;
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) k_dmma884(int iters, double* out, double seed)
{
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = seed; c[i][1] = seed + i; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) k_dmma1688(int iters, double* out, double seed)
{
    double c[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { c[i][0] = seed; c[i][1] = seed + i; c[i][2] = seed; c[i][3] = seed - i; }
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 1e-9 * threadIdx.x, b1 = b0 * 2;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

// mixed: DMMA and DFMA issued together (do the two share a pipe?)
__global__ void __launch_bounds__(256) k_mixed(int iters, double* out, double seed)
{
    double c[4][2], f[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { c[i][0] = seed; c[i][1] = seed + i; }
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = seed + i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = fma(f[i], 1.0000001, 1e-9);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[i];
    if (s == 12345.678) out[0] = s;
}

template <int W>
__global__ void __launch_bounds__(256) k_lds(int iters, double* out)
{
    __shared__ double sm[4096];
    for (int i = threadIdx.x; i < 4096; i += 256) sm[i] = i;
    __syncthreads();
    double s = 0;
    int idx = threadIdx.x * W;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (W == 1) { s += sm[(idx + u * 256) & 4095]; }
            else { double2 v = *(const double2*)&sm[(idx + u * 512) & 4095]; s += v.x + v.y; }
        }
        idx = (idx + 8) & 4095;
    }
    if (s == 12345.678) out[0] = s;
}

int main()
{
    int dev = 0, sms = 0, clk = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
    double* out;
    CK(cudaMalloc(&out, 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grid = sms * 8, iters = 20000;
    float ms;
    printf("{\"sm_count\": %d, \"clock_khz\": %d", sms, clk);
    for (int rep = 0; rep < 2; ++rep) { k_dfma<<<grid, 256>>>(100, out, 1.0); }
    CK(cudaDeviceSynchronize());
    double best;
#define TIME(launch, flops_per_thread_iter, name, scale, unit)                                            \
    best = 1e30;                                                                                          \
    for (int rep = 0; rep < 5; ++rep) {                                                                   \
        CK(cudaEventRecord(e0)); launch; CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));           \
        CK(cudaGetLastError()); CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;          \
    }                                                                                                     \
    printf(", \"%s\": %.2f", name, (double)grid * 256 * iters * (flops_per_thread_iter) / (best * 1e-3) / scale);
    TIME((k_dfma<<<grid, 256>>>(iters, out, 1.0)), 16 * 2.0, "dfma_tflops", 1e12, "TFLOP/s")
    // m8n8k4: 8*8*4*2 = 512 flop per warp instruction -> 16 flop per thread
    TIME((k_dmma884<<<grid, 256>>>(iters, out, 1.0)), 8 * 16.0, "dmma_m8n8k4_tflops", 1e12, "TFLOP/s")
    // m16n8k8: 16*8*8*2 = 2048 flop per warp instruction -> 64 per thread
    TIME((k_dmma1688<<<grid, 256>>>(iters, out, 1.0)), 4 * 64.0, "dmma_m16n8k8_tflops", 1e12, "TFLOP/s")
    TIME((k_mixed<<<grid, 256>>>(iters, out, 1.0)), 4 * 16.0 + 8 * 2.0, "mixed_dmma_dfma_tflops", 1e12, "TFLOP/s")
    TIME((k_lds<1><<<grid, 256>>>(iters, out)), 8 * 8.0, "lds64_TBps", 1e12, "TB/s")
    TIME((k_lds<2><<<grid, 256>>>(iters, out)), 8 * 16.0, "lds128_TBps", 1e12, "TB/s")
    printf("}\n");
    return 0;
}
