// gram_loop.cu -- isolates the Gram phase of K_ne: 3 LDS.64 + 6 DMMA per k-step, 16 k-steps per batch.
// Variants: V=0 fragments from shared memory each k-step (kernel-like); V=1 same with all 48 fragment loads of a batch
// hoisted before its DMMAs (register-staged); V=2 no loads (register operands; pipe roof).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int V>
__global__ void __launch_bounds__(128) k(int batches, double* out)
{
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* ws = sm + warp * 1536;
    for (int i = lane; i < 1536; i += 32) ws[i] = 1.0 + 1e-6 * i;
    __syncwarp();
    const int g8 = lane >> 2, jb = (lane >> 1) & 1, jr = lane & 1;
    const int L = 16 * jb + 8 * jr + (g8 ^ (4 * jb));
    double a00[2] = {0, 0}, a01[2] = {0, 0}, a02[2] = {0, 0}, a11[2] = {0, 0}, a12[2] = {0, 0}, a22[2] = {0, 0};
    for (int b = 0; b < batches; ++b) {
        if (V == 0) {
#pragma unroll 4
            for (int ks = 0; ks < 16; ++ks) {
                const double* f = ws + 32 * ks + (L ^ (((ks & 1) << 3) | (ks & 2)));
                const double v0 = f[0], v1 = f[512], v2 = f[1024];
                dmma(a00[0], a00[1], v0, v0); dmma(a01[0], a01[1], v0, v1); dmma(a02[0], a02[1], v0, v2);
                dmma(a11[0], a11[1], v1, v1); dmma(a12[0], a12[1], v1, v2); dmma(a22[0], a22[1], v2, v2);
            }
        } else if (V == 1) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double v0[8], v1[8], v2[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int ks = 8 * h + q;
                    const double* f = ws + 32 * ks + (L ^ (((ks & 1) << 3) | (ks & 2)));
                    v0[q] = f[0]; v1[q] = f[512]; v2[q] = f[1024];
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    dmma(a00[0], a00[1], v0[q], v0[q]); dmma(a01[0], a01[1], v0[q], v1[q]); dmma(a02[0], a02[1], v0[q], v2[q]);
                    dmma(a11[0], a11[1], v1[q], v1[q]); dmma(a12[0], a12[1], v1[q], v2[q]); dmma(a22[0], a22[1], v2[q], v2[q]);
                }
            }
        } else {
            const double v0 = 1.0 + lane, v1 = 2.0 + lane, v2 = 3.0 + lane;
#pragma unroll 4
            for (int ks = 0; ks < 16; ++ks) {
                dmma(a00[0], a00[1], v0, v0); dmma(a01[0], a01[1], v0, v1); dmma(a02[0], a02[1], v0, v2);
                dmma(a11[0], a11[1], v1, v1); dmma(a12[0], a12[1], v1, v2); dmma(a22[0], a22[1], v2, v2);
            }
        }
        __syncwarp();
    }
    const double s = a00[0] + a00[1] + a01[0] + a01[1] + a02[0] + a02[1] + a11[0] + a11[1] + a12[0] + a12[1] + a22[0] + a22[1];
    if (s == 12345.678) out[0] = s;
}

int main()
{
    int sms = 0;
    CK(cudaSetDevice(0));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    double* out; CK(cudaMalloc(&out, 8));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int batches = 2000;
    CK(cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152));
    CK(cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152));
    CK(cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152));
    printf("{\"unit\": \"TFLOP/s\", \"rows\": [\n");
    for (int ctas = 1; ctas <= 4; ++ctas) {
        double r[3];
        for (int v = 0; v < 3; ++v) {
            double best = 1e30; float ms;
            for (int rep = 0; rep < 4; ++rep) {
                CK(cudaEventRecord(e0));
                if (v == 0) k<0><<<sms * ctas, 128, 49152>>>(batches, out);
                if (v == 1) k<1><<<sms * ctas, 128, 49152>>>(batches, out);
                if (v == 2) k<2><<<sms * ctas, 128, 49152>>>(batches, out);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
                CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
            }
            r[v] = (double)sms * ctas * 4 * batches * 16 * 6 * 512.0 / (best * 1e-3) / 1e12;
        }
        printf("  {\"warps_per_sm\": %d, \"lds_per_kstep\": %.2f, \"lds_hoisted\": %.2f, \"register_operands\": %.2f}%s\n", ctas * 4, r[0], r[1], r[2], ctas < 4 ? "," : "");
    }
    printf("]}\n");
    return 0;
}
