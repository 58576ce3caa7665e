// lat_probe.cu -- dependent-issue latencies (cycles) of the instructions on the Cholesky tile-factor chain.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lat_probe tools/lat_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
__global__ void k(double* out, long long* cyc, double seed)
{
    double y = seed, z = seed * 0.5;
    long long t0, t1;
    const int N = 512;
    // MUFU.RCP64H chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(y));
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0) / N;
    // DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(z) : "d"(y));
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = (t1 - t0) / N;
    // 64-bit shuffle chain (two SHFL)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) z = __shfl_sync(0xffffffffu, z, (i * 7) & 31);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = (t1 - t0) / N;
    // fp32 MUFU.RCP chain
    float f = (float)seed;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f));
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = (t1 - t0) / N;
    // STS -> LDS round trip through shared memory
    __shared__ double sm[64];
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { sm[threadIdx.x] = z; __syncwarp(); z = sm[(threadIdx.x + 1) & 31]; __syncwarp(); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = (t1 - t0) / N;
    // rsqrt (library) chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) y = rsqrt(y + 1.5);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = (t1 - t0) / N;
    // cvt f64->f32->f64 chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { float g; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(g) : "d"(z)); asm volatile("cvt.f64.f32 %0, %1;" : "=d"(z) : "f"(g)); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = (t1 - t0) / N;
    out[threadIdx.x] = y + z + f;
}
int main()
{
    double* out; long long* cyc; long long h[7];
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 7 * 8);
    for (int r = 0; r < 3; ++r) k<<<1, 32>>>(out, cyc, 1.37);
    cudaMemcpy(h, cyc, 7 * 8, cudaMemcpyDeviceToHost);
    printf("{\"mufu_rcp64h\": %lld, \"dfma\": %lld, \"shfl64\": %lld, \"mufu_rcp_f32\": %lld, \"sts_lds_syncwarp\": %lld, \"rsqrt_f64\": %lld, \"cvt_f64_f32_f64\": %lld}\n",
           h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
    return 0;
}
