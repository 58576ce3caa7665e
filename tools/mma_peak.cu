// mma_peak.cu -- pipe rates that decide the mixed-precision normal-equation kernel (DESIGN.md "K_ne, mixed track"):
//   HMMA  tf32 m16n8k8 / m16n8k4, bf16 m16n8k16, f16 m16n8k16 (FP32 accumulate) through mma.sync on sm_100a
//   F2F   double -> float conversions (cvt.rn.f32.f64) and float -> double
//   FFMA  FP32 FMA, DFMA FP64 FMA for scale, and a DFMA + HMMA mix (do the FP64 pipe and the legacy tensor path overlap?)
//   SHFL  64-bit shuffles (2 x SHFL.BFLY) as used by the per-segment FP64 gradient reduction
// Every kernel: `warps` warps per CTA, one CTA per SM x `ctas` co-resident, fixed iteration count, independent
// accumulators so that the issue rate (not the dependent latency) is measured.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_peak tools/mma_peak.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int NACC = 8;

__global__ void k_tf32_1688(int iters, float* out, float seed)
{
    float c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = seed; c[i][1] = seed + i; c[i][2] = seed; c[i][3] = seed - i; }
    unsigned a0 = __float_as_uint(1.0f + threadIdx.x * 1e-3f), a1 = a0 + 64, a2 = a0 + 128, a3 = a0 + 256, b0 = a0 ^ 0x1000, b1 = a0 ^ 0x2000;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678f) out[0] = s;
}

__global__ void k_tf32_1684(int iters, float* out, float seed)
{
    float c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = seed; c[i][1] = seed + i; c[i][2] = seed; c[i][3] = seed - i; }
    unsigned a0 = __float_as_uint(1.0f + threadIdx.x * 1e-3f), a1 = a0 + 64, b0 = a0 ^ 0x1000;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(b0));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678f) out[0] = s;
}

template <bool BF16>
__global__ void k_h16816(int iters, float* out, float seed)
{
    float c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = seed; c[i][1] = seed + i; c[i][2] = seed; c[i][3] = seed - i; }
    unsigned a0 = 0x3c003c00u + threadIdx.x, a1 = a0 + 64, a2 = a0 + 128, a3 = a0 + 256, b0 = a0 ^ 0x10, b1 = a0 ^ 0x20;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (BF16)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678f) out[0] = s;
}

__global__ void k_f2f_down(int iters, float* out, double seed)
{
    double a[16];
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = seed + threadIdx.x * 1e-9 + i; acc[i] = 0.f; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float f;
            // the operand changes every iteration (integer add on the low word, other pipe), so nothing can be hoisted
            a[i] = __longlong_as_double(__double_as_longlong(a[i]) + it);
            asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f) : "d"(a[i]));
            acc[i] = __uint_as_float(__float_as_uint(acc[i]) + __float_as_uint(f));   // integer add on the other pipe keeps the result live
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 12345.678f) out[0] = s;
}

__global__ void k_f2f_up(int iters, float* out, float seed)
{
    float a[16];
    unsigned long long acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = seed + threadIdx.x * 1e-3f + i; acc[i] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            double d;
            a[i] = __uint_as_float(__float_as_uint(a[i]) + it);
            asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(a[i]));
            acc[i] += (unsigned long long)__double_as_longlong(d);
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 12345ull) out[0] = (float)s;
}

__global__ void k_ffma(int iters, float* out, float seed)
{
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + threadIdx.x * 1e-6f + i;
    const float b = 1.0000001f + seed, c = 1e-9f + seed;   // register operands (the immediate form issues at twice the rate)
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 12345.678f) out[0] = s;
}

__global__ void k_dfma(int iters, float* out, double seed)
{
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + threadIdx.x * 1e-9 + i;
    const double b = 1.0000001 + seed, c = 1e-9 + seed;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 12345.678) out[0] = (float)s;
}

// per iteration: 16 DFMA + 8 HMMA tf32 m16n8k8 (independent chains): do the FP64 pipe and the legacy tensor path overlap?
__global__ void k_mix_dfma_hmma(int iters, float* out, double seed)
{
    double a[16];
    float c[NACC][4];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = (float)seed; c[i][1] = (float)seed + i; c[i][2] = (float)seed; c[i][3] = (float)seed - i; }
    const double b = 1.0000001 + seed, cc = 1e-9 + seed;
    unsigned a0 = __float_as_uint(1.0f + threadIdx.x * 1e-3f), a1 = a0 + 64, a2 = a0 + 128, a3 = a0 + 256, b0 = a0 ^ 0x1000, b1 = a0 ^ 0x2000;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            a[2 * i] = fma(a[2 * i], b, cc);
            a[2 * i + 1] = fma(a[2 * i + 1], b, cc);
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = (float)s;
}

// per iteration: 16 DFMA + 8 cvt.rn.f32.f64 (independent): does the conversion unit share the FP64 pipe?
__global__ void k_mix_dfma_f2f(int iters, float* out, double seed)
{
    double a[16], s8[8];
    float acc[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s8[i] = seed + threadIdx.x * 1e-7 + i; acc[i] = 0.f; }
    const double b = 1.0000001 + seed, cc = 1e-9 + seed;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            a[2 * i] = fma(a[2 * i], b, cc);
            a[2 * i + 1] = fma(a[2 * i + 1], b, cc);
            float f;
            s8[i] = __longlong_as_double(__double_as_longlong(s8[i]) + it);
            asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f) : "d"(s8[i]));
            acc[i] = __uint_as_float(__float_as_uint(acc[i]) + __float_as_uint(f));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 12345.678) out[0] = (float)s;
}

__global__ void k_shfl64(int iters, float* out, double seed)
{
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + (i & 15));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345.678) out[0] = (float)s;
}

template <typename K, typename S>
static double run(const char* name, K kern, int warps, int ctas, int iters, double ops_per_thread_iter, S seed, int sms, float* d_out,
                  const char* unit, double scale)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    kern<<<sms * ctas, warps * 32>>>(iters / 10, d_out, seed);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        kern<<<sms * ctas, warps * 32>>>(iters, d_out, seed);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    const double total = ops_per_thread_iter * iters * (double)warps * 32 * ctas * sms;
    const double rate = total / (best * 1e-3) * scale;
    printf("{\"kernel\": \"%s\", \"warps_per_cta\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"rate\": %.4g, \"unit\": \"%s\"}\n", name, warps, ctas, best, rate, unit);
    return rate;
}

int main()
{
    int dev = 0, sms = 0, clk = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
    printf("{\"sms\": %d, \"clock_khz\": %d}\n", sms, clk);
    float* d_out; CK(cudaMalloc(&d_out, 64));
    const int it = 20000;
    // MMA: ops per thread-iteration = NACC MMAs / 32 lanes -> report warp-level MMA instructions per second and TFLOP/s
    for (int warps : {4, 8, 16}) {
        // flops per warp-MMA: m16n8k8 = 2*16*8*8 = 2048, m16n8k4 = 1024, m16n8k16 = 4096
        run("hmma_tf32_m16n8k8", k_tf32_1688, warps, 2, it, NACC * 2048.0 / 32, 0.f, sms, d_out, "TFLOP/s", 1e-12);
        run("hmma_tf32_m16n8k4", k_tf32_1684, warps, 2, it, NACC * 1024.0 / 32, 0.f, sms, d_out, "TFLOP/s", 1e-12);
        run("hmma_bf16_m16n8k16", k_h16816<true>, warps, 2, it, NACC * 4096.0 / 32, 0.f, sms, d_out, "TFLOP/s", 1e-12);
        run("hmma_f16_m16n8k16", k_h16816<false>, warps, 2, it, NACC * 4096.0 / 32, 0.f, sms, d_out, "TFLOP/s", 1e-12);
    }
    for (int warps : {8, 16}) {
        run("f2f_f64_to_f32", k_f2f_down, warps, 2, it, 16, 0.0, sms, d_out, "Gconv/s", 1e-9);
        run("f2f_f32_to_f64", k_f2f_up, warps, 2, it, 16, 0.f, sms, d_out, "Gconv/s", 1e-9);
        run("ffma", k_ffma, warps, 2, it, 16 * 2, 0.f, sms, d_out, "TFLOP/s", 1e-12);
        run("dfma", k_dfma, warps, 2, it, 16 * 2, 0.0, sms, d_out, "TFLOP/s", 1e-12);
        run("mix_16dfma_8hmma(time only)", k_mix_dfma_hmma, warps, 2, it, 16 * 2, 0.0, sms, d_out, "TFLOP/s fp64 part", 1e-12);
        run("mix_16dfma_8f2f(time only; 16 dfma alone = dfma row, 8 f2f alone = half the f2f row)", k_mix_dfma_f2f, warps, 2, it, 16 * 2, 0.0, sms, d_out, "TFLOP/s fp64 part", 1e-12);
        run("shfl64", k_shfl64, warps, 2, it, 8, 0.0, sms, d_out, "G 64-bit shuffles/s (per lane)", 1e-9);
    }
    return 0;
}
