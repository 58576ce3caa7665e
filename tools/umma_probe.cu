// umma_probe.cu -- can the per-segment 16 x 16 Gram update of K_ne go onto tcgen05 (UMMA, accumulators in TMEM)?
//
// What K_ne would need: every warp owns its own stream of (camera, pose) segments, so every warp needs its own
// accumulator and issues its own small MMAs:  D[16 x 32] (+)= hi^T [hi | lo]  over K = 8 staged rows (4 observations) per
// instruction, kind::tf32 (FP32 values split x = hi + lo into two TF32 terms), operands in shared memory in the K-major
// no-swizzle canonical layout (MN-major without swizzle produced all-zero accumulators for TF32 on this part -- the debug
// kernel below keeps that experiment).  tcgen05.mma has M >= 64: the 16 useful rows are placed at rows 16 q .. 16 q + 15
// (q = warp % 4: the TMEM lane quadrant the warp may read) by moving the A descriptor's start address back by 2 q row groups;
// the other 48 rows of A read whatever shared memory is there.
//
// The probe (1) checks that this layout / descriptor / TMEM read-back gives the right numbers, (2) checks whether four warps
// can share accumulator columns with the disable-output-lane mask, (3) measures the sustained MMA rate of an SM for this
// shape -- the number that decides whether the tensor path pays: 4 observations per instruction, ~3 KB of operand reads.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_probe tools/umma_probe.cu
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

constexpr int WARPS = 4;
constexpr int NKB = 8;              // K-blocks (of 8 staged rows = 4 observations) per batch of 32 observations
constexpr int LBO = 144;            // K-major no-swizzle: core matrix = 8 rows (MN) x 16 B (4 TF32 along K); next core matrix along K
constexpr int SBO = 288;            // next group of 8 rows (MN)
constexpr int KB_BYTES = 4 * SBO + 32;   // tile = 4 row groups [hi 0..7, hi 8..15, lo 0..7, lo 8..15]; the skews (16 B per core matrix, 32 B per
                                         // tile) make the 8-byte staging stores of a half warp bank-conflict free
constexpr int LEAD_BYTES = 4096;    // slack in front of / behind the tiles: the 64-row A operand reaches 1.5 KB back
constexpr int WARP_BYTES = NKB * KB_BYTES;
constexpr int SMEM_BYTES = 2 * LEAD_BYTES + WARPS * WARP_BYTES + 64;

#define CHECK(call)                                                                              \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); return 1; } \
    } while (0)

__host__ __device__ inline float hi_val(int w, int k, int col) { return (float)(((k * 5 + col * 3 + w) % 9) - 4) * 0.5f; }
__host__ __device__ inline float lo_val(int w, int k, int col) { return (float)(((k * 3 + col * 7 + 2 * w) % 7) - 3) * 0.125f; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, no swizzle: start address, leading / stride byte offsets in 16-byte units, version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) |
           (1ull << 46);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                          uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(m0), "r"(m1), "r"(m2), "r"(m3)
        : "memory");
}

__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity)
{
    for (int spin = 0; spin < (1 << 22); ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}

// mode bit 0: 1 = the four warps share accumulator columns and rely on the disable-output-lane mask
// rounds: timing loop length (each round = NKB MMAs + commit + wait per warp)
__global__ void __launch_bounds__(WARPS * 32) k_probe(int mode, int rounds, int active_warps, float* __restrict__ out, long long* __restrict__ clk,
                                                     int* __restrict__ err)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint32_t tmem_base_sm;
    __shared__ __align__(8) unsigned long long bars[WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, q = warp & 3;

    // fill: slack = large finite values (would show up if rows outside the quadrant leaked into the result)
    for (int i = threadIdx.x; i < SMEM_BYTES / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 1000.0f;
    __syncthreads();
    unsigned char* base_w = sm + LEAD_BYTES + warp * WARP_BYTES;
    for (int i = lane; i < NKB * 32 * 8; i += 32) {
        const int kk = i & 7, mn = (i >> 3) & 31, kb = i >> 8;   // element (mn, kk) of K-block kb; mn < 16: hi column mn, else lo column mn - 16
        const int k = 8 * kb + kk;
        *reinterpret_cast<float*>(base_w + kb * KB_BYTES + (mn >> 3) * SBO + (kk >> 2) * LBO + (mn & 7) * 16 + (kk & 3) * 4) =
            mn < 16 ? hi_val(warp, k, mn) : lo_val(warp, k, mn - 16);
    }
    if (threadIdx.x == 0)
        for (int w = 0; w < WARPS; ++w) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[w])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sm)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores above -> visible to the tensor core's reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_sm;
    const bool shared_cols = mode & 1;
    const uint32_t col0 = shared_cols ? 0u : 32u * warp;
    const uint32_t tmem_d = tmem_base + col0;                                 // lane field 0: the MMA addresses all 128 lanes
    const uint32_t tmem_rd = tmem_base + ((uint32_t)(32 * q) << 16) + col0;   // this warp's quadrant
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 32, M = 64
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((64u >> 4) << 24);
    const uint32_t mk[4] = {shared_cols && q != 0 ? 0xffffffffu : 0u, shared_cols && q != 1 ? 0xffffffffu : 0u,
                            shared_cols && q != 2 ? 0xffffffffu : 0u, shared_cols && q != 3 ? 0xffffffffu : 0u};
    const uint32_t bar = smem_u32(&bars[warp]);
    uint32_t parity = 0;
    bool ok = true;
    const uint32_t tile0 = smem_u32(base_w);

    const long long t0 = clock64();
    if (warp < active_warps) {
        for (int r = 0; r < rounds && ok; ++r) {
            if (lane == 0) {
#pragma unroll
                for (int kb = 0; kb < NKB; ++kb) {
                    const uint32_t t = tile0 + kb * KB_BYTES;
                    const uint64_t ad = make_desc(t - 2 * q * SBO, LBO, SBO);   // rows 16 q .. 16 q + 15 = the two hi row groups
                    const uint64_t bd = make_desc(t, LBO, SBO);                 // N = 32: hi groups 0, 1 | lo groups 2, 3
                    umma_tf32(tmem_d, ad, bd, idesc, kb > 0 ? 1u : 0u, mk[0], mk[1], mk[2], mk[3]);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
            }
            __syncwarp();
            ok = mbar_wait(bar, parity);
            parity ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
    }
    const long long t1 = clock64();
    if (!ok && lane == 0) atomicExch(err, 1);

    if (ok && warp < active_warps) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
            "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
              "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
              "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
              "=r"(v[31])
            : "r"(tmem_rd)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (blockIdx.x == 0)
            for (int c = 0; c < 32; ++c) out[(warp * 32 + lane) * 32 + c] = __uint_as_float(v[c]);
    }
    if (lane == 0) clk[blockIdx.x * WARPS + warp] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
}

// debug: warp 0 alone issues the 8 MMAs of its batch into columns 0..31; every warp dumps its lane quadrant of those
// columns (where do the 64 rows of D land?), after a tcgen05.st / tcgen05.ld round trip through columns 64..95
__global__ void __launch_bounds__(WARPS * 32) k_debug(int variant, float* __restrict__ out, float* __restrict__ rt, int* __restrict__ err)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint32_t tmem_base_sm;
    __shared__ __align__(8) unsigned long long bars[1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < SMEM_BYTES / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 1000.0f;
    __syncthreads();
    unsigned char* base_w = sm + LEAD_BYTES;
    for (int i = threadIdx.x; i < NKB * 8 * 8 * 4; i += blockDim.x) {
        const int e = i & 3, r = (i >> 2) & 7, j = (i >> 5) & 7, kb = i >> 8;
        const int k = 8 * kb + r, col = 4 * (j & 3) + e;
        *reinterpret_cast<float*>(base_w + kb * KB_BYTES + j * 128 + r * 16 + e * 4) = j < 4 ? hi_val(0, k, col) : lo_val(0, k, col);
    }
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sm)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_sm;
    if (threadIdx.x == 0) rt[128 * 4] = __uint_as_float(tmem_base);
    // round trip: lane l of warp w stores (1000 w + l + 0.25 c) into column 64 + c, c = 0..3
    {
        const uint32_t ta = tmem_base + ((uint32_t)(32 * warp) << 16) + 64;
        const float f0 = 1000.f * warp + lane, f1 = f0 + 0.25f, f2 = f0 + 0.5f, f3 = f0 + 0.75f;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ta), "r"(__float_as_uint(f0)), "r"(__float_as_uint(f1)),
                     "r"(__float_as_uint(f2)), "r"(__float_as_uint(f3))
                     : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        uint32_t g0, g1, g2, g3;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(g0), "=r"(g1), "=r"(g2), "=r"(g3) : "r"(ta) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float* o = rt + (warp * 32 + lane) * 4;
        o[0] = __uint_as_float(g0); o[1] = __uint_as_float(g1); o[2] = __uint_as_float(g2); o[3] = __uint_as_float(g3);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const bool kmajor = variant & 1;
    // variant bit 0: operands declared K-major instead (descriptor strides as the K-major canonical layout would read them)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (kmajor ? 0u : (1u << 15) | (1u << 16)) | ((32u >> 3) << 17) | ((64u >> 4) << 24);
    const uint32_t bar = smem_u32(&bars[0]);
    bool ok = true;
    if (warp == 0) {
        if (lane == 0) {
            const uint32_t tile0 = smem_u32(base_w);
            for (int kb = 0; kb < NKB; ++kb) {
                const uint32_t t = tile0 + kb * KB_BYTES;
                const uint64_t ad = make_desc(t, (variant & 2) ? 128 : 1024, (variant & 2) ? 1024 : 128);
                const uint64_t bd = make_desc(t, (variant & 2) ? 128 : 1024, (variant & 2) ? 1024 : 128);
                umma_tf32(tmem_base, ad, bd, idesc, kb > 0 ? 1u : 0u, 0u, 0u, 0u, 0u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        }
        __syncwarp();
    }
    ok = mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok && lane == 0) atomicExch(err, 1);
    if (ok) {
        uint32_t v[32];
        const uint32_t tmem_rd = tmem_base + ((uint32_t)(32 * warp) << 16);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
            "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
              "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
              "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
              "=r"(v[31])
            : "r"(tmem_rd)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int c = 0; c < 32; ++c) out[(warp * 32 + lane) * 32 + c] = __uint_as_float(v[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
}

int main(int argc, char** argv)
{
    int dev = 0, sms = 0;
    CHECK(cudaSetDevice(dev));
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CHECK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    float* out;
    long long* clk;
    int* err;
    CHECK(cudaMalloc(&out, WARPS * 32 * 32 * sizeof(float)));
    CHECK(cudaMalloc(&clk, (size_t)sms * 4 * WARPS * sizeof(long long)));
    CHECK(cudaMalloc(&err, sizeof(int)));
    CHECK(cudaFuncSetAttribute(k_debug, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    if (argc > 1) {
        float* rt;
        CHECK(cudaMalloc(&rt, (128 * 4 + 4) * sizeof(float)));
        for (int variant = 0; variant < 4; ++variant) {
            CHECK(cudaMemset(out, 0, WARPS * 32 * 32 * sizeof(float)));
            CHECK(cudaMemset(rt, 0, (128 * 4 + 4) * sizeof(float)));
            CHECK(cudaMemset(err, 0, sizeof(int)));
            k_debug<<<1, WARPS * 32, SMEM_BYTES>>>(variant, out, rt, err);
            CHECK(cudaDeviceSynchronize());
            std::vector<float> h(WARPS * 32 * 32), hr(128 * 4 + 4);
            int herr = 0;
            CHECK(cudaMemcpy(h.data(), out, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
            CHECK(cudaMemcpy(hr.data(), rt, hr.size() * sizeof(float), cudaMemcpyDeviceToHost));
            CHECK(cudaMemcpy(&herr, err, sizeof(int), cudaMemcpyDeviceToHost));
            uint32_t tb; memcpy(&tb, &hr[128 * 4], 4);
            printf("variant %d (bit0: K-major idesc, bit1: LBO/SBO swapped): timeout %d, tmem_base 0x%08x, st/ld round trip lane 0: %g %g, lane 33: %g, lane 127: %g\n", variant, herr, tb,
                   hr[0], hr[1], hr[33 * 4], hr[127 * 4 + 3]);
            // expected G[a][b] for a few entries
            auto ref = [&](int a, int b) { double r = 0; for (int k = 0; k < 8 * NKB; ++k) r += (double)hi_val(0, k, a) * (b < 16 ? hi_val(0, k, b) : lo_val(0, k, b - 16)); return r; };
            printf("  expected row 0: %g %g %g %g .. col16: %g ; row 1: %g %g\n", ref(0, 0), ref(0, 1), ref(0, 2), ref(0, 3), ref(0, 16), ref(1, 0), ref(1, 1));
            for (int l = 0; l < 128; ++l) {
                bool nz = false;
                for (int c = 0; c < 32; ++c) nz = nz || h[l * 32 + c] != 0.0f;
                if (nz && (l % 16 < 2 || l == 17)) printf("  lane %3d: %g %g %g %g %g %g %g %g .. c16 %g c31 %g\n", l, h[l * 32], h[l * 32 + 1], h[l * 32 + 2], h[l * 32 + 3], h[l * 32 + 4],
                                  h[l * 32 + 5], h[l * 32 + 6], h[l * 32 + 7], h[l * 32 + 16], h[l * 32 + 31]);
            }
            int nzl = 0; for (int l = 0; l < 128; ++l) { bool nz = false; for (int c = 0; c < 32; ++c) nz = nz || h[l * 32 + c] != 0.0f; nzl += nz; }
            printf("  lanes with non-zero data: %d\n", nzl);
        }
        return 0;
    }
    printf("{\"device_sms\": %d", sms);
    for (int mode = 0; mode < 2; ++mode) {
        CHECK(cudaMemset(out, 0, WARPS * 32 * 32 * sizeof(float)));
        CHECK(cudaMemset(err, 0, sizeof(int)));
        k_probe<<<1, WARPS * 32, SMEM_BYTES>>>(mode, 1, WARPS, out, clk, err);
        CHECK(cudaDeviceSynchronize());
        std::vector<float> h(WARPS * 32 * 32);
        int herr = 0;
        CHECK(cudaMemcpy(h.data(), out, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
        CHECK(cudaMemcpy(&herr, err, sizeof(int), cudaMemcpyDeviceToHost));
        double max_err = 0.0;
        int bad = 0;
        for (int w = 0; w < WARPS; ++w)
            for (int a = 0; a < 16; ++a)
                for (int b = 0; b < 32; ++b) {
                    double ref = 0.0;
                    for (int k = 0; k < 8 * NKB; ++k) ref += (double)hi_val(w, k, a) * (b < 16 ? hi_val(w, k, b) : lo_val(w, k, b - 16));
                    const double got = h[(w * 32 + a) * 32 + b];
                    const double e = std::fabs(got - ref);
                    if (e > max_err) max_err = e;
                    if (e > 1e-3) ++bad;
                }
        printf(", \"%s\": {\"barrier_timeout\": %d, \"max_abs_err\": %.3g, \"wrong_entries\": %d, \"sample_got\": [%.4g, %.4g, %.4g], \"lane16_got\": %.4g}",
               mode ? "shared_columns_lane_mask" : "private_columns", herr, max_err, bad, h[0], h[1], h[32 + 17], h[16 * 32]);
    }
    // sustained rate: `ctas` CTAs per SM worth of grid, 4 warps each issuing rounds x 8 MMAs
    const int rounds = 2000;
    for (int aw = 1; aw <= WARPS; aw *= 2) {
        CHECK(cudaMemset(err, 0, sizeof(int)));
        for (int ctas = 1; ctas <= 4; ctas *= 2) {
            if (aw < WARPS && ctas > 1) continue;
            k_probe<<<sms * ctas, WARPS * 32, SMEM_BYTES>>>(0, rounds, aw, out, clk, err);
            CHECK(cudaDeviceSynchronize());
            std::vector<long long> hc((size_t)sms * ctas * WARPS);
            CHECK(cudaMemcpy(hc.data(), clk, hc.size() * sizeof(long long), cudaMemcpyDeviceToHost));
            long long mx = 0;
            for (size_t i = 0; i < hc.size(); ++i) mx = hc[i] > mx ? hc[i] : mx;
            const double per_mma = (double)mx / ((double)rounds * NKB * aw * ctas);
            printf(", \"rate_%d_warps_x_%d_ctas_per_sm\": {\"clocks_per_mma_per_sm\": %.2f, \"clocks_per_observation_per_sm\": %.2f, \"clocks_per_round_per_warp\": %.1f}",
                   aw, ctas, per_mma, per_mma / 4.0, (double)mx / rounds);
        }
    }
    printf("}\n");
    return 0;
}
