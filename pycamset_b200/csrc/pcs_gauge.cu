// pcs_gauge.cu -- scale estimate of the self-calibration gauge transform.
//
// SelfBundleHandler.apply_gauge_transform (standard_bundle_handler.py:339-410) maps the solved target points back to the
// scale of the target model: over all pairs (i < j) of VISIBLE points whose model distance equals the target's square
// size (np.isclose: |d_ref - square| <= atol + rtol |square|), s = mean(d_ref / d_estimate).  The reference builds two
// K x K distance tables with scipy's cdist (:360-366); here one kernel walks the K (K - 1) / 2 pairs, keeps the table in
// registers and reduces sum(ratio) and the pair count -- nothing O(K^2) is ever stored.
#include "pcs_internal.cuh"

namespace pcs {

__global__ void __launch_bounds__(256)
k_gauge_scale(int64_t K, const double* __restrict__ est, const double* __restrict__ ref, const uint8_t* __restrict__ visible,
              double square, double rtol, double atol, double* __restrict__ sum_ratio, unsigned long long* __restrict__ n_pairs)
{
    double acc = 0.0;
    unsigned long long cnt = 0;
    const int64_t total = K * K;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / K, j = t % K;
        if (j <= i || !visible[i] || !visible[j]) continue;
        const double rx = ref[3 * i] - ref[3 * j], ry = ref[3 * i + 1] - ref[3 * j + 1], rz = ref[3 * i + 2] - ref[3 * j + 2];
        const double dr = sqrt(rx * rx + ry * ry + rz * rz);
        if (!(fabs(dr - square) <= atol + rtol * fabs(square))) continue;
        const double ex = est[3 * i] - est[3 * j], ey = est[3 * i + 1] - est[3 * j + 1], ez = est[3 * i + 2] - est[3 * j + 2];
        acc += dr / sqrt(ex * ex + ey * ey + ez * ez);
        ++cnt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    __shared__ double s_acc[8];
    __shared__ unsigned long long s_cnt[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_acc[warp] = acc; s_cnt[warp] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
        unsigned long long c = 0;
        for (int w = 0; w < 8; ++w) { a += s_acc[w]; c += s_cnt[w]; }
        if (c) { atomicAdd(sum_ratio, a); atomicAdd(n_pairs, c); }
    }
}

}  // namespace pcs

using namespace pcs;

extern "C" int pcs_gauge_scale(int device, int64_t n_points, const double* estimate, const double* reference, const uint8_t* visible,
                               double square_size, double rtol, double atol, double* sum_ratio, int64_t* n_pairs)
{
    PCS_REQUIRE(n_points > 0 && estimate && reference && visible && sum_ratio && n_pairs, "NULL argument or no points");
    PCS_CUDA(cudaSetDevice(device));
    double *d_est = nullptr, *d_ref = nullptr, *d_sum = nullptr;
    uint8_t* d_vis = nullptr;
    unsigned long long* d_cnt = nullptr;
    auto cleanup = [&]() { cudaFree(d_est); cudaFree(d_ref); cudaFree(d_sum); cudaFree(d_vis); cudaFree(d_cnt); };
#define G_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { set_error(std::string(#call) + " -> " + cudaGetErrorString(e__)); cleanup(); return PCS_ERR_CUDA; } } while (0)
    const size_t nb = (size_t)n_points * 24;
    G_CUDA(cudaMalloc((void**)&d_est, nb)); G_CUDA(cudaMalloc((void**)&d_ref, nb)); G_CUDA(cudaMalloc((void**)&d_vis, (size_t)n_points));
    G_CUDA(cudaMalloc((void**)&d_sum, 8)); G_CUDA(cudaMalloc((void**)&d_cnt, 8));
    G_CUDA(cudaMemcpy(d_est, estimate, nb, cudaMemcpyHostToDevice));
    G_CUDA(cudaMemcpy(d_ref, reference, nb, cudaMemcpyHostToDevice));
    G_CUDA(cudaMemcpy(d_vis, visible, (size_t)n_points, cudaMemcpyHostToDevice));
    G_CUDA(cudaMemset(d_sum, 0, 8)); G_CUDA(cudaMemset(d_cnt, 0, 8));
    int sms = 0;
    G_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int64_t total = n_points * n_points;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)sms * 8));
    k_gauge_scale<<<grid, 256>>>(n_points, d_est, d_ref, d_vis, square_size, rtol, atol, d_sum, d_cnt);
    G_CUDA(cudaGetLastError());
    unsigned long long h_cnt = 0;
    G_CUDA(cudaMemcpy(sum_ratio, d_sum, 8, cudaMemcpyDeviceToHost));
    G_CUDA(cudaMemcpy(&h_cnt, d_cnt, 8, cudaMemcpyDeviceToHost));
    *n_pairs = (int64_t)h_cnt;
#undef G_CUDA
    cleanup();
    return PCS_OK;
}
