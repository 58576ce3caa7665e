// pcs_costfn.cu -- initialiser cost evaluation.
//
// estimate_camera_relative_poses (template_handler.py:468-601) scores C + 1 candidate pose tables by running
// bundle_adjustment_costfn (compiled_helpers.py:517-549: P_c [X; 1], perspective divide, pixel-space Brown-Conrady
// distortion nb_distort_prealloc :438-460) over ALL observations once per table and summing the per-observation
// error norms per image (:550-560).  Here all tables are evaluated in one launch over the problem's resident
// observation SoA (grid.y = table), and the per-image sums are formed on the device, so the (tables x 2N) error
// array only leaves the GPU when the caller asks for it.
//   k_costfn_errors    : errors[b][2N] in dd row order (one thread per observation, 16-byte stores)   -- HBM bound
//   k_costfn_per_image : per_image[b][m] = sum of |e| over the observations of image m; runs on the (camera, pose)-
//                        sorted SoA so that a warp's observations form few runs of equal pose: segmented warp
//                        reduction, one FP64 reduction per run.
#include "pcs_internal.cuh"

namespace pcs {

struct CostfnTables {
    const double* im_points;  // [B][M][K][3]
    const double* proj;       // [C][3][4]
    const double* ints;       // [C][3][3]
    const double* dists;      // [C][5]
    int M, K;
};

__device__ __forceinline__ double2 costfn_one(const CostfnTables& t, int b, int c, int m, int k, double2 o)
{
    const double* P = t.proj + 12 * (int64_t)c;
    const double* A = t.ints + 9 * (int64_t)c;
    const double* kd = t.dists + 5 * (int64_t)c;
    const double* X = t.im_points + 3 * (((int64_t)b * t.M + m) * t.K + k);
    const double X0 = X[0], X1 = X[1], X2 = X[2];
    // same association as np.dot(P, [X; 1]) row by row (compiled_helpers.py:541)
    const double p0 = P[0] * X0 + P[1] * X1 + P[2] * X2 + P[3];
    const double p1 = P[4] * X0 + P[5] * X1 + P[6] * X2 + P[7];
    const double p2 = P[8] * X0 + P[9] * X1 + P[10] * X2 + P[11];
    const double u = p0 / p2, v = p1 / p2;
    const double c0 = A[2], c1 = A[5], f0 = A[0], f1 = A[4];
    const double x = (u - c0) / f0, y = (v - c1) / f1;
    const double r2 = x * x + y * y;
    const double kup = 1.0 + kd[0] * r2 + kd[1] * (r2 * r2) + kd[4] * (r2 * r2 * r2);
    double xD = x * kup, yD = y * kup;
    xD += 2.0 * kd[2] * x * y + kd[3] * (r2 + 2.0 * (x * x));
    yD += kd[2] * (r2 + 2.0 * (y * y)) + 2.0 * kd[3] * x * y;
    return make_double2(xD * f0 + c0 - o.x, yD * f1 + c1 - o.y);
}

__global__ void __launch_bounds__(256)
k_costfn_errors(int64_t N, const int32_t* __restrict__ cam, const int32_t* __restrict__ pose, const int32_t* __restrict__ key,
                const double2* __restrict__ uv, CostfnTables t, double2* __restrict__ errors)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int b = blockIdx.y;
    errors[(int64_t)b * N + i] = costfn_one(t, b, cam[i], pose[i], key[i], uv[i]);
}

__global__ void __launch_bounds__(256)
k_costfn_per_image(int64_t N, const int32_t* __restrict__ s_cam, const int32_t* __restrict__ s_pose,
                   const int32_t* __restrict__ s_key, const double2* __restrict__ s_uv, CostfnTables t,
                   double* __restrict__ per_image)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    int m = -1, c = -1;
    double v = 0.0;
    if (i < N) {
        m = s_pose[i];
        c = s_cam[i];
        const double2 e = costfn_one(t, b, c, m, s_key[i], s_uv[i]);
        v = sqrt(e.x * e.x + e.y * e.y);
    }
    // Segmented suffix sum over the warp's runs of ADJACENT lanes with the same (camera, pose) key; the first lane of
    // a run owns the total.  The table is (camera, pose)-sorted, so one pose recurs once per camera: runs are
    // delimited by adjacency (head flags), never by pose equality at a distance -- two runs of the same pose that
    // land in one warp (few observations per pair) stay separate sums.
    const int mp = __shfl_up_sync(0xffffffffu, m, 1), cp = __shfl_up_sync(0xffffffffu, c, 1);
    const bool head = lane == 0 || mp != m || cp != c;
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    const unsigned above = lane == 31 ? 0u : heads & (0xffffffffu << (lane + 1));   // heads of later runs
    const int run_end = above ? __ffs(above) - 1 : 32;               // one past the last lane of this lane's run
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double w = __shfl_down_sync(0xffffffffu, v, off);
        if (lane + off < run_end) v += w;
    }
    if (m >= 0 && head) atomicAdd(per_image + (int64_t)b * t.M + m, v);
}

}  // namespace pcs

using namespace pcs;

extern "C" int pcs_costfn(pcs_problem* p, int n_tables, const double* im_points, const double* proj, const double* intrinsics,
                          const double* dists, double* errors, double* per_image)
{
    PCS_REQUIRE(p && im_points && proj && intrinsics && dists, "NULL argument");
    PCS_REQUIRE(n_tables >= 1 && n_tables <= 65535, "n_tables must be in [1, 65535]");
    PCS_REQUIRE(errors || per_image, "nothing to compute: both outputs are NULL");
    PCS_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = p->stream;
    const int64_t N = p->N, B = n_tables;
    const int64_t n_pts = B * p->M * (int64_t)p->K * 3, n_cam = (int64_t)p->C * (12 + 9 + 5);
    double *d_tab = nullptr, *d_cam = nullptr, *d_err = nullptr, *d_img = nullptr;
    auto cleanup = [&]() { cudaFree(d_tab); cudaFree(d_cam); cudaFree(d_err); cudaFree(d_img); };
#define CF_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { set_error(std::string(#call) + " -> " + cudaGetErrorString(e__)); cleanup(); return PCS_ERR_CUDA; } } while (0)
    CF_CUDA(cudaMalloc((void**)&d_tab, (size_t)n_pts * 8));
    CF_CUDA(cudaMalloc((void**)&d_cam, (size_t)n_cam * 8));
    CF_CUDA(cudaMemcpyAsync(d_tab, im_points, (size_t)n_pts * 8, cudaMemcpyHostToDevice, st));
    CF_CUDA(cudaMemcpyAsync(d_cam, proj, (size_t)p->C * 12 * 8, cudaMemcpyHostToDevice, st));
    CF_CUDA(cudaMemcpyAsync(d_cam + (int64_t)p->C * 12, intrinsics, (size_t)p->C * 9 * 8, cudaMemcpyHostToDevice, st));
    CF_CUDA(cudaMemcpyAsync(d_cam + (int64_t)p->C * 21, dists, (size_t)p->C * 5 * 8, cudaMemcpyHostToDevice, st));
    CostfnTables t{d_tab, d_cam, d_cam + (int64_t)p->C * 12, d_cam + (int64_t)p->C * 21, p->M, p->K};
    const dim3 grid((unsigned)((N + 255) / 256), (unsigned)B);
    if (errors && N > 0) {
        CF_CUDA(cudaMalloc((void**)&d_err, (size_t)(B * N) * 16));
        k_costfn_errors<<<grid, 256, 0, st>>>(N, p->cam, p->pose, p->key, (const double2*)p->uv, t, (double2*)d_err);
        CF_CUDA(cudaGetLastError());
        ++p->n_launches;
        CF_CUDA(cudaMemcpyAsync(errors, d_err, (size_t)(B * N) * 16, cudaMemcpyDeviceToHost, st));
    }
    if (per_image) {
        CF_CUDA(cudaMalloc((void**)&d_img, (size_t)(B * p->M) * 8));
        CF_CUDA(cudaMemsetAsync(d_img, 0, (size_t)(B * p->M) * 8, st));
        if (N > 0) {
            k_costfn_per_image<<<grid, 256, 0, st>>>(N, p->s_cam, p->s_pose, p->s_key, (const double2*)p->s_uv, t, d_img);
            CF_CUDA(cudaGetLastError());
            ++p->n_launches;
        }
        CF_CUDA(cudaMemcpyAsync(per_image, d_img, (size_t)(B * p->M) * 8, cudaMemcpyDeviceToHost, st));
    }
    CF_CUDA(cudaStreamSynchronize(st));
#undef CF_CUDA
    cleanup();
    return PCS_OK;
}
