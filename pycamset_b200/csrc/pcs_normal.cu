// pcs_normal.cu -- K_ne: fused residual + analytic Jacobian + J^T J / J^T r (template chain), the north-star kernel.
//
// No reference counterpart computes J^T J (scipy's LSMR consumes the CSR Jacobian, optimisation_handling.py:88-98);
// the per-observation math is the reference chain projection + extrinsic3D + template_points
// (function_block_implementations.py:21-211, matmul_map.py:147-243) as restated in pcs_math.cuh.
//
// Design (B200, FP64):
//   * Observations are sorted by (camera, pose) at problem build; a "segment" is one (camera, pose) pair.
//   * A warp owns a contiguous range of WHOLE segments (balanced by observation count) and walks it in batches
//     of 32 observations, one lane per observation: the lane evaluates residual and Jacobian in registers and
//     parks its two augmented rows  J' = [pose(6) r 0 | cam(0..7) | cam(8..14) 0]  (24 doubles each) in shared
//     memory -- 12 KB per warp, swizzled so that both the 16-byte row stores and the fragment loads below are
//     bank-conflict free.
//   * The Gram update  G += J'^T J'  runs on the FP64 tensor cores: per 2 observations (4 rows = one k-step) the warp
//     loads 3 fragment values per lane and issues 6 DMMA m8n8k4 (the upper-triangular 8x8 tile pairs of the 24x24
//     Gram matrix).  A-fragment and B-fragment of a tile are the same register, so a k-step costs 3 LDS.64.
//   * Tiles (0,*) hold V_m, g_m, W_{c,m} (flushed per segment: W by plain stores -- the warp owns the segment --
//     V/g_m by FP64 reductions) plus g_c and r.r in row 6, which keep accumulating; tiles (1,1) (1,2) (2,2) hold
//     U_c and are flushed when the camera changes.
//   HBM traffic per observation: (u,v) 16 B + key 4 B + segment id 4 B = 24 B read; outputs are O(segments).
#include <algorithm>
#include <cstdlib>

#include "pcs_internal.cuh"
#include "pcs_math.cuh"

namespace pcs {

constexpr int NE_WARPS = 4;                  // warps per CTA
constexpr int NE_TILE_DOUBLES = 64 * 8;      // one 8-column tile of the 64 staged rows
constexpr int NE_WARP_DOUBLES = 3 * NE_TILE_DOUBLES;

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// Shared-memory position of staged row (observation slot o in 0..31, row r in {u, v}), first column of its tile.
// rho is a bijection (o, r) -> 0..63; `rot` (even) rotates the 8 columns of the row.  With this swizzle
//   - the 16-byte stores of a quarter warp (8 consecutive o, fixed r and column pair) hit 8 distinct 16-byte banks;
//   - the 8-byte fragment loads of a half warp (4 rows of one k-step x 4 columns) hit 16 distinct 8-byte banks.
__device__ __forceinline__ int stage_row(int o, int r) { return (2 * o + (r ^ ((o >> 1) & 1))) * 8; }
__device__ __forceinline__ int stage_rot(int o) { return 4 * (o & 1) + 2 * ((o >> 2) & 1); }

__device__ __forceinline__ void stage_store_row(double* __restrict__ ws, int o, int r, const double t0[8], const double t1[8],
                                                const double t2[8])
{
    const int row = stage_row(o, r), rot = stage_rot(o);
#pragma unroll
    for (int c = 0; c < 8; c += 2) {
        const int pos = row + ((c + rot) & 7);
        *reinterpret_cast<double2*>(ws + pos) = make_double2(t0[c], t0[c + 1]);
        *reinterpret_cast<double2*>(ws + NE_TILE_DOUBLES + pos) = make_double2(t1[c], t1[c + 1]);
        *reinterpret_cast<double2*>(ws + 2 * NE_TILE_DOUBLES + pos) = make_double2(t2[c], t2[c + 1]);
    }
}

struct NeAcc {
    double a00[2], a01[2], a02[2];  // per-segment tiles (row 6 = g_c / cost keeps accumulating)
    double a11[2], a12[2], a22[2];  // per-camera tiles
};

// V_m, g_m (reductions), W_{c,m} (plain stores: this warp owns the whole segment); resets what it flushed.
__device__ __forceinline__ void flush_segment(NeAcc& A, int lane, int64_t seg, int m, double* __restrict__ V,
                                              double* __restrict__ gp, double* __restrict__ W)
{
    const int row = lane >> 2, cp = 2 * (lane & 3);
    if (row < 6) {
        double* Vm = V + (int64_t)m * 36 + row * 6;
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (cp + i < 6) atomicAdd(Vm + cp + i, A.a00[i]);
        double* Ws = W + seg * 90 + row;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            Ws[(cp + i) * 6] = A.a01[i];
            if (8 + cp + i < 15) Ws[(8 + cp + i) * 6] = A.a02[i];
        }
        A.a00[0] = A.a00[1] = A.a01[0] = A.a01[1] = A.a02[0] = A.a02[1] = 0.0;
    } else if (row == 6) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (cp + i < 6) {
                atomicAdd(gp + (int64_t)m * 6 + cp + i, A.a00[i]);
                A.a00[i] = 0.0;
            }
    } else {
        A.a00[0] = A.a00[1] = A.a01[0] = A.a01[1] = A.a02[0] = A.a02[1] = 0.0;
    }
}

// U_c (both triangles) and g_c
__device__ __forceinline__ void flush_camera(NeAcc& A, int lane, int c, double* __restrict__ U, double* __restrict__ gc)
{
    const int row = lane >> 2, cp = 2 * (lane & 3);
    double* Uc = U + (int64_t)c * 225;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int col = cp + i;
        atomicAdd(Uc + row * 15 + col, A.a11[i]);
        if (8 + col < 15) {
            atomicAdd(Uc + row * 15 + 8 + col, A.a12[i]);
            atomicAdd(Uc + (8 + col) * 15 + row, A.a12[i]);
            if (row < 7) atomicAdd(Uc + (8 + row) * 15 + 8 + col, A.a22[i]);
        }
        if (row == 6) {
            atomicAdd(gc + (int64_t)c * 15 + col, A.a01[i]);
            if (8 + col < 15) atomicAdd(gc + (int64_t)c * 15 + 8 + col, A.a02[i]);
            A.a01[i] = A.a02[i] = 0.0;
        }
        A.a11[i] = A.a12[i] = A.a22[i] = 0.0;
    }
}

__global__ void __launch_bounds__(NE_WARPS * 32, 3)
k_normal_v2(int64_t N, int64_t n_seg, int n_warps, const int32_t* __restrict__ s_key, const double2* __restrict__ s_uv,
            const int32_t* __restrict__ s_seg, const int64_t* __restrict__ seg_start, const int32_t* __restrict__ seg_cam,
            const int32_t* __restrict__ seg_pose, const double* __restrict__ camtab, const double* __restrict__ posetab,
            const double* __restrict__ pts, double* __restrict__ U, double* __restrict__ gc, double* __restrict__ cost,
            double* __restrict__ V, double* __restrict__ gp, double* __restrict__ W)
{
    extern __shared__ __align__(16) double ne_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* ws = ne_smem + warp * NE_WARP_DOUBLES;
    const int wg = blockIdx.x * NE_WARPS + warp;
    if (wg >= n_warps) return;

    // this warp's segment range: [lower_bound(seg_start, wg N / n_warps), lower_bound(seg_start, (wg + 1) N / n_warps))
    int64_t sb, se;
    {
        const int64_t t0 = N * wg / n_warps, t1 = N * (wg + 1) / n_warps;
        int64_t lo = 0, hi = n_seg;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (seg_start[mid] < t0) lo = mid + 1; else hi = mid; }
        sb = lo;
        hi = n_seg;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (seg_start[mid] < t1) lo = mid + 1; else hi = mid; }
        se = lo;
    }
    if (sb >= se) return;
    const int64_t begin = seg_start[sb], end = seg_start[se];

    NeAcc A;
    A.a00[0] = A.a00[1] = A.a01[0] = A.a01[1] = A.a02[0] = A.a02[1] = 0.0;
    A.a11[0] = A.a11[1] = A.a12[0] = A.a12[1] = A.a22[0] = A.a22[1] = 0.0;
    int64_t cur_seg = -1;
    int cur_c = -1, cur_m = -1;

    const int g8 = lane >> 2, j = lane & 3, jb = j >> 1, jr = j & 1;

    for (int64_t base = begin; base < end; base += 32) {
        const int64_t i = base + lane;
        const int cnt = (int)min((int64_t)32, end - base);
        int seg = -1;
        if (lane < cnt) {
            seg = s_seg[i];
            const int c = seg_cam[seg], m = seg_pose[seg];
            const int k = s_key[i];
            const double2 o = s_uv[i];
            const double* pt = pts + 3 * (int64_t)k;
            const double Xt[3] = {pt[0], pt[1], pt[2]};
            const double* ct = camtab + (int64_t)c * CAM_STRIDE;
            const double* ptab = posetab + (int64_t)m * POSE_STRIDE;
            double res[2], Xw[3];
            ObsJac J;
            eval_obs(ct, ptab, Xt, o.x, o.y, res, J, Xw);
            {
                const double t0[8] = {J.Bm[0], J.Bm[1], J.Bm[2], J.N[0], J.N[1], J.N[2], res[0], 0.0};
                const double t1[8] = {J.xD, 1.0, 0.0, 0.0, J.Au[0], J.Au[1], J.Au[2], J.Au[3]};
                const double t2[8] = {J.Au[4], J.Bc[0], J.Bc[1], J.Bc[2], J.Pm[0], J.Pm[1], J.Pm[2], 0.0};
                stage_store_row(ws, lane, 0, t0, t1, t2);
            }
            {
                const double t0[8] = {J.Bm[3], J.Bm[4], J.Bm[5], J.N[3], J.N[4], J.N[5], res[1], 0.0};
                const double t1[8] = {0.0, 0.0, J.yD, 1.0, J.Av[0], J.Av[1], J.Av[2], J.Av[3]};
                const double t2[8] = {J.Av[4], J.Bc[3], J.Bc[4], J.Bc[5], J.Pm[3], J.Pm[4], J.Pm[5], 0.0};
                stage_store_row(ws, lane, 1, t0, t1, t2);
            }
        }
        // piece heads: lanes whose segment differs from the previous lane's
        const int prev = __shfl_up_sync(0xffffffffu, seg, 1);
        unsigned heads = __ballot_sync(0xffffffffu, lane < cnt && (lane == 0 || seg != prev));
        __syncwarp();
        while (heads) {
            const int a = __ffs(heads) - 1;
            heads &= heads - 1;
            const int b = heads ? __ffs(heads) - 1 : cnt;
            const int sp = __shfl_sync(0xffffffffu, seg, a);
            if (sp != cur_seg) {
                if (cur_seg >= 0) flush_segment(A, lane, cur_seg, cur_m, V, gp, W);
                const int nc = seg_cam[sp];
                if (nc != cur_c && cur_c >= 0) flush_camera(A, lane, cur_c, U, gc);
                cur_seg = sp; cur_c = nc; cur_m = seg_pose[sp];
            }
            const int k0 = a >> 1, k1 = (b - 1) >> 1;
            for (int ks = k0; ks <= k1; ++ks) {
                const int o = 2 * ks + jb;
                const int off = stage_row(o, jr) + ((g8 + stage_rot(o)) & 7);
                double v0 = ws[off], v1 = ws[NE_TILE_DOUBLES + off], v2 = ws[2 * NE_TILE_DOUBLES + off];
                if (ks == k0 || ks == k1) {  // boundary k-steps may hold rows of a neighbouring segment (or stale rows)
                    const bool ok = o >= a && o < b;
                    v0 = ok ? v0 : 0.0; v1 = ok ? v1 : 0.0; v2 = ok ? v2 : 0.0;
                }
                dmma884(A.a00[0], A.a00[1], v0, v0);
                dmma884(A.a01[0], A.a01[1], v0, v1);
                dmma884(A.a02[0], A.a02[1], v0, v2);
                dmma884(A.a11[0], A.a11[1], v1, v1);
                dmma884(A.a12[0], A.a12[1], v1, v2);
                dmma884(A.a22[0], A.a22[1], v2, v2);
            }
        }
        __syncwarp();
    }
    if (cur_seg >= 0) {
        flush_segment(A, lane, cur_seg, cur_m, V, gp, W);
        flush_camera(A, lane, cur_c, U, gc);
        if (lane == 27) atomicAdd(cost, A.a00[0]);  // row 6, column 6 of tile (0,0) = r . r
    }
}

int launch_normal_blocks_v1(pcs_problem* p);

int launch_normal_blocks(pcs_problem* p)
{
    static const int use_v1 = [] { const char* e = std::getenv("PCS_NE_KERNEL"); return e && e[0] == 'v' && e[1] == '1'; }();
    if (use_v1) return launch_normal_blocks_v1(p);
    // zero what is accumulated with reductions: [U | gc | cost | pad | V | gp]; W is fully overwritten
    const int64_t zero_doubles = (p->V - p->ne) + (int64_t)p->M * 42;
    PCS_CUDA(cudaMemsetAsync(p->ne, 0, (size_t)zero_doubles * sizeof(double), p->stream));
    if (p->N == 0) return PCS_OK;
    static bool attr_set = false;
    const size_t smem = (size_t)NE_WARPS * NE_WARP_DOUBLES * sizeof(double);
    if (!attr_set) {
        PCS_CUDA(cudaFuncSetAttribute(k_normal_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    // persistent-style grid: 3 CTAs of 4 warps per SM; at least ~64 observations per warp
    int64_t n_warps = std::min<int64_t>((p->N + 63) / 64, (int64_t)p->sm_count * 3 * NE_WARPS);
    n_warps = std::max<int64_t>(1, std::min<int64_t>(n_warps, p->n_seg));
    const int grid = (int)((n_warps + NE_WARPS - 1) / NE_WARPS);
    if (p->timing) PCS_CUDA(cudaEventRecord(p->ev_a, p->stream));
    k_normal_v2<<<grid, NE_WARPS * 32, smem, p->stream>>>(p->N, p->n_seg, (int)n_warps, p->s_key, (const double2*)p->s_uv,
                                                         p->s_seg, p->seg_start, p->seg_cam, p->seg_pose, p->camtab,
                                                         p->posetab, p->chain == PCS_CHAIN_TEMPLATE ? p->tmpl : nullptr,
                                                         p->U, p->gc, p->cost, p->V, p->gp, p->W);
    PCS_CUDA(cudaGetLastError());
    if (p->timing) PCS_CUDA(cudaEventRecord(p->ev_b, p->stream));
    return PCS_OK;
}

}  // namespace pcs
