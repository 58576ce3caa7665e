// pcs_normal.cu -- K_ne: fused residual + analytic Jacobian + J^T J / J^T r (template chain), the north-star kernel.
//
// No reference counterpart computes J^T J (scipy's LSMR consumes the CSR Jacobian, optimisation_handling.py:88-98);
// the per-observation math is the reference chain projection + extrinsic3D + template_points
// (function_block_implementations.py:21-211, matmul_map.py:147-243) as restated in pcs_math.cuh.
//
// Design (B200, FP64):
//   * Observations are sorted by (camera, pose) at problem build; a "segment" is one (camera, pose) pair.
//   * A warp owns a contiguous range of WHOLE segments (balanced by observation count) and walks it in batches
//     of 32 observations (inputs of batch n+1 are prefetched while batch n runs), one lane per observation: the
//     lane evaluates residual and Jacobian in registers and parks its two augmented rows
//         J' = [cam(0..7) | cam(8..14) r]        (16 doubles each)
//     in shared memory -- 8 KB per warp, swizzled so that both the 16-byte row stores and the fragment loads below
//     are bank-conflict free.
//   * Only the CAMERA columns are accumulated per observation.  Inside one segment the six pose columns are the
//     six extrinsic columns times a constant adjoint,  [w_m | n] = [w_c | p] T,  T = [[R_c, 0], [[R_c t_m]x R_c, R_c]]
//     (rows in the tangent parametrisation, pcs_math.cuh), so V_m, g_m and W_{c,m} follow from the SEGMENT's camera
//     Gram matrix G_s at flush time:   W_s = G_s[:, 9:15] T,   V_m += T^T G_s[9:15, 9:15] T,   g_m += T^T g_s[9:15].
//     That halves the per-observation Gram work: 16 columns = exactly two 8-column tiles.
//   * The Gram update  G += J'^T J'  runs on the FP64 tensor path: per 2 observations (4 rows = one k-step) the warp
//     loads 2 fragment values per lane and issues 3 DMMA m8n8k4 (tile pairs AA, AB, BB).  A-fragment and B-fragment
//     of a tile are the same register.
//   * Two accumulator sets: the segment's (flushed per segment through a 2 KB shared scratch: W by coalesced plain
//     stores -- the warp owns the segment -- V/g_m by FP64 reductions) and the camera's (U_c, g_c, r.r; flushed when
//     the camera changes).
//   * k_normal_epilogue applies the SO(3) left Jacobians once per block (tangent -> rvec parametrisation).
//   HBM traffic per observation: (u,v) 16 B + camera, pose, key 12 B = 28 B read; outputs are O(segments).
#include <algorithm>
#include <cstdlib>

#include "pcs_internal.cuh"
#include "pcs_math.cuh"

namespace pcs {

constexpr int NE_WARPS = 4;                  // warps per CTA
constexpr int NE_TILE_DOUBLES = 64 * 8;      // one 8-column tile of the 64 staged rows
constexpr int NE_SCRATCH_DOUBLES = 16 * 16 + 36 + 36;   // G_s | T | E T
constexpr int NE_WARP_DOUBLES = 2 * NE_TILE_DOUBLES + NE_SCRATCH_DOUBLES;

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// Shared-memory position of staged row (observation slot o in 0..31, row r in {u, v}) and column c of one tile:
//   word(o, r, c) = 8 * rho(o, r) + (c ^ rot(o)),   rho = 2 o + (r ^ bit1(o)),   rot = 4 bit0(o) + 2 bit2(o)  (even).
// rho is a bijection onto 0..63.  With this swizzle
//   - the 16-byte stores of a quarter warp (8 consecutive o, fixed r and column pair) hit 8 distinct 16-byte banks;
//   - the 8-byte fragment loads of a half warp (4 rows of one k-step x 4 columns) hit 16 distinct 8-byte banks.
// For the fragment loads of k-step ks (rows o = 2 ks + jb, r = jr; column g8) this collapses to
//   word = 32 ks + (L ^ X(ks)),   L = 16 jb + 8 jr + (g8 ^ 4 jb),   X(ks) = 8 bit0(ks) + 2 bit1(ks).
__device__ __forceinline__ int stage_row(int o, int r) { return (2 * o + (r ^ ((o >> 1) & 1))) * 8; }
__device__ __forceinline__ int stage_rot(int o) { return 4 * (o & 1) + 2 * ((o >> 2) & 1); }

__device__ __forceinline__ void stage_store_row(double* __restrict__ row, int rot, const double t0[8], const double t1[8])
{
#pragma unroll
    for (int c = 0; c < 8; c += 2) {
        const int pos = c ^ rot;
        *reinterpret_cast<double2*>(row + pos) = make_double2(t0[c], t0[c + 1]);
        *reinterpret_cast<double2*>(row + NE_TILE_DOUBLES + pos) = make_double2(t1[c], t1[c + 1]);
    }
}

// Fragment (m8n8 accumulator) layout: lane holds rows lane >> 2, columns 2 (lane & 3) + {0, 1} of an 8 x 8 tile.
// Column order of the 16 accumulated columns: 0..14 = camera (9 intrinsic, 3 rotation (tangent), 3 translation), 15 = r.
struct NeAcc {
    double aa[2], ab[2], bb[2];
};

__device__ __forceinline__ void acc_zero(NeAcc& A) { A.aa[0] = A.aa[1] = A.ab[0] = A.ab[1] = A.bb[0] = A.bb[1] = 0.0; }

// Segment flush.  S = this segment's camera Gram matrix (fragments), C = the running camera sums.
// scratch: G[16][16] | T[6][6] | ET[6][6].
__device__ __forceinline__ void flush_segment(NeAcc& S, NeAcc& C, int lane, int64_t seg, int c, int m,
                                              const double* __restrict__ camtab, const double* __restrict__ posetab,
                                              double* __restrict__ scratch, double* __restrict__ V, double* __restrict__ gp,
                                              double* __restrict__ W)
{
    double* const G = scratch;
    double* const T = scratch + 256;
    double* const ET = scratch + 292;
    const int row = lane >> 2, cp = 2 * (lane & 3);
    // (1) the segment's full symmetric 16 x 16 matrix; fold the segment into the camera sums
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int col = cp + i;
        G[row * 16 + col] = S.aa[i];
        G[row * 16 + 8 + col] = S.ab[i];
        G[(8 + col) * 16 + row] = S.ab[i];
        G[(8 + row) * 16 + 8 + col] = S.bb[i];
        C.aa[i] += S.aa[i]; C.ab[i] += S.ab[i]; C.bb[i] += S.bb[i];
    }
    acc_zero(S);
    // (2) adjoint T = [[R_c, 0], [[s]x R_c, R_c]], s = R_c t_m
    if (lane < 9) {
        const double* Rc = camtab + (int64_t)c * CAM_STRIDE + CAM_R;
        const double* tm = posetab + (int64_t)m * POSE_STRIDE + POSE_T;
        const int i = lane / 3, j = lane - 3 * i;
        const double t0 = tm[0], t1 = tm[1], t2 = tm[2];
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
        // ([s]x R_c)[i][j] = s_{i+1} R[i+2][j] - s_{i+2} R[i+1][j]  with cyclic indices
        const double s1 = Rc[3 * i1] * t0 + Rc[3 * i1 + 1] * t1 + Rc[3 * i1 + 2] * t2;
        const double s2 = Rc[3 * i2] * t0 + Rc[3 * i2 + 1] * t1 + Rc[3 * i2 + 2] * t2;
        const double r = Rc[3 * i + j];
        T[i * 6 + j] = r;
        T[i * 6 + 3 + j] = 0.0;
        T[(3 + i) * 6 + j] = s1 * Rc[3 * i2 + j] - s2 * Rc[3 * i1 + j];
        T[(3 + i) * 6 + 3 + j] = r;
    }
    __syncwarp();
    // (3) W_s = G[0:15, 9:15] T  (90 entries, stored with consecutive addresses);  ET = G[9:15, 9:15] T
    double* Ws = W + seg * 90;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const int idx = lane + 32 * t;
        if (idx < 90) {
            const int a = idx / 6, j = idx - 6 * a;
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) v = fma(G[a * 16 + 9 + k], T[k * 6 + j], v);
            Ws[idx] = v;
        }
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int idx = lane + 32 * t;
        if (idx < 36) {
            const int i = idx / 6, j = idx - 6 * i;
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) v = fma(G[(9 + i) * 16 + 9 + k], T[k * 6 + j], v);
            ET[idx] = v;
        }
    }
    __syncwarp();
    // (4) V_m += T^T ET,  g_m += T^T G[9:15, 15]
    double* Vm = V + (int64_t)m * 36;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int idx = lane + 32 * t;
        if (idx < 36) {
            const int i = idx / 6, j = idx - 6 * i;
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) v = fma(T[k * 6 + i], ET[k * 6 + j], v);
            atomicAdd(Vm + idx, v);
        } else if (idx < 42) {
            const int j = idx - 36;
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) v = fma(T[k * 6 + j], G[(9 + k) * 16 + 15], v);
            atomicAdd(gp + (int64_t)m * 6 + j, v);
        }
    }
    __syncwarp();  // scratch is reused by the next flush
}

// Camera flush: U_c (both triangles), g_c and r.r from the running camera sums
__device__ __forceinline__ void flush_camera(NeAcc& C, int lane, int c, double* __restrict__ U, double* __restrict__ gc,
                                             double* __restrict__ cost)
{
    const int row = lane >> 2, cp = 2 * (lane & 3);
    double* Uc = U + (int64_t)c * 225;
    double* gcc = gc + (int64_t)c * 15;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int col = cp + i;
        atomicAdd(Uc + row * 15 + col, C.aa[i]);
        if (col < 7) {
            atomicAdd(Uc + row * 15 + 8 + col, C.ab[i]);
            atomicAdd(Uc + (8 + col) * 15 + row, C.ab[i]);
            if (row < 7) atomicAdd(Uc + (8 + row) * 15 + 8 + col, C.bb[i]);
        } else {
            atomicAdd(gcc + row, C.ab[i]);
            if (row < 7) atomicAdd(gcc + 8 + row, C.bb[i]);
            else atomicAdd(cost, C.bb[i]);
        }
    }
    acc_zero(C);
}

template <int CTAS_PER_SM>
__global__ void __launch_bounds__(NE_WARPS * 32, CTAS_PER_SM)
k_normal(int n_warps, const int64_t* __restrict__ warp_seg, const int32_t* __restrict__ s_cam, const int32_t* __restrict__ s_pose,
         const int32_t* __restrict__ s_key, const double2* __restrict__ s_uv, const int64_t* __restrict__ seg_start,
         const double* __restrict__ camtab, const double* __restrict__ posetab, const double* __restrict__ pts,
         double* __restrict__ U, double* __restrict__ gc, double* __restrict__ cost, double* __restrict__ V,
         double* __restrict__ gp, double* __restrict__ W)
{
    extern __shared__ __align__(16) double ne_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* ws = ne_smem + warp * NE_WARP_DOUBLES;
    const int wg = blockIdx.x * NE_WARPS + warp;
    if (wg >= n_warps) return;

    // this warp's range of whole segments (k_warp_ranges)
    const int64_t sb = warp_seg[wg], se = warp_seg[wg + 1];
    if (sb >= se) return;
    const int64_t begin = seg_start[sb], end = seg_start[se];

    NeAcc S, C;   // segment / camera accumulators
    acc_zero(S);
    acc_zero(C);
    double* const scratch = ws + 2 * NE_TILE_DOUBLES;
    int64_t cur_seg = sb - 1;   // segments are visited in order: a piece head advances this counter
    int cur_c = -1, cur_m = -1;
    int last_c = -1, last_m = -1;  // (camera, pose) of the last observation of the previous batch

    // lane constants of the staging layout
    double* const st_u = ws + stage_row(lane, 0);
    double* const st_v = ws + stage_row(lane, 1);
    const int st_rot = stage_rot(lane);
    const int g8 = lane >> 2, jb = (lane >> 1) & 1, jr = lane & 1;
    const int ld_L = 16 * jb + 8 * jr + (g8 ^ (4 * jb));

    // software prefetch: the inputs of the next batch are requested before the current batch is evaluated
    int c_n = -1, m_n = -1, k_n = 0;
    double2 uv_n = make_double2(0.0, 0.0);
    if (begin + lane < end) {
        c_n = s_cam[begin + lane]; m_n = s_pose[begin + lane]; k_n = s_key[begin + lane]; uv_n = s_uv[begin + lane];
    }

    for (int64_t base = begin; base < end; base += 32) {
        const int cnt = (int)min((int64_t)32, end - base);
        const int c = c_n, m = m_n, k = k_n;
        const double2 o = uv_n;
        {
            const int64_t i = base + 32 + lane;
            if (i < end) { c_n = s_cam[i]; m_n = s_pose[i]; k_n = s_key[i]; uv_n = s_uv[i]; }
        }
        if (lane < cnt) {
            const double* pt = pts + 3 * (int64_t)k;
            const double Xt[3] = {pt[0], pt[1], pt[2]};
            const double* ct = camtab + (int64_t)c * CAM_STRIDE;
            const double* ptab = posetab + (int64_t)m * POSE_STRIDE;
            double res[2];
            ObsJacCam J;
            eval_obs_cam(ct, ptab, Xt, o.x, o.y, res, J);
            {
                const double t0[8] = {J.xD, 1.0, 0.0, 0.0, J.Au[0], J.Au[1], J.Au[2], J.Au[3]};
                const double t1[8] = {J.Au[4], J.Wc[0], J.Wc[1], J.Wc[2], J.Pm[0], J.Pm[1], J.Pm[2], res[0]};
                stage_store_row(st_u, st_rot, t0, t1);
            }
            {
                const double t0[8] = {0.0, 0.0, J.yD, 1.0, J.Av[0], J.Av[1], J.Av[2], J.Av[3]};
                const double t1[8] = {J.Av[4], J.Wc[3], J.Wc[4], J.Wc[5], J.Pm[3], J.Pm[4], J.Pm[5], res[1]};
                stage_store_row(st_v, st_rot, t0, t1);
            }
        }
        // piece heads: lanes whose (camera, pose) differs from the previous observation's
        int pc = __shfl_up_sync(0xffffffffu, c, 1), pm = __shfl_up_sync(0xffffffffu, m, 1);
        if (lane == 0) { pc = last_c; pm = last_m; }
        const unsigned heads = __ballot_sync(0xffffffffu, lane < cnt && (c != pc || m != pm));
        last_c = __shfl_sync(0xffffffffu, c, cnt - 1);
        last_m = __shfl_sync(0xffffffffu, m, cnt - 1);
        __syncwarp();
        unsigned pieces = heads | 1u;  // a batch may start in the middle of a segment
        while (pieces) {
            const int a = __ffs(pieces) - 1;
            pieces &= pieces - 1;
            const int b = pieces ? __ffs(pieces) - 1 : cnt;
            if ((heads >> a) & 1u) {
                if (cur_c >= 0) flush_segment(S, C, lane, cur_seg, cur_c, cur_m, camtab, posetab, scratch, V, gp, W);
                const int nc = __shfl_sync(0xffffffffu, c, a);
                if (nc != cur_c && cur_c >= 0) flush_camera(C, lane, cur_c, U, gc, cost);
                ++cur_seg; cur_c = nc; cur_m = __shfl_sync(0xffffffffu, m, a);
            }
            const int k0 = a >> 1, k1 = (b - 1) >> 1;
            for (int ks = k0; ks <= k1; ++ks) {
                const double* f = ws + 32 * ks + (ld_L ^ (((ks & 1) << 3) | (ks & 2)));
                double v0 = f[0], v1 = f[NE_TILE_DOUBLES];
                if (ks == k0 || ks == k1) {  // boundary k-steps may hold rows of a neighbouring segment (or stale rows)
                    const int ob = 2 * ks + jb;
                    const bool ok = ob >= a && ob < b;
                    v0 = ok ? v0 : 0.0; v1 = ok ? v1 : 0.0;
                }
                dmma884(S.aa[0], S.aa[1], v0, v0);
                dmma884(S.ab[0], S.ab[1], v0, v1);
                dmma884(S.bb[0], S.bb[1], v1, v1);
            }
        }
        __syncwarp();
    }
    if (cur_c >= 0) {
        flush_segment(S, C, lane, cur_seg, cur_c, cur_m, camtab, posetab, scratch, V, gp, W);
        flush_camera(C, lane, cur_c, U, gc, cost);
    }
}

// Segment range of every warp: [lower_bound(seg_start, w N / n_warps), lower_bound(seg_start, (w + 1) N / n_warps)).
// Depends only on the static problem structure; evaluated once per (problem, n_warps).
__global__ void k_warp_ranges(int64_t N, int64_t n_seg, int n_warps, const int64_t* __restrict__ seg_start,
                              int64_t* __restrict__ warp_seg)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w > n_warps) return;
    const int64_t target = N * w / n_warps;
    int64_t lo = 0, hi = n_seg;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (seg_start[mid] < target) lo = mid + 1; else hi = mid; }
    warp_seg[w] = w == n_warps ? n_seg : lo;
}

// Epilogue: the kernel above accumulates the rotation rows / columns in the tangent parametrisation; this maps the
// blocks to the reference's rvec parametrisation, B = T^T B' T with T_c = diag(I9, Jl_c, I3), T_m = diag(Jl_m, I3).
//   blocks [0, C)            : U_c, g_c   (one block of 64 threads per camera)
//   blocks [C, C + Pb)       : V_m, g_m   (one thread per pose)
//   blocks [C + Pb, ...)     : W_s        (15 lanes per segment, 2 segments per warp)
__global__ void __launch_bounds__(64)
k_normal_epilogue(int C, int M, int64_t n_seg, int pose_blocks, const int32_t* __restrict__ seg_cam,
                  const int32_t* __restrict__ seg_pose, const double* __restrict__ camtab, const double* __restrict__ posetab,
                  double* __restrict__ U, double* __restrict__ gc, double* __restrict__ V, double* __restrict__ gp,
                  double* __restrict__ W)
{
    const int t = threadIdx.x;
    if ((int)blockIdx.x < C) {
        const int c = blockIdx.x;
        __shared__ double u[225];
        __shared__ double jl[9];
        double* Uc = U + (int64_t)c * 225;
        for (int e = t; e < 225; e += 64) u[e] = Uc[e];
        if (t < 9) jl[t] = camtab[(int64_t)c * CAM_STRIDE + CAM_JL + t];
        __syncthreads();
        if (t < 15) {  // columns 9..11 of row t  <-  row * Jl
            const double a0 = u[t * 15 + 9], a1 = u[t * 15 + 10], a2 = u[t * 15 + 11];
#pragma unroll
            for (int i = 0; i < 3; ++i) u[t * 15 + 9 + i] = a0 * jl[i] + a1 * jl[3 + i] + a2 * jl[6 + i];
        } else if (t == 15) {  // g_c[9..11] <- Jl^T g
            double* g = gc + (int64_t)c * 15 + 9;
            const double a0 = g[0], a1 = g[1], a2 = g[2];
#pragma unroll
            for (int i = 0; i < 3; ++i) g[i] = a0 * jl[i] + a1 * jl[3 + i] + a2 * jl[6 + i];
        }
        __syncthreads();
        if (t < 15) {  // rows 9..11 of column t  <-  Jl^T * column
            const double a0 = u[9 * 15 + t], a1 = u[10 * 15 + t], a2 = u[11 * 15 + t];
#pragma unroll
            for (int i = 0; i < 3; ++i) u[(9 + i) * 15 + t] = a0 * jl[i] + a1 * jl[3 + i] + a2 * jl[6 + i];
        }
        __syncthreads();
        for (int e = t; e < 225; e += 64) Uc[e] = u[e];
        return;
    }
    if ((int)blockIdx.x < C + pose_blocks) {
        const int m = ((int)blockIdx.x - C) * 64 + t;
        if (m >= M) return;
        const double* jl = posetab + (int64_t)m * POSE_STRIDE + POSE_JL;
        const double j[9] = {jl[0], jl[1], jl[2], jl[3], jl[4], jl[5], jl[6], jl[7], jl[8]};
        double* Vm = V + (int64_t)m * 36;
        double v[36];
#pragma unroll
        for (int e = 0; e < 36; ++e) v[e] = Vm[e];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const double a0 = v[r * 6], a1 = v[r * 6 + 1], a2 = v[r * 6 + 2];
#pragma unroll
            for (int i = 0; i < 3; ++i) v[r * 6 + i] = a0 * j[i] + a1 * j[3 + i] + a2 * j[6 + i];
        }
#pragma unroll
        for (int cidx = 0; cidx < 6; ++cidx) {
            const double a0 = v[cidx], a1 = v[6 + cidx], a2 = v[12 + cidx];
#pragma unroll
            for (int i = 0; i < 3; ++i) v[i * 6 + cidx] = a0 * j[i] + a1 * j[3 + i] + a2 * j[6 + i];
        }
#pragma unroll
        for (int e = 0; e < 36; ++e) Vm[e] = v[e];
        double* g = gp + (int64_t)m * 6;
        const double a0 = g[0], a1 = g[1], a2 = g[2];
#pragma unroll
        for (int i = 0; i < 3; ++i) g[i] = a0 * j[i] + a1 * j[3 + i] + a2 * j[6 + i];
        return;
    }
    // W_s (15 x 6): lane a of a 15-lane group owns row a
    const int64_t wblock = (int64_t)blockIdx.x - C - pose_blocks;
    const int lane = t & 31, grp = lane / 15, a = lane % 15;
    const int64_t s = (wblock * 2 + (t >> 5)) * 2 + grp;
    const bool active = grp < 2 && s < n_seg;
    double r[6] = {0, 0, 0, 0, 0, 0};
    const int k = a - 9;  // 0..2 for the camera-rotation rows
    double jk0 = 0.0, jk1 = 0.0, jk2 = 0.0;  // column k of Jl_c
    if (active) {
        const double* jm = posetab + (int64_t)seg_pose[s] * POSE_STRIDE + POSE_JL;
        const double* Ws = W + s * 90 + a * 6;
#pragma unroll
        for (int i = 0; i < 6; ++i) r[i] = Ws[i];
        const double a0 = r[0], a1 = r[1], a2 = r[2];
#pragma unroll
        for (int i = 0; i < 3; ++i) r[i] = a0 * jm[i] + a1 * jm[3 + i] + a2 * jm[6 + i];
        if (k >= 0 && k < 3) {
            const double* jcp = camtab + (int64_t)seg_cam[s] * CAM_STRIDE + CAM_JL;
            jk0 = jcp[k]; jk1 = jcp[3 + k]; jk2 = jcp[6 + k];
        }
    }
    // rows 9..11 mix: new row (9 + i) = sum_i' Jl_c[i'][i] row (9 + i')
    const int src0 = grp * 15 + 9;
    double out[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const double x0 = __shfl_sync(0xffffffffu, r[i], src0 & 31), x1 = __shfl_sync(0xffffffffu, r[i], (src0 + 1) & 31),
                     x2 = __shfl_sync(0xffffffffu, r[i], (src0 + 2) & 31);
        out[i] = (k >= 0 && k < 3) ? x0 * jk0 + x1 * jk1 + x2 * jk2 : r[i];
    }
    if (active) {
        double* Ws = W + s * 90 + a * 6;
#pragma unroll
        for (int i = 0; i < 6; ++i) Ws[i] = out[i];
    }
}

static int ensure_ranges(pcs_problem* p, int64_t n_ranges)
{
    if (p->ne_warps == n_ranges) return PCS_OK;
    if (p->warp_seg) cudaFree(p->warp_seg);
    p->warp_seg = nullptr;
    p->ne_warps = 0;
    PCS_CUDA(cudaMalloc((void**)&p->warp_seg, (size_t)(n_ranges + 1) * sizeof(int64_t)));
    k_warp_ranges<<<(int)((n_ranges + 128) / 128), 128, 0, p->stream>>>(p->N, p->n_seg, (int)n_ranges, p->seg_start, p->warp_seg);
    PCS_CUDA(cudaGetLastError());
    p->ne_warps = n_ranges;
    return PCS_OK;
}

int launch_normal_blocks(pcs_problem* p)
{
    // zero what is accumulated with reductions: [U | gc | cost | pad | V | gp]; W is fully overwritten
    const int64_t zero_doubles = (p->V - p->ne) + (int64_t)p->M * 42;
    PCS_CUDA(cudaMemsetAsync(p->ne, 0, (size_t)zero_doubles * sizeof(double), p->stream));
    if (p->N == 0) return PCS_OK;
    // resident CTAs per SM: 4 x 128 registers (default) or 3 x 154; PCS_NE_CTAS=3 selects the latter for A/B runs
    static const int ctas = [] { const char* e = std::getenv("PCS_NE_CTAS"); return e && e[0] == '3' ? 3 : 4; }();
    auto kern = ctas == 3 ? k_normal<3> : k_normal<4>;
    static bool attr_set = false;
    const size_t smem = (size_t)NE_WARPS * NE_WARP_DOUBLES * sizeof(double);
    if (!attr_set) {
        PCS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    // persistent-style grid: `ctas` CTAs of NE_WARPS warps per SM; at least ~64 observations per warp
    int64_t n_warps = std::min<int64_t>((p->N + 63) / 64, (int64_t)p->sm_count * ctas * NE_WARPS);
    n_warps = std::max<int64_t>(1, std::min<int64_t>(n_warps, p->n_seg));
    PCS_TRY(ensure_ranges(p, n_warps));
    const int grid = (int)((n_warps + NE_WARPS - 1) / NE_WARPS);
    if (p->timing) PCS_CUDA(cudaEventRecord(p->ev_a, p->stream));
    kern<<<grid, NE_WARPS * 32, smem, p->stream>>>((int)n_warps, p->warp_seg, p->s_cam, p->s_pose, p->s_key,
                                                  (const double2*)p->s_uv, p->seg_start, p->camtab, p->posetab, p->tmpl,
                                                  p->U, p->gc, p->cost, p->V, p->gp, p->W);
    PCS_CUDA(cudaGetLastError());
    const int pose_blocks = (p->M + 63) / 64;
    const int64_t w_blocks = (p->n_seg + 3) / 4;
    k_normal_epilogue<<<(unsigned)(p->C + pose_blocks + w_blocks), 64, 0, p->stream>>>(
        p->C, p->M, p->n_seg, pose_blocks, p->seg_cam, p->seg_pose, p->camtab, p->posetab, p->U, p->gc, p->V, p->gp, p->W);
    PCS_CUDA(cudaGetLastError());
    p->n_launches += 2;
    if (p->timing) PCS_CUDA(cudaEventRecord(p->ev_b, p->stream));
    return PCS_OK;
}

}  // namespace pcs
