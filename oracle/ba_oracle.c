/*
 * ba_oracle.c -- CPU restatement (plain C, FP64) of pyCamSet's bundle-adjustment inner loop.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (pycamset_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity pinning: the reference ships no entry-level golden vectors, so this
 * restatement is pinned against outputs of the reference itself, generated in the build
 * container by tests/golden/make_golden.py (residuals, CSR Jacobian, J^T J, J^T r on the ccube
 * fixture and on seeded synthetic rings) and committed under tests/golden/.
 *
 * Each function cites the reference file:line (relative to the pyCamSet repository) it follows.
 * The arithmetic deliberately mirrors the reference's order of operations (e.g. the z**7 form
 * of the projection Jacobian) so that agreement is at rounding level; the reference itself is
 * JIT-compiled with fastmath, so "bit-for-bit" is not defined for its outputs.
 *
 * Parameter string layout (abstract_function_blocks.py:777-820, :669-681):
 *   [ intr C x 9 | extr C x 6 | pose M x 6 | point K x 3 (self-calibration chain only) ]
 *   intr row = [fx, px, fy, py, k1, k2, p1, p2, k3]; extr / pose row = [rvec(3), t(3)].
 * Chains (the key is the tuple of block class names, abstract_function_blocks.py:297):
 *   chain 0: projection + extrinsic3D + template_points            (P = 21 columns per row)
 *   chain 1: projection + extrinsic3D + rigidTform3d + free_point  (P = 24 columns per row)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* compiled_helpers.py:197-235  numba_flat_rodrigues_INPLACE: rvec -> row-major 3x3 */
static void rodrigues(const double *r, double *R)
{
    double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (theta < 1e-10) {
        memset(R, 0, 9 * sizeof(double));
        R[0] = R[4] = R[8] = 1.0;
        return;
    }
    double scalar = 1.0 / theta;
    double s2 = scalar * scalar;
    double ct = cos(theta);
    double st = sin(theta) * scalar;
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j) {
            R[3 * i + j] = r[i] * r[j];
            R[3 * j + i] = r[i] * r[j];
        }
    double f = (1.0 - ct) * s2;
    for (int k = 0; k < 9; ++k) R[k] *= f;
    R[0] += ct; R[4] += ct; R[8] += ct;
    R[1] -= r[2] * st; R[3] += r[2] * st;
    R[2] += r[1] * st; R[6] -= r[1] * st;
    R[5] -= r[0] * st; R[7] += r[0] * st;
}

/* compiled_helpers.py:237-286  numba_rodrigues_jac: out[i*9 + k] = d R_k / d r_i */
static void rodrigues_jac(const double *r, double *out)
{
    double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (theta < 1e-10) {
        memset(out, 0, 27 * sizeof(double));
        out[5] = -1; out[15] = -1; out[19] = -1;
        out[7] = 1;  out[11] = 1;  out[21] = 1;
        return;
    }
    double i_theta = 1.0 / theta;
    double ct = cos(theta), ct_1 = 1.0 - ct, st = sin(theta);
    double x = r[0] * i_theta, y = r[1] * i_theta, z = r[2] * i_theta;
    double rrt[9] = {x * x, x * y, x * z, x * y, y * y, y * z, x * z, y * z, z * z};
    double r_x[9] = {0, -z, y, z, 0, -x, -y, x, 0};
    double eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double drrt[27] = {x + x, y, z, y, 0, 0, z, 0, 0,
                       0, x, 0, x, y + y, z, 0, z, 0,
                       0, 0, x, 0, 0, y, x, y, z + z};
    double d_r_x[27] = {0, 0, 0, 0, 0, -1, 0, 1, 0,
                        0, 0, 1, 0, 0, 0, -1, 0, 0,
                        0, -1, 0, 1, 0, 0, 0, 0, 0};
    double rv[3] = {x, y, z};
    for (int i = 0; i < 3; ++i) {
        double ri = rv[i];
        double a0 = -st * ri;
        double a1 = (st - 2 * ct_1 * i_theta) * ri;
        double a2 = ct_1 * i_theta;
        double a3 = (ct - st * i_theta) * ri;
        double a4 = st * i_theta;
        for (int k = 0; k < 9; ++k)
            out[i * 9 + k] = a0 * eye[k] + a1 * rrt[k] + a2 * drrt[i * 9 + k] + a3 * r_x[k] + a4 * d_r_x[i * 9 + k];
    }
}

/* function_block_implementations.py:150-155 rigidTform3d.compute_fun
 * (n_e4x4_flat_INPLACE compiled_helpers.py:288-301, n_htform_prealloc :357-370): y = R x + t */
static void rigid_fun(const double *p, const double *X, double *Y)
{
    double R[9];
    rodrigues(p, R);
    for (int a = 0; a < 3; ++a)
        Y[a] = X[0] * R[3 * a] + X[1] * R[3 * a + 1] + X[2] * R[3 * a + 2] + p[3 + a];
}

/* function_block_implementations.py:157-182 rigidTform3d.compute_jac:
 * Jr (3x3) = d(RX)/dr, Jx (3x3) = R; the translation block is I3.
 * template_points.compute_jac (:194-211) is the same without Jx. */
static void rigid_jac(const double *p, const double *X, double *Jr, double *Jx)
{
    double dR[27];
    rodrigues_jac(p, dR);
    for (int op = 0; op < 3; ++op)
        for (int ang = 0; ang < 3; ++ang)
            Jr[3 * op + ang] = dR[9 * ang + 3 * op + 0] * X[0] + dR[9 * ang + 3 * op + 1] * X[1] +
                               dR[9 * ang + 3 * op + 2] * X[2];
    rodrigues(p, Jx);
}

/* function_block_implementations.py:27-47 projection.compute_fun (pinhole + Brown-Conrady) */
static void proj_fun(const double *q, const double *X, double *out)
{
    double x = X[0], y = X[1], inv_z = 1.0 / X[2];
    double u = (q[0] * x + q[1] * X[2]) * inv_z;
    double v = (q[2] * y + q[3] * X[2]) * inv_z;
    const double *k = q + 4;
    x = (u - q[1]) / q[0];
    y = (v - q[3]) / q[2];
    double r2 = x * x + y * y;
    double kup = 1 + k[0] * r2 + k[1] * (r2 * r2) + k[4] * (r2 * r2 * r2);
    double xD = x * kup, yD = y * kup;
    xD += 2 * k[2] * x * y + k[3] * (r2 + 2 * (x * x));
    yD += k[2] * (r2 + 2 * (y * y)) + 2 * k[3] * x * y;
    out[0] = xD * q[0] + q[1];
    out[1] = yD * q[2] + q[3];
}

/* function_block_implementations.py:50-140 projection.compute_jac: 2 x 12 =
 * d(u,v)/d[fx,px,fy,py,k0,k1,p0,p1,k2 | x,y,z], in the reference's z**7 / z**8 form. */
static void proj_jac(const double *q, const double *X, double *o)
{
    double f_x = q[0], f_y = q[2], k_0 = q[4], k_1 = q[5], p_0 = q[6], p_1 = q[7], k_2 = q[8];
    double x = X[0], y = X[1], z = X[2];
    double x2 = x * x, y2 = y * y, rr = x2 + y2, rr2 = rr * rr, rr3 = rr2 * rr;
    double z2 = z * z, z3 = z2 * z, z4 = z2 * z2, z5 = z4 * z, z6 = z3 * z3, z7 = z6 * z, z8 = z4 * z4;
    double radn = k_0 * z4 * rr + k_1 * z2 * rr2 + k_2 * rr3 + z6;
    double drad = k_0 * z4 + 2 * k_1 * z2 * rr + 3 * k_2 * rr2;
    o[0] = (x * radn + z5 * (2 * p_0 * x * y + p_1 * (3 * x2 + y2))) / z7; /* du/dfx */
    o[1] = 1; o[2] = 0; o[3] = 0;
    o[4] = f_x * x * rr / z3;
    o[5] = f_x * x * rr2 / z5;
    o[6] = 2 * f_x * x * y / z2;
    o[7] = f_x * (3 * x2 + y2) / z2;
    o[8] = f_x * x * rr3 / z7;
    o[9] = f_x * (k_0 * z4 * rr + k_1 * z2 * rr2 + k_2 * rr3 + 2 * x2 * drad + z6 + 2 * z5 * (p_0 * y + 3 * p_1 * x)) / z7;
    o[10] = 2 * f_x * (x * y * drad + z5 * (p_0 * x + p_1 * y)) / z7;
    o[11] = -f_x * (4 * p_0 * x * y * z5 + 2 * p_1 * z5 * (3 * x2 + y2) + 2 * x * rr * drad + x * radn) / z8;
    o[12] = 0; o[13] = 0;
    o[14] = (y * radn + z5 * (p_0 * (x2 + 3 * y2) + 2 * p_1 * x * y)) / z7; /* dv/dfy */
    o[15] = 1;
    o[16] = f_y * y * rr / z3;
    o[17] = f_y * y * rr2 / z5;
    o[18] = f_y * (x2 + 3 * y2) / z2;
    o[19] = 2 * f_y * x * y / z2;
    o[20] = f_y * y * rr3 / z7;
    o[21] = 2 * f_y * (x * y * drad + z5 * (p_0 * x + p_1 * y)) / z7;
    o[22] = f_y * (k_0 * z4 * rr + k_1 * z2 * rr2 + k_2 * rr3 + 2 * y2 * drad + z6 + 2 * z5 * (3 * p_0 * y + p_1 * x)) / z7;
    o[23] = -f_y * (2 * p_0 * z5 * (x2 + 3 * y2) + 4 * p_1 * x * y * z5 + 2 * y * rr * drad + y * radn) / z8;
}

/*
 * One observation through the whole chain.
 * Forward: generated full_loss body (abstract_function_blocks.py:365-385): blocks run last -> first,
 * template point injected at :374-375, residual = projected - observed (:384).
 * Jacobian: generated full_jac body (:573-594) followed by matflow (matmul_map.py:147-243), i.e.
 *   J = [ A | Pm [D_c | I] | Pm R_c [D_m | I] ( | Pm R_c R_m ) ]     
 * J is written dense as 2 rows of P columns (P = 21 or 24), row-major; pass J == NULL to skip it.
 */
static void eval_obs(int chain, const double *intr, const double *extr, const double *pose,
                     const double *Xt, double u, double v, double *res, double *J)
{
    double Xw[3], Xc[3], uvp[2];
    rigid_fun(pose, Xt, Xw);
    rigid_fun(extr, Xw, Xc);
    proj_fun(intr, Xc, uvp);
    res[0] = uvp[0] - u;
    res[1] = uvp[1] - v;
    if (!J) return;
    const int P = chain == 0 ? 21 : 24;
    double pj[24], Dm[9], Rm[9], Dc[9], Rc[9];
    rigid_jac(pose, Xt, Dm, Rm);
    rigid_jac(extr, Xw, Dc, Rc);
    proj_jac(intr, Xc, pj);
    for (int row = 0; row < 2; ++row) {
        const double *a = pj + 12 * row;
        double *Jr = J + P * row;
        const double *pm = a + 9;
        for (int k = 0; k < 9; ++k) Jr[k] = a[k];
        /* extrinsic block: Pm [D_c | I] */
        for (int i = 0; i < 3; ++i) Jr[9 + i] = pm[0] * Dc[i] + pm[1] * Dc[3 + i] + pm[2] * Dc[6 + i];
        for (int i = 0; i < 3; ++i) Jr[12 + i] = pm[i];
        /* N = Pm R_c */
        double n[3];
        for (int i = 0; i < 3; ++i) n[i] = pm[0] * Rc[i] + pm[1] * Rc[3 + i] + pm[2] * Rc[6 + i];
        for (int i = 0; i < 3; ++i) Jr[15 + i] = n[0] * Dm[i] + n[1] * Dm[3 + i] + n[2] * Dm[6 + i];
        for (int i = 0; i < 3; ++i) Jr[18 + i] = n[i];
        if (chain == 1)
            for (int i = 0; i < 3; ++i) Jr[21 + i] = n[0] * Rm[i] + n[1] * Rm[3 + i] + n[2] * Rm[6 + i];
    }
}

typedef struct {
    int chain, C, M, K;
    int64_t N;
    const int32_t *cam, *pose, *key;
    const double *uv;       /* N x 2 interleaved */
    const double *params;   /* full parameter string */
    const double *tmpl;     /* K x 3 template points (chain 0) */
} problem_t;

static inline const double *obs_point(const problem_t *p, int64_t i)
{
    if (p->chain == 0) return p->tmpl + 3 * (int64_t)p->key[i];
    return p->params + 15 * (int64_t)p->C + 6 * (int64_t)p->M + 3 * (int64_t)p->key[i];
}

/* global column indices of observation i in the parameter string
 * (get_block_param_inds, abstract_function_blocks.py:192-233) */
static inline void obs_columns(const problem_t *p, int64_t i, int64_t *cols)
{
    int64_t c = p->cam[i], m = p->pose[i];
    for (int k = 0; k < 9; ++k) cols[k] = 9 * c + k;
    for (int k = 0; k < 6; ++k) cols[9 + k] = 9 * (int64_t)p->C + 6 * c + k;
    for (int k = 0; k < 6; ++k) cols[15 + k] = 15 * (int64_t)p->C + 6 * m + k;
    if (p->chain == 1)
        for (int k = 0; k < 3; ++k) cols[21 + k] = 15 * (int64_t)p->C + 6 * (int64_t)p->M + 3 * (int64_t)p->key[i] + k;
}

static problem_t mk(int chain, int64_t N, const int32_t *cam, const int32_t *pose, const int32_t *key,
                    const double *uv, int C, int M, int K, const double *params, const double *tmpl)
{
    problem_t p = {chain, C, M, K, N, cam, pose, key, uv, params, tmpl};
    return p;
}

ORACLE_API int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORACLE_API void oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* residual vector, interleaved (x, y) per observation in dd row order (template_handler.py:166-170) */
ORACLE_API int oracle_residual(int chain, int64_t N, const int32_t *cam, const int32_t *pose, const int32_t *key,
                               const double *uv, int C, int M, int K, const double *params, const double *tmpl,
                               double *r_out)
{
    problem_t p = mk(chain, N, cam, pose, key, uv, C, M, K, params, tmpl);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i)
        eval_obs(chain, params + 9 * (int64_t)cam[i], params + 9 * (int64_t)C + 6 * (int64_t)cam[i],
                 params + 15 * (int64_t)C + 6 * (int64_t)pose[i], obs_point(&p, i), uv[2 * i], uv[2 * i + 1],
                 r_out + 2 * i, NULL);
    return 0;
}

/* dense per-observation Jacobian rows (the reference's dense_output, abstract_function_blocks.py:561,
 * :597): J_out is (2N) x P row-major; r_out may be NULL */
ORACLE_API int oracle_jacobian_dense(int chain, int64_t N, const int32_t *cam, const int32_t *pose,
                                     const int32_t *key, const double *uv, int C, int M, int K,
                                     const double *params, const double *tmpl, double *J_out, double *r_out)
{
    problem_t p = mk(chain, N, cam, pose, key, uv, C, M, K, params, tmpl);
    const int P = chain == 0 ? 21 : 24;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        double r[2];
        eval_obs(chain, params + 9 * (int64_t)cam[i], params + 9 * (int64_t)C + 6 * (int64_t)cam[i],
                 params + 15 * (int64_t)C + 6 * (int64_t)pose[i], obs_point(&p, i), uv[2 * i], uv[2 * i + 1], r,
                 J_out + 2 * P * i);
        if (r_out) { r_out[2 * i] = r[0]; r_out[2 * i + 1] = r[1]; }
    }
    return 0;
}

/* CSR structure with fixed parameters removed (make_jac_CSR_columns_row_pointers,
 * abstract_function_blocks.py:465-489).  free_map[L]: free column index or -1.  Returns nnz.
 * col_idx may be NULL to only count. */
ORACLE_API int64_t oracle_csr_structure(int chain, int64_t N, const int32_t *cam, const int32_t *pose,
                                        const int32_t *key, int C, int M, int K, const int32_t *free_map,
                                        int64_t *col_idx, int64_t *row_ptr)
{
    problem_t p = mk(chain, N, cam, pose, key, NULL, C, M, K, NULL, NULL);
    const int P = chain == 0 ? 21 : 24;
    int64_t nnz = 0;
    if (row_ptr) row_ptr[0] = 0;
    for (int64_t i = 0; i < N; ++i) {
        int64_t cols[24];
        obs_columns(&p, i, cols);
        for (int row = 0; row < 2; ++row) {
            for (int k = 0; k < P; ++k) {
                int32_t f = free_map[cols[k]];
                if (f >= 0) {
                    if (col_idx) col_idx[nnz] = f;
                    ++nnz;
                }
            }
            if (row_ptr) row_ptr[2 * i + row + 1] = nnz;
        }
    }
    return nnz;
}

/* CSR values in the order of oracle_csr_structure (jac_fn, abstract_function_blocks.py:644-652) */
ORACLE_API int oracle_csr_values(int chain, int64_t N, const int32_t *cam, const int32_t *pose, const int32_t *key,
                                 const double *uv, int C, int M, int K, const double *params, const double *tmpl,
                                 const int32_t *free_map, const int64_t *row_ptr, double *vals)
{
    problem_t p = mk(chain, N, cam, pose, key, uv, C, M, K, params, tmpl);
    const int P = chain == 0 ? 21 : 24;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        double r[2], J[48];
        int64_t cols[24];
        eval_obs(chain, params + 9 * (int64_t)cam[i], params + 9 * (int64_t)C + 6 * (int64_t)cam[i],
                 params + 15 * (int64_t)C + 6 * (int64_t)pose[i], obs_point(&p, i), uv[2 * i], uv[2 * i + 1], r, J);
        obs_columns(&p, i, cols);
        for (int row = 0; row < 2; ++row) {
            int64_t w = row_ptr[2 * i + row];
            for (int k = 0; k < P; ++k)
                if (free_map[cols[k]] >= 0) vals[w++] = J[P * row + k];
        }
    }
    return 0;
}

/* Dense normal equations over the free parameters: JtJ (n_free x n_free, row-major, full symmetric),
 * Jtr (n_free), cost = r.r.  The reference never forms these; they are defined
 * as J.T @ J and J.T @ r of the reference CSR.  Intended for small problems. */
ORACLE_API int oracle_normal_dense(int chain, int64_t N, const int32_t *cam, const int32_t *pose,
                                   const int32_t *key, const double *uv, int C, int M, int K, const double *params,
                                   const double *tmpl, const int32_t *free_map, int64_t n_free, double *JtJ,
                                   double *Jtr, double *cost)
{
    problem_t p = mk(chain, N, cam, pose, key, uv, C, M, K, params, tmpl);
    const int P = chain == 0 ? 21 : 24;
    memset(JtJ, 0, (size_t)(n_free * n_free) * sizeof(double));
    memset(Jtr, 0, (size_t)n_free * sizeof(double));
    double c = 0;
    for (int64_t i = 0; i < N; ++i) {
        double r[2], J[48];
        int64_t cols[24];
        eval_obs(chain, params + 9 * (int64_t)cam[i], params + 9 * (int64_t)C + 6 * (int64_t)cam[i],
                 params + 15 * (int64_t)C + 6 * (int64_t)pose[i], obs_point(&p, i), uv[2 * i], uv[2 * i + 1], r, J);
        obs_columns(&p, i, cols);
        c += r[0] * r[0] + r[1] * r[1];
        for (int a = 0; a < P; ++a) {
            int64_t fa = free_map[cols[a]];
            if (fa < 0) continue;
            Jtr[fa] += J[a] * r[0] + J[P + a] * r[1];
            for (int b = 0; b < P; ++b) {
                int64_t fb = free_map[cols[b]];
                if (fb < 0) continue;
                JtJ[fa * n_free + fb] += J[a] * J[b] + J[P + a] * J[P + b];
            }
        }
    }
    *cost = c;
    return 0;
}

/*
 * Block normal equations for chain 0 (no reference counterpart; defined as the blocks of J.T @ J,
 * J.T @ r of the reference Jacobian with NO parameter fixed):
 *   U[c]  15x15 = sum [A|B_c]^T [A|B_c]        gc[c] 15 = sum [A|B_c]^T r
 *   V[m]  6x6   = sum B_m^T B_m                gp[m] 6  = sum B_m^T r
 *   W[s]  15x6  = sum [A|B_c]^T B_m  over the observations with seg[i] == s
 * seg[i] in [0, n_seg) names the (camera, pose) pair of observation i.
 * Threading: observations are walked per thread in contiguous ranges with private U/gc and
 * atomic adds for V / gp / W, so it also serves as the multi-threaded CPU baseline.
 */
ORACLE_API int oracle_normal_blocks(int64_t N, const int32_t *cam, const int32_t *pose, const int32_t *key,
                                    const double *uv, const int32_t *seg, int C, int M, int K, int64_t n_seg,
                                    const double *params, const double *tmpl, double *U, double *gc, double *V,
                                    double *gp, double *W, double *cost)
{
    problem_t p = mk(0, N, cam, pose, key, uv, C, M, K, params, tmpl);
    memset(U, 0, (size_t)C * 225 * sizeof(double));
    memset(gc, 0, (size_t)C * 15 * sizeof(double));
    memset(V, 0, (size_t)M * 36 * sizeof(double));
    memset(gp, 0, (size_t)M * 6 * sizeof(double));
    memset(W, 0, (size_t)n_seg * 90 * sizeof(double));
    double total = 0;
#pragma omp parallel reduction(+ : total)
    {
        double *Ul = (double *)calloc((size_t)C * 240, sizeof(double));
        double Wl[90], Vl[36], gl[6];
        int64_t cur_seg = -1;
        int32_t cur_pose = -1;
        memset(Wl, 0, sizeof Wl); memset(Vl, 0, sizeof Vl); memset(gl, 0, sizeof gl);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            double r[2], J[42];
            eval_obs(0, params + 9 * (int64_t)cam[i], params + 9 * (int64_t)C + 6 * (int64_t)cam[i],
                     params + 15 * (int64_t)C + 6 * (int64_t)pose[i], obs_point(&p, i), uv[2 * i], uv[2 * i + 1], r,
                     J);
            if (seg[i] != cur_seg) {
                if (cur_seg >= 0) {
                    for (int k = 0; k < 90; ++k) {
#pragma omp atomic
                        W[cur_seg * 90 + k] += Wl[k];
                    }
                    for (int k = 0; k < 36; ++k) {
#pragma omp atomic
                        V[(int64_t)cur_pose * 36 + k] += Vl[k];
                    }
                    for (int k = 0; k < 6; ++k) {
#pragma omp atomic
                        gp[(int64_t)cur_pose * 6 + k] += gl[k];
                    }
                }
                memset(Wl, 0, sizeof Wl); memset(Vl, 0, sizeof Vl); memset(gl, 0, sizeof gl);
                cur_seg = seg[i];
                cur_pose = pose[i];
            }
            total += r[0] * r[0] + r[1] * r[1];
            double *Uc = Ul + (int64_t)cam[i] * 240;
            for (int a = 0; a < 15; ++a) {
                double ju = J[a], jv = J[21 + a];
                Uc[225 + a] += ju * r[0] + jv * r[1];
                for (int b = 0; b < 15; ++b) Uc[15 * a + b] += ju * J[b] + jv * J[21 + b];
                for (int b = 0; b < 6; ++b) Wl[6 * a + b] += ju * J[15 + b] + jv * J[36 + b];
            }
            for (int a = 0; a < 6; ++a) {
                double ju = J[15 + a], jv = J[36 + a];
                gl[a] += ju * r[0] + jv * r[1];
                for (int b = 0; b < 6; ++b) Vl[6 * a + b] += ju * J[15 + b] + jv * J[36 + b];
            }
        }
        if (cur_seg >= 0) {
            for (int k = 0; k < 90; ++k) {
#pragma omp atomic
                W[cur_seg * 90 + k] += Wl[k];
            }
            for (int k = 0; k < 36; ++k) {
#pragma omp atomic
                V[(int64_t)cur_pose * 36 + k] += Vl[k];
            }
            for (int k = 0; k < 6; ++k) {
#pragma omp atomic
                gp[(int64_t)cur_pose * 6 + k] += gl[k];
            }
        }
#pragma omp critical
        {
            for (int c = 0; c < C; ++c) {
                for (int k = 0; k < 225; ++k) U[(int64_t)c * 225 + k] += Ul[(int64_t)c * 240 + k];
                for (int k = 0; k < 15; ++k) gc[(int64_t)c * 15 + k] += Ul[(int64_t)c * 240 + 225 + k];
            }
        }
        free(Ul);
    }
    *cost = total;
    return 0;
}

/* block-level known answers are checked through these thin exports */
ORACLE_API void oracle_block_projection(const double *q, const double *X, double *fun2, double *jac24)
{
    proj_fun(q, X, fun2);
    proj_jac(q, X, jac24);
}

ORACLE_API void oracle_block_rigid(const double *p, const double *X, double *fun3, double *jac27)
{
    double Jr[9], Jx[9];
    rigid_fun(p, X, fun3);
    rigid_jac(p, X, Jr, Jx);
    for (int a = 0; a < 3; ++a) {
        for (int i = 0; i < 3; ++i) jac27[9 * a + i] = Jr[3 * a + i];
        for (int i = 0; i < 3; ++i) jac27[9 * a + 3 + i] = (a == i) ? 1.0 : 0.0;
        for (int i = 0; i < 3; ++i) jac27[9 * a + 6 + i] = Jx[3 * a + i];
    }
}

ORACLE_API void oracle_block_rodrigues_jac(const double *r, double *out27) { rodrigues_jac(r, out27); }

/* ------------------------------------------------------------------------------------------------
 * Initialiser cost evaluation.
 * compiled_helpers.py:517-549 numpy_bundle_adjustment_costfn with :438-460 nb_distort_prealloc, called once per
 * candidate pose table by estimate_camera_relative_poses (template_handler.py:510-593):
 *   p = P_c [X; 1],  (u, v) = (p0 / p2, p1 / p2),  Brown-Conrady distortion applied in PIXEL space around the
 *   principal point (x = (u - cx) / fx, ...),  error = distorted - measured.
 * im_points: [M][K][3] target points already transformed by the candidate pose of every image;
 * proj: [C][3][4] = K_c [R_c | t_c];  ints: [C][3][3];  dists: [C][5] = (k1, k2, p1, p2, k3).
 * errors_out: [2N] interleaved (x, y) in dd row order.
 * ---------------------------------------------------------------------------------------------- */
ORACLE_API int oracle_costfn(int64_t N, const int32_t *cam, const int32_t *pose, const int32_t *key,
                             const double *uv, int C, int M, int K, const double *im_points, const double *proj,
                             const double *ints, const double *dists, double *errors_out)
{
    (void)C; (void)M;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        const double *P = proj + 12 * (int64_t)cam[i];
        const double *A = ints + 9 * (int64_t)cam[i];
        const double *k = dists + 5 * (int64_t)cam[i];
        const double *X = im_points + 3 * ((int64_t)pose[i] * K + key[i]);
        double p[3];
        for (int r = 0; r < 3; ++r) p[r] = P[4 * r] * X[0] + P[4 * r + 1] * X[1] + P[4 * r + 2] * X[2] + P[4 * r + 3];
        const double u = p[0] / p[2], v = p[1] / p[2];
        const double c0 = A[2], c1 = A[5], f0 = A[0], f1 = A[4];
        const double x = (u - c0) / f0, y = (v - c1) / f1;
        const double r2 = x * x + y * y;
        const double kup = 1 + k[0] * r2 + k[1] * (r2 * r2) + k[4] * (r2 * r2 * r2);
        double xD = x * kup, yD = y * kup;
        xD += 2 * k[2] * x * y + k[3] * (r2 + 2 * (x * x));
        yD += k[2] * (r2 + 2 * (y * y)) + 2 * k[3] * x * y;
        errors_out[2 * i] = xD * f0 + c0 - uv[2 * i];
        errors_out[2 * i + 1] = yD * f1 + c1 - uv[2 * i + 1];
    }
    return 0;
}
