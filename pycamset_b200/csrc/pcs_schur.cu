// pcs_schur.cu -- reduced camera system: S -= Z Z^T (lower triangle), the rank-6M update that eliminates the poses.
//
// Part of the LM step that stands in for scipy's TRF / LSMR solve (optimisation_handling.py:88-98): with
// Z = W L^-T (15 C x 6 M, dense when every camera sees every pose) the Schur complement of the pose blocks is
// S = U + lambda D - Z Z^T.  cuBLAS DSYRK reaches a third of the FP64 pipe on this shape (n = 480, k = 12000: few
// output tiles, long k); this kernel is shaped for it:
//   * 96 x 96 output tiles of the lower triangle; the (tile, k) iteration space is linearised and cut into one
//     contiguous range per SM ("stream-K"), so all SMs finish together whatever the tile count; a CTA flushes its
//     accumulators with FP64 reductions whenever its range leaves a tile (at most twice).
//   * Operands stream global -> shared memory with 16-byte asynchronous copies through a 3-stage ring of 16-column
//     slabs; Z is column-major, so a slab row (96 consecutive matrix rows of one column) is contiguous.
//   * 8 warps, warp tile 48 x 24 = 6 x 3 DMMA m8n8k4 accumulators; A and B fragments of a k-step are plain 8-byte
//     loads from the slab (row stride 104 doubles: the minimum two wavefronts per load).
#include <algorithm>

#include "pcs_internal.cuh"

namespace pcs {

namespace {

constexpr int ST = 96;            // tile edge
constexpr int SK = 16;            // slab depth (columns of Z per stage)
constexpr int SLD = ST + 8;       // slab row stride in doubles
constexpr int S_STAGES = 3;
constexpr int S_THREADS = 256;
constexpr int SLAB_DOUBLES = SK * SLD;              // one operand slab
constexpr int STAGE_DOUBLES = 2 * SLAB_DOUBLES;     // A rows + B rows

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async_zfill(void* smem, const void* gmem, int bytes, int src_bytes)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    if (bytes == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "r"(src_bytes) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// lower-triangle tile t -> (ti, tj), ti >= tj, row-major enumeration
__device__ __forceinline__ void tile_of(int t, int& ti, int& tj)
{
    ti = 0;
    while (t > ti) { t -= ti + 1; ++ti; }
    tj = t;
}

// Per-thread constants of the slab copies: a thread moves the same (column, row) positions of every slab, so the
// shared-memory offsets and the row / column offsets inside a slab are computed once.
template <bool ALIGNED>
struct FetchPlan {
    static constexpr int PER_SLAB = ALIGNED ? (SK * (ST / 2)) / S_THREADS : (SK * ST) / S_THREADS;   // copies per thread and slab
    int kk[PER_SLAB], row[PER_SLAB];   // column inside the slab, row inside the tile
    __device__ __forceinline__ void init()
    {
#pragma unroll
        for (int q = 0; q < PER_SLAB; ++q) {
            const int r = threadIdx.x + q * S_THREADS;
            if (ALIGNED) { kk[q] = r / (ST / 2); row[q] = 2 * (r % (ST / 2)); }
            else { kk[q] = r / ST; row[q] = r % ST; }
        }
    }
};

// fetch one stage: columns [k0, k0 + SK) of the rows of tiles ti (slab A) and tj (slab B)
template <bool ALIGNED>
__device__ __forceinline__ void fetch_stage(const FetchPlan<ALIGNED>& fp, double* __restrict__ stage, const double* __restrict__ Z,
                                            int64_t nc, int64_t np, int ti, int tj, int64_t k0)
{
#pragma unroll
    for (int slab = 0; slab < 2; ++slab) {
        const int64_t row0 = (int64_t)ST * (slab ? tj : ti);
#pragma unroll
        for (int q = 0; q < FetchPlan<ALIGNED>::PER_SLAB; ++q) {
            const int64_t row = row0 + fp.row[q], k = k0 + fp.kk[q];
            int src = 0;
            if (ALIGNED) { if (k < np) src = row + 1 < nc ? 16 : (row < nc ? 8 : 0); }
            else src = (k < np && row < nc) ? 8 : 0;
            const double* g = src ? Z + k * nc + row : Z;
            cp_async_zfill(stage + slab * SLAB_DOUBLES + fp.kk[q] * SLD + fp.row[q], g, ALIGNED ? 16 : 8, src);
        }
    }
}

template <bool ALIGNED>
__global__ void __launch_bounds__(S_THREADS, 2)
k_schur_syrk(int64_t nc, int64_t np, const double* __restrict__ Z, double* __restrict__ S, int n_lower, int64_t n_slabs)
{
    extern __shared__ __align__(16) double sy_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int wr = warp >> 2, wc = warp & 3;   // 2 x 4 warps: 48 rows x 24 columns each

    const int64_t units = (int64_t)n_lower * n_slabs;
    const int64_t u0 = units * blockIdx.x / gridDim.x, u1 = units * (blockIdx.x + 1) / gridDim.x;
    if (u0 >= u1) return;
    FetchPlan<ALIGNED> fp;
    fp.init();

    double acc[6][3][2];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto flush = [&](int ti, int tj) {
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int64_t row = (int64_t)ST * ti + wr * 48 + 8 * i + g;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int64_t col = (int64_t)ST * tj + wc * 24 + 8 * j + 2 * t4 + h;
                    if (row < nc && col <= row) atomicAdd(S + col * nc + row, -acc[i][j][h]);
                    acc[i][j][h] = 0.0;
                }
            }
    };

    // (tile, slab) of the unit being computed and of the unit being fetched, advanced incrementally
    int c_tile = (int)(u0 / n_slabs), cti, ctj;
    tile_of(c_tile, cti, ctj);
    int f_tile = c_tile, fti = cti, ftj = ctj;
    int64_t f_slab = u0 % n_slabs, f_u = u0;
    auto fetch_next = [&](int stage) {
        if (f_u < u1) {
            fetch_stage<ALIGNED>(fp, sy_smem + stage * STAGE_DOUBLES, Z, nc, np, fti, ftj, f_slab * SK);
            ++f_u;
            if (++f_slab == n_slabs) { f_slab = 0; ++f_tile; tile_of(f_tile, fti, ftj); }
        }
        cp_async_commit();
    };
    for (int s = 0; s < S_STAGES - 1; ++s) fetch_next(s);
    int c_stage = 0, f_stage = S_STAGES - 1;
    int64_t c_slab = u0 % n_slabs;
    for (int64_t u = u0; u < u1; ++u) {
        cp_async_wait<S_STAGES - 2>();
        __syncthreads();   // stage u has landed for everybody; stage u - 1 has been consumed by everybody
        fetch_next(f_stage);
        f_stage = f_stage + 1 == S_STAGES ? 0 : f_stage + 1;
        const double* A = sy_smem + c_stage * STAGE_DOUBLES + wr * 48 + g;
        const double* B = sy_smem + c_stage * STAGE_DOUBLES + SLAB_DOUBLES + wc * 24 + g;
#pragma unroll
        for (int s = 0; s < SK / 4; ++s) {
            double a[6], b[3];
#pragma unroll
            for (int i = 0; i < 6; ++i) a[i] = A[(4 * s + t4) * SLD + 8 * i];
#pragma unroll
            for (int j = 0; j < 3; ++j) b[j] = B[(4 * s + t4) * SLD + 8 * j];
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        c_stage = c_stage + 1 == S_STAGES ? 0 : c_stage + 1;
        if (++c_slab == n_slabs) {   // the range leaves this tile
            flush(cti, ctj);
            c_slab = 0;
            ++c_tile;
            tile_of(c_tile, cti, ctj);
        }
    }
    if (c_slab != 0) flush(cti, ctj);
}

}  // namespace

// S (n x n, column-major, lower triangle) -= Z Z^T with Z given column-major as [k][n]
int launch_schur_syrk(cudaStream_t st, int sm_count, int64_t n, int64_t k, const double* Z, double* S)
{
    if (n <= 0 || k <= 0) return PCS_OK;
    const int n_t = (int)((n + ST - 1) / ST);
    const int n_lower = n_t * (n_t + 1) / 2;
    const int64_t n_slabs = (k + SK - 1) / SK;
    const size_t smem = (size_t)S_STAGES * STAGE_DOUBLES * sizeof(double);
    const bool aligned = (n % 2 == 0) && ((reinterpret_cast<uintptr_t>(Z) & 15) == 0);
    auto kern = aligned ? k_schur_syrk<true> : k_schur_syrk<false>;
    PCS_CUDA(ensure_dynamic_smem(kern, smem));
    const int64_t units = (int64_t)n_lower * n_slabs;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(2 * (int64_t)sm_count, units));   // two CTAs per SM
    kern<<<grid, S_THREADS, smem, st>>>(n, k, Z, S, n_lower, n_slabs);
    PCS_CUDA(cudaGetLastError());
    return PCS_OK;
}

}  // namespace pcs

extern "C" int pcs_syrk_sub(int device, int64_t n, int64_t k, const double* Z, double* S)
{
    using namespace pcs;
    PCS_REQUIRE(n > 0 && k > 0 && Z && S, "NULL argument or empty shape");
    PCS_CUDA(cudaSetDevice(device));
    double *dZ = nullptr, *dS = nullptr;
    int sms = 0;
    PCS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    cudaError_t e = cudaMalloc((void**)&dZ, (size_t)(n * k) * 8);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dS, (size_t)(n * n) * 8);
    if (e == cudaSuccess) e = cudaMemcpy(dZ, Z, (size_t)(n * k) * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dS, S, (size_t)(n * n) * 8, cudaMemcpyHostToDevice);
    int rc = PCS_OK;
    if (e == cudaSuccess) rc = launch_schur_syrk(nullptr, sms, n, k, dZ, dS);
    if (e == cudaSuccess && rc == PCS_OK) e = cudaDeviceSynchronize();
    if (e == cudaSuccess && rc == PCS_OK) e = cudaMemcpy(S, dS, (size_t)(n * n) * 8, cudaMemcpyDeviceToHost);
    cudaFree(dZ);
    cudaFree(dS);
    if (e != cudaSuccess) { set_error(std::string("pcs_syrk_sub: ") + cudaGetErrorString(e)); return PCS_ERR_CUDA; }
    return rc;
}
