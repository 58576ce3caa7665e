// chol_trace.cu -- per-phase timeline of k_chol_solve (CTA 0): load+update / factor / tiles / barrier, in ns.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pycamset_b200/csrc -o tools/chol_trace tools/chol_trace.cu
#include "pcs_chol.cu"
#include <vector>
#include <random>
namespace pcs { void set_error(const std::string& m) { fprintf(stderr, "%s\n", m.c_str()); } }
int main(int argc, char** argv)
{
    using namespace pcs;
    const int64_t n = argc > 1 ? atoll(argv[1]) : 480;
    std::vector<double> A(n * n), b(n, 1.0);
    std::mt19937_64 g(1);
    std::normal_distribution<double> nd;
    for (int64_t i = 0; i < n; ++i) for (int64_t j = 0; j <= i; ++j) { double v = 0.01 * nd(g); A[j * n + i] = v; A[i * n + j] = v; }
    for (int64_t i = 0; i < n; ++i) A[i * n + i] = n * 0.02 + 1.0;
    double *dA, *db, *Ldiag; unsigned long long* bar; unsigned long long bar_base = 0; int* info; long long* tr; int grid;
    const int nb = (int)((n + 31) / 32);
    chol_prepare(0, n, &Ldiag, &bar, &grid);
    cudaMalloc(&dA, n * n * 8); cudaMalloc(&db, n * 8); cudaMalloc(&info, 4); cudaMemset(info, 0, 4); cudaMalloc(&tr, (nb + 1) * 8 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    // warm the clocks up: ~0.5 s of back-to-back solves before the traced one
    for (int rep = 0; rep < (n > 700 ? 200 : 2000); ++rep) launch_chol_solve(nullptr, grid, n, dA, n, db, Ldiag, bar, &bar_base, info, nullptr);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 20; ++rep) {
        cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), n * 8, cudaMemcpyHostToDevice);
        cudaEventRecord(e0);
        launch_chol_solve(nullptr, grid, n, dA, n, db, Ldiag, bar, &bar_base, info, nullptr);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    printf("best of 20 (incl. 2 memsets): %.1f us\n", best * 1e3);
    cudaMemset(info, 0, 4);   // the warm-up solves re-factor an already factored matrix and raise it
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), n * 8, cudaMemcpyHostToDevice);
        cudaEventRecord(e0);
        launch_chol_solve(nullptr, grid, n, dA, n, db, Ldiag, bar, &bar_base, info, tr);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    std::vector<long long> h((nb + 1) * 8);
    cudaMemcpy(h.data(), tr, (nb + 1) * 8 * 8, cudaMemcpyDeviceToHost);
    int hi; cudaMemcpy(&hi, info, 4, cudaMemcpyDeviceToHost);
    printf("n=%lld grid=%d info=%d total %.1f us\n", (long long)n, grid, hi, ms * 1e3);
    for (int k = 0; k < nb && k < 20; ++k)
        printf("phase %2d: loads %6lld  diag update %6lld  (factor alone %6lld cycles) factor %6lld  trsm %6lld  store+rest %6lld  barrier %6lld ns\n", k, h[k*8+5]-h[k*8],
               h[k*8+1]-h[k*8+5], h[k*8+7], h[k*8+2]-h[k*8+1], h[k*8+6]-h[k*8+2], h[k*8+3]-h[k*8+6], h[k*8+4]-h[k*8+3]);
    printf("back substitution alone: %.1f us\n", (h[nb*8+1] - h[nb*8]) * 1e-3);
    printf("factorisation %.1f us, back substitution + rest: %.1f us\n", (h[(nb-1)*8+4] - h[0]) * 1e-3, ms * 1e3 - (h[(nb-1)*8+4] - h[0]) * 1e-3);
    return 0;
}
