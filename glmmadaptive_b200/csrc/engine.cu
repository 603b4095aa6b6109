// C ABI of the AGQ engine (include/glmma.h): handle lifetime, host-side theta
// algebra, kernel sequencing.  No CPU fallback: every compute entry point needs
// a CUDA device and fails with GLMMA_ERR_CUDA otherwise.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <queue>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <nvtx3/nvToolsExt.h>       // header-only NVTX 3: ranges cost nothing unless a profiler is attached
#include "../../include/glmma.h"
#include "kernels.cuh"
#include "engine.h"
#include "fastmath_tables.h"

#define SCORE_THREADS 256
#define FINAL_THREADS 256

static thread_local std::string g_create_error;
#define GLMMA_NTILES 5
#define GLMMA_TILE32 3
#define GLMMA_TILEPAIR 4
// observation tiles of the register variant (32: two-pass form; 17: the code of the pair tile, two random-intercept
// clusters of at most 8 observations each in one 16-row tile)
static const int kTile[GLMMA_NTILES] = {8, 12, 16, 32, 17};

// One host thread per device of a multi-device handle: the per-evaluation enqueue work (six kernel
// launches and a copy per device) runs in parallel instead of serialising on the caller's thread.
// The workers only ever touch CUDA and their own sub-handle -- never the caller's runtime (R).
struct PartWorker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, stop = false;
    int rc = 0;
    void loop() {
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cv.wait(lk, [&] { return has_job || stop; });
            if (stop) return;
            lk.unlock();
            const int r = job();
            lk.lock();
            rc = r;
            has_job = false;
            cv.notify_all();
        }
    }
    void submit(std::function<int()> f) {
        std::lock_guard<std::mutex> lk(m);
        job = std::move(f);
        has_job = true;
        cv.notify_all();
    }
    int wait() {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return !has_job; });
        return rc;
    }
    void shutdown() {
        { std::lock_guard<std::mutex> lk(m); stop = true; cv.notify_all(); }
        if (th.joinable()) th.join();
    }
};

struct glmma_handle {
    DevData d;                 // device pointers + sizes
    int family, link, device, nagq, q, K, y_cols;
    double family_df;          // students.t(df)
    const FamOps* ops;
    cudaStream_t stream;
    GridPar* d_grid;
    double* d_fmtab;
    double* d_ytab;
    // launch shape of the eval kernel, [0] = loglik only, [1] = with score
    int jt;
    int eval_grid[2];
    size_t eval_smem[2];
    int eval_threads[2];       // block shape of the generic kernel
    int asum_doubles;          // per-warp node sums of the generic kernel's tile-outer path (0: no cluster needs it)
    int nj;                    // register-variant tiles in use: bit v = tile kTile[v] (8, 12, 16); 0 = generic kernel only
    int fast_grid[GLMMA_NTILES][2];       // [tile][loglik only / with score]
    size_t fast_smem[GLMMA_NTILES][2];
    int fast_threads[GLMMA_NTILES][2];    // block shape of the register variant (chosen for resident warps per SM)
    int4* d_order[GLMMA_NTILES];          // cluster lists of the tiles: {cluster, first row, n_i, 0}
    int64_t n_order[GLMMA_NTILES];
    // fused mode (kernels.cuh, EvalArgs): the register variant does the prep and score-reduce work itself
    // and eval_kernel's last block finalises -- an evaluation is [ytab] + fast launches + eval_kernel
    bool fused;
    int xw;                    // staged design columns per observation (p + p_zi + offsets)
    int32_t* d_defer_list;     // clusters no tile of the register variant takes (they still need prep)
    int64_t n_defer_list;
    unsigned* d_counter;       // blocks-done counter of the in-kernel finalisation
    double* d_roww;            // per-row cluster weight, 0 on rows of clusters the register variant does not take
                               // (null: every row counts once) -- const_sums_kernel
    double* d_cpart;           // const_sums_kernel block partials
    int n_cpart;
    // peer-memory all-reduce over the ranks of a sharded job (glmma_comm_export / glmma_comm_connect)
    double* d_xchg;            // 2 (parity) x nranks x GLMMA_XSLOT doubles, exported to the other ranks
    XchgPar xchg;              // the ranks' buffers as mapped here; nranks <= 1: not connected
    std::vector<void*> ipc_opened;
    // zero-copy result delivery (kernels.cuh: deliver_host): pinned, device-mapped host memory
    double* h_out;             // result vector + flag (last slot), host address
    double* h_out_dev;         // the same memory as the device sees it
    unsigned long long out_seq;
    std::vector<double> xty_w; // sum_i w_i X_i'y1_i over the register variant's clusters (linear-eta hoist)
    double yoff_w;             // sum_i w_i y1_i'offset_i over the same clusters
    bool ytab_valid;           // d_ytab holds the table of ytab_phis
    double ytab_phis;
    int score_grid;
    double* d_partials;        // eval block partials
    double* d_score_partials;  // score_grid * (p + p_zi)
    double* d_result;          // result vector
    unsigned long long* d_stats;
    double* d_pby;             // n x K posterior node weights of the last E-step (lazily allocated)
    bool have_pby;
    bool have_modes;
    int64_t launches;
    std::vector<void*> allocs;
    std::string error;
    int sticky;                // sticky CUDA error status
    cudaEvent_t ev[4];
    // single-process multi-GPU handle (glmma_create_multi): one sub-handle per device over a
    // contiguous range of clusters; this handle itself then owns no device memory
    std::vector<glmma_handle*> parts;
    std::vector<std::unique_ptr<PartWorker>> workers;
    std::vector<int64_t> part_c0, part_r0;     // first cluster / first row of each part (+ total at the end)
    double* h_pinned = nullptr;                // sub-handle: pinned host copy of the result vector
    unsigned long long* h_stats = nullptr;     // sub-handle: pinned host copy of the mode statistics
    int multi_p = 0, multi_pzi = 0, multi_nphis = 0;
};

// ------------------------------------------------------------------------------
// small host math
// ------------------------------------------------------------------------------
// gauher(n): Gauss-Hermite nodes (descending) and weights by Newton iteration on the
// orthonormal Hermite recurrence, same start values and tolerance as the
// reference's gauher() (R/Functions.R:306-341) so that the tables agree to rounding.
static void gauher_table(int n, double* x, double* w) {
    const int m = (n + 1) / 2;
    double z = 0.0, pp = 0.0;
    for (int i = 1; i <= m; ++i) {
        if (i == 1) z = std::sqrt(2.0 * n + 1.0) - 1.85575 * std::pow(2.0 * n + 1.0, -0.16667);
        else if (i == 2) z = z - 1.14 * std::pow((double)n, 0.426) / z;
        else if (i == 3) z = 1.86 * z - 0.86 * x[0];
        else if (i == 4) z = 1.91 * z - 0.91 * x[1];
        else z = 2.0 * z - x[i - 3];
        for (int its = 0; its < 10; ++its) {
            double p1 = 0.751125544464943, p2 = 0.0;
            for (int j = 1; j <= n; ++j) {
                const double p3 = p2;
                p2 = p1;
                p1 = z * std::sqrt(2.0 / j) * p2 - std::sqrt((j - 1.0) / j) * p3;
            }
            pp = std::sqrt(2.0 * n) * p2;
            const double z1 = z;
            z = z1 - p1 / pp;
            if (std::fabs(z - z1) <= 3e-14) break;
        }
        x[i - 1] = z;
        x[n - i] = -z;
        w[i - 1] = 2.0 / (pp * pp);
        w[n - i] = w[i - 1];
    }
}

// inverse and log-determinant of a symmetric positive definite q x q matrix
static bool spd_inverse(const double* A, int q, double* Ainv, double* logdet) {
    double L[GLMMA_QMAX * GLMMA_QMAX] = {0};
    double ld = 0.0;
    for (int c = 0; c < q; ++c) {
        double s = A[c * q + c];
        for (int t = 0; t < c; ++t) s -= L[c * q + t] * L[c * q + t];
        if (!(s > 0.0) || !std::isfinite(s)) return false;
        L[c * q + c] = std::sqrt(s);
        ld += std::log(s);
        for (int r = c + 1; r < q; ++r) {
            double v = 0.5 * (A[r * q + c] + A[c * q + r]);
            for (int t = 0; t < c; ++t) v -= L[r * q + t] * L[c * q + t];
            L[r * q + c] = v / L[c * q + c];
        }
    }
    double Li[GLMMA_QMAX * GLMMA_QMAX] = {0};
    for (int c = 0; c < q; ++c)
        for (int r = c; r < q; ++r) {
            double s = (r == c) ? 1.0 : 0.0;
            for (int t = c; t < r; ++t) s -= L[r * q + t] * Li[t * q + c];
            Li[r * q + c] = s / L[r * q + r];
        }
    for (int r = 0; r < q; ++r)
        for (int c = 0; c < q; ++c) {
            double s = 0.0;
            for (int t = 0; t < q; ++t) s += Li[t * q + r] * Li[t * q + c];
            Ainv[r * q + c] = s;
        }
    if (logdet) *logdet = ld;
    return true;
}

static bool fill_fampar(int family, int n_phis, const double* phis, FamPar* fp, std::string& err, double df = 0.0) {
    std::memset(fp, 0, sizeof(*fp));
    const bool needs = family == GLMMA_STUDENTS_T || family == GLMMA_BETA_BINOMIAL || family == GLMMA_NEGBIN || family == GLMMA_ZI_NEGBIN || family == GLMMA_HURDLE_NEGBIN ||
                       family == GLMMA_HURDLE_LOGNORMAL || family == GLMMA_BETA || family == GLMMA_HURDLE_BETA ||
                       family == GLMMA_GAMMA || family == GLMMA_CENSORED_NORMAL;
    if (!needs) return true;
    if (n_phis != 1 || !phis) { err = "this family needs exactly one phis value"; return false; }
    const double ph = std::exp(phis[0]);
    switch (family) {
        case GLMMA_NEGBIN: case GLMMA_ZI_NEGBIN: case GLMMA_HURDLE_NEGBIN: case GLMMA_GAMMA:
            fp->a[0] = ph; fp->a[1] = phis[0]; fp->a[2] = std::lgamma(ph); fp->a[3] = glmma_digamma(ph); break;
        case GLMMA_HURDLE_LOGNORMAL:
            fp->a[0] = ph; fp->a[1] = phis[0]; fp->a[2] = 1.0 / (ph * ph); break;
        case GLMMA_BETA: case GLMMA_HURDLE_BETA: case GLMMA_BETA_BINOMIAL:
            fp->a[0] = ph; fp->a[1] = std::lgamma(ph); fp->a[2] = glmma_digamma(ph); break;
        case GLMMA_STUDENTS_T:
            fp->a[0] = ph; fp->a[2] = 1.0 / ph; fp->a[3] = df + 1.0; fp->a[4] = 1.0 / df; fp->a[5] = 0.5 * (df + 1.0);
            fp->a[1] = std::lgamma(0.5 * (df + 1.0)) - std::lgamma(0.5 * df) - 0.5 * std::log(df * 3.14159265358979323846) - phis[0];
            break;
        case GLMMA_CENSORED_NORMAL:
            fp->a[0] = ph; fp->a[1] = phis[0]; fp->a[2] = 1.0 / ph; break;
    }
    return true;
}

static const FamOps* lookup_ops(int family, int link, int qy, int qz) {
    switch (family) {
        case GLMMA_NEGBIN: return glmma_ops_negbin(qy, qz);
#ifndef GLMMA_DEV_ONLY      // development builds (make DEV=1) link the negative binomial kernels only
        case GLMMA_BINOMIAL: return glmma_ops_binomial(link, qy, qz);
        case GLMMA_POISSON: return (link == GLMMA_LINK_DEFAULT || link == GLMMA_LINK_LOG) ? glmma_ops_poisson(qy, qz) : nullptr;
        case GLMMA_ZI_POISSON: return glmma_ops_zi_poisson(qy, qz);
        case GLMMA_ZI_NEGBIN: return glmma_ops_zi_negbin(qy, qz);
        case GLMMA_ZI_BINOMIAL: return glmma_ops_zi_binomial(qy, qz);
        case GLMMA_HURDLE_POISSON: return glmma_ops_hurdle_poisson(qy, qz);
        case GLMMA_HURDLE_NEGBIN: return glmma_ops_hurdle_negbin(qy, qz);
        case GLMMA_HURDLE_LOGNORMAL: return glmma_ops_hurdle_lognormal(qy, qz);
        case GLMMA_BETA: return glmma_ops_beta(qy, qz);
        case GLMMA_HURDLE_BETA: return glmma_ops_hurdle_beta(qy, qz);
        case GLMMA_GAMMA: return glmma_ops_gamma(qy, qz);
        case GLMMA_CENSORED_NORMAL: return glmma_ops_censored_normal(qy, qz);
        case GLMMA_STUDENTS_T: return (link == GLMMA_LINK_DEFAULT || link == GLMMA_LINK_IDENTITY) ? glmma_ops_students_t(qy, qz) : nullptr;
        case GLMMA_BETA_BINOMIAL: return glmma_ops_beta_binomial(link, qy, qz);
#endif
    }
    return nullptr;
}

// ------------------------------------------------------------------------------
// small kernels owned by this file
// ------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* sh) {
    // fixed-order block reduction (deterministic)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    return t;   // valid on thread 0
}

// g_beta[l] = sum_j X[j,l] zbar[j],  g_gamma[l] = sum_j X_zi[j,l] zbar_zi[j]   (block partials)
__global__ void __launch_bounds__(SCORE_THREADS) score_reduce_kernel(DevData d, double* partials) {
    __shared__ double sh[SCORE_THREADS / 32];
    const int64_t chunk = (d.N + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * chunk;
    const int64_t hi = min(lo + chunk, d.N);
    const int ncol = d.p + d.p_zi;
    for (int l = 0; l < ncol; ++l) {
        const double* col = (l < d.p) ? d.X + (int64_t)l * d.N : d.X_zi + (int64_t)(l - d.p) * d.N;
        const double* z = (l < d.p) ? d.zbar : d.zbar_zi;
        double s = 0.0;
        for (int64_t r = lo + threadIdx.x; r < hi; r += SCORE_THREADS) s = fma(col[r], z[r], s);
        const double t = block_sum(s, sh);
        if (threadIdx.x == 0) partials[(int64_t)blockIdx.x * ncol + l] = t;
    }
}

// result vector from the block partials (kernels.cuh: finalize_block), one block -- the separate launch
// of the non-fused path; the fused path finalises inside eval_kernel
__global__ void __launch_bounds__(FINAL_THREADS) finalize_kernel(const double* eval_partials, int n_eval_blocks,
                                                                  const double* score_partials, int n_score_blocks,
                                                                  int p, int n_phis, int p_zi, int q, int want_score,
                                                                  double* result, XchgPar xchg, HostOut hout) {
    __shared__ double sh[GLMMA_NACC + GLMMA_PMAX + GLMMA_PZIMAX + 2];
    finalize_block(eval_partials, n_eval_blocks, score_partials, n_score_blocks, false, p, n_phis, p_zi, q, want_score,
                   result, sh, nullptr, 0, 0.0, true, &xchg, &hout);
}

// ------------------------------------------------------------------------------
// starting values: one IRLS step of stats::glm.fit (as mixed_model() calls it,
// R/mixed_model.R:162-194) as one pass over the rows: the p x p normal equations
//   A = X' W X,  r = X' W z,  deviance
// with W = mu.eta^2 / V(mu) * prior weight and z = (eta - offset) + (y - mu) / mu.eta.
// first = 1: eta from family$initialize's mustart instead of X beta.
// fam: 0 binomial(logit), 1 poisson(log), 2 Gamma(inverse), 3 gaussian(identity).
// resp: 0 the model response (y1; binomial: y1 / y2 with prior weight y2), 1 the zero indicator.
// ------------------------------------------------------------------------------
#define GLM_PMAX 8
#define GLM_NACC (GLM_PMAX * (GLM_PMAX + 1) / 2 + GLM_PMAX + 1)
#define GLM_THREADS 128
struct GlmArgs {
    const double* X; const double* y1; const double* y2; const double* offset;
    int64_t N; int p; int fam; int resp; int first; int y1_is_log;
    double beta[GLM_PMAX];
    double* partials;     // gridDim.x * GLM_NACC
};

__global__ void __launch_bounds__(GLM_THREADS) glm_irls_kernel(GlmArgs A) {
    __shared__ double sh[GLM_THREADS / 32];
    double acc[GLM_NACC];
#pragma unroll
    for (int t = 0; t < GLM_NACC; ++t) acc[t] = 0.0;
    const int p = A.p;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < A.N; row += (int64_t)gridDim.x * blockDim.x) {
        double x[GLM_PMAX];
#pragma unroll
        for (int l = 0; l < GLM_PMAX; ++l) x[l] = l < p ? A.X[row + (int64_t)l * A.N] : 0.0;
        double y = A.y1[row], pw = 1.0;
        if (A.resp == 1) {
            y = ((A.y1_is_log ? y == -INFINITY : y == 0.0)) ? 1.0 : 0.0;
        } else if (A.fam == 0 && A.y2) {
            pw = A.y2[row];
            y = pw > 0.0 ? y / pw : 0.0;
        }
        const double off = A.offset ? A.offset[row] : 0.0;
        double eta, mu;
        if (A.first) {
            mu = A.fam == 0 ? (pw * y + 0.5) / (pw + 1.0) : (A.fam == 1 ? y + 0.1 : y);
            eta = A.fam == 0 ? log(mu / (1.0 - mu)) : (A.fam == 1 ? log(mu) : (A.fam == 2 ? 1.0 / mu : mu));
        } else {
            eta = off;
#pragma unroll
            for (int l = 0; l < GLM_PMAX; ++l) eta = fma(x[l], A.beta[l], eta);
        }
        double g, var;
        if (A.fam == 0) {
            const double ec = fmin(fmax(eta, -30.0), 30.0);
            mu = 1.0 / (1.0 + exp(-ec));
            const double e = exp(-fabs(eta));
            g = fmax(e / ((1.0 + e) * (1.0 + e)), GLMMA_DBL_EPS);
            var = mu * (1.0 - mu);
        } else if (A.fam == 1) {
            mu = fmax(exp(eta), GLMMA_DBL_EPS);
            g = mu; var = mu;
        } else if (A.fam == 2) {
            mu = 1.0 / eta; g = -1.0 / (eta * eta); var = mu * mu;
        } else {
            mu = eta; g = 1.0; var = 1.0;
        }
        const double ww = pw * g * g / var;
        const double zz = (eta - off) + (y - mu) / g;
        int t = 0;
#pragma unroll
        for (int l = 0; l < GLM_PMAX; ++l) {
            const double wx = ww * x[l];
#pragma unroll
            for (int m = l; m < GLM_PMAX; ++m) { acc[t] = fma(wx, x[m], acc[t]); ++t; }
            acc[GLM_PMAX * (GLM_PMAX + 1) / 2 + l] = fma(wx, zz, acc[GLM_PMAX * (GLM_PMAX + 1) / 2 + l]);
        }
        // deviance of the CURRENT mu (glm.fit's convergence test)
        double dv;
        if (A.fam == 0) {
            dv = 2.0 * pw * ((y > 0.0 ? y * log(y / mu) : 0.0) + (y < 1.0 ? (1.0 - y) * log((1.0 - y) / (1.0 - mu)) : 0.0));
        } else if (A.fam == 1) {
            dv = 2.0 * ((y > 0.0 ? y * log(y / mu) : 0.0) - (y - mu));
        } else if (A.fam == 2) {
            dv = -2.0 * (log(y == 0.0 ? 1.0 : y / mu) - (y - mu) / mu);
        } else {
            dv = (y - mu) * (y - mu);
        }
        acc[GLM_NACC - 1] += dv;
    }
    for (int t = 0; t < GLM_NACC; ++t) {
        const double v = block_sum(acc[t], sh);
        if (threadIdx.x == 0) A.partials[(int64_t)blockIdx.x * GLM_NACC + t] = v;
    }
}

__global__ void fp64_peak_kernel(double* out, int iters) {
    // 8 independent DFMA chains per thread
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------
static int fail(glmma_handle* h, int code, const std::string& msg) {
    if (h) h->error = msg; else g_create_error = msg;
    return code;
}
static int cuda_fail(glmma_handle* h, cudaError_t e, const char* what) {
    std::string msg = std::string(what) + ": " + cudaGetErrorString(e);
    if (h) { h->error = msg; h->sticky = GLMMA_ERR_CUDA; } else g_create_error = msg;
    return GLMMA_ERR_CUDA;
}
#define CU(h, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail((h), e_, #call); } while (0)

template <class T>
static int dev_alloc(glmma_handle* h, T** out, size_t count) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaMalloc");
    h->allocs.push_back(p);
    *out = (T*)p;
    return GLMMA_OK;
}
template <class T>
static int dev_upload(glmma_handle* h, const T** out, const T* src, size_t count) {
    T* p = nullptr;
    int rc = dev_alloc(h, &p, count);
    if (rc) return rc;
    CU(h, cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    *out = p;
    return GLMMA_OK;
}

// NVTX range around an ABI call (mode search / evaluation / EM passes show up as named spans in Nsight)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// Restores the caller's current CUDA device when an ABI call returns: the host process (R, or Python
// with torch) must not find itself on another device after calling into the engine.
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); } }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

extern "C" {

int glmma_version(void) { return GLMMA_VERSION; }

int glmma_family_id(const char* name) {
    if (!name) return -1;
    static const struct { const char* n; int id; } tab[] = {
        {"binomial", GLMMA_BINOMIAL}, {"poisson", GLMMA_POISSON}, {"negative binomial", GLMMA_NEGBIN},
        {"zero-inflated poisson", GLMMA_ZI_POISSON}, {"zero-inflated negative binomial", GLMMA_ZI_NEGBIN},
        {"zero-inflated binomial", GLMMA_ZI_BINOMIAL}, {"hurdle poisson", GLMMA_HURDLE_POISSON},
        {"hurdle negative binomial", GLMMA_HURDLE_NEGBIN}, {"hurdle log-normal", GLMMA_HURDLE_LOGNORMAL},
        {"beta", GLMMA_BETA}, {"hurdle beta", GLMMA_HURDLE_BETA}, {"Gamma", GLMMA_GAMMA},
        {"censored normal", GLMMA_CENSORED_NORMAL}, {"Student's-t", GLMMA_STUDENTS_T},
        {"beta binomial", GLMMA_BETA_BINOMIAL}};
    for (auto& t : tab) if (std::strcmp(t.n, name) == 0) return t.id;
    return -1;
}

const char* glmma_last_error(const glmma_handle* h) { return h ? h->error.c_str() : g_create_error.c_str(); }

void glmma_destroy(glmma_handle* h) {
    if (!h) return;
    DeviceGuard device_guard_;
    if (!h->parts.empty()) {
        for (auto& w : h->workers) w->shutdown();
        for (glmma_handle* p : h->parts) glmma_destroy(p);
        delete h;
        return;
    }
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    if (h->h_stats) cudaFreeHost(h->h_stats);
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (void* p : h->ipc_opened) cudaIpcCloseMemHandle(p);
    if (h->h_out) cudaFreeHost(h->h_out);
    if (h->d_xchg) cudaFree(h->d_xchg);
    for (void* p : h->allocs) cudaFree(p);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    delete h;
}

int glmma_create(const glmma_data* data, int device, glmma_handle** out) {
    DeviceGuard device_guard_;
    NvtxRange nvtx_range_("glmma_create");
    if (!data || !out) return fail(nullptr, GLMMA_ERR_ARG, "null argument");
    *out = nullptr;
    const int64_t N = data->n_obs;
    const int q = data->q_y + data->q_zi;
    if (N <= 0 || data->p <= 0 || data->q_y <= 0 || !data->y || !data->X || !data->Z)
        return fail(nullptr, GLMMA_ERR_ARG, "y, X, Z and positive sizes are required");
    if (N >= (int64_t)1 << 31) return fail(nullptr, GLMMA_ERR_ARG, "n_obs must be below 2^31");
    if (data->p > GLMMA_PMAX || data->p_zi > GLMMA_PZIMAX) return fail(nullptr, GLMMA_ERR_ARG, "too many fixed-effect columns");
    if (data->nAGQ < 1 || data->nAGQ > GLMMA_NAGQ_MAX) return fail(nullptr, GLMMA_ERR_ARG, "nAGQ out of range");
    if (data->family < 0 || data->family >= GLMMA_N_FAMILIES) return fail(nullptr, GLMMA_ERR_ARG, "unknown family");
    if (data->y_cols != 1 && data->y_cols != 2) return fail(nullptr, GLMMA_ERR_ARG, "y_cols must be 1 or 2");
    if (data->family == GLMMA_CENSORED_NORMAL && data->y_cols != 2)
        return fail(nullptr, GLMMA_ERR_ARG, "censored normal needs a two-column response");
    if (data->family == GLMMA_STUDENTS_T && !(data->family_df > 0.0))
        return fail(nullptr, GLMMA_ERR_ARG, "students.t needs family_df > 0 (the df of the family object)");
    if ((data->q_zi > 0 && !data->Z_zi) || (data->p_zi > 0 && !data->X_zi))
        return fail(nullptr, GLMMA_ERR_ARG, "X_zi / Z_zi missing");
    const FamOps* ops = lookup_ops(data->family, data->link, data->q_y, data->q_zi);
    if (!ops) return fail(nullptr, GLMMA_ERR_ARG, "family / link / random-effects dimensions not supported by the engine");
    if (ops->has_zi && data->p_zi <= 0) return fail(nullptr, GLMMA_ERR_ARG, "this family needs X_zi (zi_fixed)");
    if (!ops->has_zi && (data->p_zi > 0 || data->q_zi > 0)) return fail(nullptr, GLMMA_ERR_ARG, "family has no zero part");
    double Kd = 1.0;
    for (int i = 0; i < q; ++i) Kd *= data->nAGQ;
    if (Kd > 1 << 20) return fail(nullptr, GLMMA_ERR_ARG, "nAGQ^q too large");

    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0 || device < 0 || device >= ndev)
        return fail(nullptr, GLMMA_ERR_CUDA, std::string("no usable CUDA device (the engine has no CPU fallback): ") +
                                                 (ce != cudaSuccess ? cudaGetErrorString(ce) : "bad device index"));
    ce = cudaSetDevice(device);
    if (ce != cudaSuccess) return cuda_fail(nullptr, ce, "cudaSetDevice");

    // CSR cluster index
    std::vector<int32_t> row_ptr;
    if (data->row_ptr) {
        if (data->n_clusters <= 0) return fail(nullptr, GLMMA_ERR_ARG, "n_clusters must be positive");
        row_ptr.resize(data->n_clusters + 1);
        for (int64_t i = 0; i <= data->n_clusters; ++i) row_ptr[i] = (int32_t)data->row_ptr[i];
        if (data->row_ptr[0] != 0 || data->row_ptr[data->n_clusters] != N)
            return fail(nullptr, GLMMA_ERR_ARG, "row_ptr must start at 0 and end at n_obs");
        for (int64_t i = 0; i < data->n_clusters; ++i)
            if (row_ptr[i + 1] <= row_ptr[i]) return fail(nullptr, GLMMA_ERR_ARG, "row_ptr must be strictly increasing");
    } else if (data->id) {
        row_ptr.push_back(0);
        for (int64_t r = 1; r < N; ++r)
            if (data->id[r] != data->id[r - 1]) row_ptr.push_back((int32_t)r);
        row_ptr.push_back((int32_t)N);
        if (data->n_clusters > 0 && (int64_t)row_ptr.size() - 1 != data->n_clusters)
            return fail(nullptr, GLMMA_ERR_ARG, "id is not contiguous per cluster (n_clusters mismatch)");
    } else {
        return fail(nullptr, GLMMA_ERR_ARG, "id or row_ptr is required");
    }
    const int64_t n = (int64_t)row_ptr.size() - 1;

    glmma_handle* h = new glmma_handle();
    h->family = data->family; h->link = data->link; h->device = device; h->nagq = data->nAGQ;
    h->q = q; h->K = (int)Kd; h->y_cols = data->y_cols; h->ops = ops; h->stream = nullptr;
    h->have_modes = false; h->launches = 0; h->sticky = 0; h->d_ytab = nullptr;
    h->family_df = data->family_df;
    h->d_pby = nullptr; h->have_pby = false;
    h->fused = false; h->xw = 0; h->d_defer_list = nullptr; h->n_defer_list = 0; h->d_counter = nullptr;
    h->d_roww = nullptr; h->d_cpart = nullptr; h->n_cpart = 0; h->yoff_w = 0.0;
    h->d_xchg = nullptr; std::memset(&h->xchg, 0, sizeof(h->xchg));
    h->h_out = nullptr; h->h_out_dev = nullptr; h->out_seq = 0;
    h->ytab_valid = false; h->ytab_phis = 0.0;
    for (int v = 0; v < GLMMA_NTILES; ++v) { h->d_order[v] = nullptr; h->n_order[v] = 0; }
    for (auto& e : h->ev) e = nullptr;
    std::memset(&h->d, 0, sizeof(h->d));
    DevData& d = h->d;
    d.N = N; d.n = n; d.p = data->p; d.q_y = data->q_y; d.p_zi = data->p_zi; d.q_zi = data->q_zi; d.n_phis = data->n_phis;

#define TRY(expr) do { int rc_ = (expr); if (rc_) { g_create_error = h->error; glmma_destroy(h); return rc_; } } while (0)

    // response transform (see families.cuh header)
    std::vector<double> y1(N), y2;
    const double* ya = data->y;
    const double* yb = data->y_cols == 2 ? data->y + N : nullptr;
    switch (data->family) {
        case GLMMA_BINOMIAL: case GLMMA_ZI_BINOMIAL: case GLMMA_BETA_BINOMIAL:
            y2.resize(N);
            for (int64_t r = 0; r < N; ++r) { y1[r] = ya[r]; y2[r] = yb ? ya[r] + yb[r] : 1.0; }
            break;
        case GLMMA_GAMMA: case GLMMA_HURDLE_LOGNORMAL:
            y2.resize(N);
            for (int64_t r = 0; r < N; ++r) { y1[r] = ya[r]; y2[r] = ya[r] > 0.0 ? std::log(ya[r]) : -INFINITY; }
            break;
        case GLMMA_BETA: case GLMMA_HURDLE_BETA:
            y2.resize(N);
            for (int64_t r = 0; r < N; ++r) { y1[r] = ya[r] > 0.0 ? std::log(ya[r]) : -INFINITY; y2[r] = std::log1p(-ya[r]); }
            break;
        case GLMMA_CENSORED_NORMAL:
            y2.resize(N);
            for (int64_t r = 0; r < N; ++r) { y1[r] = ya[r]; y2[r] = yb[r]; }
            break;
        default:
            for (int64_t r = 0; r < N; ++r) y1[r] = ya[r];
    }
    TRY(dev_upload(h, &d.y1, y1.data(), N));
    if (!y2.empty()) TRY(dev_upload(h, &d.y2, y2.data(), N));
    TRY(dev_upload(h, &d.X, data->X, (size_t)N * d.p));
    TRY(dev_upload(h, &d.Z, data->Z, (size_t)N * d.q_y));
    if (d.p_zi) TRY(dev_upload(h, &d.X_zi, data->X_zi, (size_t)N * d.p_zi));
    if (d.q_zi) TRY(dev_upload(h, &d.Z_zi, data->Z_zi, (size_t)N * d.q_zi));
    if (data->offset) TRY(dev_upload(h, &d.offset, data->offset, N));
    if (data->offset_zi && ops->has_zi) TRY(dev_upload(h, &d.offset_zi, data->offset_zi, N));
    if (data->weights) TRY(dev_upload(h, &d.weights, data->weights, n));
    TRY(dev_upload(h, &d.row_ptr, row_ptr.data(), n + 1));

    const int np = q * (q + 1) / 2;
    TRY(dev_alloc(h, &d.xb, N));
    if (ops->has_zi) TRY(dev_alloc(h, &d.xg, N));
    TRY(dev_alloc(h, &d.cll, N));
    if (ops->has_phi) TRY(dev_alloc(h, &d.cphi, N));
    TRY(dev_alloc(h, &d.zbar, N));
    if (ops->has_zi) TRY(dev_alloc(h, &d.zbar_zi, N));
    TRY(dev_alloc(h, &d.modes, (size_t)n * q));
    TRY(dev_alloc(h, &d.rinv, (size_t)n * np));
    TRY(dev_alloc(h, &d.chol, (size_t)n * np));
    TRY(dev_alloc(h, &d.logdet, n));
    cudaMemsetAsync(d.modes, 0, sizeof(double) * n * q, h->stream);

    // Gauss-Hermite table
    GridPar g;
    std::memset(&g, 0, sizeof(g));
    double gw[GLMMA_NAGQ_MAX];
    gauher_table(h->nagq, g.x, gw);
    for (int a = 0; a < h->nagq; ++a) g.lw[a] = std::log(gw[a]) + g.x[a] * g.x[a];
    g.lw_const = 0.5 * q * std::log(2.0);
    g.nagq = h->nagq;
    g.K = h->K;
    TRY(dev_alloc(h, &h->d_grid, 1));
    cudaMemcpyAsync(h->d_grid, &g, sizeof(g), cudaMemcpyHostToDevice, h->stream);

    {
        std::vector<double> fm(320);
        for (int t = 0; t < 256; ++t) fm[t] = GLMMA_LOGTAB_HOST[t];
        for (int t = 0; t < 64; ++t) fm[256 + t] = GLMMA_EXPTAB_HOST[t];
        TRY(dev_alloc(h, &h->d_fmtab, 320));
        cudaMemcpy(h->d_fmtab, fm.data(), sizeof(double) * 320, cudaMemcpyHostToDevice);
    }

    // launch shape: the observation tile covers (nearly) all clusters; larger ones stream
    int max_ni = 0;
    std::vector<int> cnt(GLMMA_JT_CAP + 2, 0);
    for (int64_t i = 0; i < n; ++i) {
        const int ni = row_ptr[i + 1] - row_ptr[i];
        max_ni = std::max(max_ni, ni);
        cnt[std::min(ni, GLMMA_JT_CAP + 1)]++;
    }
    d.max_ni = max_ni;
    int jt = 1;
    {
        int64_t cum = 0;
        for (int v = 1; v <= GLMMA_JT_CAP; ++v) {
            cum += cnt[v];
            jt = v;
            if (cum >= (int64_t)std::ceil(0.995 * (double)n)) break;
        }
    }
    h->jt = jt;
    // clusters larger than the staged tile: K node sums per warp (kernels.cuh, tile-outer path)
    h->asum_doubles = (max_ni > jt && h->K <= GLMMA_ASUM_MAX_K) ? ((h->K + 1) & ~1) : 0;
    cudaDeviceProp prop;
    ce = cudaGetDeviceProperties(&prop, device);
    if (ce != cudaSuccess) { cuda_fail(h, ce, "cudaGetDeviceProperties"); g_create_error = h->error; glmma_destroy(h); return GLMMA_ERR_CUDA; }
    int max_grid = 1;
    for (int s = 0; s < 2; ++s) {
        int bps = 0;
        ce = ops->eval_config(s, jt, h->nagq, h->asum_doubles, &bps, &h->eval_smem[s], &h->eval_threads[s]);
        if (ce != cudaSuccess || bps < 1) {
            cuda_fail(h, ce, "eval kernel configuration");
            g_create_error = h->error; glmma_destroy(h); return GLMMA_ERR_CUDA;
        }
        const int wpb = h->eval_threads[s] / 32;
        const int64_t want = (n + wpb - 1) / wpb;
        h->eval_grid[s] = (int)std::min<int64_t>(want, (int64_t)prop.multiProcessorCount * bps);
        max_grid = std::max(max_grid, h->eval_grid[s]);
    }
    // register variant for small clusters (eval_fast_kernel)
    // Register variant (eval_fast_kernel) by cluster size: tiles of 8, 12 and 16 observations, the
    // generic kernel beyond.  A tile is only used when it serves a worthwhile share of the clusters;
    // ragged data run several tiles back to back, each on its own list of clusters.
    h->nj = 0;
    int max_fast = 0;
    if (q <= 3 && h->nagq <= GLMMA_NAGQ_FAST && h->K <= GLMMA_K_FAST && !std::getenv("GLMMA_NO_FAST") &&
        d.p + d.p_zi <= GLMMA_FOLD_COLS && d.p + d.p_zi + 2 + 2 + GLMMA_QMAX <= GLMMA_COLTAB_MAX) {
        int64_t n_le8 = 0, n_9_12 = 0, n_13_16 = 0;
        for (int v = 1; v <= 8; ++v) n_le8 += cnt[v];
        for (int v = 9; v <= 12; ++v) n_9_12 += cnt[v];
        for (int v = 13; v <= 16; ++v) n_13_16 += cnt[v];
        const int64_t few = (int64_t)(0.005 * (double)n);
        const bool use12 = n_9_12 > 0 && n_9_12 >= (int64_t)(0.05 * (double)n);
        const bool use16 = (n_13_16 > 0 && n_13_16 >= few) || (!use12 && n_9_12 > 0 && n_9_12 >= few);
        const bool use8 = n_le8 > 0 && (n_le8 >= (int64_t)(0.1 * (double)n) || !(use12 || use16));
        int64_t n_17_up = 0;
        for (int v = 17; v <= GLMMA_JT_CAP + 1; ++v) n_17_up += cnt[v];        // (the last bin holds everything beyond)
        const bool use32 = ops->tile32 && n_17_up > 0 && n_17_up >= few && !std::getenv("GLMMA_NO_TILE32");
        // pair tile: random intercepts, nAGQ <= 16 nodes, and enough small clusters to pair up
        const bool usePair = ops->pair && q == 1 && h->K <= 16 && n_le8 >= (int64_t)(0.1 * (double)n) && n_le8 >= 2 &&
                             !std::getenv("GLMMA_NO_PAIR");
        h->nj = (use8 ? 1 : 0) | (use12 ? 2 : 0) | (use16 ? 4 : 0) | (use32 ? 8 : 0) | (usePair ? 16 : 0);
    }
    // fused mode whenever the register variant runs and the design columns fit a warp's lanes
    h->xw = d.p + d.p_zi + (d.offset ? 1 : 0) + (d.offset_zi ? 1 : 0);
    h->fused = h->nj != 0;          // the register variant is self-contained (kernels.cuh): no prep / score-reduce / finalize launches
    if (h->nj) {
        for (int v = 0; v < GLMMA_NTILES; ++v) {
            if (!(h->nj & (1 << v))) { h->fast_grid[v][0] = h->fast_grid[v][1] = 0; continue; }
            for (int s = 0; s < 2; ++s) {
                int bps = 0;
                ce = ops->fast_config(h->xw, kTile[v], h->K, h->nagq, &bps, &h->fast_smem[v][s], &h->fast_threads[v][s]);
                if (ce != cudaSuccess || bps < 1) {
                    cuda_fail(h, ce, "fast eval kernel configuration");
                    g_create_error = h->error; glmma_destroy(h); return GLMMA_ERR_CUDA;
                }
                h->fast_grid[v][s] = prop.multiProcessorCount * bps;      // trimmed to the tile's list below
            }
            max_fast += std::max(h->fast_grid[v][0], h->fast_grid[v][1]);
        }
        TRY(dev_alloc(h, &d.defer, n));
        TRY(dev_alloc(h, &d.csum, (size_t)n * GLMMA_NCSUM));
        TRY(dev_alloc(h, &d.clrec, (size_t)n * (q + np + 1 + d.q_y + 1)));
        TRY(dev_alloc(h, &h->d_ytab, 2 * GLMMA_YTAB));
        // Size classes.  Each tile in use gets the list of the clusters it is the smallest fit for (the
        // identity when all clusters fit one tile) and the clusters beyond the largest tile are flagged
        // for eval_kernel once and for all (no launch of the register variant ever visits them).
        int top = 0;
        for (int v = 0; v <= GLMMA_TILE32; ++v) if (h->nj & (1 << v)) top = kTile[v];
        {
            std::vector<int4> lists[GLMMA_NTILES];
            std::vector<int32_t> flags(n, 0), too_big;
            const bool tp = (h->nj & 8) != 0;        // two-pass tile in use: it also takes the clusters beyond 32 rows
            std::vector<int32_t> tp_clusters;
            const bool pairs = (h->nj & (1 << GLMMA_TILEPAIR)) != 0;
            for (int64_t i = 0; i < n; ++i) {
                const int ni = row_ptr[i + 1] - row_ptr[i];
                if (pairs && ni <= 8 && i + 1 < n && row_ptr[i + 2] - row_ptr[i + 1] <= 8) {
                    // clusters i and i + 1 share a warp: {i, first row, n_A + n_B, n_A}
                    const int nb = row_ptr[i + 2] - row_ptr[i + 1];
                    lists[GLMMA_TILEPAIR].push_back(make_int4((int)i, row_ptr[i], ni + nb, ni));
                    ++i;
                    continue;
                }
                if (ni > top && !tp) { flags[i] = 1; too_big.push_back((int32_t)i); continue; }
                bool placed = false;
                for (int v = 0; v < GLMMA_TILE32 && !placed; ++v)           // the smallest one-pass tile in use that holds the cluster
                    if ((h->nj & (1 << v)) && ni <= kTile[v]) { lists[v].push_back(make_int4((int)i, row_ptr[i], ni, 0)); placed = true; }
                if (!placed) tp_clusters.push_back((int32_t)i);
            }
            if (tp) {
                // Two-pass tile (kernels.cuh): a cluster is a sequence of records -- its tiles of 32 rows for pass A
                // (flags 1, +2 first tile, +4 last tile), then the same tiles for pass B (8); 15 = all in one record
                // for at most 32 rows -- that ONE warp must walk in order.  The kernel gives position p of the list to
                // warp p mod nwarps, so the records are laid out as per-warp queues (clusters dealt to the least
                // loaded warp, longest first), padded with empty records to a common length.
                const int v = GLMMA_TILE32;
                const int fwpb = h->fast_threads[v][1] / 32;
                const int64_t want = ((int64_t)tp_clusters.size() + fwpb - 1) / fwpb;
                const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, h->fast_grid[v][1]));
                h->fast_grid[v][0] = h->fast_grid[v][1] = grid;
                const int64_t nwarps = (int64_t)grid * fwpb;
                std::vector<int32_t> byrows(tp_clusters);
                std::stable_sort(byrows.begin(), byrows.end(), [&](int32_t a, int32_t b) {
                    return row_ptr[a + 1] - row_ptr[a] > row_ptr[b + 1] - row_ptr[b]; });
                std::vector<std::vector<int4>> queue((size_t)nwarps);
                typedef std::pair<int64_t, int64_t> Load;       // (rows queued, warp)
                std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
                for (int64_t wq = 0; wq < nwarps; ++wq) heap.push(Load(0, wq));
                for (int32_t ci : byrows) {
                    const int r0 = row_ptr[ci], ni = row_ptr[ci + 1] - r0;
                    Load ld = heap.top(); heap.pop();
                    std::vector<int4>& qv = queue[(size_t)ld.second];
                    if (ni <= 32) qv.push_back(make_int4(ci, r0, ni, 15));
                    else {
                        const int T = (ni + 31) / 32;
                        for (int t = 0; t < T; ++t)
                            qv.push_back(make_int4(ci, r0 + 32 * t, std::min(32, ni - 32 * t), 1 | (t == 0 ? 2 : 0) | (t == T - 1 ? 4 : 0)));
                        for (int t = 0; t < T; ++t) qv.push_back(make_int4(ci, r0 + 32 * t, std::min(32, ni - 32 * t), 8));
                    }
                    ld.first += ni;
                    heap.push(ld);
                }
                size_t qlen = 0;
                for (const auto& qv : queue) qlen = std::max(qlen, qv.size());
                lists[v].assign(qlen * (size_t)nwarps, make_int4(0, 0, 0, 0));
                for (int64_t wq = 0; wq < nwarps; ++wq)
                    for (size_t m = 0; m < queue[(size_t)wq].size(); ++m) lists[v][m * (size_t)nwarps + (size_t)wq] = queue[(size_t)wq][m];
            }
            for (int v = 0; v < GLMMA_NTILES; ++v) {
                if (!(h->nj & (1 << v))) continue;
                h->n_order[v] = (int64_t)lists[v].size();
                const int4* tmp = nullptr;
                if (lists[v].empty()) lists[v].push_back(make_int4(0, 0, 0, 0));
                TRY(dev_upload(h, &tmp, lists[v].data(), lists[v].size()));
                h->d_order[v] = const_cast<int4*>(tmp);
                for (int s = 0; s < 2 && v != GLMMA_TILE32; ++s) {       // (the two-pass tile's grid is fixed above: its list is laid out for it)
                    const int fwpb = h->fast_threads[v][s] / 32;
                    const int64_t want = (h->n_order[v] + fwpb - 1) / fwpb;
                    h->fast_grid[v][s] = (int)std::max<int64_t>(1, std::min<int64_t>(want, h->fast_grid[v][s]));
                }
            }
            cudaMemcpyAsync(d.defer, flags.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->stream);
            if (h->fused) {
                // clusters beyond the largest tile keep the prep kernel (on their list only); Z'y1 of the
                // linear-eta hoist is data, not theta: computed here once (find_modes' prep pass rewrites it)
                h->n_defer_list = (int64_t)too_big.size();
                if (!too_big.empty()) {
                    const int32_t* tmp = nullptr;
                    TRY(dev_upload(h, &tmp, too_big.data(), too_big.size()));
                    h->d_defer_list = const_cast<int32_t*>(tmp);
                }
                // Z'y1 of the linear-eta hoist is data, not theta: one pass of the prep kernel with zero
                // coefficients fills it (the same kernel -- hence the same summation order -- that a mode
                // search's prep pass uses, so an engine whose grid was injected evaluates bit-identically to
                // one that searched for it; its theta-dependent outputs are rewritten before any use)
                {
                    PrepArgs pa;
                    std::memset(&pa, 0, sizeof(pa));
                    pa.d = h->d; pa.want_score = 0; pa.ytab = nullptr; pa.list = nullptr; pa.n_list = 0;
                    CoefPar c0;
                    std::memset(&c0, 0, sizeof(c0));
                    ops->prep(pa, c0, h->stream);
                }
                // node-independent terms are summed once per evaluation over the rows of this variant's
                // clusters (const_sums_kernel): per-row weights when clusters are weighted or excluded
                std::vector<double> roww;
                if (data->weights || !too_big.empty()) {
                    roww.assign((size_t)N, 0.0);
                    for (int64_t i = 0; i < n; ++i) {
                        if (flags[i]) continue;
                        const double wi = data->weights ? data->weights[i] : 1.0;
                        for (int64_t r = row_ptr[i]; r < row_ptr[i + 1]; ++r) roww[r] = wi;
                    }
                    const double* tmp = nullptr;
                    TRY(dev_upload(h, &tmp, roww.data(), roww.size()));
                    h->d_roww = const_cast<double*>(tmp);
                }
                h->n_cpart = (int)std::min<int64_t>((N + GLMMA_CONST_THREADS * 4 - 1) / (GLMMA_CONST_THREADS * 4), (int64_t)prop.multiProcessorCount * 8);
                TRY(dev_alloc(h, &h->d_cpart, (size_t)2 * h->n_cpart));
                // linear-eta hoist: sum_i w_i sum_j y1_ij (x_ij' beta + offset_ij) = beta' (X'y1)_w + (y1'offset)_w
                h->xty_w.assign(d.p, 0.0);
                if (ops->lin_eta) {
                    for (int l = 0; l < d.p; ++l) {
                        const double* col = data->X + (size_t)l * N;
                        double acc = 0.0;
                        if (roww.empty()) for (int64_t r = 0; r < N; ++r) acc = std::fma(y1[r], col[r], acc);
                        else for (int64_t r = 0; r < N; ++r) acc = std::fma(roww[r] * y1[r], col[r], acc);
                        h->xty_w[l] = acc;
                    }
                    if (data->offset) {
                        double acc = 0.0;
                        for (int64_t r = 0; r < N; ++r) acc = std::fma((roww.empty() ? 1.0 : roww[r]) * y1[r], data->offset[r], acc);
                        h->yoff_w = acc;
                    }
                }
                TRY(dev_alloc(h, &h->d_counter, 2));       // [0] blocks done, [1] clusters deferred in this evaluation
                cudaMemsetAsync(h->d_counter, 0, 2 * sizeof(unsigned), h->stream);
            }
            cudaStreamSynchronize(h->stream);      // lists / flags are block-local; errors surface at the final check
        }
    }
    max_grid += max_fast;
    h->score_grid = (int)std::min<int64_t>((N + SCORE_THREADS * 8 - 1) / (SCORE_THREADS * 8), (int64_t)prop.multiProcessorCount * 4);
    TRY(dev_alloc(h, &h->d_partials, (size_t)max_grid * GLMMA_PSTRIDE));
    TRY(dev_alloc(h, &h->d_score_partials, (size_t)h->score_grid * (d.p + d.p_zi)));
    TRY(dev_alloc(h, &h->d_result, (size_t)glmma_result_len(h)));
    TRY(dev_alloc(h, &h->d_stats, 4));
    for (auto& e : h->ev) {
        ce = cudaEventCreate(&e);
        if (ce != cudaSuccess) { cuda_fail(h, ce, "cudaEventCreate"); g_create_error = h->error; glmma_destroy(h); return GLMMA_ERR_CUDA; }
    }
    ce = cudaStreamSynchronize(h->stream);
    if (ce != cudaSuccess) { cuda_fail(h, ce, "upload"); g_create_error = h->error; glmma_destroy(h); return GLMMA_ERR_CUDA; }
#undef TRY
    *out = h;
    return GLMMA_OK;
}

int glmma_set_stream(glmma_handle* h, void* cuda_stream) {
    if (!h) return GLMMA_ERR_ARG;
    if (!h->parts.empty()) return fail(h, GLMMA_ERR_ARG, "a multi-device handle runs on each device's default stream");
    h->stream = (cudaStream_t)cuda_stream;
    return GLMMA_OK;
}

int64_t glmma_n_clusters(const glmma_handle* h) { return !h ? 0 : (h->parts.empty() ? h->d.n : h->part_c0.back()); }
int64_t glmma_n_obs(const glmma_handle* h) { return !h ? 0 : (h->parts.empty() ? h->d.N : h->part_r0.back()); }
int glmma_nre(const glmma_handle* h) { return h ? h->q : 0; }
int glmma_n_nodes(const glmma_handle* h) { return h ? h->K : 0; }
int glmma_result_len(const glmma_handle* h) {
    if (!h) return 0;
    if (!h->parts.empty()) return glmma_result_len(h->parts[0]);
    return 4 + h->d.p + h->d.n_phis + h->d.p_zi + h->q * h->q;
}
int64_t glmma_launch_count(const glmma_handle* h) {
    if (!h) return 0;
    int64_t t = h->launches;
    for (const glmma_handle* p : h->parts) t += p->launches;
    return t;
}
int glmma_n_devices(const glmma_handle* h) { return !h ? 0 : (h->parts.empty() ? 1 : (int)h->parts.size()); }

static int enter(glmma_handle* h) {
    if (!h) return GLMMA_ERR_ARG;
    if (h->sticky) return h->sticky;
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaSetDevice");
    return GLMMA_OK;
}

static int fill_prior(glmma_handle* h, const double* M, bool is_inverse, PriorPar* pr) {
    // M is D (eval) or invD (find_modes), q x q
    const int q = h->q;
    std::memset(pr, 0, sizeof(*pr));
    if (is_inverse) {
        for (int r = 0; r < q; ++r)
            for (int c = 0; c < q; ++c) pr->Dinv[r * q + c] = 0.5 * (M[r + c * q] + M[c + r * q]);
        pr->logprior_const = 0.0;
        return GLMMA_OK;
    }
    double A[GLMMA_QMAX * GLMMA_QMAX], Ai[GLMMA_QMAX * GLMMA_QMAX], ld = 0.0;
    for (int r = 0; r < q; ++r)
        for (int c = 0; c < q; ++c) A[r * q + c] = M[r + c * q];
    if (!spd_inverse(A, q, Ai, &ld)) return fail(h, GLMMA_ERR_NUMERIC, "D is not positive definite");
    for (int i = 0; i < q * q; ++i) pr->Dinv[i] = Ai[i];
    pr->logprior_const = -0.5 * (q * GLMMA_LOG_2PI + ld);
    return GLMMA_OK;
}

static int fill_coef(glmma_handle* h, const double* betas, const double* gammas, CoefPar* c) {
    std::memset(c, 0, sizeof(*c));
    if (!betas) return fail(h, GLMMA_ERR_ARG, "betas is null");
    for (int l = 0; l < h->d.p; ++l) c->betas[l] = betas[l];
    if (h->d.p_zi) {
        if (!gammas) return fail(h, GLMMA_ERR_ARG, "gammas is null");
        for (int l = 0; l < h->d.p_zi; ++l) c->gammas[l] = gammas[l];
    }
    return GLMMA_OK;
}

// table of the node-independent terms for y = 0 .. GLMMA_YTAB - 1 (count families): a function of the
// family and phis only, so it is rebuilt only when phis changed (never for poisson / zi.poisson)
static void ensure_ytab(glmma_handle* h, const FamPar& fp, const double* phis) {
    if (!h->d_ytab || !h->ops->y_table) return;
    const double key = (h->d.n_phis > 0 && phis) ? phis[0] : 0.0;
    if (h->ytab_valid && key == h->ytab_phis) return;
    h->ops->ytab(fp, h->d_ytab, h->stream);
    h->launches++;
    h->ytab_valid = true;
    h->ytab_phis = key;
}

static int run_prep(glmma_handle* h, const double* betas, const double* phis, const double* gammas, FamPar* fp) {
    CoefPar c;
    int rc = fill_coef(h, betas, gammas, &c);
    if (rc) return rc;
    if (!fill_fampar(h->family, h->d.n_phis, phis, fp, h->error, h->family_df)) return GLMMA_ERR_ARG;
    PrepArgs pa;
    pa.d = h->d; pa.fam = *fp; pa.want_score = 1; pa.ytab = h->d.csum ? h->d_ytab : nullptr;
    pa.list = nullptr; pa.n_list = 0;
    ensure_ytab(h, *fp, phis);
    h->ops->prep(pa, c, h->stream);
    h->launches++;
    CU(h, cudaGetLastError());
    return GLMMA_OK;
}

// launches the mode search and the copy of its statistics into st_host (no synchronisation)
static int find_modes_enqueue(glmma_handle* h, const double* betas, const double* invD, const double* phis,
                              const double* gammas, int warm, unsigned long long* st_host) {
    int rc = enter(h);
    if (rc) return rc;
    if (!invD) return fail(h, GLMMA_ERR_ARG, "invD is null");
    FamPar fp;
    rc = run_prep(h, betas, phis, gammas, &fp);
    if (rc) return rc;
    ModeArgs ma;
    ma.d = h->d; ma.fam = fp; ma.fmtab = h->d_fmtab; fastmath_constants(ma.fmc); ma.warm = warm && h->have_modes; ma.maxit = 100; ma.tol = 1e-10; ma.stats = h->d_stats;
    fill_prior(h, invD, true, &ma.prior);
    CU(h, cudaMemsetAsync(h->d_stats, 0, 4 * sizeof(unsigned long long), h->stream));
    const double mean_ni = (double)h->d.N / (double)h->d.n;
    int G = mean_ni <= 12.0 ? 1 : (mean_ni <= 96.0 ? 8 : 32);
    if (const char* g_env = std::getenv("GLMMA_MODES_G")) { const int g = std::atoi(g_env); if (g == 1 || g == 8 || g == 32) G = g; }
    h->ops->modes(ma, G, h->stream);
    h->launches++;
    // the grid is DEFINED by (modes, chol): R^-1 and log_dets are re-derived from the stored factors by
    // the same kernel glmma_set_modes uses, so that get_modes -> set_modes (the round trip the R glue
    // makes when it re-injects a GH object) reproduces every later evaluation bit for bit
    h->ops->rinv(h->d, h->stream);
    h->launches++;
    if (h->d.clrec) { h->ops->pack(h->d, h->stream); h->launches++; }
    CU(h, cudaGetLastError());
    CU(h, cudaMemcpyAsync(st_host, h->d_stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    h->have_modes = true;
    h->have_pby = false;
    return GLMMA_OK;
}

static int multi_find_modes(glmma_handle* h, const double* betas, const double* invD, const double* phis,
                            const double* gammas, int warm, glmma_mode_stats* stats);

int glmma_find_modes(glmma_handle* h, const double* betas, const double* invD, const double* phis,
                     const double* gammas, int warm, glmma_mode_stats* stats) {
    DeviceGuard device_guard_;
    NvtxRange nvtx_range_("glmma_find_modes");
    if (h && !h->parts.empty()) return multi_find_modes(h, betas, invD, phis, gammas, warm, stats);
    unsigned long long st[4];
    int rc = find_modes_enqueue(h, betas, invD, phis, gammas, warm, st);
    if (rc) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    if (stats) {
        stats->n_not_converged = (int64_t)st[0];
        stats->n_chol_failed = (int64_t)st[1];
        stats->max_iterations = (int32_t)st[3];
        stats->reserved = 0;
        stats->mean_iterations = (double)st[2] / (double)h->d.n;
    }
    return GLMMA_OK;
}

static int multi_get_modes(glmma_handle* h, double* post_modes, double* chol_upper, double* log_dets);
static int multi_set_modes(glmma_handle* h, const double* post_modes, const double* chol_upper);
static int multi_eval(glmma_handle* h, const double* betas, const double* D, const double* phis, const double* gammas,
                      int want_score, double* result, int em_mode);
static int multi_eval_contrib(glmma_handle* h, const double* betas, const double* D, const double* phis,
                              const double* gammas, double* ll_i, double* zbar, double* zbar_zi, double* sphi_obs,
                              double* S_i);
static int multi_glm_pass(glmma_handle* h, int design, int glm_family, int response, int use_offset, int first,
                          const double* beta, double* out);
static int multi_fail(glmma_handle* h, const glmma_handle* p, int rc);

int glmma_get_modes(glmma_handle* h, double* post_modes, double* chol_upper, double* log_dets) {
    DeviceGuard device_guard_;
    NvtxRange nvtx_range_("glmma_get_modes");
    if (h && !h->parts.empty()) return multi_get_modes(h, post_modes, chol_upper, log_dets);
    int rc = enter(h);
    if (rc) return rc;
    if (!h->have_modes) return fail(h, GLMMA_ERR_STATE, "no modes yet: call glmma_find_modes or glmma_set_modes first");
    const int64_t n = h->d.n;
    const int q = h->q, np = q * (q + 1) / 2;
    if (post_modes) {
        // device layout is cluster-major (n rows of q); R wants an n x q column-major matrix
        std::vector<double> tmp((size_t)n * q);
        CU(h, cudaMemcpyAsync(tmp.data(), h->d.modes, sizeof(double) * n * q, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        for (int64_t i = 0; i < n; ++i)
            for (int dd = 0; dd < q; ++dd) post_modes[i + dd * n] = tmp[i * q + dd];
    }
    if (chol_upper) CU(h, cudaMemcpyAsync(chol_upper, h->d.chol, sizeof(double) * n * np, cudaMemcpyDeviceToHost, h->stream));
    if (log_dets) CU(h, cudaMemcpyAsync(log_dets, h->d.logdet, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return GLMMA_OK;
}

int glmma_set_modes(glmma_handle* h, const double* post_modes, const double* chol_upper) {
    DeviceGuard device_guard_;
    NvtxRange nvtx_range_("glmma_set_modes");
    if (h && !h->parts.empty()) return multi_set_modes(h, post_modes, chol_upper);
    int rc = enter(h);
    if (rc) return rc;
    if (!post_modes || !chol_upper) return fail(h, GLMMA_ERR_ARG, "null argument");
    const int64_t n = h->d.n;
    const int q = h->q, np = q * (q + 1) / 2;
    std::vector<double> tmp((size_t)n * q);
    for (int64_t i = 0; i < n; ++i)
        for (int dd = 0; dd < q; ++dd) tmp[i * q + dd] = post_modes[i + dd * n];
    CU(h, cudaMemcpyAsync(h->d.modes, tmp.data(), sizeof(double) * n * q, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->d.chol, chol_upper, sizeof(double) * n * np, cudaMemcpyHostToDevice, h->stream));
    h->ops->rinv(h->d, h->stream);
    h->launches++;
    if (h->d.clrec) { h->ops->pack(h->d, h->stream); h->launches++; }
    CU(h, cudaGetLastError());
    CU(h, cudaStreamSynchronize(h->stream));
    h->have_modes = true;
    h->have_pby = false;       // the stored posterior node weights belong to the previous grid
    return GLMMA_OK;
}

// enqueue one evaluation on the stream; the result vector ends up at d_out
static int enqueue_eval(glmma_handle* h, const double* betas, const double* D, const double* phis,
                        const double* gammas, int want_score, double* d_out, const DevData* dd_override,
                        cudaEvent_t k0, cudaEvent_t k1, int em_mode = 0, bool to_host = false) {
    // (the per-observation contributions -- dd_override -- stay local to the rank: no exchange)
    if (!h->have_modes) return fail(h, GLMMA_ERR_STATE, "no modes yet: call glmma_find_modes or glmma_set_modes first");
    if (!D) return fail(h, GLMMA_ERR_ARG, "D is null");
    const int s = want_score ? 1 : 0;
    const bool fused = h->fused;
    FamPar fp;
    int rc;
    EvalArgs ea;
    std::memset(&ea, 0, sizeof(ea));
    if (fused) {
        rc = fill_coef(h, betas, gammas, &ea.coef);
        if (rc) return rc;
        if (!fill_fampar(h->family, h->d.n_phis, phis, &fp, h->error, h->family_df)) return GLMMA_ERR_ARG;
        ensure_ytab(h, fp, phis);
        // prep pass only over the clusters no tile takes -- or over all of them when the per-observation
        // contributions are wanted (sphi_obs_kernel reads the work arrays)
        if (h->n_defer_list > 0 || dd_override) {
            PrepArgs pa;
            pa.d = h->d; pa.fam = fp; pa.want_score = 1; pa.ytab = h->d_ytab;
            pa.list = dd_override ? nullptr : h->d_defer_list; pa.n_list = dd_override ? 0 : h->n_defer_list;
            h->ops->prep(pa, ea.coef, h->stream);
            h->launches++;
        }
        ea.fused = 1; ea.xw = h->xw; ea.zbar_out = dd_override ? 1 : 0; ea.ytab = h->ops->y_table ? h->d_ytab : nullptr;
        ea.fin_counter = h->d_counter; ea.fin_partials = h->d_partials; ea.fin_want_score = s; ea.fin_result = d_out;
        // (per-observation contributions walk every cluster in eval_kernel's work arrays: no short cut there)
        ea.defer_count = dd_override ? nullptr : h->d_counter + 1; ea.defer_static = h->n_defer_list > 0 ? 1 : 0;
        // node-independent terms of the register variant's clusters, once per evaluation
        h->ops->consts(h->d, fp, h->d_ytab, h->d_roww, h->d_cpart, h->n_cpart, h->stream);
        h->launches++;
        ea.fin_cpart = h->d_cpart; ea.fin_ncpart = h->n_cpart;
        double llc = h->yoff_w;
        for (int l = 0; l < h->d.p; ++l) llc = std::fma(betas[l], h->xty_w[l], llc);
        ea.fin_ll_const = llc;
    } else {
        rc = run_prep(h, betas, phis, gammas, &fp);
        if (rc) return rc;
    }
    ea.em_mode = em_mode; ea.pby = h->d_pby;
    if (h->xchg.nranks > 1 && !dd_override) {
        ea.xchg = h->xchg;
        ea.xchg.seq = ++h->xchg.seq;
    }
    if (to_host && h->h_out_dev) {
        ea.hout.vec = h->h_out_dev;
        ea.hout.flag = reinterpret_cast<unsigned long long*>(h->h_out_dev + GLMMA_XSLOT - 1);
        ea.hout.seq = ++h->out_seq;
    }
    ea.d = dd_override ? *dd_override : h->d;
    ea.fam = fp; ea.grid = h->d_grid; ea.fmtab = h->d_fmtab; fastmath_constants(ea.fmc); ea.partials = h->d_partials; ea.jt = h->jt; ea.asum_doubles = h->asum_doubles;
    rc = fill_prior(h, D, false, &ea.prior);
    if (rc) return rc;
    if (k0) CU(h, cudaEventRecord(k0, h->stream));
    int n_blocks = 0;
    ea.only_deferred = 0;
    if (h->nj) {
        // register variant, smaller tile first; each launch flags what it leaves to the next one
        for (int v = 0; v < GLMMA_NTILES; ++v) {
            if (!(h->nj & (1 << v))) continue;
            ea.order = h->d_order[v];
            ea.n_order = h->n_order[v];
            ea.partials = h->d_partials + (size_t)n_blocks * GLMMA_PSTRIDE;
            h->ops->eval_fast(ea, s, kTile[v], h->fast_grid[v][s], h->fast_threads[v][s], h->fast_smem[v][s], h->stream);
            h->launches++;
            n_blocks += h->fast_grid[v][s];
        }
        ea.only_deferred = 1;
        ea.partials = h->d_partials + (size_t)n_blocks * GLMMA_PSTRIDE;
    }
    ea.fin_blocks = n_blocks + h->eval_grid[s];
    h->ops->eval(ea, s, h->eval_grid[s], h->eval_threads[s], h->eval_smem[s], h->stream);
    n_blocks += h->eval_grid[s];
    if (k1) CU(h, cudaEventRecord(k1, h->stream));
    h->launches++;
    CU(h, cudaGetLastError());
    if (fused) return GLMMA_OK;            // eval_kernel's last block wrote the result vector
    if (s) {
        score_reduce_kernel<<<h->score_grid, SCORE_THREADS, 0, h->stream>>>(ea.d, h->d_score_partials);
        h->launches++;
    }
    finalize_kernel<<<1, FINAL_THREADS, 0, h->stream>>>(h->d_partials, n_blocks, h->d_score_partials, h->score_grid,
                                                         h->d.p, h->d.n_phis, h->d.p_zi, h->q, s, d_out, ea.xchg, ea.hout);
    h->launches++;
    CU(h, cudaGetLastError());
    return GLMMA_OK;
}

// Enqueues one evaluation whose result vector the finalising block delivers into pinned host memory,
// waits for it (spinning on the flag; the stream is polled so that a failed launch cannot hang the host)
// and copies `len` doubles to `result`.
static int eval_to_host(glmma_handle* h, const double* betas, const double* D, const double* phis, const double* gammas,
                        int want_score, double* result, int len, int em_mode) {
    if (!h->h_out) {
        // mapped pinned memory; with unified addressing the host pointer is also the device pointer
        cudaError_t e = cudaHostAlloc((void**)&h->h_out, sizeof(double) * GLMMA_XSLOT, cudaHostAllocMapped);
        if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&h->h_out_dev, h->h_out, 0);
        if (e != cudaSuccess) { h->h_out = nullptr; h->h_out_dev = nullptr; cudaGetLastError(); }
        else std::memset(h->h_out, 0, sizeof(double) * GLMMA_XSLOT);
    }
    int rc = enqueue_eval(h, betas, D, phis, gammas, want_score, h->d_result, nullptr, nullptr, nullptr, em_mode, true);
    if (rc) return rc;
    if (!h->h_out) {        // no mapped memory: the plain copy
        CU(h, cudaMemcpyAsync(result, h->d_result, sizeof(double) * len, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        return GLMMA_OK;
    }
    volatile unsigned long long* flag = reinterpret_cast<volatile unsigned long long*>(h->h_out + GLMMA_XSLOT - 1);
    for (unsigned spins = 0; *flag != h->out_seq; ++spins) {
        if ((spins & 0x3fffu) == 0x3fffu) {
            const cudaError_t q = cudaStreamQuery(h->stream);
            if (q == cudaSuccess) {                  // the stream drained: the flag is there, or the kernel never ran
                if (*flag != h->out_seq) return fail(h, GLMMA_ERR_CUDA, "evaluation finished without delivering its result");
                break;
            }
            if (q != cudaErrorNotReady) return cuda_fail(h, q, "eval");
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    for (int t = 0; t < len; ++t) result[t] = h->h_out[t];
    return GLMMA_OK;
}

int glmma_eval_device(glmma_handle* h, const double* betas, const double* D, const double* phis,
                      const double* gammas, int want_score, double* d_result) {
    DeviceGuard device_guard_;
    NvtxRange nvtx_range_("glmma_eval_device");
    if (h && !h->parts.empty()) return fail(h, GLMMA_ERR_ARG, "glmma_eval_device needs a single-device handle");
    int rc = enter(h);
    if (rc) return rc;
    if (!d_result) return fail(h, GLMMA_ERR_ARG, "d_result is null");
    return enqueue_eval(h, betas, D, phis, gammas, want_score, d_result, nullptr, nullptr, nullptr);
}

int glmma_eval(glmma_handle* h, const double* betas, const double* D, const double* phis,
               const double* gammas, int want_score, double* result) {
    DeviceGuard device_guard_;
    NvtxRange nvtx_range_("glmma_eval");
    if (h && !h->parts.empty()) return multi_eval(h, betas, D, phis, gammas, want_score, result, 0);
    int rc = enter(h);
    if (rc) return rc;
    if (!result) return fail(h, GLMMA_ERR_ARG, "result is null");
    return eval_to_host(h, betas, D, phis, gammas, want_score, result, want_score ? glmma_result_len(h) : 4, 0);
}

int glmma_eval_contrib(glmma_handle* h, const double* betas, const double* D, const double* phis,
                       const double* gammas, double* ll_i, double* zbar, double* zbar_zi,
                       double* sphi_obs, double* S_i) {
    DeviceGuard device_guard_;
    NvtxRange nvtx_range_("glmma_eval_contrib");
    if (h && !h->parts.empty()) return multi_eval_contrib(h, betas, D, phis, gammas, ll_i, zbar, zbar_zi, sphi_obs, S_i);
    int rc = enter(h);
    if (rc) return rc;
    const int64_t n = h->d.n, N = h->d.N;
    const int q = h->q;
    DevData dd = h->d;
    double *d_ll = nullptr, *d_S = nullptr, *d_sphi = nullptr;
    cudaError_t e = cudaMalloc(&d_ll, sizeof(double) * n);
    if (e == cudaSuccess && S_i) e = cudaMalloc(&d_S, sizeof(double) * n * q * q);
    if (e == cudaSuccess && sphi_obs && h->ops->has_phi) e = cudaMalloc(&d_sphi, sizeof(double) * N);
    if (e != cudaSuccess) { cudaFree(d_ll); cudaFree(d_S); cudaFree(d_sphi); return cuda_fail(h, e, "cudaMalloc (contrib)"); }
    dd.ll_i = d_ll; dd.S_i = d_S; dd.sphi_obs = d_sphi;
    rc = enqueue_eval(h, betas, D, phis, gammas, 1, h->d_result, &dd, nullptr, nullptr);
    if (rc == GLMMA_OK && d_sphi) {
        EvalArgs ea;
        std::memset(&ea, 0, sizeof(ea));
        ea.d = dd; ea.grid = h->d_grid; ea.fmtab = h->d_fmtab; fastmath_constants(ea.fmc); ea.partials = h->d_partials; ea.jt = h->jt; ea.asum_doubles = h->asum_doubles; ea.only_deferred = 0;
        fill_fampar(h->family, h->d.n_phis, phis, &ea.fam, h->error, h->family_df);
        fill_prior(h, D, false, &ea.prior);
        h->ops->sphi(ea, h->eval_grid[1], h->eval_threads[1], h->eval_smem[1], h->stream);
        h->launches++;
    }
    if (rc == GLMMA_OK) {
        if (ll_i) cudaMemcpyAsync(ll_i, d_ll, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream);
        if (S_i) cudaMemcpyAsync(S_i, d_S, sizeof(double) * n * q * q, cudaMemcpyDeviceToHost, h->stream);
        if (zbar) cudaMemcpyAsync(zbar, h->d.zbar, sizeof(double) * N, cudaMemcpyDeviceToHost, h->stream);
        if (zbar_zi && h->d.zbar_zi) cudaMemcpyAsync(zbar_zi, h->d.zbar_zi, sizeof(double) * N, cudaMemcpyDeviceToHost, h->stream);
        if (sphi_obs && d_sphi) cudaMemcpyAsync(sphi_obs, d_sphi, sizeof(double) * N, cudaMemcpyDeviceToHost, h->stream);
        e = cudaStreamSynchronize(h->stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) rc = cuda_fail(h, e, "eval_contrib");
    }
    cudaFree(d_ll); cudaFree(d_S); cudaFree(d_sphi);
    return rc;
}

// ---- EM phase of mixed_fit (R/mixed_fit.R:119-229) ---------------------------------------

int glmma_em_estep(glmma_handle* h, const double* betas, const double* D, const double* phis,
                   const double* gammas, double* result) {
    DeviceGuard device_guard_;
    NvtxRange nvtx_range_("glmma_em_estep");
    if (h && !h->parts.empty()) return multi_eval(h, betas, D, phis, gammas, 1, result, 1);
    int rc = enter(h);
    if (rc) return rc;
    if (!result) return fail(h, GLMMA_ERR_ARG, "result is null");
    if (!h->d_pby) {
        rc = dev_alloc(h, &h->d_pby, (size_t)h->d.n * (size_t)h->K);
        if (rc) return rc;
    }
    rc = eval_to_host(h, betas, D, phis, gammas, 1, result, glmma_result_len(h), 1);
    if (rc) return rc;
    h->have_pby = true;
    return GLMMA_OK;
}

int glmma_em_score(glmma_handle* h, const double* betas, const double* phis, const double* gammas,
                   double* result) {
    DeviceGuard device_guard_;
    NvtxRange nvtx_range_("glmma_em_score");
    if (h && !h->parts.empty()) return multi_eval(h, betas, nullptr, phis, gammas, 1, result, 2);
    int rc = enter(h);
    if (rc) return rc;
    if (!result) return fail(h, GLMMA_ERR_ARG, "result is null");
    if (!h->have_pby) return fail(h, GLMMA_ERR_STATE, "no posterior weights yet: call glmma_em_estep first");
    // the prior only enters the node weights, which are frozen: any positive definite D will do
    double eye[GLMMA_QMAX * GLMMA_QMAX] = {0};
    for (int i = 0; i < h->q; ++i) eye[i * h->q + i] = 1.0;
    return eval_to_host(h, betas, eye, phis, gammas, 1, result, glmma_result_len(h), 2);
}

// ---- starting values of mixed_model() (R/mixed_model.R:157-194) --------------------------

int glmma_glm_pass(glmma_handle* h, int design, int glm_family, int response, int use_offset, int first,
                   const double* beta, double* out) {
    DeviceGuard device_guard_;
    NvtxRange nvtx_range_("glmma_glm_pass");
    if (h && !h->parts.empty()) return multi_glm_pass(h, design, glm_family, response, use_offset, first, beta, out);
    int rc = enter(h);
    if (rc) return rc;
    if (!out || (!first && !beta)) return fail(h, GLMMA_ERR_ARG, "null argument");
    const DevData& d = h->d;
    if (design != 0 && design != 1) return fail(h, GLMMA_ERR_ARG, "design must be 0 (X) or 1 (X_zi)");
    const int p = design == 0 ? d.p : d.p_zi;
    if (p < 1 || p > GLM_PMAX) return fail(h, GLMMA_ERR_ARG, "glmma_glm_pass supports 1..8 columns");
    if (glm_family < 0 || glm_family > 3) return fail(h, GLMMA_ERR_ARG, "glm_family must be 0..3");
    const bool y1_is_log = h->family == GLMMA_BETA || h->family == GLMMA_HURDLE_BETA;
    if (response == 0 && y1_is_log) return fail(h, GLMMA_ERR_ARG, "no glm start for the beta families' response");
    GlmArgs ga;
    std::memset(&ga, 0, sizeof(ga));
    ga.X = design == 0 ? d.X : d.X_zi;
    ga.y1 = d.y1;
    ga.y2 = (h->family == GLMMA_BINOMIAL || h->family == GLMMA_ZI_BINOMIAL || h->family == GLMMA_BETA_BINOMIAL) ? d.y2 : nullptr;
    ga.offset = use_offset ? (design == 0 ? d.offset : d.offset_zi) : nullptr;
    ga.N = d.N; ga.p = p; ga.fam = glm_family; ga.resp = response; ga.y1_is_log = y1_is_log ? 1 : 0;
    ga.first = first ? 1 : 0;
    for (int l = 0; l < GLM_PMAX; ++l) ga.beta[l] = (!first && l < p) ? beta[l] : 0.0;
    const int blocks = (int)std::min<int64_t>((d.N + GLM_THREADS - 1) / GLM_THREADS, 148 * 8);
    double* d_part = nullptr;
    cudaError_t e = cudaMalloc(&d_part, sizeof(double) * (size_t)blocks * GLM_NACC);
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaMalloc (glm)");
    ga.partials = d_part;
    std::vector<double> part((size_t)blocks * GLM_NACC), tot(GLM_NACC, 0.0);
    glm_irls_kernel<<<blocks, GLM_THREADS, 0, h->stream>>>(ga);
    h->launches++;
    e = cudaMemcpyAsync(part.data(), d_part, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_part);
    if (e != cudaSuccess) return cuda_fail(h, e, "glm_irls_kernel");
    for (int b = 0; b < blocks; ++b)          // fixed order: deterministic
        for (int t = 0; t < GLM_NACC; ++t) tot[t] += part[(size_t)b * GLM_NACC + t];
    out[0] = tot[GLM_NACC - 1];
    for (int l = 0; l < p; ++l) out[1 + l] = tot[GLM_PMAX * (GLM_PMAX + 1) / 2 + l];
    int t = 0;
    for (int l = 0; l < GLM_PMAX; ++l)
        for (int m = l; m < GLM_PMAX; ++m) {
            if (l < p && m < p) { out[1 + p + l + m * p] = tot[t]; out[1 + p + m + l * p] = tot[t]; }
            ++t;
        }
    return GLMMA_OK;
}

// ---- observed information (replaces cd_vec(score_mixed), R/mixed_fit.R:324-342) --------------------

// small dense helpers, q <= GLMMA_QMAX, row-major q x q
static void mm(const double* A, const double* B, int q, double* C) {
    double T[GLMMA_QMAX * GLMMA_QMAX];
    for (int r = 0; r < q; ++r)
        for (int c = 0; c < q; ++c) {
            double s = 0.0;
            for (int t = 0; t < q; ++t) s += A[r * q + t] * B[t * q + c];
            T[r * q + c] = s;
        }
    for (int i = 0; i < q * q; ++i) C[i] = T[i];
}

int glmma_observed_information(glmma_handle* h, const double* betas, const double* D, const double* phis,
                               const double* gammas, int diag_D, double* H) {
    NvtxRange nvtx_range_("glmma_observed_information");
    DeviceGuard device_guard_;
    if (!h || !H) return h ? fail(h, GLMMA_ERR_ARG, "null argument") : GLMMA_ERR_ARG;
    const int q = h->q;
    const int nD = diag_D ? q : q * (q + 1) / 2;
    if (!h->parts.empty()) {
        const int P = h->multi_p + nD + h->multi_nphis + h->multi_pzi;
        std::vector<double> tmp((size_t)P * P);
        for (int t = 0; t < P * P; ++t) H[t] = 0.0;
        for (glmma_handle* p : h->parts) {          // device order: deterministic
            int rc = glmma_observed_information(p, betas, D, phis, gammas, diag_D, tmp.data());
            if (rc) return multi_fail(h, p, rc);
            for (int t = 0; t < P * P; ++t) H[t] += tmp[t];
        }
        return GLMMA_OK;
    }
    int rc = enter(h);
    if (rc) return rc;
    if (!h->have_modes) return fail(h, GLMMA_ERR_STATE, "no modes yet: call glmma_find_modes or glmma_set_modes first");
    if (!D) return fail(h, GLMMA_ERR_ARG, "D is null");
    const int p = h->d.p, pz = h->d.p_zi, nphi = h->d.n_phis;
    const int P = p + nD + nphi + pz;
    if (P > GLMMA_INFO_PMAX) return fail(h, GLMMA_ERR_ARG, "more than GLMMA_INFO_PMAX parameters: use the finite-difference Hessian");
    InfoArgs ia;
    std::memset(&ia, 0, sizeof(ia));
    rc = run_prep(h, betas, phis, gammas, &ia.fam);
    if (rc) return rc;
    ia.eps_eta = 1e-4; ia.eps_phi = 1e-4;
    if (nphi) {
        const double hi = phis[0] + ia.eps_phi, lo = phis[0] - ia.eps_phi;
        if (!fill_fampar(h->family, nphi, &hi, &ia.fam_hi, h->error, h->family_df) ||
            !fill_fampar(h->family, nphi, &lo, &ia.fam_lo, h->error, h->family_df)) return GLMMA_ERR_ARG;
    } else { ia.fam_hi = ia.fam; ia.fam_lo = ia.fam; }
    rc = fill_prior(h, D, false, &ia.prior);
    if (rc) return rc;
    // log N(b; 0, D) = -log|D| / 2 - b' D^-1 b / 2 as a function of the D parameters theta_D:
    //   d/d theta_t       = -tr(D^-1 D_t) / 2 + b' (D^-1 D_t D^-1) b / 2
    //   d2/d theta_t d_u  = -tr(-D^-1 D_u D^-1 D_t + D^-1 D_tu) / 2
    //                       + b' (-D^-1 D_u D^-1 D_t D^-1 + D^-1 D_tu D^-1 - D^-1 D_t D^-1 D_u D^-1) b / 2
    // with D_t, D_tu the derivatives of the parametrisation: log(diag(D)) (random = ~ . || g), or
    // chol_transf(D) -- upper Cholesky factor U, D = U'U, diagonal logged, column-major over the upper
    // triangle (R/Functions.R:140-147).
    {
        double Dm[GLMMA_QMAX * GLMMA_QMAX], Di[GLMMA_QMAX * GLMMA_QMAX], U[GLMMA_QMAX * GLMMA_QMAX] = {0};
        for (int r = 0; r < q; ++r)
            for (int c = 0; c < q; ++c) { Dm[r * q + c] = D[r + c * q]; Di[r * q + c] = ia.prior.Dinv[r * q + c]; }
        // upper factor: D = U'U  (U = L' of the lower Cholesky factor)
        for (int c = 0; c < q; ++c) {
            double sdiag = Dm[c * q + c];
            for (int t = 0; t < c; ++t) sdiag -= U[t * q + c] * U[t * q + c];
            U[c * q + c] = std::sqrt(sdiag);
            for (int r = c + 1; r < q; ++r) {
                double v = Dm[c * q + r];
                for (int t = 0; t < c; ++t) v -= U[t * q + c] * U[t * q + r];
                U[c * q + r] = v / U[c * q + c];
            }
        }
        std::vector<std::vector<double>> dU(nD, std::vector<double>(q * q, 0.0)), Dt(nD, std::vector<double>(q * q, 0.0));
        std::vector<bool> is_diag(nD, false);
        if (diag_D) {
            for (int t = 0; t < q; ++t) { Dt[t][t * q + t] = Dm[t * q + t]; is_diag[t] = true; }
        } else {
            int t = 0;
            for (int c = 0; c < q; ++c)
                for (int r = 0; r <= c; ++r) {
                    dU[t][r * q + c] = (r == c) ? U[c * q + c] : 1.0;
                    is_diag[t] = (r == c);
                    // D_t = dU' U + U' dU
                    for (int a = 0; a < q; ++a)
                        for (int b = 0; b < q; ++b) {
                            double s = 0.0;
                            for (int k = 0; k < q; ++k) s += dU[t][k * q + a] * U[k * q + b] + U[k * q + a] * dU[t][k * q + b];
                            Dt[t][a * q + b] = s;
                        }
                    ++t;
                }
        }
        ia.nD = nD;
        double T1[GLMMA_QMAX * GLMMA_QMAX], T2[GLMMA_QMAX * GLMMA_QMAX], Gt[GLMMA_INFO_NDMAX][GLMMA_QMAX * GLMMA_QMAX];
        for (int t = 0; t < nD; ++t) {
            mm(Di, Dt[t].data(), q, T1);                 // D^-1 D_t
            double tr = 0.0;
            for (int a = 0; a < q; ++a) tr += T1[a * q + a];
            ia.dc[t] = -0.5 * tr;
            for (int i = 0; i < q * q; ++i) Gt[t][i] = T1[i];
            mm(T1, Di, q, T2);                           // D^-1 D_t D^-1
            for (int i = 0; i < q * q; ++i) ia.dM[t][i] = T2[i];
        }
        int tu = 0;
        for (int t = 0; t < nD; ++t)
            for (int u = t; u < nD; ++u) {
                double Dtu[GLMMA_QMAX * GLMMA_QMAX] = {0};
                if (diag_D) {
                    if (t == u) Dtu[t * q + t] = Dm[t * q + t];
                } else {
                    for (int a = 0; a < q; ++a)
                        for (int b = 0; b < q; ++b) {
                            double s = 0.0;
                            for (int k = 0; k < q; ++k) s += dU[t][k * q + a] * dU[u][k * q + b] + dU[u][k * q + a] * dU[t][k * q + b];
                            Dtu[a * q + b] = s;
                        }
                    if (t == u && is_diag[t])            // d2U/dtheta_t^2 = dU_t for a logged diagonal entry
                        for (int i = 0; i < q * q; ++i) Dtu[i] += Dt[t][i];
                }
                double GuGt[GLMMA_QMAX * GLMMA_QMAX], GtGu[GLMMA_QMAX * GLMMA_QMAX], Gtu[GLMMA_QMAX * GLMMA_QMAX];
                mm(Gt[u], Gt[t], q, GuGt);               // D^-1 D_u D^-1 D_t
                mm(Gt[t], Gt[u], q, GtGu);
                mm(Di, Dtu, q, Gtu);                     // D^-1 D_tu
                double tr = 0.0;
                for (int a = 0; a < q; ++a) tr += -GuGt[a * q + a] + Gtu[a * q + a];
                ia.dcc[tu] = -0.5 * tr;
                double A1[GLMMA_QMAX * GLMMA_QMAX], A2[GLMMA_QMAX * GLMMA_QMAX], A3[GLMMA_QMAX * GLMMA_QMAX];
                mm(GuGt, Di, q, A1);
                mm(Gtu, Di, q, A2);
                mm(GtGu, Di, q, A3);
                for (int i = 0; i < q * q; ++i) ia.dMM[tu][i] = -A1[i] + A2[i] - A3[i];
                ++tu;
            }
    }
    ia.d = h->d; ia.grid = h->d_grid; ia.fmtab = h->d_fmtab; fastmath_constants(ia.fmc);
    ia.ksm = h->K <= 1024 ? h->K : 0;        // 4 warps x K doubles of the kernel's (opt-in) dynamic shared memory
    const int grid = (int)std::min<int64_t>((h->d.n + 3) / 4, 148 * 8);
    double* d_part = nullptr;
    const int pm = P <= 12 ? 12 : GLMMA_INFO_PMAX;           // capacity of the instantiation that runs
    const int npair = pm * (pm + 1) / 2;
    cudaError_t e = cudaMalloc(&d_part, sizeof(double) * (size_t)grid * npair);
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaMalloc (observed information)");
    ia.partials = d_part;
    h->ops->info(ia, grid, pm, h->stream);
    h->launches++;
    std::vector<double> part((size_t)grid * npair);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(part.data(), d_part, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_part);
    if (e != cudaSuccess) return cuda_fail(h, e, "info_kernel");
    for (int r = 0; r < P; ++r)
        for (int c = r; c < P; ++c) {
            const int t = info_pair_pm(r, c, pm);
            double s = 0.0;
            for (int b = 0; b < grid; ++b) s += part[(size_t)b * npair + t];    // fixed order
            H[r + c * P] = -s;             // Hessian of the NEGATIVE log-likelihood
            H[c + r * P] = -s;
        }
    return GLMMA_OK;
}

// ---- predict()'s Metropolis step (R/methods.R:1013-1142) -----------------------------------

int glmma_log_post_b(glmma_handle* h, const double* b, const double* betas, const double* invD,
                     const double* phis, const double* gammas, double* out) {
    DeviceGuard device_guard_;
    NvtxRange nvtx_range_("glmma_log_post_b");
    if (h && !h->parts.empty()) {
        if (!b || !out) return fail(h, GLMMA_ERR_ARG, "null argument");
        const int64_t n = h->part_c0.back();
        for (size_t k = 0; k < h->parts.size(); ++k) {
            glmma_handle* p = h->parts[k];
            const int64_t c0 = h->part_c0[k], m = p->d.n;
            std::vector<double> tmp((size_t)m * h->q);
            for (int dd = 0; dd < h->q; ++dd) std::memcpy(tmp.data() + (size_t)dd * m, b + (size_t)dd * n + c0, sizeof(double) * m);
            int rc = glmma_log_post_b(p, tmp.data(), betas, invD, phis, gammas, out + c0);
            if (rc) return multi_fail(h, p, rc);
        }
        return GLMMA_OK;
    }
    int rc = enter(h);
    if (rc) return rc;
    if (!b || !out || !invD) return fail(h, GLMMA_ERR_ARG, "null argument");
    FamPar fp;
    rc = run_prep(h, betas, phis, gammas, &fp);
    if (rc) return rc;
    ModeArgs ma;
    ma.d = h->d; ma.fam = fp; ma.fmtab = h->d_fmtab; fastmath_constants(ma.fmc); ma.warm = 0; ma.maxit = 0; ma.tol = 0.0; ma.stats = h->d_stats;
    fill_prior(h, invD, true, &ma.prior);
    const int64_t n = h->d.n;
    double *d_b = nullptr, *d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_b, sizeof(double) * n * h->q);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, sizeof(double) * n);
    if (e != cudaSuccess) { cudaFree(d_b); return cuda_fail(h, e, "cudaMalloc (log_post_b)"); }
    e = cudaMemcpyAsync(d_b, b, sizeof(double) * n * h->q, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        h->ops->logpost(ma, d_b, d_out, h->stream);
        h->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_b); cudaFree(d_out);
    if (e != cudaSuccess) return cuda_fail(h, e, "log_post_b");
    return GLMMA_OK;
}

// ---- one process, several GPUs (R is single-process): glmma_create_multi ------------------------
// Clusters are conditionally independent, so the log-likelihood and every score component are plain
// sums over shards (R/Fit_Funs.R:33-37,105,194,231,244-247,271-278).  Each device holds a contiguous
// range of clusters behind its own sub-handle; an evaluation is enqueued on every device (kernels and
// the copy of the ~100-byte result into pinned host memory are asynchronous), then the host waits for
// each device in turn and adds the vectors in device order (fixed order: deterministic).  No
// collective, no peer access.

static int multi_fail(glmma_handle* h, const glmma_handle* p, int rc) {
    h->error = "device " + std::to_string(p->device) + ": " + p->error;
    return rc;
}

int glmma_create_multi(const glmma_data* data, const int* devices, int ndev, glmma_handle** out) {
    DeviceGuard device_guard_;
    NvtxRange nvtx_range_("glmma_create_multi");
    if (!data || !out || !devices || ndev < 1) return fail(nullptr, GLMMA_ERR_ARG, "null argument");
    *out = nullptr;
    if (ndev == 1) return glmma_create(data, devices[0], out);
    const int64_t N = data->n_obs;
    if (N <= 0 || data->p <= 0 || data->q_y <= 0 || !data->y || !data->X || !data->Z)
        return fail(nullptr, GLMMA_ERR_ARG, "y, X, Z and positive sizes are required");
    // scalar arguments are checked before any slice of the inputs is read
    if (N >= (int64_t)1 << 31) return fail(nullptr, GLMMA_ERR_ARG, "n_obs must be below 2^31");
    if (data->p > GLMMA_PMAX || data->p_zi > GLMMA_PZIMAX) return fail(nullptr, GLMMA_ERR_ARG, "too many fixed-effect columns");
    if (data->nAGQ < 1 || data->nAGQ > GLMMA_NAGQ_MAX) return fail(nullptr, GLMMA_ERR_ARG, "nAGQ out of range");
    if (data->family < 0 || data->family >= GLMMA_N_FAMILIES) return fail(nullptr, GLMMA_ERR_ARG, "unknown family");
    if (data->y_cols != 1 && data->y_cols != 2) return fail(nullptr, GLMMA_ERR_ARG, "y_cols must be 1 or 2");
    if ((data->q_zi > 0 && !data->Z_zi) || (data->p_zi > 0 && !data->X_zi))
        return fail(nullptr, GLMMA_ERR_ARG, "X_zi / Z_zi missing");
    if (!lookup_ops(data->family, data->link, data->q_y, data->q_zi))
        return fail(nullptr, GLMMA_ERR_ARG, "family / link / random-effects dimensions not supported by the engine");
    // global CSR index, validated before it is used to slice y / X / Z
    std::vector<int64_t> rp;
    if (data->row_ptr) {
        if (data->n_clusters <= 0) return fail(nullptr, GLMMA_ERR_ARG, "n_clusters must be positive");
        if (data->row_ptr[0] != 0 || data->row_ptr[data->n_clusters] != N)
            return fail(nullptr, GLMMA_ERR_ARG, "row_ptr must start at 0 and end at n_obs");
        for (int64_t i = 0; i < data->n_clusters; ++i)
            if (data->row_ptr[i + 1] <= data->row_ptr[i]) return fail(nullptr, GLMMA_ERR_ARG, "row_ptr must be strictly increasing");
        rp.assign(data->row_ptr, data->row_ptr + data->n_clusters + 1);
    } else if (data->id) {
        rp.push_back(0);
        for (int64_t r = 1; r < N; ++r)
            if (data->id[r] != data->id[r - 1]) rp.push_back(r);
        rp.push_back(N);
    } else {
        return fail(nullptr, GLMMA_ERR_ARG, "id or row_ptr is required");
    }
    const int64_t n = (int64_t)rp.size() - 1;
    if (n < ndev) return fail(nullptr, GLMMA_ERR_ARG, "fewer clusters than devices");
    // contiguous cluster ranges balanced on observations
    std::vector<int64_t> cut(ndev + 1, 0);
    cut[ndev] = n;
    for (int k = 1; k < ndev; ++k) {
        const int64_t target = (N * k) / ndev;
        int64_t c = std::lower_bound(rp.begin(), rp.end(), target) - rp.begin();
        c = std::max<int64_t>(c, cut[k - 1] + 1);
        c = std::min<int64_t>(c, n - (ndev - k));
        cut[k] = c;
    }
    glmma_handle* h = new glmma_handle();
    h->sticky = 0; h->launches = 0; h->ops = nullptr; h->stream = nullptr; h->device = devices[0];
    for (auto& e : h->ev) e = nullptr;
    std::memset(&h->d, 0, sizeof(h->d));
    auto slice = [&](const double* src, int ncol, int64_t r0, int64_t r1, std::vector<double>& buf) -> const double* {
        if (!src || ncol <= 0) return nullptr;
        const int64_t m = r1 - r0;
        buf.resize((size_t)m * ncol);
        for (int c = 0; c < ncol; ++c) std::memcpy(buf.data() + (size_t)c * m, src + (size_t)c * N + r0, sizeof(double) * m);
        return buf.data();
    };
    for (int k = 0; k < ndev; ++k) {
        const int64_t c0 = cut[k], c1 = cut[k + 1], r0 = rp[c0], r1 = rp[c1];
        std::vector<double> by, bX, bZ, bXz, bZz, bo, boz;
        std::vector<int64_t> prp((size_t)(c1 - c0 + 1));
        for (int64_t i = c0; i <= c1; ++i) prp[i - c0] = rp[i] - r0;
        glmma_data dp = *data;
        dp.n_obs = r1 - r0; dp.n_clusters = c1 - c0; dp.id = nullptr; dp.row_ptr = prp.data();
        dp.y = slice(data->y, data->y_cols, r0, r1, by);
        dp.X = slice(data->X, data->p, r0, r1, bX);
        dp.Z = slice(data->Z, data->q_y, r0, r1, bZ);
        dp.X_zi = slice(data->X_zi, data->p_zi, r0, r1, bXz);
        dp.Z_zi = slice(data->Z_zi, data->q_zi, r0, r1, bZz);
        dp.offset = slice(data->offset, 1, r0, r1, bo);
        dp.offset_zi = slice(data->offset_zi, 1, r0, r1, boz);
        dp.weights = data->weights ? data->weights + c0 : nullptr;
        glmma_handle* part = nullptr;
        int rc = glmma_create(&dp, devices[k], &part);
        if (rc == GLMMA_OK) {
            cudaSetDevice(devices[k]);
            if (cudaMallocHost(&part->h_pinned, sizeof(double) * glmma_result_len(part)) != cudaSuccess ||
                cudaMallocHost(&part->h_stats, 4 * sizeof(unsigned long long)) != cudaSuccess) {
                g_create_error = "cudaMallocHost failed";
                glmma_destroy(part);
                rc = GLMMA_ERR_CUDA;
            }
        }
        if (rc != GLMMA_OK) {
            const std::string msg = "device " + std::to_string(devices[k]) + ": " + g_create_error;
            glmma_destroy(h);
            g_create_error = msg;
            return rc;
        }
        h->parts.push_back(part);
        h->part_c0.push_back(c0);
        h->part_r0.push_back(r0);
    }
    h->part_c0.push_back(n);
    h->part_r0.push_back(N);
    const glmma_handle* p0 = h->parts[0];
    h->family = p0->family; h->link = p0->link; h->nagq = p0->nagq; h->q = p0->q; h->K = p0->K; h->y_cols = p0->y_cols;
    h->multi_p = data->p; h->multi_pzi = data->p_zi; h->multi_nphis = data->n_phis;
    h->have_modes = false; h->have_pby = false;
    for (size_t k = 0; k < h->parts.size(); ++k) {
        h->workers.emplace_back(new PartWorker());
        PartWorker* w = h->workers.back().get();
        w->th = std::thread([w] { w->loop(); });
    }
    *out = h;
    return GLMMA_OK;
}

// runs fn(part) for every part on the part's worker thread; returns the first non-zero status
static int run_parts(glmma_handle* h, const std::function<int(glmma_handle*)>& fn) {
    for (size_t k = 0; k < h->parts.size(); ++k) {
        glmma_handle* p = h->parts[k];
        h->workers[k]->submit([&fn, p] { return fn(p); });
    }
    int first = GLMMA_OK;
    for (size_t k = 0; k < h->parts.size(); ++k) {
        const int rc = h->workers[k]->wait();
        if (rc != GLMMA_OK && first == GLMMA_OK) first = multi_fail(h, h->parts[k], rc);
    }
    return first;
}

static int multi_find_modes(glmma_handle* h, const double* betas, const double* invD, const double* phis,
                            const double* gammas, int warm, glmma_mode_stats* stats) {
    int rc = run_parts(h, [&](glmma_handle* p) -> int {
        int r = find_modes_enqueue(p, betas, invD, phis, gammas, warm, p->h_stats);
        if (r) return r;
        cudaError_t e = cudaStreamSynchronize(p->stream);
        if (e != cudaSuccess) return cuda_fail(p, e, "find_modes");
        return GLMMA_OK;
    });
    if (rc) return rc;
    unsigned long long nc = 0, cf = 0, it = 0, mx = 0;
    for (glmma_handle* p : h->parts) {
        nc += p->h_stats[0]; cf += p->h_stats[1]; it += p->h_stats[2]; mx = std::max(mx, p->h_stats[3]);
    }
    h->have_modes = true;
    if (stats) {
        stats->n_not_converged = (int64_t)nc; stats->n_chol_failed = (int64_t)cf;
        stats->max_iterations = (int32_t)mx; stats->reserved = 0;
        stats->mean_iterations = (double)it / (double)h->part_c0.back();
    }
    return GLMMA_OK;
}

static int multi_get_modes(glmma_handle* h, double* post_modes, double* chol_upper, double* log_dets) {
    const int64_t n = h->part_c0.back();
    const int q = h->q, np = q * (q + 1) / 2;
    for (size_t k = 0; k < h->parts.size(); ++k) {
        glmma_handle* p = h->parts[k];
        const int64_t c0 = h->part_c0[k], m = p->d.n;
        std::vector<double> tmp(post_modes ? (size_t)m * q : 0);
        int rc = glmma_get_modes(p, post_modes ? tmp.data() : nullptr, chol_upper ? chol_upper + (size_t)c0 * np : nullptr,
                                 log_dets ? log_dets + c0 : nullptr);
        if (rc) return multi_fail(h, p, rc);
        if (post_modes)
            for (int dd = 0; dd < q; ++dd) std::memcpy(post_modes + (size_t)dd * n + c0, tmp.data() + (size_t)dd * m, sizeof(double) * m);
    }
    return GLMMA_OK;
}

static int multi_set_modes(glmma_handle* h, const double* post_modes, const double* chol_upper) {
    if (!post_modes || !chol_upper) return fail(h, GLMMA_ERR_ARG, "null argument");
    const int64_t n = h->part_c0.back();
    const int q = h->q, np = q * (q + 1) / 2;
    for (size_t k = 0; k < h->parts.size(); ++k) {
        glmma_handle* p = h->parts[k];
        const int64_t c0 = h->part_c0[k], m = p->d.n;
        std::vector<double> tmp((size_t)m * q);
        for (int dd = 0; dd < q; ++dd) std::memcpy(tmp.data() + (size_t)dd * m, post_modes + (size_t)dd * n + c0, sizeof(double) * m);
        int rc = glmma_set_modes(p, tmp.data(), chol_upper + (size_t)c0 * np);
        if (rc) return multi_fail(h, p, rc);
    }
    h->have_modes = true;
    h->have_pby = false;
    return GLMMA_OK;
}

static int multi_eval(glmma_handle* h, const double* betas, const double* D, const double* phis, const double* gammas,
                      int want_score, double* result, int em_mode) {
    if (!result) return fail(h, GLMMA_ERR_ARG, "result is null");
    const int rl = glmma_result_len(h), len = want_score ? rl : 4;
    double eye[GLMMA_QMAX * GLMMA_QMAX] = {0};
    for (int i = 0; i < h->q; ++i) eye[i * h->q + i] = 1.0;
    int rc = run_parts(h, [&](glmma_handle* p) -> int {
        int r = enter(p);
        if (r) return r;
        if (em_mode == 1 && !p->d_pby) {
            r = dev_alloc(p, &p->d_pby, (size_t)p->d.n * (size_t)p->K);
            if (r) return r;
        }
        if (em_mode == 2 && !p->have_pby) return fail(p, GLMMA_ERR_STATE, "no posterior weights yet: call glmma_em_estep first");
        r = enqueue_eval(p, betas, em_mode == 2 ? eye : D, phis, gammas, want_score, p->d_result, nullptr, nullptr, nullptr, em_mode);
        if (r) return r;
        cudaError_t e = cudaMemcpyAsync(p->h_pinned, p->d_result, sizeof(double) * len, cudaMemcpyDeviceToHost, p->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
        if (e != cudaSuccess) return cuda_fail(p, e, "eval");
        if (em_mode == 1) p->have_pby = true;
        return GLMMA_OK;
    });
    if (rc) return rc;
    // only the entries glmma.h promises for this call are touched (4 without the score): a caller may
    // pass a 4-double buffer when want_score = 0, exactly as with a single-device handle
    for (int t = 0; t < len; ++t) result[t] = 0.0;
    for (glmma_handle* p : h->parts)          // device order: deterministic
        for (int t = 0; t < len; ++t) result[t] += p->h_pinned[t];
    result[3] = 0.0;
    return GLMMA_OK;
}

static int multi_eval_contrib(glmma_handle* h, const double* betas, const double* D, const double* phis,
                              const double* gammas, double* ll_i, double* zbar, double* zbar_zi, double* sphi_obs,
                              double* S_i) {
    const int q = h->q;
    for (size_t k = 0; k < h->parts.size(); ++k) {
        glmma_handle* p = h->parts[k];
        const int64_t c0 = h->part_c0[k], r0 = h->part_r0[k];
        int rc = glmma_eval_contrib(p, betas, D, phis, gammas, ll_i ? ll_i + c0 : nullptr, zbar ? zbar + r0 : nullptr,
                                    zbar_zi ? zbar_zi + r0 : nullptr, sphi_obs ? sphi_obs + r0 : nullptr,
                                    S_i ? S_i + (size_t)c0 * q * q : nullptr);
        if (rc) return multi_fail(h, p, rc);
    }
    return GLMMA_OK;
}

static int multi_glm_pass(glmma_handle* h, int design, int glm_family, int response, int use_offset, int first,
                          const double* beta, double* out) {
    if (!out) return fail(h, GLMMA_ERR_ARG, "null argument");
    const int p_ = design == 0 ? h->multi_p : h->multi_pzi;
    const int len = 1 + p_ + p_ * p_;
    if (p_ < 1) return fail(h, GLMMA_ERR_ARG, "design has no columns");
    std::vector<double> tmp(len);
    for (int t = 0; t < len; ++t) out[t] = 0.0;
    for (glmma_handle* p : h->parts) {
        int rc = glmma_glm_pass(p, design, glm_family, response, use_offset, first, beta, tmp.data());
        if (rc) return multi_fail(h, p, rc);
        for (int t = 0; t < len; ++t) out[t] += tmp[t];
    }
    return GLMMA_OK;
}

// ---- cluster-sharded jobs, one process per GPU: peer-memory all-reduce inside the engine ----------

int glmma_comm_export(glmma_handle* h, void* handle_bytes) {
    DeviceGuard device_guard_;
    if (h && !h->parts.empty()) return fail(h, GLMMA_ERR_ARG, "a multi-device handle reduces on the host; no communicator needed");
    int rc = enter(h);
    if (rc) return rc;
    if (!handle_bytes) return fail(h, GLMMA_ERR_ARG, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == GLMMA_COMM_HANDLE_BYTES, "GLMMA_COMM_HANDLE_BYTES");
    if (!h->d_xchg) {
        const size_t bytes = sizeof(double) * 2 * GLMMA_MAX_RANKS * GLMMA_XSLOT;
        CU(h, cudaMalloc(&h->d_xchg, bytes));          // its own allocation: IPC handles name whole allocations
        CU(h, cudaMemset(h->d_xchg, 0, bytes));
    }
    cudaIpcMemHandle_t ih;
    CU(h, cudaIpcGetMemHandle(&ih, h->d_xchg));
    std::memcpy(handle_bytes, &ih, sizeof(ih));
    return GLMMA_OK;
}

int glmma_comm_connect(glmma_handle* h, int rank, int nranks, const void* all_handles) {
    DeviceGuard device_guard_;
    if (h && !h->parts.empty()) return fail(h, GLMMA_ERR_ARG, "a multi-device handle reduces on the host; no communicator needed");
    int rc = enter(h);
    if (rc) return rc;
    if (!all_handles || nranks < 1 || nranks > GLMMA_MAX_RANKS || rank < 0 || rank >= nranks)
        return fail(h, GLMMA_ERR_ARG, "bad rank / nranks (at most 16 ranks)");
    if (!h->d_xchg) return fail(h, GLMMA_ERR_STATE, "call glmma_comm_export first");
    if (glmma_result_len(h) > GLMMA_XSLOT - 1) return fail(h, GLMMA_ERR_ARG, "result vector too long for the exchange slot");
    if (h->xchg.nranks > 1) return fail(h, GLMMA_ERR_STATE, "already connected");
    XchgPar x;
    std::memset(&x, 0, sizeof(x));
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) { x.peer[r] = h->d_xchg; continue; }
        cudaIpcMemHandle_t ih;
        std::memcpy(&ih, (const char*)all_handles + (size_t)r * GLMMA_COMM_HANDLE_BYTES, sizeof(ih));
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (void* q : h->ipc_opened) cudaIpcCloseMemHandle(q);
            h->ipc_opened.clear();
            return cuda_fail(h, e, "cudaIpcOpenMemHandle (peer exchange buffer)");
        }
        h->ipc_opened.push_back(p);
        x.peer[r] = (double*)p;
    }
    x.rank = rank; x.nranks = nranks; x.seq = 0;
    h->xchg = x;
    return GLMMA_OK;
}

int glmma_comm_size(const glmma_handle* h) { return !h ? 0 : (h->xchg.nranks > 1 ? h->xchg.nranks : 1); }

int glmma_time_eval(glmma_handle* h, const double* betas, const double* D, const double* phis,
                    const double* gammas, int want_score, int iters, float* ms_per_eval, float* ms_main_kernel) {
    DeviceGuard device_guard_;
    if (h && !h->parts.empty()) return fail(h, GLMMA_ERR_ARG, "glmma_time_eval needs a single-device handle");
    int rc = enter(h);
    if (rc) return rc;
    if (iters < 1) return fail(h, GLMMA_ERR_ARG, "iters must be positive");
    std::vector<cudaEvent_t> k0(iters), k1(iters);
    for (int i = 0; i < iters; ++i) { cudaEventCreate(&k0[i]); cudaEventCreate(&k1[i]); }
    CU(h, cudaStreamSynchronize(h->stream));
    CU(h, cudaEventRecord(h->ev[0], h->stream));
    for (int i = 0; i < iters && rc == GLMMA_OK; ++i)
        rc = enqueue_eval(h, betas, D, phis, gammas, want_score, h->d_result, nullptr, k0[i], k1[i]);
    if (rc == GLMMA_OK) {
        cudaEventRecord(h->ev[1], h->stream);
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = cuda_fail(h, e, "time_eval");
    }
    if (rc == GLMMA_OK) {
        float tot = 0.f, ker = 0.f;
        cudaEventElapsedTime(&tot, h->ev[0], h->ev[1]);
        for (int i = 0; i < iters; ++i) { float t = 0.f; cudaEventElapsedTime(&t, k0[i], k1[i]); ker += t; }
        if (ms_per_eval) *ms_per_eval = tot / iters;
        if (ms_main_kernel) *ms_main_kernel = ker / iters;
    }
    for (int i = 0; i < iters; ++i) { cudaEventDestroy(k0[i]); cudaEventDestroy(k1[i]); }
    return rc;
}

int glmma_measure_fp64_peak(int device, double* tflops) {
    DeviceGuard device_guard_;
    if (!tflops) return GLMMA_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        g_create_error = "no usable CUDA device";
        return GLMMA_ERR_CUDA;
    }
    cudaSetDevice(device);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    double* out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return GLMMA_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    fp64_peak_kernel<<<blocks, threads>>>(out, 64);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fp64_peak_kernel<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    if (cudaGetLastError() != cudaSuccess) return GLMMA_ERR_CUDA;
    const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
    *tflops = flops / (best * 1e-3) / 1e12;
    return GLMMA_OK;
}

}  // extern "C"
