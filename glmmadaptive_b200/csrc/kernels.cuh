// Kernels of the AGQ hot path (sm_100a, fp64).  One warp owns one cluster; the
// 32 lanes own 32 quadrature nodes at a time ("a round") and walk the cluster's
// observations together, so every observation record is a shared-memory
// broadcast and the log-sum-exp over nodes / score accumulation are warp
// reductions.  Nothing of size N x K is ever materialised (the reference builds
// Ztb, eta_y, log_Lik, p_by ... as N x K matrices, R/Functions.R:128-134,
// R/Fit_Funs.R:21-33,63-90).
#pragma once
#include <type_traits>
#include "families.cuh"

#define GLMMA_EVAL_THREADS 128
#define GLMMA_EVAL_MAX_THREADS 384     // generic kernel: 128, 256 or 384 threads per block (170 registers)
#define GLMMA_JT_CAP 32
#define GLMMA_ASUM_MAX_K 1024          // eval_kernel: largest grid whose node sums are kept per warp (tile-outer path)

struct EvalArgs {
    DevData d;
    PriorPar prior;
    FamPar fam;
    const GridPar* grid;     // device copy of the Gauss-Hermite table
    const double* fmtab;     // device copy of the fastmath tables (log table, then exp table)
    double fmc[GLMMA_FM_NCONST];   // fastmath constants (constant bank)
    double* partials;        // gridDim.x * GLMMA_NACC block partial sums
    int jt;                  // observation tile held in shared memory per warp (<= GLMMA_JT_CAP)
    int only_deferred;       // eval_kernel: process only clusters flagged in d.defer
    int asum_doubles;        // eval_kernel: per-warp node sums (K doubles, 0 = none) for clusters larger than jt
    // EM phase of mixed_fit (R/mixed_fit.R:119-229), eval_kernel only:
    //   em_mode 1 (E-step): also store the un-normalised posterior node weights
    //             e_ik = exp(a_ik - shift_i) in pby (n x K; all zero for a dropped cluster);
    //   em_mode 2 (M-step): take e_ik from pby instead of from the current theta, i.e. evaluate
    //             the scores with p_by held fixed (score_betas / score_phis / score_gammas,
    //             R/Fit_Funs.R:295-427); the log-likelihood slot of the result is 0.
    int em_mode;
    double* pby;
    // eval_fast_kernel, dispatch on cluster size (ragged data): the launch walks the n_order clusters
    // listed in `order` (one size class, built once by glmma_create; the identity for data whose
    // clusters all fit one tile); the other classes belong to the launch with the other tile / to
    // eval_kernel.
    const int4* order;       // {cluster, first row, n_i, 0} per listed cluster: one 16-byte load
    int64_t n_order;
    // Fused mode (engine.cu, enqueue_eval): the register variant forms X beta (+ offset), X_zi gamma and
    // the node-independent terms itself from the staged design columns, and accumulates X'zbar /
    // X_zi'zbar_zi per warp -- no prep_cluster / score_reduce launches, no xb / zbar round trips
    // through HBM; the last block of eval_kernel (always the last launch of an evaluation) turns the
    // block partials into the result vector -- no finalize launch.
    int fused;
    int xw;                  // staged design columns per observation: p + p_zi + [offset] + [offset_zi]
    int zbar_out;            // fused mode: also write zbar / zbar_zi (per-observation contributions wanted)
    const double* ytab;      // 2 x GLMMA_YTAB table of obs_const(y) (count families) or null
    CoefPar coef;            // fixed effects (fused mode)
    unsigned* fin_counter;   // eval_kernel: blocks-done counter; null = separate finalize launch
    // clusters the register variant left to eval_kernel in THIS evaluation (counted where d.defer is raised); with
    // none -- the usual case -- eval_kernel's blocks skip their table set-up and the scan of the flags and go
    // straight to the finalisation.  defer_static: clusters no tile takes exist (always scan).  null: always scan
    unsigned* defer_count;
    int defer_static;
    const double* fin_partials;   // first block partial of this evaluation (all launches)
    int fin_blocks;          // block partials written by this evaluation's launches in total
    int fin_want_score;
    double* fin_result;      // result vector (include/glmma.h, glmma_eval)
    const double* fin_cpart; // const_sums_kernel block partials {sum w c_ll, sum w c_phi} (fused mode)
    int fin_ncpart;
    double fin_ll_const;     // beta' X'y1 + sum w y1 offset over the register variant's clusters (linear-eta hoist)
    XchgPar xchg;            // peer-memory all-reduce of the result vector over the ranks of a sharded job
    HostOut hout;            // result vector delivered straight into pinned host memory (glmma_eval)
};

struct PrepArgs {
    DevData d;
    FamPar fam;
    int want_score;
    double* ytab;            // 2 x GLMMA_YTAB table of obs_const(y) (count families) or null
    const int32_t* list;     // prep_cluster_kernel: only these clusters (null: all)
    int64_t n_list;
};

struct ModeArgs {
    DevData d;
    PriorPar prior;          // Dinv = invD as handed to find_modes
    FamPar fam;
    const double* fmtab;
    double fmc[GLMMA_FM_NCONST];
    int warm;
    int maxit;
    double tol;
    unsigned long long* stats;   // [0] not converged, [1] chol failed, [2] sum iterations, [3] max iterations
};

template <int Q> struct Tri {
    __host__ __device__ static constexpr int rm(int d, int e) { return d * Q - d * (d - 1) / 2 + (e - d); }   // row-major d<=e
    __host__ __device__ static constexpr int cm(int r, int c) { return c * (c + 1) / 2 + r; }               // R upper.tri order r<=c
};

// 8-byte asynchronous global -> shared copy (LDGSTS)
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc) : "memory");
}
// 16-byte variant (source and destination 16-byte aligned)
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// replicated fastmath tables (see fastmath.cuh): src is the compact global copy
// [128 x (invc, logc) | 64 x 2^(j/64)]
__device__ __forceinline__ void load_fmtab(double* fmtab, const double* src) {
    for (int t = threadIdx.x; t < GLMMA_FM_LOG_DOUBLES; t += blockDim.x) {
        const int entry = t >> 4, within = t & 1;
        fmtab[t] = src[entry * 2 + within];
    }
    for (int t = threadIdx.x; t < GLMMA_FM_EXP_DOUBLES; t += blockDim.x) fmtab[GLMMA_FM_LOG_DOUBLES + t] = src[256 + (t >> 4)];
    __syncthreads();
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Zero-copy delivery of the (reduced) result vector: the finalising block stores it into pinned,
// device-mapped host memory and then raises a flag the host thread is spinning on -- glmma_eval returns
// without a cudaMemcpy or a stream synchronisation (which cost more than the 100 bytes they move).
__device__ __forceinline__ void deliver_host(const HostOut& ho, const double* result, int len) {
    __syncthreads();
    for (int e = threadIdx.x; e < len; e += blockDim.x) ho.vec[e] = result[e];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(ho.flag) = ho.seq;
}

// All-reduce (sum) of the result vector over the ranks of a cluster-sharded job, done by the block that
// finalised this rank's evaluation -- compute and collective in ONE kernel, over peer memory (NVLink /
// NVSwitch P2P stores), no NCCL launch and no host round trip:
//   1. the block stores its vector into slot [parity][rank] of EVERY rank's exchange buffer (its own
//      included), fences at system scope and then raises the slot's flag to this evaluation's number;
//   2. thread r spins on the flag of slot [parity][r] in OUR buffer until rank r's vector has arrived;
//   3. every rank adds the vectors in rank order: the same sum on every rank, bit for bit.
// Ranks evaluate in lockstep (the host loops are identical), so a slot of parity p is not rewritten
// before all ranks have read evaluation seq - 2 from it.  A peer that never arrives (a crashed rank)
// poisons the result with NaN after ~2 s instead of hanging the GPU.
__device__ __forceinline__ void exchange_block(const XchgPar& x, double* result, int len) {
    __syncthreads();                           // the block's own writes of `result`
    const size_t par = (size_t)(x.seq & 1ull) * x.nranks;
    const size_t mine = (par + x.rank) * GLMMA_XSLOT;
    for (int t = threadIdx.x; t < len * x.nranks; t += blockDim.x) {
        const int r = t / len, e = t - r * len;
        x.peer[r][mine + e] = result[e];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < x.nranks)
        *reinterpret_cast<volatile unsigned long long*>(x.peer[threadIdx.x] + mine + GLMMA_XSLOT - 1) = x.seq;
    bool arrived = true;
    double* own = x.peer[x.rank];
    if ((int)threadIdx.x < x.nranks) {
        const volatile unsigned long long* f =
            reinterpret_cast<const volatile unsigned long long*>(own + (par + threadIdx.x) * GLMMA_XSLOT + GLMMA_XSLOT - 1);
        const long long t0 = clock64();
        while (*f != x.seq)
            if (clock64() - t0 > (4ll << 30)) { arrived = false; break; }
    }
    arrived = __syncthreads_and(arrived);
    __threadfence_system();
    for (int e = threadIdx.x; e < len; e += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < x.nranks; ++r) s += __ldcv(own + (par + r) * GLMMA_XSLOT + e);
        result[e] = arrived ? s : NAN;
    }
}

// Result vector (include/glmma.h, glmma_eval) from the block partials of one evaluation, by ONE block of
// any size: warp w reduces accumulator w (then w + nw, ...) -- lanes stride the blocks, a fixed shuffle
// tree finishes, so the sum order is fixed (deterministic) and nothing is serialised behind block
// barriers.  cols_in_ep: the score columns X'zbar, X_zi'zbar_zi travel in the block partials (fused
// mode); otherwise they come from score_reduce_kernel's partials.  sh: GLMMA_NACC + GLMMA_PMAX +
// GLMMA_PZIMAX + 2 doubles of shared memory.  The block must have at least two warps.
__device__ __forceinline__ void finalize_block(const double* ep, int n_eval_blocks, const double* sp, int n_score_blocks,
                                               bool cols_in_ep, int p, int n_phis, int p_zi, int q, int want_score,
                                               double* result, double* sh, const double* cpart = nullptr, int ncpart = 0,
                                               double ll_const = 0.0, bool ll_consts_on = true, const XchgPar* xchg = nullptr,
                                               const HostOut* hout = nullptr) {
    double* acc = sh;
    double* col = sh + GLMMA_NACC;
    const int np = q * (q + 1) / 2;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int ncol = p + p_zi;
    for (int a = w; a < 4 + np; a += nw) {
        double s = 0.0;
        for (int b = lane; b < n_eval_blocks; b += 32) s += __ldcg(ep + (int64_t)b * GLMMA_PSTRIDE + a);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) acc[a] = s;
    }
    if (want_score) {
        for (int l = w; l < ncol; l += nw) {
            double s = 0.0;
            if (cols_in_ep) {
                for (int b = lane; b < n_eval_blocks; b += 32) s += __ldcg(ep + (int64_t)b * GLMMA_PSTRIDE + GLMMA_NACC + l);
            } else {
                for (int b = lane; b < n_score_blocks; b += 32) s += sp[(int64_t)b * ncol + l];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) col[l] = s;
        }
    }
    // fused mode: node-independent terms of the register variant's clusters (const_sums_kernel partials:
    // warps 0 / 1 sum the two columns in a fixed order) and the linear-eta constant
    double* cadd = sh + GLMMA_NACC + GLMMA_PMAX + GLMMA_PZIMAX;
    if (w < 2) {
        double s = 0.0;
        for (int b = lane; b < ncpart; b += 32) s += __ldcg(cpart + 2 * b + w);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) cadd[w] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        result[0] = acc[0] + (ll_consts_on ? cadd[0] + ll_const : 0.0);
        result[1] = acc[1]; result[2] = acc[2]; result[3] = 0.0;
    }
    if (want_score) {
        for (int l = threadIdx.x; l < ncol; l += blockDim.x) {
            if (l < p) result[4 + l] = col[l];
            else result[4 + p + n_phis + (l - p)] = col[l];
        }
        if (threadIdx.x == 0) {
            for (int t = 0; t < n_phis; ++t) result[4 + p + t] = acc[3] + cadd[1];
            double* S = result + 4 + p + n_phis + p_zi;
            int t = 0;
            for (int dd = 0; dd < q; ++dd)
                for (int e = dd; e < q; ++e) {
                    S[dd + e * q] = acc[4 + t];
                    S[e + dd * q] = acc[4 + t];
                    ++t;
                }
        }
    }
    const int len = want_score ? 4 + p + n_phis + p_zi + q * q : 4;
    if (xchg != nullptr && xchg->nranks > 1) exchange_block(*xchg, result, len);
    if (hout != nullptr && hout->vec != nullptr) deliver_host(*hout, result, len);
}

// ------------------------------------------------------------------------------
// prep: per-observation theta-dependent quantities (X beta, X_zi gamma and the
// node-independent parts of log f and d log f / d phis).
// ------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(256) prep_kernel(PrepArgs A, CoefPar C) {
    const DevData& d = A.d;
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= d.N) return;
    double xb = d.offset ? d.offset[row] : 0.0;
    for (int l = 0; l < d.p; ++l) xb = fma(d.X[row + (int64_t)l * d.N], C.betas[l], xb);
    d.xb[row] = xb;
    if (F::HAS_ZI) {
        double xg = d.offset_zi ? d.offset_zi[row] : 0.0;
        for (int l = 0; l < d.p_zi; ++l) xg = fma(d.X_zi[row + (int64_t)l * d.N], C.gammas[l], xg);
        d.xg[row] = xg;
    }
    double cll, cphi;
    F::obs_const(d.y1[row], F::USES_Y2 ? d.y2[row] : 0.0, A.fam, cll, cphi);
    d.cll[row] = cll;
    if (F::HAS_PHI) d.cphi[row] = cphi;
}

// ------------------------------------------------------------------------------
// eval: fused logLik_mixed + score_mixed sufficient statistics
//
// Shared memory per block: [fastmath tables | GH nodes, log weights | per-warp areas].
// Per warp: rec  (jt x W)            observation records y1, y2, xb, (xg), z.., zzi..
//           tabF (jt x Q x nagq)     separable factors of exp(eta):   exp(eta_jk)    = prod_e tabF[j][e][a_e(k)]
//           tabL (jt x QZ x nagq)    separable factors of exp(eta_zi) (zero-part families)
//           scr_eta [, scr_zi] (jt x 32), ev (32)           per-lane score terms and node weights of a round
// The separable tables use that R^-1 is triangular: eta_jk = base_j + sum_e c_je x_{a_e},
// so exp(eta) over the K = nagq^Q nodes needs only Q * nagq exponentials per observation
// ("optional algorithmic shortcut").
// ------------------------------------------------------------------------------
#define GLMMA_FM_DOUBLES (GLMMA_FM_LOG_DOUBLES + GLMMA_FM_EXP_DOUBLES)   // replicated tables in shared memory

template <class F, int QY, int QZ>
struct EvalLayout {
    static constexpr int Q = QY + QZ;
    static constexpr int NP = Q * (Q + 1) / 2;
    static constexpr int OFF_XG = 3;
    static constexpr int OFF_Z = F::HAS_ZI ? 4 : 3;
    static constexpr int W0 = OFF_Z + QY + QZ;
    static constexpr int W = (W0 + 1) & ~1;                      // even: records stay 16-byte aligned
    static constexpr int NSCR = F::HAS_ZI ? 2 : 1;               // scr_eta [, scr_zi]
    static constexpr int NTJ = (GLMMA_JT_CAP + 31) / 32;         // observations per lane of a staged tile
    static constexpr int SCRW = 33;                              // row pitch of the score scratch: rows (one per
                                                                 // observation) are written by columns and summed by rows
    static constexpr int NTABL = F::HAS_ZI ? (QZ > 0 ? QZ : 1) : 0;
    __host__ __device__ static size_t warp_doubles(int jt, int nagq, bool score) {
        return (size_t)jt * W + (size_t)jt * (Q + NTABL) * nagq + (score ? (((size_t)NSCR * jt * SCRW + 1) & ~(size_t)1) + 32 : 0);   // even where it can be: the compiler pairs record loads
    }
    __host__ __device__ static size_t head_doubles() { return GLMMA_FM_DOUBLES + 2 * GLMMA_NAGQ_MAX; }
    __host__ static size_t smem_bytes(int jt, int nagq, bool score, int warps, int extra = 0) {
        return sizeof(double) * (head_doubles() + warps * (warp_doubles(jt, nagq, score) + (size_t)extra));
    }
};

// t / nagq for 0 <= t < 2^27 without an integer division: rcp = ceil(2^32 / nagq) errs by less than
// nagq <= 32 parts in 2^32 (nagq == 1, the Laplace grid, is flagged by rcp == 0)
__device__ __forceinline__ unsigned nagq_rcp(int nagq) {
    return nagq > 1 ? (unsigned)((0x100000000ull + (unsigned)nagq - 1u) / (unsigned)nagq) : 0u;
}
__device__ __forceinline__ int div_nagq(int t, unsigned rcp) { return rcp ? (int)__umulhi((unsigned)t, rcp) : t; }

// decode node k -> b_k = mode + sqrt(2) R^-1 z_k, returns log prior + log wGH; idx[e] = a_e(k)
template <int Q>
__device__ __forceinline__ double node_setup(int k, int nagq, unsigned rcp, const double* gx, const double* glw, double lw_const,
                                             const double* m, const double* ri, const PriorPar& pr, double* b, int* idx) {
    double z[Q];
    double lw = lw_const;
    int kk = k;
#pragma unroll
    for (int dd = 0; dd < Q; ++dd) {
        const int qd = div_nagq(kk, rcp);
        const int a = kk - qd * nagq;
        kk = qd;
        idx[dd] = a;
        z[dd] = gx[a];
        lw += glw[a];
    }
#pragma unroll
    for (int dd = 0; dd < Q; ++dd) {
        double s = 0.0;
#pragma unroll
        for (int e = dd; e < Q; ++e) s = fma(ri[Tri<Q>::rm(dd, e)], z[e], s);
        b[dd] = fma(GLMMA_SQRT2, s, m[dd]);
    }
    double quad = 0.0;
#pragma unroll
    for (int dd = 0; dd < Q; ++dd) {
        double s = 0.0;
#pragma unroll
        for (int e = 0; e < Q; ++e) s = fma(pr.Dinv[dd * Q + e], b[e], s);
        quad = fma(b[dd], s, quad);
    }
    return pr.logprior_const - 0.5 * quad + lw;
}

template <class F, int QY, int QZ>
__device__ __forceinline__ void stage_tile(const DevData& d, int64_t r0, int cnt, int lane, double* rec) {
    using L = EvalLayout<F, QY, QZ>;
    for (int j = lane; j < cnt; j += 32) {
        const int64_t row = r0 + j;
        double* r = rec + j * L::W;
        r[0] = d.y1[row];
        r[1] = F::USES_Y2 ? d.y2[row] : 0.0;
        r[2] = d.xb[row];
        if (F::HAS_ZI) r[L::OFF_XG] = d.xg[row];
#pragma unroll
        for (int dd = 0; dd < QY; ++dd) r[L::OFF_Z + dd] = d.Z[row + (int64_t)dd * d.N];
#pragma unroll
        for (int dd = 0; dd < QZ; ++dd) r[L::OFF_Z + QY + dd] = d.Z_zi[row + (int64_t)dd * d.N];
    }
}

// separable exp tables of one staged tile (lanes over (observation, dimension, node index)).
// Returns whether every factor's argument is within +-230, so that no product of up to three
// factors can under- or overflow (warp-uniform); the caller streams the cluster otherwise.
template <class F, int QY, int QZ>
__device__ __forceinline__ bool build_tables(const double* rec, int cnt, int lane, int nagq, unsigned rcp, const double* gx,
                                             const double* m, const double* ri, const MathCtx& cx,
                                             double* tabF, double* tabL) {
    using L = EvalLayout<F, QY, QZ>;
    constexpr int Q = L::Q;
    bool ok = true;
    if (F::NEED_EMU) {
        const int items = cnt * Q * nagq;
        for (int t = lane; t < items; t += 32) {
            const int je = div_nagq(t, rcp);
            const int a = t - je * nagq;
            const int e = je % Q;
            const int j = je / Q;
            const double* r = rec + j * L::W;
            // c_je = sqrt(2) sum_{d <= min(e, QY-1)} z_jd R^-1[d][e];  base_j = xb_j + z_j . mode
            double c = 0.0, base = r[2];
#pragma unroll
            for (int dd = 0; dd < QY; ++dd) {
                base = fma(r[L::OFF_Z + dd], m[dd], base);
#pragma unroll
                for (int ee = dd; ee < Q; ++ee)
                    if (ee == e) c = fma(r[L::OFF_Z + dd], ri[Tri<Q>::rm(dd, ee)], c);
            }
            const double arg = fma(GLMMA_SQRT2 * c, gx[a], e == 0 ? base : 0.0);
            if (!(fabs(arg) <= 230.0)) ok = false;
            tabF[t] = fast_exp(arg, cx);
        }
    }
    if (F::HAS_ZI) {
        constexpr int NT = L::NTABL > 0 ? L::NTABL : 1;
        const int items = cnt * NT * nagq;
        for (int t = lane; t < items; t += 32) {
            const int je = div_nagq(t, rcp);
            const int a = t - je * nagq;
            const int e = je % NT;            // table e <-> random-effect dimension QY + e
            const int j = je / NT;
            const double* r = rec + j * L::W;
            double c = 0.0, base = r[L::OFF_XG];
#pragma unroll
            for (int dd = 0; dd < QZ; ++dd) {
                base = fma(r[L::OFF_Z + QY + dd], m[QY + dd], base);
#pragma unroll
                for (int ee = dd; ee < QZ; ++ee)
                    if (ee == e) c = fma(r[L::OFF_Z + QY + dd], ri[Tri<Q>::rm(QY + dd, QY + ee)], c);
            }
            const double arg = fma(GLMMA_SQRT2 * c, QZ > 0 ? gx[a] : 0.0, e == 0 ? base : 0.0);
            if (!(fabs(arg) <= 230.0)) ok = false;
            tabL[t] = fast_exp(arg, cx);
        }
    }
    return __all_sync(0xffffffffu, ok);
}

// eta, eta_zi and (SEP: from the tables; else by fast_exp) exp(eta), exp(eta_zi) of one point
template <class F, int QY, int QZ, bool SEP>
__device__ __forceinline__ void lin_pred(const double* r, const double* b, const int* idx, int j, int nagq,
                                         const double* tabF, const double* tabL, const MathCtx& cx,
                                         double& eta, double& ez, double& emu, double& elam) {
    using L = EvalLayout<F, QY, QZ>;
    constexpr int Q = L::Q;
    eta = r[2];
#pragma unroll
    for (int dd = 0; dd < QY; ++dd) eta = fma(r[L::OFF_Z + dd], b[dd], eta);
    ez = 0.0; emu = 0.0; elam = 0.0;
    if (F::HAS_ZI) {
        ez = r[L::OFF_XG];
#pragma unroll
        for (int dd = 0; dd < QZ; ++dd) ez = fma(r[L::OFF_Z + QY + dd], b[QY + dd], ez);
    }
    if (F::NEED_EMU) {
        if (SEP) {
            const double* tf = tabF + j * Q * nagq;
            emu = tf[idx[0]];
#pragma unroll
            for (int e = 1; e < Q; ++e) emu *= tf[e * nagq + idx[e]];     // in range: build_tables
        } else {
            emu = fast_exp(eta, cx);
        }
    }
    if (F::HAS_ZI) {
        if (SEP) {
            constexpr int NT = L::NTABL > 0 ? L::NTABL : 1;
            const double* tl = tabL + j * NT * nagq;
            elam = tl[QZ > 0 ? idx[QY] : 0];
#pragma unroll
            for (int e = 1; e < QZ; ++e) elam *= tl[e * nagq + idx[QY + e]];
        } else {
            elam = fast_exp(fmin(fmax(ez, -690.0), 690.0), cx);
        }
    }
}

template <class F, int QY, int QZ, bool SCORE>
__global__ void __launch_bounds__(GLMMA_EVAL_MAX_THREADS, 1) eval_kernel(EvalArgs A) {
    using L = EvalLayout<F, QY, QZ>;
    constexpr int Q = L::Q;
    constexpr int NP = L::NP;
    const DevData& d = A.d;
    extern __shared__ double smem[];
    double* fmtab = smem;
    double* gx = smem + GLMMA_FM_DOUBLES;
    double* glw = gx + GLMMA_NAGQ_MAX;
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int jt = A.jt;
    const int nagq = A.grid->nagq;
    const int K = A.grid->K;
    const unsigned rcp_nagq = nagq_rcp(nagq);
    double* wbase = smem + L::head_doubles() + wib * (L::warp_doubles(jt, nagq, SCORE) + (A.asum_doubles & ~1));
    double* asum = wbase + L::warp_doubles(jt, nagq, SCORE);     // K node sums (clusters larger than jt)
    double* rec = wbase;
    double* tabF = rec + jt * L::W;
    double* tabL = tabF + jt * Q * nagq;
    double* scr_eta = tabL + jt * L::NTABL * nagq;
    double* scr_zi = scr_eta + jt * L::SCRW;
    double* ev = scr_eta + ((L::NSCR * jt * L::SCRW + 1) & ~1);

    // (grid-uniform: the counter is final before this launch starts and is reset only by the finalising block)
    const bool any_work = !(A.only_deferred && A.defer_count != nullptr && !A.defer_static && __ldcg(A.defer_count) == 0u);
    if (!any_work && A.fin_counter != nullptr) {
        // nothing left to this kernel: block 0 alone turns the register variant's block partials (complete: this
        // launch follows them on the stream) into the result vector, the other blocks leave at once
        if (blockIdx.x != 0) return;
        finalize_block(A.fin_partials, A.fin_blocks - (int)gridDim.x, nullptr, 0, true, d.p, d.n_phis, d.p_zi, Q, A.fin_want_score,
                       A.fin_result, smem, A.fin_cpart, A.fin_ncpart, A.fin_ll_const, A.em_mode != 2, &A.xchg, &A.hout);
        return;
    }
    if (any_work) {
        load_fmtab(fmtab, A.fmtab);
        if (threadIdx.x < GLMMA_NAGQ_MAX) {
            gx[threadIdx.x] = A.grid->x[threadIdx.x];
            glw[threadIdx.x] = A.grid->lw[threadIdx.x];
        }
    }
    __syncthreads();
    const MathCtx cx = make_mathctx(fmtab + 2 * (threadIdx.x & 7), fmtab + GLMMA_FM_LOG_DOUBLES + (threadIdx.x & 15), A.fmc);
    const double lw_const = A.grid->lw_const;
    const int nrounds = (K + 31) >> 5;
    const int rc = ((K - 1) >> 1) >> 5;         // the round holding the central node

    double tot[GLMMA_NACC];
#pragma unroll
    for (int a = 0; a < GLMMA_NACC; ++a) tot[a] = 0.0;
    double gcol = 0.0;      // fused mode: lane l sums column l of X'zbar | X_zi'zbar_zi over this warp's clusters

    const int64_t nwarps = (int64_t)gridDim.x * wpb;
    // only_deferred: every warp owns blocks of 32 consecutive clusters, reads their flags with
    // one coalesced load and walks the set bits (the register variant did the others)
    const int64_t span = A.only_deferred ? 32 : 1;
    const int64_t n_walk = any_work ? d.n : 0;
    for (int64_t ib = ((int64_t)blockIdx.x * wpb + wib) * span; ib < n_walk; ib += nwarps * span) {
      unsigned todo = 1u;
      if (A.only_deferred) {
          const int64_t ci = ib + lane;
          todo = __ballot_sync(0xffffffffu, ci < d.n && d.defer[ci] != 0);
      }
      while (todo) {
        const int bit = __ffs(todo) - 1;
        todo &= todo - 1;
        const int64_t i = ib + bit;
        const int64_t r0 = d.row_ptr[i];
        const int ni = (int)(d.row_ptr[i + 1] - r0);
        const double w = d.weights ? d.weights[i] : 1.0;
        double m[Q], ri[NP];
#pragma unroll
        for (int dd = 0; dd < Q; ++dd) m[dd] = d.modes[i * Q + dd];
#pragma unroll
        for (int t = 0; t < NP; ++t) ri[t] = d.rinv[i * NP + t];

        // node-independent parts of the cluster's log-likelihood / phi score
        double c_ll = 0.0, c_phi = 0.0;
        for (int j = lane; j < ni; j += 32) {
            c_ll += d.cll[r0 + j];
            if (SCORE && F::HAS_PHI) c_phi += d.cphi[r0 + j];
        }
        c_ll = warp_sum(c_ll);
        if (SCORE && F::HAS_PHI) c_phi = warp_sum(c_phi);

        bool fast = ni <= jt;
        __syncwarp();
        if (fast) {
            stage_tile<F, QY, QZ>(d, r0, ni, lane, rec);
            __syncwarp();
            // a factor that could under- or overflow: the streaming path forms exp(eta) per point
            fast = build_tables<F, QY, QZ>(rec, ni, lane, nagq, rcp_nagq, gx, m, ri, cx, tabF, tabL);
        }

        double shift = 0.0, sum_e = 0.0, gphi = 0.0, amax = -INFINITY;
        double S[NP];
        // posterior-weighted score sums of the staged observations lane, lane + 32, ..: after every
        // round lane j adds up row j of the (observation x node) scratch, so the sums need registers,
        // not a second jt x 32 array (shared memory per warp is what bounds the resident warps here)
        double zacc[L::NTJ], zacc_zi[L::NTJ];
        // Clusters larger than the staged tile, tile-outer form: pass A walks the tiles once (records
        // staged, separable tables built per tile) and sums the log-densities per node into asum;
        // then the node weights are known and pass B walks the tiles again for the scores, with the
        // same per-round row sums as the staged path.  Every point is evaluated twice, from the tables;
        // zbar is written un-normalised (the epilogue below scales it, as for the streaming path).
        static_assert(GLMMA_JT_CAP <= 32, "a lane owns one observation of a staged tile in the row sums");
        bool tiled_done = false;
        if (!fast && ni > jt && A.asum_doubles > 0) {
            bool tok = true;
            if (A.em_mode != 2) {
                for (int j0 = 0; j0 < ni; j0 += jt) {
                    const int cnt = min(jt, ni - j0);
                    __syncwarp();
                    stage_tile<F, QY, QZ>(d, r0 + j0, cnt, lane, rec);
                    __syncwarp();
                    tok = build_tables<F, QY, QZ>(rec, cnt, lane, nagq, rcp_nagq, gx, m, ri, cx, tabF, tabL);
                    __syncwarp();
                    if (!tok) break;
                    for (int r = 0; r < nrounds; ++r) {
                        const int k = (r << 5) + lane;
                        const bool valid = k < K;
                        double b[Q];
                        int idx[Q];
                        double a = node_setup<Q>(valid ? k : K - 1, nagq, rcp_nagq, gx, glw, lw_const, m, ri, A.prior, b, idx);
                        if (j0 > 0) a = 0.0;
#pragma unroll (F::GENERIC_UNROLL)
                        for (int j = 0; j < cnt; ++j) {
                            const double* rj = rec + j * L::W;
                            double eta, ez, emu, elam, ld, se, sz, sp;
                            lin_pred<F, QY, QZ, true>(rj, b, idx, j, nagq, tabF, tabL, cx, eta, ez, emu, elam);
                            F::template point<false, true>(rj[0], rj[1], eta, ez, emu, elam, A.fam, cx, ld, se, sz, sp);
                            a += ld;
                        }
                        if (valid) asum[k] = j0 > 0 ? asum[k] + a : a;
                    }
                }
            }
            if (tok) {
                // node weights (kept in asum), their sum and the shift (the true maximum)
                __syncwarp();
                sum_e = 0.0;
                if (A.em_mode == 2) {
                    for (int k = lane; k < K; k += 32) { const double e = A.pby[i * K + k]; asum[k] = e; sum_e += e; }
                } else {
                    double mx = -INFINITY;
                    for (int k = lane; k < K; k += 32) mx = fmax(mx, asum[k]);
                    mx = warp_max(mx);
                    shift = fabs(mx) < INFINITY ? mx : 0.0;
                    for (int k = lane; k < K; k += 32) {
                        const double e = fast_exp(asum[k] - shift, cx);
                        asum[k] = e;
                        sum_e += e;
                        if (A.em_mode == 1) A.pby[i * K + k] = e;
                    }
                }
                sum_e = warp_sum(sum_e);
                __syncwarp();
                gphi = 0.0;
#pragma unroll
                for (int t = 0; t < NP; ++t) S[t] = 0.0;
                if (SCORE) {
                    for (int j0 = 0; j0 < ni; j0 += jt) {
                        const int cnt = min(jt, ni - j0);
                        __syncwarp();
                        stage_tile<F, QY, QZ>(d, r0 + j0, cnt, lane, rec);
                        __syncwarp();
                        tok = build_tables<F, QY, QZ>(rec, cnt, lane, nagq, rcp_nagq, gx, m, ri, cx, tabF, tabL);
                        __syncwarp();
                        if (!tok) break;          // M-step on a cluster whose tables leave the range: streaming path
                        double za = 0.0, zz = 0.0;
                        for (int r = 0; r < nrounds; ++r) {
                            const int k = (r << 5) + lane;
                            const bool valid = k < K;
                            double b[Q];
                            int idx[Q];
                            node_setup<Q>(valid ? k : K - 1, nagq, rcp_nagq, gx, glw, lw_const, m, ri, A.prior, b, idx);
                            const double e = valid ? asum[k] : 0.0;
                            if (j0 == 0) {
                                int t = 0;
#pragma unroll
                                for (int dd = 0; dd < Q; ++dd)
#pragma unroll
                                    for (int ee = dd; ee < Q; ++ee) { S[t] = fma(e * b[dd], b[ee], S[t]); ++t; }
                            }
                            double vphi = 0.0;
#pragma unroll (F::GENERIC_UNROLL)
                            for (int j = 0; j < cnt; ++j) {
                                const double* rj = rec + j * L::W;
                                double eta, ez, emu, elam, ld, se, sz, sp;
                                lin_pred<F, QY, QZ, true>(rj, b, idx, j, nagq, tabF, tabL, cx, eta, ez, emu, elam);
                                F::template point<true, true>(rj[0], rj[1], eta, ez, emu, elam, A.fam, cx, ld, se, sz, sp);
                                scr_eta[j * L::SCRW + lane] = se;
                                if (F::HAS_ZI) scr_zi[j * L::SCRW + lane] = sz;
                                if (F::HAS_PHI) vphi += sp;
                            }
                            if (F::HAS_PHI) gphi = fma(e, vphi, gphi);
                            ev[lane] = e;
                            __syncwarp();
                            if (lane < cnt) {
                                const double* re = scr_eta + lane * L::SCRW;
                                const double* rz = scr_zi + lane * L::SCRW;
                                double p1[4] = {0.0, 0.0, 0.0, 0.0}, p2[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                                for (int c = 0; c < 32; ++c) {
                                    const double ec = ev[c];
                                    p1[c & 3] = fma(ec, re[c], p1[c & 3]);
                                    if (F::HAS_ZI) p2[c & 3] = fma(ec, rz[c], p2[c & 3]);
                                }
                                za += (p1[0] + p1[1]) + (p1[2] + p1[3]);
                                if (F::HAS_ZI) zz += (p2[0] + p2[1]) + (p2[2] + p2[3]);
                            }
                            __syncwarp();
                        }
                        if (lane < cnt) {
                            d.zbar[r0 + j0 + lane] = za;
                            if (F::HAS_ZI) d.zbar_zi[r0 + j0 + lane] = zz;
                        }
                    }
                }
                tiled_done = tok;
            }
            __syncwarp();
        }
        for (int attempt = 0; attempt < 2 && !tiled_done; ++attempt) {
            sum_e = 0.0; gphi = 0.0;
#pragma unroll
            for (int t = 0; t < NP; ++t) S[t] = 0.0;
            if (SCORE) {
                if (fast) {
#pragma unroll
                    for (int t = 0; t < L::NTJ; ++t) { zacc[t] = 0.0; zacc_zi[t] = 0.0; }
                } else {
                    for (int j = lane; j < ni; j += 32) {
                        d.zbar[r0 + j] = 0.0;
                        if (F::HAS_ZI) d.zbar_zi[r0 + j] = 0.0;
                    }
                }
            }
            __syncwarp();
            for (int rr = 0; rr < nrounds; ++rr) {
                const int r = (rr == 0) ? rc : (rr <= rc ? rr - 1 : rr);
                const int k = (r << 5) + lane;
                const bool valid = k < K;
                double b[Q];
                int idx[Q];
                double a = node_setup<Q>(valid ? k : K - 1, nagq, rcp_nagq, gx, glw, lw_const, m, ri, A.prior, b, idx);
                double vphi = 0.0;
                if (fast) {
#pragma unroll (F::GENERIC_UNROLL)
                    for (int j = 0; j < ni; ++j) {
                        const double* rj = rec + j * L::W;
                        double eta, ez, emu, elam, ld, se, sz, sp;
                        lin_pred<F, QY, QZ, true>(rj, b, idx, j, nagq, tabF, tabL, cx, eta, ez, emu, elam);
                        F::template point<SCORE, true>(rj[0], rj[1], eta, ez, emu, elam, A.fam, cx, ld, se, sz, sp);
                        a += ld;
                        if (SCORE) {
                            scr_eta[j * L::SCRW + lane] = se;
                            if (F::HAS_ZI) scr_zi[j * L::SCRW + lane] = sz;
                            if (F::HAS_PHI) vphi += sp;
                        }
                    }
                } else {
                    // large cluster: stream tiles, log-density only (scores are recomputed below)
                    for (int j0 = 0; j0 < ni; j0 += jt) {
                        const int cnt = min(jt, ni - j0);
                        __syncwarp();
                        stage_tile<F, QY, QZ>(d, r0 + j0, cnt, lane, rec);
                        __syncwarp();
                        for (int j = 0; j < cnt; ++j) {
                            const double* rj = rec + j * L::W;
                            double eta, ez, emu, elam, ld, se, sz, sp;
                            lin_pred<F, QY, QZ, false>(rj, b, idx, j, nagq, tabF, tabL, cx, eta, ez, emu, elam);
                            F::template point<false, true>(rj[0], rj[1], eta, ez, emu, elam, A.fam, cx, ld, se, sz, sp);
                            a += ld;
                        }
                    }
                }
                if (valid) amax = fmax(amax, a);
                if (rr == 0 && attempt == 0) {
                    shift = warp_max(valid ? a : -INFINITY);
                    if (!(fabs(shift) < INFINITY)) shift = 0.0;
                }
                double e = valid ? fast_exp(a - shift, cx) : 0.0;
                if (A.em_mode == 2) e = valid ? A.pby[i * K + k] : 0.0;
                else if (A.em_mode == 1 && valid) A.pby[i * K + k] = e;
                sum_e += e;
                if (SCORE) {
                    int t = 0;
#pragma unroll
                    for (int dd = 0; dd < Q; ++dd)
#pragma unroll
                        for (int ee = dd; ee < Q; ++ee) { S[t] = fma(e * b[dd], b[ee], S[t]); ++t; }
                    if (fast) {
                        if (F::HAS_PHI) gphi = fma(e, vphi, gphi);
                        ev[lane] = e;
                        __syncwarp();
#pragma unroll
                        for (int t = 0; t < L::NTJ; ++t) {
                            const int j = lane + 32 * t;
                            if (j < ni) {
                                // row j against the node weights: ev[c] is a broadcast, the odd row
                                // pitch keeps the lanes' row reads on distinct banks
                                const double* re = scr_eta + j * L::SCRW;
                                const double* rz = scr_zi + j * L::SCRW;
                                double p1[4] = {0.0, 0.0, 0.0, 0.0}, p2[4] = {0.0, 0.0, 0.0, 0.0};   // four chains: latency
#pragma unroll
                                for (int c = 0; c < 32; ++c) {
                                    const double ec = ev[c];
                                    p1[c & 3] = fma(ec, re[c], p1[c & 3]);
                                    if (F::HAS_ZI) p2[c & 3] = fma(ec, rz[c], p2[c & 3]);
                                }
                                const double s1 = (p1[0] + p1[1]) + (p1[2] + p1[3]);
                                const double s2 = (p2[0] + p2[1]) + (p2[2] + p2[3]);
                                zacc[t] += s1;
                                if (F::HAS_ZI) zacc_zi[t] += s2;
                            }
                        }
                        __syncwarp();
                    } else {
                        for (int j0 = 0; j0 < ni; j0 += jt) {
                            const int cnt = min(jt, ni - j0);
                            __syncwarp();
                            stage_tile<F, QY, QZ>(d, r0 + j0, cnt, lane, rec);
                            __syncwarp();
                            for (int j = 0; j < cnt; ++j) {
                                const double* rj = rec + j * L::W;
                                double eta, ez, emu, elam, ld, se, sz, sp;
                                lin_pred<F, QY, QZ, false>(rj, b, idx, j, nagq, tabF, tabL, cx, eta, ez, emu, elam);
                                F::template point<true, true>(rj[0], rj[1], eta, ez, emu, elam, A.fam, cx, ld, se, sz, sp);
                                if (F::HAS_PHI) vphi += sp;
                                scr_eta[j * 32 + lane] = e * se;
                                if (F::HAS_ZI) scr_zi[j * 32 + lane] = e * sz;
                            }
                            __syncwarp();
                            // un-normalised partial sums over this round's nodes, flushed per tile
                            for (int j = lane; j < cnt; j += 32) {
                                double s1 = 0.0, s2 = 0.0;
                                for (int c = 0; c < 32; ++c) {
                                    const int cc = (c + lane) & 31;
                                    s1 += scr_eta[j * 32 + cc];
                                    if (F::HAS_ZI) s2 += scr_zi[j * 32 + cc];
                                }
                                d.zbar[r0 + j0 + j] += s1;
                                if (F::HAS_ZI) d.zbar_zi[r0 + j0 + j] += s2;
                            }
                        }
                        if (F::HAS_PHI) gphi = fma(e, vphi, gphi);
                    }
                }
            }
            sum_e = warp_sum(sum_e);
            // the shift came from the central round; if that was not enough to keep the sum
            // in range, repeat once with the true maximum over all nodes
            if ((sum_e > 0.0 && sum_e < INFINITY) || A.em_mode == 2) break;
            const double mx = warp_max(amax);
            if (attempt == 1 || !(fabs(mx) < INFINITY) || mx == shift) break;
            shift = mx;
        }
        const double ll = A.em_mode == 2 ? 0.0 : shift + log(sum_e) + d.logdet[i] + c_ll;
        // false for NaN and +-Inf; M-step: the E-step zeroed the weights of the clusters it dropped
        const bool ok = A.em_mode == 2 ? (sum_e > 0.0 && sum_e < INFINITY) : fabs(ll) < INFINITY;
        if (A.em_mode == 1 && !ok)
            for (int k = lane; k < K; k += 32) A.pby[i * K + k] = 0.0;
        if (lane == 0) {
            if (ok) { tot[0] += w * ll; tot[1] += w; } else tot[2] += 1.0;
            if (d.ll_i) d.ll_i[i] = w * ll;
        }
        if (SCORE) {
            const double inv = ok ? w / sum_e : 0.0;
            if (F::HAS_PHI) {
                tot[3] = fma(inv, gphi, tot[3]);
                if (lane == 0 && ok) tot[3] = fma(w, c_phi, tot[3]);
            }
#pragma unroll
            for (int t = 0; t < NP; ++t) tot[4 + t] = fma(inv, S[t], tot[4 + t]);
            if (d.S_i) {
                double Sr[NP];
#pragma unroll
                for (int t = 0; t < NP; ++t) Sr[t] = inv * warp_sum(S[t]);
                if (lane == 0) {
                    int t = 0;
#pragma unroll
                    for (int dd = 0; dd < Q; ++dd)
#pragma unroll
                        for (int ee = dd; ee < Q; ++ee) {
                            d.S_i[i * Q * Q + dd * Q + ee] = Sr[t];
                            d.S_i[i * Q * Q + ee * Q + dd] = Sr[t];
                            ++t;
                        }
                }
            }
            __syncwarp();
            if (fast) {
#pragma unroll
                for (int t = 0; t < L::NTJ; ++t) {
                    const int j = lane + 32 * t;
                    if (j < ni) {
                        d.zbar[r0 + j] = inv * zacc[t];
                        if (F::HAS_ZI) d.zbar_zi[r0 + j] = inv * zacc_zi[t];
                    }
                }
            } else {
                for (int j = lane; j < ni; j += 32) {
                    d.zbar[r0 + j] *= inv;
                    if (F::HAS_ZI) d.zbar_zi[r0 + j] *= inv;
                }
            }
            if (A.fused) {
                // no score_reduce launch in fused mode: this cluster's part of X'zbar (X_zi'zbar_zi),
                // column l summed over the warp and kept by lane l
                __syncwarp();
                for (int j0 = 0; j0 < ni; j0 += 32) {
                    const int j = j0 + lane;
                    const double zb = j < ni ? d.zbar[r0 + j] : 0.0;
                    for (int l = 0; l < d.p; ++l) {
                        const double v = warp_sum(j < ni ? d.X[r0 + j + (int64_t)l * d.N] * zb : 0.0);
                        if (lane == l) gcol += v;
                    }
                    if (F::HAS_ZI) {
                        const double zz = j < ni ? d.zbar_zi[r0 + j] : 0.0;
                        for (int l = 0; l < d.p_zi; ++l) {
                            const double v = warp_sum(j < ni ? d.X_zi[r0 + j + (int64_t)l * d.N] * zz : 0.0);
                            if (lane == d.p + l) gcol += v;
                        }
                    }
                }
            }
        }
        __syncwarp();
      }
    }

    // block partials, fixed order (deterministic for a fixed launch shape)
    __syncthreads();
    double* red = smem;    // reuse: wpb * PSTRIDE doubles
#pragma unroll
    for (int a = 0; a < GLMMA_NACC; ++a) {
        const double v = warp_sum(tot[a]);
        if (lane == 0) red[wib * GLMMA_PSTRIDE + a] = v;
    }
    red[wib * GLMMA_PSTRIDE + GLMMA_NACC + lane] = gcol;
    __syncthreads();
    if (threadIdx.x < GLMMA_PSTRIDE) {
        double v = 0.0;
        for (int ww = 0; ww < wpb; ++ww) v += red[ww * GLMMA_PSTRIDE + threadIdx.x];
        A.partials[(int64_t)blockIdx.x * GLMMA_PSTRIDE + threadIdx.x] = v;
    }
    // Finalisation by the last block to finish (this kernel is the last launch of an evaluation, so the
    // partials of the earlier launches are complete): its fixed-order sum over ALL block partials is
    // the result vector, whichever block happens to be last -- deterministic.
    if (A.fin_counter != nullptr) {
        // (no static shared memory in this kernel: its dynamic allocation is raised to the opt-in maximum)
        int* s_last = reinterpret_cast<int*>(smem + (size_t)wpb * GLMMA_PSTRIDE);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) *s_last = (atomicAdd(A.fin_counter, 1u) == gridDim.x - 1) ? 1 : 0;
        __syncthreads();
        if (*s_last) {
            __threadfence();
            finalize_block(A.fin_partials, A.fin_blocks, nullptr, 0, true, d.p, d.n_phis, d.p_zi, Q, A.fin_want_score,
                           A.fin_result, smem + (size_t)wpb * GLMMA_PSTRIDE + 2, A.fin_cpart, A.fin_ncpart, A.fin_ll_const,
                           A.em_mode != 2, &A.xchg, &A.hout);
            if (threadIdx.x == 0) {
                *A.fin_counter = 0u;
                if (A.defer_count != nullptr) *A.defer_count = 0u;
            }
        }
    }
}

// ------------------------------------------------------------------------------
// prep + cluster sums in one pass: 8 lanes per cluster walk its rows (coalesced: a
// cluster's rows are contiguous in every column), write xb, xg, cll, cphi and reduce
//   csum[i] = { sum_j cll_j, sum_j cphi_j, sum_j y1_j xb_j, (Z'y1)[0..QY) }
// (the last two feed the linear-eta hoist of the register variant).
// ------------------------------------------------------------------------------
#define GLMMA_NCSUM (3 + GLMMA_QMAX)

// obs_const() of the count families depends on the (integer) response only: tabulate it
// for y = 0 .. GLMMA_YTAB-1 once per evaluation instead of three lgamma/digamma calls
// per observation (ncu r1: prep was 0.34 ms of a 4 ms evaluation, all of it special functions).
#define GLMMA_YTAB 1024
template <class F>
__global__ void ytab_kernel(FamPar fam, double* ytab) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= GLMMA_YTAB) return;
    double cll, cphi;
    F::obs_const((double)y, 0.0, fam, cll, cphi);
    ytab[2 * y] = cll;
    ytab[2 * y + 1] = cphi;
}
template <class F>
__device__ __forceinline__ void obs_const_tab(double y1, double y2, const FamPar& fam, const double* ytab,
                                              double& cll, double& cphi) {
    if (F::Y_TABLE && ytab != nullptr) {
        const int iy = (int)y1;
        if (y1 >= 0.0 && y1 < (double)GLMMA_YTAB && (double)iy == y1) {
            const double2 t = *reinterpret_cast<const double2*>(ytab + 2 * iy);
            cll = t.x; cphi = t.y;
            return;
        }
    }
    F::obs_const(y1, y2, fam, cll, cphi);
}

// obs_const() out of line: the register variant calls it (fused mode) only for responses the y table
// does not cover, and libm's lgamma inlined into that kernel would cost it registers everywhere
template <class F>
__device__ __noinline__ void obs_const_slow(double y1, double y2, FamPar fam, double* out) {
    double cll, cphi;
    F::obs_const(y1, y2, fam, cll, cphi);
    out[0] = cll; out[1] = cphi;
}

template <class F>
__global__ void __launch_bounds__(256) prep_cluster_kernel(PrepArgs A, CoefPar C) {
    const DevData& d = A.d;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t pos = gtid >> 3;
    const int lg = (int)(gtid & 7);
    const unsigned gmask = 0xffu << ((threadIdx.x & 31) & ~7);
    const bool live = pos < (A.list ? A.n_list : d.n);
    const int64_t i = (live && A.list) ? (int64_t)A.list[pos] : pos;
    const int64_t r0 = live ? d.row_ptr[i] : 0, r1 = live ? d.row_ptr[i + 1] : 0;
    double s[GLMMA_NCSUM];
#pragma unroll
    for (int t = 0; t < GLMMA_NCSUM; ++t) s[t] = 0.0;
    for (int64_t row = r0 + lg; row < r1; row += 8) {
        double xb = d.offset ? d.offset[row] : 0.0;
        for (int l = 0; l < d.p; ++l) xb = fma(d.X[row + (int64_t)l * d.N], C.betas[l], xb);
        d.xb[row] = xb;
        if (F::HAS_ZI) {
            double xg = d.offset_zi ? d.offset_zi[row] : 0.0;
            for (int l = 0; l < d.p_zi; ++l) xg = fma(d.X_zi[row + (int64_t)l * d.N], C.gammas[l], xg);
            d.xg[row] = xg;
        }
        const double y = d.y1[row];
        double cll, cphi;
        obs_const_tab<F>(y, F::USES_Y2 ? d.y2[row] : 0.0, A.fam, A.ytab, cll, cphi);
        d.cll[row] = cll;
        if (F::HAS_PHI) d.cphi[row] = cphi;
        s[0] += cll;
        if (F::HAS_PHI) s[1] += cphi;
        if (F::LIN_ETA) {
            s[2] = fma(y, xb, s[2]);
            for (int dd = 0; dd < d.q_y; ++dd) s[3 + dd] = fma(y, d.Z[row + (int64_t)dd * d.N], s[3 + dd]);
        }
    }
#pragma unroll
    for (int t = 0; t < GLMMA_NCSUM; ++t) s[t] = group_sum<8>(s[t], gmask);
    if (live && lg == 0) {
#pragma unroll
        for (int t = 0; t < GLMMA_NCSUM; ++t) d.csum[i * GLMMA_NCSUM + t] = s[t];
    }
}

// ------------------------------------------------------------------------------
// eval, register variant: clusters of at most NJ observations, nagq <= GLMMA_NAGQ_FAST,
// K <= GLMMA_K_FAST.  The observation loop is unrolled in branch-free chunks of four:
// per-observation score terms and their posterior-weighted sums live in registers,
// every shared-memory operand has an immediate offset, and the independent point
// evaluations give the FP64 pipe instruction-level parallelism.  The functors run with
// CAREFUL = false, which is exact when no link clamp is active on the cluster's grid and
// all separable factors are in range; clusters that fail that test (or are larger than
// NJ) are flagged in d.defer and handled by eval_kernel.
//
// The kernel is self-contained per evaluation ("fused"): it stages the cluster's design columns
// (X, X_zi, offsets) next to its records, forms xb = X beta + offset and xg = X_zi gamma + offset_zi
// itself and accumulates X'zbar and X_zi'zbar_zi per warp (lane l owns column l), so there is no prep
// pass over the data, no xb / zbar round trip through HBM and no score-reduce launch.  The
// node-independent terms of log f and d log f / d phis do not enter the posterior node weights; they
// are added to the totals once per evaluation (const_sums_kernel over the rows of this variant's
// clusters + beta' X'y1 for the linear-eta hoist; engine.cu) and only a cluster that drops out of the
// sums (non-finite, or left to eval_kernel) takes its share back here, on a rare path.
//
// Shared memory per block: [fastmath | gx, glw | column table | node table (z.., lw per node) | node
// idx | per-warp areas: 2 prefetch buffers (records, scalars, design columns), acc (also scratch), wv,
// per-lane totals, warp totals, two list records, tabF, tabL, mode-specific tail (ZP / two-pass state)].
//
// Forms of the round loop, chosen at compile time:
//   FAST_KIND 1 / 2   carrier families, chunks of four observations                     (one pass, NJ <= 16)
//   FAST_KIND 0       one functor call per observation; zero-inflated / hurdle families with the zero part
//                     hoisted to the node and the observations sorted by class ("ZP")   (one pass, NJ <= 16)
//   FAST_KIND 3       censored.normal: Gaussian part as a quadratic form per node ("GQ") (one pass, NJ <= 16)
//   NJ == 32 (TP)     two passes over tiles of 32 rows, any cluster size, carrier families
//   PAIR              two random-intercept clusters per warp, carrier families
// ------------------------------------------------------------------------------
#define GLMMA_NAGQ_FAST 16
#ifndef GLMMA_NODE_PERM
#define GLMMA_NODE_PERM 1
#endif
#define GLMMA_K_FAST 512
#define GLMMA_NTAB_MAX_K 256
// Launch shape: 128 registers per thread at any block size from GLMMA_EVAL_THREADS (four blocks per
// SM) to GLMMA_FAST_MAX_THREADS (one); the host picks the shape that keeps most warps resident.
#define GLMMA_FAST_MAX_THREADS 512
#define GLMMA_COLTAB_MAX 48        // staged columns: y1, y2, Z, Z_zi, X, X_zi, offsets

// one staged column: where it starts in HBM, where row j of it goes in the prefetch buffer
struct __align__(16) ColEnt {
    const double* base;
    int off;            // doubles from the start of the buffer
    int stride;         // doubles between consecutive rows
};

template <class F, int QY, int QZ, int NJ, bool PAIR = false>
struct FastLayout {
    using L = EvalLayout<F, QY, QZ>;
    static constexpr int Q = QY + QZ;
    static constexpr int W = L::W;
    static constexpr int NT = L::NTABL;
    static constexpr int NG = GLMMA_NAGQ_FAST;
    // separable tables: the NJ observations of one (dimension, node index) are contiguous, row pitch
    // RP: even (16-byte pairs) with RP / 2 odd, so the 16-byte loads of a quarter warp, whose lanes
    // walk consecutive node indices of the fastest dimension, fall on distinct bank groups
    static constexpr int RP = NJ + 2;
    // table rows: (dimension e, node index a) -> row e * nagq + a, i.e. only the nagq rows per dimension
    // that the grid has (nagq = 7 at q = 3 keeps 21 + 7 rows instead of 48 + 16: shared memory per warp
    // is what bounds the resident warps of the zero-inflated q = 3 kernels and of the 16-observation tile)
    // FAST_KIND 3 (Gaussian part hoisted, families.cuh): compact records of the cluster's censored
    // observations {A, C_0.., y, sign} take the place of the (unused) separable table
    static constexpr bool GQ = F::FAST_KIND == 3;
    static constexpr int CRW = (QY + 3 + 1) & ~1;
    __host__ __device__ static int tfd(int nagq) { return F::NEED_EMU ? Q * nagq * RP : (GQ ? NJ * CRW : 0); }   // doubles of tabF
    __host__ __device__ static int tld(int nagq) { return NT * (QZ > 0 ? nagq : 1) * RP; }   // doubles of tabL
    static constexpr int NACCS = (F::HAS_ZI || GQ) ? 2 : 1;
    static constexpr int NTOT = 2 + Q * (Q + 1) / 2;        // per-lane totals: g_phi, (pad), S packed
    // per-cluster scalars prefetched with the records, one packed record per cluster in HBM
    // (pack_cluster_kernel): modes, R^-1, logdet, Z'y1, weight
    static constexpr int NSC = Q + Q * (Q + 1) / 2 + 1 + QY + 1;
    // PAIR (random intercepts, kernel header): two adjacent clusters share a warp; both packed records are staged
    static constexpr int NSCP = PAIR ? 2 * NSC : NSC;
    static constexpr int XOFF = ((NJ * W + NSCP) + 1) & ~1;  // design columns follow records + scalars
    // NJ = 32: two-pass form (kernel header) -- observations reduced in chunks of eight, node terms of the cluster
    // (a_k then e_k, and the phi-score's node sums) kept per warp after the tables
    static constexpr bool TP = NJ == 32;
    static constexpr int NJR = (NJ <= 8 || TP) ? 8 : 16;    // rows of the per-observation reduction scratch (power of two)
    __host__ __device__ static int tpd(int K) { return TP ? 2 * ((K + 1) & ~1) + 8 : 0; }   // + the cluster's state between its records
    static constexpr int CG = 32 / NJ;                      // columns copied side by side by one prefetch trip
    static_assert(NJ <= 32, "a lane stages one observation");
    __host__ __device__ static int pre(int xw) { return XOFF + ((NJ * xw + 1) & ~1); }      // one prefetch buffer
    static constexpr int FIXED0 = (NACCS * NJR * 32 + NJ + NTOT * 32 + 4 + 4 + 1) & ~1;   // scratch, wv, per-lane totals, warp totals, two list records
    // "ZP" mode (zero part hoisted, kernel header): per-lane marginal node weights of the zero-part
    // dimension (nz x 32) and the per-node sum of log(1 + exp(eta_zi)) over the cluster (nz), after tabL
    static constexpr bool ZP = F::HAS_ZI && QZ <= 1 && F::FAST_KIND == 0;
    // ... and plogis(eta_zi) per (observation slot, zero-part node index), pitch nz, for the epilogue (the zero-part
    // table itself lives in the accumulator scratch, which serves as the zero responses' score stash during the rounds)
    __host__ __device__ static int zpd(int nagq) {
        return ZP ? (QZ > 0 ? nagq : 1) * 32 + GLMMA_NAGQ_FAST + ((NJ * (QZ > 0 ? nagq : 1) + 1) & ~1) : 0;
    }
    __host__ __device__ static int fixed(int nagq) { return FIXED0 + tfd(nagq) + tld(nagq) + zpd(nagq); }
    __host__ __device__ static size_t warp_stride(int xw, int nagq, int K) { return (size_t)(2 * pre(xw) + fixed(nagq) + tpd(K) + 1) & ~(size_t)1; }
    // node table: z_0.., log wGH per node (K x (Q + 1) doubles) + packed table rows (K ints).  For
    // large grids the double part is dropped (the round loop then reads gx / glw through the packed
    // rows): at K = 343 it is 11 KB of the block's shared memory and costs one resident block per SM.
    __host__ __device__ static bool has_ntab(int K) { return K <= GLMMA_NTAB_MAX_K; }
    __host__ __device__ static size_t node_doubles(int K) {
        return (((has_ntab(K) ? (size_t)K * (Q + 1) : 0) + (K + 1) / 2) + 1) & ~(size_t)1;
    }
    __host__ __device__ static size_t coltab_doubles() { return 2 * GLMMA_COLTAB_MAX; }
    __host__ static size_t smem_bytes(int warps, int K, int xw, int nagq) {
        return sizeof(double) * (L::head_doubles() + coltab_doubles() + node_doubles(K) + (size_t)warps * warp_stride(xw, nagq, K));
    }
};

// node-independent terms of one observation, out of line (rare paths of the register variant only:
// a cluster that leaves the sums, per-cluster contributions): {c_ll + [LIN_ETA] y1 xb, c_phi}
template <class F>
__device__ __noinline__ double2 obs_consts_rare(double y1, double y2, double xb, FamPar fam, const double* ytab) {
    double cll, cphi;
    obs_const_tab<F>(y1, y2, fam, ytab, cll, cphi);
    if (F::LIN_ETA) cll = fma(y1, xb, cll);
    return make_double2(cll, cphi);
}

template <class F, int QY, int QZ, bool SCORE, int NJ, bool PAIR = false>
__global__ void __launch_bounds__(GLMMA_FAST_MAX_THREADS, 1) eval_fast_kernel(EvalArgs A) {
    using FL = FastLayout<F, QY, QZ, NJ, PAIR>;
    using L = EvalLayout<F, QY, QZ>;
    constexpr int Q = L::Q;
    constexpr int NP = L::NP;
    constexpr int W = L::W;
    constexpr int NG = GLMMA_NAGQ_FAST;
    constexpr int RP = FL::RP;
    const DevData& d = A.d;
    extern __shared__ double smem[];
    double* fmtab = smem;
    double* gx = smem + GLMMA_FM_DOUBLES;
    double* glw = gx + GLMMA_NAGQ_MAX;
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int nagq = A.grid->nagq;
    const int K = A.grid->K;
    // K > GLMMA_NTAB_MAX_K needs nagq^Q > 256 with nagq <= 16, i.e. Q >= 3: for Q < 3 the test folds away
    const bool has_ntab = Q < 3 || FL::has_ntab(K);
    ColEnt* coltab = reinterpret_cast<ColEnt*>(smem + L::head_doubles());
    double* ntab = smem + L::head_doubles() + FL::coltab_doubles();       // K x (Q + 1): z_0.., log wGH (small grids)
    int* nidx = reinterpret_cast<int*>(ntab + (has_ntab ? (size_t)K * (Q + 1) : 0));   // K packed node indices
    const int xw = A.xw;
    const int PRE = FL::pre(xw);
    double* pre0 = smem + L::head_doubles() + FL::coltab_doubles() + FL::node_doubles(K) + wib * FL::warp_stride(xw, nagq, K);
    // fixed-size areas sit between the prefetch buffers and the tables, at constant offsets below tabF (the
    // table sizes depend on nagq: keeping them last leaves one anchor register for everything but tabL)
    double* tabF = pre0 + 2 * PRE + FL::FIXED0;
    double* tabL = tabF + FL::tfd(nagq);
    double* zp_m = tabL + FL::tld(nagq);                     // ZP mode: nz x 32 marginal weights, then nz sums of log(1 + lambda)
    double* zp_lz = zp_m + (QZ > 0 ? nagq : 1) * 32;
    double* zp_pi = zp_lz + GLMMA_NAGQ_FAST;                 // n_i x nz values of plogis(eta_zi), by slot
    double* acc_eta = tabF - FL::FIXED0;
    double* acc_zi = acc_eta + FL::NJR * 32;
    double* wv = acc_eta + FL::NACCS * FL::NJR * 32;         // staged second record field (y + phi, trials ...), contiguous
    double* totl = wv + NJ;                                  // NTOT x 32 per-lane totals
    double* tots = totl + FL::NTOT * 32;                     // loglik, sum_w, n_nonfinite (lane 0)
    double gcol = 0.0;                                       // lane l sums column l of X'zbar | X_zi'zbar_zi

    load_fmtab(fmtab, A.fmtab);
    if (threadIdx.x < GLMMA_NAGQ_MAX) {
        gx[threadIdx.x] = A.grid->x[threadIdx.x];
        glw[threadIdx.x] = A.grid->lw[threadIdx.x];
    }
    // staged columns: record fields first (row stride W), then the design columns (column-major per
    // cluster, as in HBM): X | X_zi | offset | offset_zi
    const int ncol_rec = 1 + (F::USES_Y2 ? 1 : 0) + QY + QZ;
    const int ncols = ncol_rec + xw;
    if (threadIdx.x < ncols) {
        const int c = threadIdx.x;
        ColEnt e;
        e.stride = W;
        int f = c;
        if (f == 0) { e.base = d.y1; e.off = 0; }
        else if (F::USES_Y2 && f == 1) { e.base = d.y2; e.off = 1; }
        else {
            f -= 1 + (F::USES_Y2 ? 1 : 0);
            if (f < QY) { e.base = d.Z + (int64_t)f * d.N; e.off = L::OFF_Z + f; }
            else if (f < QY + QZ) { e.base = d.Z_zi + (int64_t)(f - QY) * d.N; e.off = L::OFF_Z + f; }
            else {
                f -= QY + QZ;
                e.stride = 1;
                e.off = FL::XOFF + f * NJ;
                const int pz_ = F::HAS_ZI ? d.p_zi : 0;
                if (f < d.p) e.base = d.X + (int64_t)f * d.N;
                else if (f < d.p + pz_) e.base = d.X_zi + (int64_t)(f - d.p) * d.N;
                else if (f == d.p + pz_ && d.offset) e.base = d.offset;
                else e.base = d.offset_zi;
            }
        }
        coltab[c] = e;
    }
    __syncthreads();
    // Node order.  A quarter warp's 16-byte loads of table rows are conflict-free when its eight lanes read
    // the same row or rows whose indices differ mod 8 (row pitch RP with RP / 2 odd).  In the natural order
    // (index of dimension 0 fastest) eight consecutive nodes wrap around nAGQ (11: rows 8, 9, 10, 0, 1 ...) and
    // collide (ncu, C3: 2.25 excess wavefronts per table load).  So the nodes are walked in two sections:
    // first those with a_0 < 8 * (nAGQ / 8), a_0 fastest -- every aligned group of eight lanes has a_0 =
    // 8 g .. 8 g + 7 and the other indices equal --, then the remaining nAGQ % 8 values of a_0, again a_0
    // fastest: a group repeats a few rows (broadcast) that differ mod 8.  nAGQ < 8: the natural order.
    const int nfull = nagq & ~7, nrem = nagq - nfull, nhi = K / nagq;
    auto node_id = [&](int k) {          // position in the walk -> natural node index a_0 + nAGQ (a_1 + nAGQ a_2)
        if (k < nfull * nhi) return (k % nfull) + nagq * (k / nfull);
        const int k2 = k - nfull * nhi;
        return nfull + (k2 % nrem) + nagq * (k2 / nrem);
    };
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        int kk = GLMMA_NODE_PERM ? node_id(k) : k, packed = 0;
        double lw = A.grid->lw_const;
#pragma unroll
        for (int e = 0; e < Q; ++e) {
            const int a = kk % nagq;
            kk /= nagq;
            packed |= (e * nagq + a) << (8 * e);       // table row of (dimension e, node index a)
            if (has_ntab) ntab[k * (Q + 1) + e] = gx[a];
            lw += glw[a];
        }
        if (has_ntab) ntab[k * (Q + 1) + Q] = lw;
        nidx[k] = packed;
    }
    for (int t = lane; t < FL::NTOT * 32 + 4; t += 32) totl[t] = 0.0;
    __syncthreads();
    const MathCtx cx = make_mathctx(fmtab + 2 * (threadIdx.x & 7), fmtab + GLMMA_FM_LOG_DOUBLES + (threadIdx.x & 15), A.fmc);
    const int nrounds = (K + 31) >> 5;
    int rc;                                     // the round holding the central node
    {
        const int c = (K - 1) >> 1, a0c = c % nagq, hic = c / nagq;
        rc = (GLMMA_NODE_PERM ? (a0c < nfull ? hic * nfull + a0c : nfull * nhi + hic * nrem + (a0c - nfull)) : c) >> 5;
    }
    constexpr int SC_M = 0, SC_RI = Q, SC_LD = Q + NP, SC_ZTY = Q + NP + 1, SC_W = Q + NP + 1 + QY;
    const int p_x = d.p, p_z = F::HAS_ZI ? d.p_zi : 0;
    const bool has_off = d.offset != nullptr, has_offz = F::HAS_ZI && d.offset_zi != nullptr;

    // Asynchronous prefetch (cp.async, LDGSTS) of one cluster's columns and per-cluster scalars into a
    // shared-memory buffer, issued one cluster ahead so that the global-memory latency is covered by the
    // previous cluster's arithmetic.  Lane (c, j) = (lane / NJ, lane % NJ) copies row j of columns
    // c, c + CG, ...: a trip moves CG columns side by side.
    const int pj = lane % NJ, pc = lane / NJ;
    auto prefetch = [&](int ci, int c_r0, int c_ni, double* buf) {
        if (c_ni <= NJ) {
            if (pj < c_ni && pc < FL::CG) {
                for (int c = pc; c < ncols; c += FL::CG) {
                    const ColEnt e = coltab[c];
                    cp_async8(buf + e.off + pj * e.stride, e.base + c_r0 + pj);
                }
            }
            if (lane < FL::NSCP) cp_async8(buf + NJ * W + lane, d.clrec + ci * FL::NSC + lane);
        }
        cp_async_commit();
    };

    // positions in the cluster list fit 32 bits (n < 2^31)
    const int nwarps = (int)gridDim.x * wpb;
    const int p0 = (int)blockIdx.x * wpb + wib;
    const int n_pos = (int)A.n_order;
    const int4* __restrict__ order = A.order;
    // The current list record {cluster, first row, n_i, flags} is kept in registers; the record after it travels
    // through shared memory with the same cp.async group as the current cluster's data (lane 0, 16 bytes), two
    // clusters ahead -- a plain load into registers made its first consumer wait for the load in flight (ncu, C3:
    // 3-6 % of the stall samples, as a write-after-write hazard or a spill store depending on the build)
    int4 cur = make_int4(0, 0, 0, 0);
    int4* nslot = reinterpret_cast<int4*>(tots + 4);
    if (p0 < n_pos) {
        cur = order[p0];
        if (lane == 0 && p0 + nwarps < n_pos) cp_async16(reinterpret_cast<double*>(nslot), reinterpret_cast<const double*>(order + p0 + nwarps));
        prefetch(cur.x, cur.y, cur.z, pre0);
    }
    int pbuf = 0;
    for (int pos = p0; pos < n_pos; pos += nwarps) {
        double* rec = pre0 + pbuf * PRE;
        const double* xcur = rec + FL::XOFF;
        const double* sc = rec + NJ * W;
        // rotate the pipeline and start the next cluster's copies
        const int64_t i = cur.x;      // 64-bit: i * K, i * Q ... index arrays of up to n * K elements
        const int64_t r0c = cur.y;
        // (ncu, C3: 3.4 % of the stall samples sit on the first write to the dead fourth register of the 16-byte
        // `order` load while that load is in flight; keeping the register live instead -- nic = cur.z + cur.w --
        // cost 4 % through register pressure, so the stall stays)
        const int nic = cur.z;
        // two-pass tile: what this record asks for -- 1 pass A over its rows, 2 first tile of the cluster, 4 last
        // tile (the weights follow), 8 pass B over its rows; 15: a cluster of at most 32 observations in one record;
        // 0: padding of a warp's queue
        const int rflags = FL::TP ? cur.w : 0;
        // pair tile: clusters i (rows 0 .. nA - 1 of the staged tile) and i + 1 (the rest)
        const int nA = PAIR ? cur.w : nic;
        // (pos + 2 nwarps may pass 2^31: compare the remaining distance instead)
        cp_async_wait_all();
        __syncwarp();
        if (n_pos - pos > nwarps) {
            cur = nslot[pbuf];
            if (lane == 0 && n_pos - pos > 2 * nwarps)
                cp_async16(reinterpret_cast<double*>(nslot + (pbuf ^ 1)), reinterpret_cast<const double*>(order + pos + 2 * nwarps));
            prefetch(cur.x, cur.y, cur.z, pre0 + (pbuf ^ 1) * PRE);
        }
        pbuf ^= 1;
        if (FL::TP && rflags == 0) continue;
        if (nic > NJ) {
            if (lane == 0) { d.defer[i] = 1; if (A.defer_count) atomicAdd(A.defer_count, 1u); }
            continue;
        }
        // modes and sqrt(2) R^-1 (the factor of b_k = mode + sqrt(2) R^-1 z_k is folded in once)
        double m[Q], ri[NP];
        // (pair tile: an observation lane takes the record of ITS cluster)
        const double* sco = sc + ((PAIR && lane >= nA) ? FL::NSC : 0);
#pragma unroll
        for (int dd = 0; dd < Q; ++dd) m[dd] = sco[SC_M + dd];
#pragma unroll
        for (int t = 0; t < NP; ++t) ri[t] = GLMMA_SQRT2 * sco[SC_RI + t];
        // Lane j < n_i owns observation j: forms xb (xg) from the staged design columns, stages its
        // record, forms the coefficients of its separable factors
        //   exp(eta_jk) = prod_e exp(c_je x_{a_e(k)} + [e == 0] base_j)
        // and tests the range of eta (and eta_zi) over the whole grid: every factor is monotone in
        // its node, so the extremes are base -+ sum_e |c_e| x_max.  Outside the fast range (a link
        // clamp could be active, or a factor / fast_log_rcp argument could leave its range) the
        // cluster is left to eval_kernel.
        double2* coef = reinterpret_cast<double2*>(acc_eta);
        constexpr int NTc = FL::NT > 0 ? FL::NT : 1;
        // zero part tabulated (families.cuh, GLMMA_POINT_ARGS): lives in the accumulator scratch, past
        // the coefficients, until the rounds are over
        constexpr bool ZT = F::HAS_ZI && QZ <= 1;
        double2* tabZ = reinterpret_cast<double2*>(acc_eta + 2 * (NJ * Q + NJ * NTc));
        static_assert(!ZT || 2 * (NJ * Q + NJ * NTc) + 2 * NJ * NG <= FL::NACCS * FL::NJR * 32, "zero-part table does not fit");
        static_assert(!FL::ZP || NJ * 64 <= FL::NACCS * FL::NJR * 32, "score stash of the zero responses does not fit");
        // ZP mode: the cluster's observations are sorted into positive responses (slots 0 .. npos - 1) and
        // zero responses (npos .. n_i - 1); records, coefficient and table columns are laid out by slot, so
        // the round loop knows an observation's class from its position
        constexpr bool ZP = FL::ZP;
        int slot = lane, npos = nic;
        if (ZP) {
            const bool live = lane < nic;
            const bool posv = live && F::is_pos(rec[lane * W]);
            const unsigned pm = __ballot_sync(0xffffffffu, posv), zm = __ballot_sync(0xffffffffu, live && !posv);
            const unsigned lt = (1u << lane) - 1u;
            npos = __popc(pm);
            slot = posv ? __popc(pm & lt) : npos + __popc(zm & lt);
        }
        double hz0 = 0.0, hz1 = 0.0;      // hurdle families: sum of eta_zi = hz0 + hz1 b_zi over the zero responses
        bool bad = false;
        if (lane < nic) {
            double* r = rec + lane * W;
            if (!F::USES_Y2) r[1] = 0.0;
            {
                const double* xr = xcur + lane;
                double xb = has_off ? xr[(p_x + p_z) * NJ] : 0.0;
                for (int l = 0; l < p_x; ++l) xb = fma(xr[l * NJ], A.coef.betas[l], xb);
                r[2] = xb;
                if (F::HAS_ZI) {
                    double xg = has_offz ? xr[(p_x + p_z + (has_off ? 1 : 0)) * NJ] : 0.0;
                    for (int l = 0; l < p_z; ++l) xg = fma(xr[(p_x + l) * NJ], A.coef.gammas[l], xg);
                    r[L::OFF_XG] = xg;
                }
            }
            F::stage(r[0], r[1], A.fam);
            if (F::FAST_KIND == 1) wv[slot] = r[1];
            const double xmax = gx[0];
            if (F::NEED_EMU) {
                double base = r[2], amp = 0.0;
#pragma unroll
                for (int dd = 0; dd < QY; ++dd) base = fma(r[L::OFF_Z + dd], m[dd], base);
#pragma unroll
                for (int e = 0; e < Q; ++e) {
                    double c = 0.0;
#pragma unroll
                    for (int dd = 0; dd < QY; ++dd)
                        if (dd <= e) c = fma(r[L::OFF_Z + dd], ri[Tri<Q>::rm(dd, e)], c);
                    coef[slot * Q + e] = make_double2(c, e == 0 ? base : 0.0);
                    const double ae = fabs(c) * xmax;
                    if (!(ae + (e == 0 ? fabs(base) : 0.0) <= 230.0)) bad = true;
                    amp += ae;
                }
                if (!(base - amp >= F::LOG_EMU_LO && base + amp <= F::LOG_EMU_HI)) bad = true;
            } else if (F::HAS_ETA_RANGE) {
                double base = r[2], amp = 0.0;
#pragma unroll
                for (int dd = 0; dd < QY; ++dd) base = fma(r[L::OFF_Z + dd], m[dd], base);
#pragma unroll
                for (int e = 0; e < Q; ++e) {
                    double c = 0.0;
#pragma unroll
                    for (int dd = 0; dd < QY; ++dd)
                        if (dd <= e) c = fma(r[L::OFF_Z + dd], ri[Tri<Q>::rm(dd, e)], c);
                    amp += fabs(c) * xmax;
                }
                if (!(base - amp >= F::ETA_LO && base + amp <= F::ETA_HI)) bad = true;
            }
            if (F::HAS_ZI) {
                double base = r[L::OFF_XG], amp = 0.0;
#pragma unroll
                for (int dd = 0; dd < QZ; ++dd) base = fma(r[L::OFF_Z + QY + dd], m[QY + dd], base);
#pragma unroll
                for (int e = 0; e < NTc; ++e) {
                    double c = 0.0;
#pragma unroll
                    for (int dd = 0; dd < QZ; ++dd)
                        if (dd <= e) c = fma(r[L::OFF_Z + QY + dd], ri[Tri<Q>::rm(QY + dd, QY + e)], c);
                    coef[NJ * Q + slot * NTc + e] = make_double2(c, e == 0 ? base : 0.0);
                    const double ae = fabs(c) * xmax;
                    if (!(ae + (e == 0 ? fabs(base) : 0.0) <= 230.0)) bad = true;
                    amp += ae;
                }
                if (!(base - amp >= -667.0 && base + amp <= 667.0)) bad = true;
            }
        }
        bad = __any_sync(0xffffffffu, bad);
        __syncwarp();                              // coefficient stores -> table-building reads
        // The round loop reads mode and sqrt(2) R^-1 (and Z'y1) from the cluster's record in shared memory
        // instead of holding them in registers across the rounds: a few broadcast loads per trip buy
        // Q + NP (+ QY) double registers in a kernel that sits at its 128-register ceiling
        if (!PAIR && lane == 0) {
            double* scw = rec + NJ * W;
#pragma unroll
            for (int t = 0; t < NP; ++t) scw[SC_RI + t] = ri[t];
        }
        const volatile double* scv = sc;
        // (measured: pays for the larger tiles, q = 3 and the zero-inflated families; the q <= 2, 8-observation
        // kernels fit their registers and are 1.5 % faster with the values held there)
        constexpr bool NODE_SMEM = NJ > 8 || Q >= 3 || F::HAS_ZI;
        auto nd_m = [&](int dd) { return NODE_SMEM ? scv[SC_M + dd] : m[dd]; };
        auto nd_ri = [&](int t) { return NODE_SMEM ? scv[SC_RI + t] : ri[t]; };
        const double w = sc[SC_W];
        double zty[QY];
#pragma unroll
        for (int dd = 0; dd < QY; ++dd) zty[dd] = (F::LIN_ETA && !NODE_SMEM) ? sc[SC_ZTY + dd] : 0.0;
        // A cluster that leaves the sums (left to eval_kernel here, or dropped as non-finite below) takes
        // back its share of the node-independent terms, which const_sums_kernel / beta' X'y1 added for
        // every cluster of this variant; with ll_i the same terms complete the cluster's contribution.
        auto cluster_consts = [&](int ridx, double& s_ll, double& s_phi) {
            double2 cc = make_double2(0.0, 0.0);
            if (lane < nic) {
                const int64_t row = r0c + lane;
                cc = obs_consts_rare<F>(d.y1[row], F::USES_Y2 ? d.y2[row] : 0.0, rec[ridx * W + 2], A.fam, A.ytab);
            }
            s_ll = warp_sum(cc.x);
            s_phi = warp_sum(cc.y);
        };
        // two-pass tile: a cluster may span several records (tiles of 32 rows, pass A records then pass B records,
        // all in this warp's queue); it is left to eval_kernel as a whole when ANY of its tiles fails the range
        // test.  The state travels in shared memory: tp_st[0] shift, [1] 1 / sum_e, [2] w / sum_e or 0, [3] dead,
        // [4] ok, [5] log-likelihood of the cluster
        double* tp_st = zp_m + 2 * ((K + 1) & ~1);
        bool tp_dead = false;
        if (FL::TP) {
            tp_dead = (rflags & 2) ? false : tp_st[3] != 0.0;
            if ((rflags & 1) && bad) tp_dead = true;
            __syncwarp();
            if (lane == 0) {
                tp_st[3] = tp_dead ? 1.0 : 0.0;
                if (rflags & 4) { d.defer[i] = tp_dead ? 1 : 0; if (tp_dead && A.defer_count) atomicAdd(A.defer_count, 1u); }
            }
            __syncwarp();
            if (tp_dead && !(rflags & 8)) continue;       // its pass B records write the work arrays of their rows
        } else if (PAIR) {
            // (a failed range test in either cluster leaves both to eval_kernel)
            if (lane == 0) { d.defer[i] = bad ? 1 : 0; d.defer[i + 1] = bad ? 1 : 0; if (bad && A.defer_count) atomicAdd(A.defer_count, 2u); }
            if (bad) {
                double2 cc = make_double2(0.0, 0.0);
                double wrow = 0.0;
                if (lane < nic) {
                    const int64_t row = r0c + lane;
                    const double* r = rec + lane * W;
                    wrow = sco[SC_W];
                    cc = obs_consts_rare<F>(d.y1[row], F::USES_Y2 ? d.y2[row] : 0.0, r[2], A.fam, A.ytab);
                    double cll, cphi;
                    obs_const_tab<F>(d.y1[row], F::USES_Y2 ? d.y2[row] : 0.0, A.fam, A.ytab, cll, cphi);
                    d.xb[row] = r[2];
                    d.cll[row] = cll;
                    if (F::HAS_PHI) d.cphi[row] = cphi;
                }
                const double s_ll = warp_sum(wrow * cc.x), s_phi = warp_sum(wrow * cc.y);
                if (lane == 0) {
                    if (A.em_mode != 2) tots[0] -= s_ll;
                    if (F::HAS_PHI) tots[3] -= s_phi;
                }
                __syncwarp();
                continue;
            }
        } else if (lane == 0) { d.defer[i] = bad ? 1 : 0; if (bad && A.defer_count) atomicAdd(A.defer_count, 1u); }
        if (!PAIR && (FL::TP ? tp_dead : bad)) {
            // eval_kernel takes this cluster from the work arrays the prep kernel would have written
            double s_ll, s_phi;
            cluster_consts(lane, s_ll, s_phi);
            if (lane < nic) {
                const int64_t row = r0c + lane;
                const double* r = rec + lane * W;
                double cll, cphi;
                obs_const_tab<F>(d.y1[row], F::USES_Y2 ? d.y2[row] : 0.0, A.fam, A.ytab, cll, cphi);
                d.xb[row] = r[2];
                if (F::HAS_ZI) d.xg[row] = r[L::OFF_XG];
                d.cll[row] = cll;
                if (F::HAS_PHI) d.cphi[row] = cphi;
            }
            if (lane == 0) {
                if (A.em_mode != 2) tots[0] -= w * s_ll;
                if (F::HAS_PHI) tots[3] -= w * s_phi;
            }
            __syncwarp();
            continue;
        }
        if (ZP) {
            double tmp[W];
            if (lane < nic) {
#pragma unroll
                for (int t = 0; t < W; ++t) tmp[t] = rec[lane * W + t];
            }
            if (F::ZERO_FREE) {
                const bool zr = lane < nic && slot >= npos;
                hz0 = warp_sum(zr ? tmp[L::OFF_XG] : 0.0);
                if (QZ > 0) hz1 = warp_sum(zr ? tmp[L::OFF_Z + QY] : 0.0);
            }
            __syncwarp();
            if (lane < nic) {
#pragma unroll
                for (int t = 0; t < W; ++t) rec[slot * W + t] = tmp[t];
            }
            __syncwarp();
        }
        // tables: lanes stride the items, three independent exp chains per lane and trip (fast_exp_vec),
        // arguments in range by the test above.  Entry (observation j, dimension e, node index a) goes to
        // row e * nagq + a, column j.
        {
            const unsigned inv_nagq = (65536u + (unsigned)nagq - 1u) / (unsigned)nagq;
            // (a row-major walk -- consecutive lanes on consecutive columns of a row, conflict-free stores -- was
            // measured 2 % slower: its coefficient loads collide instead)
            if (F::NEED_EMU) {
                const int items = nic * Q * nagq;
                for (int t0 = lane; t0 < items; t0 += 96) {
                    double x3[3], o3[3];
                    int off[3];
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const int t = min(t0 + 32 * u, items - 1);
                        const int p = (int)(((unsigned)t * inv_nagq) >> 16);
                        const int na = t - p * nagq;
                        const double2 cb = coef[p];
                        x3[u] = fma(cb.x, gx[na], cb.y);
                        const int pj_ = p / Q, pe_ = p - pj_ * Q;
                        off[u] = (pe_ * nagq + na) * RP + pj_;
                    }
                    fast_exp_vec<3>(x3, cx, o3);
#pragma unroll
                    for (int u = 0; u < 3; ++u) tabF[off[u]] = o3[u];
                }
            }
            if (F::HAS_ZI) {
                const int nz = QZ > 0 ? nagq : 1;
                const unsigned inv_nz = QZ > 0 ? inv_nagq : 65536u;
                const int items = nic * NTc * nz;
                for (int t0 = lane; t0 < items; t0 += 96) {
                    double x3[3], o3[3];
                    int off[3];
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const int t = min(t0 + 32 * u, items - 1);
                        const int p = (int)(((unsigned)t * inv_nz) >> 16);
                        const int na = t - p * nz;
                        const double2 cb = coef[NJ * Q + p];
                        x3[u] = fma(cb.x, QZ > 0 ? gx[na] : 0.0, cb.y);
                        const int pj_ = p / NTc, pe_ = p - pj_ * NTc;
                        off[u] = (pe_ * nz + na) * RP + pj_;
                    }
                    fast_exp_vec<3>(x3, cx, o3);
#pragma unroll
                    for (int u = 0; u < 3; ++u) tabL[off[u]] = o3[u];
                }
            }
        }
        __syncwarp();
        if (ZT) {
            // zero-part pieces per (observation, zero-part node index): log(1 + lambda), plogis(eta_zi)
            const int nz = QZ > 0 ? nagq : 1;
            const unsigned inv_nz = QZ > 0 ? (65536u + (unsigned)nagq - 1u) / (unsigned)nagq : 65536u;
            for (int t = lane; t < nic * nz; t += 32) {
                const int j = (int)(((unsigned)t * inv_nz) >> 16);
                const int na = t - j * nz;
                double l1p, pi;
                fast_l1p_sigmoid<false>(tabL[na * RP + j], cx, l1p, pi);
                tabZ[j * NG + na] = make_double2(l1p, pi);
                if (ZP) zp_pi[j * nz + na] = pi;
            }
            __syncwarp();
            if (ZP) {
                if (lane < nz) {
                    double sl = 0.0;
                    for (int j = 0; j < nic; ++j) sl += tabZ[j * NG + lane].x;
                    zp_lz[lane] = sl;
                }
                __syncwarp();
            }
        }


        // FAST_KIND 3 (censored.normal): an uncensored observation contributes -t^2 / 2 with t linear in
        // the node, so the sum over the cluster's uncensored observations is ONE quadratic form in
        // delta_k = b_k - mode per node, sum_j (res_j - z_j' delta)^2 = U0 + delta' (U1 + U2 delta) with
        // res_j the residual at the mode: no per-point work at all for them -- their eta-scores are linear
        // in delta (posterior mean of delta) and their phi-scores are the same quadratic form.  Only the
        // censored observations are walked per node, from compact records (x_jk = A_j + C_j' delta_k).
        constexpr bool GQ = FL::GQ;
        constexpr int NPY = QY * (QY + 1) / 2;
        double* crec = tabF;
        double gq_u0 = 0.0, gq_u1[QY], gq_u2[NPY];
        int gq_nc = 0;
        if (GQ) {
            const bool live = lane < nic;
            const double* r = rec + lane * W;
            double resm = 0.0, zv[QY];
            bool cens = false;
#pragma unroll
            for (int dd = 0; dd < QY; ++dd) zv[dd] = live ? r[L::OFF_Z + dd] : 0.0;
            if (live) {
                resm = r[0] - r[2];
#pragma unroll
                for (int dd = 0; dd < QY; ++dd) resm = fma(-zv[dd], m[dd], resm);
                cens = r[1] != 0.0;
            }
            const unsigned cm = __ballot_sync(0xffffffffu, live && cens);
            gq_nc = __popc(cm);
            const bool unc = live && !cens;
            gq_u0 = warp_sum(unc ? resm * resm : 0.0);
            int t = 0;
#pragma unroll
            for (int dd = 0; dd < QY; ++dd) {
                gq_u1[dd] = -2.0 * warp_sum(unc ? resm * zv[dd] : 0.0);
#pragma unroll
                for (int ee = dd; ee < QY; ++ee) { gq_u2[t] = (ee == dd ? 1.0 : 2.0) * warp_sum(unc ? zv[dd] * zv[ee] : 0.0); ++t; }
            }
            if (live && cens) {
                const int slot = __popc(cm & ((1u << lane) - 1u));
                const double sg = r[1] == 1.0 ? 1.0 : -1.0;          // x = sg (y - eta) / sigma: left +, right -
                const double s = sg * A.fam.a[2];
                double* cr = crec + slot * FL::CRW;
                cr[0] = s * resm;
#pragma unroll
                for (int dd = 0; dd < QY; ++dd) cr[1 + dd] = -s * zv[dd];
                cr[1 + QY] = r[0];
                cr[2 + QY] = sg;
            }
            __syncwarp();
        }

        if constexpr (PAIR) {
            // ---- Pair tile (random intercepts: q = 1, K = nAGQ <= 16 nodes).  One cluster would use nAGQ of the
            // warp's 32 lanes for a single round; here lanes 0 .. 15 are the nodes of cluster i and lanes 16 .. 31
            // those of cluster i + 1.  The two clusters' rows are adjacent in memory, so they are staged, formed
            // and tabulated as ONE tile of n_A + n_B <= 16 observations (each observation lane with its own
            // cluster's mode and scale); a node lane then walks only its cluster's columns.  All reductions over
            // nodes stay inside a half warp (xor shuffles below 16), the exact maximum replaces the retry (one
            // round), and lanes 0 / 16 speak for their clusters in the totals.
            static_assert(Q == 1 && NJ == 16 && (F::FAST_KIND == 1 || F::FAST_KIND == 2), "pair tile: q = 1, carrier families");
            const int h = lane >> 4, kq = lane & 15;
            const bool valid = kq < K;
            const int kc = valid ? kq : K - 1;
            const double* sch = sc + h * FL::NSC;
            const double w_h = sch[SC_W];
            const int base = h ? nA : 0, cnt = h ? nic - nA : nA;
            const int64_t i_h = i + h;
            const int row_k = nidx[kc] & 0xff;                  // table row = node index (one dimension)
            const double b = fma(GLMMA_SQRT2 * sch[SC_RI], ntab[kc * 2], sch[SC_M]);
            double a = fma(-0.5 * A.prior.Dinv[0] * b, b, A.prior.logprior_const) + ntab[kc * 2 + 1];
            if (F::LIN_ETA) a = fma(sch[SC_ZTY], b, a);
            const double* pf0 = tabF + row_k * RP + base;       // this cluster's columns of the node's row
            const double* wvh = wv + base;
            double a1 = 0.0, vphi = 0.0, se8[8];
            const int cmax = max(nA, nic - nA);
#pragma unroll
            for (int c4 = 0; c4 < 2; ++c4) {
                if (4 * c4 >= cmax) {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) se8[4 * c4 + jj] = 0.0;
                    continue;
                }
                double t4[4];
                bool on[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = 4 * c4 + jj;
                    on[jj] = j < cnt;
                    const double f = pf0[on[jj] ? j : 0];
                    t4[jj] = F::FAST_KIND == 1 ? f + F::fast_addend(A.fam) : f;
                }
                if (F::FAST_KIND == 1) {
                    double L4[4], r4[4];
                    fast_log_rcp_vec<4>(t4, cx, L4, r4);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const double wj = on[jj] ? wvh[4 * c4 + jj] : 0.0;
                        if (jj & 1) a1 = fma(-wj, L4[jj], a1); else a = fma(-wj, L4[jj], a);
                        se8[4 * c4 + jj] = on[jj] ? r4[jj] : 0.0;
                        if (F::HAS_PHI) vphi += on[jj] ? L4[jj] : 0.0;
                    }
                } else {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const double ej = on[jj] ? t4[jj] : 0.0;
                        if (jj & 1) a1 -= ej; else a -= ej;
                        se8[4 * c4 + jj] = ej;
                    }
                }
            }
            a += a1;
            double mx = valid ? a : -INFINITY;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const double shift_h = fabs(mx) < INFINITY ? mx : 0.0;
            double e = valid ? fast_exp(a - shift_h, cx) : 0.0;
            if (A.em_mode) {
                if (A.em_mode == 2) e = valid ? A.pby[i_h * K + row_k] : 0.0;
                else if (valid) A.pby[i_h * K + row_k] = e;
            }
            double sum_h = e;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) sum_h += __shfl_xor_sync(0xffffffffu, sum_h, o);
            double lsum, isum;
            fast_log_rcp<true>(sum_h, cx, lsum, isum);
            const double ll = A.em_mode == 2 ? 0.0 : shift_h + lsum + sch[SC_LD];
            const bool ok = A.em_mode == 2 ? (sum_h > 0.0 && sum_h < INFINITY) : fabs(ll) < INFINITY;
            if (A.em_mode == 1 && !ok && valid) A.pby[i_h * K + row_k] = 0.0;
            const double inv = ok ? w_h * isum : 0.0;
            {
                // lanes 0 and 16 carry their cluster's terms; lane 0 adds both to the warp's totals
                double t0 = (kq == 0 && ok) ? w_h * ll : 0.0, t1 = (kq == 0 && ok) ? w_h : 0.0, t2 = (kq == 0 && !ok) ? 1.0 : 0.0;
                t0 += __shfl_down_sync(0xffffffffu, t0, 16);
                t1 += __shfl_down_sync(0xffffffffu, t1, 16);
                t2 += __shfl_down_sync(0xffffffffu, t2, 16);
                if (lane == 0) { tots[0] += t0; tots[1] += t1; tots[2] += t2; }
            }
            const bool okA = __shfl_sync(0xffffffffu, (int)ok, 0) != 0, okB = __shfl_sync(0xffffffffu, (int)ok, 16) != 0;
            if (!okA || !okB || d.ll_i != nullptr) {
                double2 cc = make_double2(0.0, 0.0);
                if (lane < nic) {
                    const int64_t row = r0c + lane;
                    cc = obs_consts_rare<F>(d.y1[row], F::USES_Y2 ? d.y2[row] : 0.0, rec[lane * W + 2], A.fam, A.ytab);
                }
                const bool inA = lane < nA;
                const double llA = warp_sum(inA ? cc.x : 0.0), llB = warp_sum(inA ? 0.0 : cc.x);
                const double phA = warp_sum(inA ? cc.y : 0.0), phB = warp_sum(inA ? 0.0 : cc.y);
                const double wA = __shfl_sync(0xffffffffu, w_h, 0), wB = __shfl_sync(0xffffffffu, w_h, 16);
                const double lA = __shfl_sync(0xffffffffu, ll, 0), lB = __shfl_sync(0xffffffffu, ll, 16);
                if (lane == 0) {
                    if (!okA) { if (A.em_mode != 2) tots[0] -= wA * llA; if (F::HAS_PHI) tots[3] -= wA * phA; }
                    if (!okB) { if (A.em_mode != 2) tots[0] -= wB * llB; if (F::HAS_PHI) tots[3] -= wB * phB; }
                    if (d.ll_i) { d.ll_i[i] = wA * (lA + llA); d.ll_i[i + 1] = wB * (lB + llB); }
                }
            }
            if (SCORE) {
                const double ebb = e * b * b;
                totl[2 * 32 + lane] = fma(inv, ebb, totl[2 * 32 + lane]);
                if (F::HAS_PHI) {
                    double gp = e * vphi;
                    if (F::FAST_KIND == 1) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j < cnt) gp = fma(wvh[j] * e, se8[j], gp);
                    }
                    totl[lane] = fma(inv * F::sp_scale(A.fam), gp, totl[lane]);
                }
                if (d.S_i) {
                    double sr = ebb;
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) sr += __shfl_xor_sync(0xffffffffu, sr, o);
                    if (kq == 0) d.S_i[i_h] = inv * sr;
                }
                // per-observation sums over the 16 node lanes of the observation's cluster
#pragma unroll
                for (int j = 0; j < 8; ++j) acc_eta[j * 32 + lane] = e * se8[j];
                __syncwarp();
                double s1 = 0.0;
                {
                    const int rw = lane & 7, sl = (lane >> 3) & 1;
#pragma unroll
                    for (int c = 0; c < 8; ++c) s1 += acc_eta[rw * 32 + 16 * h + 8 * sl + ((c + rw) & 7)];
                    s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
                }
                // observation lane `lane` (row `lane` of the tile): its cluster, its position there
                const int cl = (lane < nic && lane >= nA) ? 1 : 0;
                const int src = lane < nic ? 16 * cl + (lane - (cl ? nA : 0)) : 0;
                const double s1o = __shfl_sync(0xffffffffu, s1, src);
                const double isum_o = __shfl_sync(0xffffffffu, isum, 16 * cl);
                const double w_o = __shfl_sync(0xffffffffu, w_h, 16 * cl);
                const bool ok_o = __shfl_sync(0xffffffffu, (int)ok, 16 * cl) != 0;
                double zb = 0.0;
                if (lane < nic) {
                    const double scv1 = F::fast_score(s1o * isum_o, rec[lane * W], rec[lane * W + 1], A.fam);
                    zb = ok_o ? w_o * scv1 : 0.0;
                    if (A.zbar_out) d.zbar[r0c + lane] = zb;
                }
                __syncwarp();
                if (lane < NJ) acc_eta[lane] = zb;
                __syncwarp();
                if (lane < p_x) {
                    const double* xc = xcur + lane * NJ;
#pragma unroll
                    for (int jj = 0; jj < NJ; ++jj)
                        if (jj < nic) gcol = fma(xc[jj], acc_eta[jj], gcol);
                }
            }
            __syncwarp();
            continue;
        }
        double shift = 0.0, sum_e = 0.0, gphi = 0.0, amax = -INFINITY;
        double S[NP];
        double acc_e[NJ], acc_z[F::HAS_ZI ? NJ : 1];
        double gq_bb[QY];             // FAST_KIND 3: sum_k e_k delta_k
        constexpr bool TP = FL::TP;
        if constexpr (TP) {
            // ---- Two-pass form (FAST_KIND 1 / 2 families) for clusters of more than 16 observations.  The
            // one-pass form keeps an observation's score carrier in a register until the node weight is known:
            // 2 x NJ double registers, impossible beyond 16 observations.  Here pass A walks the observations
            // for the node terms only (log t, no reciprocal: a_k and the phi-score's sum_j log t_jk, kept per warp
            // in shared memory), the weights e_k = exp(a_k - max) follow exactly, and pass B walks the
            // observations again in chunks of eight for the carriers only (1 / t, no logarithm) with the weights
            // known, so a chunk's eight sums are plain register accumulators.  19 FP64 instructions per point
            // against 17 in one pass; the generic kernel, which this replaces, needs about 30.  A cluster of more
            // than 32 observations is a sequence of records in ONE warp's queue -- its tiles of 32 rows for pass
            // A, then the same tiles for pass B (staged and tabulated twice: Q nAGQ exponentials per observation
            // against nAGQ^Q points) -- with the node terms accumulating across the pass A records.
            static_assert(F::FAST_KIND == 1 || F::FAST_KIND == 2, "two-pass form: carrier families only");
            double* tp_a = zp_m;                          // K node terms, then the weights
            double* tp_v = tp_a + ((K + 1) & ~1);         // K sums of log t over the observations
            auto node = [&](int kc, double* b, const double** pf) {
                const int packed = nidx[kc];
                double nt[Q + 1];
                if (has_ntab) {
#pragma unroll
                    for (int e = 0; e <= Q; ++e) nt[e] = ntab[kc * (Q + 1) + e];
                } else {
                    nt[Q] = A.grid->lw_const;
#pragma unroll
                    for (int e = 0; e < Q; ++e) {
                        const int ae = ((packed >> (8 * e)) & 0xff) - e * nagq;
                        nt[e] = gx[ae];
                        nt[Q] += glw[ae];
                    }
                }
#pragma unroll
                for (int dd = 0; dd < Q; ++dd) {
                    double sdd = scv[SC_M + dd];
#pragma unroll
                    for (int e = dd; e < Q; ++e) sdd = fma(scv[SC_RI + Tri<Q>::rm(dd, e)], nt[e], sdd);
                    b[dd] = sdd;
                }
#pragma unroll
                for (int e = 0; e < Q; ++e) pf[e] = tabF + ((packed >> (8 * e)) & 0xff) * RP;
                return nt[Q];
            };
            // the chunk's four table products (+ addend); rows past the tile's end read stale columns
            auto prod4 = [&](const double** pf, int c4, double* t4) {
                double fq[Q][4];
#pragma unroll
                for (int e = 0; e < Q; ++e) {
                    const double2 f01 = *reinterpret_cast<const double2*>(pf[e] + 4 * c4);
                    const double2 f23 = *reinterpret_cast<const double2*>(pf[e] + 4 * c4 + 2);
                    fq[e][0] = f01.x; fq[e][1] = f01.y; fq[e][2] = f23.x; fq[e][3] = f23.y;
                }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    double pr = fq[0][jj];
#pragma unroll
                    for (int e = 1; e < Q - 1; ++e) pr *= fq[e][jj];
                    if (F::FAST_KIND == 1)
                        t4[jj] = Q > 1 ? fma(pr, fq[Q - 1][jj], F::fast_addend(A.fam)) : pr + F::fast_addend(A.fam);
                    else
                        t4[jj] = Q > 1 ? pr * fq[Q - 1][jj] : pr;
                    if (4 * c4 + jj >= nic) t4[jj] = 1.0;
                }
            };
            // ---- pass A over this record's rows
            if (rflags & 1) {
                const bool first = (rflags & 2) != 0;
                for (int r = 0; r < nrounds; ++r) {
                    const int k = (r << 5) + lane;
                    const bool valid = k < K;
                    const int kc = valid ? k : K - 1;
                    double b[Q];
                    const double* pf[Q];
                    const double lw = node(kc, b, pf);
                    double a, a1 = 0.0, vphi;
                    if (first) {
                        // b_k' D^-1 b_k, log weights and the linear-eta hoist enter once per cluster
                        double quad = 0.0;
#pragma unroll
                        for (int dd = 0; dd < Q; ++dd) {
                            double sdd = 0.0;
#pragma unroll
                            for (int e = 0; e < Q; ++e) sdd = fma(A.prior.Dinv[dd * Q + e], b[e], sdd);
                            quad = fma(b[dd], sdd, quad);
                        }
                        a = fma(-0.5, quad, A.prior.logprior_const) + lw;
                        if (F::LIN_ETA) {
#pragma unroll
                            for (int dd = 0; dd < QY; ++dd) a = fma(scv[SC_ZTY + dd], b[dd], a);
                        }
                        vphi = 0.0;
                    } else {
                        a = tp_a[kc];
                        vphi = tp_v[kc];
                    }
#pragma unroll 2
                    for (int c4 = 0; c4 < NJ / 4; ++c4) {
                        if (4 * c4 >= nic) break;
                        double t4[4];
                        prod4(pf, c4, t4);
                        if (F::FAST_KIND == 1) {
                            double L4[4], r4[4];
                            fast_log_rcp_vec<4>(t4, cx, L4, r4);             // the reciprocal half is dead code here
                            const double2 w01 = *reinterpret_cast<const double2*>(wv + 4 * c4);
                            const double2 w23 = *reinterpret_cast<const double2*>(wv + 4 * c4 + 2);
                            const double w4[4] = {w01.x, w01.y, w23.x, w23.y};
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const bool on = 4 * c4 + jj < nic;
                                const double wj = on ? w4[jj] : 0.0;
                                if (jj & 1) a1 = fma(-wj, L4[jj], a1); else a = fma(-wj, L4[jj], a);
                                if (F::HAS_PHI) vphi += on ? L4[jj] : 0.0;
                            }
                        } else {
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const double ej = 4 * c4 + jj < nic ? t4[jj] : 0.0;
                                if (jj & 1) a1 -= ej; else a -= ej;
                            }
                        }
                    }
                    if (valid) { tp_a[k] = a + a1; tp_v[k] = vphi; }
                }
                __syncwarp();
            }
            // ---- after the cluster's last tile: weights (exact maximum, no retry), cluster totals
            if (rflags & 4) {
                {
                    double mx = -INFINITY;
                    for (int k = lane; k < K; k += 32) mx = fmax(mx, tp_a[k]);
                    mx = warp_max(mx);
                    shift = fabs(mx) < INFINITY ? mx : 0.0;
                }
                sum_e = 0.0; gphi = 0.0;
#pragma unroll
                for (int t = 0; t < NP; ++t) S[t] = 0.0;
                for (int r = 0; r < nrounds; ++r) {
                    const int k = (r << 5) + lane;
                    const bool valid = k < K;
                    const int kc = valid ? k : K - 1;
                    double b[Q];
                    const double* pf[Q];
                    node(kc, b, pf);
                    double e = valid ? fast_exp(tp_a[kc] - shift, cx) : 0.0;
                    if (A.em_mode) {
                        const int packed = nidx[kc];
                        int kn = 0, mul = 1;
#pragma unroll
                        for (int ee = 0; ee < Q; ++ee) { kn += (int)(((packed >> (8 * ee)) & 0xff) - ee * nagq) * mul; mul *= nagq; }
                        if (A.em_mode == 2) e = valid ? A.pby[i * K + kn] : 0.0;
                        else if (valid) A.pby[i * K + kn] = e;
                    }
                    if (valid) tp_a[k] = e;
                    sum_e += e;
                    int t = 0;
#pragma unroll
                    for (int dd = 0; dd < Q; ++dd) {
                        const double eb = e * b[dd];
#pragma unroll
                        for (int ee = dd; ee < Q; ++ee) { S[t] = fma(eb, b[ee], S[t]); ++t; }
                    }
                    if (F::HAS_PHI) gphi = fma(e, valid ? tp_v[kc] : 0.0, gphi);
                }
                sum_e = warp_sum(sum_e);
                double lsum, isum;
                fast_log_rcp<true>(sum_e, cx, lsum, isum);
                const double ll = A.em_mode == 2 ? 0.0 : shift + lsum + sc[SC_LD];
                const bool ok = A.em_mode == 2 ? (sum_e > 0.0 && sum_e < INFINITY) : fabs(ll) < INFINITY;
                if (A.em_mode == 1 && !ok)
                    for (int k = lane; k < K; k += 32) A.pby[i * K + k] = 0.0;
                const double inv = ok ? w * isum : 0.0;
                if (lane == 0) {
                    if (ok) { tots[0] += w * ll; tots[1] += w; } else tots[2] += 1.0;
                    if (d.ll_i) d.ll_i[i] = w * ll;          // the pass B records add their rows' node-independent terms
                    tp_st[1] = isum; tp_st[2] = inv; tp_st[4] = ok ? 1.0 : 0.0;
                }
                if (SCORE) {
                    if (F::HAS_PHI) totl[lane] = fma(inv * F::sp_scale(A.fam), gphi, totl[lane]);
#pragma unroll
                    for (int t = 0; t < NP; ++t) totl[(2 + t) * 32 + lane] = fma(inv, S[t], totl[(2 + t) * 32 + lane]);
                    if (d.S_i) {
                        double Sr[NP];
#pragma unroll
                        for (int t = 0; t < NP; ++t) Sr[t] = inv * warp_sum(S[t]);
                        if (lane == 0) {
                            int t = 0;
#pragma unroll
                            for (int dd = 0; dd < Q; ++dd)
#pragma unroll
                                for (int ee = dd; ee < Q; ++ee) {
                                    d.S_i[i * Q * Q + dd * Q + ee] = Sr[t];
                                    d.S_i[i * Q * Q + ee * Q + dd] = Sr[t];
                                    ++t;
                                }
                        }
                    }
                }
                __syncwarp();
            }
            // ---- pass B over this record's rows: chunks of eight observations, carriers only
            if (rflags & 8) {
                const double isum = tp_st[1], inv = tp_st[2];
                const bool ok = tp_st[4] != 0.0;
                if (!ok || d.ll_i != nullptr) {
                    // node-independent terms of this record's rows: taken back when the cluster dropped out of the
                    // sums, added to the cluster's contribution when those are wanted
                    double s_ll, s_phi;
                    cluster_consts(lane, s_ll, s_phi);
                    if (lane == 0) {
                        if (!ok) {
                            if (A.em_mode != 2) tots[0] -= w * s_ll;
                            if (F::HAS_PHI) tots[3] -= w * s_phi;
                        }
                        if (d.ll_i) d.ll_i[i] += w * s_ll;
                    }
                }
                double tp_s1 = 0.0;              // this lane's observation's sum_k e_k carrier_jk
                if (SCORE) {
                    for (int c8 = 0; c8 < NJ / 8; ++c8) {
                        if (8 * c8 >= nic) break;
                        double acc8[8];
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) acc8[jj] = 0.0;
                        for (int r = 0; r < nrounds; ++r) {
                            const int k = (r << 5) + lane;
                            const bool valid = k < K;
                            const int kc = valid ? k : K - 1;
                            const int packed = nidx[kc];
                            const double* pf[Q];
#pragma unroll
                            for (int e = 0; e < Q; ++e) pf[e] = tabF + ((packed >> (8 * e)) & 0xff) * RP;
                            const double e = valid ? tp_a[kc] : 0.0;
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int c4 = 2 * c8 + h;
                                double t4[4];
                                prod4(pf, c4, t4);
                                if (F::FAST_KIND == 1) {
                                    double L4[4], r4[4];
                                    fast_log_rcp_vec<4>(t4, cx, L4, r4);         // the logarithm half is dead code here
#pragma unroll
                                    for (int jj = 0; jj < 4; ++jj) acc8[4 * h + jj] = fma(e, 4 * c4 + jj < nic ? r4[jj] : 0.0, acc8[4 * h + jj]);
                                } else {
#pragma unroll
                                    for (int jj = 0; jj < 4; ++jj) acc8[4 * h + jj] = fma(e, 4 * c4 + jj < nic ? t4[jj] : 0.0, acc8[4 * h + jj]);
                                }
                            }
                        }
                        // the chunk's eight sums over the 32 lanes (as the one-pass epilogue does for a tile of eight)
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) acc_eta[jj * 32 + lane] = acc8[jj];
                        __syncwarp();
                        {
                            const int j = lane & 7, sl = lane >> 3;
                            double s1 = 0.0;
#pragma unroll
                            for (int c = 0; c < 8; ++c) s1 += acc_eta[j * 32 + sl * 8 + ((c + j) & 7)];
                            s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
                            s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                            if ((lane >> 3) == c8) tp_s1 = s1;          // lane l owns observation l = 8 c8 + (l & 7)
                        }
                        __syncwarp();
                    }
                    // second term of the phi-score of these rows: sum_j rec1_j sum_k e_k r_jk
                    if (F::HAS_PHI && F::FAST_KIND == 1 && lane < nic)
                        totl[lane] = fma(inv * F::sp_scale(A.fam), wv[lane] * tp_s1, totl[lane]);
                    double zb = 0.0;
                    if (lane < nic) {
                        const double scv1 = F::fast_score(tp_s1 * isum, rec[lane * W], rec[lane * W + 1], A.fam);
                        zb = ok ? w * scv1 : 0.0;
                        if (A.zbar_out) d.zbar[r0c + lane] = zb;
                    }
                    __syncwarp();
                    acc_eta[lane] = zb;
                    __syncwarp();
                    if (lane < p_x) {
                        const double* xc = xcur + lane * NJ;
#pragma unroll 8
                        for (int jj = 0; jj < NJ; ++jj)
                            if (jj < nic) gcol = fma(xc[jj], acc_eta[jj], gcol);
                    }
                }
            }
            __syncwarp();
            continue;
        } else {
        for (int attempt = 0; attempt < 2; ++attempt) {
            sum_e = 0.0; gphi = 0.0;
#pragma unroll
            for (int t = 0; t < NP; ++t) S[t] = 0.0;
#pragma unroll
            for (int j = 0; j < NJ; ++j) { acc_e[j] = 0.0; if (F::HAS_ZI) acc_z[j] = 0.0; }
            if (ZP && SCORE) {
                for (int a_ = 0; a_ < (QZ > 0 ? nagq : 1); ++a_) zp_m[a_ * 32 + lane] = 0.0;
            }
            if (GQ) {
                for (int c = 0; c < gq_nc; ++c) acc_eta[c * 32 + lane] = 0.0;
#pragma unroll
                for (int dd = 0; dd < QY; ++dd) gq_bb[dd] = 0.0;
            }
            for (int rr = 0; rr < nrounds; ++rr) {
                const int r = (rr == 0) ? rc : (rr <= rc ? rr - 1 : rr);
                const int k = (r << 5) + lane;
                const bool valid = k < K;
                const int kc = valid ? k : K - 1;
                // node: b_k = mode + sqrt(2) R^-1 z_k, a = log N(b_k; 0, D) + log wGH_k
                const int packed = nidx[kc];
                double nt[Q + 1];
                if (has_ntab) {
#pragma unroll
                    for (int e = 0; e <= Q; ++e) nt[e] = ntab[kc * (Q + 1) + e];
                } else {
                    nt[Q] = A.grid->lw_const;
#pragma unroll
                    for (int e = 0; e < Q; ++e) {
                        const int ae = ((packed >> (8 * e)) & 0xff) - e * nagq;
                        nt[e] = gx[ae];
                        nt[Q] += glw[ae];
                    }
                }
                double b[Q], dl[GQ ? Q : 1];
#pragma unroll
                for (int dd = 0; dd < Q; ++dd) {
                    double s = GQ ? 0.0 : nd_m(dd);
#pragma unroll
                    for (int e = dd; e < Q; ++e) s = fma(nd_ri(Tri<Q>::rm(dd, e)), nt[e], s);
                    if (GQ) { dl[dd] = s; s += nd_m(dd); }
                    b[dd] = s;
                }
                double quad = 0.0;
#pragma unroll
                for (int dd = 0; dd < Q; ++dd) {
                    double s = 0.0;
#pragma unroll
                    for (int e = 0; e < Q; ++e) s = fma(A.prior.Dinv[dd * Q + e], b[e], s);
                    quad = fma(b[dd], s, quad);
                }
                double a = fma(-0.5, quad, A.prior.logprior_const) + nt[Q];
                const double* pf[Q];
#pragma unroll
                for (int e = 0; e < Q; ++e) pf[e] = tabF + ((packed >> (8 * e)) & 0xff) * RP;
                const double* pl[QZ > 0 ? QZ : 1];
#pragma unroll
                for (int e = 0; e < (QZ > 0 ? QZ : 1); ++e)
                    pl[e] = tabL + (QZ > 0 ? (int)((packed >> (8 * (QY + e))) & 0xff) - QY * nagq : 0) * RP;
                const double2* pz = tabZ + (QZ > 0 ? (int)((packed >> (8 * QY)) & 0xff) - QY * nagq : 0);
                double vphi = 0.0;
                double se_[NJ], sz_[F::HAS_ZI ? NJ : 1];
                if (F::LIN_ETA) {
                    // sum_j y1_j eta_jk = sum_j y1_j xb_j + (Z'y1) . b_k: the first term does not depend on
                    // the node -- it is added to the log-likelihood once per evaluation (beta' X'y1)
#pragma unroll
                    for (int dd = 0; dd < QY; ++dd) a = fma(NODE_SMEM ? scv[SC_ZTY + dd] : zty[dd], b[dd], a);
                }
                const int zp_az = (ZP && QZ > 0) ? (int)((packed >> (8 * QY)) & 0xff) - QY * nagq : 0;
                if (ZP) {
                    a -= zp_lz[zp_az];
                    if (F::ZERO_FREE) a += QZ > 0 ? fma(hz1, b[QY < Q ? QY : 0], hz0) : hz0;
                }
#define GLMMA_FAST_OBS(j)                                                                              \
    {                                                                                                  \
        const double* rj = rec + (j) * W;                                                              \
        double eta = rj[2];                                                                            \
        _Pragma("unroll") for (int dd = 0; dd < QY; ++dd) eta = fma(rj[L::OFF_Z + dd], b[dd], eta);     \
        double ez = 0.0, emu = 0.0, elam = 0.0;                                                        \
        if (F::NEED_EMU) {                                                                             \
            emu = pf[0][(j)];                                                                 \
            _Pragma("unroll") for (int e = 1; e < Q; ++e) emu *= pf[e][(j)];                   \
        }                                                                                              \
        if (F::HAS_ZI) {                                                                               \
            ez = rj[L::OFF_XG];                                                                        \
            _Pragma("unroll") for (int dd = 0; dd < QZ; ++dd)                                           \
                ez = fma(rj[L::OFF_Z + QY + dd], b[QY + dd], ez);                                      \
            elam = pl[0][(j)];                                                                \
            _Pragma("unroll") for (int e = 1; e < QZ; ++e) elam *= pl[e][(j)];                 \
        }                                                                                              \
        double ld, se, sz, sp;                                                                         \
        if (ZT) {                                                                                      \
            const double2 zt = pz[(j) * NG];                                                           \
            F::template point<SCORE, false, true>(rj[0], rj[1], eta, ez, emu, elam, A.fam, cx, ld, se, sz, sp, zt.x, zt.y); \
        } else {                                                                                       \
            F::template point<SCORE, false>(rj[0], rj[1], eta, ez, emu, elam, A.fam, cx, ld, se, sz, sp); \
        }                                                                                              \
        a += ld;                                                                                       \
        if (SCORE) {                                                                                   \
            se_[j] = se;                                                                               \
            if (F::HAS_ZI) sz_[j] = sz;                                                                \
            if (F::HAS_PHI) vphi += sp;                                                                \
        }                                                                                              \
    }
                if (GQ) {
                    const double isig = A.fam.a[2], isig2 = isig * isig;
                    double qk = gq_u0;
                    {
                        int t = 0;
#pragma unroll
                        for (int dd = 0; dd < QY; ++dd) {
                            double s = gq_u1[dd];
#pragma unroll
                            for (int ee = dd; ee < QY; ++ee) { s = fma(gq_u2[t], dl[ee], s); ++t; }
                            qk = fma(dl[dd], s, qk);
                        }
                    }
                    a = fma(-0.5 * isig2, qk, a);
                    vphi = isig2 * qk;
#pragma unroll 1
                    for (int c = 0; c < gq_nc; ++c) {
                        const double* cr = crec + c * FL::CRW;
                        double x = cr[0];
#pragma unroll
                        for (int dd = 0; dd < QY; ++dd) x = fma(cr[1 + dd], dl[dd], x);
                        double ldc, lam, dn, B;
                        fast_gauss_tail(x, cx, ldc, lam, dn, B);
                        double cv = lam, spv = -x * lam;
                        if (B < GLMMA_SQRT_DBL_EPS) {
                            // the reference's sqrt(eps) floors (R/Fit_Funs.R:1395, 1413, 1421): rare
                            const double Bf = GLMMA_SQRT_DBL_EPS;
                            if (cr[2 + QY] > 0.0) {
                                spv = -x * dn / Bf;
                            } else {
                                const double eta = fma(x, A.fam.a[0], cr[1 + QY]);
                                cv = (A.fam.a[0] * dn - eta * (Bf - B)) / Bf * isig;
                                spv = (B - Bf - x * dn) / Bf;
                            }
                        }
                        a += ldc;
                        vphi += spv;
                        acc_zi[c * 32 + lane] = cv;
                    }
                } else if (F::FAST_KIND != 0) {
                    // stage-wise chunks of four observations (interleaved chains): separable
                    // product, (log t, 1/t) for the chunk, accumulation.  Two alternating
                    // accumulators keep the sum over observations off the critical path.
                    double a1 = 0.0;
                    auto chunk = [&](const int c4, auto full_tag) {
                        constexpr bool FULL = decltype(full_tag)::value;
                        double t4[4], L4[4], r4[4];
                        // the chunk's factors: the tables keep the observations of one (dimension, node
                        // index) contiguous, so two 16-byte loads per table serve four observations
                        double fq[Q][4];
#pragma unroll
                        for (int e = 0; e < Q; ++e) {
                            const double2 f01 = *reinterpret_cast<const double2*>(pf[e] + 4 * c4);
                            const double2 f23 = *reinterpret_cast<const double2*>(pf[e] + 4 * c4 + 2);
                            fq[e][0] = f01.x; fq[e][1] = f01.y; fq[e][2] = f23.x; fq[e][3] = f23.y;
                        }
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int j = 4 * c4 + jj;
                            double pr = fq[0][jj];
#pragma unroll
                            for (int e = 1; e < Q - 1; ++e) pr *= fq[e][jj];
                            if (F::FAST_KIND == 1)
                                t4[jj] = Q > 1 ? fma(pr, fq[Q - 1][jj], F::fast_addend(A.fam)) : pr + F::fast_addend(A.fam);
                            else
                                t4[jj] = Q > 1 ? pr * fq[Q - 1][jj] : pr;
                            // rows past the cluster's end hold stale table entries: keep the argument of
                            // the unchecked log in range (their results are discarded below)
                            if (!FULL && j >= nic) t4[jj] = 1.0;
                        }
                        if (F::FAST_KIND == 1) {
                            fast_log_rcp_vec<4>(t4, cx, L4, r4);
                            const double2 w01 = *reinterpret_cast<const double2*>(wv + 4 * c4);
                            const double2 w23 = *reinterpret_cast<const double2*>(wv + 4 * c4 + 2);
                            const double w4[4] = {w01.x, w01.y, w23.x, w23.y};
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const int j = 4 * c4 + jj;
                                const bool on = FULL || j < nic;
                                const double wj = on ? w4[jj] : 0.0;
                                if (jj & 1) a1 = fma(-wj, L4[jj], a1); else a = fma(-wj, L4[jj], a);
                                if (SCORE) {
                                    se_[j] = on ? r4[jj] : 0.0;
                                    if (F::HAS_PHI) vphi += on ? L4[jj] : 0.0;
                                }
                            }
                        } else {
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const int j = 4 * c4 + jj;
                                const double ej = (FULL || j < nic) ? t4[jj] : 0.0;
                                if (jj & 1) a1 -= ej; else a -= ej;
                                if (SCORE) se_[j] = ej;
                            }
                        }
                    };
                    if (nic == NJ) {
                        // complete cluster: one branch-free block, so the chunks' chains interleave too
#pragma unroll
                        for (int c4 = 0; c4 < NJ / 4; ++c4) chunk(c4, std::true_type{});
                    } else {
#pragma unroll
                        for (int c4 = 0; c4 < NJ / 4; ++c4) {
                            if (nic >= 4 * c4 + 4) {
                                chunk(c4, std::true_type{});
                            } else if (nic > 4 * c4) {
                                chunk(c4, std::false_type{});
                            } else {
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj) se_[4 * c4 + jj] = 0.0;
                            }
                        }
                    }
                    a += a1;
                } else if (ZP) {
                    // class known from the slot; the zero part's shared terms were added at node level
                    auto obs = [&](auto j_tag, auto yc_tag) {
                        constexpr int j = decltype(j_tag)::value;
                        constexpr int YC = decltype(yc_tag)::value;
                        const double* rj = rec + j * W;
                        double eta = rj[2];
#pragma unroll
                        for (int dd = 0; dd < QY; ++dd) eta = fma(rj[L::OFF_Z + dd], b[dd], eta);
                        double emu = 0.0, elam = 0.0;
                        if (F::NEED_EMU) {
                            emu = pf[0][j];
#pragma unroll
                            for (int e = 1; e < Q; ++e) emu *= pf[e][j];
                        }
                        if (YC == 2) {
                            elam = pl[0][j];
#pragma unroll
                            for (int e = 1; e < QZ; ++e) elam *= pl[e][j];
                        }
                        double ld, se, sz, sp;
                        F::template point<SCORE, false, true, YC>(rj[0], rj[1], eta, 0.0, emu, elam, A.fam, cx, ld, se, sz, sp);
                        a += ld;
                        if (SCORE) {
                            se_[j] = se;
                            if (YC == 2) sz_[j] = sz;
                            if (F::HAS_PHI) vphi += sp;
                        }
                    };
                    auto obs_j = [&](auto j_tag) {
                        constexpr int j = decltype(j_tag)::value;
                        se_[j] = 0.0;
                        if (j < npos) obs(j_tag, std::integral_constant<int, 1>{});
                    };
                    // zero responses (zero-inflated families): ONE copy of the long body, walked by a rolled loop;
                    // the two score terms wait in shared memory (the accumulator scratch, idle during the rounds)
                    // until the node weight is known
                    if (!F::ZERO_FREE) {
#pragma unroll 1
                        for (int j = npos; j < nic; ++j) {
                            const double* rj = rec + j * W;
                            double emu = 0.0, elam;
                            if (F::NEED_EMU) {
                                emu = pf[0][j];
#pragma unroll
                                for (int e = 1; e < Q; ++e) emu *= pf[e][j];
                            }
                            elam = pl[0][j];
                            double ld, se, sz, sp;
                            F::template point<SCORE, false, true, 2>(rj[0], rj[1], 0.0, 0.0, emu, elam, A.fam, cx, ld, se, sz, sp);
                            a += ld;
                            if (SCORE) {
                                acc_eta[(j - npos) * 64 + lane] = se;
                                acc_eta[(j - npos) * 64 + 32 + lane] = sz;
                                if (F::HAS_PHI) vphi += sp;
                            }
                        }
                    }
                    obs_j(std::integral_constant<int, 0>{}); obs_j(std::integral_constant<int, 1>{});
                    obs_j(std::integral_constant<int, 2>{}); obs_j(std::integral_constant<int, 3>{});
                    obs_j(std::integral_constant<int, 4>{}); obs_j(std::integral_constant<int, 5>{});
                    obs_j(std::integral_constant<int, 6>{}); obs_j(std::integral_constant<int, 7>{});
                    if (NJ > 8) {
                        obs_j(std::integral_constant<int, (NJ > 8 ? 8 : 0)>{}); obs_j(std::integral_constant<int, (NJ > 8 ? 9 : 0)>{});
                        obs_j(std::integral_constant<int, (NJ > 8 ? 10 : 0)>{}); obs_j(std::integral_constant<int, (NJ > 8 ? 11 : 0)>{});
                    }
                    if (NJ > 12) {
                        obs_j(std::integral_constant<int, (NJ > 12 ? 12 : 0)>{}); obs_j(std::integral_constant<int, (NJ > 12 ? 13 : 0)>{});
                        obs_j(std::integral_constant<int, (NJ > 12 ? 14 : 0)>{}); obs_j(std::integral_constant<int, (NJ > 12 ? 15 : 0)>{});
                    }
                } else {
#pragma unroll
                for (int c4 = 0; c4 < NJ / 4; ++c4) {
                    if (nic >= 4 * c4 + 4) {
                        // full chunk: four independent point evaluations, no branches
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) GLMMA_FAST_OBS(4 * c4 + jj)
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            se_[4 * c4 + jj] = 0.0;
                            if (F::HAS_ZI) sz_[4 * c4 + jj] = 0.0;
                            if (4 * c4 + jj < nic) GLMMA_FAST_OBS(4 * c4 + jj)
                        }
                    }
                }
                }
#undef GLMMA_FAST_OBS
                if (valid) amax = fmax(amax, a);
                if (rr == 0 && attempt == 0) {
                    // any value near the maximum will do as the shift: single precision (a NaN or an
                    // infinite a gives shift 0; the retry below then takes the true maximum)
                    float sf = valid ? (float)a : -INFINITY;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) sf = fmaxf(sf, __shfl_xor_sync(0xffffffffu, sf, o));
                    shift = fabsf(sf) < INFINITY ? (double)sf : 0.0;
                }
                double e = valid ? fast_exp(a - shift, cx) : 0.0;
                if (A.em_mode) {             // EM phase: store (E-step) / replace (M-step) the node weights
                    // (p_by is kept in the natural node order: eval_kernel may own this cluster in another pass)
                    int kn = 0, mul = 1;
#pragma unroll
                    for (int ee = 0; ee < Q; ++ee) { kn += (int)(((packed >> (8 * ee)) & 0xff) - ee * nagq) * mul; mul *= nagq; }
                    if (A.em_mode == 2) e = valid ? A.pby[i * K + kn] : 0.0;
                    else if (valid) A.pby[i * K + kn] = e;
                }
                sum_e += e;
                if (SCORE) {
                    int t = 0;
#pragma unroll
                    for (int dd = 0; dd < Q; ++dd) {
                        const double eb = e * b[dd];
#pragma unroll
                        for (int ee = dd; ee < Q; ++ee) { S[t] = fma(eb, b[ee], S[t]); ++t; }
                    }
                    if (F::HAS_PHI) gphi = fma(e, vphi, gphi);
                    if (GQ) {
#pragma unroll 1
                        for (int c = 0; c < gq_nc; ++c) acc_eta[c * 32 + lane] = fma(e, acc_zi[c * 32 + lane], acc_eta[c * 32 + lane]);
#pragma unroll
                        for (int dd = 0; dd < QY; ++dd) gq_bb[dd] = fma(e, dl[dd], gq_bb[dd]);
                    } else if (ZP) {
                        zp_m[zp_az * 32 + lane] += e;
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            if (F::ZERO_FREE) {
                                acc_e[j] = fma(e, se_[j], acc_e[j]);
                            } else if (j < npos) {
                                acc_e[j] = fma(e, se_[j], acc_e[j]);
                            } else if (j < nic) {
                                acc_e[j] = fma(e, acc_eta[(j - npos) * 64 + lane], acc_e[j]);
                                acc_z[j] = fma(e, acc_eta[(j - npos) * 64 + 32 + lane], acc_z[j]);
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            acc_e[j] = fma(e, se_[j], acc_e[j]);
                            if (F::HAS_ZI) acc_z[j] = fma(e, sz_[j], acc_z[j]);
                        }
                    }
                }
            }
            sum_e = warp_sum(sum_e);
            if ((sum_e > 0.0 && sum_e < INFINITY) || A.em_mode == 2) break;
            const double mx = warp_max(amax);
            if (attempt == 1 || !(fabs(mx) < INFINITY) || mx == shift) break;
            shift = mx;
        }
        }
        double lsum, isum;
        fast_log_rcp<true>(sum_e, cx, lsum, isum);
        // node-independent terms are not part of ll here (added once per evaluation, see the header)
        const double ll = A.em_mode == 2 ? 0.0 : shift + lsum + sc[SC_LD];
        const bool ok = A.em_mode == 2 ? (sum_e > 0.0 && sum_e < INFINITY) : fabs(ll) < INFINITY;
        if (A.em_mode == 1 && !ok)
            for (int k = lane; k < K; k += 32) A.pby[i * K + k] = 0.0;
        if (lane == 0) {
            if (ok) { tots[0] += w * ll; tots[1] += w; } else tots[2] += 1.0;
        }
        if (!ok || d.ll_i != nullptr) {
            double s_ll, s_phi;
            cluster_consts(slot, s_ll, s_phi);
            if (lane == 0) {
                if (!ok) {
                    if (A.em_mode != 2) tots[0] -= w * s_ll;
                    if (F::HAS_PHI) tots[3] -= w * s_phi;
                }
                if (d.ll_i) d.ll_i[i] = w * (ll + s_ll);
            }
        }
        if (SCORE) {
            const double inv = ok ? w * isum : 0.0;
            if (F::HAS_PHI) {
                if (F::FAST_KIND == 1) {
                    // second term of the phi-score, sum_k e_k sum_j rec1_j r_jk, from the carrier sums
#pragma unroll
                    for (int j = 0; j < NJ; ++j)
                        if (j < nic) gphi = fma(wv[j], acc_e[j], gphi);
                }
                totl[lane] = fma(inv * F::sp_scale(A.fam), gphi, totl[lane]);
            }
#pragma unroll
            for (int t = 0; t < NP; ++t) totl[(2 + t) * 32 + lane] = fma(inv, S[t], totl[(2 + t) * 32 + lane]);
            if (d.S_i) {
                double Sr[NP];
#pragma unroll
                for (int t = 0; t < NP; ++t) Sr[t] = inv * warp_sum(S[t]);
                if (lane == 0) {
                    int t = 0;
#pragma unroll
                    for (int dd = 0; dd < Q; ++dd)
#pragma unroll
                        for (int ee = dd; ee < Q; ++ee) {
                            d.S_i[i * Q * Q + dd * Q + ee] = Sr[t];
                            d.S_i[i * Q * Q + ee * Q + dd] = Sr[t];
                            ++t;
                        }
                }
            }
            // per-observation sums over the 32 lanes: through shared memory; lane l sums the
            // (l / NJR)-th slice of columns of observation l % NJR, slices combined by shuffles
            // ZP mode: posterior mean of plogis(eta_zi) per observation from the marginal node weights of the
            // zero-part dimension (before the accumulator stores below overwrite the zero-part table)
            double zp_epi = 0.0;
            if (ZP) {
                const int nz = QZ > 0 ? nagq : 1;
                double ma = 0.0;
                if (lane < nz) {
#pragma unroll 8
                    for (int c = 0; c < 32; ++c) ma += zp_m[lane * 32 + ((c + lane) & 31)];
                }
                for (int a_ = 0; a_ < nz; ++a_) {
                    const double mv = __shfl_sync(0xffffffffu, ma, a_);
                    if (lane < nic) zp_epi = fma(zp_pi[slot * nz + a_], mv, zp_epi);
                }
                zp_epi *= isum;
                __syncwarp();
            }
            if (!GQ) {
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    acc_eta[j * 32 + lane] = acc_e[j];
                    if (F::HAS_ZI) acc_zi[j * 32 + lane] = acc_z[j];
                }
            }
            __syncwarp();
            {
                constexpr int NJR = FL::NJR;          // rows NJ .. NJR - 1 (tile 12) hold stale data: their lanes write nothing
                constexpr int NSL = 32 / NJR;         // slices (4 for 8 rows, 2 for 16)
                constexpr int CW = 32 / NSL;          // columns per slice
                const int j = lane % NJR, sl = lane / NJR;
                double s1 = 0.0, s2 = 0.0;
#pragma unroll
                for (int c = 0; c < CW; ++c) {
                    const int cc = sl * CW + ((c + j) & (CW - 1));
                    s1 += acc_eta[j * 32 + cc];
                    if (F::HAS_ZI) s2 += acc_zi[j * 32 + cc];
                }
#pragma unroll
                for (int o = NJR; o < 32; o <<= 1) {
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                    if (F::HAS_ZI) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                }
                double zb = 0.0, zz = 0.0;
                double scv_gq = 0.0;
                if (GQ) {
                    // censored observation in compact slot c: score = -sign / sigma * mean(carrier_c), the row sum
                    // held by lane c; uncensored: (res - z' mean(delta)) / sigma^2
                    const bool live = lane < nic;
                    const double* r = rec + lane * W;
                    const bool cens = live && r[1] != 0.0;
                    const unsigned cm = __ballot_sync(0xffffffffu, cens);
                    const int cslot = __popc(cm & ((1u << lane) - 1u));
                    const double sl = __shfl_sync(0xffffffffu, s1, cens ? cslot : 0);
                    double resd = live ? r[0] - r[2] : 0.0;
#pragma unroll
                    for (int dd = 0; dd < QY; ++dd) {
                        const double bbar = warp_sum(gq_bb[dd]) * isum;
                        if (live) resd = fma(-r[L::OFF_Z + dd], nd_m(dd) + bbar, resd);
                    }
                    const double isig = A.fam.a[2];
                    scv_gq = cens ? (r[1] == 1.0 ? -isig : isig) * (sl * isum) : isig * isig * resd;
                }
                if (ZP) {
                    // row sums are held by slot: fetch this observation's
                    s1 = __shfl_sync(0xffffffffu, s1, slot);
                    s2 = __shfl_sync(0xffffffffu, s2, slot);
                }
                if (lane < nic) {
                    const double scv = GQ ? scv_gq
                                     : F::FAST_KIND != 0 ? F::fast_score(s1 * isum, rec[slot * W], rec[slot * W + 1], A.fam)
                                                         : F::finish_se(s1 * isum, rec[slot * W], rec[slot * W + 1], A.fam);
                    zb = ok ? w * scv : 0.0;
                    if (F::HAS_ZI) zz = inv * s2;
                    // ZP: the hoisted -plogis(eta_zi) (and the hurdle families' +1 for a zero response)
                    if (ZP && ok) zz += w * ((F::ZERO_FREE && slot >= npos ? 1.0 : 0.0) - zp_epi);
                    if (A.zbar_out) {
                        d.zbar[r0c + lane] = zb;
                        if (F::HAS_ZI) d.zbar_zi[r0c + lane] = zz;
                    }
                }
                // X'zbar (and X_zi'zbar_zi): lane l owns column l; the posterior-mean scores of the
                // cluster's observations pass through the (now free) accumulator scratch
                __syncwarp();
                if (lane < NJ) {
                    acc_eta[lane] = zb;
                    if (F::HAS_ZI) acc_zi[lane] = zz;
                }
                __syncwarp();
                if (lane < p_x + p_z) {
                    const double* xc = xcur + lane * NJ;
                    const double* zs = (F::HAS_ZI && lane >= p_x) ? acc_zi : acc_eta;
#pragma unroll
                    for (int jj = 0; jj < NJ; ++jj)
                        if (jj < nic) gcol = fma(xc[jj], zs[jj], gcol);
                }
            }
        }
        __syncwarp();
    }

    // block partials, fixed order (deterministic for a fixed launch shape)
    __syncwarp();
    {
        // per-warp totals: [0] loglik [1] sum_w [2] n_nonfinite [3] g_phi [4..] S
        const double v3 = warp_sum(totl[lane]);
        double vs[NP];
#pragma unroll
        for (int t = 0; t < NP; ++t) vs[t] = warp_sum(totl[(2 + t) * 32 + lane]);
        __syncthreads();
        double* red = smem;    // wpb * PSTRIDE doubles (fastmath tables are no longer needed)
        if (lane == 0) {
            red[wib * GLMMA_PSTRIDE + 0] = tots[0];
            red[wib * GLMMA_PSTRIDE + 1] = tots[1];
            red[wib * GLMMA_PSTRIDE + 2] = tots[2];
            red[wib * GLMMA_PSTRIDE + 3] = v3 + tots[3];       // tots[3]: phi-score terms taken back on the rare paths
#pragma unroll
            for (int t = 0; t < GLMMA_NPACK; ++t) red[wib * GLMMA_PSTRIDE + 4 + t] = t < NP ? vs[t] : 0.0;
        }
        red[wib * GLMMA_PSTRIDE + GLMMA_NACC + lane] = gcol;     // column `lane` of X'zbar | X_zi'zbar_zi
        __syncthreads();
        if (threadIdx.x < GLMMA_PSTRIDE) {
            double v = 0.0;
            for (int ww = 0; ww < wpb; ++ww) v += red[ww * GLMMA_PSTRIDE + threadIdx.x];
            A.partials[(int64_t)blockIdx.x * GLMMA_PSTRIDE + threadIdx.x] = v;
        }
    }
}

// Node-independent terms of one evaluation for the clusters of the register variant:
//   partial[b] = { sum_rows w_row c_ll(row), sum_rows w_row c_phi(row) }
// over the rows whose weight is non-zero (roww: the cluster weight, 0 for the rows of clusters that
// belong to eval_kernel; null: every row with weight 1).  One memory-bound pass over y (8 or 16 bytes
// per observation), fixed-order block sums; finalize_block adds the partials.
#define GLMMA_CONST_THREADS 256
template <class F>
__global__ void __launch_bounds__(GLMMA_CONST_THREADS) const_sums_kernel(DevData d, FamPar fam, const double* ytab,
                                                                         const double* roww, double* partials) {
    __shared__ double sh[2][GLMMA_CONST_THREADS / 32];
    double s0 = 0.0, s1 = 0.0;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < d.N; row += (int64_t)gridDim.x * blockDim.x) {
        const double w = roww ? roww[row] : 1.0;
        if (w == 0.0) continue;
        double cll, cphi;
        obs_const_tab<F>(d.y1[row], F::USES_Y2 ? d.y2[row] : 0.0, fam, ytab, cll, cphi);
        s0 = fma(w, cll, s0);
        if (F::HAS_PHI) s1 = fma(w, cphi, s1);
    }
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    if (lane == 0) { sh[0][wi] = s0; sh[1][wi] = s1; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double v = 0.0;
        for (int t = 0; t < GLMMA_CONST_THREADS / 32; ++t) v += sh[threadIdx.x][t];
        partials[2 * blockIdx.x + threadIdx.x] = v;
    }
}

// ------------------------------------------------------------------------------
// per-observation phi-score contributions (i_contributions = TRUE path of
// score_mixed, R/Fit_Funs.R:186-187): sphi_obs[j] = w_i (sum_k pi_ik s_phi(ijk)).
// Rarely called (once per fit, for the sandwich estimator), so it simply
// recomputes the node weights from ll_i of a preceding eval.
// ------------------------------------------------------------------------------
template <class F, int QY, int QZ>
__global__ void __launch_bounds__(GLMMA_EVAL_MAX_THREADS) sphi_obs_kernel(EvalArgs A) {
    using L = EvalLayout<F, QY, QZ>;
    constexpr int Q = L::Q;
    constexpr int NP = L::NP;
    const DevData& d = A.d;
    extern __shared__ double smem[];
    double* fmtab = smem;
    double* gx = smem + GLMMA_FM_DOUBLES;
    double* glw = gx + GLMMA_NAGQ_MAX;
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int jt = A.jt;
    const int nagq = A.grid->nagq, K = A.grid->K;
    const unsigned rcp_nagq = nagq_rcp(nagq);
    double* rec = smem + L::head_doubles() + wib * (L::warp_doubles(jt, nagq, true) + (A.asum_doubles & ~1));
    double* scr = rec + jt * L::W + jt * (Q + L::NTABL) * nagq;
    load_fmtab(fmtab, A.fmtab);
    if (threadIdx.x < GLMMA_NAGQ_MAX) {
        gx[threadIdx.x] = A.grid->x[threadIdx.x];
        glw[threadIdx.x] = A.grid->lw[threadIdx.x];
    }
    __syncthreads();
    const MathCtx cx = make_mathctx(fmtab + 2 * (threadIdx.x & 7), fmtab + GLMMA_FM_LOG_DOUBLES + (threadIdx.x & 15), A.fmc);
    const int nrounds = (K + 31) >> 5;
    const int64_t nwarps = (int64_t)gridDim.x * wpb;
    for (int64_t i = (int64_t)blockIdx.x * wpb + wib; i < d.n; i += nwarps) {
        const int64_t r0 = d.row_ptr[i];
        const int ni = (int)(d.row_ptr[i + 1] - r0);
        const double w = d.weights ? d.weights[i] : 1.0;
        double m[Q], ri[NP];
#pragma unroll
        for (int dd = 0; dd < Q; ++dd) m[dd] = d.modes[i * Q + dd];
#pragma unroll
        for (int t = 0; t < NP; ++t) ri[t] = d.rinv[i * NP + t];
        double c_ll = 0.0;
        for (int j = lane; j < ni; j += 32) c_ll += d.cll[r0 + j];
        c_ll = warp_sum(c_ll);
        // log normaliser of the node weights: ll_i / w_i - logdet - c_ll
        const double lse = d.ll_i[i] / w - d.logdet[i] - c_ll;
        for (int j = lane; j < ni; j += 32) d.sphi_obs[r0 + j] = 0.0;
        __syncwarp();
        for (int r = 0; r < nrounds; ++r) {
            const int k = (r << 5) + lane;
            const bool valid = k < K;
            double b[Q];
            int idx[Q];
            double a = node_setup<Q>(valid ? k : K - 1, nagq, rcp_nagq, gx, glw, A.grid->lw_const, m, ri, A.prior, b, idx);
            for (int j0 = 0; j0 < ni; j0 += jt) {
                const int cnt = min(jt, ni - j0);
                __syncwarp();
                stage_tile<F, QY, QZ>(d, r0 + j0, cnt, lane, rec);
                __syncwarp();
                for (int j = 0; j < cnt; ++j) {
                    const double* rj = rec + j * L::W;
                    double eta, ez, emu, elam, ld, se, sz, sp;
                    lin_pred<F, QY, QZ, false>(rj, b, idx, j, nagq, nullptr, nullptr, cx, eta, ez, emu, elam);
                    F::template point<false, true>(rj[0], rj[1], eta, ez, emu, elam, A.fam, cx, ld, se, sz, sp);
                    a += ld;
                }
            }
            const double e = valid ? exp(a - lse) : 0.0;
            for (int j0 = 0; j0 < ni; j0 += jt) {
                const int cnt = min(jt, ni - j0);
                __syncwarp();
                stage_tile<F, QY, QZ>(d, r0 + j0, cnt, lane, rec);
                __syncwarp();
                for (int j = 0; j < cnt; ++j) {
                    const double* rj = rec + j * L::W;
                    double eta, ez, emu, elam, ld, se, sz, sp;
                    lin_pred<F, QY, QZ, false>(rj, b, idx, j, nagq, nullptr, nullptr, cx, eta, ez, emu, elam);
                    F::template point<true, true>(rj[0], rj[1], eta, ez, emu, elam, A.fam, cx, ld, se, sz, sp);
                    scr[j * 32 + lane] = e * sp;
                }
                __syncwarp();
                for (int j = lane; j < cnt; j += 32) {
                    double s1 = 0.0;
                    for (int c = 0; c < 32; ++c) s1 += scr[j * 32 + ((c + lane) & 31)];
                    d.sphi_obs[r0 + j0 + j] += s1;
                }
            }
        }
        __syncwarp();
        for (int j = lane; j < ni; j += 32)
            d.sphi_obs[r0 + j] = w * (d.sphi_obs[r0 + j] + d.cphi[r0 + j]);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------
// Observed information on the fixed grid: the Hessian of the NEGATIVE marginal log-likelihood the
// reference obtains as cd_vec(score_mixed) (R/mixed_fit.R:324-342: 2 * nparams evaluations), here from
// ONE pass over the data.  With a_ik(theta) the log of the k-th term of cluster i's quadrature sum and
// pi_ik its posterior node weight,
//     d^2 log p(y_i) / d theta d theta' = E_pi[ d^2 a ] + E_pi[ da da' ] - E_pi[da] E_pi[da]',
// theta = (betas, D parameters, phis, gammas) in the reference's order (R/mixed_fit.R:240-247).
//   da / d beta = X_i' s_eta(k),  da / d phis = sum_j s_phi,  da / d gamma = X_zi,i' s_zi(k)   (family scores)
//   da / d D_t  = c_t + b_k' M_t b_k / 2      (log N(b_k; 0, D): c_t, M_t from the host, engine.cu)
// The second derivatives of log f with respect to (eta, phis, eta_zi) are central differences of the
// functors' ANALYTIC first derivatives at every point (relative error ~1e-8, two orders below the
// truncation error of the reference's own cd_vec with eps = 1e-3), so no family needs hand-derived second
// derivatives.  A warp owns a cluster, lanes own nodes; two passes over the nodes (log-sum-exp, then the
// derivatives).  Run once per fit: clarity over speed (accumulators in local memory).
// ------------------------------------------------------------------------------
// capacity of the kernel in parameters: two instantiations, PM = 12 (accumulators of 78 pairs per thread in local
// memory) and PM = 24 (300 pairs: q = 4 with a full D alone has 10 D parameters) -- the small one is twice as fast
#define GLMMA_INFO_PMAX 24
#define GLMMA_INFO_NDMAX (GLMMA_QMAX * (GLMMA_QMAX + 1) / 2)
#define GLMMA_INFO_NDPAIR (GLMMA_INFO_NDMAX * (GLMMA_INFO_NDMAX + 1) / 2)
#define GLMMA_INFO_JT 16
struct InfoArgs {
    DevData d;
    PriorPar prior;
    FamPar fam, fam_hi, fam_lo;      // phis, phis + eps_phi, phis - eps_phi
    const GridPar* grid;
    const double* fmtab;
    double fmc[GLMMA_FM_NCONST];
    int nD;                                               // number of D parameters (q, or q (q + 1) / 2)
    double dc[GLMMA_INFO_NDMAX];                          // d log N / d D_t   = dc[t]  + b' dM[t]  b / 2
    double dM[GLMMA_INFO_NDMAX][GLMMA_QMAX * GLMMA_QMAX];
    double dcc[GLMMA_INFO_NDPAIR];                        // d^2 / d D_t d D_u = dcc[tu] + b' dMM[tu] b / 2   (t <= u, row-major)
    double dMM[GLMMA_INFO_NDPAIR][GLMMA_QMAX * GLMMA_QMAX];
    double eps_eta, eps_phi;
    double* partials;                                     // gridDim.x x npair(P) block partial Hessians of log p(y)
    int ksm;                                              // K when the K node terms of a cluster fit the warp's shared
                                                          // memory (pass 2 then reads them instead of recomputing), else 0
};

__host__ __device__ __forceinline__ int info_pair_pm(int r, int c, int pm) {      // r <= c, packed row-major over the upper triangle of pm x pm
    return r * (2 * pm - r + 1) / 2 + (c - r);
}

template <class F, int QY, int QZ, int PM>
__global__ void __launch_bounds__(128) info_kernel(InfoArgs A) {
    constexpr int NPAIR_ = PM * (PM + 1) / 2;
    auto info_pair = [](int r, int c) { return info_pair_pm(r, c, PM); };
    using L = EvalLayout<F, QY, QZ>;
    constexpr int Q = L::Q;
    constexpr int NP = L::NP;
    constexpr int JT = GLMMA_INFO_JT;
    const DevData& d = A.d;
    extern __shared__ double smem[];
    double* fmtab = smem;
    double* gx = smem + GLMMA_FM_DOUBLES;
    double* glw = gx + GLMMA_NAGQ_MAX;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int p = d.p, pz = F::HAS_ZI ? d.p_zi : 0, nphi = F::HAS_PHI ? 1 : 0, nD = A.nD;
    const int P = p + nD + nphi + pz;
    const int oD = p, oPhi = p + nD, oG = p + nD + nphi;
    const int xw = p + pz;
    double* rec = smem + L::head_doubles() + wib * (size_t)(JT * L::W + JT * PM + ((A.ksm + 1) & ~1));
    double* xt = rec + JT * L::W;                         // design rows of the staged tile: X | X_zi
    double* aks = xt + JT * PM;              // the cluster's K node terms a_k (pass 1 -> pass 2)
    const int nagq = A.grid->nagq, K = A.grid->K;
    const unsigned rcp_nagq = nagq_rcp(nagq);
    load_fmtab(fmtab, A.fmtab);
    if (threadIdx.x < GLMMA_NAGQ_MAX) {
        gx[threadIdx.x] = A.grid->x[threadIdx.x];
        glw[threadIdx.x] = A.grid->lw[threadIdx.x];
    }
    __syncthreads();
    const MathCtx cx = make_mathctx(fmtab + 2 * (threadIdx.x & 7), fmtab + GLMMA_FM_LOG_DOUBLES + (threadIdx.x & 15), A.fmc);
    const double lw_const = A.grid->lw_const;
    const int nrounds = (K + 31) >> 5;
    const double ie = 0.5 / A.eps_eta, ip = 0.5 / A.eps_phi;

    double tot[NPAIR_];                         // this lane's share of the block total
    for (int t = 0; t < NPAIR_; ++t) tot[t] = 0.0;

    auto point_at = [&](const double* rj, double eta, double ez, const FamPar& fam, double& ld, double& se, double& sz, double& sp) {
        const double emu = F::NEED_EMU ? fast_exp(eta, cx) : 0.0;
        const double elam = F::HAS_ZI ? fast_exp(fmin(fmax(ez, -690.0), 690.0), cx) : 0.0;
        F::template point<true, true>(rj[0], rj[1], eta, ez, emu, elam, fam, cx, ld, se, sz, sp);
    };
    auto stage = [&](int64_t r0, int cnt) {
        __syncwarp();
        stage_tile<F, QY, QZ>(d, r0, cnt, lane, rec);
        for (int t = lane; t < cnt * xw; t += 32) {
            const int j = t / xw, l = t - j * xw;
            xt[j * PM + l] = l < p ? d.X[r0 + j + (int64_t)l * d.N] : d.X_zi[r0 + j + (int64_t)(l - p) * d.N];
        }
        __syncwarp();
    };

    const int64_t nwarps = (int64_t)gridDim.x * wpb;
    for (int64_t i = (int64_t)blockIdx.x * wpb + wib; i < d.n; i += nwarps) {
        const int64_t r0 = d.row_ptr[i];
        const int ni = (int)(d.row_ptr[i + 1] - r0);
        const double w = d.weights ? d.weights[i] : 1.0;
        double m[Q], ri[NP];
#pragma unroll
        for (int dd = 0; dd < Q; ++dd) m[dd] = d.modes[i * Q + dd];
#pragma unroll
        for (int t = 0; t < NP; ++t) ri[t] = d.rinv[i * NP + t];
        // pass 1: log-sum-exp of the node terms (node-independent parts cancel in the weights)
        double mx = -INFINITY, se_sum = 0.0;
        for (int pass = 0; pass < 2; ++pass) {
            for (int r = 0; r < nrounds; ++r) {
                const int k = (r << 5) + lane;
                const bool valid = k < K;
                double b[Q];
                int idx[Q];
                double a = node_setup<Q>(valid ? k : K - 1, nagq, rcp_nagq, gx, glw, lw_const, m, ri, A.prior, b, idx);
                for (int j0 = 0; j0 < ni; j0 += JT) {
                    const int cnt = min(JT, ni - j0);
                    stage(r0 + j0, cnt);
                    for (int j = 0; j < cnt; ++j) {
                        const double* rj = rec + j * L::W;
                        double eta = rj[2], ez = F::HAS_ZI ? rj[L::OFF_XG] : 0.0;
#pragma unroll
                        for (int dd = 0; dd < QY; ++dd) eta = fma(rj[L::OFF_Z + dd], b[dd], eta);
#pragma unroll
                        for (int dd = 0; dd < QZ; ++dd) ez = fma(rj[L::OFF_Z + QY + dd], b[QY + dd], ez);
                        double ld, se, sz, sp;
                        point_at(rj, eta, ez, A.fam, ld, se, sz, sp);
                        a += ld;
                    }
                }
                if (pass == 0) { if (valid) mx = fmax(mx, a); }
                else if (valid) { se_sum += exp(a - mx); if (A.ksm) aks[k] = a; }
            }
            if (pass == 0) { mx = warp_max(mx); if (!(fabs(mx) < INFINITY)) mx = 0.0; }
        }
        se_sum = warp_sum(se_sum);
        const double lse = mx + log(se_sum);
        const bool ok = fabs(lse) < INFINITY;

        // pass 2: derivatives
        double Aacc[NPAIR_], gbar[PM];
        for (int t = 0; t < NPAIR_; ++t) Aacc[t] = 0.0;
        for (int t = 0; t < PM; ++t) gbar[t] = 0.0;
        for (int r = 0; r < nrounds && ok; ++r) {
            const int k = (r << 5) + lane;
            const bool valid = k < K;
            double b[Q];
            int idx[Q];
            double a = node_setup<Q>(valid ? k : K - 1, nagq, rcp_nagq, gx, glw, lw_const, m, ri, A.prior, b, idx);
            // node weight first: from pass 1's terms, or by its own loop over the observations
            if (A.ksm) a = valid ? aks[k] : 0.0;
            else for (int j0 = 0; j0 < ni; j0 += JT) {
                const int cnt = min(JT, ni - j0);
                stage(r0 + j0, cnt);
                for (int j = 0; j < cnt; ++j) {
                    const double* rj = rec + j * L::W;
                    double eta = rj[2], ez = F::HAS_ZI ? rj[L::OFF_XG] : 0.0;
#pragma unroll
                    for (int dd = 0; dd < QY; ++dd) eta = fma(rj[L::OFF_Z + dd], b[dd], eta);
#pragma unroll
                    for (int dd = 0; dd < QZ; ++dd) ez = fma(rj[L::OFF_Z + QY + dd], b[QY + dd], ez);
                    double ld, se, sz, sp;
                    point_at(rj, eta, ez, A.fam, ld, se, sz, sp);
                    a += ld;
                }
            }
            const double pik = valid ? exp(a - lse) : 0.0;
            double g[PM];
            for (int t = 0; t < PM; ++t) g[t] = 0.0;
            for (int j0 = 0; j0 < ni; j0 += JT) {
                const int cnt = min(JT, ni - j0);
                stage(r0 + j0, cnt);
                for (int j = 0; j < cnt; ++j) {
                    const double* rj = rec + j * L::W;
                    const double* xj = xt + j * PM;
                    double eta = rj[2], ez = F::HAS_ZI ? rj[L::OFF_XG] : 0.0;
#pragma unroll
                    for (int dd = 0; dd < QY; ++dd) eta = fma(rj[L::OFF_Z + dd], b[dd], eta);
#pragma unroll
                    for (int dd = 0; dd < QZ; ++dd) ez = fma(rj[L::OFF_Z + QY + dd], b[QY + dd], ez);
                    double ld, se, sz, sp, t0, se1, sz1, sp1, se2, sz2, sp2;
                    point_at(rj, eta, ez, A.fam, ld, se, sz, sp);
                    point_at(rj, eta + A.eps_eta, ez, A.fam, t0, se1, sz1, sp1);
                    point_at(rj, eta - A.eps_eta, ez, A.fam, t0, se2, sz2, sp2);
                    const double hee = (se1 - se2) * ie, hep = (sp1 - sp2) * ie, hez = (sz1 - sz2) * ie;
                    double hpp = 0.0, hzp = 0.0, hzz = 0.0;
                    if (F::HAS_PHI) {
                        point_at(rj, eta, ez, A.fam_hi, t0, se1, sz1, sp1);
                        point_at(rj, eta, ez, A.fam_lo, t0, se2, sz2, sp2);
                        hpp = (sp1 - sp2) * ip;
                        hzp = (sz1 - sz2) * ip;
                    }
                    if (F::HAS_ZI) {
                        point_at(rj, eta, ez + A.eps_eta, A.fam, t0, se1, sz1, sp1);
                        point_at(rj, eta, ez - A.eps_eta, A.fam, t0, se2, sz2, sp2);
                        hzz = (sz1 - sz2) * ie;
                    }
                    // first derivatives of a_k and pi_k-weighted second derivatives
                    for (int l = 0; l < p; ++l) {
                        g[l] = fma(xj[l], se, g[l]);
                        const double wx = pik * hee * xj[l];
                        for (int c = l; c < p; ++c) Aacc[info_pair(l, c)] = fma(wx, xj[c], Aacc[info_pair(l, c)]);
                        if (F::HAS_PHI) Aacc[info_pair(l, oPhi)] = fma(pik * hep, xj[l], Aacc[info_pair(l, oPhi)]);
                        if (F::HAS_ZI) {
                            const double wz = pik * hez * xj[l];
                            for (int c = 0; c < pz; ++c) Aacc[info_pair(l, oG + c)] = fma(wz, xj[p + c], Aacc[info_pair(l, oG + c)]);
                        }
                    }
                    if (F::HAS_PHI) {
                        g[oPhi] += sp;
                        Aacc[info_pair(oPhi, oPhi)] = fma(pik, hpp, Aacc[info_pair(oPhi, oPhi)]);
                    }
                    if (F::HAS_ZI) {
                        for (int l = 0; l < pz; ++l) {
                            g[oG + l] = fma(xj[p + l], sz, g[oG + l]);
                            const double wz = pik * hzz * xj[p + l];
                            for (int c = l; c < pz; ++c) Aacc[info_pair(oG + l, oG + c)] = fma(wz, xj[p + c], Aacc[info_pair(oG + l, oG + c)]);
                            if (F::HAS_PHI) Aacc[info_pair(oPhi, oG + l)] = fma(pik * hzp, xj[p + l], Aacc[info_pair(oPhi, oG + l)]);
                        }
                    }
                }
            }
            // prior part: D parameters
            for (int t = 0; t < nD; ++t) {
                double qf = 0.0;
#pragma unroll
                for (int a_ = 0; a_ < Q; ++a_)
#pragma unroll
                    for (int c_ = 0; c_ < Q; ++c_) qf = fma(b[a_] * A.dM[t][a_ * Q + c_], b[c_], qf);
                g[oD + t] = fma(0.5, qf, A.dc[t]);
            }
            {
                int tu = 0;
                for (int t = 0; t < nD; ++t)
                    for (int u = t; u < nD; ++u) {
                        double qf = 0.0;
#pragma unroll
                        for (int a_ = 0; a_ < Q; ++a_)
#pragma unroll
                            for (int c_ = 0; c_ < Q; ++c_) qf = fma(b[a_] * A.dMM[tu][a_ * Q + c_], b[c_], qf);
                        Aacc[info_pair(oD + t, oD + u)] = fma(pik, fma(0.5, qf, A.dcc[tu]), Aacc[info_pair(oD + t, oD + u)]);
                        ++tu;
                    }
            }
            // E[da da'] and E[da]
            for (int r_ = 0; r_ < P; ++r_) {
                gbar[r_] = fma(pik, g[r_], gbar[r_]);
                const double pg = pik * g[r_];
                for (int c_ = r_; c_ < P; ++c_) Aacc[info_pair(r_, c_)] = fma(pg, g[c_], Aacc[info_pair(r_, c_)]);
            }
        }
        // node-independent part of d^2 log f / d phis^2 (obs_const's c_phi differenced in phis)
        double cpp = 0.0;
        if (F::HAS_PHI && ok) {
            for (int j = lane; j < ni; j += 32) {
                double c0, c1, c2;
                const double y1 = d.y1[r0 + j], y2 = F::USES_Y2 ? d.y2[r0 + j] : 0.0;
                F::obs_const(y1, y2, A.fam_hi, c0, c1);
                F::obs_const(y1, y2, A.fam_lo, c0, c2);
                cpp += (c1 - c2) * ip;
            }
            cpp = warp_sum(cpp);
        }
        // cluster total over the lanes (nodes), then w_i (A - gbar gbar'); lane t keeps pairs t, t + 32, ...
        for (int r_ = 0; r_ < P; ++r_) gbar[r_] = warp_sum(gbar[r_]);
        for (int r_ = 0; r_ < P; ++r_)
            for (int c_ = r_; c_ < P; ++c_) {
                const int t = info_pair(r_, c_);
                double v = warp_sum(Aacc[t]) - gbar[r_] * gbar[c_];
                if (F::HAS_PHI && r_ == oPhi && c_ == oPhi) v += cpp;
                if (ok && ((r_ * P + c_) & 31) == lane) tot[t] = fma(w, v, tot[t]);
            }
        __syncwarp();
    }
    // block partials: every pair lives on exactly one lane of each warp
    __syncthreads();
    double* red = smem;        // wpb x NPAIR
    for (int r_ = 0; r_ < P; ++r_)
        for (int c_ = r_; c_ < P; ++c_) {
            const int t = info_pair(r_, c_);
            if (((r_ * P + c_) & 31) == lane) red[wib * NPAIR_ + t] = tot[t];
        }
    __syncthreads();
    for (int t = threadIdx.x; t < NPAIR_; t += blockDim.x) {
        double v = 0.0;
        for (int ww = 0; ww < wpb; ++ww) v += red[ww * NPAIR_ + t];
        A.partials[(int64_t)blockIdx.x * NPAIR_ + t] = v;
    }
}

// ------------------------------------------------------------------------------
// find_modes: per-cluster damped Newton on the reference's objective
//   g(b) = -sum_j log f(y_ij | b) + b' invD b / 2          (R/Functions.R:5-20)
// with its analytic gradient (R/Functions.R:21-65); the Newton matrix and the
// final Hessian are the reference's cd_vec of that gradient (R/Functions.R:90,
// 249-262).  G lanes share one cluster (G = 1: a thread per cluster).
// Control flow is warp-uniform: finished groups idle until the warp is done.
// ------------------------------------------------------------------------------
template <class F, int QY, int QZ, int G, bool SCORE>
__device__ __noinline__ void mode_objective(const DevData& d, const PriorPar& pr, const FamPar& fam,
                                               int64_t r0, int ni, int lg, unsigned gmask, const MathCtx& cx,
                                               const double* b, double& f, double* g) {
    constexpr int Q = QY + QZ;
    double fl = 0.0, gl[Q];
#pragma unroll
    for (int dd = 0; dd < Q; ++dd) gl[dd] = 0.0;
    for (int j = lg; j < ni; j += G) {
        const int64_t row = r0 + j;
        double z[Q];
        double eta = d.xb[row];
#pragma unroll
        for (int dd = 0; dd < QY; ++dd) { z[dd] = d.Z[row + (int64_t)dd * d.N]; eta = fma(z[dd], b[dd], eta); }
        double ez = 0.0;
        if (F::HAS_ZI) {
            ez = d.xg[row];
#pragma unroll
            for (int dd = 0; dd < QZ; ++dd) {
                z[QY + dd] = d.Z_zi[row + (int64_t)dd * d.N];
                ez = fma(z[QY + dd], b[QY + dd], ez);
            }
        }
        double ld, se, sz, sp;
        const double emu = F::NEED_EMU ? fast_exp(eta, cx) : 0.0;
        const double elam = F::HAS_ZI ? fast_exp(fmin(fmax(ez, -690.0), 690.0), cx) : 0.0;
        F::template point<SCORE, true>(d.y1[row], F::USES_Y2 ? d.y2[row] : 0.0, eta, ez, emu, elam, fam, cx, ld, se, sz, sp);
        if (ld == ld) fl -= ld;            // sum(log_dens, na.rm = TRUE), R/Functions.R:19
        if (SCORE) {
#pragma unroll
            for (int dd = 0; dd < QY; ++dd) gl[dd] = fma(-se, z[dd], gl[dd]);
#pragma unroll
            for (int dd = 0; dd < QZ; ++dd) gl[QY + dd] = fma(-sz, z[QY + dd], gl[QY + dd]);
        }
    }
    if (G > 1) {
        fl = group_sum<G>(fl, gmask);
        if (SCORE) {
#pragma unroll
            for (int dd = 0; dd < Q; ++dd) gl[dd] = group_sum<G>(gl[dd], gmask);
        }
    }
    double quad = 0.0;
#pragma unroll
    for (int dd = 0; dd < Q; ++dd) {
        double s = 0.0;
#pragma unroll
        for (int e = 0; e < Q; ++e) s = fma(pr.Dinv[dd * Q + e], b[e], s);
        quad = fma(b[dd], s, quad);
        if (SCORE) g[dd] = gl[dd] + s;
    }
    f = fl + 0.5 * quad;
}

// cd_vec(x, score_log_post_b): R/Functions.R:249-262, symmetrised
template <class F, int QY, int QZ, int G>
__device__ __forceinline__ void mode_hessian(const DevData& d, const PriorPar& pr, const FamPar& fam,
                                             int64_t r0, int ni, int lg, unsigned gmask, const MathCtx& cx,
                                             const double* x, double* H /* Q*Q */) {
    constexpr int Q = QY + QZ;
#pragma unroll
    for (int c = 0; c < Q; ++c) {
        double x1[Q], x2[Q], g1[Q], g2[Q], f;
#pragma unroll
        for (int dd = 0; dd < Q; ++dd) { x1[dd] = x[dd]; x2[dd] = x[dd]; }
        const double ex = 0.001 * fmax(fabs(x[c]), 1.0);
        x1[c] = x[c] + ex;
        x2[c] = x[c] - ex;
        mode_objective<F, QY, QZ, G, true>(d, pr, fam, r0, ni, lg, gmask, cx, x1, f, g1);
        mode_objective<F, QY, QZ, G, true>(d, pr, fam, r0, ni, lg, gmask, cx, x2, f, g2);
        const double dx = x1[c] - x2[c];
#pragma unroll
        for (int dd = 0; dd < Q; ++dd) H[dd * Q + c] = (g1[dd] - g2[dd]) / dx;
    }
#pragma unroll
    for (int a = 0; a < Q; ++a)
#pragma unroll
        for (int c = a + 1; c < Q; ++c) {
            const double v = 0.5 * (H[a * Q + c] + H[c * Q + a]);
            H[a * Q + c] = v;
            H[c * Q + a] = v;
        }
}

// lower Cholesky factor of H + lam I; returns false if not positive definite
template <int Q>
__device__ __forceinline__ bool chol_lower(const double* H, double lam, double* Lc) {
    bool ok = true;
#pragma unroll
    for (int c = 0; c < Q; ++c) {
        double s = H[c * Q + c] + lam;
#pragma unroll
        for (int t = 0; t < c; ++t) s -= Lc[c * Q + t] * Lc[c * Q + t];
        if (!(s > 0.0)) ok = false;
        const double dgl = sqrt(s);
        Lc[c * Q + c] = dgl;
#pragma unroll
        for (int r = c + 1; r < Q; ++r) {
            double v = H[r * Q + c];
#pragma unroll
            for (int t = 0; t < c; ++t) v -= Lc[r * Q + t] * Lc[c * Q + t];
            Lc[r * Q + c] = v / dgl;
        }
    }
    return ok;
}

template <class F, int QY, int QZ, int G>
__global__ void __launch_bounds__(128) modes_kernel(const __grid_constant__ ModeArgs A) {     // (references into the parameter space: no local copy)
    constexpr int Q = QY + QZ;
    constexpr int NP = Q * (Q + 1) / 2;
    const DevData& d = A.d;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = gtid / G;
    const int lg = (int)(gtid % G);
    const int lane = threadIdx.x & 31;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
    __shared__ double fmtab[GLMMA_FM_DOUBLES];
    load_fmtab(fmtab, A.fmtab);
    __syncthreads();
    const MathCtx cx = make_mathctx(fmtab + 2 * (threadIdx.x & 7), fmtab + GLMMA_FM_LOG_DOUBLES + (threadIdx.x & 15), A.fmc);
    const bool live = i < d.n;
    const int64_t r0 = live ? d.row_ptr[i] : 0;
    const int ni = live ? (int)(d.row_ptr[i + 1] - r0) : 0;

    double x[Q], g[Q], fx;
#pragma unroll
    for (int dd = 0; dd < Q; ++dd) x[dd] = (live && A.warm) ? d.modes[i * Q + dd] : 0.0;
    mode_objective<F, QY, QZ, G, true>(d, A.prior, A.fam, r0, ni, lg, gmask, cx, x, fx, g);
    bool done = !live;
    bool conv = false;
    int iters = 0;
    for (int it = 0; it < A.maxit; ++it) {
        if (__all_sync(0xffffffffu, done)) break;
        double H[Q * Q], Lc[Q * Q], dir[Q];
        mode_hessian<F, QY, QZ, G>(d, A.prior, A.fam, r0, ni, lg, gmask, cx, x, H);
        // Levenberg shift until positive definite
        double lam = 0.0, hmax = 1.0;
#pragma unroll
        for (int dd = 0; dd < Q; ++dd) hmax = fmax(hmax, fabs(H[dd * Q + dd]));
        bool pd = chol_lower<Q>(H, lam, Lc);
        for (int t = 0; t < 60 && !pd; ++t) {
            lam = fmax(10.0 * lam, 1e-3 * hmax);
            pd = chol_lower<Q>(H, lam, Lc);
        }
        // dir = -(L L')^-1 g
        double yv[Q];
#pragma unroll
        for (int r = 0; r < Q; ++r) {
            double s = g[r];
#pragma unroll
            for (int t = 0; t < r; ++t) s -= Lc[r * Q + t] * yv[t];
            yv[r] = s / Lc[r * Q + r];
        }
#pragma unroll
        for (int r = Q - 1; r >= 0; --r) {
            double s = yv[r];
#pragma unroll
            for (int t = r + 1; t < Q; ++t) s += Lc[t * Q + r] * dir[t];   // dir[t] already holds -x[t]
            dir[r] = -(s / Lc[r * Q + r]);
        }
        double gd = 0.0, dmax = 0.0;
#pragma unroll
        for (int dd = 0; dd < Q; ++dd) { gd = fma(g[dd], dir[dd], gd); dmax = fmax(dmax, fabs(dir[dd])); }
        // Armijo backtracking (c = 1e-4, halving) with a rounding slack so that full steps
        // inside the quadratic region are not rejected by noise in f
        double step = 1.0;
        bool accepted = done || !pd || !(dmax == dmax);
        bool moved = false;
        double xn[Q], gn[Q], fn = fx;
        for (int h = 0; h < 60; ++h) {
            if (__all_sync(0xffffffffu, accepted)) break;
            double xt[Q], gt[Q], ft;
#pragma unroll
            for (int dd = 0; dd < Q; ++dd) xt[dd] = fma(step, dir[dd], x[dd]);
            mode_objective<F, QY, QZ, G, true>(d, A.prior, A.fam, r0, ni, lg, gmask, cx, xt, ft, gt);
            if (!accepted) {
                const bool acc = (fabs(ft) < INFINITY) &&
                                 (ft <= fx + 1e-4 * step * gd + 1e-14 * (fabs(fx) + 1.0));
                if (acc) {
                    accepted = true; moved = true; fn = ft;
#pragma unroll
                    for (int dd = 0; dd < Q; ++dd) { xn[dd] = xt[dd]; gn[dd] = gt[dd]; }
                } else {
                    step *= 0.5;
                }
            }
        }
        if (!done) {
            ++iters;
            // converged when the Newton step is below the tolerance -- also when rounding noise in f
            // (a few 1e-14 with the table-based lgamma) makes the line search shorten or refuse such a
            // step; without an acceptable step otherwise, the cluster stays flagged as not converged
            if (!moved) {
                done = true;
                if (pd && dmax <= A.tol) conv = true;
            } else {
#pragma unroll
                for (int dd = 0; dd < Q; ++dd) { x[dd] = xn[dd]; g[dd] = gn[dd]; }
                fx = fn;
                if (dmax <= A.tol) { conv = true; done = true; }
            }
        }
    }
    // Hessian at the mode, its Cholesky factor, R^-1 and log_dets (R/Functions.R:90,115-122)
    double H[Q * Q], Lc[Q * Q];
    mode_hessian<F, QY, QZ, G>(d, A.prior, A.fam, r0, ni, lg, gmask, cx, x, H);
    const bool pd = chol_lower<Q>(H, 0.0, Lc);
    if (live && lg == 0) {
        // R = Lc' (upper); R^-1 by back substitution, column by column
        double Ri[Q * Q];
#pragma unroll
        for (int c = 0; c < Q; ++c) {
#pragma unroll
            for (int r = Q - 1; r >= 0; --r) {
                if (r > c) { Ri[r * Q + c] = 0.0; continue; }
                double s = (r == c) ? 1.0 : 0.0;
#pragma unroll
                for (int t = r + 1; t <= c; ++t) s -= Lc[t * Q + r] * Ri[t * Q + c];   // R[r][t] = Lc[t][r]
                Ri[r * Q + c] = s / Lc[r * Q + r];
            }
        }
        double ldet = 0.0;
#pragma unroll
        for (int dd = 0; dd < Q; ++dd) {
            d.modes[i * Q + dd] = x[dd];
            ldet -= log(Lc[dd * Q + dd]);
        }
        const double bad = pd ? 0.0 : NAN;
        d.logdet[i] = ldet + bad;
#pragma unroll
        for (int r = 0; r < Q; ++r)
#pragma unroll
            for (int c = r; c < Q; ++c) {
                d.chol[i * NP + Tri<Q>::cm(r, c)] = Lc[c * Q + r] + bad;
                d.rinv[i * NP + Tri<Q>::rm(r, c)] = Ri[r * Q + c] + bad;
            }
        if (!conv) atomicAdd(&A.stats[0], 1ull);
        if (!pd) atomicAdd(&A.stats[1], 1ull);
        atomicAdd(&A.stats[2], (unsigned long long)iters);
        atomicMax(&A.stats[3], (unsigned long long)iters);
    }
}

// ------------------------------------------------------------------------------
// log_post_b of predict()'s Metropolis step (R/methods.R:1020-1036, evaluated there per proposal
// and per cluster by mapply): one proposal b_i per cluster, 8 lanes per cluster,
//   out_i = sum_j log f(y_ij | eta_ij(b_i)) (na.rm) - b_i' invD b_i / 2.
// b_in is n x q column-major (R's matrix of proposals, one row per cluster).
// ------------------------------------------------------------------------------
template <class F, int QY, int QZ>
__global__ void __launch_bounds__(128) log_post_b_kernel(const __grid_constant__ ModeArgs A, const double* b_in, double* out) {
    constexpr int Q = QY + QZ;
    constexpr int G = 8;
    const DevData& d = A.d;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = gtid / G;
    const int lg = (int)(gtid % G);
    const int lane = threadIdx.x & 31;
    const unsigned gmask = ((1u << G) - 1u) << (lane & ~(G - 1));
    __shared__ double fmtab[GLMMA_FM_DOUBLES];
    load_fmtab(fmtab, A.fmtab);
    __syncthreads();
    const MathCtx cx = make_mathctx(fmtab + 2 * (threadIdx.x & 7), fmtab + GLMMA_FM_LOG_DOUBLES + (threadIdx.x & 15), A.fmc);
    const bool live = i < d.n;
    const int64_t r0 = live ? d.row_ptr[i] : 0;
    const int ni = live ? (int)(d.row_ptr[i + 1] - r0) : 0;
    double x[Q], g[Q], f;
#pragma unroll
    for (int dd = 0; dd < Q; ++dd) x[dd] = live ? b_in[i + (int64_t)dd * d.n] : 0.0;
    mode_objective<F, QY, QZ, G, false>(d, A.prior, A.fam, r0, ni, lg, gmask, cx, x, f, g);
    // node-independent parts of log f (prep kernel), dropped with the observation's density when NaN
    double c = 0.0;
    for (int j = lg; j < ni; j += G) {
        const double v = d.cll[r0 + j];
        if (v == v) c += v;
    }
    c = group_sum<G>(c, gmask);
    if (live && lg == 0) out[i] = c - f;
}

// ------------------------------------------------------------------------------
// Per-cluster record of the register variant (FastLayout::NSC doubles per cluster: modes, R^-1,
// logdet, Z'y1, weight), packed after every change of the grid so that the kernel's prefetch is one
// contiguous copy per cluster instead of five gathers.
// ------------------------------------------------------------------------------
template <int QY, int QZ>
__global__ void pack_cluster_kernel(DevData d) {
    constexpr int Q = QY + QZ, NP = Q * (Q + 1) / 2, NSC = Q + NP + 1 + QY + 1;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    double* r = d.clrec + i * NSC;
#pragma unroll
    for (int t = 0; t < Q; ++t) r[t] = d.modes[i * Q + t];
#pragma unroll
    for (int t = 0; t < NP; ++t) r[Q + t] = d.rinv[i * NP + t];
    r[Q + NP] = d.logdet[i];
#pragma unroll
    for (int t = 0; t < QY; ++t) r[Q + NP + 1 + t] = d.csum[i * GLMMA_NCSUM + 3 + t];
    r[Q + NP + 1 + QY] = d.weights ? d.weights[i] : 1.0;
}

// ------------------------------------------------------------------------------
// set_modes: R^-1 and log_dets from injected Cholesky factors
// ------------------------------------------------------------------------------
template <int Q>
__global__ void rinv_kernel(DevData d) {
    constexpr int NP = Q * (Q + 1) / 2;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    double R[Q * Q], Ri[Q * Q];
#pragma unroll
    for (int r = 0; r < Q; ++r)
#pragma unroll
        for (int c = 0; c < Q; ++c) R[r * Q + c] = (r <= c) ? d.chol[i * NP + Tri<Q>::cm(r, c)] : 0.0;
    double ldet = 0.0;
#pragma unroll
    for (int c = 0; c < Q; ++c) {
        ldet -= log(fabs(R[c * Q + c]));
#pragma unroll
        for (int r = Q - 1; r >= 0; --r) {
            if (r > c) { Ri[r * Q + c] = 0.0; continue; }
            double s = (r == c) ? 1.0 : 0.0;
#pragma unroll
            for (int t = r + 1; t <= c; ++t) s -= R[r * Q + t] * Ri[t * Q + c];
            Ri[r * Q + c] = s / R[r * Q + r];
        }
    }
    d.logdet[i] = ldet;
#pragma unroll
    for (int r = 0; r < Q; ++r)
#pragma unroll
        for (int c = r; c < Q; ++c) d.rinv[i * NP + Tri<Q>::rm(r, c)] = Ri[r * Q + c];
}
