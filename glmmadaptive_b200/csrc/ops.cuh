// Host-side launchers for one (family functor, q_y, q_zi) combination and the
// dispatch from runtime dimensions to the compiled instantiation.  Each
// fam_*.cu translation unit includes this file and exports one lookup function.
#pragma once
#include <cstdlib>
#include "kernels.cuh"
#include "engine.h"

// The dynamic shared-memory limit of a kernel is a per-device property of the FUNCTION, shared by every
// handle in the process: always raise it to the device's opt-in maximum, so that a handle created later
// with a smaller tile cannot lower it under the launches of an earlier one.
static inline int glmma_optin_smem() {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    return v;
}

template <class F, int QY, int QZ>
struct OpsImpl {
    using L = EvalLayout<F, QY, QZ>;

    static void ytab(const FamPar& fam, double* tab, cudaStream_t s) {
        if (F::Y_TABLE) ytab_kernel<F><<<GLMMA_YTAB / 256, 256, 0, s>>>(fam, tab);
    }
    static void consts(const DevData& d, const FamPar& fam, const double* tab, const double* roww, double* partials,
                       int blocks, cudaStream_t s) {
        const_sums_kernel<F><<<blocks, GLMMA_CONST_THREADS, 0, s>>>(d, fam, F::Y_TABLE ? tab : nullptr, roww, partials);
    }
    static void prep(const PrepArgs& a, const CoefPar& c, cudaStream_t s) {
        if (a.d.csum) {
            const int64_t ncl = a.list ? a.n_list : a.d.n;
            if (ncl > 0) prep_cluster_kernel<F><<<(unsigned)((ncl * 8 + 255) / 256), 256, 0, s>>>(a, c);
        } else {
            const int64_t blocks = (a.d.N + 255) / 256;
            prep_kernel<F><<<(unsigned)blocks, 256, 0, s>>>(a, c);
        }
    }
    // Block shape of the generic kernel: its per-warp area grows with the observation tile jt, and the
    // fastmath tables (24 KB) are per block, so larger blocks keep more warps resident when jt is large.
    template <bool SCORE>
    static cudaError_t eval_config_s(int jt, int nagq, int extra, int* blocks_per_sm, size_t* smem_bytes, int* threads) {
        cudaError_t e = cudaFuncSetAttribute(eval_kernel<F, QY, QZ, SCORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, glmma_optin_smem());
        if (e != cudaSuccess) return e;
        int best = 0;
        *blocks_per_sm = 0;
        for (int t = GLMMA_EVAL_THREADS; t <= GLMMA_EVAL_MAX_THREADS; t += GLMMA_EVAL_THREADS) {
            const size_t smem = L::smem_bytes(jt, nagq, SCORE, t / 32, extra);
            if (smem > (size_t)glmma_optin_smem()) break;
            int bps = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, eval_kernel<F, QY, QZ, SCORE>, t, smem);
            if (e != cudaSuccess) return e;
            if (bps * t > best) { best = bps * t; *blocks_per_sm = bps; *smem_bytes = smem; *threads = t; }
        }
        return cudaSuccess;
    }
    static cudaError_t eval_config(int score, int jt, int nagq, int extra, int* blocks_per_sm, size_t* smem_bytes, int* threads) {
        if (score) {
            // sphi_obs_kernel shares the scoring layout and launch shape
            cudaError_t e = cudaFuncSetAttribute(sphi_obs_kernel<F, QY, QZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, glmma_optin_smem());
            if (e != cudaSuccess) return e;
            return eval_config_s<true>(jt, nagq, extra, blocks_per_sm, smem_bytes, threads);
        }
        return eval_config_s<false>(jt, nagq, extra, blocks_per_sm, smem_bytes, threads);
    }
    static void eval(const EvalArgs& a, int score, int grid, int threads, size_t smem, cudaStream_t s) {
        if (score) eval_kernel<F, QY, QZ, true><<<grid, threads, smem, s>>>(a);
        else eval_kernel<F, QY, QZ, false><<<grid, threads, smem, s>>>(a);
    }
    // register variant (eval_fast_kernel): tiles of nj = 8, 12 or 16 observations per cluster.  Only the
    // fused (log-likelihood + score) form is instantiated: optim's fn and gr are answered from one fused
    // pass anyway (engine.py: _fused_eval), and a log-likelihood-only call simply ignores the score
    // outputs -- the loglik-only instantiations would add a quarter to build time and binary size.
    // Block shape of the register variant: the fastmath tables (24 KB) are per block, so families with
    // large per-warp areas (zero-inflated, q = 3) keep more warps resident with fewer, larger blocks.
    // Picks the shape with the most resident warps per SM; ties go to the larger block.
    // q = 4: nAGQ^4 nodes exceed the register variant's grids for every nAGQ > 4 -- not instantiated
    static constexpr bool HAS_FAST = QY + QZ <= 3;
    // tile of 32 observations: the register variant's two-pass form, carrier families only (kernels.cuh)
    static constexpr bool HAS_TILE32 = HAS_FAST && (F::FAST_KIND == 1 || F::FAST_KIND == 2);
    // pair tile (nj code 17): two random-intercept clusters per warp, carrier families only (kernels.cuh)
    static constexpr bool HAS_PAIR = QY == 1 && QZ == 0 && (F::FAST_KIND == 1 || F::FAST_KIND == 2);
    template <int NJ, bool PAIR = false>
    static cudaError_t fast_config_nj(int K, int nagq, int xw, int* blocks_per_sm, size_t* smem_bytes, int* threads) {
        cudaError_t e = cudaFuncSetAttribute(eval_fast_kernel<F, QY, QZ, true, NJ, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, glmma_optin_smem());
        if (e != cudaSuccess) return e;
        int best = 0;
        *blocks_per_sm = 0;
        // tuning knob (measurements only): GLMMA_FAST_THREADS=<threads per block> pins the block shape
        const char* pin = std::getenv("GLMMA_FAST_THREADS");
        const int pinned = pin ? std::atoi(pin) : 0;
        for (int t = GLMMA_EVAL_THREADS; t <= GLMMA_FAST_MAX_THREADS; t += 32) {
            if (pinned && t != pinned) continue;
            const size_t smem = FastLayout<F, QY, QZ, NJ, PAIR>::smem_bytes(t / 32, K, xw, nagq);
            if (smem > (size_t)glmma_optin_smem()) break;
            int bps = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, eval_fast_kernel<F, QY, QZ, true, NJ, PAIR>, t, smem);
            if (e != cudaSuccess) return e;
            // ties go to the LARGER block: one 16-warp block per SM measured 3 % faster than four 4-warp blocks
            // at C3 (one copy of the per-block tables, one start-up)
            if (bps > 0 && bps * t >= best) { best = bps * t; *blocks_per_sm = bps; *smem_bytes = smem; *threads = t; }
        }
        return cudaSuccess;
    }
    static cudaError_t fast_config(int xw, int nj, int K, int nagq, int* blocks_per_sm, size_t* smem_bytes, int* threads) {
        if constexpr (!HAS_FAST) { *blocks_per_sm = 0; *smem_bytes = 0; *threads = 0; return cudaSuccess; }
        else {
            if (nj == 32) {
                if constexpr (HAS_TILE32) return fast_config_nj<32>(K, nagq, xw, blocks_per_sm, smem_bytes, threads);
                else { *blocks_per_sm = 0; *smem_bytes = 0; *threads = 0; return cudaSuccess; }
            }
            if (nj == 17) {
                if constexpr (HAS_PAIR) return fast_config_nj<16, true>(K, nagq, xw, blocks_per_sm, smem_bytes, threads);
                else { *blocks_per_sm = 0; *smem_bytes = 0; *threads = 0; return cudaSuccess; }
            }
            return nj == 8 ? fast_config_nj<8>(K, nagq, xw, blocks_per_sm, smem_bytes, threads)
                 : nj == 12 ? fast_config_nj<12>(K, nagq, xw, blocks_per_sm, smem_bytes, threads)
                            : fast_config_nj<16>(K, nagq, xw, blocks_per_sm, smem_bytes, threads);
        }
    }
    static void eval_fast(const EvalArgs& a, int /*score*/, int nj, int grid, int threads, size_t smem, cudaStream_t s) {
        if constexpr (!HAS_FAST) return;
        else if (nj == 8) eval_fast_kernel<F, QY, QZ, true, 8><<<grid, threads, smem, s>>>(a);
        else if (nj == 12) eval_fast_kernel<F, QY, QZ, true, 12><<<grid, threads, smem, s>>>(a);
        else if (nj == 16) eval_fast_kernel<F, QY, QZ, true, 16><<<grid, threads, smem, s>>>(a);
        else if (nj == 17) { if constexpr (HAS_PAIR) eval_fast_kernel<F, QY, QZ, true, 16, true><<<grid, threads, smem, s>>>(a); }
        else if constexpr (HAS_TILE32) eval_fast_kernel<F, QY, QZ, true, 32><<<grid, threads, smem, s>>>(a);
    }
    static void sphi(const EvalArgs& a, int grid, int threads, size_t smem, cudaStream_t s) {
        sphi_obs_kernel<F, QY, QZ><<<grid, threads, smem, s>>>(a);
    }
    static void modes(const ModeArgs& a, int G, cudaStream_t s) {
        const int64_t threads = a.d.n * G;
        const unsigned blocks = (unsigned)((threads + 127) / 128);
        if (G == 1) modes_kernel<F, QY, QZ, 1><<<blocks, 128, 0, s>>>(a);
        else if (G == 8) modes_kernel<F, QY, QZ, 8><<<blocks, 128, 0, s>>>(a);
        else modes_kernel<F, QY, QZ, 32><<<blocks, 128, 0, s>>>(a);
    }
    static void info(const InfoArgs& a, int grid, int pm, cudaStream_t s) {
        using LL = EvalLayout<F, QY, QZ>;
        const size_t smem = sizeof(double) * (LL::head_doubles() + 4 * (size_t)(GLMMA_INFO_JT * LL::W + GLMMA_INFO_JT * pm + ((a.ksm + 1) & ~1)));
        if (pm == 12) {
            cudaFuncSetAttribute(info_kernel<F, QY, QZ, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, glmma_optin_smem());
            info_kernel<F, QY, QZ, 12><<<grid, 128, smem, s>>>(a);
        } else {
            cudaFuncSetAttribute(info_kernel<F, QY, QZ, GLMMA_INFO_PMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, glmma_optin_smem());
            info_kernel<F, QY, QZ, GLMMA_INFO_PMAX><<<grid, 128, smem, s>>>(a);
        }
    }
    static void logpost(const ModeArgs& a, const double* b, double* out, cudaStream_t s) {
        const unsigned blocks = (unsigned)((a.d.n * 8 + 127) / 128);
        log_post_b_kernel<F, QY, QZ><<<blocks, 128, 0, s>>>(a, b, out);
    }
    static void pack(const DevData& d, cudaStream_t s) {
        const unsigned blocks = (unsigned)((d.n + 127) / 128);
        pack_cluster_kernel<QY, QZ><<<blocks, 128, 0, s>>>(d);
    }
    static void rinv(const DevData& d, cudaStream_t s) {
        const unsigned blocks = (unsigned)((d.n + 127) / 128);
        rinv_kernel<QY + QZ><<<blocks, 128, 0, s>>>(d);
    }
    static const FamOps* get() {
        static const FamOps ops = {ytab, consts, prep, eval_config, eval, fast_config, eval_fast, sphi, modes, info, logpost, rinv, pack, F::HAS_ZI, F::HAS_PHI, F::USES_Y2, F::Y_TABLE, F::LIN_ETA, HAS_TILE32, HAS_PAIR};
        return &ops;
    }
};

template <class F>
static const FamOps* dispatch_ops(int qy, int qz) {
    if (!F::HAS_ZI && qz != 0) return nullptr;
    if (qz == 0) {
        if (qy == 1) return OpsImpl<F, 1, 0>::get();
        if (qy == 2) return OpsImpl<F, 2, 0>::get();
        if (qy == 3) return OpsImpl<F, 3, 0>::get();
        if (qy == 4) return OpsImpl<F, 4, 0>::get();
        return nullptr;
    }
    if (F::HAS_ZI) {
        if (qy == 1 && qz == 1) return OpsImpl<F, 1, F::HAS_ZI ? 1 : 0>::get();
        if (qy == 2 && qz == 1) return OpsImpl<F, 2, F::HAS_ZI ? 1 : 0>::get();
        if (qy == 1 && qz == 2) return OpsImpl<F, 1, F::HAS_ZI ? 2 : 0>::get();
        if (qy == 3 && qz == 1) return OpsImpl<F, 3, F::HAS_ZI ? 1 : 0>::get();
        if (qy == 2 && qz == 2) return OpsImpl<F, 2, F::HAS_ZI ? 2 : 0>::get();
    }
    return nullptr;
}
