// Internal host-side interface between engine.cu (the C ABI) and the per-family
// translation units.
#pragma once
#include "common.cuh"

struct PrepArgs;
struct EvalArgs;
struct ModeArgs;
struct InfoArgs;

struct FamOps {
    void (*ytab)(const FamPar&, double* tab, cudaStream_t);      // table of obs_const(y), count families
    // node-independent terms of the register variant's clusters: block partials {sum w c_ll, sum w c_phi}
    void (*consts)(const DevData&, const FamPar&, const double* ytab, const double* roww, double* partials, int blocks,
                   cudaStream_t);
    void (*prep)(const PrepArgs&, const CoefPar&, cudaStream_t);
    cudaError_t (*eval_config)(int score, int jt, int nagq, int extra, int* blocks_per_sm, size_t* smem_bytes, int* threads);
    void (*eval)(const EvalArgs&, int score, int grid, int threads, size_t smem, cudaStream_t);
    cudaError_t (*fast_config)(int xw, int nj, int K, int nagq, int* blocks_per_sm, size_t* smem_bytes, int* threads);
    void (*eval_fast)(const EvalArgs&, int score, int nj, int grid, int threads, size_t smem, cudaStream_t);
    void (*sphi)(const EvalArgs&, int grid, int threads, size_t smem, cudaStream_t);
    void (*modes)(const ModeArgs&, int G, cudaStream_t);
    void (*info)(const InfoArgs&, int grid, int pm, cudaStream_t);      // observed information, one pass (pm: 12 or GLMMA_INFO_PMAX)
    void (*logpost)(const ModeArgs&, const double* b, double* out, cudaStream_t);
    void (*rinv)(const DevData&, cudaStream_t);
    void (*pack)(const DevData&, cudaStream_t);       // packed per-cluster records of the register variant
    bool has_zi, has_phi, uses_y2;
    bool y_table;      // obs_const() is tabulated by ytab (count families)
    bool lin_eta;      // the register variant hoists y1 * eta out of the point loop (families.cuh: LIN_ETA)
    bool tile32;       // the register variant has its two-pass form for clusters of 17 .. 32 observations
    bool pair;         // the register variant has its pair tile (two random-intercept clusters per warp)
};

// one lookup per family translation unit (nullptr: dimensions not compiled in)
const FamOps* glmma_ops_binomial(int link, int qy, int qz);
const FamOps* glmma_ops_poisson(int qy, int qz);
const FamOps* glmma_ops_negbin(int qy, int qz);
const FamOps* glmma_ops_zi_poisson(int qy, int qz);
const FamOps* glmma_ops_zi_negbin(int qy, int qz);
const FamOps* glmma_ops_zi_binomial(int qy, int qz);
const FamOps* glmma_ops_hurdle_poisson(int qy, int qz);
const FamOps* glmma_ops_hurdle_negbin(int qy, int qz);
const FamOps* glmma_ops_hurdle_lognormal(int qy, int qz);
const FamOps* glmma_ops_beta(int qy, int qz);
const FamOps* glmma_ops_hurdle_beta(int qy, int qz);
const FamOps* glmma_ops_gamma(int qy, int qz);
const FamOps* glmma_ops_censored_normal(int qy, int qz);
const FamOps* glmma_ops_students_t(int qy, int qz);
const FamOps* glmma_ops_beta_binomial(int link, int qy, int qz);
