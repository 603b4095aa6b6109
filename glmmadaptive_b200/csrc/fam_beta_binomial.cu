// Instantiations of the AGQ kernels for beta.binomial() (logit, cloglog) (see families.cuh).
#include "ops.cuh"
#include "../../include/glmma.h"
const FamOps* glmma_ops_beta_binomial(int link, int qy, int qz) {
    if (link == GLMMA_LINK_DEFAULT || link == GLMMA_LINK_LOGIT) return dispatch_ops<FamBetaBinomialL<1>>(qy, qz);
    if (link == GLMMA_LINK_CLOGLOG) return dispatch_ops<FamBetaBinomialL<5>>(qy, qz);
    return nullptr;
}
