// Instantiations of the AGQ kernels for stats::binomial() (logit, probit, cloglog).
#include "ops.cuh"
#include "../../include/glmma.h"
const FamOps* glmma_ops_binomial(int link, int qy, int qz) {
    if (link == GLMMA_LINK_DEFAULT || link == GLMMA_LINK_LOGIT) return dispatch_ops<FamBinomialLogit>(qy, qz);
    if (link == GLMMA_LINK_PROBIT) return dispatch_ops<FamBinomialOther<4>>(qy, qz);
    if (link == GLMMA_LINK_CLOGLOG) return dispatch_ops<FamBinomialOther<5>>(qy, qz);
    return nullptr;
}
