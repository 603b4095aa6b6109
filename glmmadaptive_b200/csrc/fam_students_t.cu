// Instantiations of the AGQ kernels for one family functor (see families.cuh).
#include "ops.cuh"
const FamOps* glmma_ops_students_t(int qy, int qz) { return dispatch_ops<FamStudentsT>(qy, qz); }
