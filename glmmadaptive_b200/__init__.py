"""glmmadaptive_b200 -- B200-native engine for GLMMadaptive's adaptive
Gauss-Hermite marginal log-likelihood / score hot path.

Host-side mirror of the reference closures (GHfun, logLik_mixed, score_mixed) and of
its fitting loop (mixed_model, mixed_fit) over the C ABI in include/glmma.h.  See DESIGN.md.
"""
from .engine import Engine, GH, GHfun, Penalty, logLik_mixed, score_mixed, chol_transf, family_spec  # noqa: F401
from .fit import (mixed_model, mixed_fit, starting_values, starting_values_device, glm_fit_device,  # noqa: F401
                  predict_subject_specific)
from .sharded import ShardedEngine, shard_bounds  # noqa: F401
