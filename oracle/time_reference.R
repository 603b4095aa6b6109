# Times the UNMODIFIED reference implementation of the hot path -- logLik_mixed() + score_mixed() of
# drizopoulos/GLMMadaptive (R/Fit_Funs.R:1-293) on a fixed adaptive Gauss-Hermite grid built by its own
# GHfun() (R/Functions.R:105-138) -- on the BASELINE config-3 workload written by bench.py.
# TEST / MEASUREMENT INFRASTRUCTURE ONLY: bench.py runs it when Rscript AND the reference tree are present
# on the machine (BASELINE.md B0/B1); R is not installed in the build image, where this script cannot
# be executed (bench.py then times the numpy port and says so).
#
#   Rscript oracle/time_reference.R <reference dir> <data dir> <N> <p> <q> <nAGQ>
# prints "SECONDS_PER_EVAL <s>": best of three (logLik_mixed + score_mixed) passes, one core.
args <- commandArgs(trailingOnly = TRUE)
ref <- args[1]; dir <- args[2]
N <- as.integer(args[3]); p <- as.integer(args[4]); q <- as.integer(args[5]); nAGQ <- as.integer(args[6])
suppressMessages(library(matrixStats))
for (f in c("Functions.R", "Fit_Funs.R")) source(file.path(ref, "R", f))
rd <- function (name, n, what = "double") readBin(file.path(dir, paste0(name, ".bin")), what, n = n, size = if (what == "double") 8 else 4)
y <- rd("y", N)
X <- matrix(rd("X", N * p), N, p)
Z <- matrix(rd("Z", N * q), N, q)
id <- rd("id", N, "integer")
theta <- rd("theta", p + q * q + 1)
betas <- theta[seq_len(p)]; D <- matrix(theta[p + seq_len(q * q)], q, q); phis <- theta[p + q * q + 1]
family <- negative.binomial()
# what mixed_fit() prepares (R/mixed_fit.R:4-38) and mixed_model() wires (R/mixed_model.R:211-240)
id_unq <- unique(id)
y_lis <- split(y, id)
X_lis <- lapply(id_unq, function (i) X[id == i, , drop = FALSE])
Z_lis <- lapply(id_unq, function (i) Z[id == i, , drop = FALSE])
Zty_lis <- lapply(mapply(function (z, y) crossprod(z, y), Z_lis, y_lis, SIMPLIFY = FALSE), drop)
Xty <- drop(crossprod(X, y))
log_dens <- family$log_dens; mu_fun <- family$linkinv; var_fun <- family$variance; mu.eta_fun <- family$mu.eta
score_eta_fun <- family$score_eta_fun; score_phis_fun <- family$score_phis_fun; score_eta_zi_fun <- NULL
canonical <- FALSE; user_defined <- TRUE
n <- length(id_unq)
GH <- GHfun(matrix(0.0, n, q), y_lis, NULL, X_lis, Z_lis, NULL, NULL, NULL, NULL, betas, solve(D), phis, NULL,
            nAGQ, q, canonical, user_defined, Zty_lis, log_dens, mu_fun, var_fun, mu.eta_fun,
            score_eta_fun, score_phis_fun, score_eta_zi_fun)
list_thetas <- list(betas = betas, D = chol_transf(D), phis = phis)
tht <- unlist(as.relistable(list_thetas))
one <- function () {
    t0 <- proc.time()[["elapsed"]]
    ll <- logLik_mixed(tht, id, y, NULL, X, Z, NULL, NULL, NULL, NULL, GH, canonical, user_defined, Xty, NULL,
                       log_dens, mu_fun, var_fun, mu.eta_fun, score_eta_fun, score_eta_zi_fun, score_phis_fun,
                       list_thetas, FALSE, FALSE, NULL, NULL, NULL, NULL)
    sc <- score_mixed(tht, id, y, NULL, X, Z, NULL, NULL, NULL, NULL, GH, canonical, user_defined, Xty, NULL,
                      log_dens, mu_fun, var_fun, mu.eta_fun, score_eta_fun, score_eta_zi_fun, score_phis_fun,
                      list_thetas, FALSE, FALSE, NULL, NULL, NULL, weights = NULL)
    proc.time()[["elapsed"]] - t0
}
cat("SECONDS_PER_EVAL", min(replicate(3, one())), "\n")
