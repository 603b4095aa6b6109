# B200 engine behind GLMMadaptive's own closures (drizopoulos/GLMMadaptive v0.9-7).
#
# What a maintainer adds to the package: this file under R/, src/glmma_R.c (r/src/glmma_R.c here),
# libglmma.so on the loader path, `control$engine = "b200"` accepted by mixed_model()
# (R/mixed_model.R:150-160: add `engine = "R"` to `con`), and ONE branch at the top of mixed_fit()
# (R/mixed_fit.R:1):
#
#     if (identical(control$engine, "b200") &&
#         !is.null(eng <- b200_engine(y, X, Z, X_zi, Z_zi, id, offset, offset_zi, weights, family,
#                                     initial_values, control, Funs)))
#         return(mixed_fit_b200(eng, y, X, Z, X_zi, Z_zi, id, offset, offset_zi, family,
#                               initial_values, Funs, control, penalized, weights))
#
# Everything below keeps the ARGUMENT LISTS AND RETURN SHAPES of the reference closures it replaces, so
# the rest of mixed_fit() (EM loop, optim(), nearPD(), the Hessian, the returned list) is the
# reference's own code operating on these closures.  Nothing of size N x K is built in R: `GH` carries
# the external pointer plus post_modes / log_dets / wGH (and post_vars on demand).
#
# R is not installed in the build image, so this file is not executed by the repo's tests; the same
# call sequence against the same C ABI is exercised from Python (glmmadaptive_b200/engine.py, fit.py).

# families whose closures the engine implements, by family$family (include/glmma.h: glmma_family_id)
.b200_constructors <- c(
    "binomial" = "binomial", "poisson" = "poisson", "negative binomial" = "negative.binomial",
    "zero-inflated poisson" = "zi.poisson", "zero-inflated negative binomial" = "zi.negative.binomial",
    "zero-inflated binomial" = "zi.binomial", "hurdle poisson" = "hurdle.poisson",
    "hurdle negative binomial" = "hurdle.negative.binomial", "hurdle log-normal" = "hurdle.lognormal",
    "beta" = "beta.fam", "hurdle beta" = "hurdle.beta.fam", "Gamma" = "Gamma.fam",
    "censored normal" = "censored.normal", "Student's-t" = "students.t", "beta binomial" = "beta.binomial")

# TRUE when `family` is the package's own object for that name: a user may define a family called
# "beta" with different closures (vignettes/Custom_Models.Rmd:295-333) -- that one must stay in R.
b200_family_is_builtin <- function (family) {
    ctor <- .b200_constructors[family$family]
    if (is.na(ctor)) return(FALSE)
    if (family$family %in% c("binomial", "poisson"))
        return(is.null(family$log_dens) || identical(deparse(family$log_dens),
               deparse(get(paste0(family$family, "_log_dens"), asNamespace("GLMMadaptive")))))
    ref <- switch(ctor,
                  "students.t" = get(ctor, asNamespace("GLMMadaptive"))(df = environment(family$log_dens)$.df),
                  get(ctor, asNamespace("GLMMadaptive"))())
    same <- function (f, g) is.null(f) == is.null(g) && (is.null(f) || identical(deparse(f), deparse(g)))
    same(family$log_dens, ref$log_dens) && same(family$score_eta_fun, ref$score_eta_fun) &&
        same(family$score_phis_fun, ref$score_phis_fun) && same(family$score_eta_zi_fun, ref$score_eta_zi_fun)
}

# external pointer to the device-resident data set, or NULL to stay on the reference path
b200_engine <- function (y, X, Z, X_zi, Z_zi, id, offset, offset_zi, weights, family, initial_values,
                         control, Funs) {
    if (!b200_family_is_builtin(family)) return(NULL)
    storage.mode(X) <- "double"; storage.mode(Z) <- "double"
    if (!is.null(X_zi)) storage.mode(X_zi) <- "double"
    if (!is.null(Z_zi)) storage.mode(Z_zi) <- "double"
    yy <- if (is.matrix(y)) { storage.mode(y) <- "double"; y } else as.double(y)
    df <- if (family$family == "Student's-t") environment(family$log_dens)$.df else NULL
    devices <- if (is.null(control$devices)) 0L else as.integer(control$devices)
    .Call("glmma_R_create", yy, X, Z, X_zi, Z_zi, as.integer(id),
          if (!is.null(offset)) as.double(offset), if (!is.null(offset_zi)) as.double(offset_zi),
          if (!is.null(weights)) as.double(weights), family$family, family$link,
          length(initial_values$phis), as.integer(control$nAGQ), df, devices,
          PACKAGE = "GLMMadaptive")
}

# wGH of GHfun() (R/Functions.R:124-125)
b200_wGH <- function (k, q) {
    GH <- gauher(k)
    wGH <- as.matrix(expand.grid(lapply(seq_len(q), function (i, w) w, w = GH$w)))
    b <- as.matrix(expand.grid(lapply(seq_len(q), function (i, x) x, x = GH$x)))
    2^(q / 2) * apply(wGH, 1, prod) * exp(rowSums(b * b))
}

# unpack the result vector of glmma_eval / glmma_em_estep / glmma_em_score (include/glmma.h)
b200_unpack <- function (r, p, n_phis, p_zi, q) {
    o <- 4L
    list(loglik = r[1L], sum_w = r[2L], n_nonfinite = r[3L],
         g_beta = r[o + seq_len(p)], g_phi = r[o + p + seq_len(n_phis)],
         g_gamma = r[o + p + n_phis + seq_len(p_zi)],
         S = matrix(r[o + p + n_phis + p_zi + seq_len(q * q)], q, q))
}

# score.D of score_mixed() (R/Fit_Funs.R:237-286) from S = sum_i w_i E[b b' | y_i]: with
# cS_postVB + post_b post_b' summed over clusters equal to S, the reference's
#   out[i] = sum(D2[i, ] * cS_postVB) + sum((post_b %*% matrix(D2[i, ], nRE)) * post_b)
# is sum(D2[i, ] * S).
b200_score_D <- function (S, sum_w, n, D, diag_D) {
    nRE <- ncol(D)
    if (diag_D) {
        D <- diag(D)
        svD <- 1 / D
        svD2 <- svD^2
        return(D * 0.5 * (n * svD - diag(S) * svD2))          # R/Fit_Funs.R:244-247 (n, not sum(weights))
    }
    svD <- solve(D)
    dD <- deriv_D(D)
    ndD <- length(dD)
    D1 <- sapply(dD, function (x) sum(svD * x))
    D2 <- t(sapply(dD, function (x) c(svD %*% x %*% svD)))
    out <- numeric(ndD)
    for (i in seq_len(ndD)) out[i] <- sum(D2[i, ] * c(S))
    J <- jacobian2(attr(D, "L"), nRE)
    drop(0.5 * (sum_w * D1 - out) %*% J)                       # R/Fit_Funs.R:283-285
}

# the closures mixed_fit() uses, bound to one engine handle
b200_closures <- function (eng, p, n_phis, p_zi, q, n, nAGQ, list_thetas_skeleton, diag_D,
                           penalized, pen_mu, pen_invSigma, pen_df) {
    cache <- new.env()
    GHfun <- function (b, ..., betas, inv_D, phis, gammas, k, q, canonical, user_defined, Zty_lis,
                       log_dens, mu_fun, var_fun, mu.eta_fun, score_eta_fun, score_phis_fun,
                       score_eta_zi_fun) {                      # R/Functions.R:105-110, same names
        m <- .Call("glmma_R_find_modes", eng, betas, inv_D, phis, gammas, any(b != 0), PACKAGE = "GLMMadaptive")
        cache$key <- NULL
        structure(list(handle = eng, post_modes = m[[1L]], chol = m[[2L]], log_dets = m[[3L]],
                       wGH = b200_wGH(k, q), stats = m[[4L]]), class = "GH_b200")
    }
    fused <- function (thetas) {      # optim() calls fn then gr at the same thetas: one device pass
        key <- paste(format(thetas, digits = 17), collapse = " ")
        if (!identical(cache$key, key)) {
            th <- relist(thetas, skeleton = list_thetas_skeleton)
            D <- if (diag_D) diag(exp(th$D), length(th$D)) else chol_transf(th$D)
            cache$res <- b200_unpack(.Call("glmma_R_eval", eng, th$betas, D, th$phis, th$gammas, TRUE,
                                           PACKAGE = "GLMMadaptive"), p, n_phis, p_zi, q)
            cache$D <- D; cache$betas <- th$betas; cache$key <- key
        }
        cache$res
    }
    logLik_mixed <- function (thetas, ...) {                   # R/Fit_Funs.R:1-42
        r <- fused(thetas)
        out <- -r$loglik
        if (penalized)
            out <- out - dmvt(cache$betas, mu = pen_mu, invSigma = pen_invSigma, df = pen_df)
        out
    }
    score_mixed <- function (thetas, ...) {                    # R/Fit_Funs.R:44-293
        r <- fused(thetas)
        sb <- -r$g_beta
        if (penalized) {
            v <- cache$betas * diag(pen_invSigma) / pen_df
            sb <- sb + v * (pen_df + p) / c(1 + crossprod(cache$betas, v))
        }
        c(sb, b200_score_D(r$S, r$sum_w, n, cache$D, diag_D), -r$g_phi, -r$g_gamma)
    }
    # EM phase (R/mixed_fit.R:119-229): Ztb / p_by stay on the device
    e_step <- function (betas, D, phis, gammas)
        b200_unpack(.Call("glmma_R_em_estep", eng, betas, D, phis, gammas, PACKAGE = "GLMMadaptive"),
                    p, n_phis, p_zi, q)
    em_score <- function (betas, phis, gammas)
        b200_unpack(.Call("glmma_R_em_score", eng, betas, phis, gammas, PACKAGE = "GLMMadaptive"),
                    p, n_phis, p_zi, q)
    score_betas <- function (betas, ..., phis, gammas) {       # R/Fit_Funs.R:295-365
        out <- -em_score(betas, phis, gammas)$g_beta
        if (penalized) {
            v <- betas * diag(pen_invSigma) / pen_df
            out <- out + v * (pen_df + p) / c(1 + crossprod(betas, v))
        }
        out
    }
    score_phis <- function (phis, ..., betas, gammas) -em_score(betas, phis, gammas)$g_phi      # :367-403
    score_gammas <- function (gammas, ..., betas, phis) -em_score(betas, phis, gammas)$g_gamma  # :405-427
    # predict(type = "subject_specific", se.fit = TRUE): R/methods.R:1020-1036, all clusters at once
    log_post_b <- function (b, betas, invD, phis, gammas)
        .Call("glmma_R_log_post_b", eng, b, betas, invD, phis, gammas, PACKAGE = "GLMMadaptive")
    list(GHfun = GHfun, logLik_mixed = logLik_mixed, score_mixed = score_mixed, e_step = e_step,
         score_betas = score_betas, score_phis = score_phis, score_gammas = score_gammas,
         log_post_b = log_post_b)
}

# post_vars of GHfun() (R/Functions.R:119-120) from the packed Cholesky factors, on demand
post_vars.GH_b200 <- function (GH) {
    q <- ncol(GH$post_modes)
    lapply(seq_len(nrow(GH$post_modes)), function (i) {
        U <- matrix(0, q, q)
        U[upper.tri(U, TRUE)] <- GH$chol[, i]
        chol2inv(U)
    })
}
