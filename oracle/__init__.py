"""CPU oracle for the GLMMadaptive AGQ hot path -- TEST INFRASTRUCTURE ONLY.

This package is a numpy/scipy restatement of the reference R code
(/root/reference/R/Functions.R, R/Fit_Funs.R, R/mixed_fit.R; file:line cited on
every function).  It exists to *check* the CUDA engine.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; the product package ``glmmadaptive_b200`` never
does and has no CPU fallback.

Parity pin: the reference ships no tests.  The oracle is
pinned instead against the known-answer values printed in the reference's
rendered vignettes (``docs/articles/*.html``), reproduced
in ``tests/test_oracle_kat.py`` from data regenerated with a restatement of R's
RNG (``oracle/r_rng.py``), and against mpmath high-precision evaluations of the
family formulas (``tests/test_oracle_families.py``).  R itself is not installed
in the build container, so the reference cannot be executed here.
"""
