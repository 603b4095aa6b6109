"""CPU-side checks: the C-ABI library loads and exports every symbol include/glmma.h
declares, fails loudly without a GPU (no CPU fallback), and the host-side theta algebra
of glmmadaptive_b200.engine agrees with the oracle's restatement of the reference's
chol_transf / deriv_D / jacobian2 chain (R/Functions.R:140-202, R/Fit_Funs.R:237-286)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import agq
import glmmadaptive_b200 as gb
from glmmadaptive_b200 import _lib, engine as E, sharded

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "glmma.h")).read()
    declared = set(re.findall(r"GLMMA_API\s+[\w\s\*]+?\b(glmma_\w+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().glmma_version() == 200


def test_family_ids_match_reference_strings():
    lib = _lib.load()
    for name, fid in _lib.FAMILY_IDS.items():
        assert lib.glmma_family_id(name.encode()) == fid
    assert lib.glmma_family_id(b"Conway Maxwell Poisson") == -1      # not accelerated: stays on the R path
    with pytest.raises(ValueError):
        gb.family_spec("compoisson")
    with pytest.raises(ValueError):
        gb.family_spec("poisson", link="sqrt")


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.GlmmaError) as ei:
        gb.Engine(np.zeros(4), np.ones((4, 1)), np.ones((4, 1)), [0, 0, 1, 1], "poisson")
    assert ei.value.code == _lib.ERR_CUDA and "no CPU fallback" in str(ei.value)


def test_create_argument_validation_precedes_device_use():
    lib = _lib.load()
    d = _lib.GlmmaData()
    h = ctypes.c_void_p()
    assert lib.glmma_create(ctypes.byref(d), 0, ctypes.byref(h)) == _lib.ERR_ARG
    assert b"required" in lib.glmma_last_error(None)


def test_chol_transf_roundtrip_and_oracle_agreement():
    rng = np.random.default_rng(0)
    for q in (1, 2, 3):
        A = rng.standard_normal((q, q))
        D = A @ A.T + np.eye(q)
        v = gb.chol_transf(D)
        assert np.allclose(v, agq.chol_transf(D), atol=1e-14)
        D2, U = gb.chol_transf(v)
        assert np.allclose(D2, D, atol=1e-12)


def test_score_D_chain_rule_matches_reference_jacobian():
    """score_D(S, sum_w, D) must equal the reference's deriv_D / jacobian2 assembly for any S."""
    rng = np.random.default_rng(1)
    for q in (1, 2, 3):
        A = rng.standard_normal((q, q))
        D = A @ A.T + np.eye(q)
        v = agq.chol_transf(D)
        _, L = agq.chol_transf(v)
        _, U = gb.chol_transf(v)
        B = rng.standard_normal((q, q))
        S = B @ B.T * 7.0
        n = 13.0
        svD = np.linalg.inv(D)
        dD = agq.deriv_D(D)
        D1 = np.array([np.sum(svD * x) for x in dD])
        out = np.array([np.sum((svD @ x @ svD) * S) for x in dD])
        ref = 0.5 * (n * D1 - out) @ agq.jacobian2(L, q)
        got = E.score_D(S, n, n, D, U, False)
        assert np.allclose(got, ref, rtol=1e-12, atol=1e-12)
        dd = np.diag(D)
        ref_diag = dd * 0.5 * (n / dd - np.diag(S) / dd ** 2)
        assert np.allclose(E.score_D(S, n, n, np.diag(dd), None, True), ref_diag, rtol=1e-13)


def test_shard_bounds_balance_and_cover():
    rng = np.random.default_rng(2)
    sizes = 1 + rng.poisson(7, 1000)
    rp = np.concatenate(([0], np.cumsum(sizes)))
    for ws in (1, 2, 4, 8):
        b = sharded.shard_bounds(rp, ws)
        assert b[0] == 0 and b[-1] == 1000 and np.all(np.diff(b) >= 0)
        obs = np.diff(rp[b])
        assert obs.sum() == rp[-1]
        assert obs.max() - obs.min() <= 2 * sizes.max()
