"""CPU: the R glue (r/src/glmma_R.c) stays consistent with the C ABI (include/glmma.h).

R is not installed in the build image, so the glue cannot be loaded; what can be checked is checked:
it compiles without warnings against the minimal declarations of R's documented C API in r/stub/
(so every glmma_* call in it has the argument types include/glmma.h declares), the object's undefined
symbols are exactly R API names plus symbols the header declares, every .Call name used by the R side
(r/R/engine_b200.R) is registered with the right arity, and every entry point the R seam needs is bound."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GLUE = os.path.join(ROOT, "r", "src", "glmma_R.c")
RSIDE = os.path.join(ROOT, "r", "R", "engine_b200.R")
INC = ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "r", "stub")]


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "glmma.h")).read()
    return set(re.findall(r"GLMMA_API\s+[\w\s\*]+?\b(glmma_\w+)\s*\(", src))


def test_glue_compiles_against_the_header(tmp_path):
    subprocess.run(["gcc", "-fsyntax-only", "-std=c99", "-Wall", "-Werror", "-Wno-cast-function-type"] + INC + [GLUE],
                   check=True)
    obj = str(tmp_path / "glmma_R.o")
    subprocess.run(["gcc", "-c", "-fPIC", "-std=c99"] + INC + [GLUE, "-o", obj], check=True)
    undef = set(subprocess.run(["nm", "-u", obj], check=True, capture_output=True, text=True).stdout.split()) - {"U"}
    undef = {u.split("@")[0] for u in undef}
    ours = {u for u in undef if u.startswith("glmma_")}
    assert ours and ours <= _header_symbols(), ours - _header_symbols()
    rest = undef - ours
    allowed = re.compile(r"^(Rf_\w+|R_\w+|REAL|INTEGER|CHAR|STRING_ELT|VECTOR_ELT|SET_VECTOR_ELT|memset|_GLOBAL_OFFSET_TABLE_|__stack_chk_fail)$")
    assert all(allowed.match(u) for u in rest), [u for u in rest if not allowed.match(u)]
    # the whole reference seam is bound
    for need in ("glmma_create", "glmma_create_multi", "glmma_find_modes", "glmma_get_modes", "glmma_set_modes",
                 "glmma_eval", "glmma_eval_contrib", "glmma_em_estep", "glmma_em_score", "glmma_glm_pass",
                 "glmma_log_post_b", "glmma_destroy"):
        assert need in ours, need


def test_r_side_calls_are_registered_with_matching_arity():
    c = open(GLUE).read()
    reg = {m.group(1): int(m.group(2)) for m in re.finditer(r'\{"(glmma_R_\w+)",\s*\(DL_FUNC\)\s*&\w+,\s*(\d+)\}', c)}
    # registered arity == number of SEXP parameters of the definition
    for name, k in reg.items():
        sig = re.search(r"SEXP " + name + r"\(([^)]*)\)", c).group(1)
        assert sig.count("SEXP") == k, (name, k, sig)
    r = open(RSIDE).read()
    calls = re.findall(r'\.Call\("(glmma_R_\w+)"((?:[^()]|\((?:[^()]|\([^()]*\))*\))*)\)', r)
    assert calls
    for name, args in calls:
        assert name in reg, name
        depth, n = 0, 0
        for ch in args:
            if ch in "([":
                depth += 1
            elif ch in ")]":
                depth -= 1
            elif ch == "," and depth == 0:
                n += 1
        assert n - 1 == reg[name], (name, n - 1, reg[name])       # minus the PACKAGE = argument
