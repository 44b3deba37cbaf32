"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/*.h declares,
reports errors through status codes (no compute calls need a GPU here), and its general-nu
machinery (host-compiled copy of the same routines) matches mpmath."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import gpvecchia_b200 as G
from gpvecchia_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "gpvecchia_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gpv_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 18
    for nm in names:
        assert hasattr(G.lib, nm), f"{nm} declared in include/gpvecchia_b200.h but not exported"
    assert sorted(_lib.EXPORTED) == names


def test_version_and_counters():
    assert b"sm_100a" in G.lib.gpv_version()
    assert G.lib.gpv_launch_count() >= 0
    assert G.lib.gpv_device_count() >= 0


def test_bad_arguments_are_status_codes_not_crashes():
    h = C.c_void_p()
    st = G.lib.gpv_create(C.byref(h), 10, 3, 2, None, None, None, 0, None, 0, 10, 0)
    assert st == _lib.GPV_ERR_ARG and b"null" in G.lib.gpv_last_error()
    x = np.zeros(20)
    nn = np.zeros(30, dtype=np.int32)
    st = G.lib.gpv_create(C.byref(h), 10, 65, 2, x.ctypes.data, nn.ctypes.data, nn.ctypes.data, 0, None, 0, 10, 0)
    assert st == _lib.GPV_ERR_ARG and b"GPV_MAX_P" in G.lib.gpv_last_error()
    st = G.lib.gpv_create(C.byref(h), 10, 3, 2, x.ctypes.data, nn.ctypes.data, nn.ctypes.data, 0, None, 5, 3, 0)
    assert st == _lib.GPV_ERR_ARG
    assert G.lib.gpv_packed_len(None) == 0
    G.lib.gpv_destroy(None)
    # the entry points added for SURVEY.md 8(f): null handles / arrays are status codes too
    a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    assert G.lib.gpv_csc_dims(None, C.byref(a), C.byref(b), C.byref(c)) == _lib.GPV_ERR_ARG
    assert G.lib.gpv_u_sparsity(None, None, None) == _lib.GPV_ERR_ARG
    assert G.lib.gpv_u_csc_pattern(None, None, None) == _lib.GPV_ERR_ARG
    nf, ff = C.c_int64(0), C.c_int64(0)
    assert G.lib.gpv_u_values_csc(None, b"matern", None, 3, None, None, 0, None, C.byref(nf), C.byref(ff)) == _lib.GPV_ERR_ARG
    assert G.lib.gpv_u_nzentries_mat(None, None, None, 0, None, None, C.byref(nf), C.byref(ff)) == _lib.GPV_ERR_ARG
    assert G.lib.gpv_multi_csc_dims(None, C.byref(a), C.byref(b), C.byref(c)) == _lib.GPV_ERR_ARG
    assert G.lib.gpv_whichCondOnLatent(None, 5, 3, 0, None) == _lib.GPV_ERR_ARG


def test_whichCondOnLatent_is_host_code_and_needs_no_device():
    # 4 locations, m = 2: row k conditions on latent y only for the best earlier neighbour and its own
    # latent neighbours (R/whichCondOnLatent.R:2-27); NA stays NA, self is always TRUE
    NA = np.iinfo(np.int32).min
    NN = np.asfortranarray(np.array([[1, NA, NA], [2, 1, NA], [3, 2, 1], [4, 3, 2]], dtype=np.int32))
    out = np.empty_like(NN)
    assert G.lib.gpv_whichCondOnLatent(NN.ctypes.data, 4, 3, 0, out.ctypes.data) == _lib.GPV_OK
    assert out[0, 0] == 1 and out[0, 1] == NA and out[0, 2] == NA
    assert np.all(out[:, 0] == 1) and out[1, 2] == NA
    assert set(np.unique(out[2:, 1:])) <= {0, 1}


def test_unknown_covtype_is_an_error_before_touching_the_device():
    # the reference only prints (U_NZentries.cpp:27-29); the boundary returns GPV_ERR_COVTYPE
    with pytest.raises(G.GpvError) as ei:
        G.U_NZentries(1, 1, np.zeros((1, 2)), np.array([[1]]), np.array([[1]]), [0.1], [0.1], "gauss", [1, 1, 1])
    assert ei.value.status == _lib.GPV_ERR_COVTYPE
    assert "gauss covariance is not implemented" in str(ei.value)


@pytest.mark.skipif(G.lib.gpv_device_count() > 0, reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_a_gpu():
    with pytest.raises(G.GpvError) as ei:
        G.U_NZentries(1, 1, np.zeros((1, 2)), np.array([[1]]), np.array([[1]]), [0.1], [0.1], "matern", [1, 1, 1.5])
    assert ei.value.status == _lib.GPV_ERR_CUDA


def test_general_nu_evaluator_against_mpmath_golden():
    rows = json.load(open(os.path.join(GOLD, "matern_general_mpmath.json")))["rows"]
    worst = 0.0
    for nu, s, val in rows:
        got = G.lib.gpv_selftest_matern_general_host(s, 1.0, nu)
        if val == 0.0:
            assert abs(got) < 1e-300
            continue
        worst = max(worst, abs(got - val) / abs(val) / max(1.0, s))
    assert worst < 1e-14, worst       # relative error <= 1e-14 * max(1, s), any nu incl. near-integers


def test_general_nu_table_against_mpmath_golden():
    rows = json.load(open(os.path.join(GOLD, "matern_general_mpmath.json")))["rows"]
    rng_, wmax = 0.004, 2.0
    worst = worst_rel = 0.0
    for nu, s, val in rows:
        w = (s * rng_) ** 2
        s_eff = np.sqrt(w) / rng_
        if val == 0.0 or abs(s_eff - s) > 1e-14 * s:
            pass
        got = G.lib.gpv_selftest_table_eval_host(w, 1.0, rng_, nu, wmax)
        if val == 0.0:
            continue
        # What a covariance MATRIX needs is the absolute error relative to sigma^2 (= 1 here): the closed forms
        # carry 3e-16 of it, the table (degree 10 on quarter octaves of w, bessel_table.cuh) stays below 4e-15 for
        # any nu and s; the relative error holds to 1e-14 while the covariance is not small (s <= 3) and for
        # s >= 16, where the table stores exp(s) * cov.  (Forming w = (s*range)^2 and back perturbs s by ~2 ulp.)
        worst = max(worst, abs(got - val))
        if s <= 3.0 or s >= 16.5:
            worst_rel = max(worst_rel, abs(got - val) / abs(val) / max(1.0, s))
    assert worst < 4e-15, worst
    assert worst_rel < 2e-14, worst_rel


def test_r_shim_type_checks_against_the_c_abi():
    """r_shim/src/gpv_shim.c cannot be built here (no R), but its calls into include/gpvecchia_b200.h can
    be type-checked: `gcc -fsyntax-only` against declarations of the R API subset it uses
    (tests/r_api_mock, written from the R-extensions manual).  Catches a shim that drifts from the C ABI."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "r_api_mock"),
                        "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "r_shim", "src", "gpv_shim.c")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = open(os.path.join(ROOT, "r_shim", "src", "gpv_shim.c")).read()
    # the reference's own .Call names stay registered with the reference's arities (src/RcppExports.cpp)
    for name, arity in (("_GPvecchia_U_NZentries", 9), ("_GPvecchia_U_NZentries_mat", 9), ("_GPvecchia_ic0", 3),
                        ("_GPvecchia_createUcppM", 3), ("_GPvecchia_createUcpp", 4), ("_GPvecchia_MaternFun", 2),
                        ("_GPvecchia_EsqeFun", 2)):
        assert re.search(r'\{"%s",\s*\(DL_FUNC\)&\w+,\s*%d\}' % (name, arity), src), name
    # the caller's revNNarray is never written through (Rf_coerceVector may return its argument)
    assert not re.search(r"ip\[i\]\s*=", src) and "const int* ip = INTEGER(nn)" in src
    # every .Call name the R side (r_shim/R/createU_b200.R) uses is registered
    rsrc = open(os.path.join(ROOT, "r_shim", "R", "createU_b200.R")).read()
    for name in set(re.findall(r'\.Call\("(_GPvecchia_\w+)"', rsrc)):
        assert ('{"%s"' % name) in src, name
