"""The band-folded set kernel's LOGIC on the CPU (no GPU needed): gpvecchia_b200/csrc/u_band.cuh is compiled
unchanged with g++ against stand-ins for the CUDA builtins (tests/simt_emu/: one host thread per CUDA thread,
barriers for __syncthreads / __syncwarp, shuffles and ballots through a per-warp slot array, cp.async as an
immediate copy, uninitialised shared memory filled with NaN) and its output is compared with the oracle:
row-major and packed U values, zero fill, the fused likelihood partial sums, failure counting, missing
entries, p < P padding, three instantiations (G = 8 with four bands, G = 8 and 16 with three bands, d = 2 / 3).
This is a check of index maps, compaction, elimination order and outputs; the parity tests proper are the
`-m gpu` tests on a B200."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle as O
from gpvecchia_b200 import harness as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIBS = {}
# the warp-specialised experiment hands slots over through spin-waits: a protocol bug must fail, not hang, the run
pytestmark = pytest.mark.timeout(600)


def _build(tmpdir, defs=()):
    key = tuple(defs)
    if key in _LIBS:
        return _LIBS[key]
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    so = os.path.join(str(tmpdir), "libemu_" + "_".join(d.replace("=", "") for d in defs) + ".so")
    cmd = ["g++", "-std=c++17", "-O1", "-pthread", "-shared", "-fPIC", "-Wno-unknown-pragmas",
           "-I", os.path.join(ROOT, "tests", "simt_emu"), "-I", os.path.join(ROOT, "gpvecchia_b200", "csrc"),
           "-I", os.path.join(ROOT, "include")] + ["-D" + d for d in defs] + \
          [os.path.join(ROOT, "tests", "simt_emu", "emu_harness.cpp"), "-o", so]
    subprocess.check_call(cmd)
    L = C.CDLL(so)
    vp = C.c_void_p
    L.emu_u_band.argtypes = [C.c_int] * 4 + [C.c_int64, C.c_int, C.c_int] + [vp] * 7 + [C.c_int, vp, vp, vp, C.c_int, vp]
    L.emu_u_band.restype = C.c_int
    L.emu_u_sets.argtypes = [C.c_int] + L.emu_u_band.argtypes
    L.emu_u_sets.restype = C.c_int
    _LIBS[key] = L
    return L


@pytest.fixture(scope="module")
def emu_dir(tmp_path_factory):
    return tmp_path_factory.mktemp("simt_emu")


def _cov_consts(covType, cp):
    """What setup_cov (gpv_capi.cu) puts into UParams: kind, (c0..c4)."""
    if covType == "matern":
        sig2, rng_, nu = cp
        kind = {0.5: 0, 1.5: 1, 2.5: 2}[nu]
        c1 = {0.5: 1.0, 1.5: np.sqrt(3.0), 2.5: np.sqrt(5.0)}[nu] / rng_
        return kind, np.array([sig2, c1, 0.0, 0.0, 0.0])
    s1, r1, s2, r2 = cp
    return 3, np.array([s1 + s2, 1.0 / r1, s2, 1.0 / (r2 * r2), s1])


def _problem(n, m, d, seed, layout="z", p_drop=0.0):
    rng = np.random.default_rng(seed)
    locs = rng.random((n, d))
    NN = H.ordered_nn_kdtree(locs, m)
    Cond = H.layout_yz(NN, layout)
    va = H.make_vecchia_approx(locs, NN, Cond, np.ones(n, dtype=bool), layout)
    revNN = np.array(va["U_prep"]["revNNarray"], dtype=np.int64)
    revCond = np.array(va["U_prep"]["revCond"])
    if p_drop > 0:                       # knock out neighbours anywhere in the row (never the self entry)
        kill = rng.random(revNN[:, :-1].shape) < p_drop
        revNN[:, :-1][kill] = 0
    rcf = revCond.astype(np.float64)
    rcf[revCond < 0] = np.nan
    if p_drop > 0:
        # the reference compacts the ids (`inds.elem(find(inds))`) but reads revCond from the LAST n0 columns
        # (U_NZentries.cpp:44-47): keep those columns valid, whatever positions the holes are in
        p = revNN.shape[1]
        val = 1.0 if layout == "y" else 0.0
        for k in range(n):
            k0 = int((revNN[k] != 0).sum())
            rcf[k, :p - k0] = np.nan
            rcf[k, p - k0:] = val
            rcf[k, -1] = 1.0
    return locs, revNN, rcf


def _run(L, G, P, locs, revNN, rcf, nug, covType, cp, grid=2, packed=False, z=None, family=1, D=None):
    n, p = revNN.shape
    d = locs.shape[1]
    nn = np.ascontiguousarray(revNN.astype(np.int32) - 1)             # 0-based, -1 = missing
    cond = np.zeros(n, dtype=np.uint64)
    for j in range(p):
        cond |= (np.nan_to_num(rcf[:, j], nan=0.0) == 1.0).astype(np.uint64) << np.uint64(j)
    n0 = (revNN != 0).sum(axis=1)
    out = np.full(n * p, np.nan)
    row_off = None
    if packed:
        row_off = np.concatenate([[0], np.cumsum(n0)[:-1]]).astype(np.int64)
    zloc = None if z is None else np.ascontiguousarray(z, dtype=np.float64)
    partials = None if z is None else np.full(grid * 4, np.nan)
    nfail = np.zeros(1, dtype=np.uint64)
    first = np.full(1, np.iinfo(np.int64).max, dtype=np.int64)
    kind, c = _cov_consts(covType, cp)
    P_ = lambda a: None if a is None else a.ctypes.data
    lr = np.ascontiguousarray(locs, dtype=np.float64)
    rc = L.emu_u_sets(family, G, P, d if D is None else D, grid, n, p, d, P_(lr), P_(nn), P_(cond),
                      P_(np.ascontiguousarray(nug)), P_(out), P_(row_off), P_(zloc), 1 if z is not None else 0,
                      P_(partials), P_(nfail), P_(first), kind, P_(c))
    assert rc == 0
    return out, partials, int(nfail[0]), int(first[0]), n0


def _oracle(locs, revNN, rcf, nug, covType, cp, mode=0):
    n = locs.shape[0]
    return O.U_NZentries(O.max_threads(), n, locs, revNN, rcf, nug, nug, covType, np.asarray(cp, dtype=float), mode=mode)


@pytest.mark.parametrize("G,P,m,d,covType,cp", [
    (8, 31, 30, 2, "matern", [1.3, 0.25, 1.5]),
    (8, 31, 27, 2, "matern", [1.0, 0.3, 0.5]),          # p = 28 < P = 31: leading padding of every row
    (8, 31, 30, 2, "esqe", [0.7, 0.3, 0.4, 0.2]),
    (8, 21, 20, 3, "matern", [0.9, 0.4, 2.5]),          # three bands of eight lanes
    (16, 41, 40, 3, "matern", [1.0, 0.5, 1.5]),         # three bands of sixteen lanes
])
def test_band_kernel_logic_on_the_host_matches_the_oracle(emu_dir, G, P, m, d, covType, cp):
    L = _build(emu_dir)
    n = 150                                   # 2 blocks x 4 warps x (32 / G) sets per pass: several passes + a ragged tail
    locs, revNN, rcf = _problem(n, m, d, seed=G * 100 + P + m)
    nug = np.random.default_rng(1).uniform(0.05, 0.15, n)
    z = np.random.default_rng(2).standard_normal(n)
    ref = _oracle(locs, revNN, rcf, nug, covType, cp)
    got, partials, nfail, _, n0 = _run(L, G, P, locs, revNN, rcf, nug, covType, cp, z=z)
    got = got.reshape(n, m + 1)
    Lr = ref["Lentries"]
    assert nfail == 0 and ref["nfail"] == 0
    assert not np.isnan(got).any()                                   # every slot written (values or zero fill)
    assert np.array_equal(got == 0, Lr == 0)                         # pattern, zero fill beyond n0
    scale = np.abs(Lr).max(axis=1, keepdims=True)
    assert (np.abs(got - Lr) / scale).max() < 1e-10
    # fused likelihood partial sums (vecchia_likelihood.R:74-76 per set): sum_k (sum_j x_kj z_j)^2, sum_k log x_kk
    ps = partials.reshape(-1, 4).sum(axis=0)
    ids = revNN - 1
    quad = 0.0
    logd = 0.0
    for k in range(n):
        k0 = int(n0[k])
        xs = Lr[k, :k0]
        nb = ids[k][ids[k] >= 0]
        rc = rcf[k][~np.isnan(rcf[k])]
        quad += float((xs[rc == 0] * z[nb[rc == 0]]).sum()) ** 2
        logd += np.log(xs[-1])
    assert abs(ps[0] - quad) <= 1e-9 * abs(quad) and abs(ps[1] - logd) <= 1e-9 * abs(logd)


@pytest.mark.parametrize("G,P,m,d,D,covType,cp", [
    (16, 31, 30, 2, 2, "matern", [1.3, 0.25, 1.5]),     # the kernel the band family replaced at this size
    (4, 8, 7, 5, 0, "matern", [1.0, 0.8, 0.5]),         # run-time dimension (D = 0 instantiation), four lanes per set
    (8, 11, 9, 3, 3, "esqe", [0.7, 0.5, 0.4, 0.3]),     # p = 10 < P = 11
    (32, 51, 50, 2, 2, "matern", [1.0, 0.4, 2.5]),      # one set per warp
])
def test_two_row_kernel_logic_on_the_host_matches_the_oracle(emu_dir, G, P, m, d, D, covType, cp):
    L = _build(emu_dir)
    n = 120
    locs, revNN, rcf = _problem(n, m, d, seed=G + P)
    nug = np.random.default_rng(1).uniform(0.05, 0.15, n)
    ref = _oracle(locs, revNN, rcf, nug, covType, cp)
    got, _, nfail, _, _ = _run(L, G, P, locs, revNN, rcf, nug, covType, cp, family=0, D=D)
    got = got.reshape(n, m + 1)
    Lr = ref["Lentries"]
    assert nfail == 0 and not np.isnan(got).any()
    assert np.array_equal(got == 0, Lr == 0)
    assert (np.abs(got - Lr) / np.abs(Lr).max(axis=1, keepdims=True)).max() < 1e-10


def test_band_kernel_missing_entries_packed_order_and_failures(emu_dir):
    L = _build(emu_dir)
    n, m, d = 100, 30, 2
    locs, revNN, rcf = _problem(n, m, d, seed=7, layout="z", p_drop=0.2)     # holes anywhere in the rows
    nug = np.full(n, 0.1)
    locs[12, 0] = np.nan                      # every set that contains point 12 fails (NaN pivot: dpotrf's disnan test)
    cp = [1.0, 0.3, 1.5]
    ref = _oracle(locs, revNN, rcf, nug, "matern", cp, mode=1)   # published dpotf2 (OpenBLAS skips the NaN test)
    got, _, nfail, first, n0 = _run(L, 8, 31, locs, revNN, rcf, nug, "matern", cp, packed=True)
    # packed (createU.R:158-160) order: the n0 values of each row, rows concatenated
    exp = np.concatenate([ref["Lentries"][k, :n0[k]] for k in range(n)])
    gotp = got[:exp.size]
    failed = np.array([np.all(ref["Lentries"][k, :n0[k]] == 0) for k in range(n)])
    assert nfail == ref["nfail"] == int(failed.sum())
    if nfail:
        assert first == int(np.flatnonzero(failed)[0])
    ok = np.repeat(~failed, n0)
    scale = np.repeat(np.abs(ref["Lentries"]).max(axis=1), n0)
    assert nfail >= 1
    assert (np.abs(gotp - exp)[ok] / scale[ok]).max() < 1e-10
    assert np.all(gotp[~ok] == 0)                                    # a failed row is written as zeros (:64-66)


@pytest.mark.parametrize("defs", [()])
def test_shared_memory_protocol_is_race_free_under_thread_sanitizer(tmp_path, defs):
    """One host thread per CUDA thread, barriers only where the kernel synchronises: an exchange through shared
    memory that no __syncwarp / __syncthreads orders is a data race ThreadSanitizer reports (tests/simt_emu/
    tsan_main.cpp).  Self-test: with the kernel's __syncwarp calls compiled out the same run must be flagged."""
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    base = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-pthread", "-Wno-unknown-pragmas",
            "-I", os.path.join(ROOT, "tests", "simt_emu"), "-I", os.path.join(ROOT, "gpvecchia_b200", "csrc"),
            "-I", os.path.join(ROOT, "include")] + ["-D" + d for d in defs]
    srcs = [os.path.join(ROOT, "tests", "simt_emu", "emu_harness.cpp"), os.path.join(ROOT, "tests", "simt_emu", "tsan_main.cpp")]
    good, bad = str(tmp_path / "emu_tsan"), str(tmp_path / "emu_tsan_nosync")
    if subprocess.run(base + srcs + ["-o", good], capture_output=True).returncode != 0:
        pytest.skip("ThreadSanitizer runtime not available")
    r = subprocess.run([good], capture_output=True, text=True, timeout=600)
    if "unexpected memory mapping" in r.stderr:
        pytest.skip("ThreadSanitizer cannot run in this container")
    assert r.returncode == 0 and "ThreadSanitizer" not in r.stderr + r.stdout, (r.stdout + r.stderr)[-3000:]
    if not defs:
        subprocess.check_call(base + ["-DGPV_EMU_DROP_SYNCWARP"] + srcs + ["-o", bad])
        rb = subprocess.run([bad], capture_output=True, text=True, timeout=600)
        assert "ThreadSanitizer: data race" in rb.stderr + rb.stdout


@pytest.mark.parametrize("family,defs", [(1, ())])
def test_randomised_shapes_masks_and_nuggets_on_the_host(emu_dir, family, defs):
    """Seeded sweep through the emulated band kernel: set sizes p = 17..31 on the P = 31 instantiation, holes
    anywhere, mixed latent / response conditioning, and nuggets that include Inf on response-conditioned
    neighbours (Vecchia-Laplace's missing data, vecchia_laplace_NR.R:108: that neighbour decouples)."""
    L = _build(emu_dir, defs)
    rng = np.random.default_rng(20240601)
    for trial in range(6):
        m = int(rng.integers(16, 31))
        n = 64
        locs, revNN, rcf = _problem(n, m, 2, seed=100 + trial, layout="z", p_drop=float(rng.choice([0.0, 0.15])))
        p = m + 1
        # mixed conditioning inside the valid (last n0) columns; self stays latent
        for k in range(n):
            k0 = int((revNN[k] != 0).sum())
            flip = rng.random(k0 - 1) < 0.3
            rcf[k, p - k0:p - 1][flip] = 1.0
        nug = rng.uniform(0.05, 0.2, n)
        if trial % 2 == 1:
            nug[rng.integers(0, n, 3)] = np.inf
        cp = [float(rng.uniform(0.5, 2.0)), float(rng.uniform(0.2, 0.6)), float(rng.choice([0.5, 1.5, 2.5]))]
        ref = _oracle(locs, revNN, rcf, nug, "matern", cp, mode=1)
        got, _, nfail, _, n0 = _run(L, 8, 31, locs, revNN, rcf, nug, "matern", cp, family=family)
        got = got.reshape(n, p)
        Lr = ref["Lentries"]
        failed = np.array([np.all(Lr[k, :n0[k]] == 0) for k in range(n)])
        assert nfail == ref["nfail"] == int(failed.sum()), (trial, nfail, ref["nfail"])
        ok = ~failed
        scale = np.abs(Lr[ok]).max(axis=1, keepdims=True)
        assert (np.abs(got[ok] - Lr[ok]) / scale).max() < 1e-9, trial
        assert np.all(got[failed] == 0)


@pytest.mark.parametrize("family,G,P,m,d,nu", [(1, 8, 31, 30, 2, 0.8), (1, 8, 31, 30, 2, 1.3), (1, 16, 41, 40, 3, 2.2)])
def test_general_nu_table_path_on_the_host(emu_dir, family, G, P, m, d, nu):
    """General-nu Matern (Matern.cpp:72-83): the coefficient table built by the library's own
    build_cov_table_kernel and read by u_band_kernel<general>, both emulated, against the oracle's
    std::cyl_bessel_k restatement.  Includes duplicated locations (distance 0 -> sigma^2, :76-77) through the
    kernel's slow path."""
    L = _build(emu_dir)
    L.emu_u_band_general.argtypes = [C.c_int] * 5 + [C.c_int64, C.c_int, C.c_int] + [C.c_void_p] * 7 + [C.c_double] * 4
    L.emu_u_band_general.restype = C.c_int
    n = 80
    locs, revNN, rcf = _problem(n, m, d, seed=int(nu * 10) + P)
    locs[50] = locs[20]                      # a duplicate: zero distance inside the later sets
    nug = np.random.default_rng(1).uniform(0.05, 0.15, n)
    cp = [1.4, 0.35, nu]
    ref = _oracle(locs, revNN, rcf, nug, "matern", cp)
    p = m + 1
    nn = np.ascontiguousarray(revNN.astype(np.int32) - 1)
    cond = np.zeros(n, dtype=np.uint64)
    for j in range(p):
        cond |= (np.nan_to_num(rcf[:, j], nan=0.0) == 1.0).astype(np.uint64) << np.uint64(j)
    out = np.full(n * p, np.nan)
    nfail = np.zeros(1, dtype=np.uint64)
    first = np.full(1, np.iinfo(np.int64).max, dtype=np.int64)
    lr = np.ascontiguousarray(locs)
    w_max = float(((locs.max(axis=0) - locs.min(axis=0)) ** 2).sum())
    rc = L.emu_u_band_general(family, G, P, d, 2, n, p, d, lr.ctypes.data, nn.ctypes.data, cond.ctypes.data, nug.ctypes.data,
                              out.ctypes.data, nfail.ctypes.data, first.ctypes.data, cp[0], cp[1], cp[2], w_max)
    assert rc == 0 and int(nfail[0]) == 0 and ref["nfail"] == 0
    got = out.reshape(n, p)
    Lr = ref["Lentries"]
    assert np.array_equal(got == 0, Lr == 0)
    assert (np.abs(got - Lr) / np.abs(Lr).max(axis=1, keepdims=True)).max() < 1e-10
