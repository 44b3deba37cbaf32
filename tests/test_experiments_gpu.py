"""Opt-in GPU checks of the kernels that are in the tree as experiments and not on the default
path: skipped unless GPV_TEST_EXPERIMENTS=1, because they have only run on the host emulation so far
(tests/test_simt_emu.py) and hand work between warps through spin-waits.  First GPU contact:

    GPV_TEST_EXPERIMENTS=1 timeout 300 python -m pytest tests/test_experiments_gpu.py -m gpu -x -q

The warp-specialised kernel (csrc/u_band_ws.cuh, selected with GPV_KERNEL_FAMILY=ws) must give the default
kernel's values bit for bit: same pairs, same arithmetic per pair, same factorisation text."""
import os

import numpy as np
import pytest

import gpvecchia_b200 as G
from gpvecchia_b200 import harness as H

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("GPV_TEST_EXPERIMENTS") != "1",
                                 reason="experimental kernels: opt in with GPV_TEST_EXPERIMENTS=1"),
              pytest.mark.timeout(240)]


def _values(va, nug, covType, cp, z=None, family=None):
    prep = va["U_prep"]
    old = os.environ.pop("GPV_KERNEL_FAMILY", None)
    if family:
        os.environ["GPV_KERNEL_FAMILY"] = family
    try:
        with G.UHandle(va["locsord"], prep["revNNarray"], prep["revCond"], obs=np.ones(len(nug), dtype=bool)) as h:
            packed, nfail, _ = h.values_packed(covType, cp, nug, nug)
            name = h.last_kernel_name()
            ll = h.loglik_numerator(covType, cp, nug, nug, z) if z is not None else None
    finally:
        os.environ.pop("GPV_KERNEL_FAMILY", None)
        if old is not None:
            os.environ["GPV_KERNEL_FAMILY"] = old
    return packed, nfail, name, ll


@pytest.mark.parametrize("m,covType,cp", [(30, "matern", [1.0, 0.05, 1.5]), (30, "matern", [1.3, 0.04, 0.5]),
                                          (27, "esqe", [0.7, 0.05, 0.3, 0.04]), (25, "matern", [1.0, 0.05, 2.5]),
                                          (30, "matern", [1.0, 0.05, 0.8])])
def test_warp_specialised_kernel_matches_the_default_kernel(m, covType, cp):
    n = 40000                                   # 148 blocks x 16 sets per pass: every slot is reused many times
    locs = H.make_locs(n, 2, stream=77)
    NN = H.ordered_nn_kdtree(locs, m)
    va = H.make_vecchia_approx(locs, NN, H.layout_yz(NN, "z"), np.ones(n, dtype=bool), "z")
    nug = H.make_nuggets(n, stream=77)
    z = H.make_data(n, stream=77)
    base, nf0, name0, ll0 = _values(va, nug, covType, cp, z=z)
    got, nf1, name1, ll1 = _values(va, nug, covType, cp, z=z, family="ws")
    assert name0.startswith("u_band<") and name1.startswith("u_band_ws<"), (name0, name1)
    assert nf0 == nf1 == 0
    assert np.array_equal(base, got)
    assert np.allclose(np.asarray(ll0[:2], dtype=float), np.asarray(ll1[:2], dtype=float), rtol=1e-12, atol=0)
