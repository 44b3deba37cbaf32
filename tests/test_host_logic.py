"""Host-side logic that needs no GPU: vectorised U_sparsity vs the loop restatement of
R/U_sparsity.R, the harness' ordered neighbour search vs brute force, the zy layout vs the
restatement of R/vecchia_specify.R:191-224, and row sharding."""
import numpy as np
import pytest

import oracle as O
from gpvecchia_b200 import harness as H
from gpvecchia_b200.host import U_sparsity
from gpvecchia_b200 import shard


@pytest.mark.parametrize("cond_yz", ["y", "z", "SGV", "zy"])
def test_u_sparsity_vectorised_is_bit_identical(cond_yz):
    rng = np.random.default_rng(0)
    locs = rng.random((70, 2))
    va = O.vecchia_specify(locs, 6, cond_yz=cond_yz)
    ref = va["U_prep"]
    got = U_sparsity(va["locsord"], va["NNarray"], va["obs"], va["Cond"])
    for k in ("revNNarray", "revCond", "rowpointers", "colindices", "y_ind", "observed_map"):
        assert np.array_equal(got[k], ref[k]), k
    assert got["size"] == ref["size"]


def test_u_sparsity_with_prediction_locations():
    rng = np.random.default_rng(1)
    locs, lp = rng.random((40, 2)), rng.random((15, 2))
    for cyz in ("SGV", "zy", "y"):
        va = O.vecchia_specify(locs, 5, cond_yz=cyz, locs_pred=lp)
        got = U_sparsity(va["locsord"], va["NNarray"], va["obs"], va["Cond"])
        for k in ("rowpointers", "colindices", "y_ind", "observed_map"):
            assert np.array_equal(got[k], va["U_prep"][k]), (cyz, k)


def test_ordered_nn_kdtree_matches_brute_force():
    locs = H.make_locs(6000, 2, stream=3)
    m = 12
    got = H.ordered_nn_kdtree(locs, m)
    ref = O.find_ordered_nn_brute(locs[:700], m)
    assert np.array_equal(got[:700], ref)
    # rows beyond the brute-force block: check the defining property on a sample
    for i in (4097, 5000, 5999):
        dd = np.sqrt(((locs[:i] - locs[i]) ** 2).sum(1))
        assert np.array_equal(got[i, 1:] - 1, np.argsort(dd, kind="stable")[:m])
    sl = H.ordered_nn_kdtree(locs, m, row_begin=4500, row_end=5200)
    assert np.array_equal(sl, got[4500:5200])


def test_ordered_nn_3d():
    locs = H.make_locs(900, 3, stream=4)
    assert np.array_equal(H.ordered_nn_kdtree(locs, 7), O.find_ordered_nn_brute(locs, 7))


def test_layout_zy_matches_specify_restatement():
    locs = H.make_locs(300, 2, stream=5)
    va = O.vecchia_specify(locs, 9, cond_yz="zy")
    locs2, NN, Cond, obs = H.layout_zy(locs, 9, 300)
    assert np.array_equal(NN, va["NNarray"]) and np.array_equal(Cond, va["Cond"])
    assert np.array_equal(obs, va["obs"]) and np.array_equal(locs2, va["locsord"])


def test_row_sharding_partitions_full_rows_evenly():
    n0 = np.concatenate([np.ones(1000, dtype=np.int64), np.full(1000, 31)])   # zy-like: trivial then full
    for world in (1, 2, 3, 8):
        cuts = shard.row_cuts(n0, world)
        assert cuts[0] == 0 and cuts[-1] == n0.size and np.all(np.diff(cuts) >= 0)
        w = (n0.astype(np.float64) ** 3)
        loads = [w[cuts[r]:cuts[r + 1]].sum() for r in range(world)]
        assert max(loads) <= 1.02 * (w.sum() / world) + 31 ** 3
    assert shard.row_cuts(np.full(10, 5), 4).tolist() == [0, 3, 5, 8, 10]


def test_layout_zy_pred_matches_specify_restatement():
    locs = H.make_locs(250, 2, stream=6)
    lp = H.make_locs(60, 2, stream=7)
    va = O.vecchia_specify(locs, 8, cond_yz="zy", locs_pred=lp)
    locs2, NN, Cond, obs = H.layout_zy_pred(locs, lp, 8, use_gpu=False)
    assert np.array_equal(NN, va["NNarray"]) and np.array_equal(Cond, va["Cond"])
    assert np.array_equal(obs, va["obs"]) and np.array_equal(locs2, va["locsord"])


@pytest.mark.parametrize("seed,n,m,npred", [(0, 120, 5, 0), (1, 300, 12, 0), (2, 90, 30, 0), (3, 150, 7, 40), (4, 40, 3, 25)])
def test_native_whichCondOnLatent_matches_restated_R_loop(seed, n, m, npred):
    # gpv_whichCondOnLatent (host C++, merge counts) vs the line-by-line restatement of
    # R/whichCondOnLatent.R:2-27 in the oracle, with and without prediction locations (firstind.pred)
    rng = np.random.default_rng(seed)
    locs = rng.random((n + npred, 2))
    NN = O.find_ordered_nn_brute(locs, m)
    first = n + 1 if npred else None
    ref = O.whichCondOnLatent(NN, first)
    got = H.whichCondOnLatent(NN, first)
    assert got.shape == ref.shape and np.array_equal(got, ref)


def test_native_whichCondOnLatent_scales():
    # n = 2e5, m = 30 in a few seconds (the R loop is O(n m^3) interpreter calls): properties only
    n, m = 200_000, 30
    locs = H.make_locs(n, 2, stream=9)
    NN = H.ordered_nn_kdtree(locs, m)
    C_ = H.whichCondOnLatent(NN)
    assert np.all(C_[:, 0] == 1) and np.array_equal(C_ == -1, NN == 0)
    frac = (C_[m + 1:, 1:] == 1).mean()
    assert 0.05 < frac < 0.95          # SGV conditions on a mix of y and z


def test_uniform_cuts_cover_the_rows():
    # bench.py shards 'z' layouts by plain equal ranges (the library's locality layer makes late rows cost what
    # early ones do); zy layouts by sum n0^3 (row_cuts above)
    for world in (1, 2, 3, 8):
        cuts = shard.uniform_cuts(1_000_003, world)
        assert cuts[0] == 0 and cuts[-1] == 1_000_003 and cuts.size == world + 1
        assert np.diff(cuts).max() - np.diff(cuts).min() <= 1


def test_pair_stage_short_sqrt_exp_sequences_on_the_host(tmp_path):
    """csrc/u_kernels.cuh neg_sqrt_fast_n / exp_negarg_fast_n restated with <cmath> fma and swept over
    the MUFU.RSQ64H seed envelope against __float128 (tools/check_pair_fast.cpp): sqrt below one ulp,
    exp within 1.25 x 2^-53 absolute, Matern-1.5 covariance no worse than 1.5 x the longer sequences."""
    import os
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "check_pair_fast")
    subprocess.check_call(["g++", "-O2", "-o", exe, os.path.join(root, "tools", "check_pair_fast.cpp"), "-lquadmath"])
    out = subprocess.run([exe, "200000"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
