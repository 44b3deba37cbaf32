"""Inputs of the reference-run fixtures (tests/golden/ref_compiled.npz): small instances of every BASELINE
config shape plus the edge cases of SURVEY.md 8a, built with the numpy restatement of vecchia_specify
(oracle/vecchia_np.py) so that neither the fixture generator nor the tests depend on the product library.
Shared by tools/gen_ref_golden.py (which runs the COMPILED REFERENCE, oracle/_ref, on them) and by the tests."""
import numpy as np

import oracle as O


def _inputs(va, nug_obs):
    """Arguments 3..7 of U_NZentries as createU builds them (R/createU.R:74-80,146-147) for a scalar or
    per-observation nugget in identity ordering."""
    prep = va["U_prep"]
    locsord = np.asarray(va["locsord"], dtype=np.float64)
    N = locsord.shape[0]
    obs = np.asarray(va["obs"], dtype=bool)
    n = int(obs.sum())
    nug_obs = np.broadcast_to(np.asarray(nug_obs, dtype=np.float64), (n,)).copy()
    if va["cond_yz"] == "zy":                       # nuggets.all = c(nuggets, 0...), ord.all = c(ord[1:n], ord + n)
        nug_all = np.zeros(N)
        nug_all[:n] = nug_obs
    else:
        nug_all = np.zeros(N)
        nug_all[obs] = nug_obs
    rc = prep["revCond"].astype(np.float64)
    rc[prep["revCond"] < 0] = np.nan
    return dict(n=n, locs=locsord, revNNarray=np.asarray(prep["revNNarray"], dtype=np.int32), revCond=rc,
                nuggets=nug_all, nuggets_obsord=nug_obs)


def cases():
    """name -> dict(n, locs, revNNarray, revCond (double, NaN = NA), nuggets, nuggets_obsord, covType, covparms,
    textbook (bool: run the compiled reference with the published unblocked chol instead of OpenBLAS))."""
    out = {}
    rng = np.random.default_rng(20261017)

    def add(name, va, nug, covType, covparms, textbook=False):
        c = _inputs(va, nug)
        c.update(covType=covType, covparms=np.asarray(covparms, dtype=np.float64), textbook=textbook)
        out[name] = c

    # cfg1: response-first zy, m = 20, Matern 1.5 (BASELINE configs[0])
    locs = rng.random((250, 2))
    add("cfg1_zy_m20_nu15", O.vecchia_specify(locs, 20, cond_yz="zy"), 0.1, "matern", [1.0, 0.25, 1.5])
    # cfg2: m = 30, closed forms, z and y conditioning (configs[1])
    locs = rng.random((400, 2))
    va_z, va_y = O.vecchia_specify(locs, 30, cond_yz="z"), O.vecchia_specify(locs, 30, cond_yz="y")
    tau = 0.05 + 0.1 * rng.random(400)
    for nu in (0.5, 1.5, 2.5):
        add(f"cfg2_z_m30_nu{nu}", va_z, tau, "matern", [1.0, 0.2, nu])
    add("cfg2_y_m30_nu1.5", va_y, tau, "matern", [1.3, 0.2, 1.5])
    # cfg3: general nu (Bessel branch), nu = 0.8 and the 1.3 of the reference's own tests (configs[2])
    add("cfg3_z_m30_nu0.8", va_z, tau, "matern", [1.0, 0.2, 0.8])
    add("cfg3_z_m30_nu1.3", va_z, 0.1, "matern", [1.0, 0.2, 1.3])
    # cfg4: 3-D, m = 40, esqe (configs[3])
    locs3 = rng.random((300, 3))
    add("cfg4_z_m40_d3_esqe", O.vecchia_specify(locs3, 40, cond_yz="z"), 0.1, "esqe", [1.0, 0.4, 0.5, 0.4])
    # cfg5: obs + pred, zy, m = 30 (configs[4])
    lo, lp = rng.random((160, 2)), rng.random((50, 2))
    add("cfg5_zy_pred_m30_nu15", O.vecchia_specify(lo, 30, cond_yz="zy", locs_pred=lp), 0.1, "matern", [1.0, 0.3, 1.5])
    # SGV (the default of vecchia_specify), small m, 1-D and 5-D locations
    locs = rng.random((200, 2))
    add("sgv_m10_nu25", O.vecchia_specify(locs, 10, cond_yz="SGV"), 0.05 + 0.1 * rng.random(200), "matern", [0.9, 0.3, 2.5])
    add("sgv_m7_d1_nu05", O.vecchia_specify(np.sort(rng.random((120, 1)), axis=0), 7, cond_yz="SGV"), 0.2, "matern", [1.0, 0.1, 0.5])
    add("z_m12_d5_nu15", O.vecchia_specify(rng.random((150, 5)), 12, cond_yz="z"), 0.1, "matern", [1.0, 0.8, 1.5])
    # edge cases: a negative nugget (failing Cholesky, row left zero, U_NZentries.cpp:64-66), zero nuggets,
    # Inf nuggets (Vecchia-Laplace missing data, vecchia_laplace_NR.R:108-109), duplicated locations
    locs = rng.random((90, 2))
    locs[40] = locs[12]
    locs[77] = locs[3]
    va = O.vecchia_specify(locs, 9, cond_yz="z")
    nug = np.full(90, 0.1)
    nug[5] = -40.0
    add("edge_negative_nugget", va, nug, "matern", [1.0, 0.3, 1.5], textbook=True)
    nug = np.full(90, 0.1)
    nug[[7, 30]] = 0.0
    add("edge_zero_nugget_dups", va, nug, "matern", [1.0, 0.3, 0.5], textbook=True)
    nug = np.full(90, 0.1)
    nug[[11, 50]] = np.inf
    add("edge_inf_nugget_z", va, nug, "matern", [1.0, 0.3, 1.5], textbook=True)
    add("edge_inf_nugget_sgv", O.vecchia_specify(locs, 9, cond_yz="SGV"), nug, "matern", [1.0, 0.3, 1.5], textbook=True)
    return out
