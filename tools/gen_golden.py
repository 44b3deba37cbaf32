"""Generates the committed golden fixtures under tests/golden/.

The reference (R + Rcpp + RcppArmadillo + BH) cannot run in this image, so no fixture here is an
output of the reference itself.  They are independent known answers:
  matern_general_mpmath.json   sig2/(2^(nu-1) Gamma(nu)) s^nu K_nu(s) from mpmath at 40 digits
                               (the formula of src/Matern.cpp:73,80)
  matern_closed_mpmath.json    the three closed forms of src/Matern.cpp:39,52,68 and Esqe.cpp:34
                               from mpmath at 40 digits
  u_small_quad.npz             U_NZentries of a 60-point SGV problem (m=8) from the __float128
                               arbiter of oracle/ (textbook dpotf2 + back substitution in quad
                               precision on the fp64 inputs), plus the exact Gaussian log-density
Run:  python tools/gen_golden.py
"""
import json
import os
import sys

import mpmath as mp
import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
mp.mp.dps = 40


def matern_general(s, nu, sig2=1.0):
    s, nu = mp.mpf(s), mp.mpf(nu)
    return sig2 / (mp.mpf(2) ** (nu - 1) * mp.gamma(nu)) * s ** nu * mp.besselk(nu, s)


rng = np.random.default_rng(7)
rows = []
for nu in [0.05, 0.3, 0.8, 1.0, 1.3, 1.9999, 2.0, 2.7, 3.5, 5.2, 9.9]:
    ss = np.concatenate([10.0 ** rng.uniform(-7, 0, 12), rng.uniform(1, 3, 8), rng.uniform(3, 40, 6),
                         rng.uniform(40, 500, 3)])
    for s in ss:
        rows.append([float(nu), float(s), float(matern_general(float(s), float(nu)))])
json.dump(dict(doc="sig2=1; value = 2^(1-nu)/Gamma(nu) s^nu K_nu(s)", rows=rows),
          open(os.path.join(OUT, "matern_general_mpmath.json"), "w"))

closed = []
for s in np.concatenate([10.0 ** rng.uniform(-6, 0, 10), rng.uniform(1, 30, 10)]):
    x = mp.mpf(float(s))
    closed.append(dict(s=float(s),
                       nu05=float(mp.e ** (-x)),
                       nu15=float((1 + mp.sqrt(3) * x) * mp.e ** (-mp.sqrt(3) * x)),
                       nu25=float(mp.e ** (-mp.sqrt(5) * x) * (1 + mp.sqrt(5) * x + 5 * x * x / 3)),
                       esqe=float(mp.mpf("0.7") * mp.e ** (-x / mp.mpf("0.5")) +
                                  mp.mpf("0.4") * mp.e ** (-(x / mp.mpf("1.5")) ** 2))))
json.dump(dict(doc="sig2=1, range=1 (esqe: covparms=(0.7,0.5,0.4,1.5)); argument is the distance",
               rows=closed), open(os.path.join(OUT, "matern_closed_mpmath.json"), "w"))

import oracle as O  # noqa: E402

n, m = 60, 8
locs = rng.random((n, 2))
z = rng.standard_normal(n)
va = O.vecchia_specify(locs, m, cond_yz="SGV")
prep = va["U_prep"]
rc = prep["revCond"].astype(np.float64)
rc[prep["revCond"] < 0] = np.nan
nug = 0.05 + 0.1 * rng.random(n)
fix = dict(locs=locs, z=z, revNNarray=prep["revNNarray"], revCond=prep["revCond"], nuggets=nug,
           rowpointers=prep["rowpointers"], colindices=prep["colindices"])
for tag, ct, cp in [("m05", "matern", [1.2, 0.3, 0.5]), ("m15", "matern", [1.2, 0.3, 1.5]),
                    ("m25", "matern", [1.2, 0.3, 2.5]), ("g08", "matern", [1.2, 0.3, 0.8]),
                    ("g13", "matern", [1.2, 0.3, 1.3]), ("esqe", "esqe", [0.7, 0.5, 0.4, 1.5])]:
    r = O.U_NZentries(1, n, locs, prep["revNNarray"], rc, nug, nug, ct, np.array(cp), mode=2)
    fix["L_" + tag] = r["Lentries"]
    fix["cp_" + tag] = np.array(cp)
va_full = O.vecchia_specify(locs, n - 1, cond_yz="SGV")
fix["exact_loglik_m15"] = np.array(O.exact_loglik(z, locs, [1.2, 0.3, 1.5], nug))
np.savez_compressed(os.path.join(OUT, "u_small_quad.npz"), **fix)
print("wrote", os.listdir(OUT))
