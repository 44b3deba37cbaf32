"""Which (intervals per octave, degree) does the general-nu coefficient table need?  Chebyshev interpolation of
2^(1-nu)/Gamma(nu) s^nu K_nu(s) as a function of w = s^2 on geometric sub-intervals, converted to monomials and
evaluated by Horner in fp64, against mpmath (30 digits).  REL = False: absolute error relative to sigma^2 (what a
covariance matrix needs); REL = True: relative error.  Result (bessel_table.cuh): (2, 10), (3, 8), (4, 7) reach the
2e-15 rounding floor of the evaluation in the absolute metric for nu in [0.05, 10], s <= 16."""
import numpy as np, mpmath as mp
REL = False
from numpy.polynomial import chebyshev as Ch
mp.mp.dps = 30
def f_exact(w, nu, inv_range=1.0):
    s = mp.sqrt(w) * inv_range
    return float(mp.mpf(2) ** (1 - nu) / mp.gamma(nu) * s ** nu * mp.besselk(nu, s))
def test(nu, sub_bits, deg, octaves=range(-12, 3)):
    worst = 0.0
    nsub = 1 << sub_bits
    for e in octaves:
        for j in range(nsub):
            lo = 2.0 ** e * (1 + j / nsub); hi = 2.0 ** e * (1 + (j + 1) / nsub)
            c, h = 0.5 * (lo + hi), 0.5 * (hi - lo)
            k = np.arange(deg + 1)
            x = np.cos(np.pi * (k + 0.5) / (deg + 1))
            fv = np.array([f_exact(c + h * xx, nu) for xx in x])
            coef = Ch.chebfit(x, fv, deg)
            mono = Ch.cheb2poly(coef)
            xt = np.linspace(-1, 1, 7)
            for xx in xt:
                # Horner in double
                acc = 0.0
                for a in mono[::-1]:
                    acc = acc * xx + a
                ex = f_exact(c + h * xx, nu)
                worst = max(worst, abs(acc - ex) / (abs(ex) if REL else 1.0))
    return worst
for nu in ():
    for sb, degs in ((1, (13, 15, 17, 19)), (2, (10, 12, 14)), (3, (8, 9, 10, 11))):
        print(nu, sb, [(d, f"{test(nu, sb, d):.1e}") for d in degs])
print("---- edges")
for nu in (0.05, 0.3, 1.3, 3.5, 9.9):
    for sb, degs in ((2, (8, 9, 10)), (3, (6, 7, 8)), (4, (5, 6, 7))):
        print(nu, sb, [(d, f"{test(nu, sb, d, octaves=range(-16, 8, 2)):.1e}") for d in degs])
