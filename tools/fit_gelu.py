"""Fits the two one-MUFU GELU forms of vit-2spn_b200/csrc/tc_math.cuh (run here, CPU only: numpy + scipy).

forward     gelu(x)  = relu(x) - |x| Phi(-|x|),   Phi(-a) = 0.5 exp2(q(a)),  q(a) = a (c0 + c1 a + ... + c4 a^4)
            q is fitted minimax on the ABSOLUTE error of a Phi(-a) over a in [0, 14] (p-norm continuation from a
            least-squares fit of log2 of the normal tail); fp32 evaluation: |gelu error| < 7.1e-7 for all x.
derivative  gelu'(x) = m(|x|) for x <= 0, 1 - m(|x|) for x > 0,  m(a) = Phi(-a) - a phi(a) = exp2(q(a)) T(a)
            T (degree 5) is fitted by Lawson-reweighted least squares on the absolute error of m with q fixed;
            fp32 evaluation: |gelu' error| < 2.6e-6.

    python tools/fit_gelu.py          # prints the coefficients used in tc_math.cuh and the achieved errors
"""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import erf, erfc, log_ndtr

LOG2E = np.log2(np.e)


def horner(c, a):
    r = np.zeros_like(a)
    for cc in c[::-1]:
        r = r * a + cc
    return r


def fit_q(deg=5):
    a = np.concatenate([np.linspace(0, 6, 6001), np.linspace(6, 14, 801)[1:]])
    target = a * 0.5 * erfc(a / np.sqrt(2))
    m = a < 5
    V = np.vander(a[m], deg + 1, increasing=True)[:, 1:]
    c, *_ = np.linalg.lstsq(V, (log_ndtr(-a[m]) - np.log(0.5)) * LOG2E, rcond=None)

    def resid(c):
        return a * 0.5 * np.exp2(horner(c, a) * a) - target
    for pnorm in (2, 4, 8, 16, 32, 64):
        c = least_squares(lambda c: np.sign(resid(c)) * np.abs(resid(c) * 1e4) ** (pnorm / 2), c, method="lm",
                          max_nfev=4000, xtol=1e-15, ftol=1e-15).x
    return c


def fit_t(c, deg=5):
    a = np.concatenate([np.linspace(0, 7, 14001), np.linspace(7, 14, 701)[1:]])
    e = np.exp2(horner(c, a) * a)
    target = 0.5 * erfc(a / np.sqrt(2)) - a * np.exp(-a * a / 2) / np.sqrt(2 * np.pi)
    V = np.vander(a, deg + 1, increasing=True) * e[:, None]
    w = np.ones_like(a)
    for _ in range(300):
        t, *_ = np.linalg.lstsq(V * w[:, None], target * w, rcond=None)
        err = V @ t - target
        w = w * (np.abs(err) / np.abs(err).max() + 1e-3) ** 0.5
        w /= w.max()
    return t


def check(c, t):
    x = np.linspace(-14, 14, 560001).astype(np.float32)
    ax = np.abs(x)
    c32, t32 = c.astype(np.float32), t.astype(np.float32)
    r = np.float32(c32[-1]) * np.ones_like(ax)
    for cc in c32[-2::-1]:
        r = (r * ax + cc).astype(np.float32)
    e = np.exp2((r * ax).astype(np.float32)).astype(np.float32)
    g = (np.maximum(x, 0) - np.float32(0.5) * ax * e).astype(np.float32)
    T = np.float32(t32[-1]) * np.ones_like(ax)
    for cc in t32[-2::-1]:
        T = (T * ax + cc).astype(np.float32)
    m = (T * e).astype(np.float32)
    d = np.where(x > 0, np.float32(1) - m, m)
    xd = x.astype(np.float64)
    ref_g = 0.5 * xd * (1 + erf(xd / np.sqrt(2)))
    ref_d = 0.5 * (1 + erf(xd / np.sqrt(2))) + xd * np.exp(-xd * xd / 2) / np.sqrt(2 * np.pi)
    return np.abs(g - ref_g).max(), np.abs(d - ref_d).max(), float((r * ax).max())


if __name__ == "__main__":
    c = fit_q()
    t = fit_t(c)
    eg, ed, qmax = check(c, t)
    print("q coefficients c0..c4 (exponent in log2 units):", ", ".join(f"{v:.8e}" for v in c))
    print("T coefficients t0..t5:", ", ".join(f"{v:.8e}" for v in t))
    print(f"fp32 evaluation over [-14, 14]: max |gelu error| {eg:.3g}, max |gelu' error| {ed:.3g}, max exponent {qmax:.3g} (<= 0)")
