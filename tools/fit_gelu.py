import numpy as np
from scipy.special import erf, erfc, log_ndtr
from scipy.optimize import least_squares
a = np.concatenate([np.linspace(0, 6, 6001), np.linspace(6, 14, 801)[1:]])
target = a*0.5*erfc(a/np.sqrt(2))
LOG2E = np.log2(np.e)
def model(c, a):
    # q(a) = a*(c0 + c1 a + ... ) in log2 domain
    r = np.zeros_like(a)
    for cc in c[::-1]:
        r = r*a + cc
    return r*a
def resid(c):
    return (a*0.5*np.exp2(model(c, a)) - target)
def fit(deg):
    # init: fit log2(erfc) by LSQ on [0,5]
    m = a < 5
    lt = (log_ndtr(-a[m]) - np.log(0.5))*LOG2E
    V = np.vander(a[m], deg+1, increasing=True)[:,1:]
    c0, *_ = np.linalg.lstsq(V, lt, rcond=None)
    # minimax via p-norm continuation
    c = c0
    for pnorm in (2, 4, 8, 16, 32, 64):
        f = lambda c: np.sign(resid(c))*np.abs(resid(c)*1e4)**(pnorm/2)
        c = least_squares(f, c, method="lm", max_nfev=4000, xtol=1e-15, ftol=1e-15).x
    return c
for deg in range(3, 8):
    c = fit(deg)
    c32 = c.astype(np.float32)
    x = np.linspace(-14, 14, 560001).astype(np.float32)
    ax = np.abs(x)
    r = np.float32(c32[-1])*np.ones_like(ax)
    for cc in c32[-2::-1]:
        r = (r*ax + cc).astype(np.float32)
    q = (r*ax).astype(np.float32)
    e = np.exp2(q).astype(np.float32)
    g = (np.maximum(x, 0) - np.abs(np.float32(0.5)*x)*e).astype(np.float32)
    ref = 0.5*x.astype(np.float64)*(1+erf(x.astype(np.float64)/np.sqrt(2)))
    err = np.abs(g-ref)
    print("B deg", deg, "fp32 max abs err", err.max(), "at", x[err.argmax()], "q(14)=", q[-1], "max q", q.max())
    print("   coeffs", repr(c))
