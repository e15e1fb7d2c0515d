"""Fit of the one-MUFU GELU used by the GEMM epilogues (csrc/uav_common.cuh: gelu_erf_f): Phi(-x) = 2^p(x) on [0, 6],
p a degree-8 polynomial in t = x / 3 - 1; prints the monomial coefficients (highest first) and the error of an fp32 Horner
evaluation against float64 erf.  Needs numpy + scipy (build container only; not used at run time)."""
import numpy as np
from numpy.polynomial import chebyshev as C
from scipy.special import erf, erfc

X, DEG = 6.0, 8
xs = np.linspace(0, X, 40001)
phi = 0.5 * erfc(xs / np.sqrt(2))
mono = C.cheb2poly(C.chebfit(2 * xs / X - 1, np.log2(phi), DEG, w=phi + 1e-4))
print("coefficients (t^8 ... t^0):", [float(np.float32(m)) for m in mono[::-1]])


def gelu32(x):
    x = x.astype(np.float32)
    ax = np.minimum(np.abs(x), np.float32(X))
    t = ax * np.float32(2 / X) + np.float32(-1)
    p = np.full_like(t, np.float32(mono[-1]))
    for k in range(len(mono) - 2, -1, -1):
        p = (p * t + np.float32(mono[k])).astype(np.float32)
    s = np.exp2(p.astype(np.float64)).astype(np.float32)
    return (np.maximum(x, np.float32(0)) - (ax * s).astype(np.float32)).astype(np.float32)


xx = np.concatenate([np.linspace(-12, 12, 200001), np.random.RandomState(0).randn(200000) * 2])
ref = xx * 0.5 * (1 + erf(xx / np.sqrt(2)))
err = np.abs(gelu32(xx) - ref)
print(f"max abs error of the fp32 evaluation: {err.max():.3e} at x = {xx[err.argmax()]:.3f}")
