#!/usr/bin/env python
"""Fit + check the rational erf used by gelu_erf_fast2 (csrc/tc_common.cuh):
erf(x/sqrt2) = xc * P3(u)/Q3(u), xc = clamp(x, +-A), A = 3.2*sqrt(2), u = 2 xc^2/A^2 - 1.
Prints the coefficients (highest degree first) and the max errors of erf and GELU when evaluated in
float32 the way the kernel does."""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import erf

A = 3.2 * np.sqrt(2.0)
DP = DQ = 3
n = 6000
t = (np.cos(np.pi * (np.arange(n) + 0.5) / n) + 1) / 2 * A * A
x = np.sqrt(t)
y = np.where(x > 1e-12, erf(x / np.sqrt(2)) / np.maximum(x, 1e-300), np.sqrt(2 / np.pi))
u = 2 * t / (A * A) - 1


def model(c, u):
    return np.polyval(c[:DP + 1], u) / np.polyval(np.concatenate([c[DP + 1:], [1.0]]), u)


Am = np.concatenate([np.vander(u, DP + 1), -(y[:, None]) * np.vander(u, DQ + 1)[:, :-1]], axis=1)
c = np.linalg.lstsq(Am * x[:, None], y * x, rcond=None)[0]
c = least_squares(lambda cc: (model(cc, u) - y) * x, c, xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=20000).x
w = np.ones_like(x)
for _ in range(40):  # reweighting towards a minimax fit
    r = np.abs((model(c, u) - y) * x)
    w = w * (1 + 4 * r / r.max())
    w /= w.mean()
    c = least_squares(lambda cc: (model(cc, u) - y) * x * w, c, xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=5000).x
cf = c.astype(np.float32)
print("A =", A, " 2/A^2 =", 2 / (A * A))
print("P:", [float(v) for v in cf[:DP + 1]])
print("Q:", [float(v) for v in cf[DP + 1:]] + [1.0])
xx = np.linspace(-8, 8, 800001).astype(np.float32)
xc = np.clip(xx, np.float32(-A), np.float32(A))
uu = (xc * xc * np.float32(2 / (A * A)) - np.float32(1)).astype(np.float32)
P = np.full_like(uu, cf[0])
for k in range(1, DP + 1):
    P = (P * uu + cf[k]).astype(np.float32)
qc = np.concatenate([cf[DP + 1:], [np.float32(1)]]).astype(np.float32)
Q = np.full_like(uu, qc[0])
for k in range(1, DQ + 1):
    Q = (Q * uu + qc[k]).astype(np.float32)
e = (xc * (P * (np.float32(1) / Q))).astype(np.float32)
xd = xx.astype(np.float64)
hx = np.float32(0.5) * xx
g = (hx * e + hx).astype(np.float32)
print("max |erf error|  %.2e" % np.abs(e - erf(xd / np.sqrt(2))).max())
print("max |GELU error| %.2e" % np.abs(g - 0.5 * xd * (1 + erf(xd / np.sqrt(2)))).max())
