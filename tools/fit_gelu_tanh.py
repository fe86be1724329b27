#!/usr/bin/env python
"""Fit erf(x/sqrt2) ~= tanh(x * (a + b w + c w^2)), w = min(x^2, U): the GELU form built on MUFU.TANH
(gelu_erf_tanh2, csrc/tc_common.cuh).  Minimises the max GELU error 0.5|x| |tanh(..) - erf(..)|."""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import erf

import sys
DEG = int(sys.argv[1]) if len(sys.argv) > 1 else 2
U = float(sys.argv[2]) if len(sys.argv) > 2 else 20.25
X = np.sqrt(U)
n = 8000
x = (np.cos(np.pi * (np.arange(n) + 0.5) / n) + 1) / 2 * 6.0 + 1e-9
target = erf(x / np.sqrt(2))


def model(c, x):
    w = np.minimum(x * x, U)
    return np.tanh(x * np.polyval(c, w))


c0 = np.zeros(DEG + 1)
c0[-1] = np.sqrt(2 / np.pi)
if DEG >= 1:
    c0[-2] = np.sqrt(2 / np.pi) * 0.044715
wgt = np.ones_like(x)
c = c0
for it in range(60):
    c = least_squares(lambda cc: (model(cc, x) - target) * 0.5 * x * wgt, c, xtol=1e-15, ftol=1e-15, gtol=1e-15).x
    r = np.abs((model(c, x) - target) * 0.5 * x)
    wgt = wgt * (1 + 3 * r / r.max())
    wgt /= wgt.mean()
cf = c.astype(np.float32)
print("coeffs (highest first):", [float(v) for v in cf], "U =", U)
xx = np.linspace(-10, 10, 2000001).astype(np.float32)
w = np.minimum(xx * xx, np.float32(U)).astype(np.float32)
p = np.full_like(xx, cf[0])
for k in range(1, DEG + 1):
    p = (p * w + cf[k]).astype(np.float32)
t = np.tanh((xx * p).astype(np.float32).astype(np.float64))
hx = np.float32(0.5) * xx
g = hx * t + hx
xd = xx.astype(np.float64)
print("max |erf err| %.2e   max |GELU err| %.2e" % (np.abs(t - erf(xd / np.sqrt(2))).max(),
      np.abs(g - 0.5 * xd * (1 + erf(xd / np.sqrt(2)))).max()))
# with tanh.approx (rel err 2^-11 worst case) on top:
print("plus tanh.approx 2^-11 rel: GELU err <= %.2e" % (np.abs(g - 0.5 * xd * (1 + erf(xd / np.sqrt(2)))).max() + (np.abs(hx * t) * 2.0**-11).max()))
