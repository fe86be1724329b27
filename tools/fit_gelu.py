#!/usr/bin/env python
"""Fit + check the polynomial erf used by gelu_erf_fast (csrc/tc_common.cuh).
erf(z) = z*q(u), u = 2 z^2/a^2 - 1, |z| <= a = 3.2; prints the coefficients (highest degree first)
and the max errors of erf and GELU when evaluated in float32 exactly as the kernel does."""
import numpy as np
from numpy.polynomial import chebyshev as Ch
from scipy.special import erf

A, DEG = 3.2, 10
n = 6000
t = (np.cos(np.pi * (np.arange(n) + 0.5) / n) + 1) / 2 * A * A
z = np.sqrt(t)
y = np.where(z > 1e-12, erf(z) / np.maximum(z, 1e-300), 2 / np.sqrt(np.pi))
mono = Ch.cheb2poly(Ch.chebfit(2 * t / (A * A) - 1, y, DEG, w=z + 1e-3)).astype(np.float32)
print("coefficients, highest degree first:", [float(m) for m in mono[::-1]])
x = np.linspace(-8, 8, 800001).astype(np.float32)
zc = np.clip((x * np.float32(0.70710678)).astype(np.float32), np.float32(-A), np.float32(A))
u = (zc * zc * np.float32(2 / (A * A)) - np.float32(1)).astype(np.float32)
q = np.full_like(u, mono[-1])
for k in range(len(mono) - 2, -1, -1):
    q = (q * u + mono[k]).astype(np.float32)
e = (zc * q).astype(np.float32)
hx = (np.float32(0.5) * x).astype(np.float32)
g = (hx * e + hx).astype(np.float32)
xd = x.astype(np.float64)
print("max |erf error|  %.2e" % np.abs(e - erf(xd / np.sqrt(2))).max())
print("max |GELU error| %.2e" % np.abs(g - 0.5 * xd * (1 + erf(xd / np.sqrt(2)))).max())
