#!/usr/bin/env python
"""CPU emulation (torch, fp32 arithmetic with bf16 rounding at the engine's rounding points) of the BF16 path,
to measure what a change of the rounding points does to the logit error BEFORE writing the kernel:

  base : LN output rounded to bf16 (what layernorm_kernel feeds the qkv / fc1 GEMMs today)
  fold : LayerNorm folded into the GEMM -- A operand = bf16(x) (the raw residual row), W' = bf16(gamma * W),
         y = rstd * (A W'^T) - rstd * mu * colsum(W') + (b + W beta), statistics from the fp32 row

Both against the fp32 oracle (oracle/vit_oracle.c).  python tools/bf16_fold_emulation.py [n_images]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402
from oracle import binding  # noqa: E402

pkg = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
blobs = pkg.synth.model_blobs(os.path.join(ROOT, "oracle", "_ref", "Network"), 224, seed=0)
imgs = pkg.synth.synthetic_images(n, 224, seed=1234)
ref = binding.Oracle().forward(imgs, blobs)["logits"]
W = [torch.from_numpy(np.asarray(b, np.float32)) for b in blobs]


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def forward(img, mode, fp8=None):
    x = torch.from_numpy(img)
    # patch embedding (TF32 on the GPU: emulate with fp32), tokens [197,768]
    p = x.reshape(3, 14, 16, 14, 16).permute(1, 3, 0, 2, 4).reshape(196, 768)
    tok = p @ W[1].reshape(768, 768).T + W[2]
    xs = torch.cat([W[0][None, :], tok], 0) + W[3].reshape(197, 768)
    stats = {"mu_over_sigma": []}
    for l in range(12):
        b = 4 + 12 * l

        def ln_gemm(xs, g_, be_, w_, bias_):
            mu = xs.mean(1, keepdim=True)
            var = (xs * xs).mean(1, keepdim=True) - mu * mu
            rstd = 1.0 / torch.sqrt(var + 1e-6)
            stats["mu_over_sigma"].append(float((mu.abs() * rstd).max()))
            if mode == "base":
                a = bf((xs - mu) * rstd * g_ + be_)
                return a @ bf(w_).T + bias_
            wf = bf(w_ * g_[None, :])
            cs = wf.sum(1)
            b2 = bias_ + w_ @ be_
            acc = bf(xs) @ wf.T
            return rstd * acc - (rstd * mu) * cs[None, :] + b2[None, :]

        qkv = bf(ln_gemm(xs, W[b], W[b + 1], W[b + 2].reshape(2304, 768), W[b + 3]))
        q, k, v = qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:]
        outs = []
        for h in range(12):
            s = q[:, 64 * h:64 * h + 64] @ k[:, 64 * h:64 * h + 64].T
            m = s.max(1, keepdim=True).values
            pe = torch.exp2((s - m) * (0.125 * 1.4426950408889634))
            o = bf(pe) @ v[:, 64 * h:64 * h + 64]
            outs.append(o / pe.sum(1, keepdim=True))
        att = bf(torch.cat(outs, 1))
        xs = xs + att @ bf(W[b + 4].reshape(768, 768)).T + W[b + 5]
        h1 = ln_gemm(xs, W[b + 6], W[b + 7], W[b + 8].reshape(3072, 768), W[b + 9])
        h1 = bf(0.5 * h1 * (1.0 + torch.erf(h1 * 0.7071067811865476)))
        xs = xs + h1 @ bf(W[b + 10].reshape(768, 3072)).T + W[b + 11]
    c = xs[0]
    mu = c.mean()
    var = (c * c).mean() - mu * mu
    c = (c - mu) / torch.sqrt(var + 1e-6) * W[148] + W[149]
    return (W[150].reshape(1000, 768) @ c + W[151]).numpy(), stats


torch.set_num_threads(8)
for mode in ("base", "fold"):
    errs, ratios = [], []
    for i in range(n):
        lg, st = forward(imgs[i], mode)
        errs.append(np.abs(lg - ref[i]).max())
        ratios.append(max(st["mu_over_sigma"]))
        assert lg.argmax() == ref[i].argmax()
    print(f"{mode}: per-image max|dlogit| max {max(errs):.4e} mean {np.mean(errs):.4e}   max |mu|/sigma over all LN rows {max(ratios):.3f}")
