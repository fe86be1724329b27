#!/usr/bin/env python
"""BF16-path logit error against the oracle over several images: RMS and max per image (the max over 1000
logits is a noisy statistic; the RMS tells whether a kernel change adds error systematically)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402
from oracle import binding  # noqa: E402

pkg = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
blobs = pkg.synth.model_blobs(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "Network"))
imgs = pkg.synth.synthetic_images(n, 224, seed=4242)
ref = binding.Oracle().forward(imgs, blobs)["logits"]
with pkg.Engine(0, 224, pkg.BF16, max_batch=n) as eng:
    eng.load_weights(blobs)
    _, logits = eng.forward(imgs, want_logits=True)
d = logits - ref
print("rms per image:", " ".join(f"{v:.2e}" for v in np.sqrt((d * d).mean(1))))
print("max per image:", " ".join(f"{v:.2e}" for v in np.abs(d).max(1)))
print(f"overall rms {np.sqrt((d * d).mean()):.3e}  max {np.abs(d).max():.3e}  top-1 equal {np.array_equal(logits.argmax(1), ref.argmax(1))}")
