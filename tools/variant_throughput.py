#!/usr/bin/env python
"""Resident BF16 throughput of the model variants (not the headline bench): python tools/variant_throughput.py [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256


def gflop(img, patch, embed, depth, hidden):
    t = (img // patch) ** 2 + 1
    layer = 2 * t * embed * (3 * embed + embed + 2 * hidden) + 4 * t * t * embed
    return ((t - 1) * 2 * embed * 3 * patch * patch + depth * layer + 2 * embed * 1000) / 1e9


for name, img, b in (("b16", 224, batch), ("b32", 224, 4 * batch), ("s16", 224, 2 * batch), ("l16", 224, batch // 2), ("b32", 384, batch)):
    patch, embed, depth, heads, hidden = pkg.synth.VARIANTS[name]
    blobs = pkg.synth.variant_blobs(name, img, seed=7)
    x = pkg.synth.synthetic_images(b, img, seed=1)
    with pkg.Engine(0, img, pkg.BF16, max_batch=b, model=name) as e:
        e.load_weights(blobs)
        e.stage(x)
        for _ in range(3):
            e.forward_resident(b)
        ms = e.time_resident(b, 10) / 10
    gf = gflop(img, patch, embed, depth, hidden)
    print(f"{name} {img}x{img} batch {b}: {ms:.3f} ms/step, {b / ms * 1e3:.0f} images/s, {gf:.2f} GFLOP/image, {b / ms * gf:.0f} TFLOP/s")
