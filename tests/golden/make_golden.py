"""Generate tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref = the
reference's ViT_seq.c compiled unmodified by oracle/Makefile).  Run in the build
container, where /root/reference exists:

    python tests/golden/make_golden.py

The vectors pin oracle/vit_oracle.c (tests/test_oracle.py) and, through it, the
CUDA path.  Inputs are regenerated from seeds by the tests, only outputs are
stored.  The reference's own golden files (Data/answer_result*.txt) cannot be
reproduced because 36 weight blobs are absent from the checkout.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402
from oracle import binding  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
pkg = g.load_package()
synth = pkg.synth


def stage_inputs():
    """seeded inputs of the per-stage vectors (shared with tests/test_oracle.py)"""
    rng = np.random.default_rng(11)
    d = {}
    d["lin_x"] = rng.standard_normal((5, 768), dtype=np.float32)
    d["lin_w"] = rng.standard_normal((40, 768), dtype=np.float32) * np.float32(0.05)
    d["lin_b"] = rng.standard_normal(40, dtype=np.float32)
    d["ln_x"] = rng.standard_normal((197, 768), dtype=np.float32) * np.float32(2.0) + np.float32(0.5)
    d["ln_g"] = rng.standard_normal(768, dtype=np.float32) * np.float32(0.1) + np.float32(1.0)
    d["ln_b"] = rng.standard_normal(768, dtype=np.float32) * np.float32(0.1)
    d["mha_x"] = rng.standard_normal((197, 768), dtype=np.float32)
    d["mha_win"] = rng.standard_normal((2304, 768), dtype=np.float32) * np.float32(0.04)
    d["mha_bin"] = rng.standard_normal(2304, dtype=np.float32) * np.float32(0.1)
    d["mha_wout"] = rng.standard_normal((768, 768), dtype=np.float32) * np.float32(0.04)
    d["mha_bout"] = rng.standard_normal(768, dtype=np.float32) * np.float32(0.1)
    return d


def main():
    binding.build()
    ref = binding.Reference(224)
    out = {}

    # 1. the bundled image through the bundled + seed-0 model (label 606 / 0.007682)
    net_dir = os.path.join(ROOT, "oracle", "_ref", "Network")
    blobs = synth.model_blobs(net_dir, 224, seed=0)
    img = synth.load_image_file(os.path.join(ROOT, "oracle", "_ref", "Data", "input-1.bin"))
    out["bundled_input1_probs"] = ref.forward(img, blobs)

    # 2. fully synthetic model + images (reproducible without the reference data)
    sblobs = synth.model_blobs(None, 224, seed=7)
    simgs = synth.synthetic_images(2, 224, seed=1234)
    out["synth_probs"] = ref.forward(simgs, sblobs)

    # 3. stage functions of the reference
    d = stage_inputs()
    out["lin_y"] = ref.linear(d["lin_x"], d["lin_w"], d["lin_b"])
    out["ln_y_rows"] = ref.layer_norm(d["ln_x"], d["ln_g"], d["ln_b"])[[0, 1, 196]]
    y = ref.mha(d["mha_x"], d["mha_win"], d["mha_bin"], d["mha_wout"], d["mha_bout"])
    out["mha_y_rows"] = y[[0, 100, 196]]
    out["mha_y_sum"] = np.array([y.astype(np.float64).sum()])

    # 4. 384x384 / 577 tokens (ViT_seq.c built with img_size 384)
    ref384 = binding.Reference(384)
    b384 = synth.model_blobs(None, 384, seed=7)
    i384 = synth.synthetic_images(1, 384, seed=4321)
    out["synth384_probs"] = ref384.forward(i384, b384)

    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, float(np.abs(v).max()))


def variants():
    """model variants the reference expresses by editing its macros (oracle/Makefile streams the edit
    into gcc): ViT-B/32 (patch_size 32) and ViT-S/16 (embed_dim 384, num_heads 6), 2 synthetic images each"""
    binding.build()
    out = {}
    for name in ("b32", "s16"):
        ref = binding.Reference(224, name)
        blobs = synth.variant_blobs(name, 224, seed=7)
        imgs = synth.synthetic_images(2, 224, seed=1234)
        out[f"{name}_probs"] = ref.forward(imgs, blobs)
    np.savez_compressed(os.path.join(HERE, "variant_vectors.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, float(np.abs(v).max()), v.argmax(1))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "variants":
        variants()
    else:
        main()
