"""GPU parity tests of the whole path, through the reference-facing C-ABI
(ViT_opencl and the engine API of include/vit_b200.h), against the CPU oracle
and the reference-generated golden vectors.

Stated tolerances (BASELINE.json north_star):
  FP32 path : max|logit - ref| <= 1e-4 * max|ref logit|
  BF16 path : max|logit - ref| <= 2e-2 (absolute) and identical top-1 on every image
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = np.load(os.path.join(HERE, "golden", "reference_vectors.npz"))

FP32_REL = 1e-4
BF16_ABS = 2e-2


@pytest.fixture(scope="module")
def case224(pkg, oracle, blobs224):
    """5 images (a ragged batch for max_batch=2/4) and their oracle outputs incl. stage dump"""
    imgs = pkg.synth.synthetic_images(5, 224, seed=1234)
    ref = oracle.forward(imgs, blobs224, want_stages=True)
    return imgs, ref


def test_fp32_engine_matches_oracle(pkg, lib, blobs224, case224):
    imgs, ref = case224
    with pkg.Engine(0, 224, pkg.FP32, max_batch=2) as eng:
        eng.load_weights(blobs224)
        probs, logits = eng.forward(imgs, want_logits=True)  # chunks of 2,2,1
    scale = np.abs(ref["logits"]).max()
    assert np.abs(logits - ref["logits"]).max() <= FP32_REL * scale
    assert np.abs(probs - ref["probs"]).max() <= 1e-6
    assert np.array_equal(probs.argmax(1), ref["probs"].argmax(1))


def test_fp32_stage_by_stage(pkg, lib, blobs224, case224):
    """embedding and encoder outputs of image 0 against the oracle's stage dump"""
    imgs, ref = case224
    with pkg.Engine(0, 224, pkg.FP32, max_batch=1) as eng:
        eng.load_weights(blobs224)
        eng.stage(imgs[:1])
        for layer in (0, 1, 2, 6, 12):
            eng.stop_after_layer(layer)
            eng.forward_resident(1)
            x = eng.read_tokens(1)[0]
            want = ref["stages"][layer]
            err = np.abs(x - want).max() / np.abs(want).max()
            assert err <= 5e-5, f"stage {layer}: rel err {err}"


def test_bf16_engine_matches_oracle(pkg, lib, blobs224, case224):
    imgs, ref = case224
    with pkg.Engine(0, 224, pkg.BF16, max_batch=4) as eng:
        eng.load_weights(blobs224)
        probs, logits = eng.forward(imgs, want_logits=True)
        assert eng.kernels_per_forward == 89  # 2 (embedding) + 12 * 7 + 3 (final LN, head, softmax)
    err = np.abs(logits - ref["logits"]).max()
    assert err <= BF16_ABS, f"bf16 max|dlogit| = {err}"
    assert np.array_equal(logits.argmax(1), ref["logits"].argmax(1))
    np.testing.assert_allclose(probs.sum(1), 1.0, atol=1e-5)


def test_bf16_stage_by_stage(pkg, lib, blobs224, case224):
    imgs, ref = case224
    with pkg.Engine(0, 224, pkg.BF16, max_batch=1) as eng:
        eng.load_weights(blobs224)
        eng.stage(imgs[:1])
        for layer in (0, 1, 12):
            eng.stop_after_layer(layer)
            eng.forward_resident(1)
            x = eng.read_tokens(1)[0]
            want = ref["stages"][layer]
            err = np.abs(x - want).max() / np.abs(want).max()
            assert err <= 2e-2, f"stage {layer}: rel err {err}"


def test_vit_opencl_drop_in(pkg, lib, blobs224, case224, monkeypatch):
    """the reference's calling convention (Main.c:54): ImageData[], Network[152], float** rows"""
    imgs, ref = case224
    monkeypatch.setenv("VITB200_PRECISION", "fp32")
    probs = pkg.vit_opencl(imgs[:3], blobs224)
    # comparator.c:74-86 semantics: same label, |dprob| <= 0.01 -- and far tighter here
    assert np.array_equal(probs.argmax(1), ref["probs"][:3].argmax(1))
    assert np.abs(probs - ref["probs"][:3]).max() <= 1e-6
    monkeypatch.setenv("VITB200_PRECISION", "bf16")
    monkeypatch.setenv("VITB200_BATCH", "2")
    probs16 = pkg.vit_opencl(imgs[:3], blobs224)
    assert np.array_equal(probs16.argmax(1), ref["probs"][:3].argmax(1))
    assert np.abs(probs16 - ref["probs"][:3]).max() <= 0.01


def test_vit_opencl_persistent_context(pkg, lib, blobs224, case224, monkeypatch):
    """VITB200_PERSIST=1: later calls reuse the engine and the packed weights, notice
    changed weights, and give the same rows as the default create/upload/destroy call"""
    imgs, ref = case224
    monkeypatch.setenv("VITB200_PRECISION", "fp32")
    once = pkg.vit_opencl(imgs[:3], blobs224)
    monkeypatch.setenv("VITB200_PERSIST", "1")
    try:
        a = pkg.vit_opencl(imgs[:3], blobs224)
        b = pkg.vit_opencl(imgs[:3], blobs224)           # engine + weights reused
        np.testing.assert_allclose(a, once, rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(b, once, rtol=1e-5, atol=1e-9)
        other = [w.copy() for w in blobs224]
        other[151] = other[151][::-1].copy()             # head bias reversed: same shapes, other logits
        c = pkg.vit_opencl(imgs[:3], other)
        assert np.abs(c - once).max() > 1e-6
        edited = blobs224[151].copy()
        blobs224[151][0] += 1.0                          # in-place edit of a resident blob is noticed
        try:
            d = pkg.vit_opencl(imgs[:3], blobs224)
        finally:
            blobs224[151][...] = edited
        assert np.abs(d - once).max() > 1e-6
        e = pkg.vit_opencl(imgs[:3], blobs224)
        np.testing.assert_allclose(e, once, rtol=1e-5, atol=1e-9)
        # one element in the middle of a 2.36 M-element matrix (a sampled signature would miss it)
        k = blobs224[12].size // 2 + 12345
        keep = blobs224[12][k]
        blobs224[12][k] = keep + np.float32(8.0)
        try:
            f = pkg.vit_opencl(imgs[:3], blobs224)
        finally:
            blobs224[12][k] = keep
        assert np.abs(f - once).max() > 1e-7
        g = pkg.vit_opencl(imgs[:3], blobs224)
        np.testing.assert_allclose(g, once, rtol=1e-5, atol=1e-9)
    finally:
        lib.vitb200_release_persistent()


def test_reload_weights_drops_captured_graphs(pkg, lib, blobs224):
    """a second vitb200_load_weights on an engine whose forward is already captured in CUDA graphs:
    the graphs hold the old arena's addresses, so they must be dropped with it.  The new arena is built
    beside the old one (a failed reload keeps the old weights), so its address always differs."""
    imgs = pkg.synth.synthetic_images(32, 224, seed=21)
    other = [w.copy() for w in blobs224]
    other[6] = (other[6] * np.float32(0.5)).astype(np.float32)   # layer-0 in_proj weight: changes every logit
    other[151] = other[151][::-1].copy()
    with pkg.Engine(0, 224, pkg.BF16, max_batch=32) as fresh:   # M = 6304: no split-K, deterministic sums
        fresh.load_weights(other)
        want = fresh.forward(imgs, want_logits=True)[1]
    with pkg.Engine(0, 224, pkg.BF16, max_batch=32) as eng:
        eng.load_weights(blobs224)
        first = eng.forward(imgs, want_logits=True)[1]         # eager
        eng.forward(imgs)                                       # captured
        eng.forward(imgs)                                       # replayed
        eng.load_weights(other)
        got = [eng.forward(imgs, want_logits=True)[1] for _ in range(3)]  # eager, capture, replay
        bad = list(other)
        bad[3] = bad[3][:-768]
        with pytest.raises(pkg.VitError, match="blob 3"):
            eng.load_weights(bad)                               # rejected: the engine keeps serving `other`
        still = eng.forward(imgs, want_logits=True)[1]
        eng.load_weights(blobs224)
        back = eng.forward(imgs, want_logits=True)[1]
        assert lib.vitcu_watchdog_check() == 0
    for g in got + [still]:
        assert np.array_equal(g, want)
    assert np.abs(want - first).max() > 1e-3
    assert np.array_equal(back, first)


def test_forward_structs_equals_contiguous(pkg, lib, blobs224, case224):
    imgs, _ = case224
    with pkg.Engine(0, 224, pkg.FP32, max_batch=4) as eng:
        eng.load_weights(blobs224)
        a = eng.forward(imgs)
        b = eng.forward_structs(imgs)
    # same images, same kernels; at this small batch the residual GEMMs are split along K and meet
    # through TMA reduce-add, whose fp32 summation order is not fixed -> equal to rounding, not bitwise
    np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-9)


def test_pageable_sources_match_pinned(pkg, lib, blobs224, monkeypatch):
    """pageable images (one contiguous array, or the reference's per-image buffers, R/Network.c:84-105)
    pass through the threaded pinned ring; a pinned source is DMA'd in place -- same results.
    70 images over max_batch 32 and 1 MB slots: three chunks, ragged last chunk, ragged last group."""
    monkeypatch.setenv("VITB200_STAGE_SLOT_MB", "1")
    monkeypatch.setenv("VITB200_STAGE_THREADS", "3")
    n = 70
    imgs = pkg.synth.synthetic_images(n, 224, seed=5)
    pin = pkg.PinnedArray(imgs.shape)
    pin.array[...] = imgs
    with pkg.Engine(0, 224, pkg.BF16, max_batch=32) as eng:
        eng.load_weights(blobs224)
        ref = eng.forward(pin.array)
        a = eng.forward(imgs)
        b = eng.forward_structs(imgs)
        c = eng.forward(imgs)       # ring slots are reused across calls
        monkeypatch.setenv("VITB200_NO_STAGER", "1")
        d = eng.forward(imgs)       # driver-staged pageable copy
    for got in (a, b, c, d):
        # full chunks: same kernels in the same order -> bitwise equal, so every image landed in its slot
        assert np.array_equal(got[:64], ref[:64])
        # the ragged 6-image chunk splits its residual GEMMs along K (TMA reduce-add, unordered fp32
        # sums), and one flipped bf16 rounding downstream moves a probability by up to ~1 %
        np.testing.assert_allclose(got[64:], ref[64:], rtol=3e-2, atol=1e-7)
        assert np.array_equal(got.argmax(1), ref.argmax(1))
    assert len({tuple(ref[i].argsort()[-3:]) for i in range(n)}) > 1  # not all images alike


def test_forward_topk_matches_host_argmax(pkg, lib, blobs224, case224):
    """labels-only read-back: same top-1 as scanning the full probability rows the way
    R/Main.c:59-72 does, and as the oracle's rows"""
    imgs, ref = case224
    with pkg.Engine(0, 224, pkg.FP32, max_batch=4) as eng:
        eng.load_weights(blobs224)
        probs = eng.forward(imgs)
        labels, top = eng.forward_topk(imgs, 5)
        l1, _ = eng.forward_topk(imgs, 1)
        with pytest.raises(RuntimeError):
            eng.forward_topk(imgs, 9)
    n = imgs.shape[0]
    assert np.array_equal(labels[:, 0], probs.argmax(1))
    assert np.array_equal(l1[:, 0], labels[:, 0])
    assert np.array_equal(labels[:, 0], ref["probs"][:n].argmax(1))
    np.testing.assert_allclose(top, np.take_along_axis(probs, labels, 1), rtol=1e-5)
    assert (np.diff(top, axis=1) <= 0).all()


# ---------------------------------------------------------------- model variants
@pytest.mark.parametrize("variant,img,nimg", [("b32", 224, 3), ("s16", 224, 3), ("b32", 384, 2), ("l16", 224, 1)])
def test_model_variants_match_oracle(pkg, lib, variant, img, nimg):
    """other members of the family through run-time dims: ViT-B/32 (patch_size 32: gather + GEMM patch
    embedding, 50 tokens), ViT-S/16 (embed_dim 384, 6 heads: 1-CTA GEMM tiles, 3-vector LayerNorm),
    ViT-L/16 (1024 wide, 16 heads, 24 layers, 296 blobs).  Oracle = the -D build of the restatement
    (b32 / s16 pinned against the reference compiled with the same macro edit, tests/test_oracle.py)."""
    from oracle import binding
    blobs = pkg.synth.variant_blobs(variant, img, seed=7)
    imgs = pkg.synth.synthetic_images(nimg, img, seed=1234)
    ref = binding.Oracle(variant).forward(imgs, blobs)
    scale = np.abs(ref["logits"]).max()
    with pkg.Engine(0, img, pkg.FP32, max_batch=2, model=variant) as eng:
        eng.load_weights(blobs)
        probs, logits = eng.forward(imgs, want_logits=True)
    assert np.abs(logits - ref["logits"]).max() <= FP32_REL * scale
    assert np.array_equal(probs.argmax(1), ref["probs"].argmax(1))
    with pkg.Engine(0, img, pkg.BF16, max_batch=4, model=variant) as eng:
        eng.load_weights(blobs)
        probs16, logits16 = eng.forward(imgs, want_logits=True)
        assert lib.vitcu_watchdog_check() == 0
    assert np.abs(logits16 - ref["logits"]).max() <= BF16_ABS
    assert np.array_equal(probs16.argmax(1), ref["probs"].argmax(1))


def test_vit_opencl_infers_variant_from_blobs(pkg, lib, monkeypatch):
    """the unchanged drop-in call with ViT-B/32 weights: patch side, width and MLP width are read off the
    blob sizes (vitb200_model_from_blobs), the image side off the ImageData struct"""
    from oracle import binding
    blobs = pkg.synth.variant_blobs("b32", 224, seed=7)
    imgs = pkg.synth.synthetic_images(2, 224, seed=1234)
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "variant_vectors.npz"))
    monkeypatch.setenv("VITB200_PRECISION", "fp32")
    probs = pkg.vit_opencl(imgs, blobs)
    assert np.abs(probs - gold["b32_probs"]).max() <= 1e-6      # the reference's own output (patch_size 32 build)
    assert np.array_equal(probs.argmax(1), gold["b32_probs"].argmax(1))
    m = pkg.Model()
    nets, keep = pkg.make_network_structs(blobs)
    assert lib.vitb200_model_from_blobs(nets, None, C.byref(m)) == 0
    assert (m.img, m.patch, m.embed, m.depth, m.heads, m.hidden) == (224, 32, 768, 12, 12, 3072)


def test_model_rejects_unsupported_dims(pkg, lib):
    for bad in (dict(patch=8), dict(embed=512, heads=8), dict(heads=8), dict(depth=0), dict(hidden=100), dict(img=200)):
        m = pkg.Model.variant("b16", 224)
        for k, v in bad.items():
            setattr(m, k, v)
        h = C.c_void_p()
        assert lib.vitb200_create_model(C.byref(h), 0, C.byref(m), pkg.FP32, 1) != 0, bad


def test_golden_reference_vectors(pkg, lib, synth_blobs224):
    """probabilities the reference's own ViT_seq.c produced (tests/golden/make_golden.py)"""
    imgs = pkg.synth.synthetic_images(2, 224, seed=1234)
    with pkg.Engine(0, 224, pkg.FP32, max_batch=2) as eng:
        eng.load_weights(synth_blobs224)
        probs = eng.forward(imgs)
    want = GOLD["synth_probs"]
    assert np.abs(probs - want).max() <= 1e-4 * want.max()
    assert np.array_equal(probs.argmax(1), want.argmax(1))


def test_golden_bundled_image(pkg, lib, ref_dir):
    net, img_path = os.path.join(ref_dir, "Network"), os.path.join(ref_dir, "Data", "input-1.bin")
    if not (os.path.isdir(net) and os.path.exists(img_path)):
        pytest.skip("bundled reference data (oracle/_ref) not present on this box")
    blobs = pkg.synth.model_blobs(net, 224, seed=0)
    img = pkg.synth.load_image_file(img_path)
    want = GOLD["bundled_input1_probs"]
    for precision, tol in ((pkg.FP32, 1e-4 * want.max()), (pkg.BF16, 0.01)):
        with pkg.Engine(0, 224, precision, max_batch=1) as eng:
            eng.load_weights(blobs)
            probs = eng.forward(img)
        assert int(probs.argmax()) == 606
        assert np.abs(probs - want).max() <= tol


def test_384_matches_reference(pkg, lib):
    """577 tokens: beyond what the reference's OpenCL attention could run (256-key cap)"""
    blobs = pkg.synth.model_blobs(None, 384, seed=7)
    imgs = pkg.synth.synthetic_images(1, 384, seed=4321)
    want = GOLD["synth384_probs"]
    with pkg.Engine(0, 384, pkg.FP32, max_batch=1) as eng:
        eng.load_weights(blobs)
        probs = eng.forward(imgs)
    assert np.abs(probs - want).max() <= 1e-4 * want.max()
    assert np.array_equal(probs.argmax(1), want.argmax(1))
    with pkg.Engine(0, 384, pkg.BF16, max_batch=1) as eng:
        eng.load_weights(blobs)
        probs16 = eng.forward(imgs)
    assert np.abs(probs16 - want).max() <= 0.01


def test_edge_cases(pkg, lib, blobs224):
    with pkg.Engine(0, 224, pkg.FP32, max_batch=2) as eng:
        with pytest.raises(pkg.VitError):  # weights not loaded
            eng.forward(np.zeros((1, 3, 224, 224), np.float32))
        holed = list(blobs224)
        holed[6] = None
        with pytest.raises(pkg.VitError, match="blob 6 is missing"):
            eng.load_weights(holed)
        bad = list(blobs224)
        bad[3] = bad[3][:-768]
        with pytest.raises(pkg.VitError, match="blob 3"):
            eng.load_weights(bad)
        eng.load_weights(blobs224)
        assert eng.forward(np.zeros((0, 3, 224, 224), np.float32)).shape == (0, 1000)


def test_full_batch_properties_bf16(pkg, lib, blobs224):
    """BASELINE config 3 size (batch 256): size-independent properties instead of the 50-minute oracle run"""
    n = 256
    imgs = pkg.synth.synthetic_images(n, 224, seed=77)
    with pkg.Engine(0, 224, pkg.BF16, max_batch=n) as eng:
        eng.load_weights(blobs224)
        p1 = eng.forward(imgs)
        p2 = eng.forward(imgs)          # second call replays the captured CUDA graph
        perm = np.random.default_rng(0).permutation(n)
        p3 = eng.forward(np.ascontiguousarray(imgs[perm]))
        assert lib.vitcu_watchdog_check() == 0
    assert np.isfinite(p1).all()
    np.testing.assert_allclose(p1.sum(1), 1.0, atol=1e-5)
    assert np.array_equal(p1, p2)                       # deterministic, graph == eager
    assert np.array_equal(p3, p1[perm])                 # images are independent of their batch position
    # the same images through a small-batch engine: same math, other tiling
    with pkg.Engine(0, 224, pkg.BF16, max_batch=8) as eng:
        eng.load_weights(blobs224)
        p4 = eng.forward(imgs[:16])
    # (the 256-image chunk folds its LayerNorms into the GEMMs, the 8-image chunk runs the LayerNorm kernel: the
    # bf16 rounding points differ, the probabilities agree to a few 1e-4)
    assert np.abs(p4 - p1[:16]).max() <= 3e-4
    assert np.array_equal(p4.argmax(1), p1[:16].argmax(1))


def test_4096_images_replica_property_bf16(pkg, lib, blobs224):
    """BASELINE config 4 size on one GPU (4096 images, 16 chunks of 256) through a size-independent
    property: 64 distinct images tiled 64 times -- every replica must give bit-identical rows (same
    position inside a full chunk -> same kernels, same tiles), the 64 distinct rows must match a
    64-image run to rounding, and nothing may be lost between chunks, staging slots or graph replays"""
    base = pkg.synth.synthetic_images(64, 224, seed=99)
    imgs = np.ascontiguousarray(np.tile(base, (64, 1, 1, 1)))
    with pkg.Engine(0, 224, pkg.BF16, max_batch=256) as eng:
        eng.load_weights(blobs224)
        probs = eng.forward(imgs)               # pageable source: threaded pinned staging
        labels, top = eng.forward_topk(imgs, 1)
        small = eng.forward(base)
        assert lib.vitcu_watchdog_check() == 0
    assert probs.shape == (4096, 1000) and np.isfinite(probs).all()
    np.testing.assert_allclose(probs.sum(1), 1.0, atol=1e-5)
    tiles = probs.reshape(64, 64, 1000)
    assert np.array_equal(tiles, np.broadcast_to(tiles[0], tiles.shape))
    assert np.array_equal(labels[:, 0], probs.argmax(1))
    assert np.array_equal(small.argmax(1), tiles[0].argmax(1))
    assert np.abs(small - tiles[0]).max() <= 3e-4


def test_main_c_drop_in(pkg, lib, oracle, blobs224, ref_dir, tmp_path):
    """The reference's UNMODIFIED Main.c + Network.c + comparator.c (objects built by
    oracle/Makefile) linked against libvit_b200.so instead of ViT_opencl.c: the
    reference's own end-to-end check must print "good"."""
    objs = [os.path.join(ref_dir, f) for f in ("Main.o", "comparator.o", "Network.o")]
    if not all(os.path.exists(o) for o in objs):
        pytest.skip("oracle/_ref objects not present on this box")
    exe = tmp_path / "main_b200"
    subprocess.run(["gcc", "-o", str(exe), *objs, "-L" + os.path.dirname(pkg.LIB_PATH), "-lvit_b200",
                    "-Wl,-rpath," + os.path.dirname(pkg.LIB_PATH), "-lm"], check=True)
    (tmp_path / "Data").mkdir()
    (tmp_path / "Network").mkdir()
    for i, b in enumerate(blobs224):
        b.tofile(tmp_path / "Network" / f"Weight_{i}_blob.bin")
    # comparator.c insists on 100 lines (IMAGE_COUNT): 4 distinct images, repeated
    base = pkg.synth.synthetic_images(4, 224, seed=31)
    imgs = np.ascontiguousarray(base[np.arange(100) % 4])
    pkg.synth.write_image_file(str(tmp_path / "Data" / "input-100.bin"), imgs)
    ref = oracle.forward(base, blobs224)["probs"]
    # answer file in Main.c's format; Main.c:59-69 carries pred_idx over between images
    pred = 0
    with open(tmp_path / "Data" / "answer_result.txt", "w") as f:
        for i in range(100):
            row = ref[i % 4]
            for j in range(1, 1000):
                if row[j] > row[pred]:
                    pred = j
            f.write("[%d] label: %d / prob: %.6f\n" % (i, pred, row[pred]))
    env = dict(os.environ, VITB200_PRECISION="fp32")
    out = subprocess.run([str(exe)], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "good" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_vit_opencl_multi_gpu_split(pkg, lib, blobs224, monkeypatch):
    """the product's own multi-GPU path (vit_opencl.c: one host thread per GPU, contiguous shards,
    replicated weights, host-side gather into the caller's rows; reference loop R/ViT_opencl.c:926-965):
    one ViT_opencl call over G GPUs must return the rows a 1-GPU call returns -- bit for bit when the
    shards are whole chunks (same position in a full chunk -> same kernels, same tiles), to rounding and
    with the same labels for a ragged split -- cold and with the persistent context."""
    ndev = pkg.device_count()
    if ndev < 2:
        pytest.skip("needs at least 2 GPUs (run with gpurun --gpus 2)")
    gpus = min(ndev, 4)
    monkeypatch.setenv("VITB200_PRECISION", "bf16")
    monkeypatch.setenv("VITB200_BATCH", "64")
    base = pkg.synth.synthetic_images(96, 224, seed=11)
    n = 64 * 2 * gpus
    imgs = np.ascontiguousarray(base[np.arange(n) % 96])
    monkeypatch.setenv("VITB200_GPUS", "1")
    one = pkg.vit_opencl(imgs, blobs224)
    stats = pkg.CallStats()
    lib.vitb200_last_call_stats(C.byref(stats))
    assert (stats.images, stats.gpus) == (n, 1)
    monkeypatch.setenv("VITB200_GPUS", str(gpus))
    cold = pkg.vit_opencl(imgs, blobs224)
    lib.vitb200_last_call_stats(C.byref(stats))
    assert (stats.images, stats.gpus) == (n, gpus)
    assert np.array_equal(cold, one)
    monkeypatch.setenv("VITB200_PERSIST", "1")
    try:
        warm1 = pkg.vit_opencl(imgs, blobs224)
        warm2 = pkg.vit_opencl(imgs, blobs224)
        ragged = pkg.vit_opencl(imgs[:n - 37], blobs224)
    finally:
        lib.vitb200_release_persistent()
    assert np.array_equal(warm1, one) and np.array_equal(warm2, one)
    assert np.array_equal(ragged.argmax(1), one[:n - 37].argmax(1))
    np.testing.assert_allclose(ragged, one[:n - 37], rtol=3e-2, atol=1e-7)
    assert lib.vitcu_watchdog_check() == 0


@pytest.mark.parametrize("variant,img,batch", [("l16", 224, 64), ("b32", 224, 256)])
def test_model_variants_large_chunks_match_small_chunks(pkg, lib, variant, img, batch):
    """the other model widths at chunk sizes that take the CTA-pair GEMMs, the folded LayerNorm (6 or 8 partial-sum slots)
    and the duo attention kernel, against the same images through 4-image chunks (separate LayerNorm kernel, single-CTA
    GEMM tiles), which test_model_variants_match_oracle ties to the oracle; FP8 on the same chunk within its own contract"""
    blobs = pkg.synth.variant_blobs(variant, img, seed=7)
    imgs = pkg.synth.synthetic_images(batch, img, seed=99)
    with pkg.Engine(0, img, pkg.BF16, max_batch=4, model=variant) as eng:
        eng.load_weights(blobs)
        small, small_logits = eng.forward(imgs[:16], want_logits=True)
    with pkg.Engine(0, img, pkg.BF16, max_batch=batch, model=variant) as eng:
        eng.load_weights(blobs)
        lib.vitcu_launch_count_reset()
        big, big_logits = eng.forward(imgs, want_logits=True)
        counts = pkg.launch_counts()
        again = eng.forward(imgs)
        assert lib.vitcu_watchdog_check() == 0
    depth = pkg.synth.VARIANTS[variant][2]
    # (+ 1 for /32 patches: their patch embedding is a gather + the same GEMM)
    assert counts["gemm_bf16_tc2_kernel"] == 4 * depth + (1 if variant == "b32" else 0), counts
    assert counts["layernorm_kernel"] == 1, counts              # only the final one: the others are folded
    assert np.array_equal(big, again)
    # two BF16 paths with different rounding points (each within 2e-2 of the oracle): up to twice that apart
    # (no top-1 comparison between the two: with random-init weights some images have top-1 margins below the rounding)
    assert np.abs(big_logits[:16] - small_logits).max() <= 4e-2
    from oracle import binding
    ref = binding.Oracle(variant).forward(imgs[:3], blobs)
    err = np.abs(big_logits[:3] - ref["logits"]).max()
    print(f"\n{variant} batch {batch}: folded BF16 path vs oracle max|dlogit| = {err:.3e}")
    assert err <= BF16_ABS
    srt = np.sort(ref["logits"], 1)
    clear = (srt[:, -1] - srt[:, -2]) > 2 * BF16_ABS             # top-1 must agree wherever the oracle's margin allows it
    assert np.array_equal(big[:3].argmax(1)[clear], ref["probs"].argmax(1)[clear])
    with pkg.Engine(0, img, pkg.FP8, max_batch=batch, model=variant) as eng:
        eng.load_weights(blobs)
        p8, l8 = eng.forward(imgs, want_logits=True)
        assert lib.vitcu_watchdog_check() == 0
    assert np.isfinite(l8).all()
    assert np.abs(l8[:3] - ref["logits"]).max() <= 2.5e-1
