"""Oracle parity AT THE CONFIGURATION THE BENCH TIMES (BASELINE.json configs 2, 3 and 5).

The small-batch parity tests (tests/test_gpu_forward.py) dispatch to the single-CTA GEMM tiles; the
bench (256 images per step, M = 50 432 rows) runs the CTA-pair 256x256 kernel, CUDA-graph replay and a
3 072-item persistent attention launch.  These tests compare exactly that configuration with the CPU
oracle (oracle/vit_oracle.c, pinned bit-exact to the reference's ViT_seq.c) on the first 32 images of
the bench's own input (images N(0,1) seed 1234, bundled + seed-0 weights):

  BF16  max_batch 256  : max|logit - ref| <= 2e-2 on each of the 32 images, identical top-1 on all 32,
                         eager forward == graph replay bit for bit, launch list holds the CTA-pair GEMM
  FP32  max_batch 64   : max|logit - ref| <= 1e-4 * max|ref logit| on all 32 images
  384x384, max_batch 16: the key-blocked (flash) attention with more than one item per CTA, 4 images
                         against the img_size-384 oracle, same BF16 contract

Reference semantics: R/ViT_seq.c:402-517; the acceptance rule of the reference's own check is
R/comparator.c:74-86 (same label, |dprob| <= 0.01), which is asserted as well.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FP32_REL = 1e-4
BF16_ABS = 2e-2
N_CHECK = 32


def _histogram(err, edges=(0.0, 2.5e-3, 5e-3, 7.5e-3, 1e-2, 1.25e-2, 1.5e-2, 1.75e-2, 2e-2, 1.0)):
    h, _ = np.histogram(err, bins=np.asarray(edges))
    return " ".join(f"[{edges[i]:.4g},{edges[i + 1]:.4g}):{h[i]}" for i in range(len(h)))


@pytest.fixture(scope="module")
def bench_case(pkg, oracle, blobs224):
    """the bench's 256 images (rank 0) and the oracle's logits / probabilities of the first 32"""
    imgs = pkg.synth.synthetic_images(256, 224, seed=1234)
    ref = oracle.forward(imgs[:N_CHECK], blobs224)
    return imgs, ref


def test_bf16_batch256_pair_gemm_graph_vs_oracle(pkg, lib, blobs224, bench_case):
    imgs, ref = bench_case
    with pkg.Engine(0, 224, pkg.BF16, max_batch=256) as eng:
        eng.load_weights(blobs224)
        # exactly what bench.py times: the 256-image chunk resident in HBM, one forward per step
        eng.stage(imgs)
        lib.vitcu_launch_count_reset()
        eng.forward_resident(256)                                   # first full chunk: eager launches
        counts = pkg.launch_counts()
        p_eager, l_eager = eng.read_probs(256)
        eng.forward_resident(256)                                   # second: captured into a CUDA graph, replayed
        p_graph, l_graph = eng.read_probs(256)
        eng.forward_resident(256)                                   # third: replay of the same graph
        p_again, l_again = eng.read_probs(256)
        # the public call over host images (upload, chunk pipeline, read-back) must give the same bits
        p_call, l_call = eng.forward(imgs, want_logits=True)
        assert lib.vitcu_watchdog_check() == 0
    # the kernels the bench times are the ones checked here
    assert counts["gemm_bf16_tc2_kernel"] == 48, counts           # qkv, out_proj, fc1, fc2 x 12 as CTA pairs
    assert counts["gemm_bf16_tc_kernel"] == 0, counts
    assert counts["attention_duo_tc_kernel"] + counts["attention_tc_kernel"] == 12, counts
    assert counts["patch_embed_tc_kernel"] == 1, counts
    assert counts["layernorm_kernel"] == 1, counts                # the final one; the 24 others are folded into the GEMMs
    assert np.array_equal(l_eager, l_graph) and np.array_equal(l_graph, l_again)
    assert np.array_equal(p_eager, p_graph)
    assert np.array_equal(l_call, l_graph) and np.array_equal(p_call, p_graph)
    err = np.abs(l_graph[:N_CHECK] - ref["logits"]).max(1)         # per image
    margin = np.sort(ref["logits"], 1)
    margin = margin[:, -1] - margin[:, -2]
    print(f"\nBF16 batch-256 logit error per image over {N_CHECK} images: max {err.max():.4e} mean {err.mean():.4e} "
          f"rms(all logits) {np.sqrt(np.mean((l_graph[:N_CHECK] - ref['logits']) ** 2)):.4e}")
    print("histogram of per-image max|dlogit|:", _histogram(err))
    print(f"oracle top-1 margins: min {margin.min():.3f} max {margin.max():.3f}")
    assert err.max() <= BF16_ABS, f"worst image {int(err.argmax())}: max|dlogit| = {err.max()}"
    assert np.array_equal(l_graph[:N_CHECK].argmax(1), ref["logits"].argmax(1))
    # the reference's own acceptance rule (R/comparator.c:74-86)
    assert np.array_equal(p_graph[:N_CHECK].argmax(1), ref["probs"].argmax(1))
    assert np.abs(p_graph[:N_CHECK].max(1) - ref["probs"].max(1)).max() <= 0.01


def test_fp32_batch64_vs_oracle(pkg, lib, blobs224, bench_case):
    imgs, ref = bench_case
    with pkg.Engine(0, 224, pkg.FP32, max_batch=64) as eng:
        eng.load_weights(blobs224)
        lib.vitcu_launch_count_reset()
        p1, l1 = eng.forward(imgs[:128], want_logits=True)          # chunk 1 eager, chunk 2 captured
        counts = pkg.launch_counts()
        p2, l2 = eng.forward(imgs[:64], want_logits=True)           # graph replay
        assert lib.vitcu_watchdog_check() == 0
    assert counts["gemm_bf16_tc2_kernel"] + counts["gemm_bf16_tc_kernel"] >= 48, counts  # split-bf16 on the tensor cores
    assert np.array_equal(l1[:64], l2)
    scale = np.abs(ref["logits"]).max()
    err = np.abs(l1[:N_CHECK] - ref["logits"]).max(1)
    print(f"\nFP32 batch-64 logit error over {N_CHECK} images: max {err.max():.3e} (bound {FP32_REL * scale:.3e})")
    assert err.max() <= FP32_REL * scale
    assert np.array_equal(l1[:N_CHECK].argmax(1), ref["logits"].argmax(1))
    assert np.abs(p1[:N_CHECK] - ref["probs"]).max() <= 1e-6


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_batch1_latency_chain_vs_oracle(pkg, lib, blobs224, bench_case, precision, monkeypatch):
    """BASELINE config 2 at the configuration `latency_batch1` of the bench line times: Engine(max_batch=1), one image
    resident, eager forward then graph replays.  FP32: the K-sliced chain -- LayerNorm launches that zero the GEMM
    outputs, qkv / fc1 / patch embedding in accumulate mode with all split-bf16 pieces of a k-block in one ring slot,
    GELU in the split pass, the cp.async attention kernel -- within 1e-4 of the oracle's logits on each of 8 images, and
    again with VITB200_FP32_SPLITK=0 / VITCU_FP32_FUSED=0 (the chain it replaced).  BF16: 2e-2 and identical top-1.
    Replays are deterministic only up to the order in which K slices reduce-add, so replays are compared by tolerance
    (chunks of 32 images and more do not slice K and are bit-reproducible: test_bf16_batch256_pair_gemm_graph_vs_oracle)."""
    imgs, ref = bench_case
    n = 8
    fp32 = precision == "fp32"
    results = []
    for old_chain in ((False, True) if fp32 else (False,)):
        if old_chain:
            monkeypatch.setenv("VITB200_FP32_SPLITK", "0")
            monkeypatch.setenv("VITCU_FP32_FUSED", "0")
        with pkg.Engine(0, 224, pkg.FP32 if fp32 else pkg.BF16, max_batch=1) as eng:
            eng.load_weights(blobs224)
            logits = np.empty((n, 1000), np.float32)
            counts = None
            for i in range(n):
                eng.stage(imgs[i:i + 1])
                lib.vitcu_launch_count_reset()
                eng.forward_resident(1)                   # eager on the first image, graph replays afterwards
                if counts is None:
                    counts = pkg.launch_counts()          # (a launch captured into a graph counts once, a replay not at all)
                _, l0 = eng.read_probs(1)
                eng.forward_resident(1)
                _, l1 = eng.read_probs(1)
                # K slices add in arrival order: fp32 sums differ in the last bits from run to run, and on the BF16 path
                # those bits decide bf16 roundings downstream -- every run is another draw of the same rounding noise
                # that separates it from the oracle (observed between two runs: up to 1.04e-2 on a logit)
                assert np.abs(l1 - l0).max() <= (1e-5 * np.abs(l0).max() if fp32 else BF16_ABS)
                logits[i] = l1[0]
            launches = eng.kernels_per_forward
            assert lib.vitcu_watchdog_check() == 0
        if fp32:
            assert launches == 103 and counts["gemm_bf16_tc_kernel"] == 49 and counts["gemm_bf16_tc2_kernel"] == 0, (launches, counts)
            assert counts["attention_simt_kernel"] == 12 and counts["layernorm_kernel"] == 25, counts
        else:
            assert launches == 89 and counts["gemm_bf16_tc_kernel"] == 48 and counts["attention_duo_tc_kernel"] == 12, (launches, counts)
        results.append(logits)
        err = np.abs(logits - ref["logits"][:n]).max(1)
        scale = np.abs(ref["logits"][:n]).max(1)
        print(f"\n{precision} batch-1 chain{' (unsliced, unfused)' if old_chain else ''}: max|dlogit| per image {err.max():.3e} "
              f"(logit scale {scale.min():.2f})")
        if fp32:
            assert (err <= FP32_REL * scale).all(), err / scale
        else:
            assert err.max() <= BF16_ABS
        assert np.array_equal(logits.argmax(1), ref["logits"][:n].argmax(1))
    if fp32:  # both chains are fp32-accurate: they agree far inside the oracle tolerance
        assert np.abs(results[0] - results[1]).max() <= 2e-5 * np.abs(results[1]).max()


def test_bf16_384_flash_batch16_vs_oracle(pkg, lib, oracle):
    """BASELINE config 5 shape: 577 tokens, key-blocked attention, 16 x 12 = 192 items over 148 CTAs"""
    blobs = pkg.synth.model_blobs(None, 384, seed=7)
    imgs = pkg.synth.synthetic_images(16, 384, seed=1234)
    ref = oracle.forward(imgs[:4], blobs)
    with pkg.Engine(0, 384, pkg.BF16, max_batch=16) as eng:
        eng.load_weights(blobs)
        lib.vitcu_launch_count_reset()
        p1, l1 = eng.forward(imgs, want_logits=True)
        counts = pkg.launch_counts()
        p2, l2 = eng.forward(imgs, want_logits=True)                # graph replay
        assert lib.vitcu_watchdog_check() == 0
    assert counts["attention_flash_tc_kernel"] == 12, counts
    assert np.array_equal(l1, l2)
    err = np.abs(l1[:4] - ref["logits"]).max(1)
    print(f"\nBF16 384x384 batch-16 logit error over 4 images: {err}")
    assert err.max() <= BF16_ABS
    assert np.array_equal(l1[:4].argmax(1), ref["logits"].argmax(1))
    assert np.abs(p1[:4].max(1) - ref["probs"].max(1)).max() <= 0.01


FP8_ABS = 2.5e-1


def test_fp8_batch256_vs_oracle(pkg, lib, blobs224, bench_case):
    """VITB200_FP8: the BF16 path with fc1 / fc2 on E4M3 operands (tcgen05 kind::f8f6f4, per-tensor scales,
    activations calibrated on the first chunk).  Accuracy contract ("FP8"): on the bench's first 32
    images max|logit - ref| <= 2.5e-1 (measured 0.15 - 0.20 per image) and identical top-1 on every image (the oracle's smallest top-1 margin on this
    set is 0.53); the reference's own acceptance rule (R/comparator.c:74-86: same label, |dprob| <= 0.01) holds."""
    imgs, ref = bench_case
    with pkg.Engine(0, 224, pkg.FP8, max_batch=256) as eng:
        eng.load_weights(blobs224)
        eng.stage(imgs)
        lib.vitcu_launch_count_reset()
        eng.forward_resident(256)                         # calibration pass (bf16 chain) + eager FP8 forward
        counts = pkg.launch_counts()
        p1, l1 = eng.read_probs(256)
        eng.forward_resident(256)                         # captured graph
        p2, l2 = eng.read_probs(256)
        eng.forward_resident(256)                         # replay
        p3, l3 = eng.read_probs(256)
        assert lib.vitcu_watchdog_check() == 0
    assert counts["gemm_bf16_tc2_kernel"] == 2 * 48, counts       # calibration (bf16) + fp8 forward, all CTA pairs
    assert np.array_equal(l1, l2) and np.array_equal(l2, l3)
    err = np.abs(l2[:N_CHECK] - ref["logits"]).max(1)
    print(f"\nFP8 batch-256 logit error per image over {N_CHECK} images: max {err.max():.4e} mean {err.mean():.4e} "
          f"rms(all logits) {np.sqrt(np.mean((l2[:N_CHECK] - ref['logits']) ** 2)):.4e}")
    print("histogram of per-image max|dlogit|:", _histogram(err, (0.0, 2e-2, 4e-2, 6e-2, 8e-2, 1e-1, 1.5e-1, 2e-1, 1.0)))
    assert err.max() <= FP8_ABS
    assert np.array_equal(l2[:N_CHECK].argmax(1), ref["logits"].argmax(1))
    assert np.abs(p2[:N_CHECK].max(1) - ref["probs"].max(1)).max() <= 0.01
    np.testing.assert_allclose(p2.sum(1), 1.0, atol=1e-5)
