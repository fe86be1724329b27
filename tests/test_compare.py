"""host-side result comparison (vit-with-opencl_b200/compare.py): the reference comparator's rules
(R/comparator.c:27-96) for any image count, and the stricter row-level check"""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def cmp(pkg):
    return pkg.compare


def _rows(n=6, seed=0):
    rng = np.random.default_rng(seed)
    l = rng.standard_normal((n, 1000)).astype(np.float32) * 3
    e = np.exp(l - l.max(1, keepdims=True))
    return (e / e.sum(1, keepdims=True)).astype(np.float32), l


def test_reference_answer_file_parses(cmp, ref_dir):
    """the reference's own golden line for its bundled image (R/Data/answer_result_1.txt)"""
    import os
    path = os.path.join(ref_dir, "Data", "answer_result_1.txt")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/Data not built")
    labels, probs = cmp.parse_result_text(open(path).read())
    assert labels.tolist() == [65] and abs(float(probs[0]) - 0.919345) < 1e-6


def test_result_file_round_trip_and_rules(cmp):
    p, _ = _rows()
    text = "\n".join(cmp.format_result_lines(p)) + "\n"
    labels, probs = cmp.parse_result_text(text)
    assert np.array_equal(labels, p.argmax(1))
    assert np.abs(probs - p.max(1)).max() < 1e-6
    assert cmp.compare_result_files(text, text).ok
    # one wrong label + one probability off by 0.02 -> two errors, as comparator.c counts them
    lines = text.splitlines()
    lines[1] = f"[1] label: {(labels[1] + 1) % 1000} / prob: {probs[1]:.6f}"
    lines[4] = f"[4] label: {labels[4]} / prob: {probs[4] + 0.02:.6f}"
    rep = cmp.compare_result_files("\n".join(lines), text)
    assert rep.errors == 2 and "Label mismatch" in rep.messages[0] and "Probability mismatch" in rep.messages[1]
    # 0.005 passes the reference tolerance, fails a tightened one
    lines = text.splitlines()
    lines[2] = f"[2] label: {labels[2]} / prob: {probs[2] + 0.005:.6f}"
    assert cmp.compare_result_files("\n".join(lines), text).ok
    assert cmp.compare_result_files("\n".join(lines), text, prob_tol=1e-4).errors == 1
    # short file: reported, like the "not enough lines" branch
    assert cmp.compare_result_files("\n".join(text.splitlines()[:3]), text).errors == 1
    with pytest.raises(ValueError):
        cmp.parse_result_text("label 3 prob 0.5\n")


def test_compare_rows_strict(cmp):
    p, l = _rows()
    assert cmp.compare_rows(p, p, l, l).ok
    q = p.copy()
    q[2, 17] += 1e-4                       # not the winner: invisible to comparator.c, caught here
    rep = cmp.compare_rows(q, p)
    assert rep.errors == 1 and "dprob" in rep.messages[0]
    assert cmp.compare_result_files("\n".join(cmp.format_result_lines(q)), "\n".join(cmp.format_result_lines(p))).ok
    l2 = l.copy()
    l2[3] += 2e-3 * np.abs(l).max()
    assert cmp.compare_rows(p, p, l2, l).errors == 1          # FP32 contract: 1e-4 relative
    assert cmp.compare_rows(p, p, l2, l, logit_abs=2e-2 * np.abs(l).max()).ok
    sw = p.copy()
    a, b = np.argsort(-p[0])[:2]
    sw[0, a], sw[0, b] = p[0, b], p[0, a]  # top-1 swapped
    assert any("top-1" in m for m in cmp.compare_rows(sw, p, prob_abs=1.0).messages)
    assert not cmp.compare_rows(p[:2], p[:3]).ok
