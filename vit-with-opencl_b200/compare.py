"""Result comparison for the ViT path: the reference's comparator and a stricter one.

The reference checks an OpenCL run by re-reading the two text files Main.c wrote and comparing, per
image, the top-1 label and its probability to 0.01 (R/comparator.c:27-96, R/ = /root/reference/
MulticoreMainProject/); it insists on exactly 100 lines (IMAGE_COUNT, R/comparator.c:9).  This module
keeps those semantics for any image count (`compare_result_files`) and adds what that check cannot
see (`compare_rows`): whole probability rows, pre-softmax logits at a relative tolerance, and top-k
agreement.  Host-side test tooling only -- nothing on the forward path imports it.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

import numpy as np

_LINE = re.compile(r"^\[(-?\d+)\] label: (-?\d+) / prob: ([-+0-9.eEnaifNAIF]+)")


def format_result_lines(probs: np.ndarray) -> list[str]:
    """the lines Main.c writes (R/Main.c:59-72): first maximum wins, probability with %.6f"""
    return [f"[{i}] label: {int(np.argmax(r))} / prob: {float(r[int(np.argmax(r))]):.6f}" for i, r in enumerate(probs)]


def parse_result_text(text: str):
    """-> (labels int64 [n], probs float32 [n]); raises ValueError on a line comparator.c could not parse"""
    labels, probs = [], []
    for ln, line in enumerate(text.splitlines()):
        if not line.strip():
            continue
        m = _LINE.match(line)
        if not m:
            raise ValueError(f"line {ln}: not '[i] label: L / prob: P': {line!r}")
        labels.append(int(m.group(2)))
        probs.append(np.float32(m.group(3)))
    return np.array(labels, np.int64), np.array(probs, np.float32)


@dataclass
class Report:
    errors: int = 0
    messages: list = field(default_factory=list)
    stats: dict = field(default_factory=dict)

    def fail(self, msg: str):
        self.errors += 1
        self.messages.append(msg)

    @property
    def ok(self) -> bool:
        return self.errors == 0


def compare_result_files(result_text: str, answer_text: str, prob_tol: float = 0.01, count: int | None = None) -> Report:
    """comparator.c semantics (label equal, |dprob| <= 0.01, one error per violated rule) over `count`
    lines (default: all lines of the answer file instead of the fixed 100)"""
    rep = Report()
    lr, pr = parse_result_text(result_text)
    la, pa = parse_result_text(answer_text)
    n = len(la) if count is None else count
    if len(lr) < n or len(la) < n:
        rep.fail(f"Line {min(len(lr), len(la))}: not enough lines (result {len(lr)}, answer {len(la)}, need {n})")
        n = min(len(lr), len(la))
    for i in range(n):
        if lr[i] != la[i]:
            rep.fail(f"Line {i}: Label mismatch (Result: {lr[i]}, Answer: {la[i]})")
        if not abs(float(pr[i]) - float(pa[i])) <= prob_tol:
            rep.fail(f"Line {i}: Probability mismatch (Result: {pr[i]:.6f}, Answer: {pa[i]:.6f})")
    rep.stats = {"lines": n, "max_abs_dprob": float(np.abs(pr[:n] - pa[:n]).max()) if n else 0.0}
    return rep


def compare_rows(probs: np.ndarray, ref_probs: np.ndarray, logits: np.ndarray | None = None,
                 ref_logits: np.ndarray | None = None, prob_abs: float = 1e-6, logit_rel: float = 1e-4,
                 logit_abs: float | None = None, topk: int = 5) -> Report:
    """Stricter check of full rows [n,1000]:
    * identical top-1 on every image (ties: first maximum, as Main.c scans);
    * max |dprob| <= prob_abs over every class, not just the winner;
    * logits (when given): max|dlogit| <= logit_rel * max|ref logit| (the FP32 contract of BASELINE.json)
      or <= logit_abs when that is given instead (the BF16 contract, 2e-2);
    * the top-k sets are reported (stats), and a differing top-k ORDER counts only when the reference
      separates the two classes by more than the probability tolerance."""
    rep = Report()
    probs, ref_probs = np.asarray(probs), np.asarray(ref_probs)
    if probs.shape != ref_probs.shape:
        rep.fail(f"shape mismatch {probs.shape} vs {ref_probs.shape}")
        return rep
    n = probs.shape[0]
    if not np.isfinite(probs).all():
        rep.fail("non-finite probabilities")
    t1, r1 = probs.argmax(1), ref_probs.argmax(1)
    for i in np.nonzero(t1 != r1)[0]:
        rep.fail(f"image {i}: top-1 {t1[i]} != reference {r1[i]}")
    dp = np.abs(probs - ref_probs).max(1) if n else np.zeros(0)
    for i in np.nonzero(~(dp <= prob_abs))[0]:
        rep.fail(f"image {i}: max |dprob| {dp[i]:.3e} > {prob_abs:.1e}")
    rep.stats = {"images": n, "max_abs_dprob": float(dp.max()) if n else 0.0}
    if logits is not None and ref_logits is not None:
        dl = np.abs(np.asarray(logits) - np.asarray(ref_logits)).max(1)
        bound = logit_abs if logit_abs is not None else logit_rel * float(np.abs(ref_logits).max())
        for i in np.nonzero(~(dl <= bound))[0]:
            rep.fail(f"image {i}: max |dlogit| {dl[i]:.3e} > {bound:.3e}")
        rep.stats["max_abs_dlogit"] = float(dl.max()) if n else 0.0
        rep.stats["logit_bound"] = float(bound)
    if topk > 1 and n:
        order = np.argsort(-probs, axis=1, kind="stable")[:, :topk]
        rorder = np.argsort(-ref_probs, axis=1, kind="stable")[:, :topk]
        same_set = sum(set(a) == set(b) for a, b in zip(order, rorder))
        rep.stats[f"top{topk}_set_equal"] = int(same_set)
        for i in range(n):
            for a, b in zip(order[i], rorder[i]):
                if a != b and abs(float(ref_probs[i, a]) - float(ref_probs[i, b])) > 2 * prob_abs:
                    rep.fail(f"image {i}: top-{topk} order differs beyond tolerance ({a} vs {b})")
                    break
    return rep
