"""CPU oracle for the evaluation metrics  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy restatement of /root/reference/train.py:285-358 (``Trainer.compute_discrimination_metrics``) and :439-481
(``Trainer._compute_retrieval_metrics``): the full fp32 similarity matrix, boolean masks, per-row partial sorts. Two
deliberate differences, both only visible with exactly tied scores: ranks are canonicalised to (score descending, index
ascending) - the reference's ``argpartition`` / unstable ``argsort`` leave tie order unspecified - and AUC-ROC is written
in its Mann-Whitney form instead of calling scikit-learn (same value: ties count one half). Pinned by
tests/golden/metrics_*.json, which tests/golden/make_golden.py produces by calling the UNMODIFIED reference methods.
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np


def retrieval_metrics(embeddings: np.ndarray, video_ids: Sequence, k_values: Sequence[int] = (1, 5, 10)) -> Dict[str, float]:
    E = np.ascontiguousarray(embeddings, dtype=np.float32)
    ids = np.asarray(video_ids)
    n = len(E)
    n_videos = len(set(ids.tolist()))
    S = E @ E.T
    metrics: Dict[str, float] = {}
    order = []
    for i in range(n):                                     # train.py:454-476
        s = S[i].astype(np.float64)
        s[i] = -np.inf
        order.append(np.argsort(-s, kind="stable"))        # ties: ascending index
    for k in k_values:
        if k > n_videos - 1:
            continue
        metrics[f"R@{k}"] = float(np.mean([np.any(ids[order[i][:k]] == ids[i]) for i in range(n)])) if n else 0.0
    aps = []
    for i in range(n):
        # reference quirk kept: the row itself (score -inf, sorted LAST) carries its own video id, so it counts as one more
        # positive at rank n; every row therefore has an AP (1/n for a row without a true positive)
        positives = ids[order[i]] == ids[i]
        if positives.sum() > 0:
            precisions = np.cumsum(positives) / (np.arange(len(positives)) + 1)
            aps.append((precisions * positives).sum() / positives.sum())
    metrics["mAP"] = float(np.mean(aps)) if aps else 0.0
    return metrics


def discrimination_metrics(embeddings: np.ndarray, video_ids: Sequence, thresholds: Sequence[float] = (0.7, 0.8, 0.85, 0.9)) -> Dict[str, float]:
    E = np.ascontiguousarray(embeddings, dtype=np.float32)
    ids = np.asarray(video_ids)
    S = E @ E.T
    same = ids[None, :] == ids[:, None]
    diff = ~same
    np.fill_diagonal(same, False)
    np.fill_diagonal(diff, False)
    intra, inter = S[same], S[diff]
    both = len(intra) > 0 and len(inter) > 0
    m = {
        "intra_sim_mean": float(np.mean(intra.astype(np.float64))) if len(intra) else 0.0,
        "intra_sim_std": float(np.std(intra.astype(np.float64))) if len(intra) else 0.0,
        "inter_sim_mean": float(np.mean(inter.astype(np.float64))) if len(inter) else 0.0,
        "inter_sim_std": float(np.std(inter.astype(np.float64))) if len(inter) else 0.0,
    }
    m["separation_gap"] = m["intra_sim_mean"] - m["inter_sim_mean"] if both else 0.0
    for threshold in thresholds:
        if not both:
            continue
        t = np.float32(threshold)
        tp, fp = int(np.sum(intra >= t)), int(np.sum(inter >= t))
        fn, tn = len(intra) - tp, len(inter) - fp
        precision = tp / (tp + fp) if (tp + fp) > 0 else 0.0
        recall = tp / (tp + fn) if (tp + fn) > 0 else 0.0
        m[f"precision@{threshold:.2f}"] = precision
        m[f"recall@{threshold:.2f}"] = recall
        m[f"f1@{threshold:.2f}"] = 2 * precision * recall / (precision + recall) if (precision + recall) > 0 else 0.0
        m[f"fpr@{threshold:.2f}"] = fp / (fp + tn) if (fp + tn) > 0 else 0.0
    if both:
        srt = np.sort(intra)
        upper = np.searchsorted(srt, inter, side="right").astype(np.int64).sum()
        lower = np.searchsorted(srt, inter, side="left").astype(np.int64).sum()
        m["auc_roc"] = float((len(inter) * len(intra) - upper + 0.5 * (upper - lower)) / (len(intra) * len(inter)))
    else:
        m["auc_roc"] = 0.5
    return m


def make_metric_embeddings(seed: int, n_videos: int, clips_per_video, sigma: float, dim: int = 256):
    """Seeded synthetic validation set: one unit base vector per video, every clip = normalise(base + sigma * noise).
    ``clips_per_video`` is an int or a (lo, hi) range. Returns (embeddings fp32 (N, dim), video ids (N,) int64), shuffled."""
    rng = np.random.default_rng(seed)
    base = rng.standard_normal((n_videos, dim))
    base /= np.linalg.norm(base, axis=1, keepdims=True)
    if isinstance(clips_per_video, int):
        reps = np.full(n_videos, clips_per_video)
    else:
        reps = rng.integers(clips_per_video[0], clips_per_video[1] + 1, size=n_videos)
    ids = np.repeat(np.arange(n_videos), reps)
    E = base[ids] + sigma * rng.standard_normal((len(ids), dim))
    E /= np.linalg.norm(E, axis=1, keepdims=True)
    perm = rng.permutation(len(ids))
    return np.ascontiguousarray(E[perm], dtype=np.float32), ids[perm].astype(np.int64)


METRIC_CASES = {
    # name: (seed, n_videos, clips_per_video, sigma)
    "pairs_clean": (3, 300, 2, 0.03),          # trainer layout: clip1 + clip2 per video, well separated
    "pairs_noisy": (4, 300, 2, 0.12),          # positives often NOT rank 1 -> R@k < 1, mAP < 1, thresholds cut through both classes
    "ragged": (5, 200, (1, 4), 0.08),          # 1..4 clips per video: rows without positives, rows with several
    "tiny": (6, 3, 2, 0.05),                   # R@5 / R@10 skipped (k > n_videos - 1)
}
