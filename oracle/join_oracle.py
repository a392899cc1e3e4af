"""CPU oracle for the similarity join / top-k / duplicate grouping  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy fp32 restatement of

    /root/reference/fingerprint.py:450-480  find_duplicates            -> ``find_duplicates``
    /root/reference/fingerprint.py:482-513  _find_duplicates_direct    -> ``threshold_pairs`` + ``group_direct``
    /root/reference/fingerprint.py:515-548  _find_duplicates_faiss     -> ``topk_inner_product`` + ``group_topk``

The FAISS call sites (fingerprint.py:522-528: ``IndexFlatIP(d).add(E); search(E, k)``) bind the third-party
module **faiss-cpu 1.11.0** (uv.lock:113-114), whose source is not under /root/reference and which is not
installed here. ``IndexFlatIP.search`` is an exact brute-force fp32 inner-product search returning, per query
row, the k best scores in descending order; that published behaviour is what ``topk_inner_product`` restates.
FAISS's order among exactly tied scores is heap-implementation defined, so both sides are canonicalised to
(score descending, index ascending). No reference test or golden vector pins the FAISS boundary:
**the top-k path is "parity unpinned"** (it is checked against this restatement only). The direct path IS
pinned: tests/golden/make_golden.py runs the unmodified reference ``find_duplicates`` and stores its groups.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np


def threshold_pairs(E: np.ndarray, thr: float, Q: np.ndarray | None = None, block: int = 4096):
    """All (i, j, s) with s = <Q_i, E_j> >= thr in fp32, rows ascending then columns ascending
    (the order ``np.where(similarities[i] >= threshold)`` visits them, fingerprint.py:493-499)."""
    E = np.ascontiguousarray(E, dtype=np.float32)
    Q = E if Q is None else np.ascontiguousarray(Q, dtype=np.float32)
    ii, jj, ss = [], [], []
    for r0 in range(0, Q.shape[0], block):
        S = Q[r0 : r0 + block] @ E.T
        i, j = np.nonzero(S >= np.float32(thr))
        ii.append(i + r0)
        jj.append(j)
        ss.append(S[i, j])
    if not ii:
        return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float32)
    return np.concatenate(ii).astype(np.int64), np.concatenate(jj).astype(np.int64), np.concatenate(ss).astype(np.float32)


def topk_inner_product(Q: np.ndarray, DB: np.ndarray, k: int, block: int = 1024) -> Tuple[np.ndarray, np.ndarray]:
    """Exact flat inner-product top-k: (scores (Nq,k) fp32 descending, indices (Nq,k) int64), ties by index."""
    Q = np.ascontiguousarray(Q, dtype=np.float32)
    DB = np.ascontiguousarray(DB, dtype=np.float32)
    k = min(k, DB.shape[0])
    S_out = np.empty((Q.shape[0], k), np.float32)
    I_out = np.empty((Q.shape[0], k), np.int64)
    for r0 in range(0, Q.shape[0], block):
        S = Q[r0 : r0 + block] @ DB.T
        # stable sort on -score keeps ascending index among equal scores
        order = np.argsort(-S, axis=1, kind="stable")[:, :k]
        I_out[r0 : r0 + block] = order
        S_out[r0 : r0 + block] = np.take_along_axis(S, order, axis=1)
    return S_out, I_out


def group_direct(n: int, pi: Sequence[int], pj: Sequence[int], ps: Sequence[float]) -> List[List[Tuple[int, float]]]:
    """Greedy star grouping over a row-sorted pair list (fingerprint.py:495-511). Returns groups of
    (member index, similarity to the seed row)."""
    pi = np.asarray(pi)
    pj = np.asarray(pj)
    ps = np.asarray(ps)
    starts = np.searchsorted(pi, np.arange(n), side="left")
    ends = np.searchsorted(pi, np.arange(n), side="right")
    processed = set()
    groups = []
    for i in range(n):
        if i in processed:
            continue
        lo, hi = starts[i], ends[i]
        if hi - lo > 1:
            g = []
            for t in range(lo, hi):
                idx = int(pj[t])
                if idx not in processed:
                    processed.add(idx)
                    g.append((idx, float(ps[t])))
            if len(g) > 1:
                groups.append(g)
    return groups


def group_topk(S: np.ndarray, I: np.ndarray, thr: float) -> List[List[Tuple[int, float]]]:
    """Greedy grouping over per-row top-k lists (fingerprint.py:530-546)."""
    processed = set()
    groups = []
    for i in range(S.shape[0]):
        if i in processed:
            continue
        g = []
        for sim, idx in zip(S[i], I[i]):
            idx = int(idx)
            if sim >= thr and idx not in processed:
                processed.add(idx)
                g.append((idx, float(sim)))
        if len(g) > 1:
            groups.append(g)
    return groups


def find_duplicates(fingerprints: Dict[str, dict], similarity_threshold: float = 0.95, use_faiss: bool = True):
    """Same return structure as the reference method (fingerprint.py:450-480)."""
    if len(fingerprints) < 2:
        return []
    paths = list(fingerprints.keys())
    E = np.array([fingerprints[p]["embedding"] for p in paths]).astype("float32")
    if use_faiss and len(E) > 100:
        S, I = topk_inner_product(E, E, min(20, len(E)))
        raw = group_topk(S, I, similarity_threshold)
    else:
        pi, pj, ps = threshold_pairs(E, similarity_threshold)
        raw = group_direct(len(E), pi, pj, ps)
    out = []
    for g in raw:
        items = []
        for idx, sim in g:
            item = fingerprints[paths[idx]].copy()
            item["similarity"] = sim
            items.append(item)
        hashes = [it["file_hash"] for it in items]
        for it in items:
            it["exact_duplicate"] = hashes.count(it["file_hash"]) > 1
        out.append(items)
    return out
