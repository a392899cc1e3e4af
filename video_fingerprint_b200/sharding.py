"""Multi-GPU layer: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch on the GPU box, gloo in
the CPU tests). The reference has no distributed code at all; the hot path shards as
follows:

* forward: clips are independent units -> length-balanced (LPT) partition over ranks, weights replicated,
  NO collective while fingerprinting;
* join: ONE all-gather of the (n_r, 256) fp32 embedding shards, after which every rank owns the full matrix
  and joins its contiguous row block against all columns (`q_row0` makes the pair indices global);
* top-k: queries are sharded the same way, the database is the all-gathered matrix -> no merge step;
* the greedy grouping is sequential in the seed index by definition, so the per-rank pair lists are gathered
  and grouped on rank 0 (host).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def partition_clips(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of clips to ranks by frame count (forward cost is ~linear
    in T). Deterministic; each rank's list is in ascending clip order."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0] * world
    parts: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        parts[r].append(i)
        loads[r] += int(lengths[i])
    return [sorted(p) for p in parts]


def row_block(n: int, world: int, rank: int, align: int = 128) -> Tuple[int, int]:
    """Contiguous row block of rank `rank`; interior boundaries are multiples of the 128-row GEMM tile."""
    per = -(-n // world)
    per = -(-per // align) * align
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi


def _world(group=None) -> Tuple[int, int]:
    if not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def all_gather_rows(local: torch.Tensor, group=None) -> Tuple[torch.Tensor, List[int]]:
    """All-gather row shards with possibly different row counts. Returns (concatenation in rank order, counts)."""
    world, _ = _world(group)
    if world == 1:
        return local, [local.shape[0]]
    counts_t = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    all_counts = [torch.zeros_like(counts_t) for _ in range(world)]
    dist.all_gather(all_counts, counts_t, group=group)
    counts = [int(c.item()) for c in all_counts]
    width = max(counts)
    padded = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    buf = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    parts = [buf[r * width : r * width + counts[r]] for r in range(world)]
    return torch.cat(parts, dim=0), counts


def sharded_fingerprint(model, clips: Sequence[torch.Tensor], group=None) -> torch.Tensor:
    """Every rank fingerprints its LPT share of `clips` (all ranks pass the same list; only the local share is
    touched), then the shards are all-gathered and put back into clip order. Returns (n, D) on every rank."""
    world, rank = _world(group)
    lengths = [int(c.shape[0]) for c in clips]
    parts = partition_clips(lengths, world)
    mine = parts[rank]
    if mine:
        local = model.fingerprint_clips([clips[i] for i in mine])
    else:
        local = torch.empty((0, model.embedding_dim), dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
    gathered, counts = all_gather_rows(local, group)
    order = torch.tensor([i for p in parts for i in p], dtype=torch.int64, device=gathered.device)
    out = torch.empty_like(gathered)
    out[order] = gathered
    return out


def sharded_threshold_join(
    local_embeddings: torch.Tensor,
    thr: float,
    group=None,
    join_fn: Optional[Callable] = None,
    gather_to: Optional[int] = 0,
) -> Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]]:
    """Row-block sharded all-pairs join. `local_embeddings` is this rank's (n_r, 256) shard of the embedding
    matrix, shards being consecutive in rank order. One all-gather, then each rank joins rows
    [row0, row0+n_r) x all columns. Returns the (i, j, s) lists sorted by (i, j) on rank `gather_to`
    (None elsewhere), or on every rank if gather_to is None.
    `join_fn(db, thr, q, q_row0) -> (i, j, s)` defaults to the device join; the CPU tests inject the oracle."""
    world, rank = _world(group)
    if join_fn is None:
        from .fingerprint import threshold_join_device

        def join_fn(db, thr_, q, q_row0):  # noqa: E306
            return threshold_join_device(db, thr_, q=q, q_row0=q_row0)

    full, counts = all_gather_rows(local_embeddings.float().contiguous(), group)
    row0 = sum(counts[:rank])
    if local_embeddings.shape[0] > 0:
        i, j, s = join_fn(full, thr, local_embeddings.float().contiguous(), row0)
        i, j, s = torch.as_tensor(i), torch.as_tensor(j), torch.as_tensor(s)
    else:
        i = j = torch.zeros(0, dtype=torch.int64)
        s = torch.zeros(0, dtype=torch.float32)
    trip = torch.stack([i.to(torch.float64), j.to(torch.float64), s.to(torch.float64)], dim=1).to(full.device)
    if world > 1:
        trip, _ = all_gather_rows(trip, group)
    if gather_to is not None and rank != gather_to:
        return None
    t = trip.cpu().numpy()
    pi, pj, ps = t[:, 0].astype(np.int64), t[:, 1].astype(np.int64), t[:, 2].astype(np.float32)
    order = np.lexsort((pj, pi))
    return pi[order], pj[order], ps[order]


def sharded_topk(local_queries: torch.Tensor, local_db: torch.Tensor, k: int, group=None, topk_fn: Optional[Callable] = None):
    """Query-sharded flat inner-product top-k: the database shards are all-gathered once, every rank searches its
    own queries against the full database (no merge). Returns this rank's (scores, indices)."""
    if topk_fn is None:
        from .fingerprint import topk_inner_product_device as topk_fn
    db, _ = all_gather_rows(local_db.float().contiguous(), group)
    return topk_fn(local_queries.float().contiguous(), db, k)
