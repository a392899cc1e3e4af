"""BASELINE.json configs 3-5 on one GPU (the multi-GPU forms are the same calls under torchrun via sharding.py):

  cfg3  variable-length clips (16..300 frames), packed forward, parity of a subset against the oracle
  cfg4  all-pairs threshold join over N x 256 unit vectors with planted duplicates (N = 1M by default)
  cfg5  top-10 inner-product search, DB 10M x 256 vs 100k queries

Prints one JSON line per config. Run on the GPU box: python tests/tools/run_configs.py [--small]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import video_fingerprint_b200 as vfp  # noqa: E402
from oracle.forward_oracle import fingerprint_clips  # noqa: E402
from oracle.weights import make_state_dict  # noqa: E402


def timed(fn, reps=1):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return out, a.elapsed_time(b) / reps


def cfg3(n_clips):
    rng = np.random.default_rng(7)
    lengths = rng.integers(16, 301, n_clips).tolist()
    total = int(sum(lengths))
    sd = make_state_dict(2, "stress")
    model = vfp.create_model("attention").eval()
    model.load_state_dict(sd)
    g = torch.Generator(device="cuda").manual_seed(3)
    frames = torch.randint(0, 256, (total, 3, 64, 64), dtype=torch.uint8, device="cuda", generator=g)
    model.fingerprint_packed(frames, lengths)  # warm-up
    emb, ms = timed(lambda: model.fingerprint_packed(frames, lengths))
    # parity of 24 clips (short, median, longest) against the per-clip B=1 oracle
    order = np.argsort(lengths)
    pick = list(order[:8]) + list(order[n_clips // 2 - 4 : n_clips // 2 + 4]) + list(order[-8:])
    cu = np.concatenate([[0], np.cumsum(lengths)])
    clips = [frames[cu[i] : cu[i + 1]].float().cpu() / 255.0 for i in pick]
    want = torch.stack(fingerprint_clips(sd, clips))
    got = emb[pick].cpu()
    cos = torch.nn.functional.cosine_similarity(got.double(), want.double(), dim=1)
    flops = sum(39_806_976 * t + 4096 * t * t + 524_288 for t in lengths)
    return {"config": "cfg3 variable-length forward", "clips": n_clips, "frames": total, "ms": ms, "videos_per_s": n_clips / ms * 1e3,
            "frames_per_s": total / ms * 1e3, "tflops": flops / ms / 1e9, "min_cosine_vs_oracle_24_clips": float(cos.min())}


def planted(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    E = torch.randn((n, 256), generator=g, device="cuda")
    E /= E.norm(dim=1, keepdim=True)
    n_dup = n // 50
    src = torch.randint(0, n, (n_dup,), generator=g, device="cuda")
    dst = torch.randint(0, n, (n_dup,), generator=g, device="cuda")
    sigma = torch.tensor([0.0, 0.005, 0.0145, 0.0205, 0.03], device="cuda")[torch.randint(0, 5, (n_dup,), generator=g, device="cuda")]
    v = E[src] + sigma[:, None] * torch.randn((n_dup, 256), generator=g, device="cuda")
    E[dst] = v / v.norm(dim=1, keepdim=True)
    return E.contiguous()


def cfg4(n):
    E = planted(n, 11)
    i, j, s = vfp.threshold_join_device(E, 0.95)
    cap = int(i.numel()) + 4096
    (i, j, s), ms = timed(lambda: vfp.threshold_join_device(E, 0.95, capacity=cap))
    # property checks: symmetric pair set, full diagonal, scores in range
    ii, jj = i.long(), j.long()
    key = ii * n + jj
    key_t = jj * n + ii
    sym = bool(torch.equal(torch.sort(key).values, torch.sort(key_t).values))
    return {"config": "cfg4 all-pairs threshold join", "n": n, "threshold": 0.95, "ms": ms, "gpairs_per_s": n * n / ms / 1e6,
            "tflops": n * n * 512 / ms / 1e9, "pairs": int(i.numel()), "diagonal_complete": int((ii == jj).sum()) == n, "symmetric": sym,
            "min_score": float(s.min())}


def cfg5(n_db, n_q, k):
    g = torch.Generator(device="cuda").manual_seed(21)
    DB = torch.empty((n_db, 256), device="cuda")
    for s0 in range(0, n_db, 1 << 20):
        e0 = min(n_db, s0 + (1 << 20))
        blk = torch.randn((e0 - s0, 256), generator=g, device="cuda")
        DB[s0:e0] = blk / blk.norm(dim=1, keepdim=True)
    Q = torch.randn((n_q, 256), generator=torch.Generator(device="cuda").manual_seed(22), device="cuda")
    Q /= Q.norm(dim=1, keepdim=True)
    hit = torch.arange(0, n_q, 100, device="cuda")          # 1 % of the queries are noisy copies of database rows
    src = (hit * 97) % n_db
    v = DB[src] + 0.02 * torch.randn((hit.numel(), 256), device="cuda", generator=g)
    Q[hit] = v / v.norm(dim=1, keepdim=True)
    DB[1:1001:2] = DB[0:1000:2]                               # exact duplicates -> tied scores
    vfp.topk_inner_product_device(Q[:1024], DB, k)           # warm-up
    (S, I), ms = timed(lambda: vfp.topk_inner_product_device(Q, DB, k))
    # exact check of 256 queries against fp32 brute force on the GPU (torch as the checker)
    torch.backends.cuda.matmul.allow_tf32 = False
    sel = torch.cat([hit[:128], torch.arange(1, 129, device="cuda")])
    full = Q[sel] @ DB.T
    wS, wI = torch.sort(full, dim=1, descending=True, stable=True)
    ok_idx = bool(torch.equal(I[sel], wI[:, :k])) or bool(((I[sel] != wI[:, :k]) & ((S[sel] - wS[:, :k]).abs() > 2e-6)).sum() == 0)
    return {"config": "cfg5 flat-IP top-k", "db": n_db, "queries": n_q, "k": k, "ms": ms, "gpairs_per_s": n_db * n_q / ms / 1e6,
            "tflops": n_db * n_q * 512 / ms / 1e9, "planted_found": bool(torch.equal(I[hit, 0], src)),
            "checked_256_queries_exact": ok_idx, "max_score_err": float((S[sel] - wS[:, :k]).abs().max())}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true")
    a = ap.parse_args()
    print(json.dumps(cfg3(500 if a.small else 4000)), flush=True)
    print(json.dumps(cfg4(131072 if a.small else 1 << 20)), flush=True)
    print(json.dumps(cfg5(1_000_000 if a.small else 10_000_000, 10_000 if a.small else 100_000, 10)), flush=True)
