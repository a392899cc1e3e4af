"""Development: where the time of the direct duplicate search goes at 1 M embeddings (join on the device, pair list to the
host, (i, j) sort, greedy grouping) - SURVEY.md section 8(f1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import make_join_data
from video_fingerprint_b200 import fingerprint as fp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dev = torch.device("cuda", 0)
E = make_join_data(n, dev, seed=11)
fp.threshold_join_device(E, 0.95)
torch.cuda.synchronize()
for it in range(2):
    t0 = time.perf_counter()
    i, j, s = fp.threshold_join_device(E, 0.95)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    i, j, s = i.cpu().numpy().astype(np.int64), j.cpu().numpy().astype(np.int64), s.cpu().numpy()
    t2 = time.perf_counter()
    order = np.lexsort((j, i))
    pi, pj, ps = i[order], j[order], s[order]
    t3 = time.perf_counter()
    groups = fp.group_pairs_direct(n, pi, pj, ps)
    t4 = time.perf_counter()
    print(f"n {n}: pairs {len(pi)}, groups {len(groups)}; join {1e3 * (t1 - t0):.1f} ms, to host {1e3 * (t2 - t1):.1f} ms, "
          f"lexsort {1e3 * (t3 - t2):.1f} ms, grouping {1e3 * (t4 - t3):.1f} ms, total {1e3 * (t4 - t0):.1f} ms", flush=True)
for it in range(2):
    t0 = time.perf_counter()
    a, b, c = fp.duplicate_pairs(E, 0.95)
    t1 = time.perf_counter()
    g2 = fp.group_pairs_direct(n, a, b, c)
    print(f"duplicate_pairs (join + device filter / sort + D2H) {1e3 * (t1 - t0):.1f} ms, {len(a)} pairs; grouping {1e3 * (time.perf_counter() - t1):.1f} ms; "
          f"groups {len(g2)}, same as the full list: {[[i for i, _ in g] for g in g2] == [[i for i, _ in g] for g in groups]}")
