"""Dev helper: time the device metrics pass against the NumPy oracle (= the reference's host algorithm)."""
import sys, time
import numpy as np, torch
from oracle import metrics_oracle as mo
from video_fingerprint_b200 import metrics

n_videos = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
E, ids = mo.make_metric_embeddings(7, n_videos, 2, 0.1)
Ed = torch.from_numpy(E).cuda()
for _ in range(2):
    metrics.compute_retrieval_metrics(Ed, ids); metrics.compute_discrimination_metrics(Ed, ids)
torch.cuda.synchronize(); t0 = time.perf_counter()
r = metrics.compute_retrieval_metrics(Ed, ids); d = metrics.compute_discrimination_metrics(Ed, ids)
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"N = {len(E)}: device metrics (both functions) {1e3 * (t1 - t0):.1f} ms", {k: round(v, 4) for k, v in r.items()}, round(d["auc_roc"], 6))
if len(E) <= 12000:
    t0 = time.perf_counter(); ro = mo.retrieval_metrics(E, ids); do = mo.discrimination_metrics(E, ids); t1 = time.perf_counter()
    print(f"NumPy oracle (reference algorithm, host): {t1 - t0:.1f} s", {k: round(v, 4) for k, v in ro.items()}, round(do["auc_roc"], 6))
