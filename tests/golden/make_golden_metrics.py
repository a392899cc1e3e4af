"""Golden vectors for the evaluation metrics, produced by the UNMODIFIED reference trainer methods
(/root/reference/train.py:285-358 and :439-481). Run in the build container only:

    python tests/golden/make_golden_metrics.py

``Trainer.compute_discrimination_metrics`` / ``Trainer._compute_retrieval_metrics`` do not touch ``self``, so they are
called unbound on the seeded synthetic validation sets of oracle/metrics_oracle.py (no exactly tied scores in these sets:
the reference's tie order is unspecified).
"""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))
from oracle.metrics_oracle import METRIC_CASES, make_metric_embeddings  # noqa: E402

sys.modules.setdefault("av", types.ModuleType("av"))
sys.path.insert(0, "/root/reference")
import train as ref_train  # noqa: E402

out = {}
for name, (seed, n_videos, cpv, sigma) in METRIC_CASES.items():
    E, ids = make_metric_embeddings(seed, n_videos, cpv, sigma)
    r = ref_train.Trainer._compute_retrieval_metrics(None, torch.from_numpy(E), ids.tolist())
    d = ref_train.Trainer.compute_discrimination_metrics(None, E, ids)
    out[name] = {"n": int(len(E)), "checksum": float(E.astype(np.float64).sum()),
                 "retrieval": {k: float(v) for k, v in r.items()}, "discrimination": {k: float(v) for k, v in d.items()}}
    print(name, len(E), {k: round(float(v), 4) for k, v in r.items()}, round(float(d["auc_roc"]), 6), round(float(d["separation_gap"]), 4))
with open(os.path.join(OUT, "metrics.json"), "w") as f:
    json.dump(out, f, indent=1)
