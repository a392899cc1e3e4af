"""Golden vectors for the 3-D CNN model, produced by the UNMODIFIED reference module (/root/reference/model.py:406-512).
Run in the build container only:   python tests/golden/make_golden_3d.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))
from oracle.forward3d_oracle import CASES_3D, make_clips_3d, make_state_dict_3d  # noqa: E402

sys.path.insert(0, "/root/reference")
import model as ref_model  # noqa: E402

torch.set_num_threads(8)
out = {}
for name, (wseed, fs, cseed, n, t, stress) in CASES_3D.items():
    if name.endswith("refinit"):
        torch.manual_seed(wseed)
        m = ref_model.create_model("3d", frame_stride=fs).eval()
    else:
        m = ref_model.create_model("3d", frame_stride=fs).eval()
        m.load_state_dict(make_state_dict_3d(wseed, fs, stress=stress), strict=True)
    clips = make_clips_3d(cseed, n, t)
    with torch.no_grad():
        e = m(clips).numpy()
    out[name] = e.astype(np.float32)
    print(name, e.shape, "min pairwise cos", float((e @ e.T).min()))
np.savez_compressed(os.path.join(OUT, "forward3d.npz"), **out)
