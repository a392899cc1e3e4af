"""Golden vectors for the frame preprocessing, produced by the reference's own code path: cv2.resize(..., INTER_AREA) +
centre crop exactly as /root/reference/fingerprint.py:186-207 does it (the method body is executed unmodified through
VideoFingerprintScanner._preprocess_frames; PyAV is stubbed). Run in the build container only (needs cv2 + /root/reference):

    python tests/golden/make_golden_preprocess.py
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))
from oracle.preprocess_oracle import PREPROCESS_CASES, make_frames  # noqa: E402

sys.modules.setdefault("av", types.ModuleType("av"))
sys.path.insert(0, "/root/reference")
import cv2  # noqa: E402
import fingerprint as ref_fp  # noqa: E402

scanner = ref_fp.VideoFingerprintScanner.__new__(ref_fp.VideoFingerprintScanner)
scanner.frame_size = 64
out = {}
for name, t, h, w in PREPROCESS_CASES:
    frames = make_frames(name, t, h, w)
    clip = scanner._preprocess_frames(list(frames))                       # (T, 3, 64, 64) float32 = uint8 / 255
    u8 = (clip * 255.0).round().to(dtype=__import__("torch").uint8).permute(0, 2, 3, 1).numpy()
    assert np.array_equal(u8.astype(np.float32) / 255.0, clip.permute(0, 2, 3, 1).numpy())
    out[name] = u8
    print(name, frames.shape, "->", u8.shape, int(u8.astype(np.int64).sum()))
np.savez_compressed(os.path.join(OUT, "preprocess.npz"), cv2_version=np.array(cv2.__version__), **out)
