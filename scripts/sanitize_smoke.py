"""Small-shape run of every hot kernel for compute-sanitizer (one tool per GPU call): forward (stem, conv GEMMs, CTA-pair conv4,
fused feed-forward, attention, pooling, head), similarity join (CTA-pair kernel + re-score), top-k (heap epilogue + merge + fallback)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import video_fingerprint_b200 as vfp
from video_fingerprint_b200 import _native
from oracle.weights import make_clips, make_state_dict
lib = _native.load()
m = vfp.create_model("attention").eval(); m.load_state_dict(make_state_dict(2, "stress"))
clips = make_clips(123, [16, 24, 10, 33, 70, 130], "colour")
e = m.fingerprint_clips(clips)
rng = np.random.default_rng(0)
E = rng.standard_normal((1500, 256)).astype(np.float32); E /= np.linalg.norm(E, axis=1, keepdims=True); E[700:720] = E[10:30]
i, j, s = vfp.threshold_join(E, 0.95)
S, I = vfp.topk_inner_product(E[:300], E, 10)
E[100:260] = E[5]
S2, I2 = vfp.topk_inner_product(E[90:270], E, 20)   # flagged rows -> exact fallback
torch.cuda.synchronize()
print("ok", tuple(e.shape), len(i), S.shape, S2.shape, "device error", lib.vfp_device_error_word())
