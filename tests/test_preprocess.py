"""Frame preprocessing (/root/reference/fingerprint.py:186-214): INTER_AREA resize to a short
side of 64 + centre crop (down-scaling and, for frames with a side below 64 px, up-scaling). Integer / byte work -> bit-exact
everywhere.

CPU: the NumPy oracle against tests/golden/preprocess.npz (what the reference's unmodified ``_preprocess_frames`` returned
here, i.e. cv2 4.13) and, where cv2 is importable, live against cv2 on random sizes. GPU (-m gpu): vfp_preprocess_frames
through the C ABI against the golden file and the oracle, and end to end into the forward.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import preprocess_oracle as po  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "preprocess.npz"))
SMALL = [c for c in po.PREPROCESS_CASES if c[2] * c[3] <= 400 * 400]      # the pure-NumPy oracle is slow on big frames


@pytest.mark.parametrize("case", SMALL, ids=[c[0] for c in SMALL])
def test_oracle_matches_reference_golden(case):
    name, t, h, w = case
    frames = po.make_frames(name, t, h, w)[:1]
    assert np.array_equal(po.preprocess_frames(frames), GOLD[name][:1])


def test_oracle_matches_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for h, w in [(97, 143), (150, 100), (128, 256), (64, 90), (256, 256), (199, 64), (48, 50), (63, 64), (17, 9), (1, 1), (3, 200), (40, 40)]:
        frame = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        nh, nw = po.target_size(h, w)
        assert np.array_equal(po.resize_area(frame, nw, nh), cv2.resize(frame, (nw, nh), interpolation=cv2.INTER_AREA)), (h, w)


def test_target_size_truncates_like_the_reference():
    assert po.target_size(1080, 1920) == (64, 113) and po.target_size(640, 360) == (113, 64) and po.target_size(64, 64) == (64, 64)


@pytest.mark.gpu
@pytest.mark.parametrize("case", po.PREPROCESS_CASES, ids=[c[0] for c in po.PREPROCESS_CASES])
def test_device_preprocess_is_bit_exact(case):
    import video_fingerprint_b200 as vfp

    name, t, h, w = case
    frames = po.make_frames(name, t, h, w)
    got = vfp.preprocess_frames_device(frames).cpu().numpy()
    assert got.shape == (t, 64, 64, 3)
    diff = np.argwhere(got != GOLD[name])
    assert len(diff) == 0, (name, len(diff), diff[:4])


@pytest.mark.gpu
def test_device_preprocess_inputs_errors_and_forward():
    import video_fingerprint_b200 as vfp

    frames = po.make_frames("odd", 12, 150, 231)
    want = po.preprocess_frames(frames[:2])
    a = vfp.preprocess_frames_device(list(frames))                     # list of arrays
    b = vfp.preprocess_frames_device(torch.from_numpy(frames).cuda())  # device tensor
    assert torch.equal(a, b) and np.array_equal(a[:2].cpu().numpy(), want)
    rng = np.random.default_rng(3)
    for h, w in [(48, 100), (5, 9), (63, 63), (1, 1), (30, 500)]:                # a side below 64 px: INTER_AREA up-scales
        tiny = rng.integers(0, 256, (3, h, w, 3), dtype=np.uint8)
        assert np.array_equal(vfp.preprocess_frames_device(tiny).cpu().numpy(), po.preprocess_frames(tiny)), (h, w)
    with pytest.raises(ValueError):
        vfp.preprocess_frames_device(np.zeros((1, 100, 100, 3), np.float32))
    # decoded frames -> embedding, against the float clip the reference would build from the same preprocessing
    torch.manual_seed(0)
    model = vfp.create_model("attention").eval()
    scanner = vfp.VideoFingerprintScanner(model=model)
    e1 = scanner.extract_fingerprint_from_decoded(frames)
    clip = scanner._preprocess_frames(frames)                          # (T, 3, 64, 64) float in [0, 1]
    assert clip.shape == (12, 3, 64, 64) and float(clip.max()) <= 1.0
    e2 = model(clip.unsqueeze(0))[0].cpu().numpy()
    assert float(np.dot(e1, e2) / (np.linalg.norm(e1) * np.linalg.norm(e2))) > 0.99998
    assert scanner.extract_fingerprint_from_decoded(frames[:9]) is None
