"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports /root/reference/model.py and /root/reference/fingerprint.py as they are (PyAV is stubbed in
sys.modules because video decode is out of scope and ``av`` is not installed), feeds them the seeded
weights / clips from oracle/weights.py and records what the reference returns. The GPU box has no
/root/reference; tests there regenerate the same seeded inputs and compare with these files.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

from oracle.weights import make_clips, make_state_dict, state_dict_digest  # noqa: E402


def load_reference():
    sys.path.insert(0, REF)
    sys.modules.setdefault("av", types.ModuleType("av"))
    import fingerprint as ref_fp  # noqa
    import model as ref_model  # noqa

    return ref_model, ref_fp


def ref_embeddings(ref_model, sd, clips):
    """Scanner semantics (fingerprint.py:244-249): eval(), no_grad(), one B=1 forward per clip."""
    m = ref_model.create_model("attention")
    m.load_state_dict(sd, strict=True)
    m.eval()
    outs, pooled = [], []
    with torch.no_grad():
        for clip in clips:
            e, feats = m(clip.unsqueeze(0), return_features=True)
            outs.append(e[0].numpy())
            pooled.append(m.adaptive_pooling(feats)[0].numpy())
    return np.stack(outs).astype(np.float32), np.stack(pooled).astype(np.float32)


def fake_fingerprints(E, n_hash_dups=0):
    fps = {}
    for i, e in enumerate(E):
        fps[f"/videos/v{i:05d}.mp4"] = {
            "embedding": e,
            "path": f"/videos/v{i:05d}.mp4",
            "name": f"v{i:05d}.mp4",
            "size": 1000 + 7 * i,
            "file_hash": f"hash{i if i >= n_hash_dups else 0:05d}",
            "embedding_norm": float(np.linalg.norm(e)),
        }
    return fps


def groups_to_json(groups):
    return [[{"name": it["name"], "similarity": it["similarity"], "exact_duplicate": bool(it["exact_duplicate"])} for it in g] for g in groups]


def main():
    ref_model, ref_fp = load_reference()
    torch.set_num_threads(8)
    manifest = {}

    # ---- forward cases ------------------------------------------------------------------------
    cases = {
        # BASELINE.json configs[0] shape: 16 clips x 32 frames, uniform noise frames (not quantised)
        "cfg1_default": dict(wseed=0, wstyle="default", cseed=1234, lengths=[32] * 16, cstyle="noise", quantise=False),
        # discriminating variant: stress weights + per-video colour
        "cfg1_stress": dict(wseed=2, wstyle="stress", cseed=1235, lengths=[32] * 16, cstyle="colour", quantise=True),
        # variable-length clips, incl. the 10-frame minimum and a long clip
        "varlen_stress": dict(wseed=2, wstyle="stress", cseed=77, lengths=[10, 16, 37, 64, 23, 100, 11, 129, 300, 64, 65, 12], cstyle="colour", quantise=True),
        "t64_default": dict(wseed=5, wstyle="default", cseed=99, lengths=[64] * 8, cstyle="colour", quantise=True),
    }
    for name, c in cases.items():
        sd = make_state_dict(c["wseed"], c["wstyle"])
        clips = make_clips(c["cseed"], c["lengths"], c["cstyle"], c["quantise"])
        E, P = ref_embeddings(ref_model, sd, clips)
        np.savez_compressed(os.path.join(OUT, f"forward_{name}.npz"), embeddings=E, pooled=P, lengths=np.array(c["lengths"]))
        manifest[name] = dict(c, weights_sha256=state_dict_digest(sd), clip0_sum=float(clips[0].double().sum()))
        print(name, E.shape, "min pairwise cos", float((E @ E.T).min()))

    # the reference's own default initialisation under torch.manual_seed(0) (BASELINE cfg 1 wording)
    torch.manual_seed(0)
    m = ref_model.create_model("attention").eval()
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    clips = list(torch.rand(16, 32, 3, 64, 64, generator=torch.Generator().manual_seed(1234)))
    E0, P0 = ref_embeddings(ref_model, sd0, clips)
    np.savez_compressed(os.path.join(OUT, "forward_cfg1_refinit.npz"), embeddings=E0, pooled=P0, lengths=np.array([32] * 16))
    manifest["cfg1_refinit"] = dict(weights_sha256=state_dict_digest(sd0), keys=len(sd0), clip0_sum=float(clips[0].double().sum()),
                                    numel=int(sum(v.numel() for v in sd0.values())))

    # ---- find_duplicates (direct path) golden -------------------------------------------------
    scanner = ref_fp.VideoFingerprintScanner.__new__(ref_fp.VideoFingerprintScanner)
    dup = {}
    Es = np.load(os.path.join(OUT, "forward_cfg1_stress.npz"))["embeddings"]
    for thr in (0.95, 0.99):
        dup[f"cfg1_stress_thr{thr}"] = groups_to_json(scanner.find_duplicates(fake_fingerprints(Es, 3), thr, use_faiss=True))
    dup["cfg1_refinit_thr0.95"] = groups_to_json(scanner.find_duplicates(fake_fingerprints(E0), 0.95))
    # 150 synthetic unit vectors with planted near-duplicates; use_faiss=False forces the direct path at N>100
    rng = np.random.default_rng(11)
    X = rng.standard_normal((150, 256)).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    for src, dst, sigma in ((3, 40, 0.01), (3, 41, 0.015), (40, 90, 0.012), (7, 8, 0.0), (100, 149, 0.02), (120, 121, 0.03)):
        v = X[src] + sigma * rng.standard_normal(256).astype(np.float32)
        X[dst] = v / np.linalg.norm(v)
    np.save(os.path.join(OUT, "join_planted150.npy"), X)
    dup["planted150_thr0.95_direct"] = groups_to_json(scanner.find_duplicates(fake_fingerprints(X, 2), 0.95, use_faiss=False))
    dup["planted150_thr0.8_direct"] = groups_to_json(scanner.find_duplicates(fake_fingerprints(X, 2), 0.8, use_faiss=False))
    with open(os.path.join(OUT, "find_duplicates.json"), "w") as f:
        json.dump(dup, f, indent=1)
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("groups:", {k: [len(g) for g in v] for k, v in dup.items()})


if __name__ == "__main__":
    main()
