"""Developer check on a GPU box: forward + join parity against the oracle / golden files."""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.weights import make_state_dict, make_clips
from oracle.forward_oracle import forward_oracle, fingerprint_clips
from oracle import join_oracle
import video_fingerprint_b200 as vfp

torch.backends.cuda.matmul.allow_tf32 = False
def cos(a, b):
    a = a.double(); b = b.double()
    return (a * b).sum(-1) / (a.norm(dim=-1) * b.norm(dim=-1))

man = json.load(open(os.path.join(ROOT, "tests/golden/manifest.json")))
for name in ["cfg1_default", "cfg1_stress", "varlen_stress", "t64_default"]:
    c = man[name]
    sd = make_state_dict(c["wseed"], c["wstyle"])
    clips = make_clips(c["cseed"], c["lengths"], c["cstyle"], c["quantise"])
    gold = np.load(os.path.join(ROOT, f"tests/golden/forward_{name}.npz"))
    m = vfp.create_model("attention"); m.load_state_dict(sd); m.eval()
    t0 = time.time()
    emb = m.fingerprint_clips(clips).cpu()
    torch.cuda.synchronize()
    cs = cos(emb, torch.from_numpy(gold["embeddings"]))
    print(f"{name}: min cos vs golden {cs.min():.7f}  mean {cs.mean():.7f}  max|d| {(emb - torch.from_numpy(gold['embeddings'])).abs().max():.3e}  ({time.time()-t0:.2f}s)")
    if name == "cfg1_stress":
        # stage bisect: features after the attention blocks
        st = {}
        x = torch.stack(clips)
        ref = forward_oracle(sd, x, st)
        e2, feats = m(x.cuda(), return_features=True)
        fr = st["attn3"]
        print("   features rel err", ((feats.cpu() - fr).norm() / fr.norm()).item(), " emb cos", cos(e2.cpu(), ref).min().item())
        # centred cosine
        ec = emb - emb.mean(0, keepdim=True); gc = torch.from_numpy(gold["embeddings"]); gc = gc - gc.mean(0, keepdim=True)
        print("   centred cos min", cos(ec, gc).min().item())
# join
X = np.load(os.path.join(ROOT, "tests/golden/join_planted150.npy"))
for thr in (0.95, 0.8, 0.2):
    i, j, s = vfp.threshold_join(X, thr)
    oi, oj, os_ = join_oracle.threshold_pairs(X, thr)
    same = len(i) == len(oi) and np.array_equal(i, oi) and np.array_equal(j, oj)
    print(f"join thr={thr}: {len(i)} pairs, oracle {len(oi)}, identical={same}", (np.abs(s - os_).max() if same and len(s) else None))
rng = np.random.default_rng(0)
E = rng.standard_normal((20000, 256)).astype(np.float32); E /= np.linalg.norm(E, axis=1, keepdims=True)
E[5000:5100] = E[100:200] + 0.01 * rng.standard_normal((100, 256)).astype(np.float32)
E[5000:5100] /= np.linalg.norm(E[5000:5100], axis=1, keepdims=True)
t0 = time.time(); i, j, s = vfp.threshold_join(E, 0.9); t1 = time.time()
oi, oj, os_ = join_oracle.threshold_pairs(E, 0.9)
print("join 20000:", len(i), len(oi), np.array_equal(i, oi) and np.array_equal(j, oj), f"{t1-t0:.3f}s")
print("device error word", hex(vfp._native.load().vfp_device_error_word()))
