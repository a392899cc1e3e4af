"""Development: how much does the packing (pass size / neighbours) change an embedding? (rounding of the attention probabilities only)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import video_fingerprint_b200 as vfp
from oracle.weights import make_state_dict
from oracle.forward_oracle import forward_oracle
m = vfp.create_model("attention").eval(); m.load_state_dict(make_state_dict(2, "stress"))
g = torch.Generator().manual_seed(77)
lengths = [int(t) for t in torch.randint(10, 120, (620,), generator=g)]
frames = torch.randint(0, 256, (sum(lengths), 3, 64, 64), dtype=torch.uint8, generator=g).cuda()
m.pipelines, m.frames_per_pass = 1, 1 << 20
one = m.fingerprint_packed(frames, lengths).cpu()
m.pipelines, m.frames_per_pass = 2, 9000
m._workspaces.clear()
two = m.fingerprint_packed(frames, lengths).cpu()
d = (one - two).abs()
cos = torch.nn.functional.cosine_similarity(one.double(), two.double(), dim=1)
print("max abs diff", float(d.max()), "mean", float(d.mean()), "min cos", float(cos.min()), "clips differing > 1e-6:", int((d.max(dim=1).values > 1e-6).sum()), "of", len(lengths))
worst = int(cos.argmin()); print("worst clip", worst, "len", lengths[worst])
sd = make_state_dict(2, "stress"); cu = np.concatenate([[0], np.cumsum(lengths)])
for c in (worst, 0, 311):
    want = forward_oracle(sd, frames[cu[c]:cu[c+1]].cpu().float().div(255).unsqueeze(0))[0].double()
    for name, e in (("one", one), ("two", two)):
        print(c, name, "cos vs oracle", float(torch.dot(want, e[c].double()) / (want.norm() * e[c].double().norm())))
