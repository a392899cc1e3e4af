"""Development: a few self joins of n rows (for ncu captures). usage: dev_join_small.py [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_fingerprint_b200 as vfp
from bench import make_join_data
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
E = make_join_data(n, torch.device("cuda", 0))
for _ in range(3):
    i, j, s = vfp.threshold_join_device(E, 0.95)
torch.cuda.synchronize()
print("ok", i.numel())
