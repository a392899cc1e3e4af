"""Development A/B of the threshold-join kernels on one GPU: generic tile kernel vs A-resident panel-major kernel, full
matrix vs upper triangle. Prints Gpairs/s (algorithmic: n^2 ordered pairs) and executed TFLOP/s."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_fingerprint_b200 as vfp
from video_fingerprint_b200 import _native
from bench import make_join_data

lib = _native.load()
dev = torch.device("cuda", 0)
out = []
for n in [int(a) for a in sys.argv[1:]] or [262144, 1048576]:
    E = make_join_data(n, dev)
    ref = None
    for name, tun in [("tile_full", {10: 0, 11: 0}), ("ares_full", {10: 1, 11: 0}), ("ares_tri", {10: 1, 11: 1}),
                      ("pair_full", {10: 2, 11: 0}), ("pair_tri", {10: 2, 11: 1})]:
        for k, v in {12: 512, **tun}.items():
            assert lib.vfp_set_tuning(k, v) == 0
        i, j, s = vfp.threshold_join_device(E, 0.95)
        cap = int(i.numel()) + 4096
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            i, j, s = vfp.threshold_join_device(E, 0.95, capacity=cap)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        key = torch.sort(i.to(torch.int64) * n + j.to(torch.int64)).values
        if ref is None:
            ref = key
        same = bool(key.numel() == ref.numel() and torch.equal(key, ref))
        executed = 0.5 if tun[11] else 1.0
        rec = {"n": n, "variant": name, "ms": round(ms, 3), "gpairs_s": round(n * n / ms / 1e6, 1), "executed_tflops": round(n * n * 512 * executed / ms / 1e9, 1),
               "pairs": int(i.numel()), "same_pair_set_as_tile_full": same, "device_error": lib.vfp_device_error_word()}
        print(json.dumps(rec), flush=True)
        out.append(rec)
    del E
