"""GPU, >= 2 devices: the sharded path end to end over NCCL - LPT-partitioned forward, all-gather of the embedding
shards, row-block threshold join with global indices, query-sharded top-k. Skipped on single-GPU boxes
(tests/test_sharding.py covers the host logic over gloo)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import join_oracle
from oracle.forward_oracle import fingerprint_clips
from oracle.weights import make_clips, make_state_dict

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


VARLEN = [500, 16, 300, 47, 128, 10, 211, 64, 33, 97, 250]


def planted3000():
    rng = np.random.default_rng(17)
    E = rng.standard_normal((3000, 256)).astype(np.float32)
    E /= np.linalg.norm(E, axis=1, keepdims=True)
    src, dst = rng.integers(0, 3000, 400), rng.integers(0, 3000, 400)
    for a, b in zip(src, dst):
        v = E[a] + rng.choice([0.0, 0.005, 0.0145, 0.0205, 0.03]) * rng.standard_normal(256).astype(np.float32)
        E[b] = v / np.linalg.norm(v)
    return E


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import video_fingerprint_b200 as vfp
    from video_fingerprint_b200 import sharding

    sd = make_state_dict(2, "stress")
    model = vfp.create_model("attention").eval()
    model.load_state_dict(sd)
    clips = make_clips(55, [12, 40, 17, 64, 23, 10, 31, 48, 20], "colour")
    emb = sharding.sharded_fingerprint(model, clips)           # (9, 256) on every rank, clip order
    # join: every rank contributes a contiguous shard of a planted matrix
    X = np.load(os.path.join(os.path.dirname(__file__), "golden", "join_planted150.npy"))
    lo, hi = sharding.row_block(X.shape[0], world, rank, align=1)
    local = torch.from_numpy(X[lo:hi]).cuda()
    pairs = sharding.sharded_threshold_join(local, 0.8)
    S, I = sharding.sharded_topk(local, local, 5)
    gathered_I, _ = sharding.all_gather_rows(I)
    # a larger planted set whose shard boundaries are NOT multiples of the 128-row tile, ragged shard sizes
    Y = planted3000()
    cuts = np.linspace(0, Y.shape[0], world + 1).astype(int) + np.r_[0, np.arange(1, world) * 7 % 50, 0]
    pairs2 = sharding.sharded_threshold_join(torch.from_numpy(Y[cuts[rank] : cuts[rank + 1]]).cuda(), 0.95)
    # variable-length clips including one at the scanner's max_frames = 500
    emb2 = sharding.sharded_fingerprint(model, make_clips(56, VARLEN, "colour"))
    if rank == 0:
        np.savez(os.path.join(out_dir, "out.npz"), emb=emb.cpu().numpy(), pi=pairs[0], pj=pairs[1], ps=pairs[2], topk=gathered_I.cpu().numpy(),
                 qi=pairs2[0], qj=pairs2[1], qs=pairs2[2], emb2=emb2.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_path_matches_oracle(tmp_path, golden_dir):
    world = min(torch.cuda.device_count(), 8)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    out = np.load(tmp_path / "out.npz")
    sd = make_state_dict(2, "stress")
    want = torch.stack(fingerprint_clips(sd, make_clips(55, [12, 40, 17, 64, 23, 10, 31, 48, 20], "colour"))).numpy()
    cos = (out["emb"] * want).sum(1) / (np.linalg.norm(out["emb"], axis=1) * np.linalg.norm(want, axis=1))
    assert cos.min() >= 0.9999
    X = np.load(os.path.join(golden_dir, "join_planted150.npy"))
    wi, wj, ws = join_oracle.threshold_pairs(X, 0.8)
    assert np.array_equal(out["pi"], wi) and np.array_equal(out["pj"], wj) and np.allclose(out["ps"], ws, atol=1e-5)
    _, wI = join_oracle.topk_inner_product(X, X, 5)
    assert np.array_equal(out["topk"], wI)
    Y = planted3000()
    yi, yj, ys = join_oracle.threshold_pairs(Y, 0.95)
    got = {(int(a), int(b)): float(c) for a, b, c in zip(out["qi"], out["qj"], out["qs"])}
    exp = {(int(a), int(b)): float(c) for a, b, c in zip(yi, yj, ys)}
    for key in set(got) ^ set(exp):            # only pairs inside the fp32 summation-order band may differ
        assert abs(got.get(key, exp.get(key)) - 0.95) < 1e-5, key
    assert len(exp) > 3000 and np.all(np.diff(out["qi"]) >= 0)
    want2 = torch.stack(fingerprint_clips(sd, make_clips(56, VARLEN, "colour"))).numpy()
    cos2 = (out["emb2"] * want2).sum(1) / (np.linalg.norm(out["emb2"], axis=1) * np.linalg.norm(want2, axis=1))
    assert cos2.min() >= 0.9999
