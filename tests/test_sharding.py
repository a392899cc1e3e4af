"""CPU, world_size 2 over gloo: the host side of the multi-GPU path (variable-size all-gather, row-block join with
global indices, gather of the pair lists). The device join is replaced by the oracle's CPU join here; the GPU
tests run the same functions with the real kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import join_oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cpu_join(db, thr, q, q_row0):
    i, j, s = join_oracle.threshold_pairs(db.numpy(), thr, Q=q.numpy())
    return torch.from_numpy(i + q_row0), torch.from_numpy(j), torch.from_numpy(s)


def _worker(rank, world, port, X, thr, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from video_fingerprint_b200.sharding import all_gather_rows, sharded_threshold_join

    n = X.shape[0]
    bounds = [0, 70, n]  # deliberately uneven shards
    local = torch.from_numpy(X[bounds[rank] : bounds[rank + 1]])
    full, counts = all_gather_rows(local)
    assert counts == [70, n - 70] and torch.equal(full, torch.from_numpy(X))
    res = sharded_threshold_join(local, thr, join_fn=_cpu_join)
    if rank == 0:
        np.savez(os.path.join(out_dir, "pairs.npz"), i=res[0], j=res[1], s=res[2])
    else:
        assert res is None
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_join_world2(tmp_path, golden_dir):
    X = np.load(os.path.join(golden_dir, "join_planted150.npy"))
    mp.spawn(_worker, args=(2, _free_port(), X, 0.8, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "pairs.npz")
    wi, wj, ws = join_oracle.threshold_pairs(X, 0.8)
    assert np.array_equal(got["i"], wi) and np.array_equal(got["j"], wj)
    assert np.allclose(got["s"], ws, atol=1e-6)
