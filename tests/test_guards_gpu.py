"""GPU: out-of-bounds WRITE detection without compute-sanitizer (closed on this GPU pool, profiles/r02_sanitizer_closed.txt).
Every caller-owned buffer of the C ABI is carved out of a larger allocation with 1 MiB guard zones on both sides, filled with a
pattern; after the call the guards must be untouched and the device error word clear. Shapes are ragged on purpose (tiles,
pairs of tiles, key blocks and conv passes all end inside a partial unit)."""
import ctypes as C

import numpy as np
import pytest
import torch

import video_fingerprint_b200 as vfp
from oracle import join_oracle
from oracle.forward_oracle import fingerprint_clips
from oracle.weights import make_clips, make_state_dict
from video_fingerprint_b200 import _native
from video_fingerprint_b200.fingerprint import screen_margin

pytestmark = pytest.mark.gpu

GUARD = 1 << 20
PATTERN = 0xA5


class Guarded:
    """A device buffer of `nbytes` with guard zones; `.t` is the usable uint8 view (256-byte aligned)."""

    def __init__(self, nbytes):
        self.n = (int(nbytes) + 255) // 256 * 256
        self.raw = torch.full((self.n + 2 * GUARD,), PATTERN, dtype=torch.uint8, device="cuda")
        self.t = self.raw[GUARD : GUARD + self.n]

    def view(self, dtype, shape):
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        return self.t[:n].view(dtype).view(shape)

    def intact(self):
        return bool((self.raw[:GUARD] == PATTERN).all()) and bool((self.raw[GUARD + self.n :] == PATTERN).all())


def ptr(t):
    return C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("lengths,dtype", [([10, 33, 130, 64, 17], torch.uint8), ([64] * 5 + [11], torch.bfloat16), ([300, 12], torch.uint8)])
def test_forward_writes_stay_inside_its_buffers(lengths, dtype):
    lib = _native.load()
    sd = make_state_dict(2, "stress")
    m = vfp.create_model("attention").eval()
    m.load_state_dict(sd)
    clips = make_clips(71, lengths, "colour")
    x = torch.cat(clips)
    frames = (torch.round(x * 255).to(torch.uint8) if dtype == torch.uint8 else x.to(torch.bfloat16)).cuda().contiguous()
    total, n = sum(lengths), len(lengths)
    weights = m._ensure_native(torch.cuda.current_device())
    ws = Guarded(lib.vfp_forward_workspace_bytes(total, n))
    emb = Guarded(n * 256 * 4)
    feats = Guarded(total * 256 * 4)
    cu = (C.c_int32 * (n + 1))(*np.concatenate([[0], np.cumsum(lengths)]).tolist())
    lib.vfp_set_tuning(9, 1)
    rc = lib.vfp_forward(C.c_void_p(weights), ptr(frames), _native.FRAME_U8 if dtype == torch.uint8 else _native.FRAME_BF16, C.cast(cu, C.c_void_p), n,
                         ptr(emb.t), ptr(feats.t), ptr(ws.t), ws.n, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _native.check(rc, "vfp_forward")
    torch.cuda.synchronize()
    assert lib.vfp_device_error_word() == 0
    assert ws.intact() and emb.intact() and feats.intact()
    got = emb.view(torch.float32, (n, 256)).cpu()
    want = torch.stack(fingerprint_clips(sd, clips))
    cos = torch.nn.functional.cosine_similarity(got.double(), want.double(), dim=1)
    assert cos.min() >= 0.9999


@pytest.mark.parametrize("n", [130, 1500, 3001])
def test_join_and_topk_writes_stay_inside_their_buffers(n):
    lib = _native.load()
    rng = np.random.default_rng(n)
    E = rng.standard_normal((n, 256)).astype(np.float32)
    E /= np.linalg.norm(E, axis=1, keepdims=True)
    E[n // 2 : n // 2 + 40] = E[3:43]
    dev = torch.from_numpy(E).cuda()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    margin = float(screen_margin(dev, dev))
    cap = 8 * n
    ws = Guarded(lib.vfp_join_workspace_bytes(n, n, 4 * cap))
    oi, oj, os_, cnt = Guarded(cap * 4), Guarded(cap * 4), Guarded(cap * 4), Guarded(16)
    _native.check(lib.vfp_join_threshold(ptr(dev), ptr(dev), n, n, 256, 0, 0.95, margin, ptr(oi.t), ptr(oj.t), ptr(os_.t), cap, ptr(cnt.t),
                                         ptr(ws.t), ws.n, st), "vfp_join_threshold")
    torch.cuda.synchronize()
    assert lib.vfp_device_error_word() == 0
    assert all(g.intact() for g in (ws, oi, oj, os_, cnt))
    k = int(cnt.view(torch.int64, (2,))[0])
    gi = oi.view(torch.int32, (cap,))[:k].cpu().numpy().astype(np.int64)
    gj = oj.view(torch.int32, (cap,))[:k].cpu().numpy().astype(np.int64)
    order = np.lexsort((gj, gi))
    wi, wj, _ = join_oracle.threshold_pairs(E, 0.95)
    assert np.array_equal(gi[order], wi) and np.array_equal(gj[order], wj)

    kk = 10
    ws2 = Guarded(lib.vfp_topk_workspace_bytes(n, n, kk))
    S, I, fl = Guarded(n * kk * 4), Guarded(n * kk * 8), Guarded(16)
    _native.check(lib.vfp_topk_ip(ptr(dev), ptr(dev), n, n, 256, kk, margin, ptr(S.t), ptr(I.t), ptr(fl.t), ptr(ws2.t), ws2.n, st), "vfp_topk_ip")
    torch.cuda.synchronize()
    assert lib.vfp_device_error_word() == 0
    assert all(g.intact() for g in (ws2, S, I, fl))
    _, wI = join_oracle.topk_inner_product(E, E, kk)
    gI = I.view(torch.int64, (n, kk)).cpu().numpy()
    assert (gI == wI).mean() > 0.999   # near-ties may swap (tests/test_topk_gpu.py states the rule)


def test_repeated_runs_are_bit_identical():
    """A data race between the warp-specialised roles (TMA producer / UMMA issuer / TMEM epilogues / activation warps) would
    show up as run-to-run differences; the kernels have no atomics on floating-point data, so results must repeat bit for bit.
    More frames than SMs and more than one CTA pair per kernel, so every ring and double buffer wraps."""
    m = vfp.create_model("attention").eval()
    m.load_state_dict(make_state_dict(2, "stress"))
    g = torch.Generator().manual_seed(5)
    lengths = [int(t) for t in torch.randint(10, 200, (300,), generator=g)]
    frames = torch.randint(0, 256, (sum(lengths), 3, 64, 64), dtype=torch.uint8, generator=g).cuda()
    first = m.fingerprint_packed(frames, lengths).clone()
    for _ in range(8):
        assert torch.equal(m.fingerprint_packed(frames, lengths), first)
    E = first / first.norm(dim=1, keepdim=True)
    a = vfp.threshold_join(E, 0.9)
    b = vfp.threshold_join(E, 0.9)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    # the CTA-pair join over many pair tiles and panels (remote mbarrier arrives between the two CTAs of a pair)
    gen = torch.Generator(device="cuda").manual_seed(9)
    big = torch.randn((60_000, 256), generator=gen, device="cuda")
    big /= big.norm(dim=1, keepdim=True)
    big[30_000:30_500] = big[:500] + 0.01 * torch.randn((500, 256), generator=gen, device="cuda")
    big[30_000:30_500] /= big[30_000:30_500].norm(dim=1, keepdim=True)
    ref = vfp.threshold_join(big, 0.95)
    assert len(ref[0]) >= 60_000 + 2 * 400
    for _ in range(4):
        again = vfp.threshold_join(big, 0.95)
        assert all(np.array_equal(x, y) for x, y in zip(ref, again))
    assert _native.load().vfp_device_error_word() == 0
