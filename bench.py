#!/usr/bin/env python
"""Headline benchmark of the B200 duplicate-detection hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): batched fingerprint forward of 10 000 synthetic clips x 64 frames @ 64x64
per GPU, bf16 frames resident in HBM, random-init weights of the reference architecture. One "step" = one
pass of `model.fingerprint_packed` over all clips of the rank. Launched under torchrun for N > 1 (one rank
per GPU, clips sharded, no data-path collective in the forward => weak scaling).

One JSON line on stdout (rank 0): value = whole-job videos/s (device-resident inputs), `e2e` = the same metric
through the public API with HOST (pinned) uint8 frames, host<->device copies inside the timed region,
`roofline` for the dominant kernel (per-stage CUDA-event times from the library's stage profiler),
`cpu_baseline` = the oracle's reference-semantics B=1 loop on the box's host cores, `join` = the all-pairs
cosine threshold join on synthetic unit vectors with planted duplicates (second half of the BASELINE metric).
`--impl reference` times the reference algorithm's CPU path (oracle port; the reference itself is not
installable on the GPU box) on the same config.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CLIPS = 10_000
T_FRAMES = 64
FLOPS_PER_CLIP = 39_806_976 * T_FRAMES + 4096 * T_FRAMES * T_FRAMES + 524_288
METRIC = "fingerprint videos/s (64 frames@64^2)"

# algorithmic FLOPs per clip of each profiled stage (2 FLOP per MAC), T = 64
_T = T_FRAMES
STAGE_FLOPS = {
    "conv1_stem": 2 * 1024 * 32 * 75 * _T,
    "conv2_igemm": 2 * 256 * 64 * 288 * _T,
    "stem_fused": 2 * (1024 * 32 * 75 + 256 * 64 * 288) * _T,   # conv1 + conv2 in one kernel
    "conv3_igemm": 2 * 64 * 128 * 576 * _T,
    "conv4_igemm_pool": 2 * 16 * 256 * 1152 * _T,
    "token_embed_gemm": 2 * (256 * 128 + 128 * 256) * _T,
    "qkv_gemm": 4 * 2 * 256 * 768 * _T,
    "attention": 4 * 4 * _T * _T * 256,
    "out_proj_gemm": 4 * 2 * 256 * 256 * _T,
    "mlp1_gemm_gelu": 4 * 2 * 256 * 1024 * _T,
    "mlp2_gemm": 4 * 2 * 1024 * 256 * _T,
    "pool_logits_gemm": 2 * 256 * 256 * _T,
}


# algorithmic HBM bytes per clip of the stages whose binding roofline is memory, not the tensor pipe
STAGE_BYTES = {
    "conv1_stem": (24576 + 65536) * _T,           # bf16 frame in, bf16 32x32x32 out
    "conv2_igemm": (65536 + 32768) * _T,          # conv1 output in, 16x16x64 out
    "stem_fused": (24576 + 32768) * _T,           # bf16 frame in, 16x16x64 out (conv1's output never leaves the SM)
    "conv3_igemm": (32768 + 16384) * _T,
    "temporal_conv": 2 * (1024 + 1024) * _T,      # two blocks, fp32 stream in + out
    "layernorm": 8 * (1024 + 512 + 1024 + 512) * _T,
}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), tc_burst=float(p["bf16_tflops"]), tc_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    except Exception:
        return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms; only the samples that fall INSIDE the timed region
    (host timestamps taken around it) are reported. Started before the warm-up so nvidia-smi's start-up is not lost."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t_begin: float, t_end: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        inside = [r for t, r in self.rows if t_begin <= t <= t_end + 0.05 and len(r) >= 6]
        rows = inside if inside else [r for _, r in self.rows[-3:] if len(r) >= 6]
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "samples_inside_timed_region": len(inside)}


def dist_setup(n_gpus: int):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x: float, world: int, dev) -> float:
    if world == 1:
        return x
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# --------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port of the reference's --device cpu path: B=1 forward loop, fingerprint.py:244-249)
# --------------------------------------------------------------------------------------------------
def cpu_reference_rate(budget_s: float, max_clips: int):
    from oracle.forward_oracle import forward_oracle
    from oracle.weights import make_state_dict

    sd = make_state_dict(0, "default")
    g = torch.Generator().manual_seed(4321)
    clip = torch.rand((1, T_FRAMES, 3, 64, 64), generator=g)
    forward_oracle(sd, clip)  # warm-up
    n, t0 = 0, time.perf_counter()
    while n < max_clips and (time.perf_counter() - t0 < budget_s or n < 4):
        forward_oracle(sd, clip)
        n += 1
    dt = time.perf_counter() - t0
    return n / dt, n, dt


def run_reference(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)  # torchrun pins OMP_NUM_THREADS=1; the reference arm gets every host core
    threads = torch.get_num_threads()
    per_step = max(4.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_reference_rate(per_step / 4, 64)
    rates, clips, secs = [], 0, 0.0
    for _ in range(args.steps):
        r, n, dt = cpu_reference_rate(per_step, 512)
        rates.append(r)
        clips += n
        secs += dt
    value = clips / secs
    sample = f"{clips} clips x {T_FRAMES} frames in {secs:.1f}s, B=1 forwards (oracle port of the reference CPU path; the reference itself cannot be shipped to the GPU box)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "videos/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * secs / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": f"fingerprint forward, {T_FRAMES} frames @ 64x64 per clip, bounded sample of the 10k-clip workload"},
        "cpu_baseline": {"value": value, "unit": "videos/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "videos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "world_size_env": world,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def make_join_data(n: int, dev, seed: int = 11):
    g = torch.Generator(device=dev).manual_seed(seed)
    E = torch.randn((n, 256), generator=g, device=dev)
    E = E / E.norm(dim=1, keepdim=True)
    n_dup = n // 50
    src = torch.randint(0, n, (n_dup,), generator=g, device=dev)
    dst = torch.randint(0, n, (n_dup,), generator=g, device=dev)
    sigma = torch.tensor([0.0, 0.005, 0.0145, 0.0205, 0.03], device=dev)[torch.randint(0, 5, (n_dup,), generator=g, device=dev)]
    v = E[src] + sigma[:, None] * torch.randn((n_dup, 256), generator=g, device=dev)
    E[dst] = v / v.norm(dim=1, keepdim=True)
    return E.contiguous()


def run_ours(args):
    import video_fingerprint_b200 as vfp
    from video_fingerprint_b200 import _native

    world, rank, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    lib = _native.load()
    if args.stem_pass:
        lib.vfp_set_tuning(0, args.stem_pass)
    if args.two_kernel_stem:
        lib.vfp_set_tuning(1, 0)
    if args.stem_mode >= 0:
        lib.vfp_set_tuning(1, args.stem_mode)
    for kv in args.tuning:
        k, v = kv.split("=")
        lib.vfp_set_tuning(int(k), int(v))
    if args.conv_pass:
        lib.vfp_set_tuning(3, args.conv_pass)
    peaks = load_peaks()

    n_clips = args.clips
    torch.manual_seed(0)
    model = vfp.create_model("attention").eval()
    model.frames_per_pass = args.frames_per_pass
    lengths = [T_FRAMES] * n_clips
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    frames = torch.empty((n_clips * T_FRAMES, 3, 64, 64), dtype=torch.bfloat16, device=dev)
    chunk = 64 * T_FRAMES
    for s in range(0, frames.shape[0], chunk):  # generate in slices: torch.rand has no bf16 generator path for 15 GB at once
        e = min(frames.shape[0], s + chunk)
        frames[s:e] = torch.rand((e - s, 3, 64, 64), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    torch.cuda.synchronize()

    def step():
        return model.fingerprint_packed(frames, lengths)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        emb = step()
    barrier(world)
    stage_ms = (C.c_double * 32)()
    launches = C.c_uint64(0)
    lib.vfp_profile_read(stage_ms, 32, C.byref(launches), 1)  # reset the launch counter
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(world)
    t_begin = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        emb = step()
    ev1.record()
    barrier(world)
    t_end = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    lib.vfp_profile_read(stage_ms, 32, C.byref(launches), 1)
    gpu_launches = int(launches.value)
    # per-kernel times: one more step with the library's stage profiler on (CUDA events between the stages,
    # on the launching stream); kept out of the timed region because draining events stalls the launch thread
    lib.vfp_profile_enable(1)
    emb = step()
    torch.cuda.synchronize()
    lib.vfp_profile_read(stage_ms, 32, C.byref(launches), 1)
    lib.vfp_profile_enable(0)
    ms = max_over_ranks(ms, world, dev)
    value = world * n_clips * args.steps / (ms / 1000.0)
    n_stage = lib.vfp_profile_num_stages()
    stages = {lib.vfp_profile_stage_name(i).decode(): stage_ms[i] for i in range(n_stage)}

    # ---- end-to-end through the public API with host uint8 frames (pinned), copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        pool_clips = min(n_clips, 1000)
        host = torch.empty((pool_clips * T_FRAMES, 3, 64, 64), dtype=torch.uint8).pin_memory()
        host.random_(0, 256)
        sub = min(pool_clips, 500)  # clips per H2D chunk (two device buffers, copy of chunk i+1 overlaps compute of i)
        dbuf = [torch.empty((sub * T_FRAMES, 3, 64, 64), dtype=torch.uint8, device=dev) for _ in range(2)]
        out_host = torch.empty((n_clips, 256), dtype=torch.float32).pin_memory()
        copy_stream = torch.cuda.Stream(dev)
        main = torch.cuda.current_stream(dev)

        def e2e_step():
            done, i = 0, 0
            ready = [torch.cuda.Event(), torch.cuda.Event()]
            free = [torch.cuda.Event(), torch.cuda.Event()]
            for ev in free:
                ev.record(main)
            while done < n_clips:
                n = min(sub, n_clips - done)
                off = (done % pool_clips)
                if off + n > pool_clips:
                    off = 0
                b = i & 1
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(free[b])
                    dbuf[b][: n * T_FRAMES].copy_(host[off * T_FRAMES : (off + n) * T_FRAMES], non_blocking=True)
                    ready[b].record(copy_stream)
                main.wait_event(ready[b])
                e = model.fingerprint_packed(dbuf[b][: n * T_FRAMES], [T_FRAMES] * n)
                out_host[done : done + n].copy_(e, non_blocking=True)
                free[b].record(main)
                done += n
                i += 1
            main.synchronize()

        e2e_step()
        barrier(world)
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 3))
        for _ in range(e2e_steps):
            e2e_step()
        barrier(world)
        dt = max_over_ranks(time.perf_counter() - t0, world, dev)
        e2e = {
            "value": world * n_clips * e2e_steps / dt, "unit": "videos/s",
            "h2d_bytes_per_step": n_clips * T_FRAMES * 12288, "d2h_bytes_per_step": n_clips * 256 * 4,
            "input": "uint8 frames in pinned host memory, 500-clip chunks double-buffered on a copy stream", "steps": e2e_steps,
        }
        del dbuf, host

    # ---- similarity join (second half of the BASELINE metric) ----
    # N = 1: one GPU joins n x n. N > 1 (BASELINE configs[3] layout): every rank owns a row block of `join_n` rows of
    # the (N * join_n)-row matrix; the timed region holds the NCCL all-gather of the shards AND the row-block join
    # against all columns (global pair indices through q_row0), max over ranks.
    join = None
    if not args.no_join:
        n_local = args.join_n
        n_total = n_local * world
        E_local = make_join_data(n_local, dev, seed=11 + rank)
        if world > 1:
            import torch.distributed as dist

            E_full = torch.empty((n_total, 256), dtype=torch.float32, device=dev)

            def join_step(capacity=None):
                dist.all_gather_into_tensor(E_full, E_local)
                return vfp.threshold_join_device(E_full, 0.95, q=E_local, q_row0=rank * n_local, capacity=capacity)
        else:
            def join_step(capacity=None):
                return vfp.threshold_join_device(E_local, 0.95, capacity=capacity)

        ii, jj, ss = join_step()  # warm-up + sizes
        cap = int(ii.numel()) + 4096
        torch.cuda.synchronize()
        barrier(world)
        j0, j1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        j0.record()
        reps = 3
        for _ in range(reps):
            ii, jj, ss = join_step(cap)
        j1.record()
        torch.cuda.synchronize()
        jms = max_over_ranks(j0.elapsed_time(j1) / reps, world, dev)
        pairs_found = int(max_over_ranks(float(ii.numel()), world, dev)) if world == 1 else None
        if world > 1:
            t = torch.tensor([float(ii.numel())], dtype=torch.float64, device=dev)
            dist.all_reduce(t)
            pairs_found = int(t.item())
        gpairs = (n_total * n_total) / (jms / 1000.0) / 1e9
        tfl = gpairs * 512 / 1000.0 / world
        join = {
            "value": gpairs, "unit": "Gpairs/s", "n": n_total, "rows_per_gpu": n_local, "threshold": 0.95, "pairs_found": pairs_found, "ms": jms,
            "roofline": {"bound": "tensor", "achieved": tfl, "peak": peaks["tc_burst"], "unit": "TFLOP/s per GPU", "frac": tfl / peaks["tc_burst"], "traffic": None},
            "note": ("one GPU joins n x n" if world == 1 else
                     f"row-block sharded: NCCL all-gather of {world} x ({n_local}, 256) fp32 shards + each rank joins its {n_local} rows against all {n_total} columns, both inside the timed region"),
        }
        del E_local

    # ---- the neighbouring steps of SURVEY.md section 8(f), timed briefly (rank 0, device-resident inputs, CUDA events) ----
    extras = None
    if not args.no_extras and rank == 0:
        extras = {}

        def timed(fn, reps=3):
            fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps

        try:
            dec = torch.randint(0, 256, (64, 1080, 1920, 3), dtype=torch.uint8, device=dev)   # one video's decoded 1080p frames
            ms_pre = timed(lambda: vfp.preprocess_frames_device(dec))
            extras["preprocess_1080p"] = {"frames_per_s": 64 / ms_pre * 1e3, "ms_per_64_frames": ms_pre, "hbm_gbs": 64 * 1088 * 1080 * 3 / ms_pre / 1e6,
                                          "what": "INTER_AREA resize to a short side of 64 + centre crop (fingerprint.py:186-214), bit-exact with cv2"}
            del dec
            g3 = torch.Generator(device=dev).manual_seed(5)
            nm = 8192
            Em = torch.randn((nm, 256), generator=g3, device=dev)
            Em = Em / Em.norm(dim=1, keepdim=True)
            ids = np.repeat(np.arange(nm // 2), 2)
            ms_met = timed(lambda: (vfp.compute_retrieval_metrics(Em, ids), vfp.compute_discrimination_metrics(Em, ids)), reps=2)
            extras["trainer_metrics"] = {"embeddings": nm, "ms_both_functions": ms_met, "what": "R@k, mAP, P/R/F1/FPR, AUC-ROC (train.py:285-358, 439-481), host bookkeeping included"}
            m3 = vfp.create_model("3d").eval()
            n3 = min(2048, n_clips)
            x3 = frames[: n3 * T_FRAMES].view(n3, T_FRAMES, 3, 64, 64)
            ms_3d = timed(lambda: m3(x3))
            extras["model_3d"] = {"videos_per_s": n3 / ms_3d * 1e3, "clips": n3, "frames": T_FRAMES, "frame_stride": 16, "what": "VideoFingerprint3D forward (model.py:406-512), bf16 frames in HBM"}
            del m3
        except Exception as exc:  # these are side measurements: never fail the headline line
            extras["error"] = repr(exc)

    if rank != 0:
        return
    # ---- roofline of the dominant kernel ----
    dom = max((k for k in stages if k in STAGE_FLOPS or k in STAGE_BYTES), key=lambda k: stages[k])
    token_passes = -(-n_clips * T_FRAMES // args.frames_per_pass)
    conv_passes = -(-min(n_clips * T_FRAMES, args.frames_per_pass) // 16384) * token_passes
    launches_per_step = conv_passes if dom.startswith("conv") or dom == "stem_fused" else token_passes * (4 if dom.endswith("gemm") or dom in ("attention", "mlp1_gemm_gelu") else 1)
    dom_ms = stages[dom]
    tflops = STAGE_FLOPS.get(dom, 0) * n_clips / (dom_ms / 1000.0) / 1e12
    gbs = STAGE_BYTES.get(dom, 0) * n_clips / (dom_ms / 1000.0) / 1e9
    frac_t, frac_h = tflops / peaks["tc_sustained"], gbs / peaks["hbm"]
    hbm_bound = frac_h > frac_t  # the binding roofline is the one the kernel sits closer to
    # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/r01_v6_stem_ts_full.txt:
    # dram__bytes_read.sum + dram__bytes_write.sum of one stem launch over a 16 384-frame conv pass); below the algorithmic
    # 939.5 MB because the tail of the output is still in L2 when the kernel ends
    ncu_traffic = {"stem_fused": 402.787584e6 + 488.215040e6}
    roofline = {
        "bound": "hbm" if hbm_bound else "tensor", "kernel": dom,
        "achieved": gbs if hbm_bound else tflops, "peak": peaks["hbm"] if hbm_bound else peaks["tc_sustained"],
        "unit": "GB/s" if hbm_bound else "TFLOP/s", "frac": frac_h if hbm_bound else frac_t,
        "traffic": ncu_traffic.get(dom) if (n_clips * T_FRAMES) >= 16384 else None,
        "traffic_note": "bytes per launch (one 16 384-frame conv pass) from the committed ncu capture of this kernel; algorithmic bytes per launch = %.1f MB" % (STAGE_BYTES.get(dom, 0) / T_FRAMES * 16384 / 1e6),
        "peak_source": f"{peaks['source']} ({'HBM copy bandwidth' if hbm_bound else 'sustained bf16, kernel timed inside a long step'})",
        "other_roofline": {"tensor_tflops": tflops, "tensor_frac": frac_t, "hbm_gbs": gbs, "hbm_frac": frac_h},
        "launches_per_step": launches_per_step, "ms_per_step_in_kernel": dom_ms,
        "algorithmic_per_clip": {"flops": STAGE_FLOPS.get(dom), "bytes": STAGE_BYTES.get(dom)},
        "whole_step": {"achieved": value / world * FLOPS_PER_CLIP / 1e12, "frac": value / world * FLOPS_PER_CLIP / 1e12 / peaks["tc_sustained"], "unit": "TFLOP/s"},
        "per_stage": {
            k: {"ms": round(stages[k], 3),
                "tflops": round(STAGE_FLOPS[k] * n_clips / (stages[k] / 1000.0) / 1e12, 1) if k in STAGE_FLOPS and stages[k] > 0 else None,
                "gbs": round(STAGE_BYTES[k] * n_clips / (stages[k] / 1000.0) / 1e9, 1) if k in STAGE_BYTES and stages[k] > 0 else None}
            for k in stages
        },
    }
    line = {
        "metric": METRIC, "value": value, "unit": "videos/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {
            "workload": f"batched fingerprint forward: {n_clips} synthetic clips x {T_FRAMES} frames @ 64x64 per GPU, bf16 frames resident in HBM, random-init weights (BASELINE configs[1])",
            "frames_per_pass": args.frames_per_pass, "l2": f"inputs ({n_clips * T_FRAMES * 24576 / 1e9:.1f} GB) and per-pass activations exceed the 126 MB L2; no flush needed",
            "parallelism": f"clips sharded over {world} GPU(s), no data-path collective",
        },
        "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline, "stage_ms_per_step": stages, "join": join, "extras": extras, "clocks": clocks,
    }
    if not args.no_cpu and world == 1:
        r, n, dt = cpu_reference_rate(args.cpu_seconds, 2048)
        line["cpu_baseline"] = {
            "value": r, "unit": "videos/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} clips x {T_FRAMES} frames in {dt:.1f}s, B=1 forward loop of oracle/forward_oracle.py (restatement of the reference --device cpu path) on host cores",
        }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=N_CLIPS)
    ap.add_argument("--frames-per-pass", type=int, default=1 << 20)
    ap.add_argument("--join-n", type=int, default=262_144)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--stem-pass", type=int, default=0, help="experiment: frames per conv1+conv2 stem pass")
    ap.add_argument("--two-kernel-stem", action="store_true", help="experiment: stand-alone conv1 + conv2 kernels instead of the fused stem")
    ap.add_argument("--stem-mode", type=int, default=-1, help="experiment: 0 two kernels, 1 fused stem with mma.sync conv1, 2 fused stem with TS-mode tcgen05 conv1")
    ap.add_argument("--tuning", action="append", default=[], help="experiment: key=value passed to vfp_set_tuning")
    ap.add_argument("--conv-pass", type=int, default=0, help="experiment: frames per conv pass (<= 16384)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-join", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the brief preprocess / trainer-metrics / 3-D model timings")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    run_ours(args)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
