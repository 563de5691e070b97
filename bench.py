#!/usr/bin/env python
"""Headline benchmark: R(2+1)D (select_backbone('r21d')) SimCLR+DualVar pretraining step,
64 samples/GPU (3 views each, 16x112x112), bf16 convs — BASELINE.json configs[1].

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One "step" = what pretrain.py:400-451 does for one batch: Normalize/transpose ingest, model forward
(two encoder passes: 3B clips + B segment-shuffled clips), four-loss sum, backward, SGD(momentum).
Prints ONE JSON line (rank 0). See DESIGN.md "Measurement" for the definition of every field.
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pretrain samples/sec (1 sample = 3 views; R(2+1)D SimCLR+DualVar 16x112^2)"
# conv MACs per clip-pass for r21d (1,1,1,1) at 16x112^2: fwd 42.724 GFLOP,
# fwd+dgrad+wgrad 126.95 GFLOP; a sample = 4 clip-passes.
GFLOP_PER_SAMPLE = 507.8


def seed_all(s):
    torch.manual_seed(s)
    np.random.seed(s)
    random.seed(s)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"bf16_tflops": d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "hbm_gbs": d.get("hbm_gbs"),
                "source": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"bf16_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# --------------------------------------------------------------------------------------- reference arm
def oracle_step_fn(batch, threads):
    """The CPU port of the reference step (oracle/, fp32, oneDNN) — the reference itself is Python +
    torch and is not shipped to the GPU box; oracle/ restates it and is pinned to it by tests/golden."""
    from oracle import models as OM
    torch.set_num_threads(threads)
    seed_all(0)
    model = OM.SimCLR_TimeSeriesV4("r21d", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc",
                                   SimpleNamespace(shufflerank_theta=0.05)).train()
    opt = torch.optim.SGD([{"params": p} for p in model.parameters()], lr=0.003, weight_decay=1e-4, momentum=0.9)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1, 1)
    frames = torch.rand(batch, 3, 48, 112, 112, generator=torch.Generator().manual_seed(1234))

    def step():
        x = ((frames - mean) / std).view(batch, 3, 3, 16, 112, 112).transpose(1, 2).contiguous()
        ret = model(x)
        loss = sum(v for k, v in ret.items() if "loss" in k)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return float(loss.detach())

    return step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = 2
    step = oracle_step_fn(sample, cores)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = sample / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, sample),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} samples/step (6 input clips, 8 clip-passes) of the same workload, "
                                   f"oracle/ port of the reference step on the host CPU"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, batch):
    return {"workload": "configs[1]: R(2+1)D (select_backbone('r21d'), 14.4M) SimCLR+DualVar mode clip-sr-tc, "
                        "3 views x 16x112x112 per sample, n_series=2, T=0.07, SGD lr 0.003 m 0.9 wd 1e-4",
            "samples_per_gpu": batch, "clips_per_gpu_step": batch * 3, "clip_passes_per_gpu_step": batch * 4,
            "parallelism": f"dp{args.gpus}", "l2": "inputs (462 MB/step/GPU fp32 at 64 samples) exceed the 126 MB L2"}


def fp32_mode_leg(dev, host_frames, steps=3, batch=16):
    """The same step in the fp32 mode (engine.set_precision("fp32"): fp32 activations, convolutions as sums of bf16
    split-plane products): a side figure next to the bf16 headline, 16 samples per step, resident inputs."""
    from dualvar_b200 import engine as E, models as PM
    from dualvar_b200.engine import RawClips
    from dualvar_b200.optim import SGD
    E.set_precision("fp32", 3)
    try:
        seed_all(0)
        m = PM.SimCLR_TimeSeriesV4("r21d", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc",
                                   SimpleNamespace(shufflerank_theta=0.05)).to(dev).train()
        opt = SGD([{"params": p} for p in m.parameters()], lr=0.003, weight_decay=1e-4, momentum=0.9)
        frames = host_frames[:batch].to(dev)

        def st():
            ret = m(RawClips(frames, 3))
            loss = sum(v for k, v in ret.items() if "loss" in k)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            return loss
        for _ in range(2):
            st()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = st()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"value": batch / (ms / 1e3), "unit": "samples/s", "ms_per_step": ms, "samples_per_gpu": batch,
                "steps": steps, "split_planes": 3, "final_loss": float(loss.detach()),
                "dtype": "f32 activations; conv = 6 bf16 plane products, fp32 accumulate",
                "parity": "losses within 1e-4 of the fp32 oracle (tests/test_fp32_mode_gpu.py)"}
    except Exception as e:  # noqa: BLE001 - a side figure must not take the headline line down
        return {"error": f"{type(e).__name__}: {e}"}
    finally:
        E.set_precision("bf16")


# --------------------------------------------------------------------------------------- our arm
def run_ours(args):
    from dualvar_b200 import _lib, models as PM
    from dualvar_b200.engine import RawClips
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if not _lib.load().dv_device_ok():
        raise SystemExit("bench.py: sm_100a kernels need a compute-capability 10.x device")
    B = args.batch
    seed_all(0)
    model = PM.SimCLR_TimeSeriesV4("r21d", 128, 0.07, world > 1, True, 2, 64, 0.07, 0.07, "clip-sr-tc",
                                   SimpleNamespace(shufflerank_theta=0.05))
    if world > 1:
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)     # pretrain.py:244
    model = model.to(dev).train()
    net = model
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank])   # pretrain.py:248
    from dualvar_b200.optim import SGD
    opt = SGD([{"params": p} for p in net.parameters()], lr=0.003, weight_decay=1e-4, momentum=0.9)   # pretrain.py:262-272
    # synthetic decoded+augmented batch as the loader yields it: (B, 3, 3*16, 112, 112) in [0,1], pinned host
    gen = torch.Generator().manual_seed(1234 + rank)
    host = [torch.rand(B, 3, 48, 112, 112, generator=gen).pin_memory() for _ in range(2)]
    h2d_bytes = host[0].numel() * 4
    dev_in = [torch.empty_like(host[0], device=dev) for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    # the e2e legs' source: fp32 loader batches, uint8 crops or decoded uint8 frames (see below)
    bufs = {"host": host, "dev": dev_in, "wrap": lambda fr: RawClips(fr, 3)}

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])
            bufs["dev"][i % 2].copy_(bufs["host"][i % 2], non_blocking=True)
            ready[i % 2].record(copy_stream)

    loss_host = torch.zeros(1).pin_memory()

    def step(i, e2e):
        if e2e:
            torch.cuda.current_stream().wait_event(ready[i % 2])
            prefetch(i + 1)
        frames = bufs["dev"][i % 2]
        ret = model(bufs["wrap"](frames))
        loss = sum(v for k, v in ret.items() if "loss" in k)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        consumed[i % 2].record()
        if e2e:
            loss_host.copy_(loss.detach().view(1), non_blocking=True)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, e2e, first):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(first, first + nsteps):
            step(i, e2e)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms / nsteps

    # resident inputs for the kernel-side number
    dev_in[0].copy_(host[0]); dev_in[1].copy_(host[1])
    for i in range(args.warmup):
        step(i, False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.load().dv_launch_count()
    ms_step = timed(args.steps, False, 0)
    launches = _lib.load().dv_launch_count() - launches0
    # Kernel attribution for the roofline: the same steps again with CUDA events around every conv call on its
    # launching stream. The product path runs the weight gradients on a side stream next to the BatchNorm
    # passes and the two backbone passes on two streams, where per-launch event times overlap other kernels; for this
    # pass everything is folded back onto one
    # stream so that a launch's duration is its own (ms_step_serial is the step time of that pass).
    from dualvar_b200 import engine as _engine
    side_was, pass_was = _engine.WGRAD_SIDE_STREAM, _engine.PASS_STREAMS
    _engine.WGRAD_SIDE_STREAM = False
    _engine.PASS_STREAMS = False
    timer = _lib.KernelTimer(["dv_conv3d_fprop_bf16", "dv_conv3d_dgrad_bf16", "dv_conv3d_dgrad_bnred_bf16", "dv_conv3d_wgrad_bf16",
                              "dv_conv3d_stem_fprop_bf16", "dv_conv3d_stem_wgrad_bf16"])
    _lib.set_timer(timer)
    ms_step_serial = timed(args.steps, False, 0)
    _lib.set_timer(None)
    _engine.WGRAD_SIDE_STREAM, _engine.PASS_STREAMS = side_was, pass_was
    ksum = timer.summary()
    # end-to-end: host buffers, H2D of every step's input inside the timed region (double-buffered on a
    # copy stream), D2H of the loss every step
    for i in range(2):
        consumed[i].record()
    torch.cuda.synchronize()
    prefetch(0)
    ms_e2e = timed(args.steps, True, 0)
    final_loss = float(loss_host[0])
    # the same end-to-end step fed with uint8 frames (what a decoder yields; ToTensor's x/255 moves into the ingest
    # kernel): 4x fewer bytes over PCIe. Reported next to the fp32 figure, which stays the contract's e2e.
    u8 = [(h * 255).round().to(torch.uint8).pin_memory() for h in host]
    bufs["host"], bufs["dev"] = u8, [torch.empty_like(u8[0], device=dev) for _ in range(2)]
    for i in range(2):
        consumed[i].record()
    torch.cuda.synchronize()
    prefetch(0)
    ms_e2e_u8 = timed(args.steps, True, 0)
    # the same step fed with DECODED frames (uint8 HWC 320x240, what a JPEG decoder yields): Scale((128,171)) bicubic +
    # RandomCrop(112) of the reference loader run on the GPU (dualvar_b200/frames.py, bit-exact with Pillow) instead of in
    # 16 PIL worker processes; crop offsets drawn on the host every step as A.RandomCrop does
    ms_e2e_dec, dec_bytes, dec_error = None, 0, None
    if world == 1:
        try:                                   # a side figure: it must not take the headline line down
            from dualvar_b200 import frames as FR
            dec = torch.randint(0, 256, (B, 48, 240, 320, 3), dtype=torch.uint8, generator=gen).pin_memory()
            dec_h = [dec, dec.clone().pin_memory()]
            del u8
            bufs["host"], bufs["dev"] = dec_h, [torch.empty_like(dec, device=dev) for _ in range(2)]
            bufs["wrap"] = lambda fr: FR.stage_clips(fr, FR.draw_crops(B, 3), 3)
            for i in range(2):
                consumed[i].record()
            torch.cuda.synchronize()
            prefetch(0)
            timed(2, True, 0)                     # warm-up of the staging kernels and their tables
            prefetch(0)
            ms_e2e_dec = timed(args.steps, True, 0)
            dec_bytes = dec.numel()
        except Exception as e:  # noqa: BLE001
            ms_e2e_dec, dec_error = None, f"{type(e).__name__}: {e}"
            torch.cuda.synchronize()
        bufs["wrap"] = lambda fr: RawClips(fr, 3)
    bufs["host"], bufs["dev"] = host, dev_in
    sampler.stop_flag = True

    if rank == 0:
        peaks = measured_peaks()
        value = B * world / (ms_step / 1e3)
        e2e_value = B * world / (ms_e2e / 1e3)
        conv_ms = sum(d["ms"] for d in ksum.values())
        calls = {}
        for name, d in ksum.items():
            calls[name] = {"calls_per_step": d["calls"] / args.steps, "ms_per_step": d["ms"] / args.steps,
                           "tflops": d["flops"] / (d["ms"] * 1e-3) / 1e12 if d["ms"] > 0 else 0.0}
        # group the C-ABI calls by the CUDA kernel that serves them
        groups = {"conv_tile_kernel": ["dv_conv3d_fprop_bf16", "dv_conv3d_dgrad_bf16", "dv_conv3d_dgrad_bnred_bf16",
                                       "dv_conv3d_stem_fprop_bf16"],
                  "conv_wgrad_kernel": ["dv_conv3d_wgrad_bf16", "dv_conv3d_stem_wgrad_bf16"]}
        kern = {}
        for kname, members in groups.items():
            ms = sum(ksum[m]["ms"] for m in members if m in ksum)
            fl = sum(ksum[m]["flops"] for m in members if m in ksum)
            n = sum(ksum[m]["calls"] for m in members if m in ksum)
            if ms > 0:
                kern[kname] = {"ms_per_step": ms / args.steps, "calls_per_step": n / args.steps,
                               "tflops": fl / (ms * 1e-3) / 1e12, "avg_launch_us": ms / n * 1e3,
                               "gflop_per_launch": fl / n / 1e9}
        top = max(kern, key=lambda k: kern[k]["ms_per_step"]) if kern else None
        roof = None
        if top:
            ach = kern[top]["tflops"]
            traffic = traffic_detail = None
            tpath = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
            if os.path.exists(tpath):
                traffic_detail = json.load(open(tpath)).get(top)
                if traffic_detail:
                    traffic = traffic_detail.get("avg_dram_bytes_per_launch")   # dram read + write per launch (ncu)
            roof = {"bound": "tensor", "kernel": top, "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": ach / peaks["bf16_tflops"], "traffic": traffic, "traffic_detail": traffic_detail,
                    "peak_source": peaks["source"],
                    "share_of_step": kern[top]["ms_per_step"] / ms_step_serial, "kernels": kern, "calls": calls,
                    "conv_share_of_step": conv_ms / args.steps / ms_step_serial, "ms_step_serial": ms_step_serial,
                    "whole_step_tflops": GFLOP_PER_SAMPLE * B / ms_step,
                    "note": "achieved = algorithmic conv FLOPs (2*positions*Cout*Cin*taps, logical channels) of all "
                            "launches of the kernel / their summed CUDA-event time, taken in a second pass of the same "
                            "steps with the weight-gradient side stream folded back (single stream: ms_step_serial), "
                            "because concurrent kernels make per-launch event times overlap; traffic = avg DRAM bytes "
                            "per launch from ncu (profiles/), null if no capture"}
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, B),
            "clips_per_s": value * 3, "clip_passes_per_s": value * 4,
            "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
            "e2e_uint8_frames": {"value": B * world / (ms_e2e_u8 / 1e3), "unit": "samples/s", "ms_per_step": ms_e2e_u8,
                                 "h2d_bytes_per_step": h2d_bytes // 4, "d2h_bytes_per_step": 4},
            "e2e_decoded_frames": ({"error": dec_error} if dec_error else None) if ms_e2e_dec is None else {
                "value": B * world / (ms_e2e_dec / 1e3), "unit": "samples/s", "ms_per_step": ms_e2e_dec,
                "h2d_bytes_per_step": dec_bytes, "d2h_bytes_per_step": 4,
                "what": "48 decoded uint8 320x240 frames per sample copied from pinned host memory; Scale((128,171)) "
                        "bicubic + RandomCrop(112) + ToTensor + Normalize on the GPU (bit-exact with Pillow)"},
            "gpu_launches": int(launches), "roofline": roof, "clocks": sampler.summary(),
            "final_loss": final_loss,
        }
        if world == 1:
            line["fp32_mode"] = fp32_mode_leg(dev, host[0])
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            cstep = oracle_step_fn(2, cores)
            cstep()
            t0 = time.perf_counter()
            n = 8                       # ~10 s of host work: a bounded sample of the same workload
            for _ in range(n):
                cstep()
            dt = (time.perf_counter() - t0) / n
            line["cpu_baseline"] = {"value": 2 / dt, "unit": "samples/s", "cores": cores, "kind": "port",
                                    "sample": "2 samples/step (8 clip-passes) x 8 timed steps of the same workload, "
                                              "oracle/ port of the reference step, fp32 oneDNN"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="samples per GPU (3 views each)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
