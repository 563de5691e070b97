#!/usr/bin/env python
"""Headline benchmark: R(2+1)D (select_backbone('r21d')) SimCLR+DualVar pretraining step,
64 samples/GPU (3 views each, 16x112x112), bf16 convs — BASELINE.json configs[1].

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One "step" = what pretrain.py:400-451 does for one batch: Normalize/transpose ingest, model forward
(two encoder passes: 3B clips + B segment-shuffled clips), four-loss sum, backward, SGD(momentum).
Prints ONE JSON line (rank 0). See DESIGN.md "Measurement" for the definition of every field.
"""
import argparse
import glob
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
REF_SAMPLE = 4      # samples per CPU step (B = 4, the configs[0] batch)


def reference_modules(cpu):
    """(model package, kind): the UNMODIFIED reference installed under baseline/_ref by oracle/install_reference.py
    (kind "reference"), else the oracle/ restatement of it (kind "port"; pinned to the reference by tests/golden)."""
    from oracle import install_reference as IR
    if IR.available() and IR.verify():
        ref_model, _, _ = IR.import_reference(cpu=cpu)
        return ref_model, "reference"
    from oracle import models as OM
    return OM, "port"


def reference_step_fn(batch, threads):
    """One pretraining step of the reference on the host CPU exactly as pretrain.py:386-451 runs it (Normalize +
    view/transpose, model forward, four-loss sum, backward, SGD with one group per tensor), fp32, oneDNN."""
    torch.set_num_threads(threads)
    M, kind = reference_modules(cpu=True)
    seed_all(0)
    model = M.SimCLR_TimeSeriesV4("r21d", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc",
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

    return step, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = REF_SAMPLE
    step, kind = reference_step_fn(sample, cores)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = sample / dt
    what = ("the unmodified reference modules (baseline/_ref: backbone/, model/, utils/ of lzhangbj/DualVar)" if kind == "reference"
            else "oracle/ port of the reference step (baseline/_ref not installed)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, sample),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": kind,
                         "sample": f"{sample} samples/step (12 input clips, 16 clip-passes) of the same workload on the "
                                   f"host CPU, fp32 oneDNN, {torch.get_num_threads()} threads: {what}"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_torch_gpu(args):
    """Internal leg (--impl torch-gpu, spawned by the N = 1 run): the same step through the reference's own torch modules
    on THIS B200 - cuDNN / cuBLAS, no dualvar_b200 code - in fp32 (TF32 off) and under bf16 autocast with channels_last_3d
    + cudnn.benchmark. SURVEY.md 8(d) calls this "the honest bar on the same box". Prints one JSON object."""
    dev = torch.device("cuda", 0)
    M, kind = reference_modules(cpu=False)
    B = args.batch
    out = {"kind": kind, "samples_per_step": B, "unit": "samples/s",
           "what": "reference torch modules on the same B200 through cuDNN (no dualvar_b200 kernels), full step "
                   "(Normalize, forward, 4 losses, backward, SGD), inputs resident"}
    mean = torch.tensor([0.485, 0.456, 0.406], device=dev).view(1, 3, 1, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device=dev).view(1, 3, 1, 1, 1)
    frames = torch.rand(B, 3, 48, 112, 112, device=dev, generator=torch.Generator(device=dev).manual_seed(1234))
    for tag in ("bf16_autocast_channels_last", "fp32_tf32_off"):
        try:
            fp32 = tag.startswith("fp32")
            torch.backends.cudnn.allow_tf32 = not fp32
            torch.backends.cuda.matmul.allow_tf32 = not fp32
            torch.backends.cudnn.benchmark = not fp32
            seed_all(0)
            model = M.SimCLR_TimeSeriesV4("r21d", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc",
                                          SimpleNamespace(shufflerank_theta=0.05)).to(dev).train()
            if not fp32:
                model = model.to(memory_format=torch.channels_last_3d)
            opt = torch.optim.SGD([{"params": p} for p in model.parameters()], lr=0.003, weight_decay=1e-4, momentum=0.9)

            def step():
                x = ((frames - mean) / std).view(B, 3, 3, 16, 112, 112).transpose(1, 2).contiguous()
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=not fp32):
                    ret = model(x)
                loss = sum(v.float() for k, v in ret.items() if "loss" in k)
                opt.zero_grad(set_to_none=True)
                loss.backward()
                opt.step()
                return loss
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            n = 3 if fp32 else 6
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                loss = step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            out[tag] = {"value": B / (ms / 1e3), "ms_per_step": ms, "steps": n, "final_loss": float(loss.detach())}
            del model, opt
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            out[tag] = {"error": f"{type(e).__name__}: {e}"}
    print("TORCH_GPU_JSON " + json.dumps(out), flush=True)


def _spawn_leg(impl, extra, marker, timeout):
    """Run another leg of this script in a fresh process (the reference needs ``Tensor.cuda`` patched on the CPU, and its
    top-level package names - model, utils - must not leak into this process); returns its JSON object or an error dict."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", impl] + extra
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
        for ln in reversed(r.stdout.splitlines()):
            if marker is None and ln.startswith("{"):
                return json.loads(ln)
            if marker is not None and ln.startswith(marker):
                return json.loads(ln[len(marker):])
        return {"error": f"leg {impl}: no JSON (rc {r.returncode}): {(r.stderr or r.stdout)[-300:]}"}
    except Exception as e:  # noqa: BLE001
        return {"error": f"leg {impl}: {type(e).__name__}: {e}"}


# the newest per-layer ncu capture of the round (profiles/r02*_ncu_traffic.json, written by tests/diag/ncu_traffic.py)
NCU_TRAFFIC = (sorted(glob.glob(os.path.join(ROOT, "profiles", "r0*_ncu_traffic.json"))) or
               [os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")])[-1]


def ncu_traffic():
    """Per-layer / per-kernel DRAM bytes from this round's ncu capture (tests/diag/ncu_step.py), or None."""
    if os.path.exists(NCU_TRAFFIC):
        try:
            return json.load(open(NCU_TRAFFIC))
        except Exception:  # noqa: BLE001
            return None
    return None


def layer_table(lsum, steps, peaks):
    """Per layer (entry point + geometry): launches and time per step, algorithmic FLOPs and bytes per launch, achieved
    TFLOP/s and GB/s on the algorithmic figures, and - when this round's ncu capture has the layer - the DRAM bytes per
    launch next to them (dram / algorithmic > 1 = re-reads)."""
    tr = ncu_traffic() or {}
    per_layer = tr.get("layers", {})
    rows = []
    for key, d in lsum.items():
        n = d["calls"]
        if n == 0 or d["ms"] <= 0:
            continue
        ms = d["ms"] / n
        row = {"layer": key, "launches_per_step": n / steps, "ms_per_launch": ms, "ms_per_step": d["ms"] / steps}
        if d["flops"] > 0:
            row["gflop_per_launch"] = d["flops"] / n / 1e9
            row["tflops"] = d["flops"] / n / (ms * 1e-3) / 1e12
            row["frac_of_bf16_peak"] = row["tflops"] / peaks["bf16_tflops"]
        if d.get("bytes", 0) > 0:
            row["algorithmic_bytes_per_launch"] = d["bytes"] / n
            row["algorithmic_gbs"] = d["bytes"] / n / (ms * 1e-3) / 1e9
            row["frac_of_hbm_peak"] = row["algorithmic_gbs"] / peaks["hbm_gbs"]
        t = per_layer.get(key)
        if t:      # ncu captured ONE step: DRAM bytes of all kernel launches behind this layer's calls / calls per step
            row["dram_bytes_per_launch"] = t["dram_bytes_per_step"] / (n / steps)
            row["ncu_kernel_launches_per_step"] = t["launches"]
            if row.get("algorithmic_bytes_per_launch"):
                row["dram_over_algorithmic"] = row["dram_bytes_per_launch"] / row["algorithmic_bytes_per_launch"]
        # the layer's own roofline: the slower of its tensor time and its HBM time on the algorithmic figures
        t_roof = max(d["flops"] / n / (peaks["bf16_tflops"] * 1e12), d.get("bytes", 0) / n / (peaks["hbm_gbs"] * 1e9)) * 1e3
        if t_roof > 0:
            row["roofline_ms_per_launch"] = t_roof
            row["frac_of_layer_roofline"] = t_roof / ms
            row["bound"] = "tensor" if d["flops"] / n / (peaks["bf16_tflops"] * 1e12) * 1e3 >= t_roof else "hbm"
        rows.append(row)
    rows.sort(key=lambda r: -r["ms_per_step"])
    return rows


def roofline_summary(rows, prefixes):
    """Time-weighted fraction of the per-layer roofline over the layers whose entry point starts with one of prefixes."""
    sel = [r for r in rows if r["layer"].split(" ")[0].startswith(prefixes) and "roofline_ms_per_launch" in r]
    t = sum(r["ms_per_step"] for r in sel)
    roof = sum(r["roofline_ms_per_launch"] * r["launches_per_step"] for r in sel)
    return {"ms_per_step": t, "roofline_ms_per_step": roof, "frac": roof / t if t else None,
            "hbm_bound_ms_per_step": sum(r["ms_per_step"] for r in sel if r.get("bound") == "hbm"),
            "tensor_bound_ms_per_step": sum(r["ms_per_step"] for r in sel if r.get("bound") == "tensor")}


def workload_config(args, batch):
    return {"workload": "configs[1]: R(2+1)D (select_backbone('r21d'), 14.4M) SimCLR+DualVar mode clip-sr-tc, "
                        "3 views x 16x112x112 per sample, n_series=2, T=0.07, SGD lr 0.003 m 0.9 wd 1e-4",
            "samples_per_gpu": batch, "clips_per_gpu_step": batch * 3, "clip_passes_per_gpu_step": batch * 4,
            "parallelism": f"dp{args.gpus}" + (f" ({args.dp} gradient all-reduce)" if args.gpus > 1 and hasattr(args, "dp") else ""), "l2": "inputs (462 MB/step/GPU fp32 at 64 samples) exceed the 126 MB L2"}


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
            opt.zero_grad(set_to_none=False)
            loss.backward()
            opt.step()
            return loss.detach()      # (detached: see time_step.eager in other_configs_leg)

        def timed(fn):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                out = fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / steps, out
        for _ in range(2):
            st()
        ms, loss = timed(st)
        res = {"value": batch / (ms / 1e3), "unit": "samples/s", "ms_per_step": ms, "mode": "eager", "samples_per_gpu": batch,
               "steps": steps, "split_planes": 3, "final_loss": float(loss),
               "dtype": "f32 activations; conv = 6 bf16 plane products in one launch, fp32 accumulate",
               "parity": "losses within 1e-4 of the fp32 oracle (tests/test_fp32_mode_gpu.py)"}
        try:      # the same step replayed as one CUDA graph; the eager figure stands if the capture is refused
            from dualvar_b200.graph_step import GraphedTrainStep
            gs = GraphedTrainStep(m, opt, n_views=3, warmup=1)
            if gs.enabled:
                for _ in range(2):
                    gs(frames)
                ms_g, out = timed(lambda: gs(frames))
                res["eager"] = {"value": res["value"], "ms_per_step": ms}
                res.update(value=batch / (ms_g / 1e3), ms_per_step=ms_g, mode="cuda_graph", final_loss=float(out["loss"]))
                gs.release()
        except Exception as e:  # noqa: BLE001
            res["graph_error"] = f"{type(e).__name__}: {e}"[:300]
        return res
    except Exception as e:  # noqa: BLE001 - a side figure must not take the headline line down
        return {"error": f"{type(e).__name__}: {e}"}
    finally:
        E.set_precision("bf16")


def other_configs_leg(dev):
    """Side figures for the other BASELINE.json configurations (each parity-tested in tests/test_parity_configs_gpu.py and
    tests/test_retrieval_gpu.py): full pretraining steps with resident inputs, and the retrieval maths."""
    from dualvar_b200 import models as PM, retrieval as R
    from dualvar_b200.engine import RawClips
    from dualvar_b200.optim import SGD
    out = {}
    a = SimpleNamespace(shufflerank_theta=0.05)

    def time_step(model, frames, batch, n=4):
        """ms per full step, replayed as one CUDA graph (graph_step.GraphedTrainStep) and issued eagerly."""
        from dualvar_b200 import _lib
        from dualvar_b200.graph_step import GraphedTrainStep
        opt = SGD([{"params": p} for p in model.parameters() if p.requires_grad], lr=0.003, weight_decay=1e-4, momentum=0.9)

        def eager():
            ret = model(RawClips(frames, 3))
            loss = sum(v for k, v in ret.items() if "loss" in k)
            opt.zero_grad(set_to_none=False)
            loss.backward()
            opt.step()
            # detached: a kept loss WITH autograd history keeps the parameters' AccumulateGrad nodes (bound to this
            # stream) alive, and autograd then synchronises the graph-capture stream below with this one - which
            # invalidates the capture
            return loss.detach()

        def timed(fn, k):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(k):
                out = fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / k, out
        for _ in range(2):
            eager()
        nl0 = _lib.load().dv_launch_count()
        ms_eager, loss = timed(eager, n)
        launches = (_lib.load().dv_launch_count() - nl0) / n
        res = {"unit": "samples/s", "samples_per_gpu": batch, "steps": n, "gpu_launches_per_step": launches,
               "eager": {"value": batch / (ms_eager / 1e3), "ms_per_step": ms_eager}}
        gs = GraphedTrainStep(model, opt, n_views=3, warmup=1)
        if gs.enabled:
            for _ in range(3):
                gs(frames)
            ms_graph, out = timed(lambda: gs(frames), n)
            loss = out["loss"]
            res.update(value=batch / (ms_graph / 1e3), ms_per_step=ms_graph, mode="cuda_graph")
            gs.release()
        else:
            res.update(value=res["eager"]["value"], ms_per_step=ms_eager, mode="eager")
        res["final_loss"] = float(loss.detach())
        return res

    try:
        seed_all(0)
        B = 64
        m = PM.MoCo_TimeSeriesV4("r21d", 128, 16384, 0.999, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", a).to(dev).train()
        frames = torch.rand(B, 3, 48, 112, 112, device=dev)
        out["configs[2] MoCo+DualVar r21d K=16384 m=0.999 16x112^2"] = time_step(m, frames, B)
        del m, frames
    except Exception as e:  # noqa: BLE001
        out["configs[2]"] = {"error": f"{type(e).__name__}: {e}", "reserved_gb": torch.cuda.memory_reserved() / 1e9}
    torch.cuda.empty_cache()
    try:
        seed_all(0)
        B = 16
        m = PM.SimCLR_TimeSeriesV4("s3dg", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", a).to(dev).train()
        frames = torch.rand(B, 3, 96, 128, 128, device=dev)
        out["configs[3] S3D-G SimCLR+DualVar 32x128^2"] = time_step(m, frames, B, n=3)
        del m, frames
    except Exception as e:  # noqa: BLE001
        out["configs[3]"] = {"error": f"{type(e).__name__}: {e}"}
    torch.cuda.empty_cache()
    try:
        test = torch.randn(3783, 512, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
        train = torch.randn(9537, 512, device=dev, generator=torch.Generator(device=dev).manual_seed(8))
        for _ in range(2):
            R.retrieval_topk(test, train)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            R.retrieval_topk(test, train)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out["configs[4] retrieval 3783 x 9537 x 512 (centre, normalise, fp64 similarity, top-1/5/10/20/50)"] = {
            "value": 3783 / (ms / 1e3), "unit": "queries/s", "ms": ms}
    except Exception as e:  # noqa: BLE001
        out["configs[4]"] = {"error": f"{type(e).__name__}: {e}"}
    return out


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
        # pretrain.py:248. Default: dualvar_b200.parallel.DataParallel - DDP's contract with the gradient all-reduce
        # bucketed INSIDE the engine's backward (overlapped); --dp torch uses torch's DistributedDataParallel, whose
        # reducer only sees the backbone gradients when the one-node backbone backward has finished
        if args.dp == "overlap":
            from dualvar_b200.parallel import DataParallel
            model = DataParallel(model, device_ids=[local_rank])
        else:
            model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank])
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
    # One CUDA graph per step (dualvar_b200/graph_step.py) on a single GPU: the whole step - ingest, both encoder passes,
    # heads, the four losses, backward, fused SGD - is captured once and replayed; the batch is copied into the graph's
    # static input buffer and the host RNG draws of the step (segment permutations) are refreshed before every replay.
    # Multi-GPU runs capture per rank with the collectives inside (see graph_step.py); under --dp torch they stay eager.
    # --no-graph / DV_BENCH_GRAPH=0 times the eager step.
    from dualvar_b200.graph_step import GraphedTrainStep
    use_graph = (world == 1 or args.dp == "overlap") and not args.no_graph and os.environ.get("DV_BENCH_GRAPH", "1") != "0"
    graphed = {"step": None}

    def new_graphed():
        if graphed["step"] is not None:
            graphed["step"].release()
        graphed["step"] = GraphedTrainStep(model, opt, n_views=3, warmup=0, wrap=bufs["wrap"]) if use_graph else None

    def step(i, e2e, eager=False):
        if e2e:
            torch.cuda.current_stream().wait_event(ready[i % 2])
            prefetch(i + 1)
        frames = bufs["dev"][i % 2]
        if graphed["step"] is not None and not eager:
            loss = graphed["step"](frames)["loss"]      # D2D copy into the static input buffer + graph replay
        else:
            ret = model(bufs["wrap"](frames))
            loss = sum(v for k, v in ret.items() if "loss" in k)
            opt.zero_grad(set_to_none=not use_graph)
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

    def timed(nsteps, e2e, first, eager=False):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(first, first + nsteps):
            step(i, e2e, eager)
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
        step(i, False, eager=True)
    ms_step_eager = timed(args.steps, False, 0, eager=True) if use_graph else None     # the same step issued call by call
    new_graphed()
    if graphed["step"] is not None:
        for i in range(3):                       # one eager step on the capture stream, the capture, a first replay
            step(i, False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.load().dv_launch_count()
    ms_step = timed(args.steps, False, 0)
    launches = _lib.load().dv_launch_count() - launches0
    if graphed["step"] is not None:              # replays issue no C-ABI calls: the graph holds the step's launches
        launches = graphed["step"].launches_per_step * args.steps
    if ms_step_eager is None:
        ms_step_eager = ms_step
    # Kernel attribution for the roofline: the same steps again with CUDA events around every conv call on its
    # launching stream. The product path runs the weight gradients on a side stream next to the BatchNorm
    # passes and the two backbone passes on two streams, where per-launch event times overlap other kernels; for this
    # pass everything is folded back onto one
    # stream so that a launch's duration is its own (ms_step_serial is the step time of that pass).
    from dualvar_b200 import engine as _engine
    side_was, pass_was = _engine.WGRAD_SIDE_STREAM, _engine.PASS_STREAMS
    _engine.WGRAD_SIDE_STREAM = False
    _engine.PASS_STREAMS = False
    conv_names = ["dv_conv3d_fprop_bf16", "dv_conv3d_dgrad_bf16", "dv_conv3d_dgrad_bnred_bf16", "dv_conv3d_wgrad_bf16",
                  "dv_conv3d_stem_fprop_bf16", "dv_conv3d_stem_wgrad_bf16", "dv_conv3d_fprop_bnrelu_bf16",
                  "dv_conv3d_wgrad_bnrelu_bf16"]
    bn_names = ["dv_bn_apply", "dv_bn_bwd_reduce", "dv_bn_bwd_apply"]
    timer = _lib.KernelTimer(conv_names + bn_names, detail=True)
    _lib.set_timer(timer)
    ms_step_serial = timed(args.steps, False, 0, eager=True)
    _lib.set_timer(None)
    _engine.WGRAD_SIDE_STREAM, _engine.PASS_STREAMS = side_was, pass_was
    lsum = timer.summary()                      # per layer: "<entry point> <geometry>" -> calls, ms, flops, bytes
    ksum = {}
    for key, d in lsum.items():
        name = key.split(" ")[0]
        if name in conv_names:
            k = ksum.setdefault(name, {"calls": 0, "ms": 0.0, "flops": 0.0})
            for f in ("calls", "ms", "flops"):
                k[f] += d[f]
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
    new_graphed()                                # another input dtype: its own graph
    if graphed["step"] is not None:
        bufs["dev"][0].copy_(u8[0])
        for i in range(3):
            step(0, False)
    for i in range(2):
        consumed[i].record()
    torch.cuda.synchronize()
    prefetch(0)
    ms_e2e_u8 = timed(args.steps, True, 0)
    if graphed["step"] is not None:
        graphed["step"].release()
        graphed["step"] = None                   # the remaining legs are eager (host-drawn crop offsets per step)
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
                                       "dv_conv3d_stem_fprop_bf16", "dv_conv3d_fprop_bnrelu_bf16"],
                  "conv_wgrad_kernel": ["dv_conv3d_wgrad_bf16", "dv_conv3d_stem_wgrad_bf16", "dv_conv3d_wgrad_bnrelu_bf16"]}
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
        layers = layer_table(lsum, args.steps, peaks)
        roof = None
        if top:
            ach = kern[top]["tflops"]
            # DRAM bytes per launch of the dominant kernel from the ncu capture of this round (same launches, taken
            # with tests/diag/ncu_step.py; per-layer figures sit in roofline.layers next to the algorithmic bytes)
            traffic = traffic_detail = None
            tr = ncu_traffic()
            if tr is not None:
                traffic_detail = tr.get("kernels", {}).get(top)
                if traffic_detail:
                    traffic = traffic_detail.get("avg_dram_bytes_per_launch")
            roof = {"bound": "tensor", "kernel": top, "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": ach / peaks["bf16_tflops"], "traffic": traffic, "traffic_detail": traffic_detail,
                    "peak_source": peaks["source"],
                    "share_of_step": kern[top]["ms_per_step"] / ms_step, "kernels": kern, "calls": calls,
                    "conv_share_of_step": conv_ms / args.steps / ms_step, "ms_step_serial": ms_step_serial,
                    "layers": layers,
                    "per_layer_roofline": {
                        "what": "each layer against ITS OWN roofline = max(algorithmic FLOPs / bf16 peak, algorithmic "
                                "bytes / HBM peak): several conv layers of this network are HBM-bound even as convolutions "
                                "(K = 192 temporal convs on 144-channel maps), so a pure tensor fraction understates them",
                        "conv": roofline_summary(layers, ("dv_conv3d",)),
                        "batchnorm": roofline_summary(layers, ("dv_bn_",))},
                    "whole_step_tflops": GFLOP_PER_SAMPLE * B / ms_step,
                    "note": "achieved = algorithmic conv FLOPs (2*positions*Cout*Cin*taps, logical channels) of all "
                            "launches of the kernel / their summed CUDA-event time, taken in a second pass of the same "
                            "steps with the weight-gradient side stream folded back (single stream: ms_step_serial), "
                            "because concurrent kernels make per-launch event times overlap; share_of_step = that device "
                            "time / ms_per_step of the headline (graph-replayed) step; traffic = avg DRAM bytes "
                            "per launch from ncu (profiles/), null if no capture"}
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, B),
            "clips_per_s": value * 3, "clip_passes_per_s": value * 4,
            # the public call: model(RawClips(frames, 3)) on the loader's uint8 crops (B, 3, 48, 112, 112), copied from
            # pinned host memory every step (double-buffered on a copy stream), loss read back every step
            "e2e": {"value": B * world / (ms_e2e_u8 / 1e3), "unit": "samples/s", "ms_per_step": ms_e2e_u8,
                    "h2d_bytes_per_step": h2d_bytes // 4, "d2h_bytes_per_step": 4,
                    "input": "uint8 frames as decoded (ToTensor's x/255 and Normalize run in the ingest kernel)"},
            "e2e_fp32_frames": {"value": e2e_value, "unit": "samples/s", "ms_per_step": ms_e2e,
                                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                                "input": "fp32 frames as the reference's CPU ToTensor yields them (4x the bytes)"},
            "e2e_decoded_frames": ({"error": dec_error} if dec_error else None) if ms_e2e_dec is None else {
                "value": B * world / (ms_e2e_dec / 1e3), "unit": "samples/s", "ms_per_step": ms_e2e_dec,
                "h2d_bytes_per_step": dec_bytes, "d2h_bytes_per_step": 4,
                "what": "48 decoded uint8 320x240 frames per sample copied from pinned host memory; Scale((128,171)) "
                        "bicubic + RandomCrop(112) + ToTensor + Normalize on the GPU (bit-exact with Pillow)"},
            "gpu_launches": int(launches), "roofline": roof, "clocks": sampler.summary(),
            "step_issue": {"mode": "cuda_graph" if use_graph else "eager", "ms_per_step_eager": ms_step_eager,
                           "what": "value / e2e replay ONE CUDA graph per step (dualvar_b200/graph_step.py); "
                                   "ms_per_step_eager is the same step issued call by call from Python"},
            "final_loss": final_loss,
        }
        if world == 1:
            line["fp32_mode"] = fp32_mode_leg(dev, host[0])
        if world == 1 and not args.no_side_legs:
            del model, net, opt
            torch.cuda.empty_cache()
            line["other_configs"] = other_configs_leg(dev)
            torch.cuda.empty_cache()
            # the honest same-box bar: the reference's torch modules on this B200 through cuDNN (fresh process)
            line["gpu_baseline"] = _spawn_leg("torch-gpu", ["--batch", "32"], "TORCH_GPU_JSON ", 900)
        if world == 1 and not args.no_cpu_baseline:
            # the reference step on this box's host cores: a bounded sample (1 warm-up + 2 timed steps of 4 samples)
            ref = _spawn_leg("reference", ["--steps", "2", "--warmup", "1"], None, 900)
            line["cpu_baseline"] = ref.get("cpu_baseline", ref)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-gpu"])
    ap.add_argument("--batch", type=int, default=64, help="samples per GPU (3 views each)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dp", default=os.environ.get("DV_BENCH_DP", "overlap"), choices=["overlap", "torch"],
                    help="data-parallel wrapper for --gpus > 1")
    ap.add_argument("--no-graph", action="store_true", help="time the eager step instead of the CUDA-graph replay")
    ap.add_argument("--no-side-legs", action="store_true", help="skip the other_configs / gpu_baseline side figures")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch-gpu":
        run_torch_gpu(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
