"""GPU parity on the BASELINE.json configurations themselves and tensor-by-tensor gradient bounds.

* configs[2]: MoCo+DualVar with the 16384-entry queue and m = 0.999 (model/moco.py:482-573): losses, clip_/tc_ logits,
  queue / series_queue / queue_ptr and the momentum-updated key encoder after the enqueue.
* configs[3]: S3D-G SimCLR+DualVar at 32x128x128 in the bf16 mode: the four losses against the fp32 oracle.
* every parameter gradient, tensor by tensor, against the oracle run with the product's bf16 rounding points
  (tests/bf16_emulation.py) - the median-only comparison against torch autocast lets a wrong minority through.
* the packed-weight cache across model lifetimes (ids and addresses are recycled).

Measured figures are recorded through the ``measured`` fixture; bounds below are those figures with head-room.
"""
import copy
import gc
import os
import random
import subprocess
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
dev = "cuda:0"
ARGS = SimpleNamespace(shufflerank_theta=0.05)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


S3DG_LOSS_TOL = 0.12      # measured 8.2e-2 (clip), 7.8e-3 (tc), 2.3e-2 / 4.1e-3 (rank): profiles/r02_measured_parity.jsonl


def _seed(s):
    torch.manual_seed(s); np.random.seed(s); random.seed(s)


def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def _losses(ret):
    return {k: v for k, v in ret.items() if "loss" in k}


def _simclr_trio(net):
    """(fp32 oracle, the same oracle with the product's bf16 rounding points, product) on identical weights
    (backbone conv weights bf16-representable)."""
    from bf16_emulation import emulate_bf16, round_conv_weights
    from dualvar_b200 import models as PM
    from oracle import models as OM
    _no_tf32()
    _seed(0)
    ref = OM.SimCLR_TimeSeriesV4(net, 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", ARGS).to(dev).train()
    round_conv_weights(ref)
    emu = emulate_bf16(ref)
    prod = PM.SimCLR_TimeSeriesV4(net, 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", ARGS)
    prod.load_state_dict(ref.state_dict())
    return ref, emu, prod.to(dev).train()


# Gradient bounds, tensor by tensor. err(t) = ||g - g_fp32|| / ||g_fp32|| of tensor t against the fp32 oracle.
#   * product within RATIO x the error the same rounding points cause in the oracle itself (+ FLOOR), where a tensor's
#     own noise floor is taken no lower than the median one (the first layers of the heads see 12-24 rows: their error
#     swings between 0.14 and 0.36 from run to run on EITHER side, tests/diag/head_grad_probe.py): a wrong or missing
#     term in ONE tensor (say a downsample branch) reads ~1.0 against a floor of 0.1-0.4 and fails;
#   * the medians within MED_RATIO (the old assertion, kept as the second one).
# Measured on B200 (profiles/r02_measured_parity.jsonl): worst ratio 1.18 (r21d), 1.2 (r3d), medians within 1.05.
RATIO, FLOOR, MED_RATIO = 1.6, 0.03, 1.25


def _check_per_tensor(ref, emu, prod, measured, tag, ratio=RATIO, floor=FLOOR, med_ratio=MED_RATIO):
    from bf16_emulation import compare_to_noise_floor
    rows = compare_to_noise_floor(ref, emu, prod)
    assert len(rows) >= 10
    med_p = sorted(r[1] for r in rows)[len(rows) // 2]
    med_e = sorted(r[2] for r in rows)[len(rows) // 2]
    rows = [(n, p, max(e, med_e)) for n, p, e in rows]
    worst = max(rows, key=lambda t: t[1] / (t[2] + floor))
    measured(f"{tag}.grad_worst_tensor", {"name": worst[0], "err_product": worst[1], "err_rounding_oracle": worst[2]})
    measured(f"{tag}.grad_median", {"err_product": med_p, "err_rounding_oracle": med_e})
    bad = [(n, round(p, 4), round(e, 4)) for n, p, e in rows if p > ratio * e + floor]
    assert not bad, f"{tag}: gradient error above {ratio} x its rounding-noise floor + {floor}: {bad[:8]}"
    assert med_p <= med_ratio * med_e + 0.01, (tag, med_p, med_e)


@pytest.mark.parametrize("net,shape", [("r21d", (8, 3, 3, 8, 64, 64)), ("r3d", (8, 3, 3, 8, 64, 64)),
                                       ("r21d", (4, 3, 3, 16, 112, 112))])
def test_simclr_dualvar_gradients_tensor_by_tensor(net, shape, measured):
    """SimCLR+DualVar step: every loss against the oracle run with the product's rounding points (tight: the two differ
    by summation order only) and EVERY parameter gradient against its own bf16 noise floor; the (4, .., 16, 112, 112)
    case takes the bench geometry's launches (CTA pairs, two-region tiling, halo wgrad, fused dgrad + BN reduce)."""
    from bf16_emulation import round_input
    ref, emu, prod = _simclr_trio(net)
    x = round_input(torch.randn(*shape, device=dev))
    np.random.seed(11); rr = ref(x)
    np.random.seed(11); re_ = emu(x)
    np.random.seed(11); rp = prod(x)
    tag = f"simclr_{net}_{shape[3]}x{shape[4]}"
    for k, v in _losses(re_).items():
        err = abs(rp[k].item() - v.item()) / abs(v.item())
        measured(f"{tag}.{k}.vs_rounding_oracle", err)
        assert err <= 5e-3, (k, rp[k].item(), v.item())                       # measured <= 1.5e-3
        assert abs(rp[k].item() - rr[k].item()) <= 1e-2 * abs(rr[k].item()), (k, rp[k].item(), rr[k].item())
    for r_ in (rr, re_, rp):
        sum(_losses(r_).values()).backward()
    _check_per_tensor(ref, emu, prod, measured, tag)
    for (n, br), (_, bp) in zip(emu.named_buffers(), prod.named_buffers()):
        if br.dtype.is_floating_point:
            assert _rel(bp, br) < 5e-3, n


@pytest.mark.parametrize("name,shape", [("c3d", (8, 3, 8, 64, 64)), ("r2d3d18", (6, 3, 4, 96, 96)),
                                        ("s3d", (4, 3, 32, 128, 128)), ("s3dg", (4, 3, 32, 128, 128))])
def test_backbone_gradients_tensor_by_tensor(name, shape, measured):
    """Backbone alone, random upstream gradient: output against the rounding oracle, every parameter gradient against
    its own noise floor. S3D / S3D-G at the BASELINE clip size (32 x 128 x 128 -> 4 x 4 x 4 final map)."""
    from bf16_emulation import emulate_bf16, round_conv_weights, round_input
    from dualvar_b200 import backbones as PB
    from oracle import backbones as OB
    _no_tf32()
    _seed(0)
    ref, _ = OB.select_backbone(name)
    ref = round_conv_weights(ref.to(dev).train())
    emu = emulate_bf16(ref)
    prod, _ = PB.select_backbone(name)
    prod.load_state_dict(ref.state_dict())
    prod = prod.to(dev).train()
    x = round_input(torch.randn(*shape, device=dev))
    yr, ye, yp = ref(x), emu(x), prod(x)
    measured(f"backbone_{name}.out", {"product_vs_rounding_oracle": _rel(yp, ye), "product_vs_fp32": _rel(yp, yr),
                                      "rounding_oracle_vs_fp32": _rel(ye, yr)})
    g = torch.randn_like(yr)
    yr.backward(g); ye.backward(g); yp.backward(g)
    assert _rel(yp, yr) <= 1.5 * _rel(ye, yr) + 1e-2
    _check_per_tensor(ref, emu, prod, measured, f"backbone_{name}")


def test_moco_dualvar_queue_16384_matches_oracle(measured):
    """BASELINE configs[2] as stated: K = 16384, m = 0.999, r21d, 16x112x112 clips, 16 samples. Two steps: logits against
    the full queue, enqueue at ptr 0 then 16, queue_ptr, series_queue, momentum-updated key encoder."""
    from dualvar_b200 import models as PM
    from oracle import models as OM
    _no_tf32()
    _seed(0)
    K, NB = 16384, 16
    ref = OM.MoCo_TimeSeriesV4("r21d", 128, K, 0.999, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", ARGS).to(dev).train()
    prod = PM.MoCo_TimeSeriesV4("r21d", 128, K, 0.999, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", ARGS)
    prod.load_state_dict(ref.state_dict())
    prod = prod.to(dev).train()
    with torch.no_grad():          # query and key encoders differ, so the momentum update moves the keys
        for m in (ref, prod):
            torch.manual_seed(5)
            for p in m.encoder_q.parameters():
                p.add_(0.01 * torch.randn_like(p))
    q0, sq0 = ref.queue.clone(), ref.series_queue.clone()
    k0 = [p.detach().clone() for p in ref.encoder_k.parameters()]
    for step in range(2):
        x = torch.randn(NB, 3, 3, 16, 112, 112, device=dev, generator=torch.Generator(device=dev).manual_seed(40 + step))
        np.random.seed(20 + step); rr = ref(x)
        np.random.seed(20 + step); rp = prod(x)
        assert list(rr.keys()) == list(rp.keys())
        assert rp["clip_logits"].shape == rr["clip_logits"].shape == (NB, K + 1)
        assert rp["tc_logits"].shape == rr["tc_logits"].shape == (NB, K + 1)
        for k in rr:
            if "labels" in k:
                assert torch.equal(rr[k], rp[k]), k
            elif "loss" in k:
                err = abs(rp[k].item() - rr[k].item()) / abs(rr[k].item())
                measured(f"moco16384.step{step}.{k}", {"err": err, "oracle": rr[k].item()})
                # 1e-2 relative; the absolute term covers the clip / tc losses, which are a small log(1 + e^(neg - pos))
                # whose value moves by |d pos| * ~0.1 (same floor as the K = 64 test)
                assert abs(rp[k].item() - rr[k].item()) <= 1e-2 * abs(rr[k].item()) + 2e-3, (step, k, rp[k].item(), rr[k].item())
            else:
                # logits are cosines / T: tolerance 1e-2 of the range 1/T; margin logits are raw cosines (range 1)
                err = (rp[k].float() - rr[k].float()).abs().max().item()
                rng = 1.0 if "margin" in k else 1.0 / 0.07
                measured(f"moco16384.step{step}.{k}", err / rng)
                assert err <= 1e-2 * rng, (step, k, err)
        n_new = NB * (step + 1)
        assert int(prod.queue_ptr) == int(ref.queue_ptr) == n_new
        # enqueued keys are unit vectors from the bf16 key encoder: compare as vectors; untouched columns bit-identical
        assert _rel(prod.queue[:, :n_new], ref.queue[:, :n_new]) < 3e-2
        assert _rel(prod.series_queue[:, :n_new], ref.series_queue[:, :n_new]) < 3e-2
        assert torch.equal(prod.queue[:, n_new:], q0[:, n_new:]) and torch.equal(ref.queue[:, n_new:], q0[:, n_new:])
        assert torch.equal(prod.series_queue[:, n_new:], sq0[:, n_new:])
        for (n, pr), (_, pp), p0 in zip(ref.encoder_k.named_parameters(), prod.encoder_k.named_parameters(), k0):
            torch.testing.assert_close(pp, pr, rtol=1e-6, atol=1e-7, msg=n)
        moved = max((pr - p0).abs().max().item() for pr, p0 in zip(ref.encoder_k.parameters(), k0))
        assert moved > 0                                      # m = 0.999 really moved the key encoder
    sum(_losses(rp).values()).backward()
    assert all(p.grad is None for p in prod.encoder_k.parameters())
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in prod.encoder_q.parameters())


def test_s3dg_simclr_dualvar_bf16_losses_at_32x128x128(measured):
    """BASELINE configs[3] as stated (S3D-G, 32 x 128 x 128 clips), bf16 mode, 4 samples = 12 + 4 clips: each of the four
    losses within 1e-2 relative of the fp32 oracle (north star), logits within 1e-2 of their range."""
    from dualvar_b200 import models as PM
    from oracle import models as OM
    _no_tf32()
    _seed(0)
    ref = OM.SimCLR_TimeSeriesV4("s3dg", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", ARGS).to(dev).train()
    prod = PM.SimCLR_TimeSeriesV4("s3dg", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", ARGS)
    prod.load_state_dict(ref.state_dict())
    prod = prod.to(dev).train()
    x = torch.randn(4, 3, 3, 32, 128, 128, device=dev).bfloat16().float()
    from bf16_emulation import emulate_bf16
    emu = emulate_bf16(ref)            # keeps fp32 conv weights rounded only inside the copy
    np.random.seed(7); rr = ref(x)
    np.random.seed(7); rp = prod(x)
    with torch.no_grad():
        np.random.seed(7); re_ = emu(x)
        np.random.seed(7)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ra = ref(x)
    report = {}
    for k in rr:
        if "labels" in k:
            assert torch.equal(rr[k], rp[k])
        elif "loss" in k:
            den = abs(rr[k].item())
            report[k] = {"product": abs(rp[k].item() - rr[k].item()) / den, "rounding_oracle": abs(re_[k].item() - rr[k].item()) / den,
                         "torch_autocast": abs(ra[k].float().item() - rr[k].item()) / den, "oracle_value": rr[k].item()}
        else:
            rng = 1.0 if "margin" in k else 1.0 / 0.07
            report[k] = {"product": (rp[k].float() - rr[k].float()).abs().max().item() / rng,
                         "rounding_oracle": (re_[k].float() - rr[k].float()).abs().max().item() / rng,
                         "torch_autocast": (ra[k].float() - rr[k].float()).abs().max().item() / rng}
    measured("s3dg_32x128", report)
    # North star: 1e-2 relative. It does NOT hold for S3D-G at random init in ANY bf16 arithmetic: N(0, 0.01) weights and
    # 77 training-mode BatchNorm layers amplify the roundings - the oracle itself, run with bf16 rounding at the stored
    # tensors, is 7.7e-2 off on the clip loss, torch's autocast 5.1e-2, the product 8.2e-2 (tc 7.8e-3, rank 2.3e-2 /
    # 4.1e-3). What is asserted: the product is no worse than 1.5 x the worse of those two yardsticks, quantity by
    # quantity, and within S3DG_LOSS_TOL absolutely (DESIGN.md 4 states the measured figures).
    for k, r in report.items():
        assert r["product"] <= 1.5 * max(r["rounding_oracle"], r["torch_autocast"]) + 2e-3, (k, r)
        if "loss" in k:
            assert r["product"] <= S3DG_LOSS_TOL, (k, r)
    sum(_losses(rp).values()).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in prod.parameters())


def test_packed_weight_cache_does_not_leak_across_models():
    """Build, step, delete, rebuild: the second model's parameters commonly reuse the ids and device addresses of the
    first one's, and kernels that update parameters through raw pointers do not bump Tensor._version. The packed bf16
    copies must belong to the live parameter object (weakref-keyed cache, engine._cache_get)."""
    from dualvar_b200 import backbones as PB, engine as E
    from oracle import backbones as OB
    _no_tf32()
    x = torch.randn(2, 3, 8, 32, 32, device=dev)
    outs = []
    n_entries = []
    for seed in (1, 2, 3):
        _seed(seed)
        ref, _ = OB.select_backbone("r3d")
        ref = ref.to(dev).train()
        prod, _ = PB.select_backbone("r3d")
        prod.load_state_dict(ref.state_dict())
        prod = prod.to(dev).train()
        with torch.no_grad():
            yr, yp = ref(x), prod(x)
        outs.append(_rel(yp, yr))
        n_entries.append(len(E._weight_cache))
        del prod, ref, yr, yp
        gc.collect()
        torch.cuda.empty_cache()
    assert max(outs) < 5e-2, outs                 # a stale pack gives O(1) error
    assert n_entries[2] <= n_entries[0], n_entries  # entries of freed models are dropped, not accumulated


@pytest.mark.parametrize("which,wrapper", [("simclr", "ddp"), ("moco", "ddp"), ("simclr", "overlap"), ("moco", "overlap")])
def test_syncbn_ddp_parity_two_gpus(which, wrapper):
    """SyncBatchNorm + DistributedDataParallel on 2 GPUs against the oracle wrapped the same way (pretrain.py:244-248):
    tests/dist/ddp_parity.py under torchrun; "overlap" wraps the product in dualvar_b200.parallel.DataParallel (bucketed
    all-reduce inside the engine's backward). Needs two devices (gpurun --gpus 2); skipped on a one-GPU box."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist", "ddp_parity.py"), which, wrapper]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    tail = (out.stdout + out.stderr)[-3000:]
    assert out.returncode == 0, tail
    assert f"DDP_PARITY {which} bf16 mode PASS" in out.stdout, tail
