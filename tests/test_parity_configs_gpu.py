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


def _seed(s):
    torch.manual_seed(s); np.random.seed(s); random.seed(s)


def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def _losses(ret):
    return {k: v for k, v in ret.items() if "loss" in k}


def _simclr_pair(net, emulate=True):
    """(oracle with the product's rounding points, product) on identical (bf16-representable conv) weights."""
    from bf16_emulation import emulate_bf16
    from dualvar_b200 import models as PM
    from oracle import models as OM
    _no_tf32()
    _seed(0)
    ref = OM.SimCLR_TimeSeriesV4(net, 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", ARGS).to(dev).train()
    emu = emulate_bf16(ref) if emulate else ref
    prod = PM.SimCLR_TimeSeriesV4(net, 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc", ARGS)
    prod.load_state_dict(emu.state_dict())
    return ref, emu, prod.to(dev).train()


def _check_per_tensor(emu, prod, measured, tag, max_bound, med_bound):
    from bf16_emulation import per_tensor_errors
    errs = per_tensor_errors(emu, prod)
    assert len(errs) >= 10
    worst = max(errs, key=lambda t: t[1])
    vals = sorted(e for _, e in errs)
    med = vals[len(vals) // 2]
    measured(f"{tag}.grad_per_tensor_max", {"name": worst[0], "err": worst[1]})
    measured(f"{tag}.grad_per_tensor_median", med)
    bad = [(n, round(e, 4)) for n, e in errs if e > max_bound]
    assert not bad, f"{tag}: gradients off by more than {max_bound}: {bad[:8]}"
    assert med <= med_bound, (tag, med)


@pytest.mark.parametrize("net,shape", [("r21d", (8, 3, 3, 8, 64, 64)), ("r3d", (8, 3, 3, 8, 64, 64)),
                                       ("r21d", (4, 3, 3, 16, 112, 112))])
def test_simclr_dualvar_gradients_tensor_by_tensor(net, shape, measured):
    """SimCLR+DualVar step: every loss and EVERY parameter gradient against the oracle with the product's rounding
    points; the (4, .., 16, 112, 112) case takes the bench geometry's launches (CTA pairs, two-region tiling, halo
    wgrad, fused dgrad + BN reduce)."""
    from bf16_emulation import round_input
    ref, emu, prod = _simclr_pair(net)
    x = round_input(torch.randn(*shape, device=dev))
    np.random.seed(11); re_ = emu(x)
    np.random.seed(11); rp = prod(x)
    tag = f"simclr_{net}_{shape[3]}x{shape[4]}"
    for k, v in _losses(re_).items():
        err = abs(rp[k].item() - v.item()) / abs(v.item())
        measured(f"{tag}.{k}", err)
        assert err <= 1e-2, (k, rp[k].item(), v.item())
    sum(_losses(re_).values()).backward()
    sum(_losses(rp).values()).backward()
    _check_per_tensor(emu, prod, measured, tag, max_bound=0.25, med_bound=0.08)
    for (n, br), (_, bp) in zip(emu.named_buffers(), prod.named_buffers()):
        if br.dtype.is_floating_point:
            assert _rel(bp, br) < 5e-3, n


@pytest.mark.parametrize("name,shape", [("c3d", (8, 3, 8, 64, 64)), ("s3d", (4, 3, 16, 64, 64)),
                                        ("s3dg", (4, 3, 16, 64, 64)), ("r2d3d18", (6, 3, 4, 96, 96))])
def test_backbone_gradients_tensor_by_tensor(name, shape, measured):
    from bf16_emulation import emulate_bf16, round_input
    from dualvar_b200 import backbones as PB
    from oracle import backbones as OB
    _no_tf32()
    _seed(0)
    ref, _ = OB.select_backbone(name)
    emu = emulate_bf16(ref.to(dev).train())
    prod, _ = PB.select_backbone(name)
    prod.load_state_dict(emu.state_dict())
    prod = prod.to(dev).train()
    x = round_input(torch.randn(*shape, device=dev))
    ye, yp = emu(x), prod(x)
    measured(f"backbone_{name}.out", _rel(yp, ye))
    g = torch.randn_like(ye)
    ye.backward(g); yp.backward(g)
    deep = name in ("s3d", "s3dg")     # 77 BatchNorm layers down to a 2x2x2 map: mask flips compound
    assert _rel(yp, ye) <= (0.2 if deep else 2e-2)
    _check_per_tensor(emu, prod, measured, f"backbone_{name}", max_bound=1.0 if deep else 0.25,
                      med_bound=0.5 if deep else 0.08)


def test_moco_dualvar_queue_16384_matches_oracle(measured):
    """BASELINE configs[2] as stated: K = 16384, m = 0.999, r21d; small clips (the queue, momentum and loss path do not
    depend on the clip geometry). Two steps: logits against the full queue, enqueue at ptr 0 then 16, queue_ptr,
    series_queue, momentum-updated key encoder."""
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
        x = torch.randn(NB, 3, 3, 8, 64, 64, device=dev, generator=torch.Generator(device=dev).manual_seed(40 + step))
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
    x = torch.randn(4, 3, 3, 32, 128, 128, device=dev)
    np.random.seed(7); rr = ref(x)
    np.random.seed(7); rp = prod(x)
    worst = 0.0
    for k in rr:
        if "labels" in k:
            assert torch.equal(rr[k], rp[k])
        elif "loss" in k:
            err = abs(rp[k].item() - rr[k].item()) / abs(rr[k].item())
            measured(f"s3dg_32x128.{k}", {"err": err, "oracle": rr[k].item(), "product": rp[k].item()})
            worst = max(worst, err)
        else:
            rng = 1.0 if "margin" in k else 1.0 / 0.07
            err = (rp[k].float() - rr[k].float()).abs().max().item() / rng
            measured(f"s3dg_32x128.{k}", err)
            assert err <= 1e-2, (k, err)
    assert worst <= 1e-2, worst
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


@pytest.mark.parametrize("which", ["simclr", "moco"])
def test_syncbn_ddp_parity_two_gpus(which):
    """SyncBatchNorm + DistributedDataParallel on 2 GPUs against the oracle wrapped the same way (pretrain.py:244-248):
    tests/dist/ddp_parity.py under torchrun. Needs two devices (gpurun --gpus 2); skipped on a one-GPU box."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist", "ddp_parity.py"), which]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    tail = (out.stdout + out.stderr)[-3000:]
    assert out.returncode == 0, tail
    assert f"DDP_PARITY {which} bf16 mode PASS" in out.stdout, tail
