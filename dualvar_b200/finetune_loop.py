"""Finetune / linear-probe driver for the drop-in ``LinearClassifier`` - classifier.py restated on the
parts that carry the hot path, with the reference's conventions:

* ``build_optimizer``     classifier.py:233-261: ``train_what='last'`` freezes every ``backbone.*`` parameter and hands
  only the head to the optimizer, one param group per tensor; SGD(momentum 0.9) on the fused multi-tensor kernel, Adam
  stays torch's (a handful of (101, 512)-sized tensors).
* ``adjust_learning_rate`` classifier.py:998-1003 (x0.1 at the epochs of ``schedule``; called BEFORE the epoch).
* ``train_one_epoch``     classifier.py:422-498: ``'last'`` keeps the whole model in eval mode (frozen BatchNorm) except
  ``final_bn``; cross-entropy + top-1/top-5; ``tr()`` (Normalize, view, transpose, squeeze) is the ingest kernel's job.
* ``validate``            classifier.py:501-542.
* ``test_10crop``         classifier.py:545-654: centre / five / ten crop protocol (FiveCrop positions x flip), softmax
  averaged over a video's temporal windows per augmentation, accuracy summarised like ``summarize_probability``
  (classifier.py:762-784: every augmentation of a video is scored separately - the ``.mean(0)`` there is commented out).
  ``five_crop_offsets`` reproduces ``A.FiveCrop`` including its bottom-right quirk (``h - tw`` as the top edge,
  utils/augmentation.py:216); ``ten_crop_clips`` runs flip + Scale + FiveCrop + ToTensor on the GPU (frames.scale_crop).
* ``fit``                 classifier.py:390-417: epochs, validation every ``eval_freq``, checkpoints in the reference's
  format through pretrain_loop.save_checkpoint.

Meters stay on the device and are read back once per epoch (the reference does three ``.item()`` syncs per iteration).
"""
import os

import torch
import torch.nn.functional as F

from . import objectives as O
from .pretrain_loop import save_checkpoint

MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def build_optimizer(model, train_what="all", optim="sgd", lr=1e-3, wd=1e-3):
    target = model.module if hasattr(model, "module") else model
    params = []
    for name, p in target.named_parameters():
        if train_what == "last" and "backbone" in name:
            p.requires_grad = False
        else:
            params.append({"params": p})
    if optim == "adam":
        return torch.optim.Adam(params, lr=lr, weight_decay=wd)
    if optim != "sgd":
        raise NotImplementedError(optim)
    if all(g["params"].is_cuda for g in params):
        from .optim import SGD
        return SGD(params, lr=lr, weight_decay=wd, momentum=0.9)
    return torch.optim.SGD(params, lr=lr, weight_decay=wd, momentum=0.9)


def adjust_learning_rate(optimizer, epoch, schedule):
    ratio = 0.1 if epoch in schedule else 1.0
    for g in optimizer.param_groups:
        g["lr"] = g["lr"] * ratio


def _to_input(frames, native):
    """Loader frames (B, 3, seq_len, H, W) in [0, 1] (or uint8) -> model input; classifier.py:441-444,461 (num_seq = 1)."""
    if native and frames.is_cuda:
        from .engine import RawClips
        return RawClips(frames, 1, mean=MEAN, std=STD)
    x = frames.float() / 255.0 if frames.dtype == torch.uint8 else frames.float()
    m = torch.tensor(MEAN, device=x.device).view(1, 3, 1, 1, 1)
    s = torch.tensor(STD, device=x.device).view(1, 3, 1, 1, 1)
    return (x - m) / s


def _ce(logit, target, native):
    """(loss, top-1 hits, top-5 hits): one kernel launch on the product path, torch on the oracle's (CPU tests)."""
    if native and logit.is_cuda:
        loss, hits = O.cross_entropy(logit, target)
        return loss, hits[0], hits[1]
    loss = F.cross_entropy(logit, target)
    top = logit.topk(min(5, logit.shape[1]), dim=1)[1]
    hit = top == target.unsqueeze(1)
    return loss, hit[:, :1].any(1).sum(), hit.any(1).sum()


def _set_mode(model, train_what, use_bn):
    if train_what == "last":
        model.eval()                                   # totally freeze BN in backbone (classifier.py:435-438)
    else:
        model.train()
    target = model.module if hasattr(model, "module") else model
    if use_bn and getattr(target, "final_bn", None) is not None:
        target.final_bn.train()


def train_one_epoch(loader, model, optimizer, train_what="all", use_bn=False, native=True, device=None):
    """classifier.py:422-498. Returns {'loss', 'top1', 'top5'} (sample-weighted averages, accuracies in percent)."""
    _set_mode(model, train_what, use_bn)
    s_loss = s_t1 = s_t5 = None
    n = 0
    for batch in loader:
        frames, target = batch["seq"], torch.as_tensor(batch["vid"]).view(-1)
        if device is not None:
            frames, target = frames.to(device, non_blocking=True), target.to(device, non_blocking=True)
        B = frames.shape[0]
        logit, _ = model(_to_input(frames, native))
        loss, t1, t5 = _ce(logit, target.to(logit.device).long(), native)
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        with torch.no_grad():
            l_b = loss.detach().float() * B
            s_loss = l_b if s_loss is None else s_loss + l_b
            s_t1 = t1.float() if s_t1 is None else s_t1 + t1.float()
            s_t5 = t5.float() if s_t5 is None else s_t5 + t5.float()
        n += B
    return {"loss": float(s_loss) / n, "top1": 100.0 * float(s_t1) / n, "top5": 100.0 * float(s_t5) / n, "n": n}


@torch.no_grad()
def validate(loader, model, native=True, device=None):
    """classifier.py:501-542."""
    model.eval()
    s_loss = s_t1 = s_t5 = None
    n = 0
    for batch in loader:
        frames, target = batch["seq"], torch.as_tensor(batch["vid"]).view(-1)
        if device is not None:
            frames, target = frames.to(device, non_blocking=True), target.to(device, non_blocking=True)
        B = frames.shape[0]
        logit, _ = model(_to_input(frames, native))
        loss, t1, t5 = _ce(logit, target.to(logit.device).long(), native)
        l_b = loss.float() * B
        s_loss = l_b if s_loss is None else s_loss + l_b
        s_t1 = t1.float() if s_t1 is None else s_t1 + t1.float()
        s_t5 = t5.float() if s_t5 is None else s_t5 + t5.float()
        n += B
    return {"loss": float(s_loss) / n, "top1": 100.0 * float(s_t1) / n, "top5": 100.0 * float(s_t5) / n, "n": n}


# --------------------------------------------------------------------------------------------- 10-crop test
CROP_LISTS = {"center": ([5], [0]), "five": ([5, 1, 2, 3, 4], [0]), "ten": ([5, 1, 2, 3, 4], [0, 1])}


def five_crop_offsets(w, h, tw, th, where):
    """(left, top) of ``A.FiveCrop(size=(th, tw), where)`` on a w x h frame (utils/augmentation.py:194-220):
    1 top-left, 2 top-right, 3 bottom-left, 4 bottom-right (top edge ``h - tw`` as the reference writes it), 5 centre."""
    if th > h or tw > w:
        raise ValueError("Requested crop size {} is bigger than input size {}".format((th, tw), (h, w)))
    if where == 1:
        return 0, 0
    if where == 2:
        return w - tw, 0
    if where == 3:
        return 0, h - th
    if where == 4:
        return w - tw, h - tw
    if where == 5:
        return int(round((w - tw) / 2.0)), int(round((h - th) / 2.0))
    raise ValueError(where)


def ten_crop_clips(frames_u8, where, flip, n_views, scale_size=(128, 171), crop_size=(112, 112)):
    """Decoded uint8 frames (B, F, Hs, Ws, 3) on the GPU -> planar uint8 clips (B, 3, F, ch, cw): the test transform of
    classifier.py:593-603 - RandomHorizontalFlip(command='left' | 'right') = never / always flip, Scale, FiveCrop(where),
    ToTensor (the x/255 happens in the ingest kernel). Bit-exact with the Pillow chain (tests/test_finetune_loop.py)."""
    from . import frames as FR
    if flip:
        frames_u8 = torch.flip(frames_u8, dims=[3]).contiguous()       # Image.FLIP_LEFT_RIGHT before Scale, as the reference
    B = frames_u8.shape[0]
    sw, sh = scale_size                 # PIL order (width, height), as A.Scale hands it to Image.resize
    cw, ch = crop_size
    left, top = five_crop_offsets(sw, sh, cw, ch, where)
    crops = torch.tensor([left, top], dtype=torch.int32).view(1, 1, 2).expand(B, n_views, 2).contiguous()
    return FR.scale_crop(frames_u8, crops, n_views, scale_size=scale_size, crop_size=crop_size)


@torch.no_grad()
def test_10crop(make_loader, model, mode="ten", seq_len=16, native=True, device=None, log=None):
    """classifier.py:545-654. ``make_loader(aug_idx, flip_idx)`` yields the batches of the dataset under that test
    transform: {'seq': (B, 3, n_windows*seq_len, H, W) frames, 'vid', 'vpath' (one name per video)}.
    Returns {'center' | 'five' | 'ten': (acc@1, acc@5)} for the protocols completed, summarised per video like
    summarize_probability: every collected mean-probability row of a video is scored against its label."""
    model.eval()
    aug_list, flip_list = CROP_LISTS[mode]
    prob, label = {}, {}
    out = {}

    def summarise():
        a1 = a5 = 0.0
        for v, rows in prob.items():
            p = torch.stack(rows, 0)
            top = p.topk(min(5, p.shape[1]), dim=1)[1]
            hit = top == label[v]
            a1 += hit[:, :1].any(1).float().mean().item() * 100.0
            a5 += hit.any(1).float().mean().item() * 100.0
        return a1 / len(prob), a5 / len(prob)

    for flip_idx in flip_list:
        for aug_idx in aug_list:
            for batch in make_loader(aug_idx, flip_idx):
                frames = batch["seq"] if device is None else batch["seq"].to(device, non_blocking=True)
                B, _, L, H, W = frames.shape
                nwin = L // seq_len
                # tr(): (B, 3, nwin*seq_len, H, W) -> (B*nwin, 3, seq_len, H, W) in (b, window) order
                if native and frames.is_cuda:
                    from .engine import RawClips
                    x = RawClips(frames, nwin, mean=MEAN, std=STD)
                else:
                    x = _to_input(frames, False).view(B, 3, nwin, seq_len, H, W).permute(0, 2, 1, 3, 4, 5) \
                        .reshape(B * nwin, 3, seq_len, H, W)
                logit, _ = model(x)
                pm = F.softmax(logit.float(), dim=-1).view(B, nwin, -1).mean(1)       # average over the temporal windows
                vids = torch.as_tensor(batch["vid"]).view(-1)
                for i, v in enumerate(batch["vpath"]):
                    prob.setdefault(v, []).append(pm[i])
                    label[v] = int(vids[i])
            if mode == "ten" and flip_idx == 0 and aug_idx == 5:
                out["center"] = summarise()
        if mode == "ten" and flip_idx == 0:
            out["five"] = summarise()
    out[mode] = summarise()
    if log is not None:
        for k, (a1, a5) in out.items():
            log(f"{k}-crop: Acc@1 {a1:.4f} Acc@5 {a5:.4f}")
    return out


def fit(model, train_loader, val_loader, optimizer, epochs, schedule=(), start_epoch=0, train_what="all", use_bn=False,
        eval_freq=1, save_freq=1, model_path=None, native=True, device=None, log=print, best_acc=0.0):
    """classifier.py:390-417."""
    history = []
    iteration = 1
    for epoch in range(start_epoch, epochs):
        sampler = getattr(train_loader, "sampler", None)
        if hasattr(sampler, "set_epoch"):
            sampler.set_epoch(epoch)
        adjust_learning_rate(optimizer, epoch, schedule)
        tr = train_one_epoch(train_loader, model, optimizer, train_what, use_bn, native, device)
        iteration += len(train_loader) if hasattr(train_loader, "__len__") else 0
        rec = {"epoch": epoch, "train": tr}
        if (epoch + 1) % eval_freq == 0:
            va = validate(val_loader, model, native, device)
            rec["val"] = va
            is_best = va["top1"] > best_acc
            best_acc = max(va["top1"], best_acc)
            if model_path:
                target = model.module if hasattr(model, "module") else model
                state = {"epoch": epoch, "state_dict": target.state_dict(), "best_acc": best_acc,
                         "optimizer": optimizer.state_dict(), "iteration": iteration}
                save_checkpoint(state, is_best, os.path.join(model_path, "epoch%d.pth.tar" % epoch), keep_all=False,
                                is_save=((epoch + 1) % save_freq == 0))
        log("epoch %d  train loss %.4f acc@1 %.2f" % (epoch, tr["loss"], tr["top1"]) +
            ("  val loss %.4f acc@1 %.2f" % (rec["val"]["loss"], rec["val"]["top1"]) if "val" in rec else ""))
        history.append(rec)
    return history, best_acc
