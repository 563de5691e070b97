"""A working pretraining driver for the drop-in models.

The reference's pretrain.py cannot run as shipped (missing imports and a method the models never define);
this restates its loop on the parts that do work, with the same conventions so checkpoints interchange:

* one SGD param group per tensor, momentum 0.9, weight decay (pretrain.py:262-272) - the fused kernel of
  ``dualvar_b200.optim.SGD`` on CUDA parameters;
* ``MultiStepLR(schedule, gamma=0.1)`` stepped once per epoch (pretrain.py:328,357);
* the loss of a step is the sum of ``clip_contrast_loss``, every other ``*_contrast_loss`` and every other ``*loss``
  entry of the model's return dict, top-1 accuracy per ``*_logits`` (pretrain.py:401-445);
* checkpoints are ``{'epoch','state_dict','best_acc','optimizer','iteration'}`` written as ``epoch%d.pth.tar`` plus
  ``latest.pth.tar`` and ``model_best_epoch%d.pth.tar`` (pretrain.py:343-357, utils/utils.py:18-44), ``state_dict``
  taken from the model without its DDP wrapper.

What is deliberately different: running sums of losses and accuracies stay on the device and are read back once per
``log_every`` iterations / per epoch instead of a ``.item()`` per loss per iteration (each one a full stream sync).
"""
import glob
import os

import torch
import torch.distributed as dist


def build_optimizer(model, lr, weight_decay=1e-4, momentum=0.9):
    """pretrain.py:262-272. Fused multi-tensor SGD for CUDA parameters, torch.optim.SGD otherwise (oracle on CPU)."""
    params = [{"params": p} for _, p in model.named_parameters()]
    if all(g["params"].is_cuda for g in params):
        from .optim import SGD
        return SGD(params, lr=lr, weight_decay=weight_decay, momentum=momentum)
    return torch.optim.SGD(params, lr=lr, weight_decay=weight_decay, momentum=momentum)


def total_loss(ret):
    """Sum of the losses of one forward in the reference's order (pretrain.py:404-441)."""
    loss = 0
    if "clip_contrast_loss" in ret:
        loss = ret["clip_contrast_loss"]
    extra = [k for k in ret if "loss" in k and "clip" not in k]
    contrast = [k for k in extra if "contrast_loss" in k]
    for k in contrast:
        loss = loss + ret[k]
    for k in extra:
        if k not in contrast:
            loss = loss + ret[k]
    return loss


def _top1(logits, labels):
    """Top-1 accuracy in percent, on the device (utils/utils.py:75-92 with topk=(1,))."""
    return (logits.argmax(dim=1) == labels).float().mean() * 100.0


def train_one_epoch(loader, model, optimizer, to_input=None, log_every=0, log=print, iteration=1, graph_step=None):
    """One pass over ``loader`` (pretrain.py:364-460). ``loader`` yields either ``{'seq': tensor}`` batches like the
    reference's dataset or tensors; ``to_input(batch)`` turns a batch into what ``model`` takes (the reference block
    (B,3,C,T,H,W), or ``engine.RawClips`` for the fused ingest). ``graph_step``: a graph_step.GraphedTrainStep built on
    (model, optimizer) - the batches are then the raw loader frames (B, 3, V*T, H, W) and every step is one CUDA-graph
    replay (forward, losses, backward and optimizer inside). Returns (meters dict of floats, next iteration)."""
    model.train()
    sums, n_it = {}, 0
    for idx, batch in enumerate(loader):
        x = batch["seq"] if isinstance(batch, dict) else batch
        if graph_step is not None:
            ret = dict(graph_step(x))
            loss = ret.pop("loss")
        else:
            x = to_input(x) if to_input is not None else x
            ret = model(x)
            loss = total_loss(ret)
            optimizer.zero_grad(set_to_none=True)
            loss.backward()
            optimizer.step()
        with torch.no_grad():
            stats = {"loss": loss.detach()}
            for k, v in ret.items():
                if k.endswith("loss"):
                    stats[k] = v.detach()
                elif k.endswith("_logits"):
                    stats[k.replace("_logits", "_acc")] = _top1(v.detach(), ret[k.replace("_logits", "_labels")])
            for k, v in stats.items():
                sums[k] = sums[k] + v.float() if k in sums else v.float().clone()
        n_it += 1
        iteration += 1
        if log_every and (idx + 1) % log_every == 0:
            log("it %d  " % (idx + 1) + "  ".join(f"{k} {float(v) / n_it:.4f}" for k, v in sums.items()))
    meters = {k: float(v) / max(n_it, 1) for k, v in sums.items()}
    return meters, iteration


def save_checkpoint(state, is_best, filename, keep_all=False, is_save=True, save_latest=True, gap=0):
    """utils/utils.py:18-44 (same file names and pruning rules, minus the sleep)."""
    folder = os.path.dirname(filename)
    os.makedirs(folder, exist_ok=True)
    if not keep_all:
        try:
            os.remove(os.path.join(folder, "epoch%s.pth.tar" % str(state["epoch"] - gap)))
        except OSError:
            pass
    if is_save:
        torch.save(state, filename)
    if save_latest:
        torch.save(state, os.path.join(folder, "latest.pth.tar"))
    if is_best:
        past = sorted(glob.glob(os.path.join(folder, "model_best_*.pth.tar")),
                      key=lambda x: int("".join(filter(str.isdigit, os.path.basename(x)))))
        if len(past) >= 5:
            try:
                os.remove(past[0])
            except OSError:
                pass
        torch.save(state, os.path.join(folder, "model_best_epoch%s.pth.tar" % str(state["epoch"])))


def load_checkpoint(path, model, optimizer=None, map_location="cpu"):
    """pretrain.py:213-236 (--resume): restores weights, optimizer state, epoch, best accuracy and iteration."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    target = model.module if hasattr(model, "module") else model
    target.load_state_dict(ckpt["state_dict"])
    if optimizer is not None and "optimizer" in ckpt:
        optimizer.load_state_dict(ckpt["optimizer"])
    return ckpt["epoch"] + 1, ckpt.get("best_acc", 0.0), ckpt.get("iteration", 1)


def fit(model, loader, optimizer, epochs, schedule=(), start_epoch=0, model_path=None, save_freq=1, eval_freq=1,
        to_input=None, best_acc=0.0, iteration=1, log=print, log_every=0, graph_step=None):
    """pretrain.py:328-360: epochs of train_one_epoch with MultiStepLR(gamma=0.1) and rank-0 checkpoints."""
    sched = torch.optim.lr_scheduler.MultiStepLR(optimizer, list(schedule), gamma=0.1, last_epoch=start_epoch - 1) \
        if start_epoch == 0 else None
    if sched is None:     # resumed: initial_lr is already in the optimizer's groups
        for g in optimizer.param_groups:
            g.setdefault("initial_lr", g["lr"])
        sched = torch.optim.lr_scheduler.MultiStepLR(optimizer, list(schedule), gamma=0.1, last_epoch=start_epoch - 1)
    rank0 = not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0
    history = []
    for epoch in range(start_epoch, epochs):
        sampler = getattr(loader, "sampler", None)
        if hasattr(sampler, "set_epoch"):
            sampler.set_epoch(epoch)                  # pretrain.py:333
        meters, iteration = train_one_epoch(loader, model, optimizer, to_input, log_every, log, iteration, graph_step)
        sched.step()
        history.append(meters)
        log("epoch %d  " % epoch + "  ".join(f"{k} {v:.4f}" for k, v in meters.items()))
        if model_path and rank0 and (((epoch + 1) % eval_freq == 0) or epoch == epochs - 1):
            acc = meters.get("clip_acc", 0.0)
            is_best = acc > best_acc
            best_acc = max(acc, best_acc)
            target = model.module if hasattr(model, "module") else model
            state = {"epoch": epoch, "state_dict": target.state_dict(), "best_acc": best_acc,
                     "optimizer": optimizer.state_dict(), "iteration": iteration}
            save_checkpoint(state, is_best, os.path.join(model_path, "epoch%d.pth.tar" % epoch),
                            is_save=((epoch + 1) % save_freq == 0))
        if model_path and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            # rank 0 may spend seconds writing checkpoints: the other ranks wait here (host side) instead of inside
            # the first cross-replica BatchNorm exchange of the next epoch
            dist.barrier()
    return history, best_acc, iteration
