"""ORACLE tooling (test infrastructure): put the UNMODIFIED reference modules of the hot path where the GPU box can run them.

    python oracle/install_reference.py            # /root/reference -> baseline/_ref/  (git-ignored, travels with gpurun)

The reference (lzhangbj/DualVar) is plain Python + torch with no setup.py: "installing" it is copying the package
directories the path needs - ``backbone/``, ``model/``, ``utils/`` (*.py only) - byte for byte, plus a MANIFEST.json
with their SHA-256 so ``bench.py --impl reference`` can state what it ran. Nothing is copied into tracked paths:
``baseline/_ref/`` is listed in .gitignore (history stays free of reference sources) and NOT in .gpurunignore (it ships
to the GPU box like the built .so files). ``__graft_entry__.build()`` runs this when /root/reference is present; on the
GPU box (no /root/reference) the prebuilt copy is used as is.

``import_reference(root)`` applies the shim of SURVEY.md Appendix B at import time - two stub modules the reference
imports but never uses on this path (``IPython.embed``, ``dataloader.KVReader``), and the ``calc_contrast_loss`` alias its
forward calls but never defines. No file of the reference is edited.
"""
import hashlib
import json
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("DUALVAR_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
PACKAGES = ("backbone", "model", "utils")


def install(src=SRC, dst=DST):
    """Copy the reference packages; returns the manifest, or None when the reference tree is absent."""
    if not os.path.isdir(os.path.join(src, "model")):
        return None
    manifest = {}
    for pkg in PACKAGES:
        os.makedirs(os.path.join(dst, pkg), exist_ok=True)
        for fn in sorted(os.listdir(os.path.join(src, pkg))):
            if not fn.endswith(".py"):
                continue
            a, b = os.path.join(src, pkg, fn), os.path.join(dst, pkg, fn)
            shutil.copyfile(a, b)
            manifest[f"{pkg}/{fn}"] = hashlib.sha256(open(b, "rb").read()).hexdigest()
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": "lzhangbj/DualVar (unmodified copy of backbone/, model/, utils/ *.py)", "sha256": manifest},
                  f, indent=1, sort_keys=True)
    return manifest


def available(root=DST):
    return os.path.isfile(os.path.join(root, "model", "simclr.py")) and os.path.isfile(os.path.join(root, "MANIFEST.json"))


def verify(root=DST):
    """True when every installed file still has the recorded hash (nothing under baseline/_ref was edited)."""
    man = json.load(open(os.path.join(root, "MANIFEST.json")))["sha256"]
    return all(hashlib.sha256(open(os.path.join(root, rel), "rb").read()).hexdigest() == h for rel, h in man.items())


def import_reference(root=DST, cpu=True):
    """(model package, select_backbone, utils.utils) of the reference at ``root`` with the Appendix-B shim.
    cpu=True also maps ``Tensor.cuda`` to the identity (the reference hard-codes ``.cuda()`` on labels and indices)."""
    import torch
    for name, attr in (("IPython", "embed"), ("dataloader", "KVReader")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            setattr(m, attr, None)
            sys.modules[name] = m
    if cpu:
        torch.Tensor.cuda = lambda self, *a, **k: self
    if root not in sys.path:
        sys.path.insert(0, root)
    import model as ref_model
    from backbone.select_backbone import select_backbone as ref_select
    import utils.utils as ref_utils
    ref_model.SimCLR_TimeSeriesV4.calc_contrast_loss = ref_model.SimCLR_TimeSeriesV4.calc_clip_contrast_loss
    ref_model.MoCo_TimeSeriesV4.calc_contrast_loss = ref_model.MoCo_TimeSeriesV4.calc_clip_contrast_loss
    return ref_model, ref_select, ref_utils


if __name__ == "__main__":
    man = install()
    print("reference tree not found at", SRC) if man is None else print(f"installed {len(man)} files into {DST}")
