"""GPU diagnostic: forward-only encoder throughput (evaluation / feature extraction / MoCo key encoder) with the BatchNorm
+ ReLU of the factorised convs applied inside the consumer conv (engine.FUSE_BN_APPLY=1, default) vs stand-alone (0)."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from dualvar_b200 import backbones as PB, engine as E
dev = "cuda:0"
torch.manual_seed(0); np.random.seed(0); random.seed(0)
net, _ = PB.select_backbone("r21d")
net = net.to(dev)
x = torch.rand(192, 3, 16, 112, 112, device=dev)
for train_mode in (True, False):
    net.train(train_mode)
    for mode in (0, 1):
        E.FUSE_BN_APPLY = mode
        with torch.no_grad():
            for _ in range(3):
                net.encode(x, pooled=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                net.encode(x, pooled=True)
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"r21d forward only, 192 clips 16x112x112, {'batch statistics' if train_mode else 'running statistics'}, "
              f"FUSE_BN_APPLY={mode}: {ms:.2f} ms  {192 / ms * 1e3:.0f} clips/s", flush=True)
