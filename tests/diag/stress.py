"""GPU stress: loop each conv kernel on a small geometry, compare every iteration with the first."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # tests helpers (kernel_handles)
import torch
import kernel_handles as K
dev = "cuda:0"
which = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
geoms = [(6, 4, 16, 16, 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
         (2, 4, 14, 14, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
         (3, 4, 7, 7, 921, 512, (3, 1, 1), (2, 1, 1), (1, 0, 0))]
for (N, T, H, W, Cin, Cout, k, s, p) in geoms:
    g = K.make_geom(N, T, H, W, Cin, Cout, k, s, p)
    gen = torch.Generator(device=dev).manual_seed(0)
    x = K.to_ndhwc(torch.randn(N, Cin, T, H, W, device=dev, generator=gen))
    w = torch.randn(Cout, Cin, *k, device=dev, generator=gen) / 20
    wf, wt = K.pack_conv_weight(w, g)
    dy = K.to_ndhwc(torch.randn(N, Cout, g.To, g.Ho, g.Wo, device=dev, generator=gen))
    first = None
    bad = 0
    t0 = time.time()
    for it in range(iters):
        if which == "fprop":
            out = K.conv3d_fprop(x, wf, g).float()
        elif which == "dgrad":
            out = K.conv3d_dgrad(dy, wt, g).float()
        else:
            out = K.conv3d_wgrad_packed(x, dy, g)
        if first is None:
            first = out.clone()
        else:
            tol = 0 if which != "wgrad" else 1e-3 * first.abs().max().item()
            d = (out - first).abs().max().item()
            if d > tol or not torch.isfinite(out).all():
                bad += 1
                if bad < 4:
                    print(f"  iter {it}: maxdiff {d:.4g} (ref max {first.abs().max().item():.4g}) nonfinite={(~torch.isfinite(out)).sum().item()}", flush=True)
    torch.cuda.synchronize()
    print(f"{which} geom {Cin}->{Cout} k{k} N{N}: {iters} iters, {bad} bad, {time.time()-t0:.1f}s", flush=True)
