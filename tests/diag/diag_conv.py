"""GPU diagnostic (not a pytest test): run the tcgen05 conv trio over a shape sweep and print
error summaries against torch fp32 conv on bf16-rounded operands. Usage: python tests/diag/diag_conv.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # tests helpers (kernel_handles)
import torch
import torch.nn.functional as F
import kernel_handles as K

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda:0"

CASES = [
    # name, N,T,H,W, Cin,Cout, kernel, stride, pad
    ("1x1x1 64->64", 2, 4, 16, 16, 64, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0)),
    ("spatial 64->144", 2, 4, 14, 14, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("temporal 144->64", 2, 4, 14, 14, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("spatial s2 64->230", 2, 4, 28, 28, 64, 230, (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    ("temporal s2 230->128", 2, 8, 14, 14, 230, 128, (3, 1, 1), (2, 1, 1), (1, 0, 0)),
    ("down 1x1 s(1,2,2) 64->42", 2, 4, 28, 28, 64, 42, (1, 1, 1), (1, 2, 2), (0, 0, 0)),
    ("down 1x1 s(2,1,1) 42->128", 2, 8, 14, 14, 42, 128, (1, 1, 1), (2, 1, 1), (0, 0, 0)),
    ("stem 3->83 7x7 s2", 2, 4, 32, 32, 3, 83, (1, 7, 7), (1, 2, 2), (0, 3, 3)),
    ("3x3x3 64->64", 2, 4, 14, 14, 64, 64, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    ("3x3x3 s2 64->128", 2, 8, 14, 14, 64, 128, (3, 3, 3), (2, 2, 2), (1, 1, 1)),
    ("spatial 128->288 (2 n-tiles)", 2, 4, 14, 14, 128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("spatial 256->576 7x7", 3, 2, 7, 7, 256, 576, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("temporal 921->512 s2", 3, 4, 7, 7, 921, 512, (3, 1, 1), (2, 1, 1), (1, 0, 0)),
    ("c3d conv1 3->64 bias", 2, 4, 16, 16, 3, 64, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
]


def rel(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def run_case(name, N, T, H, W, Cin, Cout, k, s, p):
    g = K.make_geom(N, T, H, W, Cin, Cout, k, s, p)
    gen = torch.Generator(device=dev).manual_seed(hash(name) % (2**31))
    x = torch.randn(N, Cin, T, H, W, device=dev, generator=gen)
    w = torch.randn(Cout, Cin, *k, device=dev, generator=gen) / (Cin * k[0] * k[1] * k[2]) ** 0.5
    xb = x.bfloat16().float()
    wb = w.bfloat16().float()
    out = {}
    # reference
    xr = xb.clone().requires_grad_(True)
    wr = wb.clone().requires_grad_(True)
    yr = F.conv3d(xr, wr, None, s, p)
    dy = torch.randn(yr.shape, device=dev, generator=gen).bfloat16().float()
    yr.backward(dy)
    # ours
    x_nd = K.to_ndhwc(x)
    back = K.from_ndhwc(x_nd, Cin)
    out["layout_rt"] = rel(back, xb)
    wf, wt = K.pack_conv_weight(w, g)
    stats = torch.zeros(2 * g.Cout_p, dtype=torch.float64, device=dev)
    y_nd = K.conv3d_fprop(x_nd, wf, g, bn_stats=stats)
    torch.cuda.synchronize()
    y = K.from_ndhwc(y_nd, Cout)
    out["fprop"] = rel(y, yr.detach())
    pad_ok = True
    if g.Cout_p > Cout:
        pad_ok = bool((y_nd[..., Cout:] == 0).all().item())
    out["pad0"] = pad_ok
    # stats vs stored values
    ys = y_nd.float()[..., :Cout].reshape(-1, Cout).double()
    s_ref = ys.sum(0)
    q_ref = (ys * ys).sum(0)
    out["stat_sum"] = ((stats[:Cout] - s_ref).abs().max() / (s_ref.abs().max() + 1e-9)).item()
    out["stat_sq"] = ((stats[g.Cout_p:g.Cout_p + Cout] - q_ref).abs().max() / (q_ref.abs().max() + 1e-9)).item()
    # dgrad
    dy_nd = K.to_ndhwc(dy)
    dx_nd = K.conv3d_dgrad(dy_nd, wt, g)
    torch.cuda.synchronize()
    dx = K.from_ndhwc(dx_nd, Cin)
    out["dgrad"] = rel(dx, xr.grad)
    # wgrad
    dwp = K.conv3d_wgrad_packed(x_nd, dy_nd, g)
    dw = K.unpack_conv_wgrad(dwp, g)
    torch.cuda.synchronize()
    out["wgrad"] = rel(dw, wr.grad)
    return out


def main():
    print("device", torch.cuda.get_device_name(0))
    bad = 0
    for c in CASES:
        try:
            r = run_case(*c)
            flag = "OK " if (r["fprop"] < 1e-2 and r["dgrad"] < 1e-2 and r["wgrad"] < 1e-2 and r["pad0"]
                            and r["stat_sum"] < 1e-4 and r["stat_sq"] < 1e-4) else "BAD"
            bad += flag == "BAD"
            print(flag, c[0], {k: (f"{v:.3e}" if isinstance(v, float) else v) for k, v in r.items()}, flush=True)
        except Exception as e:  # noqa
            bad += 1
            print("EXC", c[0], repr(e)[:400], flush=True)
            try:
                torch.cuda.synchronize()
            except Exception as e2:
                print("device dead:", repr(e2)[:200]); break
    print("bad cases:", bad)


if __name__ == "__main__":
    main()
