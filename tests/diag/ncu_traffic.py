"""ncu CSV (tests/diag/ncu_step.py capture) -> profiles/r02_ncu_traffic.json, read by bench.py.

    python tests/diag/ncu_traffic.py <ncu.csv> <out.json>

Output: {"layers": {"<entry point> <geometry>": {launches, dram_bytes_per_launch, dram_read_bytes, dram_write_bytes,
us_per_launch}}, "kernels": {"conv_tile_kernel" | "conv_wgrad_kernel" | ...: {avg_dram_bytes_per_launch, launches, ...}},
"step": {launches, us_total, share per group}}. Launch names are NVTX range names (= KernelTimer keys with '_' for ' ').
"""
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6,
        "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
GROUPS = {"conv_tile_kernel": ("dv_conv3d_fprop_bf16", "dv_conv3d_dgrad_bf16", "dv_conv3d_dgrad_bnred_bf16",
                               "dv_conv3d_stem_fprop_bf16", "dv_conv3d_fprop_bnrelu_bf16"),
          "conv_wgrad_kernel": ("dv_conv3d_wgrad_bf16", "dv_conv3d_stem_wgrad_bf16", "dv_conv3d_wgrad_bnrelu_bf16"),
          "bn_passes": ("dv_bn_apply", "dv_bn_bwd_reduce", "dv_bn_bwd_apply")}


def main(src, dst):
    rows = []
    with open(src, newline="") as f:
        lines = f.readlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith('"ID"'))
    for r in csv.DictReader(lines[start:]):
        rows.append(r)
    launches = {}
    for r in rows:
        d = launches.setdefault(r["ID"], {"name": r["Kernel Name"], "read": 0.0, "write": 0.0, "us": 0.0})
        v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
        if r["Metric Name"] == "dram__bytes_read.sum":
            d["read"] = v
        elif r["Metric Name"] == "dram__bytes_write.sum":
            d["write"] = v
        elif r["Metric Name"] == "gpu__time_duration.sum":
            d["us"] = v
    layers, total_us = {}, 0.0
    for d in launches.values():
        total_us += d["us"]
        name = d["name"].split("/")[0]        # "<NVTX range>/<CUDA kernel>"
        if not name.startswith("dv_"):
            continue
        # NVTX names carry '_' for ' ': "<entry>_N192_16x56x56_64->144_k133_s111" or "<entry>_<rows>,<Cp>,<ld>"
        for ep in sorted((e for g in GROUPS.values() for e in g), key=len, reverse=True):
            if name.startswith(ep + "_") or name == ep:
                key = ep + " " + name[len(ep) + 1:].replace("_", " ")
                break
        else:
            continue
        L = layers.setdefault(key.strip(), {"launches": 0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "us": 0.0})
        L["launches"] += 1
        L["dram_read_bytes"] += d["read"]
        L["dram_write_bytes"] += d["write"]
        L["us"] += d["us"]
    for L in layers.values():
        # one C-ABI call may be several kernel launches (two-region tiling, one launch per stride-parity class):
        # bench.py divides the per-step totals by ITS calls per step
        L["dram_bytes_per_step"] = L["dram_read_bytes"] + L["dram_write_bytes"]
        L["dram_bytes_per_launch"] = L["dram_bytes_per_step"] / L["launches"]
        L["us_per_launch"] = L["us"] / L["launches"]
    kernels = {}
    for g, eps in GROUPS.items():
        sel = [L for k, L in layers.items() if k.split(" ")[0] in eps]
        if sel:
            n = sum(L["launches"] for L in sel)
            by = sum(L["dram_read_bytes"] + L["dram_write_bytes"] for L in sel)
            us = sum(L["us"] for L in sel)
            kernels[g] = {"launches": n, "avg_dram_bytes_per_launch": by / n, "dram_read_bytes": sum(L["dram_read_bytes"] for L in sel),
                          "dram_write_bytes": sum(L["dram_write_bytes"] for L in sel), "us": us,
                          "share_of_step": us / total_us if total_us else None}
    out = {"source": f"ncu --nvtx --print-nvtx-rename kernel --metrics dram__bytes_read.sum,dram__bytes_write.sum,"
                     f"gpu__time_duration.sum over ONE bench step (tests/diag/ncu_step.py; cold cache, serialised): {src}",
           "step": {"launches": len(launches), "us_total": total_us}, "kernels": kernels, "layers": layers}
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    print(f"{len(launches)} launches, {len(layers)} layers -> {dst}")
    for g, k in kernels.items():
        print(f"  {g}: {k['launches']} launches, {k['us'] / 1e3:.2f} ms ({100 * k['share_of_step']:.1f} % of the step), "
              f"{k['avg_dram_bytes_per_launch'] / 1e6:.1f} MB DRAM per launch")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
