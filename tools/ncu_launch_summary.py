"""Condense an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) of ONE
training step into per-kernel-class totals: launches, time, share of the step, DRAM bytes.
python tools/ncu_launch_summary.py gpurun_out/launches.csv [out.json]"""
import csv, collections, json, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; iK, iM, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
iID = hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    d = per.setdefault(r[iID], {"name": r[iK]})
    v = float(r[iV].replace(",", "")); u = r[iU]
    if r[iM].startswith("gpu__time"): d["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0) if u != "nsecond" else v * 1e-3
    else:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        d[r[iM]] = v * scale
def cls(name):
    for pat, c in (("gemm_2sm", "gemm (CTA pair)"), ("gemm_kernel", "gemm (1 CTA)"), ("attn16_fwd|attn_fwd", "attn_fwd"), ("dkv", "attn_bwd_dkv"),
                   ("bwd_dq", "attn_bwd_dq"), ("delta", "attn_delta"), ("add_ln_fwd", "add_ln_fwd"), ("add_ln_bwd", "add_ln_bwd"), ("colsum", "colsum"),
                   ("adam", "adam"), ("sumsq", "sumsq"), ("lsce", "lsce"), ("embed", "embed"), ("round_tf32|cast", "round/cast")):
        if re.search(pat, name): return c
    return "other (PyTorch glue, NCCL, memset ...)"
agg = collections.OrderedDict()
for d in per.values():
    a = agg.setdefault(cls(d["name"]), {"launches": 0, "us": 0.0, "dram_bytes": 0.0})
    a["launches"] += 1; a["us"] += d.get("us", 0.0)
    a["dram_bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
tot = sum(a["us"] for a in agg.values())
print(f"{len(per)} launches, {tot / 1e3:.3f} ms of kernel time (cold-cache, serialised: compare SHARES)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    print(f"  {k:38s} n={a['launches']:4d}  {a['us'] / 1e3:8.3f} ms  share {a['us'] / tot:6.3f}  dram {a['dram_bytes'] / 1e9:8.3f} GB  ({a['dram_bytes'] / max(a['launches'], 1) / 1e6:8.2f} MB/launch)")
if len(sys.argv) > 2:
    g = [a for k, a in agg.items() if k.startswith("gemm")]
    n = sum(a["launches"] for a in g); b = sum(a["dram_bytes"] for a in g)
    json.dump({"kernel": "gemm_kernel + gemm_2sm_kernel, EVERY launch of one training step", "launches": n, "dram_bytes_per_step": b,
               "dram_bytes_per_launch": b / max(n, 1), "source": f"{sys.argv[1]} (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, tools/ncu_launch_summary.py)"},
              open(sys.argv[2], "w"), indent=1)
