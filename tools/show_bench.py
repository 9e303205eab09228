"""Pretty-print a bench.py JSON line: python tools/show_bench.py gpurun_out/bench.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"ms/step {d['ms_per_step']:.2f}  value {d['value']:.0f} {d['unit']}  e2e {d['e2e']['value']:.0f}  launches {d['gpu_launches']}  clocks {d.get('clocks')}")
if d.get("roofline"):
    r = d["roofline"]; print(f"roofline: {r['achieved']:.1f} {r['unit']} = {r['frac']:.3f} of bf16 peak, {r['frac_of_tf32_peak']:.3f} of TF32 peak {r['peak_tf32_measured']:.0f}")
for k, v in (d.get("kernel_breakdown") or {}).items():
    print(f"  {k:16s} {v['ms_per_step']:8.3f} ms  share {v['share_of_step']:.3f}  launches {v['launches_per_step']:.0f}  rate {v['rate']:.2f} T(FLOP|B)/s")
if d.get("cpu_baseline"): print("cpu_baseline:", d["cpu_baseline"])
if d.get("encoder_layer"): e = d["encoder_layer"]; print(f"encoder layer fwd+bwd: {e['ms_fwd_bwd']:.3f} ms  {e['tflops']:.1f} TFLOP/s = {e['frac_of_tf32_peak_measured']:.3f} of TF32 peak, {e['frac_of_bf16_peak']:.3f} of bf16 peak")
