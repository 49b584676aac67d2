#!/usr/bin/env python
"""Tables for BASELINE.md / DESIGN.md from the committed evidence under profiles/r02/ (no GPU needed)."""
import glob, json, os, sys
D = os.path.join(os.path.dirname(os.path.abspath(__file__)), "r02")
def load(n):
    try: return json.load(open(os.path.join(D, n)))
    except Exception: return None
b = load("bench_f32_bs1.json"); b4 = load("bench_f32_bs4.json"); bb = load("bench_bf16_bs1.json"); b2 = load("bench_f32_bs1_2gpus.json")
ref = load("bench_reference_arm.json")
peak = b["roofline"]["peak"]
print("| Shape | dtype | GPUs | B1 ref-CUDA fwd / bwd (us) | ours fwd / bwd (us) | achieved GB/s | %% of %.1f GB/s | B2 CPU (GB/s, cores) |" % peak)
print("|---|---|---|---|---|---|---|---|")
rc = b["reference_cuda_op_same_gpu"]; pc = b["per_call_us"]
for k in ("det", "map", "plan", "ego"):
    v = pc[k]; tot = v["B_fwd"] + v["B_bwd"]; gbs = tot / ((v["fwd"] + v["bwd"]) * 1e-6) / 1e9
    print("| %s call @352x640, bs=1 | f32 | 1 | %.0f / %.0f | %.1f / %.1f | %.0f | %.1f | |" % (k, rc[k]["ref_fwd_us"], rc[k]["ref_bwd_us"], v["fwd"], v["bwd"], gbs, 100 * gbs / peak))
for name, x in (("stage-2 step (24 calls), bs=1", b), ("stage-2 step, bs=4", b4), ("stage-2 step, bs=1, bf16 features", bb), ("stage-2 step, bs=1 per GPU", b2)):
    if x is None: continue
    cpu = "%.3f, %d" % (x["cpu_baseline"]["value"], x["cpu_baseline"]["cores"]) if x.get("cpu_baseline") else ""
    print("| %s | %s | %d | | %.3f ms per step (grouped: %.3f ms) | %.0f | %.1f | %s |" % (name, x["dtype"], x["n_gpus"], x["ms_per_step"], x["group_step"]["ms_per_step"], x["value"], 100 * x["roofline"]["step_frac_of_peak"], cpu))
print()
print("| layer (det+map+plan+ego) | feature MB | fwd grouped us | bwd grouped us | fwd / bwd one call at a time us |")
print("|---|---|---|---|---|")
for f in sorted(glob.glob(os.path.join(D, "layer_*.json"))):
    x = json.load(open(f)); l = x["layer"]
    print("| bs=%d %s %dx%d plan %d | %.1f | %.1f | %.1f | %.1f / %.1f |" % (x["bs"], x["dtype"], x["input_hw"][0], x["input_hw"][1], x["plan_anchors"], x["feature_MB"], l["fwd_grouped_us"], l["bwd_grouped_us"], l["fwd_sum_group1_us"], l["bwd_sum_group1_us"]))
print()
d = b["decoder_forward"]
for hw in ("352x640", "256x704"):
    if hw in d:
        print("decoder forward %s: " % hw + ", ".join("%s %.1f ms (%.1f samples/s, DFA modules %.1f ms)" % (k, v["ms_per_forward_median"], v["samples_per_s"], v["dfa_module_ms_per_forward"]) for k, v in d[hw].items() if "ms_per_forward_median" in v))
t = b["train_step"]
print("train step bs=%d: ours %.1f ms (%.2f samples/s), DFA kernels fwd %.2f + bwd %.2f ms; reference op %s" % (t["bs_per_gpu"], t["ms_per_step"], t["samples_per_s"], t["dfa_kernel_ms_per_step"]["forward"], t["dfa_kernel_ms_per_step"]["backward"], t.get("reference_cuda_op")))
if b2: print("train step 2 GPUs:", {k: b2["train_step"].get(k) for k in ("ms_per_step", "samples_per_s", "allreduce_standalone_ms", "allreduce_bus_gbs")})
print("e2e:", b["e2e"]["ms_per_step"], "ms", b["e2e"]["value"], "GB/s; reference arm", ref and ref["value"], "GB/s")
print("roofline:", json.dumps(b["roofline"]))
for k in b["roofline_per_kernel"]: print("  ", k)
