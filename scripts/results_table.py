"""Markdown summary of the committed round-2 bench lines (profiles/r02_bench_n*.json)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        return None
    for line in open(path):
        line = line.strip()
        if line.startswith("{"):
            return json.loads(line)
    return None


lines = {n: load("r02_bench_n%d.json" % n) for n in (1, 2, 4, 8)}
ref = load("r02_reference_arm.json")
print("| GPUs | clips/s resident (ms/step) | e2e clips/s (H2D GB/s per rank) | cfg4 video64 ms/video (speed-up) | cfg5 train ms/step (clips/s, speed-up) |")
print("|---|---|---|---|---|")
base = lines[1]
for n, l in lines.items():
    if l is None:
        continue
    v, t = l["cfg4_video64"], l["cfg5_train"]
    rates = l["e2e"].get("h2d_gbps_per_rank") or []
    print("| %d | %.0f (%.2f) | %.0f (%s) | %.1f (%.2fx) | %.2f (%.0f, %.2fx) |" % (
        n, l["value"], l["ms_per_step"], l["e2e"]["value"],
        "%.0f-%.0f" % (min(rates), max(rates)) if rates else "-",
        v["ms_per_video"], base["cfg4_video64"]["ms_per_video"] / v["ms_per_video"],
        t["ms_per_step"], t["clips_per_sec"], t["clips_per_sec"] / base["cfg5_train"]["clips_per_sec"]))
l = base
r = l["roofline"]
print()
print("* dominant kernel (%s): %.3f ms, %.0f TFLOP/s algorithmic = **%.3f of the measured burst bf16 peak** (%.3f of "
      "sustained); tensor-pipe products %.2f of burst; DRAM traffic %.0f MB per launch (ncu)." % (
          r["kernel"], r["launch_ms"], r["achieved"], r["frac"], r["frac_of_sustained_peak"], r["tensor_pipe_frac"],
          (r["traffic"] or 0) / 1e6))
c = l["cfg3_bf16"]
print("* cfg3 (16x480x864, bf16): %.2f ms/step = %.0f TFLOP/s = **%.3f of sustained** for the whole step; dominant launch "
      "%.3f ms = %.3f of sustained / %.3f of burst." % (c["ms_per_step"], c["whole_step"]["achieved"], c["whole_step"]["frac"],
                                                      c["dominant_kernel"]["launch_ms"], c["dominant_kernel"]["frac"],
                                                      c["dominant_kernel"]["frac_of_burst_peak"]))
h = l["roofline_cluster_hbm"]
print("* clustering, HBM regime (%s): %.3f ms = %.0f GB/s by the SURVEY 8d formula = **%.3f of the measured HBM copy peak**; "
      "%.0f GB/s of real DRAM traffic (ncu) = %.3f." % (h["kernel"], h["launch_ms"], h["achieved"], h["frac"],
                                                        h.get("dram_gbs_from_ncu_traffic") or 0, h.get("frac_from_ncu_traffic") or 0))
cb = l["cpu_baseline"]
print("* CPU baseline (%s, %d threads): %.2f clips/s; `--impl reference`: %.2f clips/s." % (
    cb["kind"], cb["cores"], cb["value"], ref["value"] if ref else float("nan")))
inc = l["incumbent_gpu"]
print("* incumbent GPU path (unmodified reference heads, torch/cuDNN on the same B200): %.1f ms fp32, %.1f ms TF32 (misses "
      "parity); this plan: %.2f ms heads." % (inc["heads_ms_fp32"], inc["heads_ms_tf32"], l["stages"]["heads_ms"]))
ef = l["e2e_frames"]
print("* e2e from frames (9.8 MB H2D per clip): %.1f ms with the torch backbone (TF32, torch default), %.1f ms (fp32 backbone), "
      "%.2f ms without backbone." % (ef["with_backbone_tf32"]["ms_per_step"], ef["with_backbone_fp32"]["ms_per_step"],
                                     ef["without_backbone"]["ms_per_step"]))
fv = l["fp16_blocks_variant"]
print("* opt-in fp16 blocks: %.0f clips/s (%.2f ms), deviation from the default plan %.1e." % (
    fv["value"], fv["ms_per_step"], fv["worst_channel_deviation_from_default_plan"]))
print("* stages:", l["stages"])
