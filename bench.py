#!/usr/bin/env python
"""Benchmark of the STEm-Seg hot path on B200: decoder heads + foreground gather + sequential clustering.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision fp32|bf16]

A *step* is one pass of the hot path over one synthetic 8x480x854 clip (padded to 480x864 like
structures/image_list.py:93-95): FPN pyramid [1,256,8,{15x27,30x54,60x108,120x216}] fp32 -> embedding head +
seediness head (DAVIS config, davis_1.yaml) -> foreground compaction/gather (all voxels foreground, bandwidth
activation fused) -> SequentialClustering(0.5, 0.3, min_seediness_prob=0.0) over the 207 360 embedding-grid points ->
int64 labels.  BASELINE.json configs[1].  With N > 1 every rank processes its own clips (weak scaling: the path
shards at sub-clip granularity with no data-path collective).

`value` times K steps with the pyramid resident in HBM (CUDA events, max over ranks).  `e2e` times the same call with
the pyramid in pinned host memory: H2D of the 282 MB pyramid and D2H of the labels inside the timed region.
`--impl reference` times the reference's CPU path (oracle port: the same torch-CPU ATen kernels the reference's
modules call, all host threads) on the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T, H, W = 8, 480, 854
HP, WP = 480, 864                      # padded to multiples of 32
H4, W4 = HP // 4, WP // 4
IN_CH = 256
INTER = (256, 256, 128, 128)
GRID_POINTS = T * H4 * W4              # 207 360
VOXELS_PER_CLIP = T * H * W            # 3 279 360 input-resolution voxels
WORKLOAD = "8x480x854 clip (pad 480x864): DAVIS heads (embedding+seediness, [256,256,128,128]) + fg gather + " \
           "SequentialClustering over 207360 points"


def make_features_cpu(seed=0):
    import torch
    g = torch.Generator().manual_seed(seed)
    feats = {}
    for s in (32, 16, 8, 4):
        feats[s] = torch.randn(1, IN_CH, T, HP // s, WP // s, generator=g, dtype=torch.float32)
    return feats


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's CPU path
# ------------------------------------------------------------------------------------------------------------------
def build_cpu_reference(seed=42):
    import torch
    from oracle import decoder_oracle as do
    emb_shapes = do.head_parameter_shapes("embedding", IN_CH, list(INTER), embedding_size=4, dim_mode="xyff",
                                          seediness_output=False)
    seed_shapes = do.head_parameter_shapes("seediness", IN_CH, list(INTER))
    return do.seeded_state_dict(emb_shapes, seed), do.seeded_state_dict(seed_shapes, seed + 1)


def cpu_reference_step(feats, emb_sd, seed_sd):
    """One clip through the oracle (torch CPU fp32 heads + numpy gather + numpy clustering)."""
    import numpy as np
    import torch
    from oracle import cluster_oracle as co
    from oracle import decoder_oracle as do
    from oracle import gather_oracle as go
    with torch.no_grad():
        f = [feats[s] for s in (32, 16, 8, 4)]
        out = do.embedding_head(emb_sd, f, T, 4, "xyff", True, False)[0]
        seediness = do.seediness_head(seed_sd, f, T)[0]
        emb, var = out[:4], out[4:6]
        bw = var.exp() * 10.0
    mask = np.ones((T, H4, W4), dtype=bool)
    coords, _ = go.masks_to_coord_list(mask)
    e, b, s = go.gather_foreground(coords, emb.numpy(), bw.numpy(), seediness.numpy())
    labels, meta = co.sequential_cluster(e, b, s, 0.5, 0.3, 0.0, 2, [0.3, 0.3])
    return labels, meta


def best_cpu_threads(feats, emb_sd, seed_sd):
    """The reference's torch-CPU convs do not scale to every core of a big host: time one clip at a few thread
    counts and keep the fastest (this is the most favourable setting for the CPU baseline)."""
    import torch
    ncpu = os.cpu_count() or 1
    best, best_t = ncpu, None
    for threads in sorted({ncpu, max(1, ncpu // 2), max(1, ncpu // 4), min(ncpu, 16)}, reverse=True):
        torch.set_num_threads(threads)
        t0 = time.perf_counter()
        cpu_reference_step(feats, emb_sd, seed_sd)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = threads, dt
    torch.set_num_threads(best)
    return best


REFERENCE_BUDGET_S = 120.0


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    import torch
    feats = make_features_cpu()
    emb_sd, seed_sd = build_cpu_reference()
    threads = best_cpu_threads(feats, emb_sd, seed_sd)
    for _ in range(max(0, min(args.warmup, 3) - 1)):          # best_cpu_threads already ran the step 4 times
        cpu_reference_step(feats, emb_sd, seed_sd)
    # bounded: full clips, but never more than REFERENCE_BUDGET_S of CPU time (a clip takes 0.8-11 s on the hosts seen
    # so far); the number of clips actually timed is reported in `steps` / `cpu_baseline.sample`
    steps, t0 = 0, time.perf_counter()
    while steps < args.steps and (steps < 2 or time.perf_counter() - t0 < REFERENCE_BUDGET_S):
        cpu_reference_step(feats, emb_sd, seed_sd)
        steps += 1
    dt = time.perf_counter() - t0
    args.steps = steps
    value = args.steps / dt
    line = {
        "impl": "reference", "metric": "clips_per_sec", "value": value, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "timing": "host wall clock"},
        "mvoxels_per_sec": value * VOXELS_PER_CLIP / 1e6,
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": threads, "kind": "port",
                         "sample": "%d full clips (oracle port of the reference's torch-CPU heads + clustering)" % args.steps},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--id=%d" % self.gpu_index, "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for row in open(self.path):
                parts = [p.strip() for p in row.split(",")]
                if len(parts) < 6:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[2:6]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            p = json.load(open(path))
            return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                    "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def time_incumbent_gpu_heads(device, reps=5):
    """Embedding + seediness heads of the bench clip as torch ops on the GPU (oracle/decoder_oracle.py is the
    functional form of the reference modules): ms per clip, CUDA events, after warm-up (cuDNN autotune off: the
    reference does not enable it)."""
    import torch
    from oracle import decoder_oracle as do
    emb_sd, seed_sd = build_cpu_reference()
    emb_sd = {k: v.to(device) for k, v in emb_sd.items()}
    seed_sd = {k: v.to(device) for k, v in seed_sd.items()}
    feats = [f.to(device) for f in (make_features_cpu()[s] for s in (32, 16, 8, 4))]
    out = {}
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for label, tf32 in (("heads_ms_fp32", False), ("heads_ms_tf32", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            with torch.no_grad():
                for _ in range(2):
                    do.embedding_head(emb_sd, feats, T, 4, "xyff", True, False)
                    do.seediness_head(seed_sd, feats, T)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record()
                for _ in range(reps):
                    do.embedding_head(emb_sd, feats, T, 4, "xyff", True, False)
                    do.seediness_head(seed_sd, feats, T)
                b.record()
                torch.cuda.synchronize()
            out[label] = a.elapsed_time(b) / reps
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    out["note"] = "torch %s / cuDNN %s eager, heads only (no gather / clustering), same clip and weights shape" % (
        torch.__version__, torch.backends.cudnn.version())
    return out


def cluster_fullres_roofline(device, peaks, n=T * HP * WP, e=4, iters=5):
    """SequentialClustering at full resolution (--resize_embeddings, inference/main.py:242-243): the working set
    (80 MB) no longer fits comfortably next to everything else and the kernel is bandwidth-bound.
    Algorithmic bytes = N*(4E+12)*(K+1)  (SURVEY.md §8d)."""
    import numpy as np
    import torch
    from stemseg_b200.clusterers import SequentialClustering
    rng = np.random.default_rng(0)
    centres = rng.uniform(-1, 1, size=(24, e)).astype(np.float32)
    which = rng.integers(0, 24, size=n)
    emb = torch.from_numpy(centres[which] + 0.05 * rng.standard_normal((n, e)).astype(np.float32)).to(device)
    bw = torch.full((n, e - 2), 100.0, device=device)
    seed = torch.rand(n, 1, device=device)
    clusterer = SequentialClustering(0.5, 0.3, 0.0, 2, [0.3, 0.3], device)
    pend = clusterer.launch(emb, bw, seed.reshape(-1), 1)
    _, meta = clusterer.finish(pend)
    k = len(meta["instance_labels"])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        a.record()
        pend = clusterer.launch(emb, bw, seed.reshape(-1), 1)
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    ms = sorted(times)[len(times) // 2]
    bytes_alg = n * (4 * e + 12) * (k + 1)
    achieved = bytes_alg / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"], "traffic": None, "kernel": "seq_cluster_kernel<4> N=%d K=%d" % (n, k),
            "launch_ms": ms, "algorithmic_mb_per_launch": bytes_alg / 1e6,
            "peak_source": "%s hbm_gbs" % peaks["source"],
            "note": "working set 80 MB < 126 MB L2: part of the traffic is served by L2, so this can exceed the DRAM "
                    "copy peak"}


def run_gpu_arm(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from stemseg_b200 import _lib, decoder
    from stemseg_b200.pipeline import build_davis_pipeline

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    lib = _lib.load()                         # raises if the CUDA library is missing: no fallback
    _lib.check(lib.stemseg_check_device())
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)

    pipe = build_davis_pipeline(device, num_frames=T, precision=args.precision)
    host_feats = {s: f.pin_memory() for s, f in make_features_cpu(seed=rank).items()}
    dev_feats = {s: f.to(device, non_blocking=True) for s, f in host_feats.items()}
    fg_mask = torch.ones((T, H4, W4), dtype=torch.uint8, device=device)
    torch.cuda.synchronize()

    def step_resident():
        return pipe(dev_feats, fg_mask=fg_mask)

    def run_resident(steps):
        """K steps through the public submit()/result() API, software-pipelined one deep: the host work of step i
        (metadata sync, result objects) overlaps the kernels of step i+1.  Every result is complete on return."""
        queue = []
        for _ in range(steps):
            queue.append(pipe.submit(dev_feats, fg_mask=fg_mask))
            if len(queue) > pipe.steps_in_flight:
                queue.pop(0).result()
        for pend in queue:
            pend.result()

    def run_e2e(steps):
        """Same, from pinned host memory: double-buffered H2D of the pyramid, labels copied back to the host."""
        ticket = stager.submit(host_feats)
        queue = []
        for i in range(steps):
            nxt_ticket = stager.submit(host_feats) if i + 1 < steps else None   # prefetch the next clip
            pend = pipe.submit(stager.get(ticket), fg_mask=fg_mask, labels_to_host=True)
            stager.release(ticket, pend.inputs_consumed)
            queue.append(pend)
            if len(queue) > pipe.steps_in_flight:
                assert queue.pop(0).result().labels_host is not None
            ticket = nxt_ticket
        for pend in queue:
            assert pend.result().labels_host.numel() == GRID_POINTS

    from stemseg_b200.pipeline import HostFeatureStream
    stager = HostFeatureStream(device)

    def step_e2e():
        ticket = stager.submit(host_feats)
        pend = pipe.submit(stager.get(ticket), fg_mask=fg_mask, labels_to_host=True)
        stager.release(ticket, pend.inputs_consumed)
        res = pend.result()
        return res.labels_host

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, profile=False, whole=False):
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if profile:
            decoder.PROFILE_EVENTS = []
        start.record()
        if whole:
            fn(steps)
        else:
            for _ in range(steps):
                fn()
        end.record()
        barrier()
        ms = start.elapsed_time(end)
        events = None
        if profile:
            events, decoder.PROFILE_EVENTS = decoder.PROFILE_EVENTS, None
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, events

    for _ in range(args.warmup):
        step_resident()
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _lib.KERNEL_LAUNCHES[0] = 0
    ms_total, _ = timed(run_resident, args.steps, whole=True)
    launches = _lib.KERNEL_LAUNCHES[0]
    ms_e2e, _ = timed(run_e2e, args.steps, whole=True)
    clocks = sampler.stop() if rank == 0 else None          # sampled over both timed regions

    # per-stage breakdown (untimed extra pass on rank 0; informational)
    stages = {}
    if rank == 0:
        def ev_time(fn, reps=3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(reps):
                out = fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps, out
        pipe.run_heads(dev_feats)          # builds / warms the heads-only graph outside the timing
        stages["heads_ms"], (emb, var, seedi, _) = ev_time(lambda: pipe.run_heads(dev_feats))
        stages["gather_cluster_ms"], _ = ev_time(lambda: pipe.cluster(emb, var, seedi, fg_mask))

    # secondary roofline: the clustering kernel in its HBM-bound regime (full-resolution point set, N = 3 317 760)
    cluster_roofline = None
    if rank == 0:
        cluster_roofline = cluster_fullres_roofline(device, load_peaks())

    # per-launch durations of the tcgen05 conv kernel: the same plan launched eagerly (the timed region replays it
    # as a CUDA graph, where individual launches cannot be bracketed), CUDA events on the launching stream
    group = pipe._head_group()
    group.use_graph, pipe.use_step_graph = False, False
    step_resident()
    _, conv_events = timed(step_resident, args.steps, profile=True)
    group.use_graph, pipe.use_step_graph = True, True

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    clips = args.steps * world
    value = clips / (ms_total * 1e-3)
    e2e_value = clips / (ms_e2e * 1e-3)

    # roofline of the dominant kernel: the tcgen05 conv launch with the largest share of the step
    by_shape = {}
    for shape, a, b in conv_events:
        by_shape.setdefault(shape, []).append(a.elapsed_time(b))
    dom_shape, dom_times = max(by_shape.items(), key=lambda kv: sum(kv[1]))
    n, t, h, w, cin, cout, ks, planes = dom_shape
    flops = 2.0 * n * t * h * w * (27 if ks == 3 else 1) * cin * cout
    dom_ms = sum(dom_times) / len(dom_times)
    achieved = flops / (dom_ms * 1e-3) / 1e12
    conv_ms_per_step = sum(sum(v) for v in by_shape.values()) / args.steps
    dom_count_per_step = len(dom_times) / args.steps
    planes_products = 3 if planes == 2 else 1
    traffic, traffic_src = None, None
    try:            # DRAM bytes of the same kernel from the committed `ncu --set full` capture (per launch)
        prof = json.load(open(os.path.join(ROOT, "profiles", "r01_full_summary.json")))
        dom = prof["dominant_conv"]
        if args.precision == "fp32" and "256, 32, 2" in dom["Kernel Name"]:
            traffic = (float(dom["dram__bytes_read.sum"]) + float(dom["dram__bytes_write.sum"])) * 1e6
            traffic_src = "profiles/r01_full_summary.json (dram__bytes_read.sum + dram__bytes_write.sum, bytes/launch)"
    except Exception:
        pass
    roofline = {
        "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
        "frac": achieved / peaks["bf16_tflops_sustained"], "traffic": traffic, "traffic_source": traffic_src,
        "kernel": "conv_tc_kernel %dx%dx%dx%d cin=%d cout=%d k=%d planes=%d" % (n * t, h, w, 1, cin, cout, ks, planes),
        "peak_source": "%s bf16_tflops_sustained (kernel timed inside a long step)" % peaks["source"],
        "algorithmic_gflop_per_launch": flops / 1e9,
        "launch_ms": dom_ms,
        "tensor_pipe_products_per_mac": planes_products,
        "tensor_pipe_frac": achieved * planes_products / peaks["bf16_tflops_sustained"],
        # fp32-parity arithmetic issues 3 bf16 tensor-core products per algorithmic MAC (hi*hi + hi*lo + lo*hi), so the
        # algorithmic rate cannot exceed peak / 3; `frac` above is against the full bf16 peak as the contract asks
        "algorithmic_ceiling_tflops": peaks["bf16_tflops_sustained"] / planes_products,
        "frac_of_algorithmic_ceiling": achieved * planes_products / peaks["bf16_tflops_sustained"],
        "launches_per_step": dom_count_per_step,
        "share_of_step": sum(dom_times) / args.steps / (ms_total / args.steps),
        "all_conv_share_of_step": conv_ms_per_step / (ms_total / args.steps),
        "timing": "CUDA events around every conv launch in an eager pass of the same plan (the timed region replays "
                  "the plan as a CUDA graph)",
    }

    # the incumbent GPU path, for context (opt-in, --incumbent): the same heads as plain torch ops (cuDNN conv3d, native
    # GroupNorm / pool / interpolate -- what the unmodified reference modules run on this GPU), fp32 with and without TF32
    incumbent = None
    if world == 1 and args.incumbent:
        try:
            incumbent = time_incumbent_gpu_heads(device)
        except Exception as exc:                     # informational only: never lose the bench line over it
            incumbent = {"error": "%s: %s" % (type(exc).__name__, exc)}

    # CPU baseline on a bounded sample (rank 0, N == 1 only)
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cfeats = make_features_cpu()
        emb_sd, seed_sd = build_cpu_reference()
        threads = best_cpu_threads(cfeats, emb_sd, seed_sd)
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or (time.perf_counter() - t0 < 10.0 and reps < 20):
            cpu_reference_step(cfeats, emb_sd, seed_sd)
            reps += 1
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": reps / dt, "unit": "clips/s", "cores": threads, "kind": "port",
                        "sample": "%d full clips after warm-up (oracle port: torch-CPU fp32 heads + numpy gather/"
                                  "clustering; fastest of {all, 1/2, 1/4, 16} host threads)" % reps}

    h2d = sum(f.numel() * f.element_size() for f in host_feats.values())
    d2h = GRID_POINTS * 8
    line = {
        "metric": "clips_per_sec", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32" if args.precision == "fp32" else "bf16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "precision": args.precision,
                   "arithmetic": "bf16x2-split operands (hi*hi+hi*lo+lo*hi on tcgen05), fp32 accumulate"
                   if args.precision == "fp32" else "bf16 operands, fp32 accumulate",
                   "l2": "inputs larger than L2 (282 MB pyramid per step vs 126 MB L2)", "clips_per_step_per_gpu": 1,
                   "host_pipelining": "submit()/result(): up to 2 steps in flight (two graph instances on two streams); the "
                                      "result of step i is collected after step i+2 is enqueued",
                   "parallelism": "clip-parallel x%d (no data-path collective)" % world},
        "mvoxels_per_sec": value * VOXELS_PER_CLIP / 1e6,
        "grid_points_per_sec": value * GRID_POINTS,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps,
                "note": "pinned host pyramid, double-buffered H2D on a copy stream (clip i+1 uploads while clip i "
                        "computes), submit()/result() pipelined one deep, labels copied back to pinned host memory "
                        "every step"},
        "gpu_launches": launches,
        "roofline": roofline,
        "roofline_cluster_fullres": cluster_roofline,
        "cpu_baseline": cpu_baseline,
        "incumbent_gpu": incumbent,
        "stages": stages,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------------
# secondary workloads (BASELINE.json configs[2] and configs[3]); the default line is configs[1]
# ------------------------------------------------------------------------------------------------------------------
def run_cfg3(args, rank, local_rank, world):
    """configs[2]: 16x480x864 clip, bf16 decoder (tcgen05, one product per MAC) + clustering; plus the clustering
    kernel alone on 8-dim embeddings with 8 learned variances (N = 414 720 quarter-res, 6 635 520 full-res)."""
    import numpy as np
    import torch
    from stemseg_b200 import _lib
    from stemseg_b200.clusterers import SequentialClustering
    from stemseg_b200.pipeline import build_davis_pipeline
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    _lib.check(_lib.load().stemseg_check_device())
    t16 = 16
    pipe = build_davis_pipeline(device, num_frames=t16, precision="bf16")
    g = torch.Generator().manual_seed(0)
    feats = {s_: torch.randn(1, IN_CH, t16, HP // s_, WP // s_, generator=g).to(device) for s_ in (32, 16, 8, 4)}
    mask = torch.ones((t16, H4, W4), dtype=torch.uint8, device=device)
    for _ in range(max(3, args.warmup)):
        pipe(feats, fg_mask=mask)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    queue = []
    for _ in range(args.steps):
        queue.append(pipe.submit(feats, fg_mask=mask))
        if len(queue) > pipe.steps_in_flight:
            queue.pop(0).result()
    for q in queue:
        q.result()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    flops = 2 * 2 * 564.87e9                     # two heads, SURVEY §8d: 564.87 GMAC per head at 16x480x864
    extra = {}
    rng = np.random.default_rng(0)
    for n in (t16 * H4 * W4, t16 * HP * WP):
        centres = rng.uniform(-1, 1, size=(24, 8)).astype(np.float32)
        which = rng.integers(0, 24, size=n)
        emb = torch.from_numpy(centres[which] + 0.05 * rng.standard_normal((n, 8)).astype(np.float32)).to(device)
        bw = torch.from_numpy(np.exp(rng.uniform(-1, 1, size=(n, 8))).astype(np.float32) * 10).to(device)
        seed = torch.rand(n, device=device)
        clu = SequentialClustering(0.5, 0.3, 0.0, 0, [], device)
        pend = clu.launch(emb, bw, seed, 1)
        _, meta = clu.finish(pend)
        k = len(meta["instance_labels"])
        times = []
        for _ in range(5):
            a.record()
            clu.launch(emb, bw, seed, 1)
            b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
        t_ms = sorted(times)[2]
        byts = n * (4 * 8 + 12) * (k + 1)
        extra["cluster_e8_n%d" % n] = {"ms": t_ms, "clusters": k, "algorithmic_gb_per_s": byts / t_ms / 1e6,
                                       "frac_of_hbm_peak": byts / t_ms / 1e6 / load_peaks()["hbm_gbs"]}
    print(json.dumps({"metric": "clips_per_sec", "value": 1e3 / ms, "unit": "clips/s", "n_gpus": 1, "steps": args.steps,
                      "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                      "config": {"workload": "configs[2]: 16x480x854 clip (pad 480x864), DAVIS heads in bf16 + fg gather + "
                                             "clustering over 414720 points; clustering alone on E=8"},
                      "algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
                      "frac_of_sustained_bf16_peak": flops / (ms * 1e-3) / 1e12 / load_peaks()["bf16_tflops_sustained"],
                      **extra}), flush=True)


def run_video64(args, rank, local_rank, world):
    """configs[3]: a 64-frame sequence = 8 overlapping 16-frame sub-clips (get_subsequence_frames(64, 16, overlap 9)),
    clip-parallel: sub-clip i on rank i % world, NCCL all-gather of the label vectors, sequential stitch on every rank."""
    import torch
    import torch.distributed as dist
    from stemseg_b200 import _lib
    from stemseg_b200.chaining import get_subsequence_frames
    from stemseg_b200.parallel import clip_parallel_process
    from stemseg_b200.pipeline import build_davis_pipeline
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    _lib.check(_lib.load().stemseg_check_device())
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    t16 = 16
    windows, _ = get_subsequence_frames(64, t16, "davis", 9)
    pipe = build_davis_pipeline(device, num_frames=t16, min_seediness_prob=0.0)
    cache = {}

    def features_for_clip(i):
        if i not in cache:
            g = torch.Generator().manual_seed(1000 + i)
            cache[i] = {s_: torch.randn(1, IN_CH, t16, HP // s_, WP // s_, generator=g).to(device) for s_ in (32, 16, 8, 4)}
        return cache[i]

    for i in range(len(windows)):
        if i % world == rank:
            features_for_clip(i)
    masks = torch.ones((64, H4, W4), dtype=torch.uint8, device=device)

    def one_video():
        container, _, _ = clip_parallel_process(pipe, masks, windows, features_for_clip)
        return container

    for _ in range(2):
        one_video()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps = max(2, args.steps // 5)
    t0 = time.perf_counter()
    for _ in range(reps):
        container = one_video()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = (time.perf_counter() - t0) / reps
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    if rank == 0:
        labels, counts, lifetimes = container.get_track_mask_idxes()
        print(json.dumps({"metric": "subclips_per_sec", "value": len(windows) / dt, "unit": "sub-clips/s",
                          "n_gpus": world, "steps": reps, "warmup": 2, "ms_per_step": dt * 1e3,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic",
                          "config": {"workload": "configs[3]: 64-frame video, 8 sub-clips of 16x480x864 (overlap 9), "
                                                 "clip-parallel heads+gather+clustering, all-gather of labels, stitch",
                                     "timing": "host wall clock around the whole video (includes the exchange and the "
                                               "host-side stitch), max over ranks"},
                          "videos_per_sec": 1.0 / dt, "tracks": len([k for k in counts if k >= 0])}), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------------
# configs[4]: data-parallel training step of the heads (one 8x384x640 clip per GPU per step)
# ------------------------------------------------------------------------------------------------------------------
TRAIN_T, TRAIN_H, TRAIN_W = 8, 384, 640
TRAIN_WORKLOAD = "configs[4]: DDP training step, one 8x384x640 clip per GPU: embedding + seediness heads " \
                 "([256,256,128,128]) forward, embedding loss (Lovasz + seediness + variance smoothness, 3 instances), " \
                 "backward to features and parameters, gradient all-reduce, SGD-Nesterov"
TRAIN_GFLOP_PER_CLIP = 3 * 2 * 334.7          # fwd + dgrad + wgrad of two heads (334.7 GFLOP / head)


def make_train_inputs(seed):
    """Synthetic FPN pyramid + targets (three moving ellipses) at the embedding resolution."""
    import torch
    g = torch.Generator().manual_seed(seed)
    h4, w4 = TRAIN_H // 4, TRAIN_W // 4
    feats = [torch.randn(1, IN_CH, TRAIN_T, TRAIN_H // s, TRAIN_W // s, generator=g) for s in (32, 16, 8, 4)]
    yy, xx = torch.meshgrid(torch.arange(h4, dtype=torch.float32), torch.arange(w4, dtype=torch.float32), indexing="ij")
    masks = torch.zeros(3, TRAIN_T, h4, w4, dtype=torch.uint8)
    taken = torch.zeros(TRAIN_T, h4, w4, dtype=torch.bool)
    for i, (cy, cx, ry, rx, vy, vx) in enumerate(((0.3, 0.3, 0.15, 0.12, 0.01, 0.02), (0.6, 0.65, 0.2, 0.15, -0.01, 0.01),
                                                  (0.75, 0.25, 0.1, 0.1, 0.0, -0.015))):
        for f in range(TRAIN_T):
            m = ((yy - (cy + vy * f) * h4) / (ry * h4)) ** 2 + ((xx - (cx + vx * f) * w4) / (rx * w4)) ** 2 <= 1.0
            m &= ~taken[f]
            masks[i, f] = m
            taken[f] |= m
    ignore = torch.rand(TRAIN_T, h4, w4, generator=g) < 0.02
    return feats, masks, ignore


def train_reference_step(state, feats, masks, ignore):
    """The reference's CPU training step for the heads, restated with the oracles (torch-CPU fp32 autograd through
    the same ATen ops as the reference modules + torch.optim.SGD)."""
    import torch
    from oracle import decoder_oracle as do
    from oracle import loss_oracle as lo
    emb_sd, seed_sd, opt = state
    opt.zero_grad()
    f = [x.clone().requires_grad_(True) for x in feats]
    out = torch.cat((do.embedding_head(emb_sd, f, TRAIN_T, 4, "xyff", True, False),
                     do.seediness_head(seed_sd, f, TRAIN_T)), dim=1)
    losses = lo.loss_from_head_output(out, masks, ignore, 4, 2, [0.3, 0.3], w_lovasz=1.0, w_variance_smoothness=10.0,
                                      w_seediness=1.0, w=1.0)
    losses["total"].backward()
    opt.step()
    return float(losses["total"])


def build_train_reference():
    import torch
    emb_sd, seed_sd = build_cpu_reference()
    emb_sd = {k: (v.clone().requires_grad_(True) if v.dim() > 0 else v) for k, v in emb_sd.items()}
    seed_sd = {k: v.clone().requires_grad_(True) for k, v in seed_sd.items()}
    params = [v for v in list(emb_sd.values()) + list(seed_sd.values()) if v.requires_grad]
    opt = torch.optim.SGD(params, 1e-3, 0.9, weight_decay=1e-4, nesterov=True)
    return emb_sd, seed_sd, opt


def run_train_reference(args, rank, world):
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    feats, masks, ignore = make_train_inputs(0)
    state = build_train_reference()
    for _ in range(max(1, min(args.warmup, 1))):
        train_reference_step(state, feats, masks, ignore)
    steps, t0 = 0, time.perf_counter()
    while steps < args.steps and (steps < 2 or time.perf_counter() - t0 < REFERENCE_BUDGET_S):
        train_reference_step(state, feats, masks, ignore)
        steps += 1
    dt = time.perf_counter() - t0
    value = steps / dt
    print(json.dumps({
        "impl": "reference", "metric": "train_clips_per_sec", "value": value, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": 1, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": TRAIN_WORKLOAD, "timing": "host wall clock"},
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "%d full training steps (oracle port: torch-CPU fp32 autograd + SGD)" % steps},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def run_train(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import torch.nn as nn
    from stemseg_b200 import _lib, decoder, heads
    from stemseg_b200.losses import EmbeddingLoss
    from stemseg_b200.training import DecoderTrainer
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    _lib.check(_lib.load().stemseg_check_device())
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    torch.manual_seed(42)
    norm = lambda c: nn.GroupNorm(32, c)       # noqa: E731
    emb = heads.EmbeddingHead(IN_CH, list(INTER), 4, True, False, "xyff", NormType=norm, num_frames=TRAIN_T,
                              precision=args.precision).to(device)
    seedh = heads.SeedinessHead(IN_CH, list(INTER), NormType=norm, num_frames=TRAIN_T,
                                precision=args.precision).to(device)
    crit = EmbeddingLoss(4, embedding_size=4, nbr_free_dims=2, free_dim_stds=[0.3, 0.3], weight_variance_smoothness=10.0,
                         weight_lovasz=1.0, weight_regularization=0.001, weight_seediness=1.0, weight=1.0)
    trainer = DecoderTrainer({"embedding": emb, "seediness": seedh}, crit, overlap_heads=not args.no_overlap_heads)
    feats_cpu, masks, ignore = make_train_inputs(rank)
    host_feats = [f.pin_memory() for f in feats_cpu]
    dev_feats = [f.to(device).requires_grad_(True) for f in host_feats]
    targets = [{"masks": masks.to(device), "ignore_masks": ignore.to(device)}]
    host_targets = [{"masks": masks.pin_memory(), "ignore_masks": ignore.to(torch.uint8).pin_memory()}]

    def step_resident():
        for f in dev_feats:
            f.grad = None
        return trainer.step(dev_feats, targets)

    from stemseg_b200.pipeline import HostFeatureStream
    stager = HostFeatureStream(device)
    host_clip = {32: host_feats[0], 16: host_feats[1], 8: host_feats[2], 4: host_feats[3],
                 "masks": host_targets[0]["masks"], "ignore": host_targets[0]["ignore_masks"]}

    def run_e2e(steps):
        """Pinned host pyramid + targets, double-buffered: the copies of clip i+1 run on the copy stream while clip i
        trains; the loss is read back (D2H, synchronising) every step."""
        ticket = stager.submit(host_clip)
        for i in range(steps):
            nxt = stager.submit(host_clip) if i + 1 < steps else None
            buf = stager.get(ticket)
            out = trainer.step([buf[32], buf[16], buf[8], buf[4]],
                               [{"masks": buf["masks"], "ignore_masks": buf["ignore"]}])
            stager.release(ticket)
            loss = float(out["optimization_losses"]["embedding_loss"].detach())
            assert loss == loss
            ticket = nxt

    def step_e2e():
        run_e2e(1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(args.warmup):
        step_resident()
    step_e2e()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _lib.KERNEL_LAUNCHES[0] = 0
    ms_total = timed(step_resident, args.steps)
    launches = _lib.KERNEL_LAUNCHES[0]
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(lambda: run_e2e(args.steps), 1)

    # the same step driven through torch autograd (what the reference's training loop does with the B200 heads and
    # loss installed): ~500 launches issued from Python instead of 3 graph replays.  Single-GPU runs only (the
    # autograd path issues its own collectives).
    phases = {}
    if world == 1:
        trainer.use_graph = False
        for _ in range(2):
            step_resident()
        phases["autograd_mode_ms_per_step"] = timed(step_resident, 3) / 3
        trainer.use_graph = True
        phases["graph_mode_ms_per_step"] = ms_total / args.steps
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    clips = args.steps * world
    value = clips / (ms_total * 1e-3)
    ms_step = ms_total / args.steps
    achieved = TRAIN_GFLOP_PER_CLIP / ms_step            # GFLOP / ms = TFLOP/s, per GPU
    products = 3 if args.precision == "fp32" else 1
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        state = build_train_reference()
        train_reference_step(state, feats_cpu, masks, ignore)
        reps, t0 = 0, time.perf_counter()
        while reps < 2 or (time.perf_counter() - t0 < 10.0 and reps < 10):
            train_reference_step(state, feats_cpu, masks, ignore)
            reps += 1
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": reps / dt, "unit": "clips/s", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": "%d full training steps after warm-up (oracle port: torch-CPU fp32 autograd + SGD)" % reps}
    h2d = sum(f.numel() * 4 for f in host_feats) + masks.numel() + ignore.numel()
    print(json.dumps({
        "metric": "train_clips_per_sec", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
        "config": {"workload": TRAIN_WORKLOAD, "precision": args.precision, "global_batch": world,
                   "l2": "inputs + saved activations larger than L2 (pyramid 167 MB, saved conv outputs > 1 GB per step)",
                   "parallelism": "dp%d: per-head flat gradient all-reduce (NCCL, async, overlapped with the other "
                                  "head's backward), mean folded into the fused SGD pass" % world},
        "mvoxels_per_sec": value * TRAIN_T * TRAIN_H * TRAIN_W / 1e6,
        "clocks": clocks,
        "e2e": {"value": clips / (ms_e2e * 1e-3), "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps,
                "note": "pinned host pyramid + targets copied in every step (double-buffered on a copy stream), loss "
                        "read back every step"},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_tflops_sustained"], "traffic": None,
                     "kernel": "whole training step (conv_tc_kernel forward + dgrad + wgrad dominate)",
                     "algorithmic_gflop_per_step": TRAIN_GFLOP_PER_CLIP, "tensor_pipe_products_per_mac": products,
                     "tensor_pipe_frac": achieved * products / peaks["bf16_tflops_sustained"],
                     "peak_source": "%s bf16_tflops_sustained" % peaks["source"]},
        "cpu_baseline": cpu_baseline, "phases": phases}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap-heads", action="store_true",
                    help="train workload: three graphs with per-head all-reduce overlap instead of two concurrent heads")
    ap.add_argument("--incumbent", action="store_true",
                    help="also time the same heads as plain torch/cuDNN ops on the GPU (informational; off by default: "
                         "it runs the oracle's functional heads, which the default arm must not touch)")
    ap.add_argument("--no-incumbent", action="store_true", help=argparse.SUPPRESS)      # accepted for old command lines
    ap.add_argument("--workload", default="davis480p", choices=["davis480p", "cfg3", "video64", "train"],
                    help="davis480p = BASELINE configs[1] (the contract line); cfg3 / video64 / train = configs[2] / "
                         "configs[3] / configs[4]")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if args.workload == "train":
            return run_train_reference(args, rank, world)
        run_reference_arm(args, rank, world)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node %d "
                             "--master-addr 127.0.0.1 --master-port 29500 bench.py --gpus %d ..." % (args.gpus, args.gpus))
    if args.workload == "cfg3":
        return run_cfg3(args, rank, local_rank, world)
    if args.workload == "video64":
        return run_video64(args, rank, local_rank, world)
    if args.workload == "train":
        return run_train(args, rank, local_rank, world)
    run_gpu_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
