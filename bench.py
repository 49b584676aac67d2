#!/usr/bin/env python
"""bench.py — deformable-aggregation fwd+bwd throughput on B200 (BASELINE.json metric).

One "step" = one pass of the hot path over one batch of synthetic input: every deformable
aggregation call of a HiP-AD stage-2 decoder (6 layers x {det 900x13, map 100x300, plan 480x90,
ego 1x13}; 6 cameras, 4 FPN levels of a 352x640 input, 256 channels, 8 groups), forward AND
backward, on `--bs` samples per GPU.  N GPUs = N batch shards, no data-path collective (weak
scaling).

  value     algorithmic GB/s (SURVEY.md §8d byte counts) with every input already in HBM, the
            step issued through the C ABI and replayed as one CUDA graph
  e2e       same step through the public Python API (hipad_b200.deformable_aggregation_function
            + autograd) from pinned HOST buffers, H2D/D2H inside the timed region
  roofline  dominant kernel: algorithmic bytes per launch / CUDA-event duration vs measured HBM peak
  cpu_baseline / --impl reference: the reference's own CPU path (torch grid_sample branch,
            oracle/torch_path.py restatement) on the host cores, bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

# stdout carries exactly one JSON line (rank 0).  Libraries write there too (NCCL prints its version banner on
# stdout under the image's NCCL_DEBUG setting), so file descriptor 1 is pointed at stderr for the whole run and the
# JSON line goes to the saved descriptor.
_REAL_STDOUT = None


def isolate_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    text = json.dumps(line) + "\n"
    if _REAL_STDOUT is None:
        sys.stdout.write(text)
    else:
        os.write(_REAL_STDOUT, text.encode())

MODALITIES = (("det", 900, 13), ("map", 100, 300), ("plan", 480, 90), ("ego", 1, 13))
LAYERS = 6
FINAL_HW = (352, 640)
C, G, CAMS = 256, 8, 6
METRIC = "deform-agg fwd+bwd GB/s"


def level_hw(final_hw):
    return [(final_hw[0] // s, final_hw[1] // s) for s in (4, 8, 16, 32)]


# ------------------------------------------------------------------------------------- inputs
def make_calls(bs, seed, layers=LAYERS, modalities=MODALITIES):
    """numpy inputs of every DFA call of one step (locations/weights differ per layer and modality)."""
    import helpers as H
    calls = []
    for layer in range(layers):
        for mi, (kind, A, P) in enumerate(modalities):
            geo_kind = "det" if kind == "ego" else kind
            c = H.make_geo_case(seed * 1000 + layer * 10 + mi, geo_kind, bs, level_hw(FINAL_HW), FINAL_HW,
                                C=C, G=G, A=A, P=P, with_feat=False)
            if kind == "ego":
                # ego key points sit inside the ego box: visible to no camera (measured 0.000)
                c["loc"] = np.full_like(c["loc"], -1.0) + 0.1 * np.random.default_rng(seed + layer).random(
                    c["loc"].shape, dtype=np.float32)
            rng = np.random.default_rng(seed * 77 + layer * 10 + mi)
            calls.append(dict(kind=kind, layer=layer, A=A, P=P, loc=c["loc"], weights=c["weights"],
                              grad_out=rng.standard_normal((bs, A, C), dtype=np.float32),
                              key_points=c["key_points"].astype(np.float32), logits=c["logits"],
                              projection_mat=c["projection_mat"].astype(np.float32), image_wh=c["image_wh"]))
    import helpers
    shapes, starts, F = helpers.level_tables(level_hw(FINAL_HW), CAMS)
    return calls, shapes, starts, F


def algorithmic_bytes(ops, shapes_d, starts_d, loc_d, bs, F, A, P, L, elem_bytes):
    """B_fwd, B_bwd with U = distinct feature rows touched (counted on the GPU from
    the kernels' own integer indices)."""
    idx = ops.sample_indices(shapes_d, starts_d, loc_d)            # [bs,A,P,cams,L,6]
    valid = idx[..., 0] > 0
    row0 = idx[..., 5].long()
    w = shapes_d[:, :, 1].long()[None, None, None]                 # [1,1,1,cams,L]
    mask = idx[..., 4]
    rows = []
    for k, off in enumerate((0, 1, w, w + 1)):
        ok = valid & ((mask >> k) & 1).bool()
        r = row0 + off + (torch.arange(bs, device=idx.device) * F)[:, None, None, None, None]
        rows.append(r[ok])
    U = int(torch.unique(torch.cat(rows)).numel()) if rows else 0
    n_valid = int(valid[..., 0].sum())
    loc_b = bs * A * P * CAMS * 2 * 4
    w_b = bs * A * P * CAMS * L * G * 4
    out_b = bs * A * C * 4
    feat_b = U * C * elem_bytes
    b_fwd = loc_b + w_b + feat_b + out_b
    b_bwd = out_b + loc_b + w_b + feat_b + w_b + loc_b + bs * F * C * elem_bytes
    return dict(U=U, n_valid=n_valid, fwd=b_fwd, bwd=b_bwd, dense_gfeat=bs * F * C * elem_bytes,
                bwd_sample=b_bwd - bs * F * C * elem_bytes)


# ------------------------------------------------------------------------------------- multi-GPU
def shard_seed(rank):
    """Batch sharding: rank r owns its own `bs` samples (seeded by rank); the ranks share nothing and
    the data path has no collective.  Only the timing is reduced (MAX) across ranks."""
    return rank


def max_over_ranks(value, device):
    """MAX-reduce a python float across the process group (identity when not initialised)."""
    import torch.distributed as dist
    t = torch.tensor([value], device=device, dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_gbs(world, bytes_per_rank_step, ms_per_step):
    """Aggregate throughput of the job: every rank moves `bytes_per_rank_step` per step (weak scaling)."""
    return world * bytes_per_rank_step / (ms_per_step * 1e-3) / 1e9


# ------------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------- GPU arm
WORKLOAD = ("HiP-AD stage-2 decoder DFA path: 6 layers x (det 900x13 + map 100x300 + plan 480x90 + ego 1x13), "
            "6 cams, 4 levels of 352x640, C=256, G=8, fwd+bwd")


def shared_config(bs, dtype):
    """`config` of BOTH arms (the driver compares them): what the workload is, nothing about how it is run."""
    return {"workload": WORKLOAD, "bs_per_gpu": bs, "feature_dtype": dtype}


def gpu_arm(args):
    import ctypes
    import torch.distributed as dist
    import hipad_b200
    from hipad_b200 import _lib
    ops = hipad_b200.ops
    lib = _lib.get()   # raises if the CUDA library is missing: there is no fallback

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (use --impl reference for the CPU arm)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    bf16 = args.dtype == "bf16"
    elem = 2 if bf16 else 4
    bs = args.bs
    L = 4

    calls, shapes, starts, F = make_calls(bs, seed=shard_seed(rank))
    rng = np.random.default_rng(1234 + shard_seed(rank))
    feat_h = torch.from_numpy(rng.standard_normal((bs, F, C), dtype=np.float32))
    if bf16:
        feat_h = feat_h.bfloat16()
    feat = feat_h.to(dev)
    shapes_d = torch.from_numpy(shapes).to(dev)
    starts_d = torch.from_numpy(starts).to(dev)
    dims = lambda c: (bs, CAMS, F, C, L, c["A"], c["P"], G)
    n_mod = len(MODALITIES)
    n_lanes = 1 if args.streams <= 1 else n_mod
    lane_of = {m[0]: (i % n_lanes) for i, m in enumerate(MODALITIES)}
    total_fwd = total_bwd = 0
    for c in calls:
        c["loc_d"] = torch.from_numpy(c["loc"]).to(dev)
        c["w_d"] = torch.from_numpy(c["weights"]).to(dev)
        c["go_d"] = torch.from_numpy(c["grad_out"]).to(dev)
        c["out_d"] = torch.empty((bs, c["A"], C), dtype=torch.float32, device=dev)
        c["g_loc_d"] = torch.empty_like(c["loc_d"])
        c["g_w_d"] = torch.empty_like(c["w_d"])
        c["bytes"] = algorithmic_bytes(ops, shapes_d, starts_d, c["loc_d"], bs, F, c["A"], c["P"], L, elem)
        total_fwd += c["bytes"]["fwd"]
        total_bwd += c["bytes"]["bwd"]
    step_bytes = total_fwd + total_bwd
    dense = bs * F * C * elem
    layers = [calls[i:i + n_mod] for i in range(0, len(calls), n_mod)]

    def table(cs, bwd):
        t = _lib.call_table([(c["loc_d"].data_ptr(), c["w_d"].data_ptr(), c["g_loc_d"].data_ptr() if bwd else None,
                              c["g_w_d"].data_ptr() if bwd else None, c["A"], c["P"]) for c in cs])
        return t, ctypes.cast(t, ctypes.c_void_p)

    def fwd_ws_bytes(cs):
        _, tp = table(cs, False)
        return max(256, lib.hipad_dfa_group_forward_workspace_bytes(tp, len(cs), bs, CAMS, C))

    def bwd_ws_bytes(cs):
        _, tp = table(cs, True)
        return max(256, lib.hipad_dfa_group_backward_workspace_bytes(tp, len(cs), bs, CAMS, F, C, L, G))

    # ---- per-call step (the reference's own contract: 24 calls, each with its own dense feature gradient).  One
    # forward / backward workspace and one dense g_feat per MODALITY (reused across layers, written in full by every
    # call): the four calls of a decoder layer are independent, so with --streams 4 they run as parallel graph branches.
    ws_f = [torch.empty(max(fwd_ws_bytes([c]) for c in calls), dtype=torch.uint8, device=dev) for _ in range(n_lanes)]
    ws_b = [torch.empty(max(bwd_ws_bytes([c]) for c in calls), dtype=torch.uint8, device=dev) for _ in range(n_lanes)]
    g_feat_l = [torch.empty_like(feat) for _ in range(n_lanes)]

    def fwd_call(c, stream):
        lane = lane_of[c["kind"]]
        _, tp = table([c], False)
        _lib.check(lib.hipad_dfa_group_forward(int(bf16), c["out_d"].data_ptr(), feat.data_ptr(), shapes_d.data_ptr(),
                                               starts_d.data_ptr(), tp, 1, bs, CAMS, F, C, L, G, ws_f[lane].data_ptr(),
                                               ws_f[lane].numel(), stream), "forward")

    def bwd_call(c, stream, mask=7):
        lane = lane_of[c["kind"]]
        _lib.check(lib.hipad_dfa_backward_stages(
            1 if bf16 else 0, mask, feat.data_ptr(), shapes_d.data_ptr(), starts_d.data_ptr(),
            c["loc_d"].data_ptr(), c["w_d"].data_ptr(), c["go_d"].data_ptr(), g_feat_l[lane].data_ptr(),
            c["g_loc_d"].data_ptr(), c["g_w_d"].data_ptr(), *dims(c), ws_b[lane].data_ptr(), ws_b[lane].numel(), stream),
            "backward")

    side = [torch.cuda.Stream(device=dev) for _ in range(n_lanes - 1)]

    def layer_group(group, fn, main):
        # the modality calls of one decoder layer: serial on `main`, or one per stream (fork / join with events,
        # which CUDA-graph capture turns into parallel branches)
        if n_lanes == 1:
            for c in group:
                fn(c, main.cuda_stream)
            return
        fork = torch.cuda.Event()
        fork.record(main)
        joins = []
        for c in group:
            lane = lane_of[c["kind"]]
            st = main if lane == 0 else side[lane - 1]
            if st is not main:
                st.wait_event(fork)
            fn(c, st.cuda_stream)
            if st is not main:
                ev = torch.cuda.Event()
                ev.record(st)
                joins.append(ev)
        for ev in joins:
            main.wait_event(ev)

    def issue_step(main):
        for group in layers:                    # decoder forward: 6 layers x 4 modalities
            layer_group(group, fwd_call, main)
        for group in reversed(layers):          # autograd order
            layer_group(list(reversed(group)), bwd_call, main)

    # ---- grouped step ("next" row f1): ONE launch per decoder layer forward, ONE backward chain per layer, ONE dense
    # fp32 feature gradient for the whole step (first layer of the backward writes it, the others accumulate)
    a_total = sum(c["A"] for c in layers[0])
    for li, group in enumerate(layers):
        out_p = torch.empty((bs, a_total, C), dtype=torch.float32, device=dev)
        go_p = torch.cat([c["go_d"] for c in group], dim=1).contiguous()
        for c in group:
            c["out_p"], c["go_p"] = out_p, go_p
    ws_fg = torch.empty(fwd_ws_bytes(layers[0]), dtype=torch.uint8, device=dev)
    ws_bg = torch.empty(bwd_ws_bytes(layers[0]), dtype=torch.uint8, device=dev)
    g_feat_step = torch.empty((bs, F, C), dtype=torch.float32, device=dev)

    def group_fwd(group, stream):
        _, tp = table(group, False)
        _lib.check(lib.hipad_dfa_group_forward(int(bf16), group[0]["out_p"].data_ptr(), feat.data_ptr(), shapes_d.data_ptr(),
                                               starts_d.data_ptr(), tp, len(group), bs, CAMS, F, C, L, G, ws_fg.data_ptr(),
                                               ws_fg.numel(), stream), "group forward")

    def group_bwd(group, stream, flags, mask=7):
        _, tp = table(group, True)
        _lib.check(lib.hipad_dfa_group_backward_stages(int(bf16), flags, mask, feat.data_ptr(), shapes_d.data_ptr(),
                                                       starts_d.data_ptr(), tp, len(group), group[0]["go_p"].data_ptr(),
                                                       g_feat_step.data_ptr(), bs, CAMS, F, C, L, G, ws_bg.data_ptr(),
                                                       ws_bg.numel(), stream), "group backward")

    def issue_group_step(main):
        for group in layers:
            group_fwd(group, main.cuda_stream)
        for i, group in enumerate(reversed(layers)):
            group_bwd(group, main.cuda_stream, 2 if i == 0 else 3)

    # launches per step: per call 1 forward (+1 ticket memset when rows are sliced: map, plan) and 5 backward kernels
    # (+1 zero-fill kernel when the grid is too small to fold the fill in: the ego call)
    def sliced(c):
        return c["P"] * CAMS > 96
    launches_per_step = sum(1 + (1 if sliced(c) else 0) + 5 + (1 if bs * c["A"] * max(1, -(-c["P"] * CAMS // 96)) < 2 * 148 else 0)
                            for c in calls)
    launches_group_step = len(layers) * (2 + 5)
    flush = torch.zeros(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def flush_l2():
        # read-only sweep of 256 MiB: evicts everything, leaves only clean lines behind
        return flush.view(torch.int64).sum()

    stream = torch.cuda.Stream(device=dev)

    def capture(issue):
        with torch.cuda.stream(stream):
            issue(stream)
            torch.cuda.synchronize()
            if args.no_graph:
                return None
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=stream):
                    issue(torch.cuda.current_stream())
                return g
            except Exception as e:  # keep measuring, eagerly
                print("graph capture failed, timing eager launches:", e, file=sys.stderr)
                return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(graph, issue, steps, warmup, sample_clocks=False):
        with torch.cuda.stream(stream):
            def run():
                if graph is not None:
                    graph.replay()
                else:
                    issue(torch.cuda.current_stream())
            for _ in range(max(warmup, 3)):
                flush_l2()
                run()
            barrier()
            sampler = ClockSampler(local_rank) if (sample_clocks and rank == 0) else None
            if sampler:
                sampler.start()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            barrier()
            for a, b in ev:
                flush_l2()                  # L2 flush between timed iterations (outside the timed interval)
                a.record()
                run()
                b.record()
            barrier()
            total = float(sum(a.elapsed_time(b) for a, b in ev))
            clocks = sampler.stop() if sampler else None
        return max_over_ranks(total, dev) / steps, clocks

    graph = capture(issue_step)
    torch.cuda.synchronize()
    ms_per_step, clocks = timed_steps(graph, issue_step, args.steps, args.warmup, sample_clocks=True)
    value = whole_job_gbs(world, step_bytes, ms_per_step)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"

    # ---- grouped step
    group_step = None
    try:
        g2 = capture(issue_group_step)
        ms2, _ = timed_steps(g2, issue_group_step, args.steps, 3)
        # bytes the grouped step has to move: everything of the per-call step except 23 of the 24 dense g_feat writes
        # (one fp32 buffer for the step; the read-modify-write of touched rows by later layers is not counted)
        bytes2 = step_bytes - len(calls) * dense + bs * F * C * 4
        group_step = {"ms_per_step": round(ms2, 4), "samples_per_s": round(world * bs / (ms2 * 1e-3), 1),
                      "algorithmic_bytes_per_step": int(bytes2), "value": round(whole_job_gbs(world, bytes2, ms2), 2),
                      "unit": "GB/s", "frac_of_peak": round(whole_job_gbs(1, bytes2, ms2) / peak, 4),
                      "gpu_launches_per_step": launches_group_step,
                      "speedup_vs_per_call_step": round(ms_per_step / ms2, 3),
                      "note": "hipad_dfa_group_forward / _backward: one launch per decoder layer forward, one backward chain per "
                              "layer, ONE dense fp32 feature gradient per step (zero-filled once, later layers accumulate); one "
                              "stream, one CUDA graph; same outputs and gradients as the 24 calls (sum of their g_feat)"}
    except Exception as e:
        group_step = {"error": str(e)[:200]}

    # ---- BASELINE configs[1], kernel-only view: every DFA call of a decoder forward through the fused kernel
    # (key-point projection + group softmax + aggregation in one launch, raw weights_fc logits in)
    inference = None
    if not args.no_graph:
        try:
            fused_fn = lib.hipad_dfa_fused_forward_bf16 if bf16 else lib.hipad_dfa_fused_forward_f32
            for c in calls:
                c["kp_d"] = torch.from_numpy(np.ascontiguousarray(c["key_points"])).to(dev)
                c["lg_d"] = torch.from_numpy(np.ascontiguousarray(c["logits"])).to(dev)
                c["pm_d"] = torch.from_numpy(np.ascontiguousarray(c["projection_mat"])).to(dev)
                c["wh_d"] = torch.from_numpy(np.ascontiguousarray(c["image_wh"])).to(dev)

            def fused_call(c, s_):
                _lib.check(fused_fn(c["out_d"].data_ptr(), feat.data_ptr(), shapes_d.data_ptr(), starts_d.data_ptr(),
                                    c["kp_d"].data_ptr(), c["pm_d"].data_ptr(), c["wh_d"].data_ptr(), c["lg_d"].data_ptr(),
                                    None, *dims(c), s_), "fused forward")

            def issue_inference(main):
                for group in layers:
                    layer_group(group, fused_call, main)

            g3 = capture(issue_inference)
            ms3, _ = timed_steps(g3, issue_inference, args.steps, 3)
            inference = {"ms_per_forward": round(ms3, 4), "dfa_only_samples_per_s": round(world * bs / (ms3 * 1e-3), 1),
                         "note": "the 24 DFA calls of one stage-2 decoder forward through the fused kernel (projection + softmax + "
                                 "aggregation per launch), one CUDA graph, L2 flushed between replays; the decoder itself is "
                                 "measured in `decoder_forward`"}
        except Exception as e:
            inference = {"error": str(e)[:200]}

    # ---- per-kernel timing (CUDA events on the launching stream), for the roofline objects
    names = ["forward sample kernel", "backward sample kernel + zero fill", "compaction + band sort", "classify + reduce"]
    kern = {k: [] for k in names}
    kbytes = {k: [] for k in names}
    per_mod = {}
    with torch.cuda.stream(stream):
        s = stream.cuda_stream
        for rep in range(3):
            flush_l2()
            for c in calls:
                e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
                e[0].record(); fwd_call(c, s); e[1].record()
                bwd_call(c, s, 1); e[2].record()
                bwd_call(c, s, 2); e[3].record()
                bwd_call(c, s, 4); e[4].record()
                c["_ev"] = e
            stream.synchronize()
            if rep == 0:
                continue
            for c in calls:
                e = c["_ev"]
                d = [e[i].elapsed_time(e[i + 1]) * 1e3 for i in range(4)]   # microseconds
                # algorithmic bytes per stage (a partition of B_fwd + B_bwd): forward B_fwd; backward sample kernel
                # everything of B_bwd except the touched rows of the dense g_feat write (it zero-fills the rest);
                # the reduce writes those touched rows; the sort only re-reads locations (no algorithmic bytes)
                touched = c["bytes"]["U"] * C * elem
                for name, dur, nb in zip(names, d, (c["bytes"]["fwd"], c["bytes"]["bwd"] - touched, 0, touched)):
                    kern[name].append(dur)
                    kbytes[name].append(nb)
                m = per_mod.setdefault(c["kind"], dict(fwd_us=[], bwd_us=[], bytes=c["bytes"]))
                m["fwd_us"].append(d[0]); m["bwd_us"].append(d[1] + d[2] + d[3])
        # the grouped layer, stage by stage
        gl = {"forward (1 launch)": [], "backward sample kernel + zero fill": [], "compaction + band sort": [],
              "classify + reduce": []}
        for rep in range(3):
            for group in layers[:2]:
                flush_l2()
                e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
                e[0].record(); group_fwd(group, s); e[1].record()
                group_bwd(group, s, 2, 1); e[2].record()
                group_bwd(group, s, 2, 2); e[3].record()
                group_bwd(group, s, 2, 4); e[4].record()
                stream.synchronize()
                if rep:
                    for k, i in zip(gl, range(4)):
                        gl[k].append(e[i].elapsed_time(e[i + 1]) * 1e3)
    share = {k: float(np.sum(v)) for k, v in kern.items()}
    per_kernel = []
    for k in names:
        us = float(np.mean(kern[k])); nb = float(np.mean(kbytes[k]))
        per_kernel.append({"kernel": k, "avg_launch_us": round(us, 2), "algorithmic_bytes_per_launch": int(nb),
                           "achieved_gbs": round(nb / (us * 1e-6) / 1e9, 1), "frac": round(nb / (us * 1e-6) / 1e9 / peak, 4),
                           "share_of_step": round(share[k] / max(sum(share.values()), 1e-9), 3)})
    dominant = max((k for k in share if np.sum(kbytes[k]) > 0), key=share.get)
    dom = next(p_ for p_ in per_kernel if p_["kernel"] == dominant)
    traffic = None
    try:   # dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/ncu_traffic.py), if any
        traffic = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json"))).get(dominant)
    except Exception:
        pass
    zero_fill_bytes = len(calls) * dense
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": dom["frac"], "traffic": traffic, "peak_source": peak_src,
                "avg_launch_us": dom["avg_launch_us"], "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
                "step_frac_of_peak": round(value / world / peak, 4),
                "step_frac_excluding_zero_fill": round(whole_job_gbs(1, step_bytes - zero_fill_bytes, ms_per_step) / peak, 4),
                "note": "per-call step: 24 dense g_feat zero fills are %d of the %d algorithmic bytes" % (zero_fill_bytes, step_bytes)}
    grouped_layer_us = {k: round(float(np.mean(v)), 1) for k, v in gl.items() if v}

    # ---- reference CUDA op (oracle/_ref, rebuilt from the reference sources) on the same inputs
    ref_cuda = None
    if not bf16 and not args.skip_ref_op:
        try:
            from oracle import build_ref
            if build_ref.available():
                ext = build_ref.load()
                rf = feat
                res = {}
                for c in calls[:4]:
                    z = lambda x: torch.zeros_like(x)
                    for _ in range(2):
                        ext.deformable_aggregation_forward(rf, shapes_d, starts_d, c["loc_d"], c["w_d"])
                    torch.cuda.synchronize()
                    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                    n = 5
                    e0.record()
                    for _ in range(n):
                        ext.deformable_aggregation_forward(rf, shapes_d, starts_d, c["loc_d"], c["w_d"])
                    e1.record()
                    for _ in range(n):
                        gf, gl_, gw = z(rf), z(c["loc_d"]), z(c["w_d"])
                        ext.deformable_aggregation_backward(rf, shapes_d, starts_d, c["loc_d"], c["w_d"], c["go_d"],
                                                            gf, gl_, gw)
                    e2.record()
                    torch.cuda.synchronize()
                    res[c["kind"]] = {"ref_fwd_us": round(e0.elapsed_time(e1) * 1e3 / n, 1),
                                      "ref_bwd_us": round(e1.elapsed_time(e2) * 1e3 / n, 1),
                                      "ours_fwd_us": round(float(np.mean(per_mod[c["kind"]]["fwd_us"])), 1),
                                      "ours_bwd_us": round(float(np.mean(per_mod[c["kind"]]["bwd_us"])), 1)}
                ref_cuda = res
        except Exception as e:
            ref_cuda = {"error": str(e)[:200]}

    # ---- end to end through the public Python API, host buffers, H2D + D2H inside the timed region
    e2e = None
    if not args.skip_e2e:
        e2e = e2e_leg(args, hipad_b200, calls, layers, feat_h, shapes_d, starts_d, dev, world, step_bytes, max_over_ranks, barrier)

    # ---- the callers of the path (BASELINE configs[1] / [2]): the unmodified reference decoder through the harness
    decoder_forward = train_step = None
    if not args.skip_decoder and not bf16:
        del g_feat_l, ws_b, ws_f, flush
        torch.cuda.empty_cache()
        decoder_forward, train_step = decoder_legs(args, dev, world, rank, max_over_ranks)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cpu_baseline = cpu_reference(steps=2, warmup=1)["cpu_baseline"]

    if rank == 0:
        cfg = shared_config(bs, args.dtype)
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if bf16 else "f32", "data": "synthetic",
            "config": cfg,
            "run": {"parallelism": "batch-sharded x%d" % world, "l2": "256 MiB read sweep (flush) between timed steps",
                    "cuda_graph": graph is not None, "streams": n_lanes,
                    "api": "C ABI, one call at a time (hipad_dfa_group_forward with 1 call + hipad_dfa_backward_*), "
                           "the 4 calls of a layer as parallel graph branches",
                    "locations": "B2D camera geometry, det/map/plan visible fraction ~0.20/0.19/0.13, ego 0"},
            "samples_per_s": round(world * bs / (ms_per_step * 1e-3), 2),
            "algorithmic_bytes_per_step": int(step_bytes),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "roofline_per_kernel": per_kernel, "cpu_baseline": cpu_baseline,
            "group_step": group_step, "grouped_layer_stage_us": grouped_layer_us,
            "decoder_forward": decoder_forward, "train_step": train_step,
            "decoder_forward_dfa_only": inference,
            "per_call_us": {k: {"fwd": round(float(np.mean(v["fwd_us"])), 1), "bwd": round(float(np.mean(v["bwd_us"])), 1),
                                "B_fwd": v["bytes"]["fwd"], "B_bwd": v["bytes"]["bwd"], "U_rows": v["bytes"]["U"]}
                            for k, v in per_mod.items()},
            "reference_cuda_op_same_gpu": ref_cuda,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def e2e_leg(args, hipad_b200, calls, layers, feat_h, shapes_d, starts_d, dev, world, step_bytes, max_over_ranks, barrier):
    """The same step through the public Python API (hipad_b200.deformable_aggregation_group per decoder layer +
    share_feature_gradient + autograd) from pinned HOST memory: one staging arena per step, THREE large host->device
    copies (features / all locations + weights / all output gradients), two device arenas so that step k+1's copies
    run under step k's compute; outputs and gradient checksums are read back every step."""
    bs = feat_h.shape[0]
    f32 = torch.float32
    sizes_lw = []
    for c in calls:
        sizes_lw += [c["loc"].size, c["weights"].size]
    n_lw = int(sum(sizes_lw))
    n_go = int(sum(c["grad_out"].size for c in calls))
    host_lw = torch.empty(n_lw, dtype=f32).pin_memory()
    host_go = torch.empty(n_go, dtype=f32).pin_memory()
    o = 0
    for c in calls:
        for key in ("loc", "weights"):
            n = c[key].size
            host_lw[o:o + n] = torch.from_numpy(c[key]).reshape(-1)
            o += n
    o = 0
    for c in calls:
        n = c["grad_out"].size
        host_go[o:o + n] = torch.from_numpy(c["grad_out"]).reshape(-1)
        o += n
    feat_pin = feat_h.pin_memory()
    h2d = feat_pin.numel() * feat_pin.element_size() + (n_lw + n_go) * 4
    a_total = sum(c["A"] for c in layers[0])
    out_host = torch.empty((len(layers), bs, a_total, C), dtype=f32).pin_memory()
    gsum_host = torch.empty(3, dtype=f32).pin_memory()
    d2h = out_host.numel() * 4 + 12

    class Arena:
        def __init__(self):
            self.feat = torch.empty_like(feat_h, device=dev)
            self.lw = torch.empty(n_lw, dtype=f32, device=dev)
            self.go = torch.empty(n_go, dtype=f32, device=dev)
            self.ready = torch.cuda.Event()
            self.free = torch.cuda.Event()
            self.free.record()

    arenas = [Arena(), Arena()]
    copy_s = torch.cuda.Stream(device=dev)

    def upload(a):
        copy_s.wait_event(a.free)           # the step that last used this arena is done with it
        with torch.cuda.stream(copy_s):
            a.feat.copy_(feat_pin, non_blocking=True)
            a.lw.copy_(host_lw, non_blocking=True)
            a.go.copy_(host_go, non_blocking=True)
            a.ready.record(copy_s)

    def compute(a):
        cur = torch.cuda.current_stream()
        cur.wait_event(a.ready)
        f_in = a.feat.detach().requires_grad_(True)
        f = hipad_b200.share_feature_gradient(f_in)
        outs, gos, leaves = [], [], []
        o_lw = o_go = 0
        for group in layers:
            pairs = []
            for c in group:
                n1, n2 = c["loc"].size, c["weights"].size
                loc = a.lw[o_lw:o_lw + n1].view(c["loc"].shape).detach().requires_grad_(True)
                w = a.lw[o_lw + n1:o_lw + n1 + n2].view(c["weights"].shape).detach().requires_grad_(True)
                o_lw += n1 + n2
                pairs.append((loc, w))
                n3 = c["grad_out"].size
                gos.append(a.go[o_go:o_go + n3].view(c["grad_out"].shape))
                o_go += n3
            leaves.append(pairs[0])
            outs += hipad_b200.deformable_aggregation_group(f, shapes_d, starts_d, pairs)
        torch.autograd.backward(outs, gos)
        n_mod = len(layers[0])
        for li in range(len(layers)):
            out_host[li].copy_(torch.cat([x.detach() for x in outs[li * n_mod:(li + 1) * n_mod]], dim=1), non_blocking=True)
        gsum_host.copy_(torch.stack([f_in.grad.float().abs().sum(), leaves[0][0].grad.abs().sum(),
                                     leaves[0][1].grad.abs().sum()]), non_blocking=True)
        a.free.record(cur)

    n_e2e = max(3, min(args.steps, 8))
    upload(arenas[0]); compute(arenas[0]); torch.cuda.synchronize()      # warm-up
    barrier()
    t0 = time.perf_counter()
    upload(arenas[0])
    for k in range(n_e2e):
        if k + 1 < n_e2e:
            upload(arenas[(k + 1) % 2])      # next step's copies run under this step's compute
        compute(arenas[k % 2])
    torch.cuda.synchronize()
    dt_s = max_over_ranks((time.perf_counter() - t0) / n_e2e, dev)
    return {"value": round(whole_job_gbs(world, step_bytes, dt_s * 1e3), 2), "unit": "GB/s",
            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": round(dt_s * 1e3, 3), "steps": n_e2e,
            "h2d_gbs_per_rank": round(h2d / dt_s / 1e9, 1),
            "limiter": "host->device copies: %.0f MB per step per rank from pinned memory (PCIe); compute is hidden under "
                       "them" % (h2d / 1e6),
            "api": "hipad_b200.deformable_aggregation_group per decoder layer + share_feature_gradient + autograd; one pinned "
                   "arena, 3 copies per step, double-buffered device arenas (H2D of step k+1 under compute of step k)"}


def decoder_legs(args, dev, world, rank, max_over_ranks):
    """BASELINE configs[1] and [2] through the UNMODIFIED reference decoder (harness/): decoder forward bs=1 with the
    reference CUDA op vs hipad_b200, and the training step (ResNet-50 + FPN stand-in, decoder, surrogate loss, AdamW,
    DDP gradient all-reduce when world > 1) at --train-bs samples per GPU."""
    try:
        from harness import bench_decoder, train_step as ts, vendor
        if vendor.vendor() is None:
            return {"unavailable": "baseline/_ref/hipad not vendored"}, None
    except Exception as e:
        return {"error": repr(e)[:200]}, None
    fwd = tr = None
    try:
        variants = (("reference", "ours", "ours_module", "ours_module_graph") if (rank == 0 and world == 1)
                    else ("ours_module_graph",))
        r = bench_decoder.run(hw_list=((352, 640), (256, 704)) if world == 1 else ((352, 640),), frames=args.decoder_frames,
                              bs=1, variants=variants)
        for hw, d in r.items():
            for v in d.values():
                if "ms_per_forward_median" in v:
                    v["ms_per_forward_median"] = round(max_over_ranks(v["ms_per_forward_median"], dev), 3)
                    v["samples_per_s"] = round(world * 1e3 / v["ms_per_forward_median"], 1)
        fwd = r
        fwd["note"] = ("unmodified reference SparseOneDecoder (70.8 M parameters, stage-2 config) through harness/; bs=1 per GPU, "
                       "eager PyTorch, host-bound (hundreds of small launches per forward); samples_per_s is the whole job")
    except Exception as e:
        fwd = {"error": repr(e)[:300]}
    try:
        t = ts.time_train_step("ours_module", bs=args.train_bs, steps=args.train_steps, warmup=2, device=dev, world=world,
                               rank=rank)
        t["ms_per_step"] = round(max_over_ranks(t["ms_per_step"], dev), 2)
        t["samples_per_s"] = round(world * args.train_bs / (t["ms_per_step"] * 1e-3), 2)
        t["n_gpus"] = world
        t["collective"] = ("DistributedDataParallel bucketed NCCL all-reduce of %.1f M fp32 gradients per step" % t["trainable_params_M"]
                           if world > 1 else "none (1 GPU)")
        tr = t
        if rank == 0 and world == 1 and not args.skip_train_reference:
            try:
                tref = ts.time_train_step("reference", bs=args.train_bs, steps=max(2, args.train_steps // 2), warmup=1, device=dev)
                tr["reference_cuda_op"] = {"ms_per_step": tref["ms_per_step"],
                                           "samples_per_s": round(args.train_bs / (tref["ms_per_step"] * 1e-3), 2)}
            except Exception as e:
                tr["reference_cuda_op"] = {"error": repr(e)[:200]}
    except Exception as e:
        tr = {"error": repr(e)[:300]}
    return fwd, tr


# ------------------------------------------------------------------------------------- CPU arm
def cpu_reference(steps, warmup):
    """The reference's CPU implementation of the path (torch grid_sample branch) on the host cores.

    One step = the four DFA calls of ONE decoder layer at bs=1, forward + backward (1/6 of a GPU step): a FIXED,
    bounded sample of the workload (never shrunk), `steps` timed runs after `warmup` untimed ones."""
    from oracle import torch_path as tp
    import oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    calls, shapes, starts, F = make_calls(1, seed=0, layers=1)
    rng = np.random.default_rng(1234)
    lv = level_hw(FINAL_HW)
    fmaps = [torch.from_numpy(rng.standard_normal((1, CAMS, C, h, w), dtype=np.float32)).requires_grad_(True)
             for h, w in lv]

    def bytes_of(c):
        # same definition as the GPU arm, U counted with the oracle's integer indices
        U = oracle.unique_rows(shapes, starts, c["loc"], F)
        A, P = c["A"], c["P"]
        loc_b, w_b, out_b = A * P * CAMS * 8, A * P * CAMS * 4 * G * 4, A * C * 4
        return (loc_b + w_b + U * C * 4 + out_b) + (out_b + 2 * loc_b + 2 * w_b + U * C * 4 + F * C * 4)

    def run(sample):
        for c in sample:
            p2d = torch.from_numpy(c["loc"]).permute(0, 3, 1, 2, 4).contiguous().requires_grad_(True)
            w = torch.from_numpy(c["weights"]).permute(0, 1, 3, 4, 2, 5).contiguous().requires_grad_(True)
            out = tp.grid_sample_path(fmaps, p2d, w)
            out.backward(torch.from_numpy(c["grad_out"]))

    desc = "one decoder layer (det+map+plan+ego calls), bs=1, fwd+bwd, torch grid_sample path"
    for _ in range(warmup):
        run(calls)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter(); run(calls); times.append(time.perf_counter() - t0)
    nbytes = sum(bytes_of(c) for c in calls)
    sec = float(np.mean(times))
    val = nbytes / sec / 1e9
    return {"value": val, "ms_per_step": sec * 1e3, "steps": steps,
            "cpu_baseline": {"value": round(val, 4), "unit": "GB/s", "cores": cores, "kind": "port",
                             "sample": desc + "; %d timed runs, %.2f s each" % (steps, sec)}}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # same --steps / --warmup as the GPU arm; each step is the fixed bounded sample (one decoder layer, ~1-2 s)
    steps, warmup = args.steps, max(args.warmup, 3)
    if args.ref_max_steps > 0:
        steps = min(steps, args.ref_max_steps)
    r = cpu_reference(steps=steps, warmup=warmup if args.ref_max_steps <= 0 else 1)
    line = {"impl": "reference", "metric": METRIC, "value": round(r["value"], 4), "unit": "GB/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": max(args.warmup, 3),
            "timed_steps": r["steps"],
            "ms_per_step": round(r["ms_per_step"], 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": shared_config(args.bs, args.dtype),
            "run": {"what": "the reference's own CPU implementation of the path (torch grid_sample branch, oracle/torch_path.py "
                            "restatement), all host cores; each step = the four calls of ONE decoder layer at bs=1 (a bounded "
                            "sample: 1/6 of a GPU step), GB/s from the same byte definition"},
            "cpu_baseline": r["cpu_baseline"],
            "e2e": {"value": round(r["value"], 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bs", type=int, default=1, help="samples per GPU")
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"], help="feature-map storage type")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--streams", type=int, default=4,
                    help="4: the four modality calls of a decoder layer run as parallel branches; 1: strictly serial")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-decoder", action="store_true", help="skip the decoder-forward and training-step legs")
    ap.add_argument("--skip-ref-op", action="store_true", help="skip timing the reference CUDA op on the same GPU")
    ap.add_argument("--skip-train-reference", action="store_true")
    ap.add_argument("--decoder-frames", type=int, default=20)
    ap.add_argument("--train-bs", type=int, default=4, help="samples per GPU of the training step (BASELINE configs[2])")
    ap.add_argument("--train-steps", type=int, default=4)
    ap.add_argument("--ref-max-steps", type=int, default=0,
                    help="reference arm: cap on timed steps (0 = exactly --steps, what the driver compares)")
    args = ap.parse_args()
    isolate_stdout()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
