#!/usr/bin/env python
"""bench.py — 256^3 volume-pair flow + interpolation throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

A "step" = one pass of the hot path (Model.inference: 3-scale 3-D IFNet + warps + blend) over one batch of
synthetic droplet-shaped volume pairs per GPU.  Pairs are batch-sharded over ranks, no data-path collective
("scaling": "weak").  One JSON line is printed by rank 0; see DESIGN.md §Measurement for every field.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nd, spatial, pairs per GPU per step, BASELINE.json config it stands for)
    "flow3d_droplet256": (3, (256, 256, 256), 4, "Flow-3D droplet-shaped 256^3 byte volume pairs, ensemble batch-sharded"),
    "flow3d_rect128": (3, (128, 128, 128), 4, "Flow-3D IFNet on synthetic 3D textured rectangle 128^3, batch 4"),
    "flow2d_droplet": (2, (160, 224), 64, "Flow-2D droplet-shaped 160x224 monochrome, batch 64"),
    "flow2d_rect_b1": (2, (160, 224), 1, "Flow-2D RIFE IFNet inference on synthetic textured rectangle 160x224, batch 1"),
}
# a separate operator-level workload (BASELINE.json configs[4]): see run_upflow_ops()
UPFLOW_LEVELS = ((196, 4, 13), (128, 8, 26), (96, 16, 52), (64, 32, 104), (32, 64, 208))     # (C, H, W) of a 256x832 pair, SURVEY.md §8 a8
UPFLOW_PARAMS = 3_354_146                                                                  # UPFlow parameter count (13.4 MB fp32)


def ifnet_macs(nd, sp):
    """Algorithmic multiply-accumulates of one IFNet inference on one pair (logical channels)."""
    widths = {2: (128, 96, 64), 3: (128, 64, 64)}[nd]
    k0 = 3 if nd == 2 else 4
    nf = 2 * nd
    vox = 1
    for s in sp:
        vox *= s
    total = 0
    for (c, cin, scale) in zip(widths, (2, 5 + nf, 5 + nf), (4, 2, 1)):
        v_in = vox // scale ** nd
        v1, v2 = v_in // 2 ** nd, v_in // 4 ** nd
        total += v1 * cin * (c // 2) * k0 ** nd + v2 * (c // 2) * c * k0 ** nd          # conv0
        total += 8 * v2 * c * c * 3 ** nd                                                 # convblock0..3
        total += 2 * v2 * c * (c // 2) * 4 ** nd                                          # conv1.0, conv2.0 (ConvT: in positions)
        total += v1 * (c // 2) * (nf + 1) * 4 ** nd                                       # conv1.2, conv2.2
    return total


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 3.0:      # nvidia-smi needs 0.1-1 s to print its first line
                time.sleep(0.02)
            self.first = len(self.rows)                            # samples before the load starts are dropped
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def keep_loaded(self, fn, min_samples=4, max_s=2.0):
        """The timed region of the default run lasts ~0.1 s, shorter than nvidia-smi's sampling period under load: continue the SAME
        work untimed until `min_samples` samples were taken under it."""
        t0 = time.time()
        while self.proc and len(self.rows) - self.first < min_samples and time.time() - t0 < max_s:
            fn()

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            self.t.join(timeout=2)
        return False

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows[getattr(self, "first", 0):]:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "window": "the timed region and an untimed continuation of the same steps (nvidia-smi -lms 50)"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


# ------------------------------------------------------------------------------------------------ CPU arms
def _workload_inputs(workload, pairs, seed=1234):
    """Synthetic host inputs of a workload as numpy arrays (img0, gt, img1)."""
    from opticalflowscivis_b200 import synth
    nd, sp, _, _ = WORKLOADS[workload]
    if workload == "flow3d_droplet256":
        return synth.droplet3d_u8(pairs, sp[0], seed=seed)
    if workload == "flow3d_rect128":
        return synth.rectangle3d(pairs, sp[0], seed=seed)
    if workload == "flow2d_rect_b1":
        return synth.rectangle2d(pairs, *sp, seed=seed)
    return synth.droplet2d(pairs, *sp, seed=seed)


def base_config(workload, pairs):
    """`config`: printed identically by both arms (ours / --impl reference); arm-specific settings go to `impl_config`."""
    nd, sp, _, desc = WORKLOADS[workload]
    return {"workload": workload, "describes": desc, "spatial": list(sp), "pairs_per_gpu_per_step": pairs,
            "weights": "random init, seed 1234",
            "only_last": "3-D inference returns merged[2] only (Flow-3D/model/RIFE.py:75): the blends of scales 0 and 1 are skipped, "
                         "their flows are still returned" if nd == 3 else "all three blends run (2-D returns merged[0..2])",
            "l2": "working set per step (>1 GB) exceeds the 126 MB L2; no explicit flush" if nd == 3 else
                  "inputs + activations of a step exceed L2 only at batch 64; no explicit flush"}


def reference_model(nd, device="cpu"):
    """oracle/ifnet_ref.py: the bit-exact torch restatement of the reference's Model.inference (pinned in tests/golden)."""
    import torch
    from oracle.ifnet_ref import ModelRef
    torch.manual_seed(1234)
    m = ModelRef(nd).eval()
    m.flownet.to(device)
    return m


def cpu_reference_rate(workload, pairs, steps, warmup):
    """The reference's CPU path on the SAME workload (same spatial size; `pairs` pairs per call) with all host threads.
    Returns (pairs_per_s, seconds_per_call, threads)."""
    import torch
    nd = WORKLOADS[workload][0]
    torch.set_num_threads(os.cpu_count() or 1)
    m = reference_model(nd)
    a, _, b = _workload_inputs(workload, pairs)
    if a.dtype.kind == "u":
        img0, img1 = torch.from_numpy(a).float() / 255.0, torch.from_numpy(b).float() / 255.0
    else:
        img0, img1 = torch.from_numpy(a), torch.from_numpy(b)
    for _ in range(warmup):
        m.inference(img0, img1)
    t0 = time.perf_counter()
    for _ in range(steps):
        m.inference(img0, img1)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return pairs / dt, dt, torch.get_num_threads()


# pairs per CPU call: ONE real pair of the workload's size for the 3-D volumes (a 256^3 call takes 10-25 s), the whole batch in 2-D
CPU_SAMPLE = {"flow3d_droplet256": (1, "one full 256^3 pair per call (a quarter of the 4-pair step; rate = 1 / seconds per call)"),
              "flow3d_rect128": (1, "one full 128^3 pair per call"),
              "flow2d_droplet": (64, "the full batch of 64 pairs of 160x224 per call"),
              "flow2d_rect_b1": (1, "the one 160x224 pair per call")}


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port, bit-exact against the imported reference)
    on this box's host cores, same workload / metric / unit; every step is a bounded sample of the step (see CPU_SAMPLE)."""
    if rank != 0:
        return
    if args.workload in ("upflow_ops", "train3d", "train2d", "upflow_net"):
        print(json.dumps({"impl": "reference", "unavailable": f"{args.workload} is a next-tier workload (SURVEY.md §8f) without a CPU arm "
                          "(train3d reports the reference's eager-CUDA update as `cuda_eager_reference`)"}), flush=True)
        return
    nd, sp, pairs, desc = WORKLOADS[args.workload]
    s_pairs, sample = CPU_SAMPLE[args.workload]
    with contextlib.redirect_stdout(sys.stderr):
        rate, dt, threads = cpu_reference_rate(args.workload, s_pairs, args.steps, args.warmup)
    unit = "pairs/s"
    line = {
        "impl": "reference", "metric": metric_name(args.workload), "value": rate, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": base_config(args.workload, pairs),
        "cpu_baseline": {"value": rate, "unit": unit, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cuda_eager_rates(workload, pairs, dev, reps=3):
    """Context number: the reference's own modules (oracle port = same torch calls: cuDNN convs, ATen grid_sample /
    interpolate) in eager PyTorch ON THIS GPU, one call of `pairs` pairs, strict fp32 and with TF32 allowed."""
    import torch
    nd = WORKLOADS[workload][0]
    m = reference_model(nd, dev)
    a, _, b = _workload_inputs(workload, pairs)
    cvt = (lambda v: torch.from_numpy(v).to(dev).float() / 255.0) if a.dtype.kind == "u" else (lambda v: torch.from_numpy(v).to(dev))
    img0, img1 = cvt(a), cvt(b)
    out = {}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for name, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            for _ in range(2):
                m.inference(img0, img1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                m.inference(img0, img1)
            e1.record()
            torch.cuda.synchronize()
            out[name] = pairs * reps / (e0.elapsed_time(e1) / 1e3)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    del m, img0, img1
    torch.cuda.empty_cache()
    return {"value_fp32": out["fp32"], "value_tf32": out["tf32"], "unit": "pairs/s", "kind": "port",
            "what": f"oracle/ifnet_ref.ModelRef (the reference's torch calls) in eager PyTorch on this GPU, {pairs} pair(s) per call, "
                    f"resident inputs, {reps} calls after 2 warm-ups; cudnn.allow_tf32 False / True"}


def metric_name(workload):
    return {"flow3d_droplet256": "256^3 volume-pair interps/sec", "flow3d_rect128": "128^3 volume-pair interps/sec",
            "flow2d_droplet": "160x224 frame-pair interps/sec", "flow2d_rect_b1": "160x224 frame-pair interps/sec",
            "upflow_ops": "UPFlow flow-path operator training steps/sec"}[workload]


# ------------------------------------------------------------------------------------------------ GPU arm
def _timed(fn, steps, barrier):
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    fn(steps)
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


# DRAM bytes per 256^3 pair from the committed `ncu --set full` capture of one inference (dram__bytes_read.sum + dram__bytes_write.sum
# summed over the launches of a class; file named in `traffic_source`).  None until a capture of the current kernels is committed.
NCU_TRAFFIC = {
    "source": "profiles/r02i_ncu_full_step_1x256.csv (ncu --set full --clock-control none of one 256^3 inference; per pair, scaled by the batch)",
    "conv_all": 3.052e9,            # the 36 conv launches (algorithmic FLOP-side traffic is not defined; weights + activations)
    "conv_hbm_layers": 1.979e9,     # block2 conv0.0 (0.672 GB) + block2 heads (1.307 GB); algorithmic 2.013 GB
    "stage": 3.488e9,               # the three stage launches; algorithmic 3.355 GB
}


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from opticalflowscivis_b200 import ops
    from opticalflowscivis_b200.pipeline import StreamedInterpolator
    from opticalflowscivis_b200.rife import Model2D, Model3D

    nd, sp, pairs, desc = WORKLOADS[args.workload]
    if args.pairs:
        pairs = args.pairs
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from opticalflowscivis_b200.shard import bind_to_gpu_numa
    numa = bind_to_gpu_numa(local_rank) if world > 1 else {"bound": False}      # before the first pinned allocation
    print(f"[bench rank {rank}] cpu binding: {numa}", file=sys.stderr, flush=True)
    torch.manual_seed(1234)
    model = (Model3D if nd == 3 else Model2D)(local_rank=local_rank, precision=args.precision, engine=args.engine)
    model.eval()
    graphs = args.workload == "flow2d_rect_b1"      # one 160x224 pair is host-enqueue-bound: replay the call from a CUDA graph

    # synthetic inputs: member seed = 1234 + global pair index
    a, _, b = _workload_inputs(args.workload, pairs, seed=1234 + rank * pairs)
    h0, h1 = torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()      # droplet volumes: uint8 on the host; others fp32
    bytes_in = h0.dtype == torch.uint8
    as_f32 = (lambda t: t.float().div_(255.0)) if bytes_in else (lambda t: t)
    d0, d1 = as_f32(h0.to(dev)), as_f32(h1.to(dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(k=1):
        for _ in range(k):
            model.inference(d0, d1)

    launches_per_call = None
    if graphs:
        c0 = ops.launch_count()
        model.inference(d0, d1)                         # one eager call: the launches a replay of the captured graph repeats
        launches_per_call = ops.launch_count() - c0
        model.enable_cuda_graphs()
    step_resident(args.warmup)
    barrier()

    # ---- timed region 1 (`value`): inputs resident in HBM, nothing but the K inference calls between the two events
    n0 = ops.launch_count()
    with ClockSampler(local_rank) as clk:
        ms = _timed(step_resident, args.steps, barrier)
        launches = ops.launch_count() - n0 if launches_per_call is None else launches_per_call * args.steps
        clk.keep_loaded(lambda: (step_resident(2), torch.cuda.synchronize()))

    # ---- profile pass (the same K steps again): CUDA events around every launch, per kernel class, on the launching stream
    classes, layer_ms, ms_prof = {}, None, None
    if not graphs:
        timer = ops.LaunchTimer()
        ops.TIMER = timer
        ms_prof = _timed(step_resident, args.steps, barrier)
        ops.TIMER = None
        classes = timer.totals()
        conv_seq = [(a_.elapsed_time(b_)) for (name, a_, b_) in timer.seq if name.startswith("conv_")]
        if len(conv_seq) == 36 * args.steps:       # 3 blocks x 12 conv launches, fixed order: per-layer mean time over the steps
            layer_ms = [sum(conv_seq[i::36]) / args.steps for i in range(36)]

    # ---- standalone warp kernel (the "warp HBM GB/s vs peak" half of the metric): a1 / a2 on the workload's shape
    g = torch.Generator(device="cpu").manual_seed(7)
    wflow = (torch.randn((pairs, nd) + tuple(max(1, s // 8) for s in sp), generator=g) * 2.0).to(dev)
    wflow = torch.nn.functional.interpolate(wflow, size=tuple(sp), mode="trilinear" if nd == 3 else "bilinear").contiguous()
    warp_fn = ops.warp3d if nd == 3 else ops.warp2d
    wreps = 10

    def warp_loop(k):
        for _ in range(k):
            warp_fn(d0, wflow)

    warp_loop(3)
    warp_ms = _timed(warp_loop, wreps, lambda: torch.cuda.synchronize()) / wreps
    del wflow

    # ---- timed region 2 (`e2e`): pinned host pairs -> H2D -> Model.inference -> D2H of the interpolated volume through the
    #      package's streaming front end (pipeline.StreamedInterpolator: upload / compute / download on three streams).  Byte
    #      volumes (cfg 4) are exported the way the reference's inference driver exports them, `(merged * 255).byte()` computed on
    #      the device before the `.cpu()` (Flow-3D/inference_img.py:105): bytes in, bytes out.  `e2e_f32` is the same run with the
    #      fp32 result downloaded instead (Flow-3D/train.py:270 evaluates on it): the conservative figure.
    def e2e_run(out_u8):
        st = StreamedInterpolator(model, dev, out_u8=out_u8, depth=3)
        nbuf = 4
        outs = [torch.empty((pairs, 1) + tuple(sp), dtype=torch.uint8 if out_u8 else torch.float32).pin_memory() for _ in range(nbuf)]

        def run(k):
            for _ in st.run(((h0, h1) for _ in range(k)), (outs[i % nbuf] for i in range(k))):
                pass            # returns when the last result has landed in host memory

        run(min(args.warmup, 3))
        h2d0, d2h0 = st.h2d_bytes, st.d2h_bytes
        t = _timed(run, args.steps, barrier)
        return t, (st.h2d_bytes - h2d0) // args.steps, (st.d2h_bytes - d2h0) // args.steps

    ms_e2e_f32, h2d_step, d2h_f32_step = e2e_run(False)
    ms_e2e_u8 = d2h_u8_step = None
    if bytes_in:
        ms_e2e_u8, _, d2h_u8_step = e2e_run(True)

    if world > 1:
        t = torch.tensor([ms, ms_e2e_f32, ms_e2e_u8 if ms_e2e_u8 is not None else 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e_f32 = float(t[0]), float(t[1])
        ms_e2e_u8 = float(t[2]) if ms_e2e_u8 is not None else None
    if rank != 0:
        return

    peaks = load_peaks()
    total_pairs = pairs * world * args.steps
    value = total_pairs / (ms / 1e3)
    vox = 1
    for s in sp:
        vox *= s

    def e2e_entry(t_ms, d2h, export):
        return {"value": total_pairs / (t_ms / 1e3), "unit": "pairs/s", "h2d_bytes_per_step": int(h2d_step), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": t_ms / args.steps, "export": export}

    e2e_f32 = e2e_entry(ms_e2e_f32, d2h_f32_step, "fp32 merged volume (Flow-3D/train.py:270)")
    e2e = e2e_entry(ms_e2e_u8, d2h_u8_step, "(merged*255).byte() on the device (Flow-3D/inference_img.py:105); input was uint8") if bytes_in else e2e_f32

    flops_pair = 2.0 * ifnet_macs(nd, sp)
    roofline = roofline_hbm = roofline_stage = None
    share, dominant = {}, None
    if classes:
        conv_cls = [k for k in classes if k.startswith("conv_")]
        conv_ms = sum(classes[k][1] for k in conv_cls)
        conv_n = sum(classes[k][0] for k in conv_cls)
        share = {k: round(v[1] / ms_prof, 4) for k, v in classes.items()}
        dominant = max(classes, key=lambda k: classes[k][1])
        at256 = nd == 3 and tuple(sp) == (256, 256, 256)
        # tensor roofline of the conv engine: algorithmic FLOPs of ALL conv launches of the profile pass / their summed time, against
        # the BURST bf16 peak (the timed region is ~0.1 s at full clock; the sustained figure was taken at a 1342 MHz median)
        conv_tf = flops_pair * pairs * args.steps / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else None
        roofline = {"bound": "tensor", "achieved": conv_tf, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                    "frac": (conv_tf / peaks["tf_burst"]) if conv_tf else None,
                    "traffic": (NCU_TRAFFIC["conv_all"] * pairs) if (at256 and NCU_TRAFFIC["conv_all"]) else None,
                    "kernel": "+".join(sorted(conv_cls)), "frac_of_sustained_peak": (conv_tf / peaks["tf_sustained"]) if conv_tf else None,
                    "peak_source": peaks["src"] + " (burst bf16; sustained %.0f)" % peaks["tf_sustained"], "launches": conv_n,
                    "share_of_step": round(conv_ms / ms_prof, 4), "algorithmic_flops_per_pair": flops_pair,
                    "traffic_source": NCU_TRAFFIC["source"]}
        if layer_ms is not None and nd == 3:
            # block2's first conv and final heads are HBM-bound at full resolution -> also on the HBM roofline.
            # conv0.0: reads the 16-channel bf16 block input (32 B/voxel), writes 32 channels bf16 at 1/8 of the voxels (8 B/voxel);
            # heads: read 64 channels bf16 at 1/8 of the voxels (16 B/voxel), read + write the fp32 flow/mask state (32 + 32 B/voxel)
            hb = (32 + 8 + 16 + 32 + 32) * vox * pairs
            ht = layer_ms[24] + layer_ms[35]
            roofline_hbm = {"kernel": "block2.conv0.0 + block2 heads (conv_halo)", "bound": "hbm", "achieved": hb / (ht / 1e3) / 1e9,
                            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hb / (ht / 1e3) / 1e9 / peaks["hbm_gbs"],
                            "traffic": (NCU_TRAFFIC["conv_hbm_layers"] * pairs) if (at256 and NCU_TRAFFIC["conv_hbm_layers"]) else None,
                            "ms_per_step": [round(layer_ms[24], 4), round(layer_ms[35], 4)], "algorithmic_bytes_per_step": hb}
        bs = classes.get("block_stage")
        if bs and nd == 3:
            per_vox = [32 + 8 + 32 / 8.0,              # block0 -> 1: write state, read imgs, pooled bf16 input of block1 (head0 is L2-resident)
                       32 + 32 + 8 + 32 + 4,           # block1 -> 2: read + write state, imgs, full-res bf16 input of block2, head1 (32 B / 8 voxels)
                       32 + 8 + 8]                     # final: read the state (accumulated by the head conv), imgs, write merged + mask
            stage_bytes = sum(per_vox) * vox * pairs * args.steps
            sg = stage_bytes / (bs[1] / 1e3) / 1e9
            roofline_stage = {"kernel": "stage3d_hfast_kernel", "bound": "hbm", "achieved": sg, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                              "frac": sg / peaks["hbm_gbs"],
                              "traffic": (NCU_TRAFFIC["stage"] * pairs) if (at256 and NCU_TRAFFIC["stage"]) else None,
                              "traffic_unit": "bytes per step (3 launches)", "launches": bs[0], "share_of_step": round(bs[1] / ms_prof, 4),
                              "algorithmic_bytes_per_voxel_per_scale": per_vox}
    # HBM roofline of the standalone warp kernel: (nd flow + 1 src + 1 out) * 4 B = 20 B/voxel in 3-D, 16 B/px in 2-D
    warp_bytes = (nd + 2) * 4 * vox * pairs
    warp_gbs = warp_bytes / (warp_ms / 1e3) / 1e9
    # DRAM traffic per launch from the committed `ncu --set full` capture of this kernel (profiles/r01q_ncu_full_warp3d_slab_4x256.csv:
    # dram__bytes_read 1118.0 MB + write 254.1 MB for a 4 x 256^3 launch = 279.5 + 63.5 MB per volume; algorithmic 335.5 MB)
    warp_traffic = (279.5e6 + 63.5e6) * pairs if (nd == 3 and tuple(sp) == (256, 256, 256)) else None
    slab = nd == 3 and sp[0] == sp[1] == sp[2] and sp[0] % 32 == 0      # ofsv_warp3d_f32 picks the TMA slab kernel on such volumes
    roofline_warp = {"kernel": "warp3d_slab_kernel" if slab else "warp%dd_kernel" % nd, "bound": "hbm", "achieved": warp_gbs,
                     "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": warp_gbs / peaks["hbm_gbs"], "traffic": warp_traffic,
                     "algorithmic_bytes_per_launch": warp_bytes, "ms_per_launch": warp_ms, "timed": "standalone, 10 launches"}

    cpu_baseline = cuda_eager = None
    if not args.no_cpu_baseline and world == 1:
        try:
            cuda_eager = cuda_eager_rates(args.workload, CPU_SAMPLE[args.workload][0], dev)
        except Exception as e:  # noqa: BLE001  (a context figure must never take the bench line down)
            cuda_eager = {"unavailable": repr(e)[:200]}
        with contextlib.redirect_stdout(io.StringIO()):
            s_pairs, sample = CPU_SAMPLE[args.workload]
            big = args.workload == "flow3d_droplet256"
            rate, dt, thr = cpu_reference_rate(args.workload, s_pairs, 1 if big else 2, 0 if big else 1)
        cpu_baseline = {"value": rate, "unit": "pairs/s", "cores": thr, "kind": "port",
                        "sample": sample + ("; 1 call, no warm-up" if big else "; 2 calls after 1 warm-up")}

    cfg = base_config(args.workload, pairs)
    line = {
        "metric": metric_name(args.workload), "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "e2e": e2e, "e2e_f32": e2e_f32 if bytes_in else None,
        "roofline": None if roofline is None else {k: roofline[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic")},
        "roofline_stage_frac": None if roofline_stage is None else round(roofline_stage["frac"], 4),
        "roofline_conv_hbm_frac": None if roofline_hbm is None else round(roofline_hbm["frac"], 4),
        "roofline_warp_frac": round(roofline_warp["frac"], 4),
        "gpu_launches": int(launches), "clocks": clk.summary(),
        "cpu_baseline": cpu_baseline, "cuda_eager": cuda_eager,
        "config": cfg,
        "impl_config": {"precision": args.precision, "conv_engine": model.flownet._engine(), "cuda_graphs": graphs},
        "e2e_path": "pipeline.StreamedInterpolator(model).run(pinned host pairs): H2D / Model.inference / D2H on three streams",
        "roofline_detail": roofline, "roofline_conv_hbm": roofline_hbm, "roofline_stage": roofline_stage, "roofline_warp": roofline_warp,
        "kernel_time_share": share, "dominant_kernel_class": dominant,
        "conv_layer_ms_per_step": None if layer_ms is None else [round(v, 4) for v in layer_ms],
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ UPFlow operator workload
def run_upflow_ops(args, rank, world, local_rank):
    """BASELINE.json configs[4] at the operator level: one step = forward AND backward of the §8a UPFlow operators the package
    implements (WarpingLayer_no_div a11, correlation cost volume + LeakyReLU a8, upsample2d_flow_as a10) on the five pyramid levels of
    a 256x832 pair at B = 16 (8 pairs x 2 directions per GPU), then the step's one collective — a SUM all-reduce of a flat fp32
    gradient bucket of UPFlow's size (3 354 146 parameters = 13.4 MB) over NCCL — and the fused AdamW step on it.  The estimator /
    context convolutions of UPFlow are NOT part of it."""
    import torch
    import torch.distributed as dist

    from opticalflowscivis_b200 import ops
    from opticalflowscivis_b200.optim import FusedAdamW, GradientBucket, allreduce_gradients

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B = args.pairs or 16
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    lv = []
    for (c, h, w) in UPFLOW_LEVELS:
        r = lambda *shape: torch.randn(*shape, generator=g).to(dev)      # noqa: E731
        lv.append({"c": c, "h": h, "w": w, "x": r(B, c, h, w), "f2": r(B, c, h, w), "fl": r(B, 2, h, w) * 2, "go": r(B, c, h, w),
                   "g81": r(B, 81, h, w), "gup": r(B, 2, 2 * h, 2 * w)})
    params = [torch.nn.Parameter(torch.zeros(UPFLOW_PARAMS, device=dev))]
    bucket = GradientBucket(params)
    opt = FusedAdamW(params, lr=1e-4, weight_decay=1e-4, bucket=bucket)
    bucket.flat.normal_(generator=None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = {"corr_fwd": [], "corr_bwd": [], "allreduce": []}

    def timed(key, fn, on):
        if not on:
            return fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        ev[key].append((a, b))
        return out

    def step(k=1, prof=False):
        for _ in range(k):
            for L in lv:
                ops.warping_no_div(L["x"], L["fl"])
                timed("corr_fwd", lambda: ops.corr81_fwd(L["x"], L["f2"], leaky_slope=0.1), prof)
                ops.upsample_flow_ac(L["fl"], 2 * L["h"], 2 * L["w"])
            for L in reversed(lv):
                ops.upsample_flow_ac_bwd(L["gup"], L["h"], L["w"])
                timed("corr_bwd", lambda: ops.corr81_bwd(L["x"], L["f2"], L["g81"]), prof)
                ops.warping_no_div_bwd(L["x"], L["fl"], L["go"])
            scale = timed("allreduce", lambda: allreduce_gradients(bucket), prof)
            opt.step(grad_scale=scale)

    step(args.warmup)
    n0 = ops.launch_count()
    with ClockSampler(local_rank) as clk:
        ms = _timed(step, args.steps, barrier)
        launches = ops.launch_count() - n0 + args.steps      # + the optimizer launch per step
        clk.keep_loaded(lambda: (step(20), torch.cuda.synchronize()))
    ms_prof = _timed(lambda k: step(k, True), args.steps, barrier)
    tot = {k: sum(a.elapsed_time(b) for a, b in v) / args.steps for k, v in ev.items()}
    if world > 1:
        t = torch.tensor([ms, tot["allreduce"]], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, tot["allreduce"] = float(t[0]), float(t[1])
    if rank != 0:
        return
    peaks = load_peaks()
    corr_bytes = sum((2 * c + 81) * h * w * 4 * B for (c, h, w) in UPFLOW_LEVELS)
    gbs = corr_bytes / (tot["corr_fwd"] / 1e3) / 1e9
    line = {
        "metric": metric_name("upflow_ops"), "value": world * args.steps / (ms / 1e3), "unit": "steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "e2e": None,
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "traffic": None},
        "allreduce_ms_per_step": tot["allreduce"], "allreduce_bytes": UPFLOW_PARAMS * 4,
        "corr_fwd_ms_per_step": tot["corr_fwd"], "corr_bwd_ms_per_step": tot["corr_bwd"],
        "gpu_launches": int(launches), "clocks": clk.summary(),
        "config": {"workload": "upflow_ops", "describes": run_upflow_ops.__doc__.split("\n\n")[0].replace("\n    ", " "),
                   "levels_CHW": [list(v) for v in UPFLOW_LEVELS], "batch_per_gpu": B,
                   "roofline_kernel": "corr81_fwd, five launches, algorithmic bytes (2C+81)*H*W*4*B = %d" % corr_bytes},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ UPFlow network workload
def run_upflow_net(args, rank, world, local_rank):
    """one step = `UPFlow_net.forward_2_frame_v3` (UPFlow/model/upflow.py:580-665: feature pyramid, five decode
    levels x two directions with warp + feature normalisation + 81-channel correlation + dense estimator + dilated context net) on
    `pairs` synthetic 256x832 image pairs per GPU (BASELINE.json configs[4] shapes), inference, random MSRA-scale weights."""
    import torch
    import torch.distributed as dist

    from opticalflowscivis_b200 import ops
    from opticalflowscivis_b200.upflow.net import UPFlowNet

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    b = args.pairs or 8
    torch.manual_seed(1234)
    net = UPFlowNet().to(dev)
    net.enable_cuda_graphs(not args.no_train_graph)
    g = torch.Generator().manual_seed(1234 + rank)
    base = torch.nn.functional.avg_pool2d(torch.rand((b, 3, 256 + 16, 832 + 16), generator=g), 7, 1, 3)
    base = (base - base.mean()) / base.std() * 0.2
    im1, im2 = base[:, :, 8:-8, 8:-8].contiguous().to(dev), base[:, :, 8:-8, 2:-14].contiguous().to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(k=1):
        for _ in range(k):
            net.forward_2_frame_v3(im1, im2)

    step(args.warmup)
    n0 = ops.launch_count()
    with ClockSampler(local_rank) as clk:
        ms = _timed(step, args.steps, barrier)
        launches = ops.launch_count() - n0
        clk.keep_loaded(lambda: (step(5), torch.cuda.synchronize()))
    net.enable_cuda_graphs(False)                      # second pass, eager and event-instrumented: time per kernel class
    step(2)
    n0 = ops.launch_count()
    ms_eager = _timed(step, args.steps, barrier)
    launches = ops.launch_count() - n0                 # the graph replays exactly these launches
    ops.TIMER = timer = ops.LaunchTimer()
    ms_prof = _timed(step, args.steps, barrier)
    ops.TIMER = None
    classes = {k: {"launches": c // args.steps, "ms_per_step": t / args.steps} for k, (c, t) in timer.totals().items()}
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    if rank != 0:
        return
    line = {
        "metric": "UPFlow 256x832 flow pairs/sec (forward_2_frame_v3, both directions)", "value": world * b * args.steps / (ms / 1e3),
        "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "e2e": None,
        "gpu_launches": int(launches), "kernel_classes": classes, "ms_per_step_eager": ms_eager / args.steps,
        "cuda_graph": not args.no_train_graph, "clocks": clk.summary(),
        "config": {"workload": "upflow_net", "describes": run_upflow_net.__doc__.split("\n\n")[0].replace("\n    ", " "),
                   "spatial": [256, 832], "pairs_per_gpu_per_step": b,
                   "note": "convolutions on the tcgen05 engines (bf16), correlation / warps / normalisation / flow resizes in fp32 "
                           "libofsv kernels; layout changes and concatenations are torch copies; the call is replayed from a CUDA graph"},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ training-step workload
def run_train(args, rank, world, local_rank):
    """one step = `Model.update` (forward with the teacher block, losses, backward, gradient all-reduce, AdamW).
    train3d: the 3-D model (Flow-3D/model/RIFE.py:81-275, L1 + L1(teacher) + 0.1 distillation) on `pairs` synthetic 64^3 (img0, img1,
    gt) triplets per GPU — the volume size the reference trains on (Flow-3D/train.py:513-546 model names, `--batch_size` 15-30).
    train2d: the 2-D model (Flow-2D/model/RIFE.py:80-336, LapLoss + 0.01 distillation + 1e-5 photometric) on 160x224 triplets."""
    import torch
    import torch.distributed as dist

    from opticalflowscivis_b200 import ops
    from opticalflowscivis_b200.rife import Model2D, Model3D

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    nd = 3 if args.workload == "train3d" else 2
    sp = (64, 64, 64) if nd == 3 else (160, 224)
    n = args.pairs or (8 if nd == 3 else 16)
    torch.manual_seed(1234)
    model = (Model3D if nd == 3 else Model2D)(local_rank=local_rank if world > 1 else -1)
    model.enable_training_graph(not args.no_train_graph)
    g = torch.Generator().manual_seed(1234 + rank)
    pool = torch.nn.functional.avg_pool3d if nd == 3 else torch.nn.functional.avg_pool2d
    base = pool(torch.rand((n, 1) + tuple(v + 8 for v in sp), generator=g), 5, 1, 2)
    crop = lambda o: base[(slice(None), slice(None)) + tuple(slice(4, 4 + v) for v in sp[:-1]) + (slice(o, o + sp[-1]),)].contiguous().to(dev)  # noqa: E731
    img0, gt, img1 = crop(2), crop(4), crop(6)
    imgs = torch.cat((img0, img1), 1)
    losses = []
    LR = 3e-6          # the reference warms up linearly from 0 to 3e-4 over 2000 steps (Flow-3D/train.py:50-54): step 20 of that ramp

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(k=1):
        for _ in range(k):
            _, info = model.update(imgs, gt, learning_rate=LR, training=True) if nd == 3 else \
                model.update(imgs, gt, "droplet2d", learning_rate=LR, training=True)
            losses.append(info["loss_G"])

    step(args.warmup)
    tr = model._trainer
    n0 = ops.launch_count()
    with ClockSampler(local_rank) as clk:
        ms = _timed(step, args.steps, barrier)
        launches = ops.launch_count() - n0
        clk.keep_loaded(lambda: (step(3), torch.cuda.synchronize()))
    tr.allreduce_events = []
    # second pass, eager and event-instrumented: time per kernel class (and the eager step time next to the graphed one)
    model.enable_training_graph(False)
    step(2)
    n0 = ops.launch_count()
    ms_eager = _timed(step, args.steps, barrier)
    launches = ops.launch_count() - n0                    # the graph replays exactly these launches (counted where Python issues them)
    ops.TIMER = timer = ops.LaunchTimer()
    ms_prof = _timed(step, args.steps, barrier)
    ops.TIMER = None
    classes = {k: {"launches": c // args.steps, "ms_per_step": t / args.steps} for k, (c, t) in timer.totals().items()}
    ar = sum(a.elapsed_time(b) for a, b in tr.allreduce_events) / max(1, len(tr.allreduce_events))
    tr.allreduce_events = None
    lossv = [float(v) for v in losses]
    if world > 1:
        t = torch.tensor([ms, ar], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ar = float(t[0]), float(t[1])
    # context: the reference's own update in eager PyTorch on this GPU (oracle/train_ref.py is the pinned restatement), rank 0 only
    eager = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            from oracle.train_ref import TrainerRef
            torch.manual_seed(1234)
            orc = TrainerRef(nd)
            orc.flownet.to(dev)
            orc.optimG = torch.optim.AdamW(orc.flownet.parameters(), lr=1e-6, weight_decay=1e-3)
            eager = {"loss_G": []}
            for name, tf32 in (("fp32", False), ("tf32", True)):
                old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
                torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
                try:
                    for _ in range(2):
                        eager["loss_G"].append(float(orc.update(imgs, gt, learning_rate=LR, training=True)[1]["loss_G"]))
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for _ in range(3):
                        orc.update(imgs, gt, learning_rate=LR, training=True)
                    torch.cuda.synchronize()
                    eager[name + "_triplets_per_s"] = 3 * n / (time.perf_counter() - t0)
                finally:
                    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
        except Exception as e:  # noqa: BLE001
            eager = {"unavailable": str(e)[:200]}
    if rank != 0:
        return
    nparam = sum(p.numel() for p in model.flownet.parameters())
    macs = ifnet_macs(nd, sp)
    line = {
        "metric": ("Flow-3D Model.update training throughput (64^3 triplets/sec)" if nd == 3 else
                   "Flow-2D Model.update training throughput (160x224 triplets/sec)"), "value": world * n * args.steps / (ms / 1e3),
        "unit": "triplets/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "e2e": None,
        "allreduce_ms_per_step": ar, "allreduce_bytes": nparam * 4, "gpu_launches": int(launches),
        "loss_G_first_steps": lossv[:4], "loss_G_last": lossv[-1], "cuda_eager_reference": eager,
        "kernel_classes": classes, "ms_per_step_eager": ms_eager / args.steps, "train_graph": not args.no_train_graph, "clocks": clk.summary(),
        "config": {"workload": args.workload, "describes": run_train.__doc__.replace("\n    ", " "),
                   "spatial": list(sp), "triplets_per_gpu_per_step": n, "student_inference_macs_per_triplet": macs,
                   "note": "conv stacks forward/backward, warp forward/backward and AdamW run in libofsv; interpolate/cat/sigmoid/"
                           "blend/loss glue is torch autograd"},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="flow3d_droplet256", choices=list(WORKLOADS) + ["upflow_ops", "train3d", "train2d", "upflow_net"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--engine", default="auto", choices=["auto", "tc", "simt"])
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU per step (default: workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train-graph", action="store_true", help="train3d / upflow_net: enqueue every launch from Python instead of replaying from a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the ONE JSON line: NCCL writes its version banner to fd 1 when the communicator is created (at the first
        # collective), so stdout points at stderr until that has happened
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    try:
        {"upflow_ops": run_upflow_ops, "train3d": run_train, "train2d": run_train, "upflow_net": run_upflow_net}.get(args.workload, run_ours)(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
