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
}


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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)
        return False

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_reference_rate(nd, sp, pairs, steps, warmup):
    """The reference's CPU path for the same workload: oracle/ifnet_ref.py (bit-identical restatement of the reference
    modules, pinned in tests/golden) with all host threads.  Returns (pairs_per_s, seconds_per_step, threads)."""
    import torch
    from oracle.ifnet_ref import ModelRef
    from opticalflowscivis_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(1234)
    m = ModelRef(nd).eval()
    if nd == 3:
        a, _, b = synth.droplet3d_u8(pairs, sp[0])
        img0, img1 = torch.from_numpy(a).float() / 255.0, torch.from_numpy(b).float() / 255.0
    else:
        a, _, b = synth.droplet2d(pairs, *sp)
        img0, img1 = torch.from_numpy(a), torch.from_numpy(b)
    for _ in range(warmup):
        m.inference(img0, img1)
    t0 = time.perf_counter()
    for _ in range(steps):
        m.inference(img0, img1)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return pairs / dt, dt, torch.get_num_threads()


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    nd, sp, pairs, desc = WORKLOADS[args.workload]
    full_vox = 1
    for s in sp:
        full_vox *= s
    # bounded sample per step: a 128^3 sub-volume pair (1/8 of a 256^3 pair) for the 3-D headline workload
    if args.workload == "flow3d_droplet256":
        s_sp, s_pairs, frac, sample = (128, 128, 128), 1, 1.0 / 8.0, "one 128^3 sub-volume pair per step = 1/8 of a 256^3 pair"
    elif args.workload == "flow3d_rect128":
        s_sp, s_pairs, frac, sample = (64, 64, 64), 1, 1.0 / 8.0, "one 64^3 sub-volume pair per step = 1/8 of a 128^3 pair"
    else:
        s_sp, s_pairs, frac, sample = sp, 8, 8.0, "8 pairs of 160x224 per step"
    rate, dt, threads = cpu_reference_rate(nd, s_sp, s_pairs, args.steps, min(args.warmup, 1))
    value = frac / dt if nd == 3 else rate
    unit = "pairs/s"
    line = {
        "impl": "reference", "metric": metric_name(args.workload), "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "describes": desc, "spatial": list(sp), "pairs_per_gpu_per_step": pairs},
        "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def metric_name(workload):
    return {"flow3d_droplet256": "256^3 volume-pair interps/sec", "flow3d_rect128": "128^3 volume-pair interps/sec",
            "flow2d_droplet": "160x224 frame-pair interps/sec"}[workload]


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from opticalflowscivis_b200 import ops, synth
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

    # synthetic inputs: member seed = 1234 + global pair index (SURVEY.md §8d cfg 4)
    if nd == 3:
        a, _, b = synth.droplet3d_u8(pairs, sp[0], seed=1234 + rank * pairs)
    else:
        a, _, b = synth.droplet2d(pairs, *sp, seed=1234 + rank * pairs)
    h0, h1 = torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()      # 3-D: uint8 host volumes; 2-D: fp32
    as_f32 = (lambda t: t.float().div_(255.0)) if nd == 3 else (lambda t: t)
    d0, d1 = as_f32(h0.to(dev)), as_f32(h1.to(dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        return model.inference(d0, d1)

    for _ in range(args.warmup):
        step_resident()
    barrier()

    # ---- timed region 1 (`value`): inputs resident in HBM, nothing but the K inference calls between the two events
    n0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        e0.record()
        for _ in range(args.steps):
            step_resident()
        e1.record()
        barrier()
    launches = ops.launch_count() - n0
    ms = e0.elapsed_time(e1)

    # ---- profile pass (the same K steps again): CUDA events around every launch, per kernel class, on the launching stream
    timer = ops.LaunchTimer()
    ops.TIMER = timer
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record()
    for _ in range(args.steps):
        step_resident()
    e5.record()
    torch.cuda.synchronize()
    ops.TIMER = None
    ms_prof = e4.elapsed_time(e5)
    classes = timer.totals()

    # ---- standalone warp kernel (the "warp HBM GB/s vs peak" half of the metric): a1 / a2 on the workload's shape
    g = torch.Generator(device="cpu").manual_seed(7)
    wflow = (torch.randn((pairs, nd) + tuple(max(1, s // 8) for s in sp), generator=g) * 2.0).to(dev)
    wflow = torch.nn.functional.interpolate(wflow, size=tuple(sp), mode="trilinear" if nd == 3 else "bilinear").contiguous()
    warp_fn = ops.warp3d if nd == 3 else ops.warp2d
    for _ in range(3):
        warp_fn(d0, wflow)
    w0e, w1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wreps = 10
    w0e.record()
    for _ in range(wreps):
        warp_fn(d0, wflow)
    w1e.record()
    torch.cuda.synchronize()
    warp_ms = w0e.elapsed_time(w1e) / wreps
    del wflow

    # ---- timed region 2 (`e2e`): pinned host pairs -> H2D -> Model.inference -> D2H of the interpolated volume through the
    #      package's streaming front end (pipeline.StreamedInterpolator: upload / compute / download on three streams)
    streamer = StreamedInterpolator(model, dev)
    nbuf = 3
    out_host = [torch.empty((pairs, 1) + tuple(sp), dtype=torch.float32).pin_memory() for _ in range(nbuf)]

    def run_e2e(k):
        for _ in streamer.run(((h0, h1) for _ in range(k)), (out_host[i % nbuf] for i in range(k))):
            pass

    run_e2e(min(args.warmup, 3))
    barrier()
    h2d0, d2h0 = streamer.h2d_bytes, streamer.d2h_bytes
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    run_e2e(args.steps)            # returns when the last result has landed in host memory
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    h2d_step = (streamer.h2d_bytes - h2d0) // args.steps
    d2h_step = (streamer.d2h_bytes - d2h0) // args.steps

    # ---- the same end-to-end path with the result exported as bytes on the device, `(merged * 255).byte()` — what the
    #      reference's inference driver does before its `.cpu()` (Flow-3D/inference_img.py:105).  Reported NEXT TO `e2e` (which
    #      stays the fp32 download): with 8 ranks sharing one host, the 268 MB/step/rank fp32 download is what halves `e2e`.
    ms_e2e_u8 = d2h_u8_step = None
    try:
        streamer8 = StreamedInterpolator(model, dev, out_u8=True)
        out_host8 = [torch.empty((pairs, 1) + tuple(sp), dtype=torch.uint8).pin_memory() for _ in range(nbuf)]

        def run_e2e8(k):
            for _ in streamer8.run(((h0, h1) for _ in range(k)), (out_host8[i % nbuf] for i in range(k))):
                pass

        run_e2e8(min(args.warmup, 3))
        barrier()
        d8 = streamer8.d2h_bytes
        e6, e7 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e6.record()
        run_e2e8(args.steps)
        e7.record()
        barrier()
        ms_e2e_u8 = e6.elapsed_time(e7)
        d2h_u8_step = (streamer8.d2h_bytes - d8) // args.steps
    except Exception as e:  # noqa: BLE001  (the extra figure must never take the bench line down)
        print(f"[bench rank {rank}] e2e_u8 skipped: {e!r}", file=sys.stderr, flush=True)
        ms_e2e_u8 = None

    if world > 1:
        t = torch.tensor([ms, ms_e2e, ms_e2e_u8 if ms_e2e_u8 is not None else float("inf")], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        ms_e2e_u8 = float(t[2]) if float(t[2]) != float("inf") else None
    if rank != 0:
        return

    peaks = load_peaks()
    total_pairs = pairs * world * args.steps
    value = total_pairs / (ms / 1e3)
    e2e_value = total_pairs / (ms_e2e / 1e3)
    vox = 1
    for s in sp:
        vox *= s
    flops_pair = 2.0 * ifnet_macs(nd, sp)
    conv_cls = [k for k in classes if k.startswith("conv_")]
    conv_ms = sum(classes[k][1] for k in conv_cls)
    conv_n = sum(classes[k][0] for k in conv_cls)
    share = {k: round(v[1] / ms_prof, 4) for k, v in classes.items()}
    dominant = max(classes, key=lambda k: classes[k][1]) if classes else None
    # tensor roofline of the conv engine: algorithmic FLOPs of all conv launches of the profile pass / their summed time
    conv_tf = flops_pair * pairs * args.steps / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else None
    roofline = {"kernel": "+".join(sorted(conv_cls)), "bound": "tensor", "achieved": conv_tf, "peak": peaks["tf_sustained"],
                "unit": "TFLOP/s", "frac": (conv_tf / peaks["tf_sustained"]) if conv_tf else None, "traffic": None,
                "peak_source": peaks["src"] + " (sustained bf16)", "launches": conv_n, "share_of_step": round(conv_ms / ms_prof, 4),
                "algorithmic_flops_per_pair": flops_pair}
    # HBM roofline of the standalone warp kernel: (nd flow + 1 src + 1 out) * 4 B = 20 B/voxel in 3-D, 16 B/px in 2-D
    warp_bytes = (nd + 2) * 4 * vox * pairs
    warp_gbs = warp_bytes / (warp_ms / 1e3) / 1e9
    # DRAM traffic per launch from the committed `ncu --set full` capture of this kernel (profiles/r01q_ncu_full_warp3d_slab_4x256.csv:
    # dram__bytes_read 1118.0 MB + write 254.1 MB for a 4 x 256^3 launch = 279.5 + 63.5 MB per volume; algorithmic 335.5 MB),
    # scaled to this launch's volume count
    warp_traffic = (279.5e6 + 63.5e6) * pairs if (nd == 3 and tuple(sp) == (256, 256, 256)) else None
    slab = nd == 3 and sp[0] == sp[1] == sp[2] and sp[0] % 32 == 0      # ofsv_warp3d_f32 picks the TMA slab kernel on such volumes
    roofline_warp = {"kernel": "warp3d_slab_kernel" if slab else "warp%dd_kernel" % nd, "bound": "hbm", "achieved": warp_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": warp_gbs / peaks["hbm_gbs"], "traffic": warp_traffic, "peak_source": peaks["src"],
                     "algorithmic_bytes_per_launch": warp_bytes, "ms_per_launch": warp_ms, "timed": "standalone, 10 launches"}
    # HBM roofline of the fused block output stage (3-D): state read/write, img gathers, merged/mask, next block's input
    roofline_stage = None
    bs = classes.get("block_stage")
    if bs and nd == 3:
        per_vox = [32 + 8 + 32 / 8.0,                  # block0 -> 1: write state, read imgs, pooled bf16 input of block1 (head0 is L2-resident)
                   32 + 32 + 8 + 32 + 4,               # block1 -> 2: read + write state, imgs, full-res bf16 input of block2, head1 (32 B / 8 voxels)
                   32 + 8 + 8]                         # final: read the state (accumulated by the head conv), imgs, write merged + mask
        stage_bytes = sum(per_vox) * vox * pairs * args.steps
        sg = stage_bytes / (bs[1] / 1e3) / 1e9
        # ncu --set full (profiles/r01n_ncu_full_block2_flow3d_256.csv, r01n_launches_flow3d_256.csv): the three stage launches of
        # one 256^3 pair move 1.77 GB of DRAM reads + 1.68 GB of writes (algorithmic: 3.36 GB)
        stage_traffic = 3.45e9 * pairs if tuple(sp) == (256, 256, 256) else None
        roofline_stage = {"kernel": "block_stage_3d_kernel", "bound": "hbm", "achieved": sg, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                          "frac": sg / peaks["hbm_gbs"], "traffic": stage_traffic, "traffic_unit": "bytes per step (3 launches)",
                          "peak_source": peaks["src"], "launches": bs[0],
                          "share_of_step": round(bs[1] / ms_prof, 4), "algorithmic_bytes_per_voxel_per_scale": per_vox}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        with contextlib.redirect_stdout(io.StringIO()):
            if args.workload == "flow3d_droplet256":
                # bounded sample: a 128^3 sub-volume pair = 1/8 of the voxels (and conv FLOPs) of one 256^3 pair
                rate, dt, thr = cpu_reference_rate(3, (128, 128, 128), 1, 2, 1)
                rate, sample = rate / 8.0, "128^3 sub-volume pair (1/8 of a 256^3 pair) x 2 calls after 1 warm-up; rate scaled by 1/8"
            elif args.workload == "flow3d_rect128":
                rate, dt, thr = cpu_reference_rate(3, (128, 128, 128), 1, 2, 1)
                sample = "one 128^3 pair x 2 calls after 1 warm-up"
            else:
                rate, dt, thr = cpu_reference_rate(2, sp, 64, 3, 1)
                sample = "64 pairs of 160x224 x 3 calls after 1 warm-up"
        cpu_baseline = {"value": rate, "unit": "pairs/s", "cores": thr, "kind": "port", "sample": sample}

    line = {
        "metric": metric_name(args.workload), "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": args.workload, "describes": desc, "spatial": list(sp), "pairs_per_gpu_per_step": pairs,
                   "precision": args.precision, "conv_engine": model.flownet._engine(),
                   "l2": "working set per step (>1 GB) exceeds the 126 MB L2; no explicit flush",
                   "weights": "random init, seed 1234"},
        "clocks": clk.summary(),
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d_step), "d2h_bytes_per_step": int(d2h_step),
                "ms_per_step": ms_e2e / args.steps,
                "path": "pipeline.StreamedInterpolator(model).run(pinned host pairs): H2D / Model.inference / D2H on three streams"},
        "e2e_u8": None if ms_e2e_u8 is None else {
            "value": total_pairs / (ms_e2e_u8 / 1e3), "unit": "pairs/s", "h2d_bytes_per_step": int(h2d_step),
            "d2h_bytes_per_step": int(d2h_u8_step), "ms_per_step": ms_e2e_u8 / args.steps,
            "path": "as e2e, result downloaded as (merged*255).byte() computed on the device (Flow-3D/inference_img.py:105)"},
        "gpu_launches": int(launches),
        "roofline": roofline, "roofline_warp": roofline_warp, "roofline_stage": roofline_stage, "kernel_time_share": share,
        "dominant_kernel_class": dominant, "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="flow3d_droplet256", choices=list(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--engine", default="auto", choices=["auto", "tc", "simt"])
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU per step (default: workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        with contextlib.redirect_stdout(sys.stderr):
            pass
        run_reference_arm(args, rank, world)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "ERROR")      # keep stdout to the one JSON line (NCCL prints its version banner there)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
