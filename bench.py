#!/usr/bin/env python
"""bench.py - MC-dropout images/sec of the hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun)
    python bench.py --impl reference ...                     (the CPU oracle port on host cores)

A step = one pass of the hot path over one batch of synthetic BiFPN features:
  T MC-dropout samples of the class / box(+sigma) heads -> fused decode + MC moments ->
  global gaussian soft-NMS -> detections (udal_run, one C-ABI call).
Workload at every N: BASELINE.json configs[1] - EfficientDet-D0 at the KITTI shape 1280x384, T=10,
batch 64 per GPU (weak scaling: images are sharded, no collective on the data path).
`value` is timed with the inputs resident in HBM; `e2e` goes through the reference-facing Python
entry point with HOST buffers (pinned), H2D and D2H inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mc_dropout_images_per_sec_effdet_d0_T10"
UNIT = "images/s"
IMAGE_SIZE = (384, 1280)  # H, W  ("1280x384")
NUM_CLASSES = 8           # BASELINE.json: "KITTI 8-class label map"
T = 10
BATCH = 64


def workload_params(heads_mode="fp32"):
    import udal_b200 as u
    return u.hparams_config.get_detection_config(
        "efficientdet-d0", image_size=IMAGE_SIZE, num_classes=NUM_CLASSES, enable_softmax=True,
        loss_attenuation=True, mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05,
        mc_dropoutsamp=T, heads_mode=heads_mode)


def algorithmic_work(eng, batch):
    """SURVEY 8(d) formulas for this geometry (per step)."""
    P, F, R, A, C, Tn = eng.P, eng.F, eng.R, eng.A, eng.C, eng.T
    cc, cb = A * C, eng.box_channels
    dense = Tn * P * (2 * R * 2 * F * F + 2 * F * cc + 2 * F * cb)           # pointwise + predict 1x1
    depthwise = Tn * P * (2 * (R + 1) * 18 * F)                               # both heads
    k2_bytes = Tn * eng.N * (8 * 4 + C * 4) + eng.N * (3 * 16 + 4 + 4 + 2 * C * 4)
    # per tower layer, as launched (one kernel per layer and head): activations in + out, sepconv FLOPs
    act = P * F * 2  # one bf16 activation map of one (sample, image)
    layers = []
    for head, cout in (("class", cc), ("box", cb)):
        for layer in range(R + 1):
            predict = layer == R
            n_in = 1 if layer <= 1 else Tn           # layer 0 / 1 read sample-invariant inputs
            n_out = 1 if layer == 0 else Tn
            b_in = n_in * (P * F * 4 if layer == 0 else act)
            b_out = n_out * (P * cout * 4 if predict else act)
            fl = n_out * P * (2 * F * (cout if predict else F) + 18 * F)
            layers.append(dict(name="%s/%s" % (head, "predict" if predict else "layer%d" % layer),
                               bytes=float(batch) * (b_in + b_out), flops=float(batch) * fl))
    return dict(heads_flops=float(batch) * (dense + depthwise), heads_dense_flops=float(batch) * dense,
                decode_bytes=float(batch) * k2_bytes, layers=layers)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except OSError:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, reasons, mx = [], set(), None
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
    (profiles/ncu_traffic.json: {kernel function name: bytes}); {} if absent."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(path))
    except (OSError, ValueError):
        return {}


def lib_int(eng, name):
    import ctypes
    try:
        return ctypes.c_int.in_dll(eng.lib, name).value
    except ValueError:
        return 0


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


# -------------------------------------------------------------------------------------------------
# CPU oracle arm (bench.py --impl reference, and the cpu_baseline leg of the GPU arm)
# -------------------------------------------------------------------------------------------------
def cpu_port_step(images, seed=0):
    """The oracle port (torch-CPU conv heads + NumPy/C post-processing) on `images` images of the
    same workload; returns (seconds, segment times)."""
    import torch
    from oracle import heads_ref, ref_np
    p = ref_np.default_params(image_size=IMAGE_SIZE, num_classes=NUM_CLASSES, mc_dropoutsamp=T)
    levels = ref_np.level_shapes(p)
    A = ref_np.num_anchors_per_location(p)
    w = cpu_port_step.w
    if w is None:
        w = cpu_port_step.w = heads_ref.init_head_weights(64, 3, 5, A, NUM_CLASSES, True)
    feats = heads_ref.make_features(levels, images, 64, seed=1234 + seed)
    masks = heads_ref.make_masks(T, 5, 3, images, 64, 0.05, 0.05, seed=7 + seed)
    t0 = time.perf_counter()
    cls, box = heads_ref.heads_sample(feats, w, masks, 0.05, 0.05, T)
    t1 = time.perf_counter()
    ref_np.postprocess_global(p, cls, box, np.ones(images, np.float32))
    t2 = time.perf_counter()
    return t2 - t0, {"heads_s": t1 - t0, "post_s": t2 - t1, "threads": torch.get_num_threads()}


cpu_port_step.w = None


WORKLOAD = ("EfficientDet-D0 1280x384 (BASELINE configs[1]): BiFPN feats -> T=10 MC-dropout heads -> "
            "decode+moments (fused into the predict layers) -> global gaussian soft-NMS")


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    from oracle import build as oracle_build
    oracle_build.build()
    sample_images = 2
    for i in range(args.warmup):
        cpu_port_step(sample_images, i)
    times = []
    for i in range(args.steps):
        dt, seg = cpu_port_step(sample_images, 100 + i)
        times.append(dt)
    total = float(np.sum(times))
    value = sample_images * len(times) / total
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        # the GPU arm's config (same workload, classes, T, anchors); each CPU step is a bounded sample of it
        "config": {"workload": WORKLOAD, "batch_per_gpu": args.batch, "num_classes": NUM_CLASSES, "T": T,
                   "sample_images_per_step": sample_images, "heads_mode": "fp32 (torch-CPU conv oracle)",
                   "parallelism": "host cores of rank 0"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d images per step x %d steps of the same workload (torch-CPU conv heads x T "
                                   "+ NumPy/C post-processing oracle); TensorFlow 2.10 reference not installable"
                                   % (sample_images, args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# -------------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local_rank):
    """Best effort: run this rank (and first-touch its pinned staging buffers) on the CPU socket the GPU hangs off, so that
    the H2D copies of the end-to-end leg read local memory.  Returns the node or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # NVML prints an 8-digit domain, sysfs uses 4
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            bind_to_gpu_numa_node.original = allowed
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


bind_to_gpu_numa_node.original = None


def run_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    numa_node = bind_to_gpu_numa_node(local_rank)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        # NCCL announces its version on stdout when the communicator comes up: create it now, with fd 1 pointing at
        # stderr, so that stdout carries the one JSON line and nothing else
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    import udal_b200 as u

    p = workload_params(args.heads_mode)
    geom = u.engine.get_engine(p, device_id=local_rank)   # geometry only (the cached, weight-free engine)
    weights = u.synthetic.init_head_weights(geom.F, geom.R, len(geom.level_hw), geom.A, geom.C, True, seed=2024)  # SURVEY 8d seeds
    sampler = u.heads.HeadSampler(p, weights, device_id=local_rank)
    eng = sampler.engine                                   # the sampler's own context holds the weights
    ctx = eng.ctx
    L = len(eng.level_hw)
    batch = args.batch
    rng = np.random.default_rng(1234 + rank)
    # host (pinned) and device copies of the synthetic BiFPN features
    pinned = [u.device.PinnedArray((batch, h, w, eng.F)) for h, w in eng.level_hw]
    for pa in pinned:
        pa.array[...] = rng.standard_normal(pa.shape, dtype=np.float32)
    feats_dev = [ctx.to_device(pa.array) for pa in pinned]
    scales_host = np.ones(batch, np.float32)
    scales_dev = ctx.to_device(scales_host)
    ctx.sync()
    h2d = sum(pa.nbytes for pa in pinned) + scales_host.nbytes

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ---------------------------------------------------------
    out = None
    for i in range(args.warmup):
        out = eng.run(feats_dev, scales_dev, None, seed=i)
    barrier()
    launches0 = ctx.launch_count()
    clocks = ClockSampler(local_rank)
    clocks.start()
    step_ms = []
    t_all0 = time.perf_counter()
    ctx.timer_start()
    for i in range(args.steps):
        out = eng.run(feats_dev, scales_dev, None, seed=1000 + i)
    total_ms = ctx.timer_stop()
    barrier()
    wall_s = time.perf_counter() - t_all0
    launches = ctx.launch_count() - launches0
    total_ms = max_over_ranks(total_ms)
    d2h = sum(o.nbytes for o in out.values())

    # per-step latency distribution (separate loop: a sync per step)
    for i in range(min(args.steps, 10)):
        ctx.timer_start()
        eng.run(feats_dev, scales_dev, None, seed=2000 + i)
        step_ms.append(ctx.timer_stop())

    # ---- kernel-level timing on the launching stream ----------------------------------------
    cls_bufs, box_bufs = eng.head_output_buffers(batch)
    reps = max(2, min(args.steps, 5))
    eng.heads_sample(feats_dev, None, 1, out=(cls_bufs, box_bufs))
    ctx.sync()
    l0 = ctx.launch_count()
    ctx.timer_start()
    for i in range(reps):
        eng.heads_sample(feats_dev, None, i, out=(cls_bufs, box_bufs))
    heads_ms = ctx.timer_stop() / reps
    heads_launches = (ctx.launch_count() - l0) // reps
    # per-layer times of the kernels udal_run actually launches (serving configuration: predict layers fused
    # with K2), and of the stand-alone head sampler (predict layers writing the [T,...] outputs)
    def run_layers(i):
        ctx.sync()  # the previous run's NMS tail (post stream) would otherwise overlap the first layers
        return ctx.layer_times(lambda: eng.run(feats_dev, scales_dev, None, seed=40 + i))

    layer_ms = [run_layers(i) for i in range(3)]
    layer_ms = [float(np.median(x)) for x in zip(*layer_ms)] if layer_ms and layer_ms[0] else []
    layer_ms_unfused = [ctx.layer_times(lambda: eng.heads_sample(feats_dev, None, 40 + i, out=(cls_bufs, box_bufs)))
                        for i in range(3)]
    layer_ms_unfused = [float(np.median(x)) for x in zip(*layer_ms_unfused)] if layer_ms_unfused and layer_ms_unfused[0] else []
    pre = eng.decode_moments(cls_bufs, box_bufs, batch)
    ctx.timer_start()
    for i in range(reps):
        pre = None  # release to the pool first: no allocation inside the timed loop
        pre = eng.decode_moments(cls_bufs, box_bufs, batch)
    decode_ms = ctx.timer_stop() / reps
    ctx.timer_start()
    for i in range(reps):
        eng.nms_v5(pre["boxes"], pre["scores"])
    nms_ms = ctx.timer_stop() / reps
    ctx.timer_start()
    for i in range(reps):
        eng.postprocess_global(cls_bufs, box_bufs, batch, scales_dev.ptr)
    post_ms = ctx.timer_stop() / reps
    clock_info = clocks.stop()
    del pre

    # ---- end to end through the public entry point with host buffers --------------------------
    # every step copies its own inputs host->device (pinned) and its detections device->host inside
    # the timed region; two contexts alternate so that the copies of step i+1 overlap the kernels of
    # step i (PipelinedSampler).  Timed with the host clock around fully synchronised work.
    host_feats = [pa.array for pa in pinned]
    pipe = u.heads.PipelinedSampler(p, weights, device_id=local_rank, heads_mode=args.heads_mode, depth=4)
    for _ in pipe.map([host_feats] * 8, [scales_host] * 8, seed=1):  # warm-up: every context twice
        pass
    barrier()
    e2e_steps = max(24, args.steps)  # the first batch of a map() cannot hide its H2D copy: amortised over the run
    t0 = time.perf_counter()
    for det in pipe.map([host_feats] * e2e_steps, [scales_host] * e2e_steps, seed=3000):
        pass
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    barrier()
    e2e_ms = max_over_ranks(e2e_ms)
    # un-pipelined reference point: one blocking call per step
    sampler.detect(host_feats, scales_host, seed=1)
    ctx.timer_start()
    for i in range(3):
        sampler.detect(host_feats, scales_host, seed=10 + i)
    e2e_blocking_ms = ctx.timer_stop() / 3

    # ---- B=1 latency (p50 ms/img of the metric string) ----------------------------------------
    f1 = [f.slice0(0, 1) for f in feats_dev]
    s1 = scales_dev.slice0(0, 1)
    lat = []
    for i in range(13):
        ctx.timer_start()
        eng.run(f1, s1, None, seed=i)
        ms = ctx.timer_stop()
        if i >= 3:
            lat.append(ms)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    work = algorithmic_work(eng, batch)
    hbm_peak, tf_burst, tf_sustained, how = measured_peaks()
    ms_per_step = total_ms / args.steps
    value = world * batch * args.steps / (total_ms / 1e3)
    heads_tflops = work["heads_flops"] / (heads_ms / 1e3) / 1e12
    n_dom = max(1, heads_launches - 2)  # the tower layers (two small mask kernels excluded)
    # per-kernel rooflines: every kernel of the step against the bound that applies to it
    traffic = ncu_traffic()
    kernels = []
    kernel_names = {0: "heads_wide_kernel<64>", 1: "heads_l1_kernel", 2: "heads_ig_kernel<tower>"}
    fused_run = args.heads_mode != "fp32" and lib_int(eng, "udal_run_fused") and eng.C == 8 and eng.A == 9
    n_anchor = float(batch) * eng.N
    for i, ms in enumerate(layer_ms if args.heads_mode != "fp32" else []):
        lay = dict(work["layers"][i])
        r_idx = i % (eng.R + 1)
        kn = "heads_ig_kernel<predict>" if r_idx == eng.R else kernel_names.get(r_idx, "heads_ig_kernel<tower>")
        if r_idx == eng.R and fused_run:
            # predict layer fused with K2: reads the T activation maps, writes the per-anchor tensors only
            is_cls = i < eng.R + 1
            kn = "heads_fused_kernel<%s>" % ("class" if is_cls else "box")
            out_b = n_anchor * ((2 * eng.C * 4 + 8) if is_cls else 48)
            lay["bytes"] = float(batch) * eng.T * eng.P * eng.F * 2 + out_b
            lay["name"] += " + %s" % ("logit moments, argmax, sigmoid" if is_cls else "decode, box moments")
        gbs = lay["bytes"] / (ms / 1e3) / 1e9
        tfl = lay["flops"] / (ms / 1e3) / 1e12
        kernels.append({"kernel": kn, "what": lay["name"], "ms": ms, "bound": "hbm", "achieved_GBs": gbs,
                        "frac_of_hbm": gbs / hbm_peak, "algorithmic_TFLOPs": tfl, "frac_of_bf16": tfl / tf_sustained})
    kernels.append({"kernel": "topk_* + nms_v5_sorted_kernel", "what": "score pre-filter + global soft-NMS (one warp per image)",
                    "ms": nms_ms, "bound": "latency", "us_per_image": 1e3 * nms_ms / batch})
    # kernels of the stand-alone entry points (not launched by udal_run in the serving configuration)
    standalone = []
    for i, ms in enumerate(layer_ms_unfused if fused_run else []):
        if i % (eng.R + 1) == eng.R:
            lay = work["layers"][i]
            gbs = lay["bytes"] / (ms / 1e3) / 1e9
            standalone.append({"kernel": "heads_ig_kernel<predict>", "what": lay["name"] + " (HeadSampler.__call__)", "ms": ms,
                               "bound": "hbm", "achieved_GBs": gbs, "frac_of_hbm": gbs / hbm_peak})
    dec_gbs = work["decode_bytes"] / (decode_ms / 1e3) / 1e9
    dec = {"kernel": "decode_moments_kernel<%d,1>" % eng.T, "what": "decode + MC moments (postprocess.* entry points)",
           "ms": decode_ms, "bound": "hbm", "achieved_GBs": dec_gbs, "frac_of_hbm": dec_gbs / hbm_peak}
    (standalone if fused_run else kernels).append(dec)
    # dominant kernel = the kernel function with the largest total time inside one step
    totals = {}
    for k in kernels:
        if k["bound"] == "hbm":
            totals.setdefault(k["kernel"], []).append(k)
    dom_name = max(totals, key=lambda n: sum(k["ms"] for k in totals[n]))
    dom_launches = totals[dom_name]
    dominant = {"kernel": dom_name, "what": " | ".join(k["what"] for k in dom_launches),
                "ms": float(np.mean([k["ms"] for k in dom_launches])), "launches_per_step": len(dom_launches),
                "achieved_GBs": float(np.mean([k["achieved_GBs"] for k in dom_launches]))}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if args.heads_mode == "fp32" else "bf16", "data": "synthetic",
        "config": {
            "workload": WORKLOAD,
            "batch_per_gpu": batch, "num_classes": NUM_CLASSES, "T": T, "anchors": eng.N,
            "heads_mode": args.heads_mode, "parallelism": "image-sharded x%d, no collective" % world,
            "l2": "inputs larger than L2 (features %.0f MB + GBs of head activations per step)" % (h2d / 1e6),
        },
        "p50_ms_per_image_batch1": float(np.median(lat)),
        "p50_ms_per_step": float(np.median(step_ms)),
        "wall_s_timed_region": wall_s,
        "clocks": clock_info,
        "e2e": {"value": world * batch / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": e2e_steps,
                "how": "PipelinedSampler.map: HeadSampler.detect(host arrays) on 4 contexts, H2D of the next steps "
                       "overlaps the kernels of step i; host clock over fully synchronised work",
                "blocking_ms_per_step": e2e_blocking_ms, "numa_node": numa_node},
        "gpu_launches": int(launches),
        "roofline": {
            "kernel": dominant["kernel"], "what": dominant["what"], "bound": "hbm",
            "achieved": dominant["achieved_GBs"], "peak": hbm_peak, "unit": "GB/s",
            "frac": dominant["achieved_GBs"] / hbm_peak,
            "traffic": traffic.get(dominant["kernel"], traffic.get(dominant["kernel"].split("<")[0])),
            "launch_ms": dominant["ms"], "launches_per_step": dominant["launches_per_step"],
            "peak_source": how + " (HBM copy)",
            "note": "dominant kernel = largest total time per step; achieved = algorithmic bytes of one launch (the "
                    "layer's activations in + out, SURVEY 8d accounting) / CUDA-event time of that launch on the launching "
                    "stream, mean over its launches of the step; traffic = ncu dram bytes per launch (profiles/, null if "
                    "no capture for this kernel)",
        },
        "kernels": kernels,
        "standalone_kernels": standalone,
        "phases_ms": {"heads": heads_ms, "decode_moments": decode_ms, "nms_topk": nms_ms, "post_total": post_ms,
                      "heads_TFLOPs_algorithmic": heads_tflops, "heads_frac_of_bf16_sustained": heads_tflops / tf_sustained},
    }
    if not args.no_cpu_baseline:
        import torch
        from oracle import build as oracle_build
        oracle_build.build()
        if bind_to_gpu_numa_node.original:  # the CPU baseline uses every core of the box again
            os.sched_setaffinity(0, bind_to_gpu_numa_node.original)
        cores = len(os.sched_getaffinity(0)) or 1
        torch.set_num_threads(cores)
        cpu_port_step(1, 0)
        n_img, t_cpu, seg = 0, 0.0, {}
        while t_cpu < 12.0 and n_img < 16:
            dt, seg = cpu_port_step(2, 50 + n_img)
            t_cpu += dt
            n_img += 2
        line["cpu_baseline"] = {
            "value": n_img / t_cpu, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d images of the same workload in %.1f s (torch-CPU conv heads x T: %.2f s, NumPy/C "
                      "post-processing: %.2f s per 2 images)" % (n_img, t_cpu, seg.get("heads_s", 0), seg.get("post_s", 0)),
        }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="udal", choices=["udal", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--heads-mode", default=os.environ.get("UDAL_HEADS_MODE", "bf16"), help="bf16 (tcgen05 tensor cores, default) | fp32 (CUDA-core parity mode)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
