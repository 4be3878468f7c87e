#!/usr/bin/env python
"""bench.py - MC-dropout images/sec of the hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--config 1|2|4]     (N > 1: launched by torchrun)
    python bench.py --impl reference ...                                (the CPU oracle port on host cores)

A step = one pass of the hot path over one batch of synthetic BiFPN features:
  T MC-dropout samples of the class / box(+sigma) heads -> decode + MC moments (fused into the predict layers) ->
  global gaussian soft-NMS -> detections (udal_run, one C-ABI call).
Workload (`--config`, default 1 = the configuration BASELINE.json's metric is quoted on; weak scaling at every N:
images are sharded, no collective on the data path):
  1  configs[1]  EfficientDet-D0 at the KITTI shape 1280x384, C = 8, T = 10, batch 64 per GPU
  2  configs[2]  EfficientDet-D0 at the BDD100K shape 1280x720, C = 10, T = 20, batch 64 per GPU
  4  configs[4]  EfficientDet-D2 768x768, C = 10, T = 30, batch 16 per GPU + the auto-label threshold pass per batch
`value` is timed with the inputs resident in HBM; `e2e` goes through the reference-facing Python entry point with HOST
buffers (pinned), H2D and D2H inside the timed region; `sustained` repeats the device-resident loop for >= 2.5 s (the
1 kW power cap lowers the clocks after ~0.4 s); `roofline` follows SURVEY 8(d): ALGORITHMIC work of the reference
formulation / measured time against the measured peaks (MEASURED_PEAKS.json).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "images/s"
CONFIGS = {
    1: dict(tag="configs[1]", metric="mc_dropout_images_per_sec_effdet_d0_T10", model="efficientdet-d0",
            size=(384, 1280), C=8, T=10, batch=64, autolabel=False,
            workload="EfficientDet-D0 1280x384 (BASELINE configs[1]): BiFPN feats -> T=10 MC-dropout heads -> decode+moments "
                     "(fused into the predict layers) -> global gaussian soft-NMS"),
    2: dict(tag="configs[2]", metric="mc_dropout_images_per_sec_effdet_d0_bdd_T20", model="efficientdet-d0",
            size=(720, 1280), C=10, T=20, batch=64, autolabel=False,
            workload="EfficientDet-D0 1280x720, 10 classes (BASELINE configs[2]): BiFPN feats -> T=20 MC-dropout heads -> "
                     "decode+moments (fused) -> global gaussian soft-NMS; images sharded over the GPUs"),
    4: dict(tag="configs[4]", metric="mc_dropout_images_per_sec_effdet_d2_T30_autolabel", model="efficientdet-d2",
            size=(768, 768), C=10, T=30, batch=16, autolabel=True,
            workload="EfficientDet-D2 768x768, 10 classes (BASELINE configs[4]): BiFPN feats -> T=30 MC-dropout heads -> "
                     "decode+moments -> global soft-NMS -> auto-label threshold pass (entropy + relative aleatoric std); "
                     "4096 images = 512 per GPU on 8 GPUs, in batches of 16"),
}


def workload_params(cfg, heads_mode):
    import udal_b200 as u
    return u.hparams_config.get_detection_config(
        cfg["model"], image_size=cfg["size"], num_classes=cfg["C"], enable_softmax=True,
        loss_attenuation=True, mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05,
        mc_dropoutsamp=cfg["T"], heads_mode=heads_mode)


def algorithmic_work(eng, batch):
    """work of the REFERENCE formulation per step (executed work may be lower - layer 0 and the layer-1
    depthwise are hoisted out of the T loop - the denominators stay these)."""
    P, F, R, A, C, Tn = eng.P, eng.F, eng.R, eng.A, eng.C, eng.T
    cc, cb = A * C, eng.box_channels
    b = float(batch)
    dense = Tn * P * (2 * R * 2 * F * F + 2 * F * cc + 2 * F * cb)           # pointwise + predict 1x1, both heads
    depthwise = Tn * P * (2 * (R + 1) * 18 * F)                               # both heads
    k2_bytes = Tn * eng.N * (8 * 4 + C * 4) + eng.N * (3 * 16 + 4 + 4 + 2 * C * 4)
    layers = []
    for head, cout in (("class", cc), ("box", cb)):
        for layer in range(R + 1):
            predict = layer == R
            layers.append(dict(
                name="%s/%s" % (head, "predict" if predict else "layer%d" % layer),
                flops=b * Tn * P * (2 * F * (cout if predict else F) + 18 * F),      # reference formulation: every layer x T
                swish=0.0 if predict else b * Tn * P * F,
                dw_fma=b * Tn * P * 9 * F))
    return dict(
        heads_flops=b * (dense + depthwise), heads_dense_flops=b * dense,
        swish_evals=b * Tn * P * F * R * 2, dw_fmas=b * Tn * P * 9 * F * (R + 1) * 2,
        # fused K1 + K2 target design: features in + per-anchor outputs only
        step_bytes=b * (P * F * 4 + eng.N * (3 * 16 + 4 + 4 + 2 * C * 4)),
        decode_bytes=b * k2_bytes, layers=layers)


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

    def mark(self):
        return len(self.rows)

    def summary(self, start=0, stop=None):
        sm, reasons, mx, pw = [], set(), None, []
        for r in self.rows[start:stop]:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw) if pw else None}

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        return self.summary()


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
def cpu_port_step(cfg, images, seed=0):
    """The oracle port (torch-CPU conv heads + NumPy/C post-processing) on `images` images of the
    same workload; returns (seconds, segment times)."""
    import torch
    from oracle import heads_ref, ref_np
    fnum = 112 if cfg["model"] == "efficientdet-d2" else 64
    p = ref_np.default_params(image_size=cfg["size"], num_classes=cfg["C"], mc_dropoutsamp=cfg["T"])
    levels = ref_np.level_shapes(p)
    A = ref_np.num_anchors_per_location(p)
    w = cpu_port_step.w.get(cfg["tag"])
    if w is None:
        w = cpu_port_step.w[cfg["tag"]] = heads_ref.init_head_weights(fnum, 3, 5, A, cfg["C"], True)
    feats = heads_ref.make_features(levels, images, fnum, seed=1234 + seed)
    masks = heads_ref.make_masks(cfg["T"], 5, 3, images, fnum, 0.05, 0.05, seed=7 + seed)
    t0 = time.perf_counter()
    cls, box = heads_ref.heads_sample(feats, w, masks, 0.05, 0.05, cfg["T"])
    t1 = time.perf_counter()
    ref_np.postprocess_global(p, cls, box, np.ones(images, np.float32))
    t2 = time.perf_counter()
    return t2 - t0, {"heads_s": t1 - t0, "post_s": t2 - t1, "threads": torch.get_num_threads()}


cpu_port_step.w = {}


def run_reference(args, cfg):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = len(os.sched_getaffinity(0)) or 1
    torch.set_num_threads(cores)
    from oracle import build as oracle_build
    oracle_build.build()
    sample_images = 2 if cfg["tag"] == "configs[1]" else 1
    for i in range(args.warmup):
        cpu_port_step(cfg, sample_images, i)
    times, seg = [], {}
    for i in range(args.steps):
        dt, seg = cpu_port_step(cfg, sample_images, 100 + i)
        times.append(dt)
    total = float(np.sum(times))
    value = sample_images * len(times) / total
    line = {
        "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        # the GPU arm's config (same workload, classes, T, anchors); each CPU step is a bounded sample of it
        "config": {"workload": cfg["workload"], "batch_per_gpu": args.batch or cfg["batch"], "num_classes": cfg["C"], "T": cfg["T"],
                   "sample_images_per_step": sample_images, "heads_mode": "fp32 (torch-CPU conv oracle)",
                   "parallelism": "host cores of rank 0"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d image(s) per step x %d steps of the same workload (torch-CPU conv heads x T: %.2f s + "
                                   "NumPy/C post-processing: %.2f s per step, %d host cores); TensorFlow 2.10 reference not "
                                   "installable" % (sample_images, args.steps, seg.get("heads_s", 0), seg.get("post_s", 0), cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# -------------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local_rank):
    """Run this rank (and first-touch its pinned staging buffers) on the CPU socket the GPU hangs off, so that the H2D
    copies of the end-to-end leg read local memory.  Returns (node or None, why): the reason is reported in the JSON
    line when no binding took place (round 1 reported null without saying why)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # NVML prints an 8-digit domain, sysfs uses 4
            bus = bus[4:]
        path = "/sys/bus/pci/devices/%s/numa_node" % bus
        if not os.path.exists(path):
            return None, "no %s" % path
        node = int(open(path).read())
        if node < 0:
            return None, "sysfs reports numa_node %d for %s (single-node or virtualised topology)" % (node, bus)
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return None, "no allowed CPU on node %d (affinity mask of the container)" % node
        bind_to_gpu_numa_node.original = allowed
        os.sched_setaffinity(0, cpus)
        return node, "bound to %d CPUs of node %d" % (len(cpus), node)
    except Exception as e:  # noqa: BLE001
        return None, "%s: %s" % (type(e).__name__, e)


bind_to_gpu_numa_node.original = None


def run_gpu(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    numa_node, numa_why = bind_to_gpu_numa_node(local_rank)
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

    mode = args.heads_mode
    p = workload_params(cfg, mode)
    geom = u.engine.get_engine(p, device_id=local_rank)   # geometry only (the cached, weight-free engine)
    weights = u.synthetic.init_head_weights(geom.F, geom.R, len(geom.level_hw), geom.A, geom.C, True, seed=2024)  # SURVEY 8d seeds
    if cfg["autolabel"]:
        u.synthetic.autolabel_variant(weights)   # scores around min_score: the decision rule takes both branches
    sampler = u.heads.HeadSampler(p, weights, device_id=local_rank)
    eng = sampler.engine                                   # the sampler's own context holds the weights
    ctx = eng.ctx
    batch = args.batch or cfg["batch"]
    rng = np.random.default_rng(1234 + rank)
    # host (pinned) and device copies of the synthetic BiFPN features
    pinned = [u.device.PinnedArray((batch, h, w, eng.F)) for h, w in eng.level_hw]
    amp = u.synthetic.autolabel_amplitudes(batch)[:, None, None, None] if cfg["autolabel"] else np.float32(1.0)
    for pa in pinned:
        pa.array[...] = rng.standard_normal(pa.shape, dtype=np.float32) * amp
    feats_dev = [ctx.to_device(pa.array) for pa in pinned]
    scales_host = np.ones(batch, np.float32)
    scales_dev = ctx.to_device(scales_host)
    ctx.sync()
    h2d = sum(pa.nbytes for pa in pinned) + scales_host.nbytes
    labeler = None
    if cfg["autolabel"]:
        labeler = u.autolabel.AutoLabeler(dict(num_classes=cfg["C"], thr_sel_uncert=["ENT", "ALBOX"], calib_method_box=None,
                                               min_score=0.4), opt_params=[0.5, 0.5], opt_thrs=[0.5])

    def step(seed, feats=feats_dev, scales=scales_dev):
        det = eng.run(feats, scales, None, seed=seed)
        if labeler is not None:
            return labeler.decide((det["boxes"], det["scores"], det["classes"], det["valid"], det["logits"]))
        return det

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
        out = step(i)
    barrier()
    launches0 = ctx.launch_count()
    clocks = ClockSampler(local_rank)
    clocks.start()
    time.sleep(0.25)
    c0 = clocks.mark()
    t_all0 = time.perf_counter()
    ctx.timer_start()
    for i in range(args.steps):
        out = step(1000 + i)
    total_ms = ctx.timer_stop()
    barrier()
    wall_s = time.perf_counter() - t_all0
    c1 = clocks.mark()
    launches = ctx.launch_count() - launches0
    total_ms = max_over_ranks(total_ms)
    d2h = sum(o.nbytes for o in out.values())

    # ---- sustained: the same loop for >= 2.5 s (power-capped clocks), every rank at the same time ----
    ms_burst = total_ms / args.steps
    sus_steps = int(max(args.steps, math.ceil(2500.0 / max(ms_burst, 1e-3))))
    barrier()
    s0 = clocks.mark()
    ctx.timer_start()
    for i in range(sus_steps):
        step(5000 + i)
    sus_ms = ctx.timer_stop()
    barrier()
    s1 = clocks.mark()
    sus_ms = max_over_ranks(sus_ms)

    # per-step latency distribution (separate loop: a sync per step)
    step_ms = []
    for i in range(min(args.steps, 10)):
        ctx.timer_start()
        step(2000 + i)
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

    # per-layer times of the kernels udal_run actually launches, and of the stand-alone head sampler (predict layers
    # writing the [T,...] outputs)
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
    # the same kernel with decode_precision = "fp32" (closed form in fp32; own context on the same GPU, same buffers)
    eng32 = u.engine.Engine(dict(eng.params, decode_precision="fp32"), eng.cfg.device, "fp32")
    ctx.sync()
    pre32 = eng32.decode_moments(cls_bufs, box_bufs, batch)
    eng32.ctx.sync()
    eng32.ctx.timer_start()
    for i in range(reps):
        pre32 = None
        pre32 = eng32.decode_moments(cls_bufs, box_bufs, batch)
    decode32_ms = eng32.ctx.timer_stop() / reps
    pre32 = None
    ctx.timer_start()
    for i in range(reps):
        eng.nms_v5(pre["boxes"], pre["scores"])
    nms_ms = ctx.timer_stop() / reps
    ctx.timer_start()
    for i in range(reps):
        eng.postprocess_global(cls_bufs, box_bufs, batch, scales_dev.ptr)
    post_ms = ctx.timer_stop() / reps
    # worst case of the sequential soft-NMS: every candidate overlaps every selection (one cluster of jittered boxes), so
    # nearly every pop is a decay + re-insertion (the bench's own boxes are the benign case)
    wc_n = min(eng.N, 20000)
    wrng = np.random.default_rng(5)
    ctr = np.float32([200.0, 600.0]) + wrng.normal(0, 4.0, (batch, wc_n, 2)).astype(np.float32)
    hw_ = wrng.uniform(60, 90, (batch, wc_n, 2)).astype(np.float32)
    wboxes = ctx.to_device(np.concatenate([ctr - hw_ / 2, ctr + hw_ / 2], -1))
    wscores = ctx.to_device(wrng.uniform(0.05, 1.0, (batch, wc_n)).astype(np.float32))
    eng.nms_v5(wboxes, wscores)
    ctx.timer_start()
    for i in range(3):
        eng.nms_v5(wboxes, wscores)
    nms_worst_ms = ctx.timer_stop() / 3
    del pre, wboxes, wscores, cls_bufs, box_bufs

    # ---- end to end through the public entry point with host buffers --------------------------
    # every step copies its own inputs host->device (pinned) and its detections device->host inside
    # the timed region; `depth` contexts alternate so that the copies of step i+1 overlap the kernels of
    # step i (PipelinedSampler).  Timed with the host clock around fully synchronised work.
    host_feats = [pa.array for pa in pinned]
    e2e_blocking_ms = None
    e2e_f32 = None
    auto_fraction = None
    feat16 = args.feature_dtype == "f16" or (args.feature_dtype == "auto" and mode == "fp16" and eng.F == 64)
    e2e_steps = max(24, args.steps)
    if not cfg["autolabel"]:
        depth = 4 if cfg["tag"] == "configs[1]" else 2

        def measure_e2e(feats_h):
            pipe = u.heads.StreamingSampler(p, weights, device_id=local_rank, heads_mode=mode, depth=3)
            for _ in pipe.map([feats_h] * 6, [scales_host] * 6, seed=1):  # warm-up: every staging slot twice
                pass
            barrier()
            t0 = time.perf_counter()
            for det in pipe.map([feats_h] * e2e_steps, [scales_host] * e2e_steps, seed=3000):
                pass
            ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
            barrier()
            pipe.close()
            return max_over_ranks(ms)

        if feat16:
            # the feature maps cross the boundary as fp16 (udal_set_feature_format; what the reference's mixed_float16 GPU
            # graphs hand over): pinned fp16 copies of the same synthetic maps, made outside the timed region
            pinned16 = [u.device.PinnedArray(pa.shape, np.float16) for pa in pinned]
            for a, b in zip(pinned16, pinned):
                a.array[...] = b.array.astype(np.float16)
            host_feats16 = [pa.array for pa in pinned16]
            e2e_ms = measure_e2e(host_feats16)
            e2e_f32 = dict(value=world * batch / (measure_e2e(host_feats) / 1e3), unit=UNIT, h2d_bytes_per_step=h2d,
                           note="the same measurement with fp32 feature maps at the boundary")
            h2d = sum(pa.nbytes for pa in pinned16) + scales_host.nbytes
            host_feats = host_feats16
        else:
            e2e_ms = measure_e2e(host_feats)
        # un-pipelined reference point: one blocking call per step
        sampler.detect(host_feats, scales_host, seed=1)
        ctx.timer_start()
        for i in range(3):
            sampler.detect(host_feats, scales_host, seed=10 + i)
        e2e_blocking_ms = ctx.timer_stop() / 3
    else:
        # auto-label pass: features H2D, detections stay on the device, the per-image decisions come back
        def host_step(i):
            f = [ctx.to_device(x) for x in host_feats]
            o = step(7000 + i, f, scales_dev)
            return o["auto_label"].numpy()
        host_step(0)
        barrier()
        decisions = []
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            decisions.append(host_step(i))
        e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        auto_fraction = float(np.mean(np.concatenate(decisions)))
        barrier()
        e2e_ms = max_over_ranks(e2e_ms)
        d2h = batch  # one decision byte per image

    # ---- B=1 latency (p50 ms/img of the metric string) ----------------------------------------
    f1 = [f.slice0(0, 1) for f in feats_dev]
    s1_ = scales_dev.slice0(0, 1)
    lat = []
    for i in range(23):
        ctx.timer_start()
        eng.run(f1, s1_, None, seed=i)
        ms = ctx.timer_stop()
        if i >= 3:
            lat.append(ms)
    clock_all = clocks.stop()
    clock_info = clocks.summary(c0, max(c1, c0 + 1))
    if not clock_info["samples"]:
        clock_info = clock_all
    clock_sus = clocks.summary(s0, max(s1, s0 + 1))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    work = algorithmic_work(eng, batch)
    hbm_peak, tf_burst, tf_sustained, how = measured_peaks()
    ms_per_step = total_ms / args.steps
    value = world * batch * args.steps / (total_ms / 1e3)
    sm_mhz = clock_info.get("sm_mhz") or clock_info.get("sm_max_mhz") or 1965.0
    mufu_peak = 16.0 * 148 * sm_mhz * 1e6           # MUFU results / s  (16 / clk / SM)
    fma_peak = 128.0 * 148 * sm_mhz * 1e6           # fp32 FMAs / s     (128 / clk / SM)

    # ---- per-kernel rooflines: algorithmic work of the layer / CUDA-event time of its launch ----
    traffic = ncu_traffic()
    wide = eng.F != 64
    fused_run = bool(mode not in ("fp32", "fp32x3") and lib_int(eng, "udal_run_fused") and eng.C in (7, 8, 10) and eng.A == 9 and not wide)
    if mode == "fp32x3":
        x3 = "heads_wide_kernel<64,f32in,fp16 x 3>"
        names = {0: x3, 1: x3, "tower": x3, "predict": x3, "fused": x3 + " %s"}
    elif mode == "fp16":
        names = {0: "heads_wide_kernel<64,f32in>", 1: "heads_l1_kernel", "tower": "heads_dw_kernel<tower>",
                 "predict": "heads_dw_kernel<predict>", "fused": "heads_dwf_kernel<%s>"}
    else:
        names = {0: "heads_wide_kernel<64,f32in>", 1: "heads_l1_kernel", "tower": "heads_ig_kernel<tower>",
                 "predict": "heads_ig_kernel<predict>", "fused": "heads_fused_kernel<%s>"}

    def kernel_row(i, ms, fused):
        lay = dict(work["layers"][i])
        r_idx = i % (eng.R + 1)
        head = "class" if i < eng.R + 1 else "box"
        if wide:
            kn = "heads_wide_kernel<128>"
        elif r_idx == eng.R:
            kn = names["fused"] % head if fused else names["predict"]
        else:
            kn = names.get(r_idx, names["tower"])
        what = lay["name"]
        if r_idx == eng.R and fused:
            what += " + %s" % ("logit moments, argmax, sigmoid" if head == "class" else "decode, box moments")
        tfl = lay["flops"] / (ms / 1e3) / 1e12
        row = {"kernel": kn, "what": what, "ms": ms, "bound": "tensor", "algorithmic_TFLOPs": tfl,
               "frac_of_tensor": tfl / tf_burst,
               "frac_of_sfu": lay["swish"] / (ms / 1e3) / mufu_peak, "frac_of_fma": lay["dw_fma"] / (ms / 1e3) / fma_peak}
        row["binding"] = max(("tensor", row["frac_of_tensor"]), ("sfu", row["frac_of_sfu"]), ("fma", row["frac_of_fma"]),
                             key=lambda kv: kv[1])[0]
        return row

    kernels = [kernel_row(i, ms, fused_run) for i, ms in enumerate(layer_ms if mode != "fp32" else [])]
    kernels.append({"kernel": "nms_epoch_cta_kernel (+ nms_epoch_full_kernel for flagged images)",
                    "what": "in-CTA score select + global soft-NMS, one cooperative CTA per image",
                    "ms": nms_ms, "bound": "latency", "us_per_image": 1e3 * nms_ms / batch,
                    "worst_case_ms": nms_worst_ms,
                    "worst_case": "%d mutually overlapping candidates per image (every pop decays and re-inserts; every image is "
                                  "flagged and redone exactly over all candidates)" % wc_n})
    standalone = [kernel_row(i, ms, False) for i, ms in enumerate(layer_ms_unfused if fused_run else [])
                  if i % (eng.R + 1) == eng.R]
    dec_gbs = work["decode_bytes"] / (decode_ms / 1e3) / 1e9
    dec = {"kernel": "decode_moments_kernel<%d>" % eng.T, "what": "decode + MC moments (postprocess.* entry points)",
           "ms": decode_ms, "bound": "hbm", "achieved_GBs": dec_gbs, "frac_of_hbm": dec_gbs / hbm_peak}
    (standalone if fused_run else kernels).append(dec)
    dec32_gbs = work["decode_bytes"] / (decode32_ms / 1e3) / 1e9
    standalone.append({"kernel": "decode_moments_f32_kernel", "what": "decode + MC moments, decode_precision=fp32 (closed form in "
                       "fp32, 1e-4 contract)", "ms": decode32_ms, "bound": "hbm", "achieved_GBs": dec32_gbs,
                       "frac_of_hbm": dec32_gbs / hbm_peak})
    # dominant kernel = the kernel function with the largest total time inside one step
    totals = {}
    for k in kernels:
        if k["bound"] == "tensor":
            totals.setdefault(k["kernel"].split("<")[0], []).append(k)
    if totals:
        dom_fn = max(totals, key=lambda n: sum(k["ms"] for k in totals[n]))
        dom_launches = totals[dom_fn]
        dom_flops = float(np.mean([k["algorithmic_TFLOPs"] for k in dom_launches]))
        roof = {"kernel": dom_fn + " (" + ", ".join(sorted({k["kernel"] for k in dom_launches})) + ")",
                "what": " | ".join(k["what"] for k in dom_launches), "bound": "tensor",
                "achieved": dom_flops, "peak": tf_burst, "unit": "TFLOP/s", "frac": dom_flops / tf_burst,
                "traffic": traffic.get(dom_fn),
                "launch_ms": float(np.mean([k["ms"] for k in dom_launches])), "launches_per_step": len(dom_launches),
                "frac_of_sfu": float(np.mean([k["frac_of_sfu"] for k in dom_launches])),
                "frac_of_fma": float(np.mean([k["frac_of_fma"] for k in dom_launches])),
                "peak_source": how + " (cuBLAS bf16 burst; MUFU 16/clk/SM and fp32 FMA 128/clk/SM at the sampled SM clock)"}
    else:
        roof = {"kernel": dec["kernel"], "bound": "hbm", "achieved": dec_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": dec_gbs / hbm_peak, "traffic": traffic.get("decode_moments_kernel")}
    # step level (fused K1 + K2 target design): algorithmic FLOPs / bytes / SFU / FMA work of the whole step
    step_s = ms_per_step / 1e3
    step_tr = traffic.get("step_total")
    roof["step"] = {
        "ms": ms_per_step, "algorithmic_bytes": work["step_bytes"], "GBs": work["step_bytes"] / step_s / 1e9,
        "frac_of_hbm": work["step_bytes"] / step_s / 1e9 / hbm_peak, "traffic_bytes": step_tr,
        "traffic_over_algorithmic": (step_tr / work["step_bytes"]) if step_tr else None,
        "algorithmic_TFLOPs": work["heads_flops"] / step_s / 1e12,
        "frac_of_tensor": work["heads_flops"] / step_s / 1e12 / tf_burst,
        "frac_of_sfu": work["swish_evals"] / step_s / mufu_peak,
        "frac_of_fma": work["dw_fmas"] / step_s / fma_peak,
    }
    roof["note"] = ("SURVEY 8(d) accounting: achieved = ALGORITHMIC FLOPs of the reference formulation (every layer x T; pointwise "
                    "2*F*Cout + depthwise 18*F per output pixel) of one launch / CUDA-event time of that launch, mean over the "
                    "dominant kernel function's launches of a step; frac_of_sfu / frac_of_fma = algorithmic swish evaluations / "
                    "depthwise FMAs against the MUFU / fp32-FMA rates - the realistic ceilings at K = 64; step = the whole udal_run "
                    "against features-in + per-anchor-outputs bytes; traffic = ncu dram bytes (profiles/ncu_traffic.json, null if "
                    "no capture)")
    sus_value = world * batch * sus_steps / (sus_ms / 1e3)
    line = {
        "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"fp32": "f32", "fp32x3": "f32 (3 x fp16 tensor-core passes)", "bf16": "bf16", "fp16": "f16"}[mode], "data": "synthetic",
        "config": {
            "workload": cfg["workload"],
            "batch_per_gpu": batch, "num_classes": cfg["C"], "T": cfg["T"], "anchors": eng.N,
            "heads_mode": mode, "parallelism": "image-sharded x%d, no collective" % world,
            "l2": "inputs larger than L2 (features %.0f MB + GBs of head activations per step)" % (h2d / 1e6),
        },
        "sustained": {"value": sus_value, "unit": UNIT, "ms_per_step": sus_ms / sus_steps, "steps": sus_steps,
                      "seconds": sus_ms / 1e3, "clocks": clock_sus,
                      "frac_of_tensor_sustained": work["heads_flops"] / (sus_ms / sus_steps / 1e3) / 1e12 / tf_sustained},
        "p50_ms_per_image_batch1": float(np.median(lat)),
        "p50_ms_per_step": float(np.median(step_ms)),
        "wall_s_timed_region": wall_s,
        "clocks": clock_info,
        "auto_labelled_fraction": auto_fraction,
        "e2e": {"value": world * batch / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": e2e_steps,
                "how": ("StreamingSampler.map(host batches): per step the feature maps go pinned host -> device on a copy stream "
                        "(3 staging slots), udal_run, detections device -> pinned host behind the run's NMS tail, NumPy results "
                        "yielded in order; host clock over the whole loop, fully synchronised at both ends") if not cfg["autolabel"] else
                       "per batch: features H2D -> udal_run -> auto-label pass -> decisions D2H; host clock",
                "blocking_ms_per_step": e2e_blocking_ms, "h2d_GBs": h2d / (e2e_ms / 1e3) / 1e9,
                "feature_dtype": "f16" if (feat16 and not cfg["autolabel"]) else "f32",
                "feature_note": ("feature maps cross the boundary as fp16 (udal_set_feature_format; the reference's mixed_float16 "
                                 "exports): half the upload; fp32 maps: see fp32_features") if (feat16 and not cfg["autolabel"]) else None,
                "fp32_features": e2e_f32,
                "numa_node": numa_node, "numa": numa_why},
        "gpu_launches": int(launches),
        "roofline": roof,
        "kernels": kernels,
        "standalone_kernels": standalone,
        "phases_ms": {"heads": heads_ms, "heads_launches": heads_launches, "decode_moments": decode_ms, "nms_topk": nms_ms,
                      "nms_worst_case": nms_worst_ms, "post_total": post_ms},
    }
    if not args.no_cpu_baseline:
        import torch
        from oracle import build as oracle_build
        oracle_build.build()
        if bind_to_gpu_numa_node.original:  # the CPU baseline uses every core of the box again
            os.sched_setaffinity(0, bind_to_gpu_numa_node.original)
        cores = len(os.sched_getaffinity(0)) or 1
        torch.set_num_threads(cores)
        per = 2 if cfg["tag"] == "configs[1]" else 1
        cpu_port_step(cfg, 1, 0)
        n_img, t_cpu, seg = 0, 0.0, {}
        while t_cpu < 12.0 and n_img < 16:
            dt, seg = cpu_port_step(cfg, per, 50 + n_img)
            t_cpu += dt
            n_img += per
        line["cpu_baseline"] = {
            "value": n_img / t_cpu, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d images of the same workload in %.1f s on %d host cores (torch-CPU conv heads x T: %.2f s, NumPy/C "
                      "post-processing: %.2f s per %d image(s)); the TensorFlow 2.10 reference is not installable"
                      % (n_img, t_cpu, cores, seg.get("heads_s", 0), seg.get("post_s", 0), per),
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
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS), help="BASELINE.json configs[] index (1 = the metric's)")
    ap.add_argument("--batch", type=int, default=0, help="images per GPU and step (0: the config's)")
    ap.add_argument("--heads-mode", default=os.environ.get("UDAL_HEADS_MODE", "fp16"),
                    help="fp16 (tcgen05 + packed-fp16 depthwise, default) | bf16 (tcgen05 implicit GEMM) | fp32 (CUDA-core parity mode)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--feature-dtype", default="auto", choices=["auto", "f32", "f16"],
                    help="element type of the feature maps in the end-to-end (host buffer) measurement; auto = f16 with the fp16 "
                         "tensor-core heads (64 filters), f32 otherwise.  The device-resident `value` always reads fp32 maps")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        return run_reference(args, cfg)
    return run_gpu(args, cfg)


if __name__ == "__main__":
    sys.exit(main())
