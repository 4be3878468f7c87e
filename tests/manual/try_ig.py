"""Debug helper: run the bf16 head sampler with the implicit-GEMM kernels (predictions by per-thread
row stores or through the staged TMA store),
report the error against the CPU oracle and the per-layer CUDA-event times at the bench shape."""
import ctypes
import sys

import numpy as np

sys.path.insert(0, ".")
import udal_b200 as u
from oracle import heads_ref

lib = u._lib.load()
use_ig = ctypes.c_int.in_dll(lib, "udal_heads_tc_use_ig")
tma_store = ctypes.c_int.in_dll(lib, "udal_ig_tma_store")
debug = ctypes.c_int.in_dll(lib, "udal_ig_debug")
VARIANTS = (0, 1)


def layer_times(eng, fn):
    lib.udal_profile_layers(eng.ctx.handle, 1)
    fn()
    ms = (ctypes.c_float * 64)()
    n = ctypes.c_int(0)
    lib.udal_get_layer_times(eng.ctx.handle, ms, 64, ctypes.byref(n))
    lib.udal_profile_layers(eng.ctx.handle, 0)
    return [round(ms[i], 4) for i in range(n.value)]


def parity(size, C, T, batch):
    p = u.hparams_config.get_detection_config(
        "efficientdet-d0", image_size=size, num_classes=C, enable_softmax=True, loss_attenuation=True,
        mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05, mc_dropoutsamp=T, heads_mode="bf16")
    eng = u.engine.get_engine(p)
    L = len(eng.level_hw)
    w = heads_ref.init_head_weights(eng.F, eng.R, L, eng.A, C, True, seed=9, randomize_bn=True)
    feats = heads_ref.make_features(eng.level_hw, batch, eng.F, seed=11)
    masks = heads_ref.make_masks(T, L, eng.R, batch, eng.F, 0.05, 0.05, seed=5)
    rcls, rbox = heads_ref.heads_sample(feats, w, masks, 0.05, 0.05, T)
    for tv in VARIANTS:
        tma_store.value = tv
        sampler = u.heads.HeadSampler(p, w)
        cls, box = sampler(feats, masks=masks)
        err = max(max(float(np.abs(a - b).max()) for a, b in zip(cls, rcls)),
                  max(float(np.abs(a - b).max()) for a, b in zip(box, rbox)))
        print("size=%s C=%d tma_store=%d  max abs err %.5f" % (size, C, tv, err), flush=True)


def timing(batch=64):
    p = u.hparams_config.get_detection_config(
        "efficientdet-d0", image_size=(384, 1280), num_classes=8, enable_softmax=True, loss_attenuation=True,
        mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05, mc_dropoutsamp=10, heads_mode="bf16")
    eng = u.engine.get_engine(p)
    L = len(eng.level_hw)
    w = heads_ref.init_head_weights(eng.F, eng.R, L, eng.A, 8, True, seed=2024)
    rng = np.random.default_rng(1)
    feats = [eng.ctx.to_device(rng.standard_normal((batch, h, ww, eng.F), dtype=np.float32)) for h, ww in eng.level_hw]
    out = eng.head_output_buffers(batch)
    for tv in VARIANTS:
        tma_store.value = tv
        eng.set_head_weights(w)
        for i in range(2):
            eng.heads_sample(feats, None, i, out=out)
        eng.ctx.sync()
        t = layer_times(eng, lambda: eng.heads_sample(feats, None, 7, out=out))
        print("tma_store=%d  layer ms (class L0..predict, box L0..predict): %s  total %.3f"
              % (tv, t, sum(t)), flush=True)
    for d in (1, 2, 4, 3, 5, 6, 7):
        debug.value = d
        eng.heads_sample(feats, None, 1, out=out)
        eng.ctx.sync()
        t = layer_times(eng, lambda: eng.heads_sample(feats, None, 7, out=out))
        print("debug=%d (1: one tap, 2: no epilogue math/stores, 4: no TMA loads)  %s" % (d, t), flush=True)
    debug.value = 0


if __name__ == "__main__":
    if "--no-parity" not in sys.argv:
        parity((64, 96), 8, 3, 2)
        parity((256, 256), 8, 6, 3)   # 234 work items: every CTA of the persistent kernels loops
        parity((40, 200), 8, 2, 3)
        parity((64, 96), 7, 2, 2)
    if "--time" in sys.argv:
        b = [int(a.split("=")[1]) for a in sys.argv if a.startswith("--batch=")]
        timing(b[0] if b else 64)
