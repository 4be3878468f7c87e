"""Development helper: per-layer times of the wide (F = 112) head kernels under the udal_wide_debug switches."""
import ctypes
import sys

import numpy as np

sys.path.insert(0, ".")
import udal_b200 as u

batch, T, C = 8, 30, 10
p = u.hparams_config.get_detection_config(
    "efficientdet-d2", image_size=(768, 768), num_classes=C, enable_softmax=True, loss_attenuation=True,
    mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05, mc_dropoutsamp=T, heads_mode="bf16")
eng = u.engine.get_engine(p)
eng.set_head_weights(u.synthetic.init_head_weights(eng.F, eng.R, len(eng.level_hw), eng.A, C, True, seed=2024))
rng = np.random.default_rng(1)
feats = [eng.ctx.to_device(rng.standard_normal((batch, h, w, eng.F), dtype=np.float32)) for h, w in eng.level_hw]
cls, box = eng.head_output_buffers(batch)
dbg = ctypes.c_int.in_dll(eng.lib, "udal_wide_debug")
for d in (0, 1, 2, 3):
    dbg.value = d
    for i in range(2):
        eng.heads_sample(feats, None, i, out=(cls, box))
    t = eng.ctx.layer_times(lambda: eng.heads_sample(feats, None, 9, out=(cls, box)))
    print("debug=%d (1: no depthwise math, 2: no epilogue)  layer ms %s" % (d, [round(x, 3) for x in t]), flush=True)
dbg.value = 0
