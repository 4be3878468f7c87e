"""Development helper: udal_run at the bench shape under an integer tuning switch of libudal.so (default udal_tower_variant):
step time and per-layer CUDA-event times for each value.   python tools/time_variants.py [switch] [values] [batch]"""
import ctypes
import sys

import numpy as np

sys.path.insert(0, ".")
import udal_b200 as u

name = sys.argv[1] if len(sys.argv) > 1 else "udal_tower_variant"
values = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,1,2,3").split(",")]
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 64
p = u.hparams_config.get_detection_config(
    "efficientdet-d0", image_size=(384, 1280), num_classes=8, enable_softmax=True, loss_attenuation=True,
    mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05, mc_dropoutsamp=10, heads_mode="fp16")
eng = u.engine.get_engine(p)
eng.set_head_weights(u.synthetic.init_head_weights(eng.F, eng.R, len(eng.level_hw), eng.A, 8, True, seed=2024))
rng = np.random.default_rng(1)
feats = [eng.ctx.to_device(rng.standard_normal((batch, h, w, eng.F), dtype=np.float32)) for h, w in eng.level_hw]
scales = eng.ctx.to_device(np.ones(batch, np.float32))
sw = ctypes.c_int.in_dll(eng.lib, name)
default = sw.value
for v in values + values:
    sw.value = v
    for i in range(3):
        eng.run(feats, scales, None, seed=i)
    eng.ctx.sync()
    t = eng.ctx.layer_times(lambda: eng.run(feats, scales, None, seed=9))
    eng.ctx.sync()
    eng.ctx.timer_start()
    for i in range(20):
        eng.run(feats, scales, None, seed=20 + i)
    ms = eng.ctx.timer_stop() / 20
    print("%s=%d: %.3f ms/step = %.0f images/s; layers %s sum %.3f" % (name, v, ms, batch / ms * 1e3, [round(x, 3) for x in t], sum(t)))
sw.value = default
