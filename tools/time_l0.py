import ctypes, sys
import numpy as np
sys.path.insert(0, ".")
import udal_b200 as u
batch = 64
p = u.hparams_config.get_detection_config(
    "efficientdet-d0", image_size=(384, 1280), num_classes=8, enable_softmax=True, loss_attenuation=True,
    mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05, mc_dropoutsamp=10, heads_mode="bf16")
eng = u.engine.get_engine(p)
eng.set_head_weights(u.synthetic.init_head_weights(eng.F, eng.R, len(eng.level_hw), eng.A, 8, True, seed=2024))
rng = np.random.default_rng(1)
feats = [eng.ctx.to_device(rng.standard_normal((batch, h, w, eng.F), dtype=np.float32)) for h, w in eng.level_hw]
scales = eng.ctx.to_device(np.ones(batch, np.float32))
dbg = ctypes.c_int.in_dll(eng.lib, "udal_wide_debug")
for d in (0, 1, 2, 3):
    dbg.value = d
    for i in range(2):
        eng.run(feats, scales, None, seed=i)
    eng.ctx.sync()
    t = eng.ctx.layer_times(lambda: eng.run(feats, scales, None, seed=9))
    print("debug=%d layer0 ms %.3f %.3f" % (d, t[0], t[4]), flush=True)
dbg.value = 0
