"""Development helper: per-layer CUDA-event times of udal_run at the bench shape (fused and unfused)."""
import ctypes
import sys

import numpy as np

sys.path.insert(0, ".")
import udal_b200 as u

batch = 64
p = u.hparams_config.get_detection_config(
    "efficientdet-d0", image_size=(384, 1280), num_classes=8, enable_softmax=True, loss_attenuation=True,
    mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05, mc_dropoutsamp=10, heads_mode=sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "fp16")
eng = u.engine.get_engine(p)
L = len(eng.level_hw)
eng.set_head_weights(u.synthetic.init_head_weights(eng.F, eng.R, L, eng.A, 8, True, seed=2024))
rng = np.random.default_rng(1)
feats = [eng.ctx.to_device(rng.standard_normal((batch, h, w, eng.F), dtype=np.float32)) for h, w in eng.level_hw]
scales = eng.ctx.to_device(np.ones(batch, np.float32))
fused = ctypes.c_int.in_dll(eng.lib, "udal_run_fused")
unst = ctypes.c_int.in_dll(eng.lib, "udal_nms_post_unstaged")
rsv = ctypes.c_int.in_dll(eng.lib, "udal_run_reserved_sms")
tl = ctypes.c_int.in_dll(eng.lib, "udal_run_debug_timeline")
tl.value = 1 if "--timeline" in sys.argv else 0
ovl = ctypes.c_int.in_dll(eng.lib, "udal_run_overlap")
for f, un, rs, ov in ((1, 0, 0, 1), (1, 0, 0, 0), (1, 0, 8, 1)):
    fused.value = f
    unst.value = un
    rsv.value = rs
    ovl.value = ov
    for i in range(3):
        eng.run(feats, scales, None, seed=i)
    eng.ctx.sync()
    t = eng.ctx.layer_times(lambda: eng.run(feats, scales, None, seed=9))
    eng.ctx.timer_start()
    for i in range(8):
        eng.run(feats, scales, None, seed=20 + i)
    # per-layer times of a run issued right behind another one (its tail overlaps these layers)
    lib = eng.lib
    lib.udal_profile_layers(eng.ctx.handle, 1)
    eng.run(feats, scales, None, seed=77)
    ms8 = (ctypes.c_float * 64)()
    n8 = ctypes.c_int(0)
    ms = eng.ctx.timer_stop() / 9
    lib.udal_get_layer_times(eng.ctx.handle, ms8, 64, ctypes.byref(n8))
    lib.udal_profile_layers(eng.ctx.handle, 0)
    print("   pipelined layer ms %s sum %.3f" % ([round(ms8[i], 3) for i in range(n8.value)], sum(ms8[i] for i in range(n8.value))))
    print("overlap=%d reserved=%d" % (ov, rs), end="  ")
    print("fused=%d  layer ms %s  sum %.3f  step %.3f ms" % (f, [round(x, 3) for x in t], sum(t), ms), flush=True)
