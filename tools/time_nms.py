"""Development helper: global soft-NMS time on the per-anchor tensors of a synthetic udal_run (bench geometry), cycle
breakdown of the cooperative CTA kernel (udal_nms_debug) and the round-1 one-warp path beside it.

    python tools/time_nms.py [H W C T B]
"""
import ctypes
import sys

import numpy as np

sys.path.insert(0, ".")
import udal_b200 as u

H, W, C, T, batch = [int(x) for x in sys.argv[1:6]] if len(sys.argv) > 5 else (384, 1280, 8, 10, 64)
p = u.hparams_config.get_detection_config(
    "efficientdet-d0", image_size=(H, W), num_classes=C, enable_softmax=True, loss_attenuation=True,
    mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05, mc_dropoutsamp=T, heads_mode="fp16")
eng = u.engine.get_engine(p)
eng.set_head_weights(u.synthetic.init_head_weights(eng.F, eng.R, len(eng.level_hw), eng.A, C, True, seed=2024))
rng = np.random.default_rng(1)
feats = [eng.ctx.to_device(rng.standard_normal((batch, h, w, eng.F), dtype=np.float32)) for h, w in eng.level_hw]
pre = eng.run_prenms(feats, None, seed=3)
boxes, scores = pre["boxes"], pre["scores"]
cta = ctypes.c_int.in_dll(eng.lib, "udal_nms_cta")
dbg = ctypes.c_int.in_dll(eng.lib, "udal_nms_debug")


def time_it(b, reps=5):
    bx = eng.ctx.to_device(boxes.numpy()[:b])
    sc = eng.ctx.to_device(scores.numpy()[:b])
    eng.nms_v5(bx, sc)
    eng.ctx.sync()
    eng.ctx.timer_start()
    for _ in range(reps):
        eng.nms_v5(bx, sc)
    return eng.ctx.timer_stop() / reps


for mode in (1, 0):
    cta.value = mode
    print("udal_nms_cta=%d: B=%d %.3f ms, B=1 %.3f ms" % (mode, batch, time_it(batch), time_it(1)))
cta.value = 1
dbg.value = 1
out = (ctypes.c_ulonglong * 8)()
eng.lib.udal_nms_debug_read(out, 1)
idx, ss, valid = eng.nms_v5(boxes, scores)
eng.ctx.sync()
eng.lib.udal_nms_debug_read(out, 1)
dbg.value = 0
print("image 0 cycles: select %d, round1 %d (loop alone: thread 0 %d, slowest warp %d), round2 %d, commit %d; round-2 evaluations %d; "
      "candidates %d; valid %s" % (out[0], out[1], out[6], out[7], out[2], out[3], out[4], out[5], valid.numpy()[:4]))
