"""Development helper: udal_run time of an arbitrary configuration.

    python tools/time_config.py H W C T B [model=efficientdet-d0] [heads_mode=bf16]
"""
import sys

import numpy as np

sys.path.insert(0, ".")
import udal_b200 as u

H, W, C, T, batch = [int(x) for x in sys.argv[1:6]]
model = sys.argv[6] if len(sys.argv) > 6 else "efficientdet-d0"
mode = sys.argv[7] if len(sys.argv) > 7 else "bf16"
p = u.hparams_config.get_detection_config(
    model, image_size=(H, W), num_classes=C, enable_softmax=True, loss_attenuation=True,
    mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05, mc_dropoutsamp=T, heads_mode=mode)
eng = u.engine.get_engine(p)
eng.set_head_weights(u.synthetic.init_head_weights(eng.F, eng.R, len(eng.level_hw), eng.A, C, True, seed=2024))
rng = np.random.default_rng(1)
feats = [eng.ctx.to_device(rng.standard_normal((batch, h, w, eng.F), dtype=np.float32)) for h, w in eng.level_hw]
scales = eng.ctx.to_device(np.ones(batch, np.float32))
for i in range(3):
    out = eng.run(feats, scales, None, seed=i)
eng.ctx.sync()
t = eng.ctx.layer_times(lambda: eng.run(feats, scales, None, seed=9))
eng.ctx.sync()
eng.ctx.timer_start()
n = 5
for i in range(n):
    out = eng.run(feats, scales, None, seed=20 + i)
ms = eng.ctx.timer_stop() / n
print("%s %s " % (model, mode), end="")
print("%dx%d C=%d T=%d B=%d anchors=%d: %.3f ms/step = %.0f images/s; head layers ms %s; scratch %.1f GB; valid %s"
      % (W, H, C, T, batch, eng.N, ms, batch / ms * 1e3, [round(x, 3) for x in t], eng.ctx.scratch_bytes() / 1e9,
         out["valid"].numpy()[:4]))
