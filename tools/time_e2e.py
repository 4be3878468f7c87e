"""Development helper: end-to-end (host buffers in, detections out) throughput of PipelinedSampler vs depth."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import udal_b200 as u

batch = 64
p = u.hparams_config.get_detection_config(
    "efficientdet-d0", image_size=(384, 1280), num_classes=8, enable_softmax=True, loss_attenuation=True,
    mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05, mc_dropoutsamp=10, heads_mode="bf16")
eng = u.engine.get_engine(p)
w = u.synthetic.init_head_weights(eng.F, eng.R, len(eng.level_hw), eng.A, 8, True, seed=2024)
rng = np.random.default_rng(1)
pinned = [u.device.PinnedArray((batch, h, ww, eng.F)) for h, ww in eng.level_hw]
for pa in pinned:
    pa.array[...] = rng.standard_normal(pa.shape, dtype=np.float32)
host_feats = [pa.array for pa in pinned]
scales = np.ones(batch, np.float32)
for depth in (1, 2, 3, 4):
    pipe = u.heads.PipelinedSampler(p, w, heads_mode="bf16", depth=depth)
    for _ in pipe.map([host_feats] * (2 * depth), [scales] * (2 * depth), seed=1):
        pass
    n = 24
    t0 = time.perf_counter()
    for det in pipe.map([host_feats] * n, [scales] * n, seed=100):
        pass
    ms = (time.perf_counter() - t0) * 1e3 / n
    print("depth %d: %.3f ms/batch = %.0f images/s" % (depth, ms, batch / ms * 1e3), flush=True)
