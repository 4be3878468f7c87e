"""Development helper: tensor-core head sampler (python tests/manual/stress_heads.py [bf16 | fp16]) vs the CPU oracle over
geometries that make the persistent kernels loop many times per CTA, three repetitions each - identical errors across the
repetitions = no race (the pytest cases keep the oracle cost small)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import udal_b200 as u
from oracle import heads_ref

MODE = sys.argv[1] if len(sys.argv) > 1 else "fp16"
CASES = [  # size, C, T, batch
    ((384, 1280), 8, 10, 1),
    ((384, 1280), 8, 7, 1),
    ((384, 1280), 8, 6, 2),
    ((384, 640), 8, 10, 2),
    ((192, 1280), 7, 5, 2),
    ((256, 256), 8, 6, 3),
    ((720, 1280), 8, 4, 1),
]
worst_all = 0.0
for size, C, T, batch in CASES:
    p = u.hparams_config.get_detection_config(
        "efficientdet-d0", image_size=size, num_classes=C, enable_softmax=True, loss_attenuation=True,
        mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05, mc_dropoutsamp=T, heads_mode=MODE)
    eng = u.engine.get_engine(p)
    L = len(eng.level_hw)
    w = heads_ref.init_head_weights(eng.F, eng.R, L, eng.A, C, True, seed=9, randomize_bn=True)
    feats = heads_ref.make_features(eng.level_hw, batch, eng.F, seed=11)
    masks = heads_ref.make_masks(T, L, eng.R, batch, eng.F, 0.05, 0.05, seed=5)
    t0 = time.time()
    rcls, rbox = heads_ref.heads_sample(feats, w, masks, 0.05, 0.05, T)
    t1 = time.time()
    sampler = u.heads.HeadSampler(p, w)
    for rep in range(3):  # repeated: a race would not fail every time
        cls, box = sampler(feats, masks=masks)
        err = max(max(float(np.abs(a - b).max()) for a, b in zip(cls, rcls)),
                  max(float(np.abs(a - b).max()) for a, b in zip(box, rbox)))
        worst_all = max(worst_all, err)
        print("size=%s C=%d T=%d B=%d rep %d: max abs err %.5f (oracle %.1fs)" % (size, C, T, batch, rep, err, t1 - t0), flush=True)
print("WORST", worst_all)
sys.exit(0 if worst_all < 0.06 else 1)
