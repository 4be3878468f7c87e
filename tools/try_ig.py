"""Debug helper: run the bf16 head sampler with the implicit-GEMM kernel under both descriptor
base-offset conventions and report the error against the CPU oracle."""
import ctypes
import sys

import numpy as np

sys.path.insert(0, ".")
import udal_b200 as u
from oracle import heads_ref

lib = u._lib.load()
use_ig = ctypes.c_int.in_dll(lib, "udal_heads_tc_use_ig")
size, C, T, batch = (64, 96), 8, 3, 2
p = u.hparams_config.get_detection_config(
    "efficientdet-d0", image_size=size, num_classes=C, enable_softmax=True, loss_attenuation=True,
    mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05, mc_dropoutsamp=T, heads_mode="bf16")
eng = u.engine.get_engine(p)
L = len(eng.level_hw)
w = heads_ref.init_head_weights(eng.F, eng.R, L, eng.A, C, True, seed=9, randomize_bn=True)
feats = heads_ref.make_features(eng.level_hw, batch, eng.F, seed=11)
masks = heads_ref.make_masks(T, L, eng.R, batch, eng.F, 0.05, 0.05, seed=5)
sampler = u.heads.HeadSampler(p, w)
rcls, rbox = heads_ref.heads_sample(feats, w, masks, 0.05, 0.05, T)
for ig, m in ((0, 0), (1, 0)):
    use_ig.value = ig
    cls, box = sampler(feats, masks=masks)
    err = max(max(float(np.abs(a - b).max()) for a, b in zip(cls, rcls)),
              max(float(np.abs(a - b).max()) for a, b in zip(box, rbox)))
    print("use_ig=%d base_offset_mode=%d  max abs err %.5f" % (ig, m, err), flush=True)
