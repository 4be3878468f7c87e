"""BiFPN (bifpn.FPNCells) on the device: time of one call at a BASELINE geometry, features resident in HBM.
    python tools/time_bifpn.py [H W batch model]"""
import sys

import numpy as np

sys.path.insert(0, ".")
import udal_b200 as u

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (384, 1280)
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 64
model = sys.argv[4] if len(sys.argv) > 4 else "efficientdet-d0"
p = u.hparams_config.get_detection_config(model, image_size=(H, W), num_classes=8)
f = p["fpn_num_filters"]
cin = [40, 112, 320, f, f]      # EfficientNet-B0 levels 3..5 + the two resampled levels
nodes = u.fpn_configs.bifpn_config(3, 7, "fastattn")["nodes"]
w = u.synthetic.init_bifpn_weights(f, p["fpn_cell_repeats"], cin, nodes, weight_method="fastattn", seed=3)
cells = u.bifpn.FPNCells(p, w)
sizes = []
h, ww = H, W
for _ in range(3):
    h, ww = (h - 1) // 2 + 1, (ww - 1) // 2 + 1
for _ in range(5):
    sizes.append((h, ww))
    h, ww = (h - 1) // 2 + 1, (ww - 1) // 2 + 1
rng = np.random.default_rng(1)
feats = [cells.ctx.to_device(rng.standard_normal((batch, a, b, c), dtype=np.float32)) for (a, b), c in zip(sizes, cin)]
for _ in range(2):
    out = cells(feats)
cells.ctx.sync()
l0 = cells.ctx.launch_count()
cells.ctx.timer_start()
for _ in range(3):
    out = cells(feats)
ms = cells.ctx.timer_stop() / 3
print("%s %dx%d batch %d: FPNCells %.2f ms per call (%d launches), %.0f images/s; outputs %s"
      % (model, W, H, batch, ms, (cells.ctx.launch_count() - l0) // 3, batch / ms * 1e3, [tuple(o.shape) for o in out]))
