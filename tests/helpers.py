"""Shared helpers for the test-suite: synthetic inputs of SURVEY 8(d) and golden loading."""
import ast
import os

import numpy as np

from oracle import ref_np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)


def golden_params(g):
    p = ast.literal_eval(str(g["params_repr"]).replace("-inf", "-1e999"))
    return p


def golden_inputs(g):
    return [g["cls%d" % i] for i in range(5)], [g["box%d" % i] for i in range(5)]


def synth_head_outputs(params, batch, seed, mc_cls=True, mc_box=True, la=True):
    """config-4 distributions: logits ~ N(-4.6, 2), t_yx ~ N(0, .5), t_hw ~ N(0, .25),
    sigma = clip(|N(0, .3)|, .01, 2)."""
    rng = np.random.default_rng(seed)
    a = ref_np.num_anchors_per_location(params)
    c = params["num_classes"]
    T = params["mc_dropoutsamp"]
    cls, box = [], []
    for h, w in ref_np.level_shapes(params):
        lead_c = (T, batch) if mc_cls else (batch,)
        lead_b = (T, batch) if mc_box else (batch,)
        cls.append(rng.normal(-4.6, 2.0, lead_c + (h, w, a * c)).astype(np.float32))
        t = rng.normal(0, 1, lead_b + (h, w, a, 4))
        t[..., :2] *= 0.5
        t[..., 2:] *= 0.25
        parts = [t.reshape(lead_b + (h, w, a * 4))]
        if la:
            parts.append(np.clip(np.abs(rng.normal(0, 0.3, lead_b + (h, w, a * 4))), 0.01, 2.0))
        box.append(np.concatenate(parts, -1).astype(np.float32))
    return cls, box


def random_boxes(rng, n, extent=300.0, lo=8.0, hi=80.0):
    ctr = rng.uniform(0, extent, (n, 2))
    wh = rng.uniform(lo, hi, (n, 2))
    return np.concatenate([ctr - wh / 2, ctr + wh / 2], 1).astype(np.float32)
