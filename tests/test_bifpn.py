"""BiFPN: the node graph against the reference's own fpn_configs (golden), the oracle's TF resampling
rules against plain NumPy loops, and - on the GPU - the device FPNCells against the oracle."""
import json
import os

import numpy as np
import pytest

from oracle import bifpn_ref, heads_ref

HERE = os.path.dirname(os.path.abspath(__file__))


def test_bifpn_graph_matches_the_reference_fpn_configs():
    import udal_b200 as u
    golden = json.load(open(os.path.join(HERE, "golden", "bifpn_graph.json")))
    assert len(golden) == 10
    for key, g in golden.items():
        lo, hi, wm = key.split("_")
        cfg = u.fpn_configs.bifpn_config(int(lo), int(hi), None if wm == "None" else wm)
        assert cfg["weight_method"] == g["weight_method"]
        assert cfg["nodes"] == g["nodes"]
        assert bifpn_ref.bifpn_nodes(int(lo), int(hi)) == g["nodes"]      # the oracle's own restatement
    with pytest.raises(ValueError):
        u.fpn_configs.get_fpn_config("qufpn", 3, 7, None)


def test_oracle_resampling_rules_against_numpy_loops():
    import torch
    rng = np.random.default_rng(0)
    for h, w, th, tw in [(48, 160, 24, 80), (23, 40, 12, 20), (45, 80, 23, 40), (3, 10, 2, 5), (6, 20, 3, 10), (64, 64, 32, 32)]:
        x = rng.normal(size=(2, h, w, 3)).astype(np.float32)
        for avg in (False, True):
            np.testing.assert_allclose(bifpn_ref.pool_same(torch.from_numpy(x), th, tw, avg).numpy(),
                                       bifpn_ref.pool_same_np(x, th, tw, avg), atol=1e-6)
        small = bifpn_ref.pool_same_np(x, th, tw)
        np.testing.assert_array_equal(bifpn_ref.nearest(torch.from_numpy(small), h, w).numpy(), bifpn_ref.nearest_np(small, h, w))
    # known answers: 3x3 / stride-2 SAME pooling of an even size pads only at the END (TF), of an odd size on both sides
    x = np.arange(16, dtype=np.float32).reshape(1, 4, 4, 1)
    np.testing.assert_array_equal(bifpn_ref.pool_same_np(x, 2, 2)[0, :, :, 0], [[10, 11], [14, 15]])
    x = np.arange(25, dtype=np.float32).reshape(1, 5, 5, 1)
    np.testing.assert_array_equal(bifpn_ref.pool_same_np(x, 3, 3)[0, :, :, 0], [[6, 8, 9], [16, 18, 19], [21, 23, 24]])
    x = np.arange(6, dtype=np.float32).reshape(1, 2, 3, 1)
    np.testing.assert_array_equal(bifpn_ref.nearest_np(x, 4, 5)[0, :, :, 0], [[0, 0, 1, 1, 2], [0, 0, 1, 1, 2], [3, 3, 4, 4, 5], [3, 3, 4, 4, 5]])


def _level_sizes(h, w, n=5):
    out = []
    for _ in range(3):          # strides 2, 4, 8 -> level 3
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    for _ in range(n):
        out.append((h, w))
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    return out


CASES = [
    # name, image (H, W), model, in_channels of the first cell, batch, overrides
    ("d0_512", (512, 512), "efficientdet-d0", [40, 112, 320, 64, 64], 2, {}),
    ("d0_kitti", (384, 1280), "efficientdet-d0", [40, 112, 320, 64, 64], 1, {}),               # ragged: 3 x 10 at level 7
    ("d0_bdd", (720, 1280), "efficientdet-d0", [40, 112, 320, 64, 64], 1, {}),                 # 90, 45, 23, 12, 6 rows: odd sizes
    ("d1_sum", (256, 320), "efficientdet-d1", [40, 112, 320, 88, 88], 2, {"fpn_weight_method": "sum"}),
    ("d0_attn", (128, 192), "efficientdet-d0", [64] * 5, 2, {"fpn_weight_method": "attn"}),    # no 1x1 convs at all
    ("d0_chan", (128, 192), "efficientdet-d0", [40, 112, 320, 64, 64], 1, {"fpn_weight_method": "channel_fastattn"}),
    ("d0_chattn_cbn", (128, 128), "efficientdet-d0", [40, 112, 320, 64, 64], 1,
     {"fpn_weight_method": "channel_attn", "conv_bn_act_pattern": True, "conv_after_downsample": True, "apply_bn_for_resampling": False}),
]


@pytest.mark.gpu
@pytest.mark.parametrize("name,size,model,cin,batch,over", CASES)
def test_fpn_cells_vs_oracle(name, size, model, cin, batch, over):
    import udal_b200 as u
    p = u.hparams_config.get_detection_config(model, image_size=size, num_classes=8, **over)
    f = p["fpn_num_filters"]
    method = p.get("fpn_weight_method") or "fastattn"
    nodes = u.fpn_configs.bifpn_config(3, 7, method)["nodes"]
    w = u.synthetic.init_bifpn_weights(f, p["fpn_cell_repeats"], cin, nodes, weight_method=method, seed=3)
    rng = np.random.default_rng(1)
    feats = [rng.normal(size=(batch, h, ww, c)).astype(np.float32) for (h, ww), c in zip(_level_sizes(*size), cin)]
    cells = u.bifpn.FPNCells(p, w)
    got = cells(feats)
    ref = bifpn_ref.fpn_cells(feats, w, weight_method=method, apply_bn_for_resampling=p["apply_bn_for_resampling"],
                              conv_after_downsample=p["conv_after_downsample"], conv_bn_act_pattern=p["conv_bn_act_pattern"])
    assert len(got) == 5
    for g, r in zip(got, ref):
        assert g.shape == r.shape
        # two fp32 conv implementations with different summation orders, 3-5 cells deep
        np.testing.assert_allclose(g, r, rtol=2e-4, atol=2e-4)
    # device arrays in -> device arrays out, same numbers
    dev = cells([cells.ctx.to_device(x) for x in feats])
    for g, d in zip(got, dev):
        np.testing.assert_array_equal(g, d.numpy())


@pytest.mark.gpu
def test_bifpn_feeds_the_head_sampler():
    """backbone-level features -> FPNCells -> HeadSampler.detect, everything on the device, against the two oracles chained"""
    import udal_b200 as u
    from oracle import ref_np
    size, cin, batch = (128, 192), [40, 112, 320, 64, 64], 2
    p = u.hparams_config.get_detection_config("efficientdet-d0", image_size=size, num_classes=7, enable_softmax=True,
                                              loss_attenuation=True, mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05,
                                              mc_dropoutsamp=3)
    nodes = u.fpn_configs.bifpn_config(3, 7)["nodes"]
    wf = u.synthetic.init_bifpn_weights(64, 3, cin, nodes, seed=5)
    wh = heads_ref.init_head_weights(64, 3, 5, 9, 7, True, randomize_bn=True)
    rng = np.random.default_rng(2)
    feats = [rng.normal(size=(batch, h, ww, c)).astype(np.float32) for (h, ww), c in zip(_level_sizes(*size), cin)]
    masks = heads_ref.make_masks(3, 5, 3, batch, 64, 0.05, 0.05)
    cells, sampler = u.bifpn.FPNCells(p, wf), u.heads.HeadSampler(p, wh)
    fpn = cells([cells.ctx.to_device(x) for x in feats])
    fpn_host = [x.numpy() for x in fpn]
    cls, box = sampler(fpn_host, masks=masks)
    ref_fpn = bifpn_ref.fpn_cells(feats, wf)
    rcls, rbox = heads_ref.heads_sample(ref_fpn, wh, masks, 0.05, 0.05, 3)
    for a, b in zip(cls + box, rcls + rbox):
        np.testing.assert_allclose(a, b, rtol=1e-3, atol=1e-3)
    det = sampler.detect(fpn_host, np.ones(batch, np.float32), masks=masks)
    assert det[0].shape == (batch, 100, 12) and int(det[3].min()) > 0
    with pytest.raises(ValueError, match="Incompatible Resampling"):
        bad = list(feats)
        bad[1] = rng.normal(size=(batch, feats[1].shape[1] + 9, 4, cin[1])).astype(np.float32)
        cells(bad)
