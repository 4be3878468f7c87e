import sys, time
import numpy as np
sys.path.insert(0, ".")
import udal_b200 as u
combos = [((64, 64), 8, 2, 1, "efficientdet-d0"), ((64, 64), 7, 1, 5, "efficientdet-d0"), ((384, 1280), 8, 10, 200, "efficientdet-d0"),
          ((384, 1280), 10, 3, 7, "efficientdet-d0"), ((40, 200), 8, 4, 33, "efficientdet-d0"), ((512, 512), 8, 30, 9, "efficientdet-d0"),
          ((64, 64), 10, 2, 1, "efficientdet-d2"), ((768, 768), 10, 5, 3, "efficientdet-d2"), ((96, 320), 90, 2, 2, "efficientdet-d0"),
          ((96, 320), 90, 2, 2, "efficientdet-d1")]
for size, C, T, B, model in combos:
    p = u.hparams_config.get_detection_config(model, image_size=size, num_classes=C, enable_softmax=True, loss_attenuation=True,
        mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05, mc_dropoutsamp=T, heads_mode="bf16")
    eng = u.engine.get_engine(p)
    eng.set_head_weights(u.synthetic.init_head_weights(eng.F, eng.R, len(eng.level_hw), eng.A, C, True, seed=1, randomize_bn=True))
    rng = np.random.default_rng(1)
    feats = [eng.ctx.to_device(rng.standard_normal((B, h, w, eng.F), dtype=np.float32)) for h, w in eng.level_hw]
    t0 = time.perf_counter()
    outs = [eng.run(feats, None, None, seed=i) for i in range(6)]   # back to back: tails overlap the next heads
    v = [o["valid"].numpy() for o in outs]
    a = eng.run(feats, None, None, seed=3)["boxes"].numpy()
    assert np.array_equal(a, outs[3]["boxes"].numpy()), "not deterministic"
    print(model, size, "C", C, "T", T, "B", B, "ok valid[0]=%d  %.1f ms" % (v[0][0], (time.perf_counter() - t0) * 1e3), flush=True)
    del feats, outs
    u.engine.clear_engines()
print("ALL OK")
