import sys, time, ctypes
import numpy as np
sys.path.insert(0, ".")
import udal_b200 as u
batch=64
p = u.hparams_config.get_detection_config("efficientdet-d0", image_size=(384, 1280), num_classes=8, enable_softmax=True, loss_attenuation=True,
    mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05, mc_dropoutsamp=10, heads_mode="fp16")
eng = u.engine.get_engine(p)
eng.set_head_weights(u.synthetic.init_head_weights(eng.F, eng.R, len(eng.level_hw), eng.A, 8, True, seed=2024))
rng = np.random.default_rng(1)
feats = [eng.ctx.to_device(rng.standard_normal((batch, h, w, eng.F), dtype=np.float32)) for h, w in eng.level_hw]
scales = eng.ctx.to_device(np.ones(batch, np.float32))
for i in range(3): eng.run(feats, scales, None, seed=i)
eng.ctx.sync()
tr = ctypes.c_int.in_dll(eng.lib, "udal_host_trace"); tr.value = 1
ts=[]
t0=time.perf_counter()
for i in range(5):
    a=time.perf_counter(); out=eng.run(feats, scales, None, seed=10+i); ts.append(time.perf_counter()-a)
t1=time.perf_counter(); eng.ctx.sync(); t2=time.perf_counter()
tr.value = 0
print("host ms per call:", [round(x*1e3,3) for x in ts])
print("enqueue total %.2f ms, after sync %.2f ms -> %.3f ms/step" % ((t1-t0)*1e3, (t2-t0)*1e3, (t2-t0)*1e3/20))
ovl = ctypes.c_int.in_dll(eng.lib, "udal_run_overlap")
for ov in (0,1):
    ovl.value=ov
    for i in range(3): eng.run(feats, scales, None, seed=i)
    eng.ctx.sync(); t0=time.perf_counter()
    for i in range(20): out=eng.run(feats, scales, None, seed=10+i)
    eng.ctx.sync(); print("overlap", ov, "%.3f ms/step" % ((time.perf_counter()-t0)*1e3/20))
