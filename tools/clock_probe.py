"""Development helper: SM clock / power under a sustained udal_run loop (NVML samples every 50 ms for ~4 s)."""
import sys
import threading
import time

import numpy as np
import pynvml

sys.path.insert(0, ".")
import udal_b200 as u

batch = 64
p = u.hparams_config.get_detection_config(
    "efficientdet-d0", image_size=(384, 1280), num_classes=8, enable_softmax=True, loss_attenuation=True,
    mc_dropout=True, mc_classheadrate=0.05, mc_boxheadrate=0.05, mc_dropoutsamp=10, heads_mode="bf16")
eng = u.engine.get_engine(p)
L = len(eng.level_hw)
eng.set_head_weights(u.synthetic.init_head_weights(eng.F, eng.R, L, eng.A, 8, True, seed=2024))
rng = np.random.default_rng(1)
feats = [eng.ctx.to_device(rng.standard_normal((batch, h, w, eng.F), dtype=np.float32)) for h, w in eng.level_hw]
scales = eng.ctx.to_device(np.ones(batch, np.float32))
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples = []
stop = False


def sampler():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        time.sleep(0.05)


th = threading.Thread(target=sampler)
th.start()
time.sleep(0.3)
n_idle = len(samples)
for rep in range(4):
    eng.ctx.timer_start()
    for i in range(250):
        eng.run(feats, scales, None, seed=i)
    ms = eng.ctx.timer_stop() / 250
    print("rep %d: %.3f ms / step" % (rep, ms), flush=True)
stop = True
th.join()
print("idle  :", samples[:n_idle])
print("loaded:", samples[n_idle:])
