"""dev timing of the water / veg kernels at C2 size (10801^2), device resident"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from pydrodem_b200.tools import water
from pydrodem_b200 import veg, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10801
g = torch.Generator(device="cuda").manual_seed(1)
dem = torch.rand((n, n), device="cuda", generator=g) * 50
burn = dem - torch.rand((n, n), device="cuda", generator=g) * 3
wm = (torch.rand((n, n), device="cuda", generator=g) * 40).to(torch.uint8)
wm[wm > 15] = 0
swo = (torch.rand((n, n), device="cuda", generator=g) * 100).to(torch.int16)
tc = (torch.rand((n, n), device="cuda", generator=g) * 100).to(torch.uint8)
th = (torch.rand((n, n), device="cuda", generator=g) * 40).to(torch.uint8)
cube = np.random.default_rng(0).random((101, 41, 20))
def t(name, f, reps=3):
    f(); f(); f(); torch.cuda.synchronize()   # three warm-ups: the arena settles after its second call
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:28s} {ms:8.2f} ms  {n*n/ms/1e3:9.0f} Mcells/s", flush=True)
t("minimum_filter f32 k=51", lambda: water.minimum_filter(burn, 51))
t("maximum_filter u8 k=35", lambda: water.maximum_filter(wm, 35))
t("river_minimize(2,5)", lambda: water.river_minimize(burn, wm, swo, 2, 5))
t("raise_small_rivers fs=31", lambda: water.raise_small_rivers(dem, burn.clone(), wm, swo, -9999, filter_size=31))
t("raise_small_streams fs=51", lambda: water.raise_small_streams(dem, burn.clone(), wm, -9999))
t("smooth_rivers fs=21", lambda: water.smooth_rivers(dem, burn.clone(), wm, swo, -9999))
t("compute_veg_bias_3D", lambda: veg.compute_veg_bias_3D(tc, th, dem, cube, 30.0, 30.0))
_lib.profile_enable(True)
water.raise_small_streams(dem, burn.clone(), wm, -9999); veg.compute_veg_bias_3D(tc, th, dem, cube, 30.0, 30.0)
for k, v in sorted(_lib.profile_read().items(), key=lambda kv: -kv[1][1]): print(f"  {k:20s} n={v[0]:3d} {v[1]:8.3f} ms")
