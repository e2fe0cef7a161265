import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from pydrodem_b200 import stencils, _lib
n = 20000
z = torch.randn((n, n), device="cuda")
for w in (9, 7, 5, 3):
    k = stencils.build_kernel(stencils.DW_WEIGHTS[w], dtype=np.float32, normalize=True)
    for _ in range(2): stencils.apply_convolve_filter(z, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): stencils.apply_convolve_filter(z, k)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"conv {w}x{w}: {ms:.3f} ms  {n*n*w*w/ms/1e9:.1f} TFMA/s  {n*n*8/ms/1e6:.0f} GB/s algorithmic")
for name, fn in (("slope_D4", lambda: stencils.slope_D4(z, 30.0, 30.0)), ("curv", lambda: stencils.curvature_D2(z, 30.0, 30.0))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, e0.elapsed_time(e1) / 3, "ms")
