import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rfb200
from oracle import rf_oracle as O
ctx = rfb200.Context(0)
for n in (512, 1024, 2048, 4096, 8192):
    a0 = np.asfortranarray(np.random.default_rng([12, n]).random((n, n), dtype=np.float32))
    for mode in (0, 1):
        F = rfb200.lu(a0, ctx=ctx, f32_mode=mode)
        r = O.residual_inf(a0, F.factors, F.ipiv) if n <= 4096 else float('nan')
        rf = O.residual_fro_rel(a0, F.factors, F.ipiv) if n <= 4096 else float('nan')
        print(n, "mode", mode, "res_inf %.4e" % r, "bound %.4e" % (20*n*1.19e-7), "fro_rel %.3e" % rf, flush=True)
