"""gaussjacobi(n, a, b) on the device against the chunked oracle: where the largest node / weight deviations sit.
FGQ_LIB_PATH selects the build (e.g. one compiled with -DJAC_ROTATE=0)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fgq_b200  # noqa: E402
from oracle import jacobi_oracle as jo  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 7
a, b = 0.3, -0.7
cache = f"/tmp/jac_oracle_{n}.npz"
if os.path.exists(cache):
    z = np.load(cache)
    xo, wo = z["x"], z["w"]
else:
    xo, wo = jo.jacobi_asy_chunked(n, a, b, chunk=250_000)
    np.savez(cache, x=xo, w=wo)
x, w, ws = fgq_b200.gaussjacobi_device(n, a, b)
torch.cuda.synchronize()
x, w = x.cpu().numpy(), w.cpu().numpy()
eps = np.finfo(float).eps
ex = np.abs(x - xo) / eps
rw = np.abs(w - wo) / wo
print(os.environ.get("FGQ_LIB_PATH", "default"), f"n={n}: node err max {ex.max():.2f} ulp(1) at {ex.argmax()}, weight rel err max {rw.max():.3e} at {rw.argmax()} "
      f"(x = {xo[rw.argmax()]:.6f}); weights above 5e-14: {np.count_nonzero(rw > 5e-14)}, above 1e-13: {np.count_nonzero(rw > 1e-13)}")
top = np.argsort(rw)[-5:]
print("  worst weights at indices", top, "rel", rw[top])
h, _ = np.histogram(np.log10(np.maximum(rw, 1e-18)), bins=[-18, -16, -15, -14.5, -14, -13.5, -13, -12.5, -12])
print("  log10 rel-err histogram [-18,-16,-15,-14.5,-14,-13.5,-13,-12.5,-12]:", h)
