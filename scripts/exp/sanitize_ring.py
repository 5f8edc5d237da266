"""Small invocations of the persistent ring kernel (claimed and dealt rows) and of the row-block matmul entry points for
compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, ".")
import numpy as np
import trueno_b200 as trn
rng = np.random.default_rng(1)
f32 = np.float32
for dyn in ("1", "0"):
    os.environ["TRN_RING_DYN"] = dyn
    for rows, cols in [(1, 32000), (5, 28680), (160, 32768), (300, 31992)]:
        x = (rng.standard_normal((rows, cols)) * 4).astype(f32)
        for log in (False, True):
            y = trn.softmax_rows(x, rows, cols, log=log)
            assert np.all(np.isfinite(y))
print("sanitize ring workload ok")
