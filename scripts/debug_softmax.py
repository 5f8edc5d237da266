import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trueno_b200 as trn
f32 = np.float32
def ulp(x): return np.spacing(np.abs(x).astype(f32)).astype(np.float64)
for rows, cols in [(700, 9000), (450, 32768), (149, 8196), (3, 20000), (300, 32000)]:
    rng = np.random.default_rng(rows * 131 + cols)
    x = (rng.standard_normal((rows, cols)) * 4).astype(f32)
    arg = (x - x.max(1, keepdims=True)).astype(f32).astype(np.float64)
    e64 = np.exp(arg)
    tlog = arg - np.log(e64.sum(1, keepdims=True))
    truth = e64 / e64.sum(1, keepdims=True)
    for rep in range(3):
        glog = trn.softmax_rows(x, rows, cols, log=True)
        got = trn.softmax_rows(x, rows, cols)
        viol = np.abs(glog - tlog) - (4 * ulp(tlog) + 2.0 ** -22)
        bad = np.argwhere(viol > 0)
        v2 = np.abs(got - truth) - np.minimum(1e-6, 8 * ulp(truth) + 1e-45)
        bad2 = np.argwhere(v2 > 0)
        print(rows, cols, "rep", rep, "log bad", len(bad), "sm bad", len(bad2))
        if len(bad):
            rws = np.unique(bad[:, 0])
            print("  bad rows", rws[:20], "n", len(rws), "rows mod 148", (rws % 148)[:20])
            r, c = bad[0]
            print("  first", r, c, glog[r, c], tlog[r, c], "x", x[r, c], "rowmax", x[r].max(), "cols in row", np.unique(bad[bad[:,0]==r][:,1])[:10], len(bad[bad[:,0]==r]))
            print("  nan/inf?", np.isnan(glog).sum(), np.isinf(glog).sum())
