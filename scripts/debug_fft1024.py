"""debug helper: error structure of the 1024^2 transform (not part of the test-suite)"""
import os, sys
import numpy as np, scipy.fft as sf
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msm_b200 as m
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rng = np.random.default_rng(n * 10 + 2)
a = rng.standard_normal((3, n, n)) + 1j * rng.standard_normal((3, n, n))
f = m.forward(a, 2)
ref = sf.fftn(a, axes=(1, 2), norm="ortho")
err = np.abs(f - ref)
print("rel", np.linalg.norm(err) / np.linalg.norm(ref), "max", err.max(), "count>1e-10", (err > 1e-10).sum())
bad = np.argwhere(err > 1e-10)
if len(bad):
    print("batches", np.unique(bad[:, 0]), "rows", np.unique(bad[:, 1])[:40], len(np.unique(bad[:, 1])), "cols", np.unique(bad[:, 2])[:40], len(np.unique(bad[:, 2])))
    # is the error confined to one axis?  undo the column transform and look again
    g = sf.ifft(f, axis=1, norm="ortho"); gr = sf.fft(a, axis=2, norm="ortho")
    e2 = np.abs(g - gr); b2 = np.argwhere(e2 > 1e-10)
    print("after undoing axis-1 (strided pass): bad", len(b2), "rows", np.unique(b2[:, 1])[:20], "cols", np.unique(b2[:, 2])[:20])
    g = sf.ifft(f, axis=2, norm="ortho"); gr = sf.fft(a, axis=1, norm="ortho")
    e2 = np.abs(g - gr); b2 = np.argwhere(e2 > 1e-10)
    print("after undoing axis-2 (contiguous pass): bad", len(b2), "rows", np.unique(b2[:, 1])[:20], "cols", np.unique(b2[:, 2])[:20])
if len(bad):
    i = np.unravel_index(np.argmax(err), err.shape)
    print("worst", i, f[i], ref[i], f[i] / ref[i])
    eb = err[1]
    print("row-wise max err (rows 888..1023 step 8):", [float("%.1e" % eb[r].max()) for r in range(888, 1024, 8)])
    print("col-wise max err, cols 0..63:", [float("%.0e" % eb[:, c].max()) for c in range(64)])
    # compare with the exact column transform of our own (correct?) first pass
    g1 = sf.fft(a[1], axis=1, norm="ortho")           # contiguous pass done exactly
    col = sf.fft(g1, axis=0, norm="ortho")
    d = f[1] - col
    print("delta rows>=896 norm", np.linalg.norm(d[896:]), "rows<896 norm", np.linalg.norm(d[:896]))
    # is the delta itself a transform of something sparse?  look at d along rows for one bad column
    c = int(np.argmax(eb.max(axis=0)))
    print("bad col", c, "delta magnitudes rows 896..:", np.abs(d[896:, c])[:8], "...", np.abs(d[1016:, c]))
