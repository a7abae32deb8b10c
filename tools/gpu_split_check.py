import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import zopfli_b200 as zb, zref
from zopfli_b200 import corpus
L = zb.library(); R = zref.Ref()
d = corpus.synth_text(1_000_000, 2)
ll, dd = R.lz77(d, 0, 1000000, 3)
rng = np.random.default_rng(0)
n = len(ll)
for nreq in (50, 5000, 70000):
    a = rng.integers(0, n - 1, nreq); w = rng.choice([3, 50, 600, 5000, 100000], nreq)
    b = np.minimum(n, a + 1 + rng.integers(0, w))
    t = time.time(); dev = L.device_auto_type_bits(ll, dd, a, b); td = time.time() - t
    k = min(nreq, 3000)
    host = np.array([L.host_block_size(ll, dd, int(x), int(y), -1) for x, y in zip(a[:k], b[:k])]).astype(np.uint64)
    bad = np.nonzero(dev[:k] != host)[0]
    dev2 = L.device_auto_type_bits(ll, dd, a, b)
    print(nreq, "dev %.3fs" % td, "mismatch", len(bad), "nondeterministic", int((dev != dev2).sum()))
    for i in bad[:5]: print("   ", a[i], b[i], dev[i], host[i])
