"""Developer diagnostic (GPU box): seam-by-seam comparison with first-mismatch details."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import zopfli_b200 as zb, zref
from zopfli_b200 import corpus

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
R = zref.Ref(); L = zb.library()
print(L.version(), flush=True)
TXT = corpus.synth_text(2300000, 2)
cases = [("text-head", TXT, 0, 20000), ("text-mid", TXT, 100000, 125000), ("collide", corpus.adv_collide(), 40000, 52000),
         ("runs", corpus.adv_runs(), 0, 40000), ("longrun", corpus.adv_longrun(), 0, 150000), ("tiny3", b"abcabcabc", 0, 9),
         ("binary", corpus.synth_binary(200000), 60000, 100000)]
if quick: cases = [("text-head", TXT, 0, 4000), ("runs", corpus.adv_runs(), 0, 6000)]
for name, data, s, e in cases:
    t = time.time()
    a = R.match_table(data, s, e); b = L.match_table(data, s, e)
    for x, y, what in zip(a, b, ("length", "dist", "sublen", "same", "hv", "hv2")):
        if what == "length": x = np.where(x < 3, 0, x); y = np.where(y < 3, 0, y)
        if what == "dist": x = np.where(a[0] < 3, 0, x); y = np.where(a[0] < 3, 0, y)
        if not np.array_equal(x, y):
            bad = np.argwhere(x != y)
            print("  MISMATCH", name, what, "count", len(bad), "first", bad[0], "ref", x[tuple(bad[0])], "got", y[tuple(bad[0])])
    print("match", name, "done %.2fs" % (time.time() - t), flush=True)
    for mode, it in [(2, 0), (1, 0), (0, 1), (0, 15)]:
        a = R.lz77(data, s, e, mode, it); b = L.lz77(data, s, e, mode, it)
        ok = np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        msg = ""
        if not ok:
            n = min(len(a[0]), len(b[0])); d = np.nonzero((a[0][:n] != b[0][:n]) | (a[1][:n] != b[1][:n]))[0]
            msg = "sizes %d %d first diff %s ref %s got %s" % (len(a[0]), len(b[0]), d[:1], [(a[0][i], a[1][i]) for i in d[:1]], [(b[0][i], b[1][i]) for i in d[:1]])
        print("  lz77", name, mode, it, "OK" if ok else "MISMATCH " + msg, flush=True)
if not quick:
    for name, data in [("empty", b""), ("a", b"a"), ("text40k", TXT[:40000]), ("mixed", corpus.mixed_small(50000)), ("text1.2M", TXT[:1200000])]:
        for fmt in (0, 2):
            kw = {"numiterations": 5}
            t = time.time(); g = L.compress(data, fmt, **kw); tg = time.time() - t
            t = time.time(); r = R.compress(data, fmt, **kw); tr = time.time() - t
            print("compress", name, fmt, "OK" if g == r else "MISMATCH %d vs %d" % (len(g), len(r)), "gpu %.3fs ref %.3fs" % (tg, tr), flush=True)
    print(L.stats())
