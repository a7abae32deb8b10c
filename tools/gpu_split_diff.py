import os, sys, subprocess, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mb = sys.argv[1] if len(sys.argv) > 1 else "5"
if len(sys.argv) > 2 and sys.argv[2] == "child":
    sys.path.insert(0, ROOT)
    import zopfli_b200 as zb
    from zopfli_b200 import corpus
    d = corpus.synth_text(int(float(mb) * 1e6), 2)
    L = zb.library()
    for rep in range(2):
        z = L.compress(d, 2, numiterations=2, verbose=1)
        print("OUT", rep, len(z), hashlib.sha256(z).hexdigest()[:16], flush=True)
    sys.exit(0)
res = {}
for mode in ("1", "0"):
    env = dict(os.environ, ZOPFLI_B200_HOST_SPLIT=mode)
    p = subprocess.run([sys.executable, __file__, mb, "child"], env=env, capture_output=True, text=True)
    pts = [l for l in p.stderr.splitlines() if l.startswith("block split points")]
    outs = [l for l in p.stdout.splitlines() if l.startswith("OUT")]
    res[mode] = (pts, outs)
    print("mode host_split=%s" % mode, outs, "nlines", len(pts))
    if p.returncode: print(p.stderr[-2000:])
a, b = res["1"][0], res["0"][0]
n = len(a) // 2
print("host rep0 vs rep1 equal:", sorted(a[:n]) == sorted(a[n:]))
print("gpu  rep0 vs rep1 equal:", sorted(b[:n]) == sorted(b[n:]))
sa, sb = sorted(a[:n]), sorted(b[:n])
diff = [(x, y) for x, y in zip(sa, sb) if x != y]
print("host vs gpu differing master blocks:", len(diff), "of", n)
for x, y in diff[:3]:
    print(" host:", x[:300]); print(" gpu :", y[:300])
