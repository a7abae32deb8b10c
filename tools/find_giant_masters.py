"""Which master blocks of the C2 bench text hold the largest deflate blocks?  Runs the REFERENCE's
first-pass splitter (ZopfliBlockSplit, blocksplitter.c:279-330, through oracle/_ref) on every master block
of synth_text(100e6, seed 2) in a process pool and prints the block lengths.  The result is frozen in
tests/test_gpu_parity_at_size.py (GIANT_MASTERS); rerun this if the corpus generator changes.
  python tools/find_giant_masters.py [nbytes] [seed]
"""
import os
import sys
from multiprocessing import Pool

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

MB = 1000000
_data = None


def _one(m):
    import zref
    ref = zref.Ref(ndebug=True)
    a, b = m * MB, min(len(_data), (m + 1) * MB)
    lo = max(0, a - 32768)
    pts = ref.block_split(_data[lo:b], a - lo, b - lo, 15)
    cuts = [a] + [int(p) + lo for p in pts] + [b]
    return m, [cuts[i + 1] - cuts[i] for i in range(len(cuts) - 1)]


def main():
    global _data
    from zopfli_b200 import corpus
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    _data = corpus.synth_text(n, seed)
    with Pool(min(8, os.cpu_count() or 1)) as p:
        res = p.map(_one, range((n + MB - 1) // MB))
    big = sorted(((max(b), m) for m, b in res), reverse=True)
    for sz, m in big[:12]:
        print("master %3d  largest block %7d" % (m, sz))
    print("GIANT_MASTERS =", [m for _, m in big[:5]])


if __name__ == "__main__":
    main()
