"""One giant deflate block through the product library (block splitting off): the critical-path
case of the iterate kernel, for ncu captures.  usage: one_block.py [bytes] [iterations]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zopfli_b200 as zb
from zopfli_b200 import corpus

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
its = int(sys.argv[2]) if len(sys.argv) > 2 else 3
data = corpus.synth_text(n, 5)
lib = zb.Library(os.environ["ZB_LIB"]) if os.environ.get("ZB_LIB") else zb.Library()
t = time.time()
out = lib.compress(data, zb.ZOPFLI_FORMAT_DEFLATE, numiterations=its, blocksplitting=0)
print("one block: %d -> %d bytes, %.3fs" % (n, len(out), time.time() - t))
st = lib.stats()
print({k: v for k, v in st.items() if k in ("ms_iterate", "cyc_max", "iterate_steps")})
