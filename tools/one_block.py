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
import zlib
print("one block: %d -> %d bytes (crc %08x), %.3fs" % (n, len(out), zlib.crc32(out), time.time() - t))
st = lib.stats()
print({k: v for k, v in st.items() if k in ("ms_iterate", "cyc_max", "iterate_steps")})
kinds = ["int", "magic", "plain", "ring", "general"]
print("DP %.1f cycles/step;" % (st["cyc_max"][1] / max(1, st["iterate_steps"])),
      {k: "%.1f%% of cycles, %.1f cyc/step" % (100.0 * c / max(1, sum(st["dp_cyc_max"])), c / max(1, m) / 32)
       for k, c, m in zip(kinds, st["dp_cyc_max"], st["dp_cnt_max"])}, "per-step loop positions:", st["dp_cnt_max"][5])
