"""Developer tool (GPU box): one engine parse() over N master blocks, greedy only (stage A) or
optimal with K iterations on 50 KB blocks (stage C shape), with the per-kernel event timers --
no lane overlap, so the numbers add up.  usage: gpu_stage.py MB mode iters [blockbytes]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zopfli_b200 as zb
from zopfli_b200 import corpus
mb = int(sys.argv[1]); mode = int(sys.argv[2]); its = int(sys.argv[3])
bs = int(sys.argv[4]) if len(sys.argv) > 4 else 1000000
data = corpus.synth_text(mb * 1000000, 2)
lib = zb.Library()
ranges = [(a, min(a + bs, len(data))) for a in range(0, len(data), bs)]
for rep in range(2):
    lib.reset_stats()
    t = time.time()
    lib.lz77_batch(data, ranges, mode, its)
    dt = time.time() - t
    st = lib.stats()
    print("rep %d: %.3fs" % (rep, dt), {k: round(v, 2) for k, v in st.items() if k.startswith("ms_")})
