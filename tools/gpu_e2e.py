import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, zopfli_b200 as zb
from zopfli_b200 import corpus
data = corpus.synth_text(100_000_000, 2)
n = len(data)
host = torch.zeros(n + 64, dtype=torch.uint8).pin_memory()
host[:n] = torch.frombuffer(bytearray(data), dtype=torch.uint8)
dev = host.cuda()
lib = zb.library()
for i in range(3):
    t = time.perf_counter()
    out = lib.compress_ptr(host.data_ptr(), n, zb.ZOPFLI_FORMAT_GZIP, dev_ptr=dev.data_ptr(), numiterations=15)
    print("compress_ptr resident: %.1f ms" % ((time.perf_counter() - t) * 1e3), len(out), file=sys.stderr)
