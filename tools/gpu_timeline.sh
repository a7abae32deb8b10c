#!/bin/bash
# timeline of one bench-sized call under scheduling variants (debug marks of driver.cpp)
# usage: tools/gpu_timeline.sh  (on the GPU box; writes gpurun_out/timeline_*.txt)
mkdir -p gpurun_out
cat > /tmp/tl.py <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import zopfli_b200 as zb
from zopfli_b200 import corpus
data = corpus.synth_text(100000000, 2)
lib = zb.Library(os.environ['ZB_LIB']) if os.environ.get('ZB_LIB') else zb.library()
host = np.frombuffer(data, dtype=np.uint8).copy()
for i in range(3):
    t = time.perf_counter()
    ob = lib.compress_ptr_nocopy(host.ctypes.data, len(data), 0, numiterations=15)
    print("call %d: %.1f ms, %d bytes" % (i, (time.perf_counter() - t) * 1e3, len(ob)), file=sys.stderr)
    ob.close()
st = lib.stats()
print({k: round(v, 1) for k, v in st.items() if k.startswith("ms_")}, file=sys.stderr)
PY
run() { name=$1; shift; env ZOPFLI_B200_DEBUG=1 "$@" python /tmp/tl.py 2> gpurun_out/timeline_$name.txt; grep "call\|ms_" gpurun_out/timeline_$name.txt | sed "s/^/$name: /"; }
run base ZB_X=0
for v in $TL_VARIANTS; do run "$(echo $v | tr -c 'A-Za-z0-9\n' '_')" $v; done
[ -n "$TL_DEFERRED" ] && run deferred ZOPFLI_B200_SYNC_TOC=0
exit 0
