#!/bin/bash
# compute-sanitizer over a small end-to-end call (all kernels incl. k_emit / k_block_plan / k_scatter / k_greedy):
# memcheck for out-of-bounds / misaligned accesses, racecheck for shared-memory hazards.
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import zopfli_b200 as zb, zref
from zopfli_b200 import corpus
lib = zb.library()
ref = zref.Ref()
for name, d, kw in [("text", corpus.synth_text(70000, 3), dict(numiterations=2)),
                    ("mixed", corpus.synth_binary(30000) + corpus.random_bytes(20000) + corpus.adv_runs()[:30000], dict(numiterations=2)),
                    ("tiny", b"abcabcabcabc", dict(numiterations=1)), ("empty", b"", dict(numiterations=1))]:
    for fmt in (0, 2):
        got = lib.compress(d, fmt, **kw)
        assert got == ref.compress(d, fmt, **kw), (name, fmt)
print("sanitize workload ok")
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload ok|Error|hazard" gpurun_out/sanitize_$tool.log | head -12
done
