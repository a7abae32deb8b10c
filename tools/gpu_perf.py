"""Developer tool (GPU box): compress N MB of the bench text and print the engine's timers and the
k_iterate phase-cycle breakdown (sum over blocks and for the critical-path block)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zopfli_b200 as zb
from zopfli_b200 import corpus
mb = float(sys.argv[1]) if len(sys.argv) > 1 else 20
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 15
kind = sys.argv[3] if len(sys.argv) > 3 else "text"
n = int(mb * 1e6)
data = corpus.synth_text(n, 2) if kind == "text" else corpus.synth_binary(n, 4)
L = zb.library()
L.compress(data[:200000], 0, numiterations=2)  # warm up (context, log table)
for rep in range(2):
    L.reset_stats()
    t = time.time(); z = L.compress(data, 0, numiterations=iters); dt = time.time() - t
    st = L.stats()
    import zlib
    print("rep %d: %.1f MB in %.3fs = %.2f MiB/s, out %d (crc %08x)" % (rep, mb, dt, n / 1048576 / dt, len(z), zlib.crc32(z)))
names = ["model", "dp", "trace", "follow", "cost", "stats"]
print({k: round(v, 2) for k, v in st.items() if k.startswith("ms_")})
tot = sum(st["cyc_sum"]) or 1
print("phase share (sum over blocks):", {nm: "%.1f%%" % (100.0 * c / tot) for nm, c in zip(names, st["cyc_sum"])})
mx = sum(st["cyc_max"]) or 1
print("critical block: %d positions, %.1f Mcycles total = %.3fs @1.965GHz:" % (st["max_block_positions"], mx / 1e6, mx / 1.965e9),
      {nm: "%.1f%%" % (100.0 * c / mx) for nm, c in zip(names, st["cyc_max"])})
print("cycles per DP step (critical block): %.1f ; all blocks avg: %.1f" % (
    st["cyc_max"][1] / max(1, st["max_block_positions"] * iters), st["cyc_sum"][1] / max(1, st["iterate_steps"])))
print("launches", st["launches"], "steps", st["iterate_steps"], "integer-window share of DP steps: %.1f%%" % (100.0 * st["int_steps"] / max(1, st["iterate_steps"])))
kinds = ["int", "magic", "plain", "ring", "general"]
for tag in ("sum", "max"):
    cy, cn = st["dp_cyc_" + tag], st["dp_cnt_" + tag]
    tot = sum(cy) or 1
    print("DP by group kind (%s):" % ("all blocks" if tag == "sum" else "critical block"),
          {k: "%.1f%% of cycles, %d groups, %.1f cyc/step" % (100.0 * c / tot, n, c / max(1, n) / 32) for k, c, n in zip(kinds, cy, cn)},
          "positions in the per-step loop:", cn[5])
