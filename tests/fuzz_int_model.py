"""Test infrastructure (not collected by pytest): random inputs (small alphabets, noisy periods, text-like, binary-like, byte runs, repeated
phrases) x random ranges x {fixed tree, 1..16 iterations} through oracle/dp_int_model.c, which checks the integer
formulation of the forward DP against the reference arithmetic on every pass.  usage: tests/fuzz_int_model.py [seed] [seconds]"""
import sys, ctypes as C, numpy as np, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/ -> repo root
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from zopfli_b200 import corpus
lib = C.CDLL(os.path.join(ROOT, 'oracle', '_build', 'libdp_int_model.so'))
lib.zo_dp_int_check.restype = C.c_uint64
lib.zo_dp_int_check.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 1)
def gen(kind, n):
    if kind == 0:   # small alphabet
        k = int(rng.integers(2, 40)); return rng.integers(0, k, n, dtype=np.uint8).tobytes()
    if kind == 1:   # periodic with noise
        p = int(rng.integers(1, 700)); base = rng.integers(0, 256, p, dtype=np.uint8)
        a = np.tile(base, n // p + 1)[:n].copy(); m = rng.random(n) < rng.choice([0.001, 0.01, 0.05]); a[m] = rng.integers(0, 256, int(m.sum()), dtype=np.uint8); return a.tobytes()
    if kind == 2:   # text-like
        return corpus.synth_text(n, int(rng.integers(1, 1000)))
    if kind == 3:   # binary-like
        return corpus.synth_binary(n, int(rng.integers(1, 1000)))
    if kind == 4:   # runs of random lengths
        out = bytearray()
        while len(out) < n:
            if rng.random() < 0.5: out += bytes([int(rng.integers(0, 256))]) * int(rng.integers(1, 2000))
            else: out += rng.integers(0, 256, int(rng.integers(1, 400)), dtype=np.uint8).tobytes()
        return bytes(out[:n])
    if kind == 5:   # repeated phrases (long matches) with edits
        words = [rng.integers(97, 123, int(rng.integers(3, 80)), dtype=np.uint8).tobytes() for _ in range(60)]
        out = bytearray()
        while len(out) < n: out += words[int(rng.integers(0, len(words)))]
        return bytes(out[:n])
tot = 0; t0 = time.time(); runs = 0
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 600
while time.time() - t0 < budget:
    kind = int(rng.integers(0, 6)); n = int(rng.integers(20000, 400000)); its = int(rng.choice([0, 1, 3, 8, 16]))
    data = gen(kind, n); s = int(rng.integers(0, min(32768, n // 4))); e = n - int(rng.integers(0, 100))
    buf = np.frombuffer(data, dtype=np.uint8).copy(); out = np.zeros(8, dtype=np.uint64)
    mm = lib.zo_dp_int_check(buf.ctypes.data, s, e, its, out.ctypes.data)
    runs += 1; tot += mm
    if mm: print("MISMATCH kind", kind, "n", n, "its", its, "s", s, "e", e, "mm", mm, flush=True)
print("runs", runs, "total mismatches", tot, "ties seen", "-", flush=True)
