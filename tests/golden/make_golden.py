"""Generates tests/golden/zopfli_golden.json with the UNMODIFIED reference (oracle/_ref, built from
/root/reference).  The reference ships no byte-level golden vectors; these are its outputs on small
deterministic inputs, recorded so that the GPU box (which has no /root/reference) can compare
against fixed bytes as well as against the prebuilt oracle/_ref.
Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import zref  # noqa: E402
from zopfli_b200 import corpus  # noqa: E402

CASES = {
    "empty": lambda: b"",
    "a": lambda: b"a",
    "abc_x3": lambda: b"abcabcabc",
    "go_foobar": corpus.go_case_foobar,
    "rand3000": lambda: corpus.random_bytes(3000, 1),
    "text32k": lambda: corpus.synth_text(32768, 1),          # config C1 stand-in
    "mixed50k": lambda: corpus.mixed_small(50000),
    "runs60k": lambda: corpus.adv_runs()[:60000],
    "collide30k": lambda: corpus.adv_collide()[:30000],
    "zeros70k": lambda: b"\0" * 70000,
}


def main():
    ref = zref.Ref()
    out = {}
    for name, gen in CASES.items():
        data = gen()
        entry = {"n": len(data), "sha256": hashlib.sha256(data).hexdigest(), "streams": {}}
        for fmt, fname in ((0, "gzip"), (1, "zlib"), (2, "deflate")):
            for iters in (1, 15):
                z = ref.compress(data, fmt, numiterations=iters)
                entry["streams"]["%s_i%d" % (fname, iters)] = z.hex() if len(z) <= 256 else \
                    {"len": len(z), "sha256": hashlib.sha256(z).hexdigest()}
        out[name] = entry
    with open(os.path.join(ROOT, "tests", "golden", "zopfli_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
