"""SURVEY 8(f)#4 / BASELINE config C5: the UNMODIFIED zopflipng (zopflipng_lib.cc + LodePNG, compiled from
/root/reference by oracle/Makefile) linked against the product's libzopfli.so.1 must write the same PNG
as zopflipng over the reference's own zopfli sources.  zopflipng reaches the library through LodePNG's
custom_deflate hook: CustomPNGDeflate -> ZopfliDeflate(&options, 2, 1, ...) (zopflipng_lib.cc:47-63)."""
import os
import subprocess
import tempfile

import pytest

from zopfli_b200 import corpus

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "zopflipng_ref")
OUR_BIN = os.path.join(ROOT, "oracle", "_ref", "zopflipng_b200")
ZOIDBERG = os.path.join(ROOT, "oracle", "_ref", "zoidberg.png")


def _run(binary, args, src, dst):
    subprocess.check_call([binary, "-y"] + args + [src, dst], stdout=subprocess.DEVNULL)
    return open(dst, "rb").read()


@pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.exists(OUR_BIN)), reason="oracle/_ref not built")
@pytest.mark.parametrize("case,args", [("zoidberg", []), ("zoidberg", ["--iterations=3", "--filters=0me"]),
                                       ("synth512", ["--filters=01"]), ("synth1024", [])])
def test_zopflipng_over_the_product_library_writes_the_reference_png(case, args):
    with tempfile.TemporaryDirectory() as td:
        if case == "zoidberg":
            src = ZOIDBERG  # go/zopflipng/testdata/zoidberg.png, the reference's own test image
        else:
            n = int(case[5:])
            src = os.path.join(td, "in.png")
            open(src, "wb").write(corpus.write_png_rgba(corpus.synth_image_rgba(n, n, 5)))
        want = _run(REF_BIN, args, src, os.path.join(td, "ref.png"))
        got = _run(OUR_BIN, args, src, os.path.join(td, "b200.png"))
        assert got == want, (case, args, len(got), len(want))
        assert len(got) < os.path.getsize(src)  # go/zopflipng/zopflipng_test.go:32-34
