"""The N>1 path of bench.py on CPU: world_size-2 `gloo`, each rank owns a shard of master blocks,
receives the 32 KiB halo of its left neighbour, produces a span, rank 0 gathers and splices
(SURVEY 8(e)).  The per-rank compressor is the host-logic test library (product host sources +
oracle-backed mock engine), so this checks the sharding protocol, not the kernels."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch, torch.distributed as dist
import zopfli_b200 as zb
from zopfli_b200 import corpus
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
lib = zb.Library(os.path.join(%(root)r, "tests", "_build", "libzopfli_hostmock.so"))
SHARD = 1000000  # one master block per rank keeps the CPU test short
data = corpus.synth_text(SHARD, 40 + rank)
halo = 32768 if rank > 0 else 0
buf = np.zeros(halo + SHARD + 64, np.uint8)
buf[halo:halo + SHARD] = np.frombuffer(data, np.uint8)
tail = torch.from_numpy(buf[halo + SHARD - 32768: halo + SHARD].copy())
recv = torch.empty(32768, dtype=torch.uint8)
reqs = []
if rank + 1 < world: reqs.append(dist.isend(tail, rank + 1))
if rank > 0: reqs.append(dist.irecv(recv, rank - 1))
for r in reqs: r.wait()
if rank > 0: buf[:halo] = recv.numpy()
span = lib.deflate_span_ptr(buf.ctypes.data, halo + SHARD, halo, halo + SHARD, final=int(rank == world - 1), numiterations=1)
crc = lib.crc32(buf.ctypes.data + halo, SHARD)
meta = torch.tensor([len(span), crc, SHARD], dtype=torch.int64)
metas = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
dist.all_gather(metas, meta)
mx = max(int(m[0]) for m in metas)
sp = torch.zeros(mx, dtype=torch.uint8); sp[:len(span)] = torch.frombuffer(bytearray(span), dtype=torch.uint8)
gathered = [torch.zeros(mx, dtype=torch.uint8) for _ in range(world)] if rank == 0 else None
dist.gather(sp, gathered, dst=0)
if rank == 0:
    spans = [g[:int(m[0])].numpy().tobytes() for g, m in zip(gathered, metas)]
    body, _ = lib.splice_spans(spans, prefix=bytes([31, 139, 8, 0, 0, 0, 0, 0, 2, 3]))
    c, tot = 0, 0
    for i, m in enumerate(metas):
        c = int(m[1]) if i == 0 else lib.crc32_combine(c, int(m[1]), int(m[2]))
        tot += int(m[2])
    out = body + int(c).to_bytes(4, "little") + int(tot & 0xffffffff).to_bytes(4, "little")
    open(os.environ["ZB_OUT"], "wb").write(out)
dist.destroy_process_group()
'''


def test_two_rank_shards_equal_single_stream(ref, tmp_path):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "hostmock")])
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    out = tmp_path / "out.gz"
    env = dict(os.environ, ZB_OUT=str(out), ZOPFLI_B200_THREADS="2")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                           "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)], env=env, timeout=900)
    from zopfli_b200 import corpus
    whole = corpus.synth_text(1000000, 40) + corpus.synth_text(1000000, 41)
    got = out.read_bytes()
    assert got == ref.compress(whole, 0, numiterations=1)
    import gzip
    assert gzip.decompress(got) == whole


WORKER2 = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch, torch.distributed as dist
import zopfli_b200 as zb
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
lib = zb.Library(os.path.join(%(root)r, "tests", "_build", "libzopfli_hostmock.so"))
data = open(os.environ["ZB_IN"], "rb").read() if rank == 0 else None
insize = int(os.environ["ZB_INSIZE"])
# --- scatter: rank 0 sends every rank its shard + dictionary (csrc/dist.cpp, dist_layout.hpp) ---
a, b, base = lib.dist_shard(insize, world, rank)
if rank == 0:
    for q in range(1, world):
        qa, qb, qbase = lib.dist_shard(insize, world, q)
        if qb > qbase:
            dist.send(torch.frombuffer(bytearray(data[qbase:qb]), dtype=torch.uint8), q)
    mine = data[:b]
    base = 0
else:
    t = torch.empty(b - base, dtype=torch.uint8)
    if b > base:
        dist.recv(t, 0)
    mine = t.numpy().tobytes()
buf = np.zeros(len(mine) + 64, np.uint8)
buf[:len(mine)] = np.frombuffer(mine, np.uint8)
span = lib.deflate_span_ptr(buf.ctypes.data, len(mine), a - base, b - base, final=int(b == insize), numiterations=1) if b > a else b""

def place(phase):   # this rank's blocks with their first bit at bit `phase` of byte 0 -> (bytes, bits)
    if not span:
        return b"", 0
    out, bp = lib.splice_spans([span], prefix=b"\0", bp0=phase) if phase else lib.splice_spans([span])
    end = (len(out) - 1) * 8 + bp if bp else len(out) * 8
    return out, end - phase

# --- placement: lengths for all 8 phases, all_gather, absolute bit offsets ---
len8 = torch.tensor([place(p)[1] for p in range(8)], dtype=torch.int64)
alls = [torch.zeros(8, dtype=torch.int64) for _ in range(world)]
dist.all_gather(alls, len8)
start = lib.dist_placement(np.stack([t.numpy() for t in alls]).astype(np.uint64))
local, nbits = place(start[rank] & 7)
assert nbits == start[rank + 1] - start[rank]
# --- gather: bytes go to their final place; the byte two ranks share is ORed ---
sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
dist.all_gather(sizes, torch.tensor([len(local)], dtype=torch.int64))
mx = max(int(s) for s in sizes)
pad = torch.zeros(mx, dtype=torch.uint8)
pad[:len(local)] = torch.frombuffer(bytearray(local), dtype=torch.uint8)
got = [torch.zeros(mx, dtype=torch.uint8) for _ in range(world)] if rank == 0 else None
dist.gather(pad, got, dst=0)
if rank == 0:
    out = bytearray((start[world] + 7) // 8)
    for q in range(world):
        nb = int(sizes[q])
        if nb == 0:
            continue
        piece = got[q][:nb].numpy().tobytes()
        gs = start[q] >> 3
        out[gs] |= piece[0]
        out[gs + 1: gs + nb] = piece[1:]
    open(os.environ["ZB_OUT"], "wb").write(bytes(out))
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 3])
def test_library_protocol_scatter_placement_gather(ref, tmp_path, world):
    """The protocol of csrc/dist.cpp over gloo instead of NCCL: shard ranges, per-phase stream lengths,
    absolute bit offsets (dist_layout.hpp through the C ABI) and the in-place gather with the shared
    boundary byte -- stored blocks included, whose padding makes a rank's length depend on its phase."""
    from zopfli_b200 import corpus
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "hostmock")])
    data = corpus.synth_text(1100000, 50) + corpus.random_bytes(500000, 3) + corpus.synth_text(700001, 51)
    src = tmp_path / "in.bin"
    src.write_bytes(data)
    script = tmp_path / "worker2.py"
    script.write_text(WORKER2 % {"root": ROOT})
    out = tmp_path / "out.deflate"
    env = dict(os.environ, ZB_OUT=str(out), ZB_IN=str(src), ZB_INSIZE=str(len(data)), ZOPFLI_B200_THREADS="2")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world,
                           "--master-addr", "127.0.0.1", "--master-port", str(29540 + world), str(script)], env=env, timeout=900)
    assert out.read_bytes() == ref.compress(data, 2, numiterations=1)
