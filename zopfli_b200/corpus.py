"""Deterministic synthetic inputs for parity tests and bench.py (no corpora ship with the box).

Shapes follow SURVEY.md section 8(d) / Appendix C:
  synth_text    -- Zipf-distributed word soup with punctuation, ~3 % markup tokens and topic
                   drift every 64 KiB (so the block splitter finds real split points); the
                   stand-in for enwik8 (config C2) and the 1 GiB web-text corpus (C3).
  synth_binary  -- ELF-like mix of constant runs, mutated records and text (config C4).
  adv_*         -- adversarial inputs that exercise the exact walk semantics of
                   ZopfliFindLongestMatch (/root/reference/src/zopfli/lz77.c:407-542): the
                   8192-hop cap, the hash-chain switch, `same` saturation and the long-run
                   shortcut of squeeze.c:251-271.
All generators are pure functions of their seed.
"""
from __future__ import annotations

import random

import numpy as np

_LETTERS = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
_LETTER_W = np.array([12.7, 9.1, 8.2, 7.5, 7.0, 6.7, 6.3, 6.1, 6.0, 4.3, 4.0, 2.8, 2.8, 2.4, 2.4,
                      2.2, 2.0, 2.0, 1.9, 1.5, 1.0, 0.8, 0.15, 0.15, 0.1, 0.07])
_MARKUP = [b"<p>", b"</p>", b"[[", b"]]", b"&amp;", b"<ref>", b"</ref>", b"{{cite", b"}}", b"==",
           b"&quot;", b"<br />", b"|", b"''", b"*"]


def _vocab(rng: np.random.Generator, nwords: int):
    lens = 2 + rng.poisson(4.0, nwords)
    lens = np.clip(lens, 1, 18).astype(np.int64)
    starts = np.zeros(nwords + 1, dtype=np.int64)
    np.cumsum(lens, out=starts[1:])
    p = _LETTER_W / _LETTER_W.sum()
    flat = _LETTERS[rng.choice(len(_LETTERS), size=int(starts[-1]), p=p)]
    return flat, starts, lens


def synth_text(nbytes: int, seed: int = 2) -> bytes:
    """~2.7-3 bits/byte under zopfli, enwik8-like structure. Vectorised: ~10 MB/s."""
    rng = np.random.default_rng(seed)
    nwords = 65536
    flat, starts, lens = _vocab(rng, nwords)
    # append markup tokens and punctuation "words" to the vocabulary
    extra = _MARKUP + [b".", b",", b";", b"\n\n", b"\n", b"?", b"(", b")", b":", b"1", b"19", b"200"]
    ex_flat = np.frombuffer(b"".join(extra), dtype=np.uint8)
    ex_lens = np.array([len(e) for e in extra], dtype=np.int64)
    ex_starts = len(flat) + np.concatenate([[0], np.cumsum(ex_lens)[:-1]])
    flat = np.concatenate([flat, ex_flat])
    all_starts = np.concatenate([starts[:-1], ex_starts])
    all_lens = np.concatenate([lens, ex_lens])
    n_markup = len(_MARKUP)
    # Zipf(s=1.1) rank weights
    ranks = np.arange(1, nwords + 1, dtype=np.float64)
    w = ranks ** -1.1
    cdf = np.cumsum(w / w.sum())
    perm = rng.permutation(nwords)  # rank -> word id, re-ranked by topic drift
    out = np.empty(nbytes + 64, dtype=np.uint8)
    pos = 0
    chunk = 64 * 1024
    while pos < nbytes:
        # topic drift: re-rank 5 % of the vocabulary every 64 KiB
        k = nwords // 20
        a = rng.integers(0, nwords, k)
        b = rng.integers(0, min(nwords, 4096), k)  # promote some words into the head
        perm[a], perm[b] = perm[b].copy(), perm[a].copy()
        nw = chunk // 5 + 64
        r = np.searchsorted(cdf, rng.random(nw))
        ids = perm[np.minimum(r, nwords - 1)]
        kind = rng.random(nw)
        # 3 % markup, ~12 % punctuation, rest words
        tok = ids.copy()
        mk = kind < 0.03
        tok[mk] = nwords + rng.integers(0, n_markup, int(mk.sum()))
        pk = (kind >= 0.03) & (kind < 0.15)
        tok[pk] = nwords + n_markup + rng.choice(len(extra) - n_markup, int(pk.sum()),
                                                 p=_punct_p(len(extra) - n_markup))
        tl = all_lens[tok]
        # words are followed by a space unless the next token is punctuation
        space = np.ones(nw, dtype=np.int64)
        space[:-1][pk[1:]] = 0
        tot = tl + space
        offs = np.concatenate([[0], np.cumsum(tot)[:-1]])
        total = int(offs[-1] + tot[-1])
        buf = np.full(total, 32, dtype=np.uint8)
        # gather word bytes
        idx_tok = np.repeat(np.arange(nw), tl)
        within = np.arange(int(tl.sum())) - np.repeat(np.cumsum(tl) - tl, tl)
        buf[offs[idx_tok] + within] = flat[all_starts[tok][idx_tok] + within]
        take = min(total, nbytes - pos)
        out[pos:pos + take] = buf[:take]
        pos += take
    return out[:nbytes].tobytes()


def _punct_p(n):
    p = np.array([6, 6, 1, 1, 1.5, 0.5, 0.6, 0.6, 0.8, 0.6, 0.5, 0.5][:n], dtype=np.float64)
    return p / p.sum()


def synth_binary(nbytes: int, seed: int = 4) -> bytes:
    """ELF-like mix: 40 % constant runs (260-4096), 30 % repeated 16-64 byte records with 1-2
    mutated bytes, 30 % text fragments."""
    rng = np.random.default_rng(seed)
    text = np.frombuffer(synth_text(1 << 20, seed + 100), dtype=np.uint8)
    parts = []
    total = 0
    while total < nbytes:
        u = rng.random()
        if u < 0.4:
            n = int(rng.integers(260, 4097))
            v = 0 if rng.random() < 0.7 else int(rng.integers(1, 256))
            part = np.full(n, v, dtype=np.uint8)
        elif u < 0.7:
            rl = int(rng.integers(16, 65))
            reps = int(rng.integers(8, 200))
            rec = rng.integers(0, 256, rl, dtype=np.uint8)
            part = np.tile(rec, reps)
            nm = reps * int(rng.integers(1, 3))
            part[rng.integers(0, len(part), nm)] = rng.integers(0, 256, nm, dtype=np.uint8)
        else:
            n = int(rng.integers(256, 8192))
            o = int(rng.integers(0, len(text) - n))
            part = text[o:o + n]
        parts.append(part)
        total += len(part)
    return np.concatenate(parts)[:nbytes].tobytes()


# ---- SURVEY Appendix C adversarial generators (Python `random`, exact recipes) ----

def adv_collide() -> bytes:
    """Hash collisions (first bytes differ only in bits 5-7) + chain cap; diverges from brute force."""
    random.seed(11)
    vs = [0x01, 0x21, 0x41, 0x61, 0x81, 0xa1, 0xc1, 0xe1]
    b = bytearray()
    while len(b) < 150000:
        b.append(random.choice(vs))
        b += b"bc"
    return bytes(b)


def adv_chain_and_runs():
    """adv_chain then adv_runs from the same interpreter state (App. C)."""
    random.seed(7)
    b = bytearray()
    while len(b) < 120000:
        b += b"abc" + bytes([random.choice(b"ABCDEFGHIJKLMNOP")])
    chain = bytes(b)
    b = bytearray()
    while len(b) < 200000:
        b += bytes(random.randint(260, 900)) + bytes([random.randint(1, 255)])
    return chain, bytes(b)


def adv_chain() -> bytes:
    return adv_chain_and_runs()[0]


def adv_runs() -> bytes:
    return adv_chain_and_runs()[1]


def adv_longrun() -> bytes:
    """Runs > 65535 so `same` saturates; heavy long-run shortcut."""
    random.seed(3)
    return (b"x" * 70000 + bytes(random.getrandbits(8) for _ in range(500)) + b"x" * 66000 +
            b"end" + b"\0" * 140000)


def go_case_foobar() -> bytes:
    """go/zopfli/zopfli_test.go:36-38"""
    return b"compressthis" + b"_foobar" * 1000 + b"$"


def random_bytes(n: int, seed: int = 1) -> bytes:
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8).tobytes()


def mixed_small(n: int, seed: int = 9) -> bytes:
    """Small mixed sample: text, a byte run, a few repeated records, random tail."""
    rng = np.random.default_rng(seed)
    t = synth_text(max(n, 4096), seed)
    parts = [t[: n // 2], b"\0" * (n // 8), bytes(rng.integers(0, 256, 24, dtype=np.uint8)) * (n // 96 + 1),
             rng.integers(0, 256, n // 8 + 8, dtype=np.uint8).tobytes(), t[n // 2:]]
    return b"".join(parts)[:n]
