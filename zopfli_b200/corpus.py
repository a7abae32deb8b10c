"""Deterministic synthetic inputs for parity tests and bench.py (no corpora ship with the box).

Shapes follow SURVEY.md section 8(d) / Appendix C:
  synth_text    -- a stream of wiki-like pages (XML page headers, per-page topic vocabulary over a
                   Zipf word soup, prose / link-list / table / citation / foreign-language
                   sections), so the statistics drift and the block splitter finds real split
                   points; the stand-in for enwik8 (config C2) and the 1 GiB web-text corpus (C3).
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


_SECTION_KINDS = ("prose", "links", "table", "refs", "foreign")


def _tokens_to_bytes(tok, space, flat, all_starts, all_lens):
    tl = all_lens[tok]
    tot = tl + space
    offs = np.concatenate([[0], np.cumsum(tot)[:-1]])
    total = int(offs[-1] + tot[-1]) if len(tok) else 0
    buf = np.full(total, 32, dtype=np.uint8)
    idx_tok = np.repeat(np.arange(len(tok)), tl)
    within = np.arange(int(tl.sum())) - np.repeat(np.cumsum(tl) - tl, tl)
    buf[offs[idx_tok] + within] = flat[all_starts[tok][idx_tok] + within]
    return buf


def synth_text(nbytes: int, seed: int = 2) -> bytes:
    """enwik8-like stand-in: a stream of "wiki pages" of very different sizes, each with its own
    topic vocabulary and a mix of section kinds (prose, link lists, tables, citation blocks,
    foreign-language text) wrapped in XML page headers.  The statistics drift from page to page and
    section to section, so the block splitter finds real split points (uniform word soup leaves
    1 MB blocks unsplit and overstates the critical path).  ~2.6-3 bits/byte."""
    rng = np.random.default_rng(seed)
    nwords = 65536
    flat, starts, lens = _vocab(rng, nwords)
    extra = _MARKUP + [b".", b",", b";", b"\n\n", b"\n", b"?", b"(", b")", b":", b"1", b"19", b"200"]
    ex_flat = np.frombuffer(b"".join(extra), dtype=np.uint8)
    ex_lens = np.array([len(e) for e in extra], dtype=np.int64)
    ex_starts = len(flat) + np.concatenate([[0], np.cumsum(ex_lens)[:-1]])
    flat = np.concatenate([flat, ex_flat])
    all_starts = np.concatenate([starts[:-1], ex_starts])
    all_lens = np.concatenate([lens, ex_lens])
    n_markup = len(_MARKUP)
    n_punct = len(extra) - n_markup
    ranks = np.arange(1, nwords + 1, dtype=np.float64)
    cdf = np.cumsum(ranks ** -1.1)
    cdf /= cdf[-1]
    foreign_map = np.arange(256, dtype=np.uint8)
    foreign_map[97:123] = rng.permutation(np.arange(97, 123)).astype(np.uint8)  # another letter distribution
    out = np.empty(nbytes + 1024, dtype=np.uint8)
    pos = 0
    page_id = 1000
    while pos < nbytes:
        page_id += int(rng.integers(1, 40))
        page_len = int(min(200000, max(600, rng.lognormal(9.0, 1.1))))
        topic = rng.integers(0, nwords, 300)
        title = flat[all_starts[topic[0]]: all_starts[topic[0]] + all_lens[topic[0]]].tobytes()
        hdr = (b"  <page>\n    <title>" + title.capitalize() + b"</title>\n    <id>%d</id>\n    <revision>\n"
               b"      <id>%d</id>\n      <timestamp>200%d-%02d-%02dT%02d:%02d:%02dZ</timestamp>\n"
               b"      <contributor>\n        <username>" % (page_id, page_id * 7 + 13, rng.integers(2, 7),
                                                              rng.integers(1, 13), rng.integers(1, 29), rng.integers(0, 24),
                                                              rng.integers(0, 60), rng.integers(0, 60))
               + flat[all_starts[topic[1]]: all_starts[topic[1]] + all_lens[topic[1]]].tobytes()
               + b"</username>\n        <id>%d</id>\n      </contributor>\n      <text xml:space=\"preserve\">" % rng.integers(1, 99999))
        parts = [np.frombuffer(hdr, dtype=np.uint8)]
        made = len(hdr)
        # page-level mixture of section kinds
        kind_p = rng.dirichlet([4.0, 0.7, 0.5, 0.6, 0.25])
        while made < page_len:
            kind = _SECTION_KINDS[int(rng.choice(5, p=kind_p))]
            sec_len = int(min(page_len - made + 200, max(200, rng.lognormal(7.3, 0.9))))
            nw = sec_len // 5 + 8
            r = np.searchsorted(cdf, rng.random(nw))
            ids = np.minimum(r, nwords - 1)
            tsel = rng.random(nw) < (0.45 if kind != "foreign" else 0.2)
            ids[tsel] = topic[rng.integers(0, len(topic), int(tsel.sum()))]
            tok = ids.copy()
            space = np.ones(nw, dtype=np.int64)
            u = rng.random(nw)
            if kind in ("prose", "foreign"):
                mk = u < 0.02
                pk = (u >= 0.02) & (u < 0.14)
                tok[mk] = nwords + rng.integers(0, n_markup, int(mk.sum()))
                tok[pk] = nwords + n_markup + rng.choice(n_punct, int(pk.sum()), p=_punct_p(n_punct))
                space[:-1][pk[1:]] = 0
                buf = _tokens_to_bytes(tok, space, flat, all_starts, all_lens)
                if kind == "foreign":
                    buf = foreign_map[buf]
            elif kind == "links":
                # "* [[word word]]\n" lists
                nl = nw // 3
                w = _tokens_to_bytes(tok[: nl * 2], np.tile([1, 0], nl), flat, all_starts, all_lens)
                lines = []
                o = 0
                l2 = (all_lens[tok[: nl * 2]] + np.tile([1, 0], nl)).reshape(nl, 2).sum(1)
                for ln in l2[: min(nl, 400)]:
                    lines.append(b"* [[" + w[o:o + ln].tobytes() + b"]]\n")
                    o += ln
                buf = np.frombuffer(b"".join(lines), dtype=np.uint8)
            elif kind == "table":
                rows = min(300, nw // 4)
                cells = rng.integers(0, 100000, (rows, 3))
                words = _tokens_to_bytes(tok[:rows], np.zeros(rows, dtype=np.int64), flat, all_starts, all_lens)
                wl = all_lens[tok[:rows]]
                o = 0
                lines = [b"{| class=\"wikitable\"\n"]
                for i in range(rows):
                    lines.append(b"|-\n| %d || %d.%d || " % (cells[i, 0], cells[i, 1] % 1000, cells[i, 2] % 10)
                                 + words[o:o + wl[i]].tobytes() + b"\n")
                    o += wl[i]
                lines.append(b"|}\n")
                buf = np.frombuffer(b"".join(lines), dtype=np.uint8)
            else:  # refs
                nr = min(120, nw // 8)
                w = _tokens_to_bytes(tok[: nr * 4], np.zeros(nr * 4, dtype=np.int64), flat, all_starts, all_lens)
                wl = all_lens[tok[: nr * 4]].reshape(nr, 4)
                o = 0
                lines = []
                for i in range(nr):
                    a0, a1, a2, a3 = (int(x) for x in wl[i])
                    lines.append(b"<ref>{{cite web|url=http://www." + w[o:o + a0].tobytes() + b".com/" + w[o + a0:o + a0 + a1].tobytes()
                                 + b"|title=" + w[o + a0 + a1:o + a0 + a1 + a2].tobytes().capitalize() + b" "
                                 + w[o + a0 + a1 + a2:o + a0 + a1 + a2 + a3].tobytes()
                                 + b"|accessdate=200%d-%02d-%02d}}</ref>\n" % (rng.integers(2, 7), rng.integers(1, 13), rng.integers(1, 29)))
                    o += a0 + a1 + a2 + a3
                buf = np.frombuffer(b"".join(lines), dtype=np.uint8)
            buf = buf[:sec_len]
            parts.append(buf)
            made += len(buf)
            if rng.random() < 0.5:
                head = b"\n\n== " + flat[all_starts[topic[2]]: all_starts[topic[2]] + all_lens[topic[2]]].tobytes().capitalize() + b" ==\n"
                parts.append(np.frombuffer(head, dtype=np.uint8))
                made += len(head)
        parts.append(np.frombuffer(b"</text>\n    </revision>\n  </page>\n", dtype=np.uint8))
        page = np.concatenate(parts)
        take = min(len(page), nbytes - pos)
        out[pos:pos + take] = page[:take]
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


def synth_image_rgba(w: int, h: int, seed: int = 5) -> np.ndarray:
    """SURVEY 8(d) C5 stand-in: smooth gradients + 8x8-tile noise + flat alpha regions, (h, w, 4) uint8."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    img = np.zeros((h, w, 4), np.uint8)
    img[..., 0] = (x * 255 // max(1, w - 1)).astype(np.uint8)
    img[..., 1] = (y * 255 // max(1, h - 1)).astype(np.uint8)
    img[..., 2] = (((x + y) // 3) & 255).astype(np.uint8)
    tiles = rng.integers(0, 24, ((h + 7) // 8, (w + 7) // 8, 3), dtype=np.uint8)
    img[..., :3] += np.kron(tiles, np.ones((8, 8, 1), np.uint8))[:h, :w]
    alpha = np.full((h, w), 255, np.uint8)
    alpha[h // 4: h // 2, w // 8: w // 2] = 0
    alpha[h // 2:, 3 * w // 4:] = 128
    img[..., 3] = alpha
    return img


def write_png_rgba(img: np.ndarray) -> bytes:
    """minimal PNG writer (8-bit RGBA, filter type 0 on every scanline, zlib level 1)"""
    import struct
    import zlib
    h, w, _ = img.shape
    raw = np.zeros((h, 1 + w * 4), np.uint8)
    raw[:, 1:] = img.reshape(h, w * 4)

    def chunk(tag, body):
        return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xffffffff)

    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)) +
            chunk(b"IDAT", zlib.compress(raw.tobytes(), 1)) + chunk(b"IEND", b""))
