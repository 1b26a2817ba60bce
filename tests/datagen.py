"""Seeded synthetic inputs for the CPU-side tests (numpy; the GPU bench has its own device generator).

text_like   : "enwik-style" Zipf word text (SURVEY.md section 8d, config C2/C5 generator, host version)
binary_records : repeating fixed-size records with counters (config C4 'binary-structured')
"""
import numpy as np


def _vocab(n_words, rng):
    lens = rng.integers(2, 11, n_words)
    letters = rng.integers(0, 26, int(lens.sum())).astype(np.uint8) + ord("a")
    offs = np.concatenate(([0], np.cumsum(lens)))
    return [letters[offs[i]:offs[i + 1]].tobytes() for i in range(n_words)]


def text_like(n_bytes, seed=1234, n_words=50000):
    rng = np.random.default_rng(seed)
    vocab = _vocab(n_words, np.random.default_rng(1234))
    out = bytearray()
    # Zipf(1)-like ranks: rank = floor(N**u)
    while len(out) < n_bytes:
        u = rng.random(4096)
        ranks = np.minimum((n_words ** u).astype(np.int64) - 1, n_words - 1)
        marks = rng.random(4096)
        for r, m in zip(ranks, marks):
            out += vocab[r]
            if m < 0.02:
                out += b".\n" if m < 0.01 else b" <" + str(int(m * 1e6)).encode() + b">"
            out += b" "
    return bytes(out[:n_bytes])


def binary_records(n_bytes, seed=5, rec=48):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, rec, dtype=np.uint8)
    n = (n_bytes + rec - 1) // rec
    arr = np.tile(base, (n, 1))
    arr[:, 0:4] = np.arange(n, dtype=np.uint32).view(np.uint8).reshape(n, 4)
    arr[:, 8] = rng.integers(0, 4, n, dtype=np.uint8)
    return arr.tobytes()[:n_bytes]


def random_bytes(n_bytes, seed=1):
    return np.random.default_rng(seed).integers(0, 256, n_bytes, dtype=np.uint8).tobytes()


def mixed(n_bytes, seed=3):
    """70% text, 20% records, 10% incompressible, in 4 KiB runs."""
    rng = np.random.default_rng(seed)
    parts, total, i = [], 0, 0
    while total < n_bytes:
        k = rng.random()
        ln = int(rng.integers(512, 8192))
        if k < 0.7:
            parts.append(text_like(ln, seed=seed * 1000 + i))
        elif k < 0.9:
            parts.append(binary_records(ln, seed=seed * 1000 + i))
        else:
            parts.append(random_bytes(ln, seed=seed * 1000 + i))
        total += ln
        i += 1
    return b"".join(parts)[:n_bytes]


def near_period_traps(seed=5, sites=400):
    """Sites where 7 of the 8 bytes at a position continue a period of 3 ("abc" + "abcabcaZ") and the same 8 bytes occurred earlier
    followed by a different byte, so the table's match is exactly 8 long: a near-source rule that trusts 7 bytes of the period would
    claim an 8-byte match at distance 3 whose last byte is wrong (found by tests/emu/fuzz_kernels.py, round 2)."""
    import random
    rng = random.Random(seed)
    out = bytearray(random_bytes(6000, seed + 1))
    for k in range(sites):
        a, b, c, z = (rng.randrange(256) for _ in range(4))
        blk = bytes([a, b, c, a, b, c, a, z])
        out += random_bytes(rng.randrange(5, 40), 100 + k) + blk + bytes([rng.randrange(256)]) + random_bytes(rng.randrange(5, 300), 900 + k)
        out += bytes([a, b, c]) + blk + bytes([rng.randrange(256)]) + random_bytes(rng.randrange(5, 40), 1900 + k)
    return bytes(out)
