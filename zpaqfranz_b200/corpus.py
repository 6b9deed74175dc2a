"""Deterministic synthetic corpora of the shapes BASELINE.json names.

Pure numpy, vectorised so that the 10 000 x 64 KiB text corpus builds in seconds.  The same bytes
feed the CUDA path, the oracle and the reference, so only determinism matters, not the generator.
"""
import numpy as np

_DICT_WORDS = 4096


def _dictionary():
    rng = np.random.Generator(np.random.PCG64(20260323))
    # word lengths 1..12, skewed short; letters skewed like English
    lens = np.clip(rng.geometric(0.28, _DICT_WORDS) + 1, 2, 14)
    letters = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
    p = 1.0 / np.arange(1, 27) ** 0.9
    p /= p.sum()
    words = []
    for L in lens:
        w = letters[rng.choice(26, size=int(L), p=p)]
        words.append(w.tobytes())
    # each word is stored twice: followed by ' ' and followed by '\n'
    flat = b"".join(w + b" " for w in words) + b"".join(w + b"\n" for w in words)
    wl = np.array([len(w) + 1 for w in words], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(wl)[:-1]])
    off2 = off + wl.sum()
    return np.frombuffer(flat, dtype=np.uint8), np.concatenate([off, off2]), np.concatenate([wl, wl])


_FLAT, _OFF, _LEN = _dictionary()


def text_bytes(seed, nbytes):
    """Word soup: Zipf-ish picks (min of two uniform draws) from a fixed 4096-word dictionary,
    separated by spaces, ~1 in 12 by newlines."""
    rng = np.random.Generator(np.random.PCG64([seed, 77]))
    nwords = nbytes // 3 + 16
    idx = np.minimum(rng.integers(0, _DICT_WORDS, nwords), rng.integers(0, _DICT_WORDS, nwords))
    idx = idx + _DICT_WORDS * (rng.integers(0, 12, nwords) == 0)
    lens = _LEN[idx]
    ends = np.cumsum(lens)
    k = int(np.searchsorted(ends, nbytes) + 1)
    idx, lens, ends = idx[:k], lens[:k], ends[:k]
    starts = ends - lens
    src = np.repeat(_OFF[idx] - starts, lens) + np.arange(int(ends[-1]))
    return _FLAT[src[:nbytes]]


def text_unit(seed, nbytes=65536):
    return text_bytes(seed, nbytes).tobytes()


def random_unit(seed, nbytes=65536):
    return np.random.Generator(np.random.PCG64([seed, 99])).integers(0, 256, nbytes, dtype=np.uint8).tobytes()


def repeats_unit(seed, nbytes=65536):
    """LZ-friendly: a handful of random phrases of 8..400 bytes repeated at random."""
    rng = np.random.Generator(np.random.PCG64([seed, 55]))
    phrases = [rng.integers(0, 256, int(rng.integers(8, 400)), dtype=np.uint8) for _ in range(24)]
    out, tot = [], 0
    while tot < nbytes:
        ph = phrases[int(rng.integers(0, len(phrases)))]
        out.append(ph)
        tot += len(ph)
    return np.concatenate(out)[:nbytes].tobytes()


def mixed_unit(seed, nbytes=65536):
    kind = seed % 4
    if kind == 0:
        return text_unit(seed, nbytes)
    if kind == 1:
        return random_unit(seed, nbytes)
    if kind == 2:
        return repeats_unit(seed, nbytes)
    return bytes(nbytes)


def text_corpus(nunits, unit=65536, seed0=0, out=None):
    """C2: nunits x unit bytes of text, unit u seeded by seed0+u, laid out back to back."""
    if out is None:
        out = np.empty(nunits * unit, dtype=np.uint8)
    group = max(1, (8 << 20) // unit)

    def fill(g):
        k = min(group, nunits - g)
        # one long stream per group, then cut: units stay deterministic per (seed0, g)
        out[g * unit:(g + k) * unit] = text_bytes(1000003 * seed0 + g, k * unit)

    starts = list(range(0, nunits, group))
    if len(starts) > 2:
        from concurrent.futures import ThreadPoolExecutor
        import os
        with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
            list(ex.map(fill, starts))
    else:
        for g in starts:
            fill(g)
    return out
