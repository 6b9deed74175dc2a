"""The decoder fixture the reference ships (AUTOTEST/sha256.zpaq, README.txt:1-45: -m5 archive written on Windows, 256
files of 37 000 bytes, every file NAMED by the SHA-256 of its content) through the device decoder.
The journaling blocks (c, h, i: unmodeled / LZ77 with PCOMP programs) are decoded whole; of the one 9.47 MB data block
(23-component model, a single segment: one warp owns it) the first 150 000 bytes are decoded with zq_decompress_prefix --
enough for the first four files.  Their names must equal SHA-256(content), computed on the device (zq_sha256); every
decoded block must equal what the reference's own decoder produced (tests/golden/sha256_zpaq.json)."""
import hashlib
import json
import os
import struct

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_autotest_archive_decodes_on_device(ctx):
    gold = json.load(open(os.path.join(HERE, "golden", "sha256_zpaq.json")))
    a = np.fromfile(os.path.join(HERE, "golden", "sha256.zpaq"), dtype=np.uint8)
    assert hashlib.sha256(a.tobytes()).hexdigest() == gold["archive_sha256"]
    blocks = gold["blocks"]
    small = [k for k, b in enumerate(blocks) if b["decoded_len"] < (1 << 20)]
    big = [k for k, b in enumerate(blocks) if b["decoded_len"] >= (1 << 20)]
    assert len(big) == 1
    offs = np.array([blocks[k]["offset"] for k in small], dtype=np.uint64)
    lens = np.array([blocks[k]["length"] for k in small], dtype=np.uint32)
    out, ooff, olen = ctx.decompress_blocks(a, offs, lens)
    dec = {}
    for j, k in enumerate(small):
        dec[k] = out[int(ooff[j]): int(ooff[j]) + int(olen[j])].tobytes()
        assert len(dec[k]) == blocks[k]["decoded_len"] and hashlib.sha256(dec[k]).hexdigest() == blocks[k]["sha256"], k
    kb = big[0]
    pre = gold["prefix"]
    out, ooff, olen = ctx.decompress_prefix(a, [blocks[kb]["offset"]], [blocks[kb]["length"]], [pre])
    d = out[: int(olen[0])].tobytes()
    assert len(d) == pre and hashlib.sha256(d).hexdigest() == blocks[kb]["prefix_sha256"]
    # the journal: h block = compressed size, then (SHA-1, size) per fragment; i blocks = (date, name, attr, fragment ids)
    h = dec[small[1]]
    sizes = [struct.unpack("<I", h[4 + 24 * k + 20: 4 + 24 * k + 24])[0] for k in range((len(h) - 4) // 24)]
    starts = np.concatenate([[0], np.cumsum(sizes)])
    assert hashlib.sha1(d[: sizes[0]]).digest() == h[4:24]
    files, idx, p = [], dec[small[2]], 0
    while p < len(idx):
        date = struct.unpack("<q", idx[p:p + 8])[0]
        p += 8
        q = idx.index(0, p)
        name, p = idx[p:q], q + 1
        if date:
            na = struct.unpack("<I", idx[p:p + 4])[0]
            p += 4 + na
            ni = struct.unpack("<I", idx[p:p + 4])[0]
            p += 4
            files.append((name.decode(), struct.unpack("<%dI" % ni, idx[p:p + 4 * ni])))
            p += 4 * ni
    whole = [(n, fr) for n, fr in files if all(int(starts[f]) <= pre for f in fr)]     # starts[f] = end of fragment f (ids are 1-based)
    assert len(whole) >= 4
    content = [b"".join(d[int(starts[f - 1]): int(starts[f])] for f in fr) for _, fr in whole]
    arena = np.frombuffer(b"".join(content), dtype=np.uint8)
    ln = np.array([len(c) for c in content], dtype=np.uint64)
    of = np.concatenate([[0], np.cumsum(ln)[:-1]]).astype(np.uint64)
    dg = ctx.sha256(arena, of, ln)
    for (name, _), c, g in zip(whole, content, dg):
        assert len(c) == 37000
        assert g.tobytes().hex().upper() == name == hashlib.sha256(c).hexdigest().upper()


def test_journal_blocks_of_the_fixture_are_rewritten_byte_for_byte(ctx):
    """The "c" and "i" blocks of AUTOTEST/sha256.zpaq, parsed back into (date, name, attribute bytes, fragment ids)
    records and written again by zq_journal_header / zq_journal_index, are the archive's own bytes (the writer of
    Jidac::add, zpaqfranz.cpp:71521-71540 and 122915-123100; 16 000-byte block rule included: three index blocks)."""
    gold = json.load(open(os.path.join(HERE, "golden", "sha256_zpaq.json")))
    a = np.fromfile(os.path.join(HERE, "golden", "sha256.zpaq"), dtype=np.uint8)
    blocks = gold["blocks"]
    raw = a.tobytes()
    names = []
    for b in blocks:
        blk = raw[b["offset"]: b["offset"] + b["length"]]
        i = blk.find(b"jDC")
        names.append(blk[i: i + 28])
    kinds = [n[17:18] for n in names]
    assert kinds == [b"c", b"d", b"h", b"i", b"i", b"i"]
    date14 = names[0][3:17].decode()
    small = [k for k in range(len(blocks)) if kinds[k] != b"d"]
    offs = np.array([blocks[k]["offset"] for k in small], dtype=np.uint64)
    lens = np.array([blocks[k]["length"] for k in small], dtype=np.uint32)
    out, ooff, olen = ctx.decompress_blocks(a, offs, lens)
    dec = {k: out[int(ooff[j]): int(ooff[j]) + int(olen[j])].tobytes() for j, k in enumerate(small)}
    cdata = struct.unpack("<q", dec[0])[0]
    first = int(names[0][18:28])
    assert ctx.journal_header(date14, cdata, first) == raw[: blocks[0]["length"]]
    records = []
    for k in (3, 4, 5):
        idx, p = dec[k], 0
        while p < len(idx):
            date = struct.unpack("<q", idx[p:p + 8])[0]
            p += 8
            q = idx.index(0, p)
            name, p = idx[p:q], q + 1
            attr, fr = b"", ()
            if date:
                na = struct.unpack("<I", idx[p:p + 4])[0]
                attr = idx[p + 4: p + 4 + na]
                p += 4 + na
                ni = struct.unpack("<I", idx[p:p + 4])[0]
                fr = struct.unpack("<%dI" % ni, idx[p + 4: p + 4 + 4 * ni])
                p += 4 + 4 * ni
            records.append((date, name, attr, fr))
    assert len(records) >= 256
    got, nb = ctx.journal_index(date14, records)
    assert nb == 3
    assert got == raw[blocks[3]["offset"]:]
