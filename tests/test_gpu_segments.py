"""Caller-supplied models (libzpaq::Compressor driven directly) and the C++ mirror of the libzpaq classes.

zq_compress_segments == writeTag / startBlock(hcomp) / startSegment / postProcess / compress / endSegment / endBlock
(Z:15970-16187) once per unit; checked bit for bit against the reference's own Compressor class."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import zpaqfranz_b200 as zqmod
from zpaqfranz_b200 import corpus

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

UNITS = [b"", b"x", b"abracadabra" * 40, bytes(5000), corpus.text_unit(3, 30000), corpus.random_unit(5, 2000),
         corpus.mixed_unit(7, 20000), corpus.text_unit(9, 70000)]

# hand-written models in ZPAQL: every component type, HCOMP with if/else, loops and computed contexts
CONFIGS = {
    "cm_o1": "comp 2 2 0 0 1 0 cm 18 255 hcomp *d=a a<<= 8 *d=a halt end",
    "icm_chain_mix2_sse": ("comp 3 3 0 0 7 0 icm 12 1 isse 14 0 2 cm 16 32 3 match 14 16 4 mix2 8 1 2 20 255 "
                           "5 sse 10 4 16 255 6 avg 4 5 96 "
                           "hcomp c++ *c=a b=c a=0 d= 0 hash *d=a d++ b-- hash *d=a d++ a=*c a<<= 9 *d=a d++ "
                           "b=c a=0 hash b-- hash b-- hash *d=a d++ a=*c *d=a d++ a=*c a>>= 3 *d=a halt end"),
    "branches": ("comp 2 4 0 0 3 0 cm 14 8 1 cm 16 20 2 mix 8 0 2 30 255 "
                 "hcomp *c=a c++ a== 32 if d= 0 *d=0 else d= 0 a+=*d a*= 73 *d=a endif "
                 "d= 1 a=*c a^= 255 *d=a d= 2 a=c a&= 7 *d=a halt end"),
    "const_only_mix": "comp 0 0 0 0 3 0 const 200 1 cm 10 4 2 mix 0 0 2 14 0 hcomp halt end",
}


def _pack(blobs):
    lens = np.array([len(b) for b in blobs], dtype=np.uint32)
    offs = np.concatenate([[0], np.cumsum(lens.astype(np.uint64))[:-1]]).astype(np.uint64)
    return np.frombuffer(b"".join(blobs) + b"\0", dtype=np.uint8), offs, lens


@pytest.mark.parametrize("name", ["level1", "level2"] + sorted(CONFIGS))
def test_segments_match_reference_compressor(ctx, ref, name):
    if name.startswith("level"):
        header = zqmod.assemble_config(zqmod.model_config(int(name[-1])))["header"]
    else:
        header = zqmod.assemble_config(CONFIGS[name])["header"]
    arena, offs, lens = _pack(UNITS)
    digests = np.frombuffer(b"".join(hashlib.sha1(u).digest() for u in UNITS), dtype=np.uint8).reshape(-1, 20)
    for sha, tag in ((digests, True), (None, False)):
        out, ooff, olen = ctx.compress_segments(arena, offs, lens, header, filename="seg", comment="as given", sha1=sha, tag=tag)
        for i, u in enumerate(UNITS):
            want = ref.compress_segment(u, header=header, filename="seg", comment="as given",
                                        sha1=None if sha is None else digests[i].tobytes(), tag=tag)
            got = out[int(ooff[i]): int(ooff[i]) + int(olen[i])].tobytes()
            assert got == want, (name, i, len(u), len(got), len(want))
    # and the device decoder restores what it wrote (size passed explicitly: the comment carries none)
    out, ooff, olen = ctx.compress_segments(arena, offs, lens, header, sha1=digests)
    dec, doff, dlen, used, tr = ctx.decompress_blocks(out, ooff, olen, expect_len=lens, details=True)
    for i, u in enumerate(UNITS):
        assert dec[int(doff[i]): int(doff[i]) + int(dlen[i])].tobytes() == u
        assert int(used[i]) == int(olen[i]) - 1 and tr[i, 0] == 1 and tr[i, 1:].tobytes() == digests[i].tobytes()


def test_segments_with_pcomp_and_unmodeled_header(ctx, ref):
    # a PCOMP program announced to the decoder (taken from the "-m3" plan) in front of data the caller prepared;
    # and a header without components (level 2 block: length-prefixed chunks)
    plan = zqmod.plan_block("x0,2,12,0,7,21,1c0,0,511i2")
    arena, offs, lens = _pack(UNITS)
    for header in (plan["header"], zqmod.plan_block("x0,2,12,0,7,21,1")["header"]):
        out, ooff, olen = ctx.compress_segments(arena, offs, lens, header, pcomp=plan["pcomp"], comment="c")
        for i, u in enumerate(UNITS):
            want = ref.compress_segment(u, header=header, pcomp=plan["pcomp"], comment="c")
            assert out[int(ooff[i]): int(ooff[i]) + int(olen[i])].tobytes() == want, (i, len(u))


def test_cpp_facade_driver(ref, tmp_path):
    """Compressor / Decompresser / SHA1 / SHA256 / compressBlock through include/libzpaq_b200.h, as a C++ caller."""
    exe = tmp_path / "facade_driver"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests/cpp/facade_driver.cpp"),
                    "-o", str(exe), "-L" + os.path.join(ROOT, "zpaqfranz_b200"), "-lzqb200",
                    "-Wl,-rpath," + os.path.join(ROOT, "zpaqfranz_b200")], check=True)
    data = corpus.text_unit(11, 50000) + corpus.random_unit(12, 700) + corpus.repeats_unit(13, 9000)
    (tmp_path / "in.bin").write_bytes(data)
    # a stream the reference's Compressor wrote: a block of five segments (-m2 model), a one-segment block, a block of three
    multi = [[corpus.text_unit(21, 9000), b"", corpus.random_unit(22, 3000), corpus.text_unit(21, 4000), corpus.repeats_unit(23, 5000)],
             [corpus.text_unit(24, 20000)], [b"x" * 100, corpus.mixed_unit(25, 7000), b"tail"]]
    stream = (ref.compress_multi(multi[0], level=2, filename="m", comment="c") + ref.compress_multi(multi[1], level=1, filename="s", comment="20000")
              + b"junk" + ref.compress_multi(multi[2], level=1, filename="t", comment="", sha=False))
    (tmp_path / "multi.zpaq").write_bytes(stream)
    r = subprocess.run([str(exe), str(tmp_path / "in.bin"), str(tmp_path), str(tmp_path / "multi.zpaq")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    want = b"".join(bytes(x) for blk in multi for x in blk)
    assert (tmp_path / "out2.bin").read_bytes() == want and (tmp_path / "out3.bin").read_bytes() == want
    segrows = [l.split() for l in r.stdout.splitlines() if l.startswith("seg ")]
    assert [row[1] for row in segrows] == ["0.0", "0.1", "0.2", "0.3", "0.4", "1.0", "2.0", "2.1", "2.2"]
    assert [row[3] for row in segrows] == ["m0", "m1", "m2", "m3", "m4", "s0", "t0", "t1", "t2"]
    assert [int(row[5]) for row in segrows] == [len(x) for blk in multi for x in blk]
    assert all(row[7] == "1" and row[9] == "1" for row in segrows[:6]) and all(row[7] == "0" for row in segrows[6:])
    lines = dict(l.split(" ", 1) for l in r.stdout.strip().splitlines() if not l.startswith("seg "))
    # the 9.5 MB stream of the driver, regenerated here
    big, x, total = bytearray(), 12345, 0
    while total < 9500000:
        x = (x * 6364136223846793005 + 1442695040888963407) & (2 ** 64 - 1)
        k = 1 + (x >> 33) % 300000
        big += ((((np.arange(total, total + k, dtype=np.uint64) * np.uint64(2654435761)) & np.uint64(0xffffffff)) >> np.uint64(13)) & np.uint64(255)).astype(np.uint8).tobytes()
        total += k
    big.append(7)
    assert int(lines["bigsize"]) == len(big)
    assert lines["bigsha1"] == hashlib.sha1(big).hexdigest() and lines["bigsha256"] == hashlib.sha256(big).hexdigest()
    assert lines["sha1"] == hashlib.sha1(data).hexdigest()
    assert lines["sha256"] == hashlib.sha256(data).hexdigest()
    assert (tmp_path / "a.zpaq").read_bytes() == ref.compress_block(data, "2", "file_a", "jDC\x01")
    assert (tmp_path / "b.zpaq").read_bytes() == ref.compress_segment(data, level=2, filename="file_b", comment="a comment",
                                                                      sha1=hashlib.sha1(data).digest())
    assert (tmp_path / "out.bin").read_bytes() == data + data
    assert lines["blocks"] == "2"
    rows = [l for l in r.stdout.splitlines() if l.startswith("block ")]
    assert "name file_a" in rows[0] and "stored_sha1 1 match 1" in rows[0] and "size %d" % len(data) in rows[0]
    assert "name file_b" in rows[1] and "stored_sha1 1 match 1" in rows[1] and "size %d" % len(data) in rows[1]
    assert lines["error"] != "none"


def test_blocks_of_several_segments_through_the_c_abi(ctx, ref):
    """zq_decompress_blocks on blocks of several segments (reference-written): output = the segments concatenated, every
    stored SHA-1 verified on the device, zq_decompress_last_segments lists the pieces; a corrupted inner segment is caught."""
    segs = [[corpus.text_unit(31, 12000), corpus.mixed_unit(32, 5000), b"", corpus.text_unit(31, 800)],
            [corpus.repeats_unit(33, 30000)],
            [corpus.random_unit(34, 2000), corpus.text_unit(35, 2000)]]
    hdr3 = zqmod.assemble_config("comp 1 1 0 0 2 0 icm 12 1 isse 14 0 hcomp *d=a d++ hash *d=a halt end")["header"]
    blocks = [ref.compress_multi(segs[0], level=2, comment="c"), ref.compress_multi(segs[1], level=1, comment="c"),
              ref.compress_multi(segs[2], header=hdr3, comment="c")]
    arena, offs, lens = _pack(blocks)
    expect = np.array([sum(len(x) for x in b) for b in segs], dtype=np.uint32)
    dec, doff, dlen, used, tr = ctx.decompress_blocks(arena, offs, lens, expect_len=expect, details=True)
    for i, b in enumerate(segs):
        assert dec[int(doff[i]): int(doff[i]) + int(dlen[i])].tobytes() == b"".join(bytes(x) for x in b)
        assert blocks[i][int(used[i])] == 255
    table = ctx.last_segments()
    assert [(s[0], s[2] - s[1]) for s in table] == [(i, len(x)) for i, b in enumerate(segs) for x in b]
    for s in table:
        assert blocks[s[0]][s[3]] == 253
    bad = bytearray(blocks[0])
    t1 = [s for s in table if s[0] == 0][1][3]          # SHA-1 stored after the second segment
    bad[t1 + 5] ^= 1
    arena2, offs2, lens2 = _pack([bytes(bad)])
    with pytest.raises(Exception, match="SHA-1"):
        ctx.decompress_blocks(arena2, offs2, lens2, expect_len=expect[:1])
