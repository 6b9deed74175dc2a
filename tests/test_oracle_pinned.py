"""Pins the plain-C oracle (oracle/zq_oracle.c) against (a) the reference's own known-answer vectors
and (b) the reference compiled here (oracle/_ref/libzpaqref.so)."""
import pytest

from zpaqfranz_b200 import corpus

CASES = [
    corpus.text_unit(1, 65536), corpus.random_unit(2, 20000), corpus.repeats_unit(3, 65536), bytes(70000), b"", b"a",
    b"abcabcabcabcabc" * 10, corpus.text_unit(5, 200000), bytes([0, 0, 1, 0, 0, 0, 1, 0]) * 500,
]


def test_sha1_known_answers(oracle):
    # reference autotest KAT: SHA-1("ABCDE") (Z:77129-77160); FIPS 180 "abc"; empty input (Z:20670-20724)
    assert oracle.sha1(b"ABCDE").hex().upper() == "7BE07AAF460D593A323D0DB33DA05B64BFDCB3A5"
    assert oracle.sha1(b"abc").hex() == "a9993e364706816aba3e25717850c26c9cd0d89d"
    assert oracle.sha1(b"").hex() == "da39a3ee5e6b4b0d3255bfef95601890afd80709"


def test_sha1_and_suffix_array_vs_reference(oracle, ref):
    for d in CASES:
        assert oracle.sha1(d) == ref.sha1(d)
        if d:
            assert (oracle.suffix_array(d) == ref.divsufsort(d)).all()


@pytest.mark.parametrize("method", ["2", "1", "3", "1,10,0", "1,20,0", "2,10,0", "4,8,0", "0", "26,100,0"])
def test_lz_stream_and_block_vs_reference(zq, oracle, ref, method):
    for d in CASES:
        p = zq.plan_block(method, d)
        a = p["args"]
        if (a[1] & 3) == 3 or a[1] >= 4:
            continue
        s = d
        if a[1] & 3:
            s = oracle.lz_stream(d, a)
            assert s == ref.lz_stream(d, a)
        if p["header"][6] == 0:
            blk = oracle.block_unmodeled(p["header"], p["pcomp"], b"nm", ("%d cm" % len(d)).encode(), s, oracle.sha1(d))
            assert blk == ref.compress_block(d, method, "nm", "cm")
            assert ref.decompress(blk, len(d)) == d


def test_fragmenter_restatement_agrees(oracle, ref):
    d = corpus.text_unit(9, 400000) + corpus.random_unit(9, 300000) + bytes(600000)
    for frag in (6, 4, 0):
        a, b = oracle.fragment(d, frag), ref.fragment(d, frag)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
        assert int(a[0].sum()) == len(d)


def test_cm_tables_match_reference_checksums(oracle):
    # the reference's own KAT for the stretch/squash tables (Z:14949-14950)
    assert oracle.table_sums() == (3887533746, 2278286169)


@pytest.mark.parametrize("method", ["3", "36,200,1", "4", "46,200,1", "5", "3,100,0", "x0,0c0,0,255i2,13m8,24s",
                                     "x0,0c0,7i1c1004,0,1256i1s8,32,255", "x0,0c2,1100,255,0,128a24,1,1t16,20",
                                     "s0,0c0,0,255,255i3", "x0,0c8,0,255c0,0,255,255a16,2,2t8"])
def test_cm_restatement_vs_reference(zq, oracle, ref, method):
    cases = [corpus.text_unit(1, 12000), corpus.random_unit(2, 2000), corpus.repeats_unit(3, 9000), bytes(3000), b"", b"a",
             b"abcabcabcabcabc" * 10]
    if method == "5":
        cases = cases[3:] + [corpus.text_unit(1, 2500)]
    for d in cases:
        p = zq.plan_block(method, d)
        a = p["args"]
        if a[1] >= 4:
            continue
        s = oracle.lz_stream(d, a) if (a[1] & 3) else d
        blk = oracle.block_modeled(p["header"], p["pcomp"], b"nm", ("%d cm" % len(d)).encode(), s, oracle.sha1(d))
        assert blk == ref.compress_block(d, method, "nm", "cm"), (method, len(d))


def test_builtin_model_sources_assemble_to_the_reference_tables(zq, ref):
    # Compressor::startBlock(int level) takes its header from a byte table (Z:15992-16026); ours assembles the
    # published min.cfg / mid.cfg sources.  Same bytes, and the reference accepts ours as startBlock(hcomp).
    for level in (1, 2):
        header = zq.assemble_config(zq.model_config(level))["header"]
        blk = ref.compress_segment(b"some data, some data", level=level, tag=False)
        hs = blk[5] + 256 * blk[6] + 2
        assert blk[5:5 + hs] == header
        assert ref.compress_segment(b"some data, some data", header=header, tag=False) == blk
    assert zq.model_config(3) is None


HAND_MODELS = {
    "cm_o1": "comp 2 2 0 0 1 0 cm 18 255 hcomp *d=a a<<= 8 *d=a halt end",
    "all_types": ("comp 3 3 0 0 7 0 icm 12 1 isse 14 0 2 cm 16 32 3 match 14 16 4 mix2 8 1 2 20 255 5 sse 10 4 16 255 6 avg 4 5 96 "
                  "hcomp c++ *c=a b=c a=0 d= 0 hash *d=a d++ b-- hash *d=a d++ a=*c a<<= 9 *d=a d++ "
                  "b=c a=0 hash b-- hash b-- hash *d=a d++ a=*c *d=a d++ a=*c a>>= 3 *d=a halt end"),
    "branches": ("comp 2 4 0 0 3 0 cm 14 8 1 cm 16 20 2 mix 8 0 2 30 255 "
                 "hcomp *c=a c++ a== 32 if d= 0 *d=0 else d= 0 a+=*d a*= 73 *d=a endif "
                 "d= 1 a=*c a^= 255 *d=a d= 2 a=c a&= 7 *d=a halt end"),
}


@pytest.mark.parametrize("name", sorted(HAND_MODELS))
def test_cm_restatement_on_hand_written_models(zq, oracle, ref, name):
    # the Compressor class driven directly with a caller's model: comment verbatim, caller's checksum
    a = zq.assemble_config(HAND_MODELS[name])
    assert a["header"] == ref.compile(HAND_MODELS[name], [0] * 9)[0]
    for d in (corpus.text_unit(1, 9000), corpus.random_unit(2, 1500), bytes(2000), b"", b"q"):
        want = ref.compress_segment(d, header=a["header"], filename="f", comment="verbatim", sha1=oracle.sha1(d))
        assert oracle.block_modeled(a["header"], b"", b"f", b"verbatim", d, oracle.sha1(d)) == want, (name, len(d))
