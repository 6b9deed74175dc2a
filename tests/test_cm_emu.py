"""The context-mixing coder's DEVICE code (zpaqfranz_b200/csrc/zq_cm.cuh, zq_decode.cuh) compiled for the host
through the SIMT emulator (tests/emu/simt_emu.h) and checked bit for bit against the oracle, no GPU needed.

This pins the warp-level logic -- lane = component mapping, the (coder, context) warp pair and its ring, the
flat-switch ZPAQL interpreter, shared-memory row cache -- on CPU; the GPU tests then only have to show the
hardware runs the same program.  The emulator is test infrastructure: nothing on the product path uses it."""
import ctypes as C
import os
import subprocess

import pytest

import zpaqfranz_b200 as zq
from zpaqfranz_b200 import corpus

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "zpaqfranz_b200", "csrc")

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
    # four mixers (the third and fourth take the generic weight path), a mixer feeding a mixer, H beyond shared memory
    "four_mixers_big_h": ("comp 11 8 0 0 8 0 cm 12 16 1 icm 10 2 isse 10 1 3 mix 6 0 3 20 255 4 mix 0 0 4 24 0 "
                          "5 mix 8 1 4 16 15 6 mix 4 0 6 28 255 7 mix2 6 5 6 12 255 "
                          "hcomp *c=a c++ d= 0 *d=a d++ b=c b-- a=*b hash *d=a d++ hash *d=a d++ a=*c a<<= 4 *d=a d++ "
                          "a=c *d=a d++ a=*c a>>= 2 *d=a d++ *d=0 d++ a=*c *d=a a= 3 a<<= 8 d=a *d=c halt end"),
}


@pytest.fixture(scope="module")
def emu():
    out = os.path.join(EMU, "_build")
    os.makedirs(out, exist_ok=True)
    lib = os.path.join(out, "libcmemu.so")
    srcs = [os.path.join(EMU, "cm_emu.cpp"), os.path.join(CSRC, "zq_cm_host.cpp"), os.path.join(CSRC, "zq_config.cpp")]
    deps = srcs + [os.path.join(EMU, "simt_emu.h"), os.path.join(CSRC, "zq_cm.cuh"), os.path.join(CSRC, "zq_decode.cuh"),
                   os.path.join(CSRC, "zq_common.cuh")]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O1", "-std=c++17", "-w", "-I" + os.path.join(EMU, "shim"), "-I" + CSRC,
                        "-I" + os.path.join(ROOT, "include"), "-shared", "-fPIC", "-o", lib] + srcs, check=True)
    h = C.CDLL(lib)
    h.emu_cm_encode.restype = C.c_long
    h.emu_cm_decode.restype = C.c_long
    return h


def _coded_by_oracle(oracle, header, pcomp, stream):
    blk = oracle.block_modeled(header, pcomp, b"", b"", stream, None)
    pre = 13 + 5 + len(header) + 4          # tag, zPQ level 1, header, 01 "" 00 "" 00 00
    return blk[pre:-6]                       # ... coded ... 00 00 00 00 FE FF


def _emu_encode(emu, header, pcomp, stream, threads=64, prefetch=1, fast=1, vm=0, ctx=0):
    payload = (bytes([1, len(pcomp) & 255, len(pcomp) >> 8]) + pcomp) if pcomp else b"\0"
    cap = len(stream) * 2 + len(payload) * 2 + 4096
    out = (C.c_uint8 * cap)()
    n = emu.emu_cm_encode(header, len(header), payload, len(payload), stream, len(stream), out, cap, threads, prefetch | fast << 1 | vm << 2 | ctx << 4)
    assert n >= 0, n
    return bytes(out[:n])


def _emu_decode(emu, header, coded, cap, fast=1, vm=0):
    out = (C.c_uint8 * (cap + 16))()
    n = emu.emu_cm_decode(header, len(header), coded, len(coded), out, cap + 16, fast | vm << 2)
    assert n >= 0, n
    return bytes(out[:n])


# every built-in model with the default interpreter; the other two ZPAQL interpreters (selects / selects + predicated
# loads) on the models whose HCOMP/PCOMP programs branch, loop and jump the most
BUILTIN = [(m, 0) for m in ["36,200,1", "3", "4", "5", "46,200,1", "412,100,0"]] + [(m, vm) for vm in (1, 2) for m in ["36,200,1", "3"]]


@pytest.mark.parametrize("method,vm", BUILTIN)
def test_builtin_models_encode_and_decode(emu, oracle, method, vm):
    data = corpus.text_unit(3, 1200) if method != "412,100,0" else corpus.mixed_unit(5, 1200)
    plan = zq.plan_block(method, data)
    header, pcomp = bytes(plan["header"]), bytes(plan["pcomp"])
    stream = oracle.lz_stream(data, plan["args"]) if (plan["args"][1] & 3) else data
    want = _coded_by_oracle(oracle, header, pcomp, stream)
    assert _emu_encode(emu, header, pcomp, stream, vm=vm) == want
    assert _emu_encode(emu, header, pcomp, stream, fast=0, vm=vm) == want     # chain models: the generic lane engine too
    # decoder: coded data + end-of-stream zeros -> post-processed original (PCOMP runs on the same interpreter)
    assert _emu_decode(emu, header, want + b"\0\0\0\0", len(data), vm=vm) == data
    assert _emu_decode(emu, header, want + b"\0\0\0\0", len(data), fast=0, vm=vm) == data


CUSTOM = [(n, 0) for n in sorted(CONFIGS)] + [(n, vm) for vm in (1, 2) for n in ("branches", "icm_chain_mix2_sse")]


@pytest.mark.parametrize("name,vm", CUSTOM)
def test_custom_models_encode_and_decode(emu, oracle, name, vm):
    header = bytes(zq.assemble_config(CONFIGS[name])["header"])
    for data in (b"", b"x", b"abracadabra" * 30, corpus.text_unit(9, 700), corpus.random_unit(5, 300)):
        want = _coded_by_oracle(oracle, header, b"", data)
        assert _emu_encode(emu, header, b"", data, vm=vm) == want, (name, len(data))
        assert _emu_decode(emu, header, want + b"\0\0\0\0", len(data), vm=vm) == data, (name, len(data))


def test_pairs_share_a_cta_and_prefetch_is_transparent(emu, oracle):
    data = corpus.text_unit(11, 600)
    plan = zq.plan_block("4", data)
    header = bytes(plan["header"])
    want = _coded_by_oracle(oracle, header, b"", data)
    assert _emu_encode(emu, header, b"", data, threads=192, prefetch=0) == want   # idle pairs leave through the queue
    assert _emu_encode(emu, header, b"", data, threads=64, prefetch=1) == want
    # contexts handed over precomputed (what the translated context program will provide): coder output unchanged
    assert _emu_encode(emu, header, b"", data, ctx=1) == want
    plan3 = zq.plan_block("3", data)
    s3 = oracle.lz_stream(data, plan3["args"])
    h3, p3 = bytes(plan3["header"]), bytes(plan3["pcomp"])
    assert _emu_encode(emu, h3, p3, s3, ctx=1) == _coded_by_oracle(oracle, h3, p3, s3)


@pytest.mark.parametrize("kind", ["level1", "level2", "icm_chain", "stored"])
def test_blocks_of_several_segments_decode(emu, oracle, ref, kind):
    """A block of four segments written by the reference's Compressor class: the device decoder carries the model, the
    coder's range and the post-processor from one segment into the next (Decompresser::decompress, Z:15481-15508),
    reads the trailers and segment headers in between itself and reports where every segment ends."""
    if ref is None:
        pytest.skip("oracle/_ref not built")
    segs = [bytes(corpus.text_unit(1, 3000)), b"", bytes(corpus.mixed_unit(2, 2500)), bytes(corpus.text_unit(1, 1200))]
    if kind.startswith("level"):
        header = zq.assemble_config(zq.model_config(int(kind[-1])))["header"]
    elif kind == "icm_chain":
        header = zq.assemble_config("comp 1 1 0 0 2 0 icm 12 1 isse 14 0 hcomp *d=a d++ hash *d=a halt end")["header"]    # the one-lane fast path
    else:
        header = zq.assemble_config("comp 0 0 0 0 0 hcomp end")["header"]
    for sha in (True, False):
        blk = ref.compress_multi(segs, header=header, comment="c", sha=sha)
        assert ref.decompress(blk, 1 << 20) == b"".join(segs)
        pre = 13 + 5 + len(header) + len(b"\x01seg0\0c\0\0")
        coded = blk[pre:]
        for fast in (1, 0):
            got = _emu_decode(emu, header, coded, 20000, fast=fast)
            assert got == b"".join(segs), (kind, sha, fast)
            buf = (C.c_uint32 * (3 * 16))()
            used = C.c_uint32(0)
            k = emu.emu_cm_decode_segments(buf, 16, C.byref(used))
            assert k == 3
            ends = [buf[3 * i + 1] for i in range(3)]
            assert ends == [len(segs[0]), len(segs[0]), len(segs[0]) + len(segs[2])]
            for i in range(3):
                assert coded[buf[3 * i + 2]] == (253 if sha else 254)
            assert coded[used.value] == (253 if sha else 254) and coded[-1] == 255


def test_models_wider_than_a_warp(emu, oracle):
    """33 components: method 5 on text in which two periods are detected (fixed-width records).  One lane evaluates the
    components in index order (zq_cm_wide.cuh) -- coded bytes equal the oracle's, the decoder restores the input."""
    data = b"".join(b"%06d,abcde,%08d,xyzxyz\n" % (i, i * 7) for i in range(60))
    plan = zq.plan_block("56,200,1", data)
    header, pcomp = bytes(plan["header"]), bytes(plan["pcomp"])
    assert header[6] == 33
    stream = oracle.lz_stream(data, plan["args"]) if (plan["args"][1] & 3) else data
    want = _coded_by_oracle(oracle, header, pcomp, stream)
    assert _emu_encode(emu, header, pcomp, stream) == want
    assert _emu_decode(emu, header, want + b"\0\0\0\0", len(data)) == data
    # a hand-written model of 40 components with every type beyond lane 31
    comps = " ".join("%d cm 10 16" % i if i % 3 == 0 else "%d icm 8" % i if i % 3 == 1 else "%d isse 8 %d" % (i, i - 1) for i in range(32))
    cfg = ("comp 3 3 0 0 40 " + comps + " 32 match 10 12 33 avg 0 1 100 34 mix2 6 32 33 20 255 35 mix 8 0 35 20 255 "
           "36 sse 8 35 8 255 37 const 150 38 mix 0 30 8 24 0 39 mix2 0 36 38 16 0 "
           "hcomp c++ *c=a b=c a=0 d= 0 hash *d=a d++ b-- hash *d=a d++ a=*c a<<= 9 *d=a d++ hash *d=a d++ "
           "a=*c *d=a d++ a=c *d=a d++ hash *d=a d++ a+=*c *d=a halt end")
    h40 = bytes(zq.assemble_config(cfg)["header"])
    assert h40[6] == 40
    for d in (b"", b"abracadabra" * 20, corpus.text_unit(4, 500)):
        want = _coded_by_oracle(oracle, h40, b"", d)
        assert _emu_encode(emu, h40, b"", d) == want, len(d)
        assert _emu_decode(emu, h40, want + b"\0\0\0\0", len(d)) == d, len(d)
