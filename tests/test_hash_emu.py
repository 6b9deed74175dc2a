"""The digest kernels' DEVICE code (SHA-1, SHA-256, XXH3-128, BLAKE3: zq_sha1.cuh, zq_hashes.cuh) on the host through the
SIMT emulator, against hashlib and the reference's own XXH3 / BLAKE3 (oracle/_ref) plus the reference's "ABCDE"
known answers (Z:77129-77160).  Test infrastructure only."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

from zpaqfranz_b200 import corpus

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "zpaqfranz_b200", "csrc")
SIZES = [0, 1, 3, 5, 55, 56, 63, 64, 65, 119, 128, 129, 240, 241, 1023, 1024, 1025, 2048, 3000, 4097, 9000, 70000]


@pytest.fixture(scope="module")
def emu():
    out = os.path.join(EMU, "_build")
    os.makedirs(out, exist_ok=True)
    lib = os.path.join(out, "libhashemu.so")
    deps = [os.path.join(EMU, "hash_emu.cpp"), os.path.join(EMU, "simt_emu.h"), os.path.join(CSRC, "zq_sha1.cuh"),
            os.path.join(CSRC, "zq_hashes.cuh"), os.path.join(CSRC, "zq_hashes2.cuh"), os.path.join(CSRC, "zq_hashes3.cuh"),
            os.path.join(CSRC, "zq_common.cuh")]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O1", "-std=c++17", "-w", "-I" + os.path.join(EMU, "shim"), "-I" + CSRC,
                        "-I" + os.path.join(ROOT, "include"), "-shared", "-fPIC", "-o", lib, deps[0]], check=True)
    return C.CDLL(lib)


def _run(emu, kind, width, bufs, shift=0):
    blob = b"\0" * shift + b"".join(bufs) + b"\0" * 64          # shift: unaligned starts
    lens = np.array([len(b) for b in bufs], dtype=np.uint64)
    offs = (np.concatenate([[0], np.cumsum(lens)[:-1]]) + shift).astype(np.uint64)
    out = (C.c_uint8 * (width * len(bufs)))()
    emu.emu_hash(kind, blob, offs.ctypes.data_as(C.c_void_p), lens.ctypes.data_as(C.c_void_p), len(bufs), out)
    raw = bytes(out)
    return [raw[i * width:(i + 1) * width] for i in range(len(bufs))]


BUFS = [bytes(corpus.random_unit(s + 1, s)) if s else b"" for s in SIZES] + [b"ABCDE", bytes(5000)]


@pytest.mark.parametrize("shift", [0, 1, 7])
def test_sha1_sha256(emu, shift):
    assert _run(emu, 0, 20, BUFS, shift) == [hashlib.sha1(b).digest() for b in BUFS]
    assert _run(emu, 1, 32, BUFS, shift) == [hashlib.sha256(b).digest() for b in BUFS]


def test_xxh3_blake3_against_reference(emu, ref):
    if ref is None:
        pytest.skip("oracle/_ref not built")
    assert _run(emu, 2, 16, BUFS, 3) == [ref.xxh3_128(b) for b in BUFS]
    assert _run(emu, 3, 32, BUFS, 5) == [ref.blake3(b) for b in BUFS]


def test_reference_known_answers(emu):
    abcde = [b"ABCDE"]
    assert _run(emu, 0, 20, abcde)[0].hex().upper() == "7BE07AAF460D593A323D0DB33DA05B64BFDCB3A5"
    assert _run(emu, 2, 16, abcde)[0].hex().upper() == "1C8288B6013152D97B4A5D7E6C7893D4"
    assert _run(emu, 3, 32, abcde)[0].hex().upper().startswith("61274278")


@pytest.mark.parametrize("shift", [0, 1, 2, 3])
def test_crc32_xxh64(emu, ref, shift):
    import zlib
    bufs = BUFS + [bytes(corpus.random_unit(77, 4096)), bytes(corpus.random_unit(78, 8192)), bytes(corpus.random_unit(79, 12289))]
    assert _run(emu, 4, 4, bufs, shift) == [zlib.crc32(b).to_bytes(4, "little") for b in bufs]
    if ref is not None:
        assert _run(emu, 4, 4, bufs, shift) == [ref.crc32(b) for b in bufs]
        assert _run(emu, 5, 8, bufs, shift) == [ref.xxh64(b) for b in bufs]


@pytest.mark.parametrize("shift", [0, 1, 5])
def test_md5_sha3(emu, ref, shift):
    bufs = BUFS + [bytes(corpus.random_unit(80, k)) for k in (55, 56, 63, 64, 65, 119, 120, 135, 136, 137, 272, 4097)]
    assert _run(emu, 6, 16, bufs, shift) == [hashlib.md5(b).digest() for b in bufs]
    assert _run(emu, 7, 32, bufs, shift) == [hashlib.sha3_256(b).digest() for b in bufs]
    if ref is not None:       # the reference's own classes (MD5 Z:21432, SHA3 Z:21189) agree
        assert [ref.md5(b) for b in bufs[:12]] == [hashlib.md5(b).digest() for b in bufs[:12]]
        assert [ref.sha3_256(b) for b in bufs[:12]] == [hashlib.sha3_256(b).digest() for b in bufs[:12]]


def test_fragment_index(emu):
    """k_dedup_insert / k_dedup_lookup: first[i] = the earliest fragment with fragment i's digest (HTIndex, Z:71567-71604);
    colliding table slots (digests sharing their first 8 bytes) included."""
    rng = np.random.default_rng(5)
    pool = rng.integers(0, 256, size=(300, 20), dtype=np.uint8)
    pool[100:200, :8] = pool[0, :8]                          # same home slot, different digests
    pick = rng.integers(0, 300, size=5000)
    dg = np.ascontiguousarray(pool[pick])
    first = np.zeros(5000, dtype=np.uint32)
    emu.emu_dedup_first(dg.ctypes.data_as(C.c_void_p), 5000, first.ctypes.data_as(C.c_void_p))
    seen, want = {}, []
    for i, k in enumerate(pick):
        want.append(seen.setdefault(dg[i].tobytes(), i))
    assert first.tolist() == want


@pytest.mark.parametrize("words,name", [(5, "sha1"), (8, "sha256")])
def test_streamed_sha_continues_from_a_chaining_value(emu, words, name):
    """k_sha1_continue / k_sha256_continue (the streaming SHA1 / SHA256 classes of the boundary): whole blocks fed in
    uneven pieces from the initial chaining value, padded by the caller, equal hashlib."""
    init = {5: [0x67452301, 0xEFCDAB89, 0x98BADCFE, 0x10325476, 0xC3D2E1F0],
            8: [0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19]}[words]
    for size in (0, 1, 55, 56, 64, 1000, 4096 + 17):
        msg = bytes(corpus.random_unit(size + 3, size)) if size else b""
        padded = msg + b"\x80" + b"\0" * ((55 - len(msg)) % 64) + (8 * len(msg)).to_bytes(8, "big")
        st = (C.c_uint32 * words)(*init)
        pos, step = 0, 1
        while pos < len(padded):                      # 1, 2, 3 ... blocks at a time
            k = min(step, (len(padded) - pos) // 64)
            emu.emu_sha_continue(words, st, padded[pos: pos + 64 * k], k)
            pos += 64 * k
            step += 1
        assert b"".join(int(x).to_bytes(4, "big") for x in st) == getattr(hashlib, name)(msg).digest(), size
