"""ctypes views of the two CHECKERS under oracle/_ref/ (test infrastructure; never used by the product)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        lib.zqo_lz_stream.restype = C.c_longlong
        lib.zqo_block_unmodeled.restype = C.c_longlong
        lib.zqo_fragment.restype = C.c_longlong
        lib.zqo_block_modeled.restype = C.c_longlong

    def sha1(self, data):
        out = C.create_string_buffer(20)
        self.lib.zqo_sha1(bytes(data), C.c_uint64(len(data)), out)
        return out.raw

    def suffix_array(self, data):
        data = bytes(data)
        sa = np.zeros(max(len(data), 1), dtype=np.uint32)
        self.lib.zqo_suffix_array(data, sa.ctypes.data_as(C.c_void_p), C.c_uint32(len(data)))
        return sa[: len(data)]

    def lz_stream(self, data, args, sa=None, tokens=False):
        data = bytes(data)
        a = (C.c_int * 9)(*args)
        cap = len(data) + len(data) // 16 + 1024
        out = C.create_string_buffer(cap)
        tok = np.zeros(3 * (len(data) + 1), dtype=np.uint32)
        ntok = C.c_uint64(0)
        sap = sa.ctypes.data_as(C.c_void_p) if sa is not None else None
        r = self.lib.zqo_lz_stream(data, C.c_uint32(len(data)), a, sap, out, C.c_uint64(cap),
                                   tok.ctypes.data_as(C.c_void_p), C.c_uint64(tok.size), C.byref(ntok))
        if r < 0:
            raise RuntimeError("zqo_lz_stream failed %d" % r)
        if tokens:
            return out.raw[:r], tok[: 3 * ntok.value].reshape(-1, 3)
        return out.raw[:r]

    def e8e9(self, data):
        b = C.create_string_buffer(bytes(data), len(data))
        self.lib.zqo_e8e9(b, C.c_int(len(data)))
        return b.raw

    def block_unmodeled(self, header, pcomp, filename, comment_full, stream, sha1):
        cap = len(stream) + len(stream) // 1000 + len(header) + len(pcomp) + 1024
        out = C.create_string_buffer(cap)
        r = self.lib.zqo_block_unmodeled(bytes(header), C.c_uint32(len(header)), bytes(pcomp), C.c_uint32(len(pcomp)),
                                         filename, comment_full, bytes(stream), C.c_uint64(len(stream)), sha1, out,
                                         C.c_uint64(cap))
        if r < 0:
            raise RuntimeError("zqo_block_unmodeled failed %d" % r)
        return out.raw[:r]

    def block_modeled(self, header, pcomp, filename, comment_full, stream, sha1):
        cap = len(stream) + len(stream) // 8 + len(header) + len(pcomp) + 4096
        out = C.create_string_buffer(cap)
        r = self.lib.zqo_block_modeled(bytes(header), C.c_uint32(len(header)), bytes(pcomp), C.c_uint32(len(pcomp)),
                                       filename, comment_full, bytes(stream), C.c_uint64(len(stream)), sha1, out,
                                       C.c_uint64(cap))
        if r < 0:
            raise RuntimeError("zqo_block_modeled failed %d" % r)
        return out.raw[:r]

    def table_sums(self):
        a, b = C.c_uint32(0), C.c_uint32(0)
        self.lib.zqo_table_sums(C.byref(a), C.byref(b))
        return a.value, b.value

    def fragment(self, data, fragment=6, blocksize=(1 << 26) - 4096):
        data = bytes(data)
        cap = len(data) // 64 + 16
        fl = np.zeros(cap, dtype=np.uint32)
        fh = np.zeros(cap, dtype=np.uint32)
        k = self.lib.zqo_fragment(data, C.c_uint64(len(data)), C.c_int(fragment), C.c_uint32(blocksize),
                                  fl.ctypes.data_as(C.c_void_p), fh.ctypes.data_as(C.c_void_p), C.c_uint64(cap))
        return fl[:k].copy(), fh[:k].copy()


class Ref:
    """The unmodified reference (libzpaq inside zpaqfranz.cpp) behind oracle/ref_shim.cpp."""

    def __init__(self, lib):
        self.lib = lib
        for f in ("zref_compress_block", "zref_decompress", "zref_lz_stream", "zref_fragment", "zref_compress_units_mt", "zref_compress_segment"):
            getattr(lib, f).restype = C.c_longlong
        lib.zref_last_error.restype = C.c_char_p

    def compress_block(self, data, method, filename=None, comment=None, dosha1=True):
        data = bytes(data)
        cap = len(data) + len(data) // 8 + 100000
        out = C.create_string_buffer(cap)
        fn = filename.encode() if isinstance(filename, str) else filename
        cm = comment.encode() if isinstance(comment, str) else comment
        r = self.lib.zref_compress_block(data, C.c_uint(len(data)), method.encode(), fn, cm, C.c_int(1 if dosha1 else 0),
                                         out, C.c_ulonglong(cap))
        if r < 0:
            raise RuntimeError("reference compressBlock failed: %s" % self.lib.zref_last_error().decode(errors="replace"))
        return out.raw[:r]

    def compress_segment(self, data, header=None, level=0, pcomp=b"", filename=None, comment=None, sha1=None, tag=True):
        """libzpaq::Compressor driven directly: writeTag/startBlock(level | header)/startSegment/postProcess/compress/
        endSegment(sha1)/endBlock."""
        data = bytes(data)
        cap = len(data) + len(data) // 8 + 200000
        out = C.create_string_buffer(cap)
        fn = filename.encode() if isinstance(filename, str) else filename
        cm = comment.encode() if isinstance(comment, str) else comment
        r = self.lib.zref_compress_segment(C.c_int(level), bytes(header) if header else None, bytes(pcomp) if pcomp else None,
                                           C.c_int(len(pcomp)), data, C.c_uint(len(data)), fn, cm,
                                           bytes(sha1) if sha1 is not None else None, C.c_int(1 if tag else 0), out, C.c_ulonglong(cap))
        if r < 0:
            raise RuntimeError("reference Compressor failed: %s" % self.lib.zref_last_error().decode(errors="replace"))
        return out.raw[:r]

    def compress_multi(self, segments, header=None, level=0, pcomp=b"", filename="seg", comment="", sha=True):
        """One block with several segments through the reference's Compressor class."""
        import numpy as _np
        data = b"".join(bytes(x) for x in segments)
        lens = _np.array([len(x) for x in segments], dtype=_np.uint32)
        offs = (_np.cumsum(lens, dtype=_np.uint64) - lens).astype(_np.uint64)
        cap = len(data) + len(data) // 8 + 200000 + 400 * len(segments)
        out = C.create_string_buffer(cap)
        self.lib.zref_compress_multi.restype = C.c_longlong
        r = self.lib.zref_compress_multi(C.c_int(level), bytes(header) if header else None, bytes(pcomp) if pcomp else None,
                                         C.c_int(len(pcomp)), C.c_int(len(segments)), data + b"\0", offs.ctypes.data_as(C.c_void_p),
                                         lens.ctypes.data_as(C.c_void_p), filename.encode(), comment.encode(), C.c_int(1 if sha else 0),
                                         out, C.c_ulonglong(cap))
        if r < 0:
            raise RuntimeError("reference Compressor failed: %s" % self.lib.zref_last_error().decode(errors="replace"))
        return out.raw[:r]

    def decompress(self, blob, cap):
        out = C.create_string_buffer(cap + 16)
        r = self.lib.zref_decompress(bytes(blob), C.c_ulonglong(len(blob)), out, C.c_ulonglong(cap + 16))
        if r < 0:
            raise RuntimeError("reference decompress failed: %s" % self.lib.zref_last_error().decode(errors="replace"))
        return out.raw[:r]

    def _digest(self, fn, n, data):
        out = C.create_string_buffer(n)
        fn(bytes(data), C.c_ulonglong(len(data)), out)
        return out.raw

    def sha1(self, d):
        return self._digest(self.lib.zref_sha1, 20, d)

    def sha256(self, d):
        return self._digest(self.lib.zref_sha256, 32, d)

    def xxh3_128(self, d):
        return self._digest(self.lib.zref_xxh3_128, 16, d)

    def blake3(self, d):
        return self._digest(self.lib.zref_blake3, 32, d)

    def md5(self, d):
        return self._digest(self.lib.zref_md5, 16, d)

    def sha3_256(self, d):
        return self._digest(self.lib.zref_sha3_256, 32, d)

    def xxh64(self, d):
        return self._digest(self.lib.zref_xxh64, 8, d)

    def crc32(self, d):
        return self._digest(self.lib.zref_crc32, 4, d)

    def divsufsort(self, data):
        data = bytes(data)
        sa = np.zeros(max(len(data), 1), dtype=np.int32)
        self.lib.zref_divsufsort(data, sa.ctypes.data_as(C.c_void_p), C.c_int(len(data)))
        return sa[: len(data)].astype(np.uint32)

    def make_config(self, method):
        args = (C.c_int * 9)()
        out = C.create_string_buffer(1 << 16)
        r = self.lib.zref_make_config(method.encode(), args, out, 1 << 16)
        if r < 0:
            raise RuntimeError(self.lib.zref_last_error().decode(errors="replace"))
        return out.value.decode(), list(args)

    def compile(self, config, args):
        a = (C.c_int * 9)(*args)
        hdr = C.create_string_buffer(70000)
        pc = C.create_string_buffer(70000)
        hl, pl = C.c_int(0), C.c_int(0)
        r = self.lib.zref_compile(config.encode(), a, hdr, C.byref(hl), pc, C.byref(pl))
        if r < 0:
            raise RuntimeError(self.lib.zref_last_error().decode(errors="replace"))
        return hdr.raw[: hl.value], pc.raw[: pl.value]

    def lz_stream(self, data, args):
        data = bytes(data)
        a = (C.c_int * 9)(*args)
        cap = len(data) + len(data) // 16 + 1024
        out = C.create_string_buffer(cap)
        r = self.lib.zref_lz_stream(data, C.c_uint(len(data)), a, out, C.c_ulonglong(cap))
        if r < 0:
            raise RuntimeError(self.lib.zref_last_error().decode(errors="replace"))
        return out.raw[:r]

    def fragment(self, data, fragment=6):
        data = bytes(data)
        cap = len(data) // 64 + 16
        fl = np.zeros(cap, dtype=np.uint32)
        fh = np.zeros(cap, dtype=np.uint32)
        k = self.lib.zref_fragment(data, C.c_ulonglong(len(data)), C.c_int(fragment), fl.ctypes.data_as(C.c_void_p),
                                   fh.ctypes.data_as(C.c_void_p), C.c_ulonglong(cap))
        return fl[:k].copy(), fh[:k].copy()


def load_oracle():
    return Oracle(C.CDLL(os.path.join(REF_DIR, "libzqoracle.so")))


def load_ref():
    p = os.path.join(REF_DIR, "libzpaqref.so")
    if not os.path.exists(p):
        return None
    return Ref(C.CDLL(p))
