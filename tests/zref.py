"""ctypes access to the *reference's own* zlib 1.2.8 (oracle/_ref/libz128.so, built
from /root/reference by oracle/build_ref.sh) and to the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY - nothing under antiz_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libz128.so")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "uncomp_ref")
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")

Z_OK, Z_STREAM_END, Z_NEED_DICT, Z_DATA_ERROR, Z_BUF_ERROR = 0, 1, 2, -3, -5
Z_SYNC_FLUSH, Z_FINISH = 2, 4


class ZStream(C.Structure):
    _fields_ = [
        ("next_in", C.c_void_p), ("avail_in", C.c_uint), ("total_in", C.c_ulong),
        ("next_out", C.c_void_p), ("avail_out", C.c_uint), ("total_out", C.c_ulong),
        ("msg", C.c_char_p), ("state", C.c_void_p),
        ("zalloc", C.c_void_p), ("zfree", C.c_void_p), ("opaque", C.c_void_p),
        ("data_type", C.c_int), ("adler", C.c_ulong), ("reserved", C.c_ulong),
    ]


_ref = None
_oracle = None


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        _ref = C.CDLL(REF_SO)
        _ref.zlibVersion.restype = C.c_char_p
        assert _ref.zlibVersion() == b"1.2.8", _ref.zlibVersion()
    return _ref


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        o = C.CDLL(ORACLE_SO)
        o.oracle_deflate.restype = C.c_longlong
        o.oracle_deflate.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32)]
        o.oracle_adler32.restype = C.c_uint32
        o.oracle_adler32.argtypes = [C.c_char_p, C.c_uint64]
        _oracle = o
    return _oracle


def ref_deflate(data: bytes, level: int, wbits: int, memlevel: int, strategy: int = 0) -> bytes:
    """deflateInit2(level, 8, wbits, memlevel, strategy) + deflate(Z_FINISH) with zlib 1.2.8 (strategy 1 = Z_FILTERED)."""
    z = ref()
    s = ZStream()
    rc = z.deflateInit2_(C.byref(s), level, 8, wbits, memlevel, strategy, b"1.2.8", C.sizeof(ZStream))
    assert rc == Z_OK, rc
    z.deflateBound.restype = C.c_ulong
    cap = z.deflateBound(C.byref(s), C.c_ulong(len(data))) + 64
    src = C.create_string_buffer(data, len(data) + 1)
    dst = C.create_string_buffer(cap)
    s.next_in = C.addressof(src); s.avail_in = len(data)
    s.next_out = C.addressof(dst); s.avail_out = cap
    rc = z.deflate(C.byref(s), Z_FINISH)
    assert rc == Z_STREAM_END, rc
    n = s.total_out
    z.deflateEnd(C.byref(s))
    return dst.raw[:n]


def ref_deflate_ex(data: bytes, level: int, wbits: int, memlevel: int, strategy: int = 0, flushes=(), tune=None) -> bytes:
    """zlib 1.2.8 driven the way other encoders drive it: `flushes` = [(input offset, flush mode)] (1 Z_PARTIAL_FLUSH, 2 Z_SYNC_FLUSH,
    3 Z_FULL_FLUSH, 5 Z_BLOCK) applied once the input up to that offset has been fed, then Z_FINISH; `tune` = deflateTune's
    (good_length, max_lazy, nice_length, max_chain).  Streams made like this inflate to `data` but are not what any plain
    deflateInit2 + deflate(Z_FINISH) produces: the search finds an imperfect winner or none."""
    z = ref()
    s = ZStream()
    rc = z.deflateInit2_(C.byref(s), level, 8, wbits, memlevel, strategy, b"1.2.8", C.sizeof(ZStream))
    assert rc == Z_OK, rc
    if tune:
        assert z.deflateTune(C.byref(s), *tune) == Z_OK
    cap = len(data) + len(data) // 4 + 1024 + 64 * (len(flushes) + 1)
    src = C.create_string_buffer(data, len(data) + 1)
    dst = C.create_string_buffer(cap)
    s.next_out = C.addressof(dst); s.avail_out = cap
    fed = 0
    for off, mode in list(flushes) + [(len(data), Z_FINISH)]:
        s.next_in = C.addressof(src) + fed; s.avail_in = off - fed; fed = off
        rc = z.deflate(C.byref(s), mode)
        assert rc in (Z_OK, Z_STREAM_END) and s.avail_in == 0, rc
    assert rc == Z_STREAM_END
    n = s.total_out
    z.deflateEnd(C.byref(s))
    return dst.raw[:n]


def oracle_deflate(data: bytes, level: int, wbits: int, memlevel: int, limit_out: int = 0):
    o = oracle()
    cap = len(data) + len(data) // 8 + 1024
    dst = C.create_string_buffer(cap)
    at = C.c_uint32(0)
    n = o.oracle_deflate(data, len(data), level, wbits, memlevel, dst, cap, limit_out, C.byref(at))
    assert 0 <= n <= cap, n
    if limit_out:
        return dst.raw[:n], at.value
    return dst.raw[:n]


def ref_inflate_scan(buf: bytes, start: int, outcap: int, segments=None):
    """Re-enact ZBuffSearcher's calls on one candidate (main.cpp:228-239): inflateInit,
    inflate(Z_SYNC_FLUSH) with avail_in = len(buf)-start, avail_out = outcap, then drain
    while avail_out == 0.  Returns (first_total_in, ret, total_in, total_out, avail_in).
    """
    z = ref()
    s = ZStream()
    rc = z.inflateInit_(C.byref(s), b"1.2.8", C.sizeof(ZStream))
    assert rc == Z_OK
    src = C.create_string_buffer(buf, len(buf) + 1)
    dst = C.create_string_buffer(outcap)
    s.next_in = C.addressof(src) + start; s.avail_in = len(buf) - start
    s.next_out = C.addressof(dst); s.avail_out = outcap
    ret = z.inflate(C.byref(s), Z_SYNC_FLUSH)
    first_in = s.total_in
    while s.avail_out == 0:
        s.next_out = C.addressof(dst); s.avail_out = outcap
        ret = z.inflate(C.byref(s), Z_SYNC_FLUSH)
    res = (first_in, ret, s.total_in, s.total_out, s.avail_in)
    z.inflateEnd(C.byref(s))
    return res


def ref_inflate(data: bytes, outlen: int):
    """doInflate (main.cpp:461-486): one-shot inflate(Z_FINISH)."""
    z = ref()
    s = ZStream()
    assert z.inflateInit_(C.byref(s), b"1.2.8", C.sizeof(ZStream)) == Z_OK
    src = C.create_string_buffer(data, len(data) + 1)
    dst = C.create_string_buffer(max(outlen, 1))
    s.next_in = C.addressof(src); s.avail_in = len(data)
    s.next_out = C.addressof(dst); s.avail_out = outlen
    ret = z.inflate(C.byref(s), Z_FINISH)
    n, ti = s.total_out, s.total_in
    z.inflateEnd(C.byref(s))
    return ret, dst.raw[:n], ti


class OIResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("err", C.c_int32), ("total_in", C.c_uint64),
                ("total_out", C.c_uint64), ("in_at_outcap", C.c_uint64), ("adler", C.c_uint32)]


class OISeg(C.Structure):
    _fields_ = [("p", C.c_void_p), ("n", C.c_uint64)]


OI_END, OI_NEED_INPUT, OI_DATA_ERROR, OI_NEED_DICT, OI_OUT_FULL = 0, 1, 2, 3, 4


def oracle_inflate(data: bytes, out_cap=None, first_out_cap=0):
    """Returns (OIResult, output bytes or None).  out_cap=None -> discard mode (32K ring)."""
    o = oracle()
    r = OIResult()
    src = C.create_string_buffer(data, len(data) + 1)
    if out_cap is None:
        o.oracle_inflate(src, C.c_uint64(len(data)), None, C.c_uint64(0), C.c_uint64(first_out_cap), C.byref(r))
        return r, None
    dst = C.create_string_buffer(max(out_cap, 1))
    o.oracle_inflate(src, C.c_uint64(len(data)), dst, C.c_uint64(out_cap), C.c_uint64(first_out_cap), C.byref(r))
    return r, dst.raw[:min(r.total_out, out_cap)]


def oracle_deflate_listmode(data: bytes, level: int, wbits: int, memlevel: int) -> bytes:
    """CPU design model of the GPU 'list mode' deflate (oracle/zdeflate.c, bottom)."""
    o = oracle()
    o.oracle_deflate_listmode.restype = C.c_longlong
    cap = len(data) + len(data) // 8 + 1024
    dst = C.create_string_buffer(cap)
    n = o.oracle_deflate_listmode(data, C.c_uint32(len(data)), level, wbits, memlevel, dst, C.c_uint64(cap))
    assert 0 <= n <= cap, n
    return dst.raw[:n]
