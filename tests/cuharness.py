"""Test harness: drives the product C-ABI (libmz_strm_cuda.so) the way the reference's tests drive mz_strm_zlib.

Streams are driven through the vtbl (tests/support/libmztest.so: mzt_* dispatch helpers + a 64-bit memory
base stream), so the same helpers work for the product stream and for the reference's own codec stream
(oracle/_ref/libmzref.so) chained over the same base -- test_stream_compress.cc:50-127 in Python.
"""
import ctypes as C
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

import refshim  # noqa: E402
from refshim import (MZ_OK, MZ_OPEN_MODE_READ, MZ_OPEN_MODE_WRITE, PROP_COMPRESS_LEVEL, PROP_COMPRESS_WINDOW,  # noqa: E402,F401
                     PROP_TOTAL_IN, PROP_TOTAL_IN_MAX, PROP_TOTAL_OUT)


def pkg():
    import __graft_entry__ as ge
    return ge._load_pkg()


class TestLib:
    """tests/support/libmztest.so"""

    def __init__(self):
        self.lib = L = C.CDLL(os.path.join(HERE, "support", "libmztest.so"))
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        L.mz_stream_mem64_create.restype = vp
        L.mz_stream_mem64_delete.argtypes = [C.POINTER(vp)]
        L.mz_stream_mem64_set_buffer.argtypes = [vp, vp, i64]
        L.mz_stream_mem64_set_sink.argtypes = [vp, vp, i64]
        L.mz_stream_mem64_set_discard.argtypes = [vp, i32]
        L.mz_stream_mem64_set_copy_threads.argtypes = [vp, i32]
        L.mz_stream_mem64_get_buffer.restype = i64
        L.mz_stream_mem64_get_buffer.argtypes = [vp, C.POINTER(vp)]
        for name, res, args in (("mzt_open", i32, [vp, C.c_char_p, i32]), ("mzt_is_open", i32, [vp]), ("mzt_read", i32, [vp, vp, i32]),
                                ("mzt_write", i32, [vp, vp, i32]), ("mzt_tell", i64, [vp]), ("mzt_seek", i32, [vp, i64, i32]),
                                ("mzt_close", i32, [vp]), ("mzt_error", i32, [vp]), ("mzt_get_prop", i32, [vp, i32, C.POINTER(i64)]),
                                ("mzt_set_prop", i32, [vp, i32, i64]), ("mzt_set_base", None, [vp, vp]),
                                ("mzt_delete", None, [C.POINTER(vp)]), ("mzt_write_all", i64, [vp, vp, i64, i32]),
                                ("mzt_read_all", i64, [vp, vp, i64, i32])):
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args

    def get_prop(self, s, prop):
        v = C.c_int64(-999)
        err = self.lib.mzt_get_prop(s, prop, C.byref(v))
        return err, v.value

    def delete(self, s):
        p = C.c_void_p(s)
        self.lib.mzt_delete(C.byref(p))
        return p.value

    def source(self, data):
        m = self.lib.mz_stream_mem64_create()
        keep = C.create_string_buffer(bytes(data), len(data)) if len(data) else C.create_string_buffer(1)
        self.lib.mz_stream_mem64_set_buffer(m, keep, len(data))
        return m, keep

    def sink(self):
        return self.lib.mz_stream_mem64_create()

    def sink_bytes(self, m):
        p = C.c_void_p()
        n = self.lib.mz_stream_mem64_get_buffer(m, C.byref(p))
        return C.string_at(p, n) if n else b""

    # ---- test_stream_compress.cc:64-86: write side -----------------------------------------------------------
    def compress(self, create_fn, data, level=6, window_bits=-15, write_size=16384, open_before_base=False):
        sink = self.sink()
        s = create_fn()
        assert s
        assert self.lib.mzt_set_prop(s, PROP_COMPRESS_LEVEL, level) == MZ_OK
        assert self.lib.mzt_set_prop(s, PROP_COMPRESS_WINDOW, window_bits) == MZ_OK
        if open_before_base:  # minigzip.c:92-93 order
            err = self.lib.mzt_open(s, None, MZ_OPEN_MODE_WRITE)
            self.lib.mzt_set_base(s, sink)
        else:
            self.lib.mzt_set_base(s, sink)
            err = self.lib.mzt_open(s, None, MZ_OPEN_MODE_WRITE)
        if err != MZ_OK:
            self.delete(s)
            self.delete(sink)
            return None, {"open": err}
        buf = C.create_string_buffer(bytes(data), len(data)) if len(data) else C.create_string_buffer(1)
        wrote = self.lib.mzt_write_all(s, buf, len(data), write_size)
        cerr = self.lib.mzt_close(s)
        info = {"open": err, "wrote": wrote, "close": cerr, "total_in": self.get_prop(s, PROP_TOTAL_IN)[1],
                "total_out": self.get_prop(s, PROP_TOTAL_OUT)[1], "sink_tell": self.lib.mzt_tell(sink),
                "is_open_after_close": self.lib.mzt_is_open(s), "error": self.lib.mzt_error(s)}
        out = self.sink_bytes(sink)
        self.delete(s)
        self.delete(sink)
        return out, info

    # ---- test_stream_compress.cc:88-117: read side --------------------------------------------------------------
    def decompress(self, create_fn, comp, out_cap, window_bits=-15, read_size=16384, total_in_max=0):
        src, keep = self.source(comp)
        s = create_fn()
        assert self.lib.mzt_set_prop(s, PROP_COMPRESS_WINDOW, window_bits) == MZ_OK
        if total_in_max:
            assert self.lib.mzt_set_prop(s, PROP_TOTAL_IN_MAX, total_in_max) == MZ_OK
        self.lib.mzt_set_base(s, src)
        err = self.lib.mzt_open(s, None, MZ_OPEN_MODE_READ)
        if err != MZ_OK:
            self.delete(s)
            self.delete(src)
            return None, {"open": err}
        buf = C.create_string_buffer(max(out_cap, 1) + 1024)
        got = self.lib.mzt_read_all(s, buf, out_cap + 1024, read_size)
        again = self.lib.mzt_read(s, buf, 1) if got >= 0 else self.lib.mzt_read(s, C.create_string_buffer(16), 16)
        info = {"read": got, "read_again": again, "error": self.lib.mzt_error(s)}
        info["close"] = self.lib.mzt_close(s)
        info["total_in"] = self.get_prop(s, PROP_TOTAL_IN)[1]
        info["total_out"] = self.get_prop(s, PROP_TOTAL_OUT)[1]
        info["base_tell"] = self.lib.mzt_tell(src)
        out = buf.raw[:got] if got >= 0 else b""
        self.delete(s)
        self.delete(src)
        return out, info
