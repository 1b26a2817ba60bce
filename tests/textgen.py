"""Synthetic enwik-style text for tests and the bench (tests/support/textgen.h): the same bytes from the host generator
(libmztest.so -- gcc only, used by the reference arm of bench.py, which must not map the GPU library) and from the device
generator (libmztextgen.so). TEST / BENCH SUPPORT, not product code."""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_host = None
_dev = None


def _hostlib():
    global _host
    if _host is None:
        _host = C.CDLL(os.path.join(ROOT, "tests/support/libmztest.so"))
        _host.mzt_textgen_host.restype = None
        _host.mzt_textgen_host.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int]
    return _host


def _devlib():
    global _dev
    if _dev is None:
        _dev = C.CDLL(os.path.join(ROOT, "tests/support/libmztextgen.so"))
        _dev.mzt_textgen_device.restype = C.c_int
        _dev.mzt_textgen_device.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
    return _dev


def host_into(addr, nbytes, seed=1, threads=0):
    _hostlib().mzt_textgen_host(addr, nbytes, seed, threads)


def host_buffer(nbytes, seed=1, threads=0):
    """ctypes uint8 array holding the text (no further copy)"""
    buf = (C.c_uint8 * max(nbytes, 1))()
    host_into(C.addressof(buf), nbytes, seed, threads)
    return buf


def host(nbytes, seed=1, threads=0):
    """bytes of text generated on the CPU"""
    buf = host_buffer(nbytes, seed, threads)  # keep it alive across the copy
    return C.string_at(C.addressof(buf), nbytes)


def device_into(ptr, nbytes, seed=1, stream=None):
    rc = _devlib().mzt_textgen_device(ptr, nbytes, seed, stream)
    if rc != 0:
        raise RuntimeError("mzt_textgen_device failed (%d)" % rc)


def device(nbytes, seed=1, out=None):
    """uint8 CUDA tensor of text generated on the current device"""
    import torch
    if out is None:
        out = torch.empty(nbytes, dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
    device_into(out.data_ptr(), nbytes, seed, torch.cuda.current_stream().cuda_stream)
    return out
