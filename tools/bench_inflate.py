#!/usr/bin/env python3
"""Secondary measurements (not the headline bench): K5 inflate and K1 CRC throughput on one GPU.

  single   : one reference-compressed raw stream (zlib level 6) of N MiB text decoded by one warp (config C3 shape)
  batch    : M independent 64 KiB entries decoded in one launch (config C4 extract shape)
  crc      : whole-buffer CRC-32 of a device buffer (config C1 shape, device resident)
Prints one JSON line per measurement; times are CUDA events on the launching stream, after warm-up.
"""
import ctypes as C
import json
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

import __graft_entry__ as ge
import textgen

pkg = ge._load_pkg()
lib = pkg.load()
pkg.check(lib.mz_cuda_init())
torch.cuda.set_device(0)
MiB = 1 << 20


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def dev(data):
    return torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()).cuda()


def single(n_mib):
    n = n_mib * MiB
    host = bytes(textgen.device(n, seed=3).cpu().numpy().tobytes())
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = co.compress(host) + co.flush()
    d_in = torch.zeros(len(comp) + 64, dtype=torch.uint8, device="cuda")
    d_in[:len(comp)] = dev(comp)
    d_out = torch.empty(n + 512, dtype=torch.uint8, device="cuda")
    job = pkg.InflateJob(d_in.data_ptr(), 0, len(comp), d_out.data_ptr(), 0, n, 1, 0)
    d_job = dev(bytes(job))
    d_st = torch.zeros(C.sizeof(pkg.InflateState), dtype=torch.uint8, device="cuda")

    def run():
        d_st.zero_()
        pkg.check(lib.mz_cuda_inflate_streams(d_job.data_ptr(), d_st.data_ptr(), 1, None))

    ms = timed(run, reps=2, warm=1)
    st = pkg.InflateState.from_buffer_copy(d_st.cpu().numpy().tobytes())
    ok = st.status == 1 and st.out_pos == n and zlib.crc32(d_out[:n].cpu().numpy().tobytes()) == zlib.crc32(host)
    print(json.dumps({"what": "inflate_single_stream", "out_mib": n_mib, "ms": round(ms, 2), "out_GBps": round(n / ms / 1e6, 3),
                      "blocks": st.blocks, "ok": bool(ok)}), flush=True)


def batch(m):
    ent = 65536
    src = textgen.device(m * ent, seed=4)
    b = pkg.DeflateBatch(m * ent)
    s = pkg._stream_ptr()
    d_off = torch.arange(m, dtype=torch.int64, device="cuda") * ent
    d_len = torch.full((m,), ent, dtype=torch.int32, device="cuda")
    d_flags = torch.ones(m, dtype=torch.uint8, device="cuda")
    pkg.check(lib.mz_cuda_deflate_chunks(src.data_ptr(), 0, 0, d_off.data_ptr(), d_len.data_ptr(), d_flags.data_ptr(), m, 0, 6,
                                         b.slots.data_ptr(), b.stride, b.out_len.data_ptr(), s))
    torch.cuda.synchronize()
    lens = b.out_len[:m].cpu().numpy()
    padded = torch.zeros(m * b.stride + 64, dtype=torch.uint8, device="cuda")
    padded[:m * b.stride] = b.slots[:m * b.stride]
    d_out = torch.empty(m * (ent + 512), dtype=torch.uint8, device="cuda")
    jobs = (pkg.InflateJob * m)()
    for i in range(m):
        jobs[i] = pkg.InflateJob(padded.data_ptr() + i * b.stride, 0, int(lens[i]), d_out.data_ptr() + i * (ent + 512), 0, ent, 1, 0)
    d_jobs = dev(bytes(jobs))
    d_st = torch.zeros(C.sizeof(pkg.InflateState) * m, dtype=torch.uint8, device="cuda")

    def run():
        d_st.zero_()
        pkg.check(lib.mz_cuda_inflate_streams(d_jobs.data_ptr(), d_st.data_ptr(), m, None))

    ms = timed(run, reps=3, warm=1)
    print(json.dumps({"what": "inflate_batch_64KiB_entries", "entries": m, "ms": round(ms, 2), "out_GBps": round(m * ent / ms / 1e6, 2),
                      "ratio": round(float(lens.sum()) / (m * ent), 4)}), flush=True)


def crc(n_mib):
    n = n_mib * MiB
    src = textgen.device(n, seed=5)
    nseg = (n + 65535) // 65536
    res = torch.empty(nseg, dtype=torch.int32, device="cuda")
    out2 = torch.empty(2, dtype=torch.int32, device="cuda")

    def run():
        pkg.check(lib.mz_cuda_crc32_segments(src.data_ptr(), n, 65536, None, None, nseg, res.data_ptr(), None, None))
        pkg.check(lib.mz_cuda_crc32_fold(res.data_ptr(), nseg, 65536, n, out2.data_ptr(), None))

    ms = timed(run, reps=10, warm=3)
    print(json.dumps({"what": "crc32_device", "mib": n_mib, "ms": round(ms, 4), "GBps": round(n / ms / 1e6, 1)}), flush=True)


def vtbl_long(n_mib, read_size=1 << 20, spec=True):
    """config C3 shape through the drop-in API: one foreign gzip member (zlib level 6) read with mz_stream_cuda_read from a
    host memory stream into a preallocated host buffer; the timed region is open + the read loop + close, host<->device
    copies included. Beside it: zlib's own inflate of the same member on one host core (what mz_strm_zlib runs)."""
    import time
    import cuharness
    tl = cuharness.TestLib()
    n = n_mib * MiB
    host = bytes(textgen.device(n, seed=9).cpu().numpy().tobytes())
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    comp = co.compress(host) + co.flush()
    os.environ["MZ_CUDA_SPEC"] = "1" if spec else "0"
    buf = C.create_string_buffer(n + 4096)
    C.memset(buf, 1, n + 4096)  # touch the pages outside the timed region
    best, got, total_in = None, -1, -1
    for _ in range(3):
        src, keep = tl.source(comp)
        s = lib.mz_stream_cuda_create()
        tl.lib.mzt_set_prop(s, pkg.MZ_STREAM_PROP_COMPRESS_WINDOW, 31)
        tl.lib.mzt_set_base(s, src)
        t0 = time.perf_counter()
        assert tl.lib.mzt_open(s, None, pkg.MZ_OPEN_MODE_READ) == 0
        got = tl.lib.mzt_read_all(s, buf, n + 1024, read_size)
        cerr = tl.lib.mzt_close(s)
        dt = time.perf_counter() - t0
        total_in = tl.get_prop(s, pkg.MZ_STREAM_PROP_TOTAL_IN)[1]
        tl.delete(s)
        tl.delete(src)
        assert cerr == 0
        best = dt if best is None else min(best, dt)
    ok = got == n and zlib.crc32(buf.raw[:n]) == zlib.crc32(host) and total_in == len(comp)
    t0 = time.perf_counter()
    zlib.decompress(comp, 31)
    cpu = time.perf_counter() - t0
    print(json.dumps({"what": "vtbl_read_one_gzip_member", "out_mib": n_mib, "spec": spec, "read_size": read_size, "s": round(best, 4),
                      "out_GBps": round(n / best / 1e9, 3), "zlib_1core_GBps": round(n / cpu / 1e9, 3), "ratio": round(len(comp) / n, 4),
                      "ok": bool(ok)}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "single":
        single(int(sys.argv[2]))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "long":
        vtbl_long(int(sys.argv[2]), spec=True)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "batch":
        batch(int(sys.argv[2]))
        sys.exit(0)
    crc(64)
    crc(4096)
    single(16)
    batch(8192)
    vtbl_long(32, spec=False)
    vtbl_long(16, spec=True)
    vtbl_long(256, spec=True)
    vtbl_long(1024, spec=True)
