"""BASELINE.json configs[1..3] at the stated sizes (VERDICT r1 task 1d): what the smaller parity tests cannot reach --
output offsets beyond 4 GiB, ISIZE wrapping to 0, a 256 MiB level-6 stream through the reference's own CLI, a
100 000-entry archive checked entry by entry by the reference's reader."""
import ctypes as C
import json
import os
import subprocess
import sys
import zlib

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFDIR = os.path.join(ROOT, "oracle", "_ref")
GiB = 1 << 30

pytestmark = pytest.mark.gpu


def _bin(name):
    p = os.path.join(REFDIR, name)
    if not os.path.exists(p):
        pytest.skip(f"{name} not built (needs /root/reference at build time)")
    return p


def _shm(tmp_path):
    return "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else str(tmp_path)


def test_c3_four_gib_member_through_the_vtbl(built):
    """One gzip member of exactly 4 GiB of text, zlib level 6 blocks (built in parallel pieces by zlib itself, see
    bench.make_gzip_member): read with mz_stream_cuda_read in 1 MiB calls. ISIZE in the trailer is 0 (4 GiB mod 2^32), output
    positions pass 2^32, TOTAL_IN must equal the member's length and the CRC of the 4 GiB must match."""
    import torch
    import bench
    import cuharness
    import textgen
    p = cuharness.pkg()
    lib = p.load()
    assert lib.mz_cuda_init() == 0
    tl = cuharness.TestLib()
    n = 4 * GiB
    text = textgen.host_buffer(n, seed=77)
    member, crc = bench.make_gzip_member(text, n, 6)
    assert int.from_bytes(member[-4:], "little") == 0 and int.from_bytes(member[-8:-4], "little") == crc
    del text
    clen = len(member)
    msrc = C.create_string_buffer(member, clen)
    del member
    out = torch.empty(n + 4096, dtype=torch.uint8)
    srcs = tl.lib.mz_stream_mem64_create()
    tl.lib.mz_stream_mem64_set_buffer(srcs, msrc, clen)
    z = lib.mz_stream_cuda_create()
    tl.lib.mzt_set_prop(z, p.MZ_STREAM_PROP_COMPRESS_WINDOW, 31)
    tl.lib.mzt_set_base(z, srcs)
    assert tl.lib.mzt_open(z, None, p.MZ_OPEN_MODE_READ) == 0
    got = tl.lib.mzt_read_all(z, out.data_ptr(), n + 1024, 1 << 20)
    assert got == n
    assert tl.get_prop(z, p.MZ_STREAM_PROP_TOTAL_IN)[1] == clen     # final as soon as the last byte is out (mz_zip.c:2100-2112)
    assert tl.get_prop(z, p.MZ_STREAM_PROP_TOTAL_OUT)[1] == n
    assert tl.lib.mzt_close(z) == 0
    tl.delete(z)
    tl.delete(srcs)
    assert zlib.crc32(memoryview(out.numpy())[:n]) == crc


def test_c2_256_mib_level6_minigzip_cuda_read_back_by_the_reference(built, tmp_path):
    """configs[1]: the reference's own minigzip program linked against the cuda stream compresses 256 MiB of text at level 6;
    the unmodified reference build decompresses it; the bytes must be the input's and the ratio in zlib's neighbourhood."""
    import textgen
    exe, ref = _bin("minigzip_cuda"), _bin("minigzip_ref")
    d = os.path.join(_shm(tmp_path), "mz_c2_%d" % os.getpid())
    os.makedirs(d, exist_ok=True)
    try:
        n = 256 << 20
        data = textgen.host(n, seed=55)
        with open(os.path.join(d, "doc.txt"), "wb") as f:
            f.write(data)
        r = subprocess.run([exe, "-6", "doc.txt"], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
        assert r.returncode == 0 and b"Operation completed successfully" in r.stdout, (r.stdout[-400:], r.stderr[-400:])
        gz = os.path.join(d, "doc.txt.gz")
        size = os.path.getsize(gz)
        with open(gz, "rb") as f:
            f.seek(-8, 2)
            tr = f.read(8)
        assert int.from_bytes(tr[:4], "little") == zlib.crc32(data) and int.from_bytes(tr[4:], "little") == n
        assert size < 0.50 * n, size / n
        r = subprocess.run([ref, "-x", "-d", "x", "doc.txt.gz"], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
        assert r.returncode == 0, (r.stdout[-400:], r.stderr[-400:])
        with open(os.path.join(d, "x", "doc.txt"), "rb") as f:
            back = f.read()
        assert back == data
    finally:
        subprocess.run(["rm", "-rf", d])


def test_c4_hundred_thousand_entries_checked_by_the_reference_reader(built, tmp_path):
    """configs[3] on one GPU: 100 000 x 64 KiB entries written by the batch writer; the archive is then walked by the
    REFERENCE's reading loop (mz_zip_entry_read_open raw=0 -> mz_stream_zlib -> CRC check in mz_zip_entry_close,
    oracle/_ref/zipbatch_ref mode extract_ref: no product code in that process) and by the product's batch extractor."""
    exe, ref = _bin("zipbatch_cuda"), _bin("zipbatch_ref")
    d = os.path.join(_shm(tmp_path), "mz_c4_%d" % os.getpid())
    os.makedirs(d, exist_ok=True)
    try:
        arc = os.path.join(d, "c4.zip")
        r = subprocess.run([exe, arc, "100000", "65536", "6", "cuda"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
        j = json.loads(r.stdout.strip().splitlines()[-1])
        assert r.returncode == 0 and j["err"] == 0 and j["close_err"] == 0 and j["entries"] == 100000, (j, r.stderr[-400:])
        r = subprocess.run([ref, arc, "100000", "65536", "6", "extract_ref"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=1800)
        k = json.loads(r.stdout.strip().splitlines()[-1])
        # every entry's CRC is checked by mz_zip_entry_close (err == 0); every 101st entry is also regenerated and compared byte by byte
        assert r.returncode == 0 and k["err"] == 0 and k["entries"] == 100000 and k["mismatches"] == 0 and k["verified"] >= 990, (k, r.stderr[-400:])
        assert k["bytes_out"] == j["bytes_in"]
        r = subprocess.run([exe, arc, "100000", "65536", "6", "extract"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
        m = json.loads(r.stdout.strip().splitlines()[-1])
        assert r.returncode == 0 and m["err"] == 0 and m["entries"] == 100000 and m["mismatches"] == 0, (m, r.stderr[-400:])
    finally:
        subprocess.run(["rm", "-rf", d])


def test_c4_native_archive_zip64_paths(built, tmp_path):
    """The native writer's zip64 branches: more than 65535 entries (zip64 end record + locator) AND an archive larger than 4 GiB
    (zip64 extra field with the 64-bit header offset in the late central-directory records): 70 000 x 64 KiB at level 0. Read back
    by the reference's own loop (every CRC checked by mz_zip_entry_close) and opened by CPython's zipfile."""
    import zipfile
    exe, ref = _bin("zipbatch_cuda"), _bin("zipbatch_ref")
    d = os.path.join(_shm(tmp_path), "mz_c4n_%d" % os.getpid())
    os.makedirs(d, exist_ok=True)
    try:
        arc = os.path.join(d, "big.zip")
        r = subprocess.run([exe, arc, "70000", "65536", "0", "native"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=1200)
        j = json.loads(r.stdout.strip().splitlines()[-1])
        assert r.returncode == 0 and j["err"] == 0 and j["entries"] == 70000, (j, r.stderr[-400:])
        assert os.path.getsize(arc) > (1 << 32)
        r = subprocess.run([ref, arc, "70000", "65536", "0", "extract_ref"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=1800)
        k = json.loads(r.stdout.strip().splitlines()[-1])
        assert r.returncode == 0 and k["err"] == 0 and k["entries"] == 70000 and k["mismatches"] == 0 and k["bytes_out"] == j["bytes_in"], (k, r.stderr[-400:])
        with zipfile.ZipFile(arc) as z:
            infos = z.infolist()
            assert len(infos) == 70000 and infos[-1].header_offset > (1 << 32)
            assert z.read(infos[-1]) is not None and z.read(infos[0]) is not None
    finally:
        subprocess.run(["rm", "-rf", d])


def test_c4_native_hundred_thousand_entries(built, tmp_path):
    """configs[3] through the native writer: 100 000 x 64 KiB at level 6, checked entry by entry by the reference's reader"""
    exe, ref = _bin("zipbatch_cuda"), _bin("zipbatch_ref")
    d = os.path.join(_shm(tmp_path), "mz_c4m_%d" % os.getpid())
    os.makedirs(d, exist_ok=True)
    try:
        arc = os.path.join(d, "c4n.zip")
        r = subprocess.run([exe, arc, "100000", "65536", "6", "native"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
        j = json.loads(r.stdout.strip().splitlines()[-1])
        assert r.returncode == 0 and j["err"] == 0 and j["entries"] == 100000, (j, r.stderr[-400:])
        r = subprocess.run([ref, arc, "100000", "65536", "6", "extract_ref"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=1800)
        k = json.loads(r.stdout.strip().splitlines()[-1])
        assert r.returncode == 0 and k["err"] == 0 and k["entries"] == 100000 and k["mismatches"] == 0 and k["bytes_out"] == j["bytes_in"], (k, r.stderr[-400:])
    finally:
        subprocess.run(["rm", "-rf", d])
