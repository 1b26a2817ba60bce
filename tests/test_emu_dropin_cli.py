"""The drop-in proof without a GPU: the REFERENCE's `minigzip` / `minizip` programs and container code, compiled where they
lie with `mz_strm_zlib.c` left out and its names aliased to `mz_stream_cuda_*` (oracle/Makefile `emu` targets), linked against
the product built on the CPU execution-model emulator (tests/emu/libmz_strm_emu.so = real host C + real API shim + kernel
sources). Same checks as tests/test_gpu_dropin_cli.py on smaller inputs: both directions against the unmodified reference
builds and Python's gzip / zipfile; the batch zip writer / extractor on the raw-entry seam."""
import gzip
import json
import os
import subprocess
import zipfile
import zlib

import pytest

import datagen

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFDIR = os.path.join(ROOT, "oracle", "_ref")
ENV = dict(os.environ, MZ_CUDA_BATCH_KB="512", MZ_CUDA_SPEC_SEG_KB="4")


@pytest.fixture(scope="module")
def emu_bins(built):
    if not os.path.exists("/root/reference/mz_strm_zlib.c"):
        pytest.skip("needs the reference sources at build time")
    r = subprocess.run(["make", "-s", "-C", os.path.join(HERE, "emu"), "libmz_strm_emu.so"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    r = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "emu"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    return {n: os.path.join(REFDIR, n) for n in ("minigzip_emu", "minizip_emu", "zipbatch_emu", "minigzip_ref", "minizip_ref")}


def _run(args, cwd, ok=True):
    r = subprocess.run(args, cwd=cwd, env=ENV, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
    if ok:
        assert r.returncode == 0, (args, r.returncode, r.stdout[-600:], r.stderr[-600:])
    return r


def _corpus():
    return {"text.txt": datagen.text_like(900_000, seed=21), "records.bin": datagen.binary_records(300_000, seed=22),
            "random.bin": datagen.random_bytes(100_000, seed=23), "tiny.txt": b"hello, hello, hello\n", "empty.bin": b""}


@pytest.mark.parametrize("level", [1, 6])
def test_minigzip_both_directions(emu_bins, tmp_path, level):
    data = datagen.text_like(1_200_000, seed=30 + level) + datagen.random_bytes(50_000, seed=3)
    (tmp_path / "doc.txt").write_bytes(data)
    r = _run([emu_bins["minigzip_emu"], f"-{level}", "doc.txt"], tmp_path)
    assert b"Operation completed successfully" in r.stdout
    gz = (tmp_path / "doc.txt.gz").read_bytes()
    assert gzip.decompress(gz) == data and int.from_bytes(gz[-8:-4], "little") == zlib.crc32(data)
    _run([emu_bins["minigzip_ref"], "-x", "-d", "x", "doc.txt.gz"], tmp_path)
    assert (tmp_path / "x" / "doc.txt").read_bytes() == data
    (tmp_path / "a.bin").write_bytes(data)
    _run([emu_bins["minigzip_ref"], "-9", "a.bin"], tmp_path)
    _run([emu_bins["minigzip_emu"], "-x", "-d", "o1", "a.bin.gz"], tmp_path)
    assert (tmp_path / "o1" / "a.bin").read_bytes() == data
    bad = bytearray((tmp_path / "a.bin.gz").read_bytes())
    bad[len(bad) // 2] ^= 0x10
    (tmp_path / "c.bin.gz").write_bytes(bytes(bad))
    r = _run([emu_bins["minigzip_emu"], "-x", "-d", "o3", "c.bin.gz"], tmp_path, ok=False)
    assert r.returncode != 0 and b"Error" in r.stdout


def test_minizip_both_directions(emu_bins, tmp_path):
    files = _corpus()
    for name, blob in files.items():
        (tmp_path / name).write_bytes(blob)
    _run([emu_bins["minizip_emu"], "-o", "-6", "a.zip"] + sorted(files), tmp_path)
    with zipfile.ZipFile(tmp_path / "a.zip") as z:
        assert z.testzip() is None and sorted(z.namelist()) == sorted(files)
        for name, blob in files.items():
            assert z.read(name) == blob and z.getinfo(name).CRC == zlib.crc32(blob)
    _run([emu_bins["minizip_ref"], "-x", "-o", "-d", "out", "a.zip"], tmp_path)  # the unmodified reference extracts and CRC-checks
    _run([emu_bins["minizip_ref"], "-o", "-9", "r.zip"] + sorted(files), tmp_path)
    _run([emu_bins["minizip_emu"], "-x", "-o", "-d", "o1", "r.zip"], tmp_path)   # ... and we extract what it wrote
    for name, blob in files.items():
        assert (tmp_path / "out" / name).read_bytes() == blob and (tmp_path / "o1" / name).read_bytes() == blob


def test_zip_batch_writer_and_extractor(emu_bins, tmp_path):
    exe = emu_bins["zipbatch_emu"]
    n, esz = 260, 30_000
    (tmp_path / "dump").mkdir()
    st = json.loads(_run([exe, "c.zip", str(n), str(esz), "6", "cuda", "dump", "13"], tmp_path).stdout.decode().strip().splitlines()[-1])
    assert st["err"] == 0 and st["close_err"] == 0 and st["entries"] == n
    _run([exe, "r.zip", str(n), str(esz), "6", "ref"], tmp_path)
    with zipfile.ZipFile(tmp_path / "c.zip") as zc, zipfile.ZipFile(tmp_path / "r.zip") as zr:
        assert zc.testzip() is None and zc.namelist() == zr.namelist() == ["e/%06d" % i for i in range(n)]
        for i in range(n):
            a, b = zc.getinfo("e/%06d" % i), zr.getinfo("e/%06d" % i)
            assert a.CRC == b.CRC and a.file_size == b.file_size
        for i in range(0, n, 13):
            assert zc.read("e/%06d" % i) == (tmp_path / "dump" / ("%06d" % i)).read_bytes()
    _run([emu_bins["minizip_ref"], "-x", "-o", "-d", "out", "c.zip"], tmp_path)
    for name in ("c.zip", "r.zip"):
        got = json.loads(_run([exe, name, str(n), str(esz), "0", "extract"], tmp_path).stdout.decode().strip().splitlines()[-1])
        want = json.loads(_run([exe, name, str(n), str(esz), "0", "extract_ref"], tmp_path).stdout.decode().strip().splitlines()[-1])
        assert got["err"] == 0 and got["entries"] == want["entries"] == n and got["bytes_out"] == want["bytes_out"] and got["mismatches"] == 0
    raw = bytearray((tmp_path / "r.zip").read_bytes())
    with zipfile.ZipFile(tmp_path / "r.zip") as z:
        info = z.getinfo("e/000100")
    raw[info.header_offset + 30 + len(info.filename) + info.compress_size // 2] ^= 0x40
    (tmp_path / "bad.zip").write_bytes(bytes(raw))
    r = _run([exe, "bad.zip", str(n), str(esz), "0", "extract"], tmp_path, ok=False)
    assert r.returncode != 0 and json.loads(r.stdout.decode().strip().splitlines()[-1])["err"] in (-3, -105)
