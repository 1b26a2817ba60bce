"""Drop-in proof: the REFERENCE's own `minigzip` / `minizip` programs and zip-container code, compiled where they lie
with `mz_strm_zlib.c` left out of the link and its names aliased to `mz_stream_cuda_*` (oracle/Makefile targets
`_ref/minigzip_cuda`, `_ref/minizip_cuda`; recipe = INTEGRATION.md 1b), must interoperate in both directions with the
unmodified reference builds (`_ref/minigzip_ref`, `_ref/minizip_ref`) and with Python's gzip / zipfile.

Call paths exercised: `minigzip.c:79-120` (create/set_prop/open/set_base/copy/close/delete, 16 KiB copy loop),
`mz_zip.c:1771-1852` (entry open), `:2047-2064` (entry CRC via mz_crypt_crc32_update), `:2116-2160` (close + verify).
"""
import gzip
import io
import os
import subprocess
import zipfile
import zlib

import pytest

import datagen

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def _bin(name):
    p = os.path.join(REFDIR, name)
    if not os.path.exists(p):
        pytest.skip(f"{name} not built (needs /root/reference at build time)")
    return p


def _run(args, cwd, ok=True):
    r = subprocess.run(args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    if ok:
        assert r.returncode == 0, (args, r.returncode, r.stdout[-600:], r.stderr[-600:])
    return r


def _corpus():
    return {
        "text.txt": datagen.text_like(3_000_000, seed=21),
        "records.bin": datagen.binary_records(1_500_000, seed=22),
        "random.bin": datagen.random_bytes(400_000, seed=23),
        "mixed.dat": datagen.mixed(2_000_000, seed=24),
        "tiny.txt": b"hello, hello, hello\n",
        "one.bin": b"\x00",
        "empty.bin": b"",
    }


@pytest.mark.gpu
@pytest.mark.parametrize("level", [1, 6, 9])
def test_minigzip_cuda_writes_what_everyone_reads(built, tmp_path, level):
    exe, ref = _bin("minigzip_cuda"), _bin("minigzip_ref")
    data = datagen.text_like(5_000_000, seed=30 + level) + datagen.random_bytes(100_000, seed=3)
    src = tmp_path / "doc.txt"
    src.write_bytes(data)
    r = _run([exe, f"-{level}", "doc.txt"], tmp_path)
    assert b"Operation completed successfully" in r.stdout
    gz = (tmp_path / "doc.txt.gz").read_bytes()
    assert gz[:3] == b"\x1f\x8b\x08" and int.from_bytes(gz[-4:], "little") == len(data)
    assert int.from_bytes(gz[-8:-4], "little") == zlib.crc32(data)
    assert gzip.decompress(gz) == data
    assert len(gz) < 0.62 * len(data)
    out = tmp_path / "x"
    _run([ref, "-x", "-d", "x", "doc.txt.gz"], tmp_path)
    assert (out / "doc.txt").read_bytes() == data


@pytest.mark.gpu
def test_minigzip_cuda_reads_what_the_reference_and_gzip_write(built, tmp_path):
    exe, ref = _bin("minigzip_cuda"), _bin("minigzip_ref")
    data = datagen.mixed(6_000_000, seed=41)
    (tmp_path / "a.bin").write_bytes(data)
    _run([ref, "-6", "a.bin"], tmp_path)
    _run([exe, "-x", "-d", "o1", "a.bin.gz"], tmp_path)
    assert (tmp_path / "o1" / "a.bin").read_bytes() == data
    buf = io.BytesIO()
    with gzip.GzipFile(filename="named-member.bin", fileobj=buf, mode="wb", compresslevel=9, mtime=1) as f:  # FNAME header field
        f.write(data)
    (tmp_path / "b.bin.gz").write_bytes(buf.getvalue())
    _run([exe, "-x", "-d", "o2", "b.bin.gz"], tmp_path)
    assert (tmp_path / "o2" / "b.bin").read_bytes() == data
    # corrupt one payload byte: the reference prints the stream error and exits non-zero
    bad = bytearray(buf.getvalue())
    bad[len(bad) // 2] ^= 0x10
    (tmp_path / "c.bin.gz").write_bytes(bytes(bad))
    r = _run([exe, "-x", "-d", "o3", "c.bin.gz"], tmp_path, ok=False)
    assert r.returncode != 0 and b"Error" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("level", [1, 6])
def test_minizip_cuda_archive_is_valid_everywhere(built, tmp_path, level):
    exe, ref = _bin("minizip_cuda"), _bin("minizip_ref")
    files = _corpus()
    for name, blob in files.items():
        (tmp_path / name).write_bytes(blob)
    _run([exe, "-o", f"-{level}", "a.zip"] + sorted(files), tmp_path)
    with zipfile.ZipFile(tmp_path / "a.zip") as z:
        assert z.testzip() is None  # CRC of every member
        assert sorted(z.namelist()) == sorted(files)
        for name, blob in files.items():
            info = z.getinfo(name)
            assert z.read(name) == blob and info.CRC == zlib.crc32(blob) and info.file_size == len(blob)
            if len(blob) > 100:
                assert info.compress_type == zipfile.ZIP_DEFLATED
        assert z.getinfo("text.txt").compress_size < 0.6 * len(files["text.txt"])
    _run([ref, "-x", "-o", "-d", "out", "a.zip"], tmp_path)  # the unmodified reference extracts and CRC-checks
    for name, blob in files.items():
        assert (tmp_path / "out" / name).read_bytes() == blob


@pytest.mark.gpu
def test_minizip_cuda_extracts_foreign_archives(built, tmp_path):
    exe, ref = _bin("minizip_cuda"), _bin("minizip_ref")
    files = _corpus()
    for name, blob in files.items():
        (tmp_path / name).write_bytes(blob)
    _run([ref, "-o", "-9", "r.zip"] + sorted(files), tmp_path)
    _run([exe, "-x", "-o", "-d", "o1", "r.zip"], tmp_path)
    with zipfile.ZipFile(tmp_path / "p.zip", "w", zipfile.ZIP_DEFLATED, compresslevel=6) as z:
        for name, blob in files.items():
            z.writestr(name, blob)
    _run([exe, "-x", "-o", "-d", "o2", "p.zip"], tmp_path)
    for name, blob in files.items():
        assert (tmp_path / "o1" / name).read_bytes() == blob
        assert (tmp_path / "o2" / name).read_bytes() == blob
    # flip a bit inside a member's compressed data: extraction must report an error (data error or CRC mismatch)
    raw = bytearray((tmp_path / "p.zip").read_bytes())
    with zipfile.ZipFile(tmp_path / "p.zip") as z:
        info = z.getinfo("text.txt")
    raw[info.header_offset + 30 + len("text.txt") + info.compress_size // 2] ^= 0x04
    (tmp_path / "bad.zip").write_bytes(bytes(raw))
    r = _run([exe, "-x", "-o", "-d", "o3", "bad.zip"], tmp_path, ok=False)
    assert r.returncode != 0 or b"Error" in r.stdout


def test_dropin_binaries_refuse_to_run_without_a_gpu(built, tmp_path):
    """No CPU fallback: on a box without a GPU the drop-in CLI must fail loudly rather than produce output."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    exe = _bin("minigzip_cuda")
    (tmp_path / "f.txt").write_bytes(b"abc" * 1000)
    r = _run([exe, "-6", "f.txt"], tmp_path, ok=False)
    assert r.returncode != 0 and b"no CPU fallback" in r.stderr
    assert not (tmp_path / "f.txt.gz").exists() or (tmp_path / "f.txt.gz").stat().st_size == 0


@pytest.mark.gpu
@pytest.mark.parametrize("level", [1, 6])
def test_zip_batch_writer_on_the_raw_seam(built, tmp_path, level):
    """config C4 shape: thousands of in-memory entries compressed + CRC'd in one GPU batch and written by
    the REFERENCE's container code through mz_zip_entry_write_open(raw=1)/close_raw; the archive must be valid for
    zipfile and for the unmodified reference extractor, and agree entry by entry (CRC, size) with the archive the
    reference's own zlib path writes from the same buffers."""
    exe, ref = _bin("zipbatch_cuda"), _bin("minizip_ref")
    n, esz = 3000, 65536
    (tmp_path / "dump").mkdir()
    r = _run([exe, "c.zip", str(n), str(esz), str(level), "cuda", "dump", "37"], tmp_path)
    import json
    stats = json.loads(r.stdout.decode().strip().splitlines()[-1])
    assert stats["err"] == 0 and stats["close_err"] == 0 and stats["entries"] == n and stats["rounds"] >= 1
    _run([exe, "r.zip", str(n), str(esz), str(level), "ref"], tmp_path)
    with zipfile.ZipFile(tmp_path / "c.zip") as zc, zipfile.ZipFile(tmp_path / "r.zip") as zr:
        assert zc.testzip() is None
        assert zc.namelist() == zr.namelist() == ["e/%06d" % i for i in range(n)]
        tot_c = tot_r = 0
        for i in range(n):
            a, b = zc.getinfo("e/%06d" % i), zr.getinfo("e/%06d" % i)
            assert a.CRC == b.CRC and a.file_size == b.file_size, i
            assert a.compress_type == zipfile.ZIP_DEFLATED
            tot_c += a.compress_size
            tot_r += b.compress_size
        assert zc.getinfo("e/000005").file_size == 0 and zc.getinfo("e/000006").file_size == 1
        assert tot_c < 1.25 * tot_r  # compression ratio in the reference's neighbourhood
        for i in range(0, n, 37):
            assert zc.read("e/%06d" % i) == (tmp_path / "dump" / ("%06d" % i)).read_bytes()
    _run([ref, "-x", "-o", "-d", "out", "c.zip"], tmp_path)  # the reference's extractor CRC-checks every entry
    for i in range(0, n, 37):
        assert (tmp_path / "out" / "e" / ("%06d" % i)).read_bytes() == (tmp_path / "dump" / ("%06d" % i)).read_bytes()


@pytest.mark.gpu
def test_zip_batch_extractor_reads_every_kind_of_archive(built, tmp_path):
    """scope row f2: the central directory is walked by the reference's reader, the compressed bytes come through its raw
    seam, all entries are inflated + CRC'd in one launch each. Archives written by the batch writer (data descriptors), by the
    reference's zlib path, and by zipfile (stored + deflated entries) must all extract with matching content; a damaged
    archive must be refused."""
    import json
    exe = _bin("zipbatch_cuda")
    n, esz = 2500, 65536
    _run([exe, "c.zip", str(n), str(esz), "6", "cuda"], tmp_path)
    _run([exe, "r.zip", str(n), str(esz), "9", "ref"], tmp_path)
    for name in ("c.zip", "r.zip"):
        got = json.loads(_run([exe, name, str(n), str(esz), "0", "extract"], tmp_path).stdout.decode().strip().splitlines()[-1])
        want = json.loads(_run([exe, name, str(n), str(esz), "0", "extract_ref"], tmp_path).stdout.decode().strip().splitlines()[-1])
        assert got["err"] == 0 and got["entries"] == want["entries"] == n and got["bytes_out"] == want["bytes_out"], (name, got, want)
        assert got["verified"] >= n // 101 and got["mismatches"] == 0 and got["rounds"] >= 1
    # zipfile: stored and deflated members side by side, names that are not of the e/%06d form (no regeneration check)
    with zipfile.ZipFile(tmp_path / "p.zip", "w") as z:
        for i in range(300):
            blob = datagen.mixed(1000 + 137 * i, seed=i) if i % 3 else datagen.random_bytes(500 + i, seed=i)
            z.writestr("dir/f%04d.bin" % i, blob, zipfile.ZIP_DEFLATED if i % 3 else zipfile.ZIP_STORED, compresslevel=6)
        z.writestr("dir/empty.bin", b"")
        total = sum(zi.file_size for zi in z.infolist())
    got = json.loads(_run([exe, "p.zip", "301", "200000", "0", "extract"], tmp_path).stdout.decode().strip().splitlines()[-1])
    assert got["err"] == 0 and got["entries"] == 301 and got["bytes_out"] == total
    # damage a deflated member: the batch must be refused with an error, like the reference's close() refuses it
    raw = bytearray((tmp_path / "p.zip").read_bytes())
    with zipfile.ZipFile(tmp_path / "p.zip") as z:
        info = z.getinfo("dir/f0200.bin")
    raw[info.header_offset + 30 + len(info.filename) + info.compress_size // 2] ^= 0x40
    (tmp_path / "bad.zip").write_bytes(bytes(raw))
    r = _run([exe, "bad.zip", "301", "200000", "0", "extract"], tmp_path, ok=False)
    bad = json.loads(r.stdout.decode().strip().splitlines()[-1])
    assert r.returncode != 0 and bad["err"] in (-3, -105)  # MZ_DATA_ERROR or MZ_CRC_ERROR


@pytest.mark.gpu
def test_zip_batch_sha256_extrafield_on_the_gpu(built, tmp_path):
    """scope row f3 on the device: K7 hashes stored in MZ_ZIP_EXTENSION_HASH are verified by hashlib, by the reference built with
    its crypto provider (minizip_refc) and by the batch extractor; the reference's own hashes are verified by K7; a damaged
    digest is refused (same scenario as the CPU suite runs on the emulator)."""
    import test_emu_dropin_cli as emu_cli
    refc = os.path.join(REFDIR, "minizip_refc")
    if not os.path.exists(refc):
        pytest.skip("oracle/_ref/minizip_refc not built (no OpenSSL headers)")

    def run(args, cwd, ok=True):
        return _run([str(a) for a in args], cwd, ok)
    emu_cli._sha_roundtrip(_bin("zipbatch_cuda"), refc, tmp_path, run)


@pytest.mark.gpu
def test_native_archive_writer_on_the_gpu(built, tmp_path):
    """mz_zip_cuda_write_archive on the device (region assembly by K4 gather + header scatter), same checks as on the emulator,
    plus the hash variant verified by the crypto-enabled reference"""
    import test_emu_dropin_cli as emu_cli

    def run(args, cwd, ok=True):
        return _run([str(a) for a in args], cwd, ok)
    bins = {"minizip_ref": _bin("minizip_ref")}
    emu_cli._native_archive_checks(_bin("zipbatch_cuda"), bins, tmp_path, run, n=3000, esz=65536)
    refc = os.path.join(REFDIR, "minizip_refc")
    if os.path.exists(refc):
        run([_bin("zipbatch_cuda"), "ns.zip", "500", "40000", "6", "native_sha"], tmp_path)
        run([refc, "-x", "-o", "-d", "out_ns", "ns.zip"], tmp_path)
