"""The whole product on the CPU: tests/emu/libmz_strm_emu.so is the REAL host-side C (mz_strm_cuda.c, mz_crypt_cuda.c,
mz_zip_cuda.c) and the REAL extern "C" shim (mz_cuda_api.cu, compiled as C++) linked against the execution-model emulator
and a host implementation of the handful of CUDA runtime calls (tests/emu/shim/cuda_runtime.h). The vtbl streams are then
driven exactly as the GPU tests drive them -- so the host logic (batching, framing, window sliding, K5/K6 hand-over, error
taxonomy of mz_strm_zlib.c:116-305) is covered without a GPU. One scenario per process: workspaces are pooled per process.
This is test infrastructure: nothing here is part of, or a fallback for, the product library."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")


@pytest.fixture(scope="module")
def emulib(built):
    r = subprocess.run(["make", "-s", "-C", EMU, "libmz_strm_emu.so"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    return os.path.join(EMU, "libmz_strm_emu.so")


def _scenario(name, **env):
    e = dict(os.environ)
    e.update({k: str(v) for k, v in env.items()})
    r = subprocess.run([sys.executable, os.path.join(EMU, "product_run.py"), name], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       timeout=1500)
    assert r.returncode == 0 and ("scenario %s ok" % name).encode() in r.stdout, r.stdout[-4000:].decode(errors="replace")
    return r.stdout.decode(errors="replace")


def test_write_path_batches_framing_and_reference_readback(emulib):
    _scenario("write", MZ_CUDA_BATCH_KB=256)


@pytest.mark.parametrize("spec", [1, 0])
def test_read_path_small_windows_with_and_without_k6(emulib, spec):
    out = _scenario("read", MZ_CUDA_SPEC=spec, MZ_CUDA_SPEC_SEG_KB=4, MZ_CUDA_BATCH_KB=1024, MZ_CUDA_TRACE=1)
    rounds = out.count("K6 round")
    assert (rounds >= 4) if spec else (rounds == 0)


def test_read_path_long_member_switches_windows_and_overlaps_rounds(emulib):
    out = _scenario("long", MZ_CUDA_SPEC_SEG_KB=4, MZ_CUDA_BATCH_KB=1024, MZ_CUDA_READ_WINDOW_KB=2048, MZ_CUDA_TRACE=1)
    assert out.count("K6 round") >= 3


def test_read_path_long_member_alternates_output_windows(emulib):
    """output window as small as the compressed one: every other round finds it full while bytes are still being delivered, and
    decoding moves on in the second window (history carried over)"""
    out = _scenario("long", MZ_CUDA_SPEC_SEG_KB=4, MZ_CUDA_BATCH_KB=1024, MZ_CUDA_READ_WINDOW_KB=1100, MZ_CUDA_READ_OUT_MULT=1, MZ_CUDA_TRACE=1)
    assert out.count("decoding on in window") >= 4


def test_read_path_compressed_window_ring_with_read_ahead(emulib):
    import re
    out = _scenario("ring", MZ_CUDA_SPEC_SEG_KB=4, MZ_CUDA_BATCH_KB=1024, MZ_CUDA_READ_WINDOW_KB=1100, MZ_CUDA_READ_AHEAD=2, MZ_CUDA_READ_STATS=1)
    m = re.findall(r"for (\d+) bytes read ahead under rounds in flight; (\d+) window uploads wrapped", out)
    assert m and max(int(a) for a, _ in m) > (1 << 20) and max(int(b) for _, b in m) >= 1, out[-2000:]


def test_crc_symbol_device_path(emulib):
    _scenario("crc", MZ_CUDA_CRC_MIN_BYTES=65536)


def test_sharded_deflate_layout_pieces_rows_and_crc(emulib):
    """mz_cuda_deflate_sharded (the multi-GPU entry point of the C library) with three shards on the emulator's device"""
    _scenario("sharded")


def test_host_paths_under_sanitizers(emulib):
    """the same scenarios with the whole library (host C included) built with AddressSanitizer + UBSan: window arithmetic that
    runs off a staging buffer, a stale pointer after the workspace switches windows, a shift by 32 ... abort the run"""
    r = subprocess.run(["make", "-s", "-C", EMU, "libmz_strm_emu_san.so"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    asan = subprocess.run(["/usr/bin/gcc", "-print-file-name=libasan.so"], stdout=subprocess.PIPE).stdout.decode().strip()
    if r.returncode != 0 or not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("no sanitizer runtime in this toolchain")
    san = dict(MZ_EMU_LIB=os.path.join(EMU, "libmz_strm_emu_san.so"), LD_PRELOAD=asan,
               ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1")
    _scenario("write", MZ_CUDA_BATCH_KB=256, **san)
    _scenario("read", MZ_CUDA_SPEC=1, MZ_CUDA_SPEC_SEG_KB=4, MZ_CUDA_BATCH_KB=1024, **san)
    _scenario("long", MZ_CUDA_SPEC_SEG_KB=4, MZ_CUDA_BATCH_KB=1024, MZ_CUDA_READ_WINDOW_KB=2048, **san)
    _scenario("long", MZ_CUDA_SPEC_SEG_KB=4, MZ_CUDA_BATCH_KB=1024, MZ_CUDA_READ_WINDOW_KB=1100, MZ_CUDA_READ_OUT_MULT=1, **san)
    _scenario("ring", MZ_CUDA_SPEC_SEG_KB=4, MZ_CUDA_BATCH_KB=1024, MZ_CUDA_READ_WINDOW_KB=1100, MZ_CUDA_READ_AHEAD=2, MZ_TEST_RING_SMALL=1, **san)
