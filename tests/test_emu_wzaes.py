"""Scope row f4 (the part that follows the codec) on the CPU: K8 = PBKDF2-HMAC-SHA1 / AES-CTR / HMAC-SHA1 for a batch of zip
entries (minizip-ng_b200/csrc/wzaes_kernel.cuh) run on the execution-model emulator, against hashlib and the `cryptography`
package; then whole AES archives both ways against the reference built with its WinZip AES stream (oracle/_ref/minizip_refe =
mz_strm_wzaes.c over mz_crypt_openssl.c, compiled where they lie)."""
import ctypes as C
import os
import subprocess

import pytest

import wzaes_checks

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFDIR = os.path.join(ROOT, "oracle", "_ref")


@pytest.fixture(scope="module")
def emuprod():
    r = subprocess.run(["make", "-s", "-C", os.path.join(HERE, "emu"), "libmz_strm_emu.so"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    import cuharness
    return cuharness.pkg().configure(C.CDLL(os.path.join(HERE, "emu", "libmz_strm_emu.so")))


def test_k8_known_answers_on_the_emulator(emuprod):
    pytest.importorskip("cryptography")
    lib = emuprod

    def buf(b):
        m = C.create_string_buffer(bytes(b), max(len(b), 1))
        return C.cast(m, C.c_void_p), m

    def back(p, n):
        return C.string_at(p, n)

    def call(name, *args):
        return getattr(lib, name)(*args)
    wzaes_checks.run(call, buf, back, lens=(0, 1, 15, 16, 17, 4097, 70001))


def test_aes_archives_both_ways_against_the_reference(emuprod, tmp_path):
    if not os.path.exists("/root/reference/mz_strm_wzaes.c") or not os.path.exists("/usr/include/openssl/sha.h"):
        pytest.skip("needs the reference sources and OpenSSL headers at build time")
    r = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_ref/minizip_refe", "_ref/zipbatch_emue"], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]

    def run_cmd(args, cwd, ok=True):
        r = subprocess.run([str(a) for a in args], cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
        if ok:
            assert r.returncode == 0, (args, r.returncode, r.stdout[-600:], r.stderr[-600:])
        return r
    wzaes_checks.aes_roundtrip(os.path.join(REFDIR, "zipbatch_emue"), os.path.join(REFDIR, "minizip_refe"), tmp_path, run_cmd)
