"""Scope row f4 (WinZip AES after the codec) on the device: K8's three kernels against hashlib / `cryptography`, and whole AES
archives both ways against the reference built with mz_strm_wzaes.c + OpenSSL -- the scenarios of tests/test_emu_wzaes.py on the
real library, with more and larger entries."""
import ctypes as C
import os
import subprocess

import pytest

import wzaes_checks

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFDIR = os.path.join(ROOT, "oracle", "_ref")


@pytest.mark.gpu
def test_k8_known_answers_on_the_gpu(built):
    pytest.importorskip("cryptography")
    import torch
    import __graft_entry__ as ge
    pkg = ge._load_pkg()
    lib = pkg.load()
    pkg.check(lib.mz_cuda_init())
    torch.cuda.set_device(0)

    def buf(b):
        t = torch.zeros(max(len(b), 1) + 16, dtype=torch.uint8, device="cuda")
        if len(b):
            t[:len(b)] = torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
        return C.c_void_p(t.data_ptr()), t

    tensors = {}

    def buf2(b):
        p, t = buf(b)
        tensors[p.value] = t
        return p, t

    def back(p, n):
        torch.cuda.synchronize()
        return bytes(tensors[p.value][:n].cpu().numpy().tobytes())

    def call(name, *args):
        r = getattr(lib, name)(*args)
        torch.cuda.synchronize()
        return r
    wzaes_checks.run(call, buf2, back, lens=(0, 1, 15, 16, 17, 4097, 65536, 1_000_003))


@pytest.mark.gpu
def test_aes_archives_both_ways_on_the_gpu(built, tmp_path):
    exe, refe = os.path.join(REFDIR, "zipbatch_cudae"), os.path.join(REFDIR, "minizip_refe")
    if not os.path.exists(exe) or not os.path.exists(refe):
        pytest.skip("oracle/_ref/zipbatch_cudae / minizip_refe not built (need /root/reference and OpenSSL headers at build time)")

    def run_cmd(args, cwd, ok=True):
        r = subprocess.run([str(a) for a in args], cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
        if ok:
            assert r.returncode == 0, (args, r.returncode, r.stdout[-600:], r.stderr[-600:])
        return r
    wzaes_checks.aes_roundtrip(exe, refe, tmp_path, run_cmd, n=2500, esz=65536)
