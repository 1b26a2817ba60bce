#!/usr/bin/env python3
"""K7 (per-entry SHA-256) on one GPU: N entries of S bytes, device resident; checked against hashlib on a sample.
usage: bench_sha.py [entries] [entry_bytes]"""
import ctypes as C, hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as ge
import textgen
pkg = ge._load_pkg(); lib = pkg.load(); pkg.check(lib.mz_cuda_init())
torch.cuda.set_device(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
s = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
src = textgen.device(n * s, seed=12)
off = torch.arange(n, dtype=torch.int64, device="cuda") * s
ln = torch.full((n,), s, dtype=torch.int64, device="cuda")
dig = torch.zeros(n * 32, dtype=torch.uint8, device="cuda")
run = lambda: pkg.check(lib.mz_cuda_sha256_batch(src.data_ptr(), off.data_ptr(), ln.data_ptr(), n, dig.data_ptr(), None))
for _ in range(3): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
host = src[: 4 * s].cpu().numpy().tobytes(); d = dig[: 4 * 32].cpu().numpy().tobytes()
ok = all(d[32 * i:32 * i + 32] == hashlib.sha256(host[i * s:(i + 1) * s]).digest() for i in range(4))
print(json.dumps({"what": "sha256_batch", "entries": n, "entry_bytes": s, "ms": round(ms, 3), "GBps": round(n * s / ms / 1e6, 1), "ok": ok}))
