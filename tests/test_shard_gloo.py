"""World-size-2 gloo test of the N>1 host path (no GPU): shard -> per-rank streams -> one all-gather -> join.

Per-rank compression is done by the product's kernel sources on the CPU emulator (tests/emu), so the bytes that
travel are exactly what the GPUs would produce: non-final shards end with the sync marker, the last carries BFINAL.
Rank 0 checks that the rank-ordered concatenation is ONE valid stream that the oracle inflates to the whole input,
and that the host crc32_combine fold of the gathered per-chunk table equals the CRC of the whole input.
"""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

WORKER = r'''
import os, sys, zlib
sys.path.insert(0, %(root)r); sys.path.insert(0, %(here)r)
import torch, torch.distributed as dist
import __graft_entry__ as ge
import emushim, datagen, refshim
pkg = ge._load_pkg()
import importlib
shard = importlib.import_module("minizip_ng_b200.shard")
lib = pkg.load()
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
data = datagen.mixed(7 * 65536 + 1234, 42)          # 8 chunks, ragged last one
nchunks = (len(data) + 65535) // 65536
lo, hi = shard.unit_range(rank, world, nchunks)
mine = data[lo * 65536:min(hi * 65536, len(data))]
emu = emushim.EmuLib()
comp, lens = emu.deflate(mine, level=1, final=shard.is_last_owner(rank, world, nchunks))
_, crcs = emu.crc32(mine, 65536)
rows = [[crcs[i], min(65536, len(mine) - i * 65536), lens[i]] for i in range(len(lens))]
stream = torch.frombuffer(bytearray(comp), dtype=torch.uint8)
streams, tables = shard.all_gather_streams(stream, len(comp), torch.tensor(rows, dtype=torch.int64))
joined = b"".join(bytes(s.numpy().tobytes()) for s in streams)
orc = refshim.OracleLib()
err, out, cons = orc.inflate(joined, len(data) + 8)
assert err == 0 and out == data and cons == len(joined), (rank, err)
crc, total = shard.fold_crc(tables, lib.mz_cuda_crc32_combine)
assert total == len(data) and crc == zlib.crc32(data) == orc.crc32(0, data), (rank, hex(crc))
assert sum(int(t[:, 2].sum()) for t in tables) == len(joined)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_unit_range_partition():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge._load_pkg()
    import importlib
    shard = importlib.import_module("minizip_ng_b200.shard")
    for n in (0, 1, 7, 8, 262144, 100000):
        for w in (1, 2, 4, 8):
            covered = []
            last = [r for r in range(w) if shard.is_last_owner(r, w, n)]
            assert len(last) == 1
            for r in range(w):
                lo, hi = shard.unit_range(r, w, n)
                covered += list(range(lo, hi)) if n < 100 else []
                assert 0 <= lo <= hi <= n
            if n < 100:
                assert covered == list(range(n))


def test_two_rank_gather_and_join(built):
    r = subprocess.run(["make", "-s"], cwd=os.path.join(HERE, "emu"), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    port = _free_port()
    src = WORKER % {"root": ROOT, "here": HERE, "port": port}
    procs = [subprocess.Popen([sys.executable, "-c", src, str(rank)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for rank in range(2)]
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            p.kill()
            o, _ = p.communicate()
        outs.append(o)
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and ("rank %d ok" % rank) in o, o[-2000:]
