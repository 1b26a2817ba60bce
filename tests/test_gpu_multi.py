"""Multi-GPU entry point of the C library (include/mz_cuda_batch.h, mz_cuda_deflate_sharded): one process, several devices,
all-gather by peer-to-peer copies on the copy engines. Needs >= 2 GPUs (gpurun --gpus 2); skipped otherwise."""
import ctypes as C
import zlib

import pytest

pytestmark = pytest.mark.gpu


def test_deflate_sharded_across_devices(built):
    import torch
    import cuharness
    import textgen
    p = cuharness.pkg()
    lib = p.load()
    assert lib.mz_cuda_init() == 0
    ndev = min(torch.cuda.device_count(), 4)
    if ndev < 2:
        pytest.skip("needs at least two GPUs")
    n = 96 * 1024 * 1024 + 70_001  # ragged last chunk
    host = textgen.host(n, seed=31)
    nch = (n + 65535) // 65536
    cuts = [(r * nch // ndev) * 65536 for r in range(ndev)] + [n]
    bounds = [int(lib.mz_cuda_gather_region_bound(cuts[r + 1] - cuts[r])) for r in range(ndev)]
    cap = sum(bounds)
    ins, gath, rows = [], [], []
    shards = (p.Shard * ndev)()
    for r in range(ndev):
        with torch.cuda.device(r):
            t = torch.frombuffer(bytearray(host[cuts[r]:cuts[r + 1]]), dtype=torch.uint8).cuda()
            g = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda")
            rw = torch.zeros(3 * nch, dtype=torch.int32, device="cuda")
        ins.append(t); gath.append(g); rows.append(rw)
        shards[r].device = r
        shards[r].d_in = t.data_ptr()
        shards[r].len = cuts[r + 1] - cuts[r]
        shards[r].d_gathered = g.data_ptr()
        shards[r].gathered_cap = cap
        shards[r].d_rows = rw.data_ptr()
    for r in range(ndev):
        torch.cuda.synchronize(r)
    roff = (C.c_uint64 * ndev)()
    slen = (C.c_uint64 * ndev)()
    crc = C.c_uint32(0)
    for pieces in (1, 4):
        err = lib.mz_cuda_deflate_sharded(shards, ndev, 1, pieces, roff, slen, C.byref(crc))
        assert err == 0, (err, lib.mz_cuda_last_error())
        assert crc.value == zlib.crc32(host)
        streams = []
        for r in range(ndev):
            gh = gath[r].cpu().numpy().tobytes()
            streams.append(b"".join(gh[roff[i]:roff[i] + slen[i]] for i in range(ndev)))
        assert all(s == streams[0] for s in streams)       # every device holds every stream
        d = zlib.decompressobj(-15)
        assert d.decompress(streams[0]) == host and d.eof and d.unused_data == b""
        r0 = rows[0].cpu().numpy().astype("uint32").tolist()
        for r in range(1, ndev):
            assert rows[r].cpu().numpy().astype("uint32").tolist() == r0
        assert sum(r0[1::3]) == n and sum(r0[2::3]) == sum(slen)
        assert r0[0] == zlib.crc32(host[:65536])
