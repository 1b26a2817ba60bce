"""Multi-GPU plumbing for the chunk-sharded path.

One process per GPU. Units (64 KiB chunks or zip entries) are independent, so rank r simply owns the
contiguous unit range [r*n/W, (r+1)*n/W): concatenation order = rank order, input never crosses NVLink.
The only collective is an all-gather of (a) every rank's joined bitstream, padded to the largest, and (b) the
per-unit {crc32, in_len, out_len} table; the CRC of the whole is a host fold of crc32_combine in order.
Works with any torch.distributed backend (NCCL on GPUs; gloo on CPU for the tests).
"""
import torch
import torch.distributed as dist


def unit_range(rank, world, n_units):
    """Contiguous block partition: rank -> [lo, hi)."""
    return rank * n_units // world, (rank + 1) * n_units // world


def is_last_owner(rank, world, n_units):
    """Only the rank that owns the globally last unit sets BFINAL (it is the last rank with a non-empty range)."""
    lo, hi = unit_range(rank, world, n_units)
    return hi == n_units and hi > lo or (n_units == 0 and rank == world - 1)


def all_gather_streams(local_stream, local_len, table, group=None):
    """Gather every rank's byte stream and per-unit table.

    local_stream : uint8 tensor holding at least local_len valid bytes (device = backend's device)
    table        : int64 tensor [n_local_units, k] (e.g. columns crc32, in_len, out_len)
    returns (streams: list of uint8 tensors, one per rank, exact lengths; tables: list of int64 tensors)
    """
    world = dist.get_world_size(group)
    dev = local_stream.device
    meta = torch.tensor([int(local_len), int(table.shape[0])], dtype=torch.int64, device=dev)
    metas = torch.empty(2 * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(metas, meta, group=group)
    metas = metas.view(world, 2).cpu()
    max_len = int(metas[:, 0].max())
    max_units = int(metas[:, 1].max())
    pad_len = max(16, (max_len + 255) // 256 * 256)
    send = torch.zeros(pad_len, dtype=torch.uint8, device=dev)
    send[:int(local_len)] = local_stream[:int(local_len)]
    recv = torch.empty(pad_len * world, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)   # the one bulk collective of the path
    k = table.shape[1] if table.dim() == 2 else 1
    tsend = torch.zeros((max(1, max_units), k), dtype=torch.int64, device=dev)
    tsend[:table.shape[0]] = table.view(-1, k)
    trecv = torch.empty(world * tsend.numel(), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(trecv, tsend.view(-1), group=group)
    trecv = trecv.view((world,) + tuple(tsend.shape))
    streams = [recv[r * pad_len:r * pad_len + int(metas[r, 0])] for r in range(world)]
    tables = [trecv[r, :int(metas[r, 1])] for r in range(world)]
    return streams, tables


def fold_crc(tables, combine):
    """CRC-32 of the whole input from per-unit rows (crc32, in_len, ...) in rank/unit order.
    `combine(crc_a, crc_b, len_b)` is mz_cuda_crc32_combine (host arithmetic)."""
    crc, total = 0, 0
    for t in tables:
        for row in t.cpu().tolist():
            c, n = int(row[0]) & 0xFFFFFFFF, int(row[1])
            if n == 0:
                continue
            crc = c if total == 0 else combine(crc, c, n)
            total += n
    return crc, total
