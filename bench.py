#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 DEFLATE + CRC-32 backend (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (mz_strm_zlib over zlib 1.3)

Workload (config C5, `configs[4]`, the configuration the metric is quoted on): a 16 GiB synthetic enwik-style
buffer, cut into independent 64 KiB chunks, DEFLATE level 1 + per-chunk CRC-32 + CRC fold + join (K2+K3, K1, K4).
A step = one pass over the whole buffer. With N GPUs the SAME 16 GiB is sharded by contiguous chunk ranges
(strong scaling), each rank compresses its shard, and one NCCL all-gather returns every rank's joined bitstream
and per-chunk {crc, length} table.

value      whole-job GiB/s of uncompressed input, inputs resident in HBM, CUDA-event timed, max over ranks
e2e        same metric through the reference-facing vtbl call (mz_stream_cuda_write/close) with HOST buffers:
           pinned host input, host->device copies, device->host of the stream, base-stream sink, all timed
roofline   dominant kernel (deflate_chunks_kernel): algorithmic bytes (input bytes read per launch) / mean launch time
           (CUDA events on the launching stream) vs the measured HBM copy bandwidth (MEASURED_PEAKS.json)
cpu_baseline the reference path (oracle/_ref: mz_strm_zlib.c + mz_crypt.c + zlib 1.3) on the host cores, bounded sample
"""
import argparse
import ctypes as C
import faulthandler
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GiB = 1 << 30
METRIC = "deflate_l1_crc32_input_throughput"
UNIT = "GiB/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size-gib", type=float, default=16.0, help="total buffer (default: the 16 GiB of the metric)")
    ap.add_argument("--level", type=int, default=1)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-sample-mib", type=int, default=2048)
    ap.add_argument("--sub-batches", type=int, default=0, help="pieces per shard when N>1 (all-gather of piece j overlaps compression of j+1)")
    return ap.parse_args()


# ---- clocks --------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every few milliseconds from a thread
    (a multi-GPU step is tens of milliseconds, too short for `nvidia-smi -lms`); falls back to nvidia-smi when the
    NVML bindings are missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None
        self.thread = None
        self.samples = []
        self.mask = 0
        self.max_mhz = None
        self.stop_flag = False

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)

    def _poll(self, nv, h):
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        try:
            import threading
            nv, h = self._nvml_handle()
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._poll, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread:
            self.stop_flag = True
            self.thread.join(timeout=2)
            sm = sorted(self.samples)
            reasons = sorted(name for bit, name in self.REASONS if self.mask & bit)
            return {"sm_mhz": float(sm[len(sm) // 2]) if sm else None, "sm_max_mhz": float(self.max_mhz) if self.max_mhz else None, "reasons": reasons,
                    "samples": len(sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        except Exception:
            pass
        finally:
            try:
                os.unlink(self.path)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi"}


# ---- the reference's CPU path (oracle/_ref) -------------------------------------------------------------------------
def cpu_reference_throughput(sample, level, threads=None):
    """Time mz_stream_zlib (level `level`, raw window) + mz_crypt_crc32_update over `sample` (bytes) split across host threads.
    Each thread drives the reference's own loop mz_stream_copy_stream_to_end (mz_strm.c:191-206: 16 KiB writes) into a
    reference memory stream. Returns (GiB/s, threads, compressed_bytes)."""
    import refshim
    ref = refshim.RefLib()
    n = len(sample)
    threads = threads or os.cpu_count() or 1
    piece = max(4 << 20, min(32 << 20, (n // (threads * 4)) >> 20 << 20))
    jobs = [(o, min(piece, n - o)) for o in range(0, n, piece)]
    base = C.addressof(sample)
    lock = threading.Lock()
    state = {"next": 0, "comp": 0, "err": 0}

    def work():
        L = ref.lib
        while True:
            with lock:
                i = state["next"]
                state["next"] += 1
            if i >= len(jobs):
                return
            off, ln = jobs[i]
            src = L.mz_stream_mem_create()
            L.mz_stream_mem_set_buffer(src, base + off, ln)
            L.mz_stream_open(src, None, refshim.MZ_OPEN_MODE_READ)
            sink = ref.mem_sink(grow=8 << 20)
            z = L.mz_stream_zlib_create()
            L.mz_stream_set_prop_int64(z, refshim.PROP_COMPRESS_LEVEL, level)
            L.mz_stream_set_base(z, sink)
            ok = L.mz_stream_open(z, None, refshim.MZ_OPEN_MODE_WRITE) == 0
            ok = ok and L.mz_stream_copy_stream_to_end(z, None, src, None) == 0
            L.mz_stream_close(z)
            crc = L.mz_crypt_crc32_update(0, base + off, ln)  # the CRC the zip path computes beside the codec (mz_zip.c:2064)
            tout = ref.get_prop(z, refshim.PROP_TOTAL_OUT)[1]
            ref.delete(z)
            L.mz_stream_close(sink)
            ref.delete(sink)
            ref.delete(src)
            with lock:
                state["comp"] += tout
                if not ok or crc is None:
                    state["err"] += 1

    t0 = time.perf_counter()
    ts = [threading.Thread(target=work) for _ in range(min(threads, len(jobs)))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    assert state["err"] == 0
    return n / GiB / dt, len(ts), state["comp"]


def host_text_sample(nbytes, seed=99):
    """The same generator as the GPU run when a GPU is present; numpy text otherwise (reference arm on a CPU-only box)."""
    try:
        import torch
        if torch.cuda.is_available():
            import __graft_entry__ as ge
            pkg = ge._load_pkg()
            t = pkg.textgen(nbytes, seed=seed)
            torch.cuda.synchronize()
            buf = (C.c_uint8 * nbytes)()
            arr = t.cpu().numpy()  # keep the array alive across the copy
            C.memmove(buf, arr.ctypes.data, nbytes)
            return buf
    except Exception:
        pass
    import datagen
    piece = datagen.text_like(min(nbytes, 8 << 20), seed)
    buf = (C.c_uint8 * nbytes)()
    for o in range(0, nbytes, len(piece)):
        k = min(len(piece), nbytes - o)
        C.memmove(C.addressof(buf) + o, piece, k)
    return buf


def run_reference(args, rank, world):
    if rank != 0:
        return
    import refshim
    if not refshim.ref_available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libmzref.so missing (reference sources were not present at build time)"}))
        return
    nbytes = args.cpu_sample_mib << 20
    sample = host_text_sample(nbytes)
    vals = []
    for i in range(args.warmup + args.steps):
        v, thr, comp = cpu_reference_throughput(sample, args.level)
        if i >= args.warmup:
            vals.append(v)
        if i == 0 and v * 1.0 > 0 and nbytes / GiB / v > 60:  # keep the whole arm within a few minutes
            break
    vals = vals or [v]
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
        "warmup": args.warmup, "ms_per_step": round(1000 * nbytes / GiB / value, 3), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "C5: 16 GiB enwik-style buffer, chunked DEFLATE level %d + CRC-32" % args.level,
                   "reference_path": "mz_strm_zlib.c + mz_crypt.c over system zlib 1.3 (zlib-ng is not vendored / not buildable offline)",
                   "sample": "%d MiB of the same generator per step, <=32 MiB per stream, all host threads" % args.cpu_sample_mib},
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": thr, "kind": "reference",
                         "sample": "%d MiB per step, one independent mz_stream_zlib per <=32 MiB piece, raw window, level %d, + mz_crypt_crc32_update" % (args.cpu_sample_mib, args.level)},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ratio": round(comp / nbytes, 4),
    }
    print(json.dumps(line), flush=True)


# ---- our arm ---------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    pkg = ge._load_pkg()
    lib = pkg.load()
    torch.cuda.set_device(local_rank)
    pkg.check(lib.mz_cuda_init(), "mz_cuda_init")
    dev = torch.device("cuda", local_rank)
    total = int(args.size_gib * GiB) // 65536 * 65536
    nchunks_total = total // 65536
    c0 = rank * nchunks_total // world
    c1 = (rank + 1) * nchunks_total // world
    shard = (c1 - c0) * 65536
    # ---- input: this rank's shard of the one global buffer (generator is position-independent) -----------------------------
    src = torch.empty(shard, dtype=torch.uint8, device=dev)
    seg = 64 << 20
    for o in range(0, shard, seg):
        k = min(seg, shard - o)
        pkg.check(lib.mz_cuda_textgen(src.data_ptr() + o, k, 1000 + (c0 * 65536 + o) // seg, None), "textgen")
    torch.cuda.synchronize()
    # ---- sub-batches: with N>1 the shard is compressed in NB pieces so the all-gather of piece j overlaps the compression of j+1
    NB = 1 if world == 1 else (args.sub_batches if args.sub_batches > 0 else (8 if world >= 8 else 4))  # measured: 4 pieces best at N=2/4, 8 at N=8
    nshard_chunks = c1 - c0
    bounds = [nshard_chunks * j // NB for j in range(NB + 1)]
    subs = [(bounds[j] * 65536, (bounds[j + 1] - bounds[j]) * 65536) for j in range(NB)]  # (byte offset in shard, bytes)
    batches = [pkg.DeflateBatch(max(nb, 65536)) for _, nb in subs]
    stream = torch.cuda.current_stream()
    comm_stream = torch.cuda.Stream() if world > 1 else None
    ev_done = [torch.cuda.Event() for _ in range(NB)]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    state = {"cap": None, "gathered": [None] * NB, "tbl_all": None}

    def compress_sub(j, kev=None):
        off, nb = subs[j]
        b = batches[j]
        n = b.nchunks(nb)
        s = pkg._stream_ptr()
        final = pkg.FLAG_FINAL if (rank == world - 1 and j == NB - 1) else 0  # only the globally last chunk carries BFINAL
        base = src.data_ptr() + off
        if kev is not None:
            kev[0].record(stream)
        pkg.check(lib.mz_cuda_deflate_chunks(base, nb, 65536, None, None, None, n, final, args.level, b.slots.data_ptr(), b.stride,
                                             b.out_len.data_ptr(), s), "deflate")
        if kev is not None:
            kev[1].record(stream)
        pkg.check(lib.mz_cuda_crc32_segments(base, nb, 65536, None, None, n, b.residue.data_ptr(), b.chunk_crc.data_ptr(), s), "crc")
        pkg.check(lib.mz_cuda_crc32_fold(b.residue.data_ptr(), n, 65536, nb, b.crc_out.data_ptr(), s), "fold")
        pkg.check(lib.mz_cuda_concat(b.slots.data_ptr(), b.stride, b.out_len.data_ptr(), n, b.offsets.data_ptr(), b.joined.data_ptr(), s), "concat")
        return 5

    def step(kevs=None):
        launches = 0
        for j in range(NB):
            launches += compress_sub(j, kevs[j] if kevs else None)
            if world > 1 and state["cap"] is not None:
                ev_done[j].record(stream)
                with torch.cuda.stream(comm_stream):
                    comm_stream.wait_event(ev_done[j])
                    # THE collective of the path: every rank's joined bitstream of piece j (fixed capacity, sized in warm-up)
                    dist.all_gather_into_tensor(state["gathered"][j], batches[j].joined[:state["cap"]])
                launches += 1
        if world > 1 and state["cap"] is not None:
            with torch.cuda.stream(comm_stream):
                comm_stream.wait_event(ev_done[NB - 1])
                tbl = torch.cat([torch.stack([b.chunk_crc[:b.nchunks(nb)], b.out_len[:b.nchunks(nb)], b.offsets[:b.nchunks(nb)].to(torch.int32)], 1)
                                 for b, (_, nb) in zip(batches, subs)]).contiguous()
                if state["tbl_all"] is None:
                    state["tbl_all"] = torch.empty((world * tbl.shape[0], 3), dtype=torch.int32, device=dev)
                dist.all_gather_into_tensor(state["tbl_all"], tbl)  # per-chunk {crc32, out_len, offset} table
            stream.wait_stream(comm_stream)
            launches += 1
        return launches

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def piece_totals():
        return [int(b.offsets[b.nchunks(nb)].item()) for b, (_, nb) in zip(batches, subs)]

    # warm-up: first pass sizes the gather slabs (max piece over all ranks + margin), the rest run the full pipeline
    step()
    barrier()
    if world > 1:
        mx = torch.tensor([max(piece_totals())], dtype=torch.int64, device=dev)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        state["cap"] = (int(mx.item()) * 51 // 50 + 4096 + 255) // 256 * 256
        for j in range(NB):
            state["gathered"][j] = torch.empty(state["cap"] * world, dtype=torch.uint8, device=dev)
    for i in range(max(0, args.warmup - 1) + (1 if world > 1 else 0)):
        step()
    barrier()
    totals = piece_totals()
    comp_bytes = sum(totals)
    # CRC of the shard = fold of the pieces' CRCs (host arithmetic); rank 0 reports its own shard's
    crc_whole = 0
    for j, (b, (_, nb)) in enumerate(zip(batches, subs)):
        c = int(b.crc_out[1].item()) & 0xFFFFFFFF
        crc_whole = c if j == 0 else lib.mz_cuda_crc32_combine(crc_whole, c, nb)
    clocks = ClockSampler(local_rank)
    clocks.start()
    kev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(NB)] for _ in range(args.steps)]
    barrier()
    ev[0].record(stream)
    launches = 0
    for i in range(args.steps):
        launches += step(kev[i])
    ev[1].record(stream)
    barrier()
    ms = ev[0].elapsed_time(ev[1])
    kms = sum(a.elapsed_time(b) for row in kev for a, b in row) / max(1, args.steps)
    clk = clocks.stop()
    if world > 1:
        assert max(piece_totals()) <= state["cap"], "gather slab too small"
        # the gathered slabs really hold every rank's stream: spot-check piece 0 of the next rank against the table
        nxt = (rank + 1) % world
        rows = state["tbl_all"].view(world, -1, 3)[nxt]
        assert int(rows[0, 2]) == 0 and int(rows[0, 1]) > 0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = (total / GiB) * args.steps / (ms / 1000.0)

    # ---- e2e through the vtbl with host buffers ----------------------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        import cuharness
        tl = cuharness.TestLib()
        hsrc = torch.empty(shard, dtype=torch.uint8, pin_memory=True)
        hsrc.copy_(src)
        torch.cuda.synchronize()
        sink_cap = shard * 9 // 16 + (64 << 20)  # level-1 text lands near 0.503 of the input; leave room for other data
        hsink = torch.empty(sink_cap, dtype=torch.uint8, pin_memory=True)
        times, times_discard = [], []
        out_bytes = 0
        # the last pass repeats the measurement with a sink that drops the bytes: what is left is upload + kernels + download
        # into the stream's pinned staging, i.e. the codec without the consumer's single-threaded memcpy
        SINK_THREADS = 8
        times_single = []
        for i in range(args.e2e_steps + 3):
            single = i == args.e2e_steps + 1
            discard = i == args.e2e_steps + 2
            sink = tl.lib.mz_stream_mem64_create()
            tl.lib.mz_stream_mem64_set_sink(sink, hsink.data_ptr(), sink_cap)
            tl.lib.mz_stream_mem64_set_copy_threads(sink, 1 if single else SINK_THREADS)
            if discard:
                tl.lib.mz_stream_mem64_set_discard(sink, 1)
            s = lib.mz_stream_cuda_create()
            lib.mz_stream_cuda_set_prop_int64(s, pkg.MZ_STREAM_PROP_COMPRESS_LEVEL, args.level)
            tl.lib.mzt_set_base(s, sink)
            barrier()
            t0 = time.perf_counter()
            assert lib.mz_stream_cuda_open(s, None, pkg.MZ_OPEN_MODE_WRITE) == 0
            wrote = tl.lib.mzt_write_all(s, hsrc.data_ptr(), shard, 1 << 30)
            cerr = lib.mz_stream_cuda_close(s)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            assert wrote == shard and cerr == 0, (wrote, cerr)
            out_bytes = tl.lib.mzt_tell(sink)
            ps = C.c_void_p(s)
            lib.mz_stream_cuda_delete(C.byref(ps))
            tl.delete(sink)
            if discard:
                times_discard.append(dt)
            elif single:
                times_single.append(dt)
            elif i > 0:
                times.append(dt)
            if not discard:
                out_bytes_kept = out_bytes
        out_bytes = out_bytes_kept
        tt = torch.tensor([sum(times) / len(times), times_discard[0], times_single[0]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": round((total / GiB) / float(tt[0].item()), 4), "unit": UNIT, "h2d_bytes_per_step": shard, "d2h_bytes_per_step": int(out_bytes),
               "steps": len(times), "api": "mz_stream_cuda_open/write(1 GiB calls)/close over a 64-bit memory base stream, pinned host input",
               "value_single_thread_sink": round((total / GiB) / float(tt[2].item()), 4),
               "value_discarding_sink": round((total / GiB) / float(tt[1].item()), 4),
               "sink": "host memory stream (tests/support/mem64.c) copying every compressed byte with %d threads" % SINK_THREADS,
               "note": "value_single_thread_sink: the same sink with one memcpy thread (the consumer, not the codec, is then the limit); "
                       "value_discarding_sink: the base stream drops the bytes (upload + kernels + download into pinned staging only)"}
        del hsrc, hsink

    if rank != 0:
        return
    # ---- roofline of the dominant kernel -----------------------------------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = shard / (kms / 1000.0) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r1_deflate_kernel.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_input_byte")
            traffic = None if traffic is None else round(traffic * shard / NB)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "deflate_chunks_kernel", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 5), "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": shard // NB, "ms_per_launch": round(kms / NB, 4), "launches_per_step": NB,
                "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum per input byte x bytes per launch (profiles/r1_deflate_kernel.json)"}
    cpu = None
    if not args.no_cpu and world == 1:  # reported baseline: rank 0 at N=1 only
        import refshim
        if refshim.ref_available():
            nb = min(args.cpu_sample_mib << 20, shard)
            sample = (C.c_uint8 * nb)()
            arr = src[:nb].cpu().numpy()  # keep the array alive across the copy
            C.memmove(sample, arr.ctypes.data, nb)
            del arr
            v, thr, comp = cpu_reference_throughput(sample, args.level)
            cpu = {"value": round(v, 4), "unit": UNIT, "cores": thr, "kind": "reference",
                   "sample": "first %d MiB of the same buffer, mz_strm_zlib over zlib 1.3 level %d, one stream per <=32 MiB piece, all host threads; ratio %.4f" % (nb >> 20, args.level, comp / nb)}
        else:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": "C5: %.2f GiB enwik-style buffer, independent 64 KiB chunks, DEFLATE level %d + CRC-32 per chunk + fold + join" % (total / GiB, args.level),
                   "chunk_bytes": 65536, "level": args.level, "l2": "inputs (%.1f GiB per GPU) are far larger than L2; no flush needed" % (shard / GiB),
                   "parallelism": ("chunk-sharded x%d; per step one NCCL all-gather of the bitstreams (in %d pieces overlapped with compression) + one of the per-chunk {crc,len,offset} table" % (world, NB)) if world > 1 else "single GPU"},
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clk, "e2e": e2e, "gpu_launches": launches,
        "ratio": round(comp_bytes / shard, 4), "crc32": "%08x" % crc_whole,
    }
    print(json.dumps(line), flush=True)


def main():
    faulthandler.enable()
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
