/* inflate_kernel.cuh -- K5: RFC1951 decode, one warp per independent stream (sm_100a).
 *
 * Replaces what zlib's inflate() does behind mz_stream_zlib_read (mz_strm_zlib.c:116-193): per block,
 * Huffman table build + symbol decode + LZ77 back-reference copy, with exact consumed-bytes accounting
 * (TOTAL_IN semantics, mz_strm_zlib.c:168-175) and zlib's error taxonomy (data error / truncated).
 *
 * Round-1 shape: block starts inside one foreign stream are only known by decoding, so a stream is
 * walked serially; parallelism comes from (a) many independent streams (zip entries) = many warps and
 * (b) inside a warp: table construction and match copies are warp-cooperative, lane 0 runs the symbol
 * loop. The decoder is RESUMABLE at symbol granularity through InflateState, so the host can feed a
 * long stream through a bounded device window (input refills, output drains) -- the read path of the
 * vtbl stream.
 */
#ifndef MZ_INFLATE_KERNEL_CUH
#define MZ_INFLATE_KERNEL_CUH

#include "mzcuda_common.cuh"

namespace mzc {

constexpr int INF_PB = 11;  /* primary lit/len table bits */
constexpr int INF_DB = 9;   /* primary distance table bits */
constexpr int INF_THREADS = 32;
constexpr uint32_t INF_IN_RING = 4096;    /* compressed-input window in shared memory (bytes, power of two) */
constexpr uint32_t INF_OUT_RING = 65536;  /* output ring in shared memory: 32 KiB of LZ77 history + unflushed output */
constexpr uint32_t INF_FLUSH = 16384;     /* flush the ring to global memory when this much output is pending */

enum { INF_ST_RUN = 0, INF_ST_END = 1, INF_ST_DATA_ERROR = -3, INF_ST_BUF_ERROR = -5 };
enum { INF_PH_HEADER = 0, INF_PH_CODES = 1, INF_PH_STORED = 2 };
enum { INF_WHY_NONE = 0, INF_WHY_INPUT = 1, INF_WHY_OUTPUT = 2 };

struct InflateState {
    uint64_t in_bitpos;  /* absolute bit position in the raw deflate stream */
    uint64_t out_pos;    /* absolute bytes produced */
    int32_t status;      /* INF_ST_* */
    int32_t why;         /* INF_WHY_* when status == RUN */
    uint32_t phase;
    uint32_t last_block;
    uint32_t stored_remaining;
    uint32_t nlit, ndist;
    uint32_t blocks;     /* blocks completed */
    uint8_t lens[320];   /* current block's code lengths (lit/len then dist) */
};

struct InflateJob {
    const uint8_t *in;   /* device bytes; in[0] is absolute stream byte in_base; padded >= 16 bytes past in_avail */
    uint64_t in_base;
    uint64_t in_avail;   /* valid bytes at `in` */
    uint8_t *out;        /* out[0] is absolute output byte out_base; must hold 32 KiB of history if out_pos > 0 */
    uint64_t out_base;
    uint64_t out_cap;    /* bytes available at `out` */
    uint32_t in_final;   /* no more input will follow */
    uint32_t pad;
};

struct InfTables {
    uint16_t lit[1 << INF_PB];  /* sym | len << 9, 0 = long code */
    uint16_t dist[1 << INF_DB];
    uint16_t lsym[288];         /* symbols sorted by (len, sym) */
    uint16_t dsym[32];
    uint16_t lcount[16], dcount[16];
    uint32_t scratch[32];
    uint8_t lens[384];
};

__device__ __forceinline__ uint32_t inf_len_base(uint32_t s, uint32_t &eb) { /* s = sym - 257 in 0..28 */
    if (s < 8) { eb = 0; return 3 + s; }
    if (s == 28) { eb = 0; return 258; }
    eb = (s >> 2) - 1;
    return 3 + ((4 + (s & 3)) << eb);
}
__device__ __forceinline__ uint32_t inf_dist_base(uint32_t s, uint32_t &eb) { /* s in 0..29 */
    if (s < 4) { eb = 0; return 1 + s; }
    eb = (s >> 1) - 1;
    return 1 + ((2 + (s & 1)) << eb);
}

/* Build decode structures for one alphabet from code lengths. Warp-cooperative.
 * Returns <0 over-subscribed, >0 incomplete (unused code space), 0 complete. */
__device__ inline int inf_build(const uint8_t *lens, int n, uint16_t *count, uint16_t *sorted, uint16_t *table, int tbits,
                                uint32_t *scratch) {
    const unsigned lane = lane_id();
    if (lane < 16) scratch[lane] = 0;
    for (int i = (int)lane; i < (1 << tbits); i += 32) table[i] = 0;
    __syncwarp();
    for (int i = (int)lane; i < n; i += 32) atomicAdd(&scratch[lens[i]], 1u);
    __syncwarp();
    uint32_t cnt = lane < 16 ? scratch[lane] : 0;
    __syncwarp();
    if (lane < 16) count[lane] = (uint16_t)cnt;
    /* Kraft check (uniform across the warp) */
    int left = 1;
    for (int len = 1; len <= 15; len++) {
        uint32_t c = __shfl_sync(MZ_FULL_MASK, cnt, len);
        left = (left << 1) - (int)c;
        if (left < 0) break;
    }
    if (left < 0) return left;
    /* lane = code length: first insert index and first canonical code of that length */
    uint32_t offs = 0, code = 0;
    if (lane >= 1 && lane < 16) {
        for (unsigned b = 1; b < lane; b++) offs += scratch[b];
        for (unsigned b = 1; b <= lane; b++) code = (code + (b == 1 ? 0u : scratch[b - 1])) << 1;
    }
    __syncwarp();
    if (lane < 16) {
        scratch[lane] = offs;       /* running insert index per length */
        scratch[16 + lane] = code;  /* running canonical code per length */
    }
    __syncwarp();
    for (int base = 0; base < n; base += 32) {
        int i = base + (int)lane;
        uint32_t l = i < n ? lens[i] : 0u;
        unsigned m = __match_any_sync(MZ_FULL_MASK, l);
        uint32_t rank = (uint32_t)__popc(m & ((1u << lane) - 1));
        if (l) {
            sorted[scratch[l] + rank] = (uint16_t)i;
            if ((int)l <= tbits) {
                uint32_t c = scratch[16 + l] + rank;
                uint32_t r = __brev(c) >> (32 - l);
                uint16_t ent = (uint16_t)((uint32_t)i | (l << 9));
                for (uint32_t e = r; e < (1u << tbits); e += 1u << l) table[e] = ent;
            }
        }
        __syncwarp();
        if (l && rank == 0) {
            uint32_t k = (uint32_t)__popc(m);
            scratch[l] += k;
            scratch[16 + l] += k;
        }
        __syncwarp();
    }
    return left;
}

struct InfBits { /* lane 0 only; reads the compressed bytes from the shared-memory input ring */
    const uint8_t *ring;
    uint64_t pos;    /* next byte index (relative to job.in) to load into the bit buffer */
    uint64_t avail;  /* valid bytes of the stream window (job.in_avail) */
    uint64_t bb;
    uint32_t bc;
    __device__ __forceinline__ void init(const uint8_t *r, uint64_t avail_bytes, uint64_t rel_bitpos) {
        ring = r; avail = avail_bytes; pos = rel_bitpos >> 3; bb = 0; bc = 0;
        refill();
        uint32_t skip = (uint32_t)(rel_bitpos & 7);
        bb >>= skip; bc -= skip;
    }
    __device__ __forceinline__ void refill() { /* keep >= 32 bits; the ring holds zero padding past avail */
        while (bc <= 32) {
            const uint32_t i = (uint32_t)pos;
            uint32_t w = (uint32_t)ring[i & (INF_IN_RING - 1)] | ((uint32_t)ring[(i + 1) & (INF_IN_RING - 1)] << 8) |
                         ((uint32_t)ring[(i + 2) & (INF_IN_RING - 1)] << 16) | ((uint32_t)ring[(i + 3) & (INF_IN_RING - 1)] << 24);
            bb |= (uint64_t)w << bc;
            pos += 4; bc += 32;
        }
    }
    __device__ __forceinline__ uint32_t peek(uint32_t n) const { return (uint32_t)(bb & ((1ull << n) - 1)); }
    __device__ __forceinline__ void drop(uint32_t n) { bb >>= n; bc -= n; }
    __device__ __forceinline__ uint32_t get(uint32_t n) { uint32_t v = peek(n); drop(n); refill(); return v; }
    /* bits consumed so far, relative to in[0] */
    __device__ __forceinline__ uint64_t bitpos() const { return pos * 8 - bc; }
    __device__ __forceinline__ int64_t bits_left() const { return (int64_t)(avail * 8) - (int64_t)bitpos(); }
};

/* Warp-cooperative: bring compressed bytes [loaded, ...) into the input ring, keeping everything from
 * the byte the bit reader may still need (`keep_from`). Reads up to 16 bytes of the job's zero padding. */
__device__ __forceinline__ uint64_t inf_fill_input(uint8_t *ring, const uint8_t *in, uint64_t keep_from, uint64_t loaded, uint64_t padded_total) {
    uint64_t target = (keep_from & ~15ull) + INF_IN_RING;
    if (target > padded_total) target = padded_total;
    for (uint64_t a = loaded + lane_id(); a < target; a += 32) ring[(uint32_t)a & (INF_IN_RING - 1)] = in[a];
    __syncwarp();
    return target > loaded ? target : loaded;
}

/* Warp-cooperative: copy finished output [from, to) out of the shared ring into global memory (coalesced bytes) */
__device__ __forceinline__ void inf_flush_output(const uint8_t *oring, uint8_t *out, uint64_t from, uint64_t to) {
    for (uint64_t a = from + lane_id(); a < to; a += 32) out[a] = oring[(uint32_t)a & (INF_OUT_RING - 1)];
    __syncwarp();
}

/* canonical bit-by-bit decode for codes longer than the primary table (lane 0) */
__device__ inline int inf_slow_decode(InfBits &b, const uint16_t *count, const uint16_t *sorted) {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; len++) {
        code |= (int)((b.bb >> (len - 1)) & 1);
        int c = count[len];
        if (code - c < first) {
            b.drop((uint32_t)len);
            b.refill();
            return sorted[index + (code - first)];
        }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

/* warp-cooperative copy of a match inside the shared output ring; all source bytes of one round are already written */
__device__ __forceinline__ void inf_copy_match(uint8_t *oring, uint64_t dst, uint32_t len, uint32_t dist) {
    const unsigned lane = lane_id();
    const uint32_t d0 = (uint32_t)dst, m = INF_OUT_RING - 1;
    if (dist < 32) {
        for (uint32_t i = lane; i < len; i += 32) oring[(d0 + i) & m] = oring[(d0 - dist + (i % dist)) & m];
    } else {
        uint32_t round = dist & ~31u;
        for (uint32_t done = 0; done < len; done += round) {
            uint32_t n = len - done < round ? len - done : round;
            for (uint32_t i = lane; i < n; i += 32) oring[(d0 + done + i) & m] = oring[(d0 + done + i - dist) & m];
            __syncwarp();
        }
    }
    __syncwarp();
}

constexpr int INF_SMEM_BYTES = ((int)sizeof(InfTables) + 15) / 16 * 16 + (int)INF_IN_RING + (int)INF_OUT_RING;

__global__ void __launch_bounds__(INF_THREADS) inflate_streams_kernel(const InflateJob *jobs, InflateState *states, uint32_t nstreams) {
    MZ_DYN_SMEM(smem);
    InfTables &T = *reinterpret_cast<InfTables *>(smem);
    uint8_t *iring = smem + (sizeof(InfTables) + 15) / 16 * 16;
    uint8_t *oring = iring + INF_IN_RING;
    const unsigned lane = lane_id();
    for (uint32_t sidx = blockIdx.x; sidx < nstreams; sidx += gridDim.x) {
        const InflateJob job = jobs[sidx];
        InflateState *st = &states[sidx];
        if (st->status != INF_ST_RUN) continue;
        InfBits b;
        uint64_t out_pos = st->out_pos;
        uint32_t phase = st->phase, last = st->last_block, stored_rem = st->stored_remaining;
        uint32_t nlit = st->nlit, ndist = st->ndist, blocks = st->blocks;
        int status = INF_ST_RUN, why = INF_WHY_NONE;
        const uint64_t out_end = job.out_base + job.out_cap;
        uint8_t *out = job.out - job.out_base; /* index with absolute positions */
        const uint64_t padded_total = job.in_avail + 16; /* the job guarantees 16 readable zero bytes past in_avail */
        const uint64_t start_bit = st->in_bitpos - job.in_base * 8;
        /* compressed-input window */
        uint64_t loaded = (start_bit >> 3) & ~15ull;
        loaded = inf_fill_input(iring, job.in, start_bit >> 3, loaded, padded_total);
        /* output ring: bring back the LZ77 history (up to 32 KiB already in global memory) */
        uint64_t flushed = out_pos;
        {
            uint64_t hist = out_pos - job.out_base < 32768 ? out_pos - job.out_base : 32768;
            for (uint64_t a = out_pos - hist + lane; a < out_pos; a += 32) oring[(uint32_t)a & (INF_OUT_RING - 1)] = out[a];
            __syncwarp();
        }
        if (lane == 0) b.init(iring, job.in_avail, start_bit);
        if (phase == INF_PH_CODES) { /* resume inside a Huffman block: rebuild the tables */
            for (int i = (int)lane; i < 320; i += 32) T.lens[i] = st->lens[i];
            __syncwarp();
            inf_build(T.lens, (int)nlit, T.lcount, T.lsym, T.lit, INF_PB, T.scratch);
            inf_build(T.lens + nlit, (int)ndist, T.dcount, T.dsym, T.dist, INF_DB, T.scratch);
            __syncwarp();
        }
        while (status == INF_ST_RUN && why == INF_WHY_NONE) {
            if (phase == INF_PH_HEADER) {
                /* the header is parsed by lane 0 alone: make sure the window holds it (<= ~600 bytes) */
                {
                    uint64_t bp = 0;
                    if (lane == 0) bp = b.bitpos() >> 3;
                    bp = __shfl_sync(MZ_FULL_MASK, bp, 0);
                    if (loaded < padded_total && loaded - bp < 2048) loaded = inf_fill_input(iring, job.in, bp, loaded, padded_total);
                }
                /* a dynamic header is at most 14 + 57 + 320*(7+7) bits ~ 570 bytes: wait for it unless final */
                int ok = 1, err = 0;
                uint32_t type = 0;
                if (lane == 0) {
                    if (!job.in_final && b.bits_left() < 8 * 1024) ok = 0;
                    else if (b.bits_left() < 3) err = INF_ST_BUF_ERROR;
                    else {
                        last = b.get(1);
                        type = b.get(2);
                    }
                }
                ok = __shfl_sync(MZ_FULL_MASK, ok, 0);
                err = __shfl_sync(MZ_FULL_MASK, err, 0);
                if (!ok) { why = INF_WHY_INPUT; break; }
                if (err) { status = err; break; }
                type = __shfl_sync(MZ_FULL_MASK, type, 0);
                last = __shfl_sync(MZ_FULL_MASK, last, 0);
                if (type == 0) {
                    uint32_t len = 0;
                    if (lane == 0) {
                        b.drop(b.bc & 7); /* to byte boundary */
                        b.refill();
                        if (b.bits_left() < 32) err = INF_ST_BUF_ERROR;
                        else {
                            len = b.get(16);
                            uint32_t nlen = b.get(16);
                            if ((len ^ 0xffffu) != nlen) err = INF_ST_DATA_ERROR;
                        }
                    }
                    err = __shfl_sync(MZ_FULL_MASK, err, 0);
                    if (err) { status = err; break; }
                    stored_rem = __shfl_sync(MZ_FULL_MASK, len, 0);
                    phase = INF_PH_STORED;
                } else if (type == 1) {
                    for (int i = (int)lane; i < 288; i += 32) T.lens[i] = i < 144 ? 8 : (i < 256 ? 9 : (i < 280 ? 7 : 8));
                    if (lane < 30) T.lens[288 + lane] = 5;
                    nlit = 288; ndist = 30;
                    __syncwarp();
                    inf_build(T.lens, 288, T.lcount, T.lsym, T.lit, INF_PB, T.scratch);
                    inf_build(T.lens + 288, 30, T.dcount, T.dsym, T.dist, INF_DB, T.scratch);
                    __syncwarp();
                    phase = INF_PH_CODES;
                } else if (type == 2) {
                    /* code-length code, then the literal/length and distance lengths (lane 0 reads bits) */
                    uint32_t hl = 0, hd = 0, hc = 0;
                    if (lane == 0) {
                        hl = b.get(5) + 257; hd = b.get(5) + 1; hc = b.get(4) + 4;
                        if (hl > 286 || hd > 30) err = INF_ST_DATA_ERROR;
                    }
                    err = __shfl_sync(MZ_FULL_MASK, err, 0);
                    if (err) { status = err; break; }
                    nlit = __shfl_sync(MZ_FULL_MASK, hl, 0);
                    ndist = __shfl_sync(MZ_FULL_MASK, hd, 0);
                    hc = __shfl_sync(MZ_FULL_MASK, hc, 0);
                    if (lane < 19) T.lens[lane] = 0;
                    __syncwarp();
                    if (lane == 0) {
                        for (uint32_t i = 0; i < hc; i++) {
                            uint32_t pos;
                            switch (i) {
                                case 0: pos = 16; break; case 1: pos = 17; break; case 2: pos = 18; break; case 3: pos = 0; break;
                                case 4: pos = 8; break; case 5: pos = 7; break; case 6: pos = 9; break; case 7: pos = 6; break;
                                case 8: pos = 10; break; case 9: pos = 5; break; case 10: pos = 11; break; case 11: pos = 4; break;
                                case 12: pos = 12; break; case 13: pos = 3; break; case 14: pos = 13; break; case 15: pos = 2; break;
                                case 16: pos = 14; break; case 17: pos = 1; break; default: pos = 15; break;
                            }
                            T.lens[pos] = (uint8_t)b.get(3);
                        }
                    }
                    __syncwarp();
                    int left = inf_build(T.lens, 19, T.lcount, T.lsym, T.lit, 7, T.scratch);
                    __syncwarp();
                    if (left != 0) { status = INF_ST_DATA_ERROR; break; } /* code-length code must be complete */
                    if (lane == 0) {
                        uint32_t idx = 0, total = nlit + ndist;
                        uint8_t *lens = T.lens + 32; /* decoded lengths, staged after the 19 cl lengths */
                        while (idx < total && !err) {
                            uint32_t e = T.lit[b.peek(7)];
                            if (e == 0) { err = INF_ST_DATA_ERROR; break; }
                            b.drop(e >> 9);
                            b.refill();
                            uint32_t sym = e & 511;
                            if (sym < 16) {
                                lens[idx++] = (uint8_t)sym;
                            } else {
                                uint32_t rep, val = 0;
                                if (sym == 16) {
                                    if (idx == 0) { err = INF_ST_DATA_ERROR; break; }
                                    val = lens[idx - 1];
                                    rep = 3 + b.get(2);
                                } else if (sym == 17) rep = 3 + b.get(3);
                                else rep = 11 + b.get(7);
                                if (idx + rep > total) { err = INF_ST_DATA_ERROR; break; }
                                while (rep--) lens[idx++] = (uint8_t)val;
                            }
                        }
                        if (!err && b.bits_left() < 0) err = INF_ST_BUF_ERROR;
                        if (!err && lens[256] == 0) err = INF_ST_DATA_ERROR; /* no end-of-block code */
                    }
                    err = __shfl_sync(MZ_FULL_MASK, err, 0);
                    if (err) { status = err; break; }
                    __syncwarp();
                    /* move into place: lit/len lengths at T.lens[0..], distance lengths right after */
                    uint8_t tmp[10];
                    for (int k = 0; k < 10; k++) {
                        uint32_t i = lane + 32 * k;
                        tmp[k] = i < nlit + ndist ? T.lens[32 + i] : 0;
                    }
                    __syncwarp();
                    for (int k = 0; k < 10; k++) {
                        uint32_t i = lane + 32 * k;
                        if (i < 320) T.lens[i] = tmp[k];
                    }
                    __syncwarp();
                    int l1 = inf_build(T.lens, (int)nlit, T.lcount, T.lsym, T.lit, INF_PB, T.scratch);
                    __syncwarp();
                    int l2 = inf_build(T.lens + nlit, (int)ndist, T.dcount, T.dsym, T.dist, INF_DB, T.scratch);
                    __syncwarp();
                    /* zlib: incomplete sets are only allowed when there is a single code (of length 1) */
                    uint32_t usedl = nlit - T.lcount[0], usedd = ndist - T.dcount[0];
                    if (l1 < 0 || l2 < 0 || (l1 > 0 && !(usedl == 1 && T.lcount[1] == 1)) || (l2 > 0 && usedd != 0 && !(usedd == 1 && T.dcount[1] == 1))) {
                        status = INF_ST_DATA_ERROR;
                        break;
                    }
                    phase = INF_PH_CODES;
                } else {
                    status = INF_ST_DATA_ERROR;
                    break;
                }
            } else if (phase == INF_PH_STORED) {
                /* raw copy, bounded by the input and output windows; the bytes go to global memory directly and
                 * into the ring (they are history for later matches) */
                inf_flush_output(oring, out, flushed, out_pos);
                uint64_t ib = 0;
                if (lane == 0) ib = b.bitpos() >> 3; /* byte aligned here */
                ib = __shfl_sync(MZ_FULL_MASK, ib, 0);
                uint64_t in_left = job.in_avail > ib ? job.in_avail - ib : 0;
                uint64_t out_left = out_end - out_pos;
                uint32_t n = stored_rem;
                if (n > in_left) n = (uint32_t)in_left;
                if (n > out_left) n = (uint32_t)out_left;
                for (uint32_t i = lane; i < n; i += 32) {
                    uint8_t v = job.in[ib + i];
                    out[out_pos + i] = v;
                    if (n - i <= 32768) oring[(uint32_t)(out_pos + i) & (INF_OUT_RING - 1)] = v; /* only the last 32 KiB can matter */
                }
                __syncwarp();
                out_pos += n;
                flushed = out_pos;
                stored_rem -= n;
                loaded = ((ib + n) & ~15ull);
                loaded = inf_fill_input(iring, job.in, ib + n, loaded, padded_total);
                if (lane == 0) b.init(iring, job.in_avail, (ib + n) * 8);
                if (stored_rem == 0) {
                    phase = INF_PH_HEADER;
                    blocks++;
                    if (last) status = INF_ST_END;
                } else if (out_pos >= out_end) {
                    why = INF_WHY_OUTPUT;
                } else if (job.in_final) {
                    status = INF_ST_BUF_ERROR;
                } else {
                    why = INF_WHY_INPUT;
                }
            } else { /* INF_PH_CODES */
                /* ev: 1 match, 2 end of block, 3 need input, 4 need output, 5 refill the input window, 6 flush the ring, <0 error */
                uint32_t ev = 0, mlen = 0, mdist = 0;
                if (lane == 0) {
                    for (;;) {
                        if (b.bits_left() < 0) { ev = (uint32_t)INF_ST_BUF_ERROR; break; } /* truncated */
                        if (!job.in_final && b.bits_left() < 64) { ev = 3; break; }
                        if (b.pos + 16 > loaded && loaded < padded_total) { ev = 5; break; }
                        if (out_pos - flushed >= INF_FLUSH) { ev = 6; break; }
                        /* near the end of the output window: remember the reader so a symbol that does
                         * not fit can be un-read */
                        const bool tight = out_pos + 258 > out_end;
                        InfBits saved = b;
                        uint32_t e = T.lit[b.peek(INF_PB)];
                        int sym;
                        if (e) {
                            b.drop(e >> 9);
                            b.refill();
                            sym = (int)(e & 511);
                        } else {
                            sym = inf_slow_decode(b, T.lcount, T.lsym);
                            if (sym < 0) { ev = (uint32_t)INF_ST_DATA_ERROR; break; }
                        }
                        if (sym < 256) {
                            if (tight && out_pos >= out_end) { b = saved; ev = 4; break; }
                            oring[(uint32_t)out_pos & (INF_OUT_RING - 1)] = (uint8_t)sym;
                            out_pos++;
                            continue;
                        }
                        if (sym == 256) { ev = 2; break; }
                        if (sym > 285) { ev = (uint32_t)INF_ST_DATA_ERROR; break; }
                        uint32_t eb;
                        mlen = inf_len_base((uint32_t)sym - 257, eb);
                        mlen += b.get(eb);
                        uint32_t de = T.dist[b.peek(INF_DB)];
                        int ds;
                        if (de) {
                            b.drop(de >> 9);
                            b.refill();
                            ds = (int)(de & 511);
                        } else {
                            ds = inf_slow_decode(b, T.dcount, T.dsym);
                            if (ds < 0) { ev = (uint32_t)INF_ST_DATA_ERROR; break; }
                        }
                        if (ds > 29) { ev = (uint32_t)INF_ST_DATA_ERROR; break; }
                        mdist = inf_dist_base((uint32_t)ds, eb);
                        mdist += b.get(eb);
                        if (mdist > out_pos) { ev = (uint32_t)INF_ST_DATA_ERROR; break; } /* too far back */
                        if (tight && out_pos + mlen > out_end) { b = saved; ev = 4; break; }
                        ev = 1;
                        break;
                    }
                    if ((ev == 1 || ev == 2) && b.bits_left() < 0) ev = (uint32_t)INF_ST_BUF_ERROR; /* ran past the end */
                }
                ev = __shfl_sync(MZ_FULL_MASK, ev, 0);
                out_pos = __shfl_sync(MZ_FULL_MASK, out_pos, 0);
                __syncwarp();
                if (ev == 1) {
                    mlen = __shfl_sync(MZ_FULL_MASK, mlen, 0);
                    mdist = __shfl_sync(MZ_FULL_MASK, mdist, 0);
                    inf_copy_match(oring, out_pos, mlen, mdist);
                    out_pos += mlen;
                } else if (ev == 2) {
                    phase = INF_PH_HEADER;
                    blocks++;
                    if (last) status = INF_ST_END;
                } else if (ev == 3) {
                    why = INF_WHY_INPUT;
                } else if (ev == 4) {
                    why = INF_WHY_OUTPUT;
                } else if (ev == 5) {
                    uint64_t bp = 0;
                    if (lane == 0) bp = b.bitpos() >> 3;
                    bp = __shfl_sync(MZ_FULL_MASK, bp, 0);
                    loaded = inf_fill_input(iring, job.in, bp, loaded, padded_total);
                } else if (ev == 6) {
                    inf_flush_output(oring, out, flushed, out_pos);
                    flushed = out_pos;
                } else {
                    status = (int)ev;
                }
            }
        }
        /* ---- write back what is still in the ring, save state ----------------------------------------- */
        __syncwarp();
        inf_flush_output(oring, out, flushed, out_pos);
        if (phase == INF_PH_CODES && status == INF_ST_RUN)
            for (int i = (int)lane; i < 320; i += 32) st->lens[i] = T.lens[i];
        if (lane == 0) {
            uint64_t bp = b.bitpos();
            if (status == INF_ST_END) bp = (bp + 7) & ~7ull; /* the partial last byte is consumed */
            st->in_bitpos = job.in_base * 8 + bp;
            st->out_pos = out_pos;
            st->status = status;
            st->why = why;
            st->phase = phase;
            st->last_block = last;
            st->stored_remaining = stored_rem;
            st->nlit = nlit;
            st->ndist = ndist;
            st->blocks = blocks;
        }
        __syncwarp();
    }
}

} // namespace mzc
#endif
