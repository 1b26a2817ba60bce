/* deflate_kernel.cuh -- K2+K3: per-chunk LZ77 match + dynamic-Huffman RFC1951 encode (sm_100a).
 *
 * Replaces, on the write path, what zlib's deflate() does behind mz_stream_zlib_write /
 * mz_stream_zlib_close (mz_strm_zlib.c:203-240, :243-264, :280-305).  The compressed bytes are NOT
 * zlib's; parity = the reference inflate (mz_stream_zlib_read) reproduces the input bit-exactly.
 *
 * Work decomposition (one CTA of 1024 threads per chunk; persistent grid-stride loop over chunks):
 *   chunk      <= 64 KiB of input, independent LZ77 history, resident in shared memory (TMA bulk load)
 *   sub-block  32 KiB = 1024 threads x 32-byte segments; one DEFLATE block per sub-block
 *   A parse    every thread runs a greedy hash-table parser over its own 32-byte segment; matches may
 *              overrun the segment (up to 258 B, clamped at the sub-block end); the hash table
 *              (shared memory, 16-bit chunk-relative positions, racy by design: every candidate is
 *              validated by comparing bytes, and any earlier position is a legal LZ77 source)
 *   B cover    exclusive prefix-max over the threads' parse end positions: a thread drops / trims the
 *              tokens that an earlier thread's overrunning match already covers
 *   C hist     literal/length + distance histograms (shared-memory atomics)
 *   D codes    warp-parallel length-limited code construction (bisection on a global rounding offset
 *              of the ideal -log2 p lengths, then exact Kraft completion), canonical codes, block header
 *   E count    per-thread bit totals -> block exclusive scan -> bit offsets
 *   F emit     every thread packs its tokens at its bit offset into the staging buffer
 *   G flush    staging -> global in 16-byte units; partial tail carried into the next sub-block
 * Blocks that would not shrink are emitted as stored blocks.  A non-final chunk ends with an empty
 * stored block (00 00 FF FF after bit padding) so chunks join byte-wise; the final chunk carries BFINAL.
 */
#ifndef MZ_DEFLATE_KERNEL_CUH
#define MZ_DEFLATE_KERNEL_CUH

#include "mzcuda_common.cuh"

namespace mzc {

constexpr int DF_THREADS = 1024;
constexpr int DF_WARPS = DF_THREADS / 32;
constexpr int DF_CHUNK_MAX = 65536;
constexpr int DF_SEG = 32;
constexpr int DF_SB = DF_THREADS * DF_SEG; /* 32768 */
constexpr int DF_MINMATCH = 4;
constexpr int DF_MAXREC = DF_SEG / DF_MINMATCH; /* 8 match records per thread per sub-block */
constexpr int DF_HASH_ENTRIES = 16384;          /* u16 entries: 32 KiB */
constexpr int DF_STAGE_WORDS = DF_SB / 4 + 64;
constexpr int DF_HDR_WORDS = 96; /* dynamic header <= 17 + 57 + 316*7 bits = 2286 bits = 72 words */

constexpr uint32_t DF_FLAG_FINAL = 1u; /* chunk ends the stream: BFINAL on its last block, no sync marker */

/* shared-memory carve-up (bytes) */
constexpr int DF_OFF_IN = 0;
constexpr int DF_OFF_HASH = DF_OFF_IN + DF_CHUNK_MAX + 64;
constexpr int DF_OFF_REC = DF_OFF_HASH + DF_HASH_ENTRIES * 2;
constexpr int DF_OFF_STAGE = DF_OFF_REC + DF_MAXREC * DF_THREADS * 4;
constexpr int DF_OFF_HDR = DF_OFF_STAGE + DF_STAGE_WORDS * 4;
constexpr int DF_OFF_END = DF_OFF_HDR + DF_HDR_WORDS * 4;       /* u16[1024] parse end */
constexpr int DF_OFF_BITOFF = DF_OFF_END + DF_THREADS * 2;      /* u32[1024] */
constexpr int DF_OFF_HIST = DF_OFF_BITOFF + DF_THREADS * 4;     /* u32[288 + 32 + 32] */
constexpr int DF_OFF_CODE = DF_OFF_HIST + (288 + 32 + 32) * 4;  /* u32[288 + 32 + 32] code | len<<16 */
constexpr int DF_OFF_LENS = DF_OFF_CODE + (288 + 32 + 32) * 4;  /* u8[288 + 32 + 32] */
constexpr int DF_OFF_SCAN = DF_OFF_LENS + (288 + 32 + 32);      /* u32[64] */
constexpr int DF_OFF_MISC = DF_OFF_SCAN + 64 * 4;               /* u32[32] + mbarrier */
constexpr int DF_SMEM_BYTES = DF_OFF_MISC + 32 * 4 + 16;

enum { MISC_HDRBITS = 1, MISC_TOKBITS = 2, MISC_NLIT = 3, MISC_NDIST = 4, MISC_BLCNT = 8 /* 16 words */ };

struct DeflateParams {
    const uint8_t *in;       /* device base of the uncompressed bytes */
    const uint64_t *in_off;  /* per-chunk byte offset into `in`, or NULL for a uniform partition */
    const uint32_t *in_len;  /* per-chunk length (<= 65536), or NULL */
    const uint8_t *flags;    /* per-chunk DF_FLAG_*, or NULL */
    uint64_t total_len;      /* uniform partition: total bytes */
    uint32_t chunk_size;     /* uniform partition: bytes per chunk (<= 65536) */
    uint32_t nchunks;
    uint32_t last_flags;     /* uniform partition: flags of the last chunk */
    int32_t level;           /* 0 stored, 1..9 */
    uint8_t *out;            /* slot i at out + i * slot_stride (16-byte aligned) */
    uint64_t slot_stride;
    uint32_t *out_len;       /* per-chunk compressed bytes */
};

__host__ __device__ inline uint64_t deflate_slot_bound(uint32_t chunk_size) {
    return (((uint64_t)chunk_size + 5ull * ((chunk_size + DF_SB - 1) / DF_SB + 1) + 64 + 15) & ~15ull);
}

/* ---- symbol mapping (RFC1951 3.2.5) without tables ------------------------------------------- */
__device__ __forceinline__ void length_symbol(uint32_t len, uint32_t &sym, uint32_t &ebits, uint32_t &eval) {
    uint32_t l = len - 3;
    if (l < 8) {
        sym = 257 + l; ebits = 0; eval = 0;
    } else if (len == 258) {
        sym = 285; ebits = 0; eval = 0;
    } else {
        uint32_t msb = 31 - __clz((int)l); /* 3..7 */
        ebits = msb - 2;
        sym = 257 + 4 * (ebits + 1) + ((l >> ebits) & 3);
        eval = l & ((1u << ebits) - 1);
    }
}
__device__ __forceinline__ void dist_symbol(uint32_t dist, uint32_t &sym, uint32_t &ebits, uint32_t &eval) {
    uint32_t d = dist - 1;
    if (d < 4) {
        sym = d; ebits = 0; eval = 0;
    } else {
        uint32_t msb = 31 - __clz((int)d); /* 2..14 */
        ebits = msb - 1;
        sym = 2 * msb + ((d >> ebits) & 1);
        eval = d & ((1u << ebits) - 1);
    }
}

/* OR `n` (<=32) bits of v into the staging bit string at bit position pos */
__device__ __forceinline__ void stage_put(uint32_t *stage, uint32_t pos, uint32_t v, uint32_t n) {
    if (n == 0) return;
    if (n < 32) v &= (1u << n) - 1;
    uint32_t w = pos >> 5, s = pos & 31;
    atomicOr(&stage[w], v << s);
    if (s + n > 32) atomicOr(&stage[w + 1], v >> (32 - s));
}

/* ---- token walk --------------------------------------------------------------------------------
 * Thread t's parse tiles [seg_start, e_t) with literals and the recorded matches. `cover` is where
 * earlier threads' matches end; everything before it is dropped, a straddling match is trimmed. */
template <typename V>
__device__ __forceinline__ void walk_tokens(const uint8_t *s_in, const uint32_t *s_rec, uint32_t tid, uint32_t nrec,
                                            uint32_t seg_start, uint32_t seg_end, uint32_t cover, V &vis) {
    uint32_t pos = seg_start;
    for (uint32_t r = 0; r < nrec; r++) {
        uint32_t rec = s_rec[r * DF_THREADS + tid];
        uint32_t rs = seg_start + (rec & 31);
        uint32_t len = ((rec >> 5) & 255) + 3;
        uint32_t dist = (rec >> 13) + 1;
        for (uint32_t q = pos > cover ? pos : cover; q < rs; q++) vis.lit(s_in[q]);
        uint32_t mend = rs + len;
        if (mend > cover) {
            uint32_t s = rs > cover ? rs : cover;
            uint32_t rem = mend - s;
            if (rem >= 3) {
                vis.match(rem, dist);
            } else {
                for (uint32_t q = s; q < mend; q++) vis.lit(s_in[q]);
            }
        }
        pos = mend;
    }
    for (uint32_t q = pos > cover ? pos : cover; q < seg_end; q++) vis.lit(s_in[q]);
}

struct HistVisitor {
    uint32_t *hist_ll, *hist_d;
    __device__ __forceinline__ void lit(uint32_t b) { atomicAdd(&hist_ll[b], 1u); }
    __device__ __forceinline__ void match(uint32_t len, uint32_t dist) {
        uint32_t s, eb, ev;
        length_symbol(len, s, eb, ev);
        atomicAdd(&hist_ll[s], 1u);
        dist_symbol(dist, s, eb, ev);
        atomicAdd(&hist_d[s], 1u);
    }
};

struct CountVisitor {
    const uint32_t *code_ll, *code_d;
    uint32_t bits;
    __device__ __forceinline__ void lit(uint32_t b) { bits += code_ll[b] >> 16; }
    __device__ __forceinline__ void match(uint32_t len, uint32_t dist) {
        uint32_t s, eb, ev;
        length_symbol(len, s, eb, ev);
        bits += (code_ll[s] >> 16) + eb;
        dist_symbol(dist, s, eb, ev);
        bits += (code_d[s] >> 16) + eb;
    }
};

struct EmitVisitor {
    const uint32_t *code_ll, *code_d;
    uint32_t *stage;
    uint64_t acc;
    uint32_t nb, w;
    bool first;
    __device__ __forceinline__ void init(uint32_t bitoff) {
        w = bitoff >> 5; nb = bitoff & 31; acc = 0; first = true;
    }
    __device__ __forceinline__ void put(uint32_t v, uint32_t n) {
        acc |= (uint64_t)v << nb;
        nb += n;
        if (nb >= 32) {
            uint32_t lo = (uint32_t)acc;
            if (first) { atomicOr(&stage[w], lo); first = false; } else { stage[w] = lo; }
            acc >>= 32; nb -= 32; w++;
        }
    }
    __device__ __forceinline__ void finish() {
        if (nb > 0) atomicOr(&stage[w], (uint32_t)acc);
    }
    __device__ __forceinline__ void lit(uint32_t b) {
        uint32_t c = code_ll[b];
        put(c & 0xffff, c >> 16);
    }
    __device__ __forceinline__ void match(uint32_t len, uint32_t dist) {
        uint32_t s, eb, ev;
        length_symbol(len, s, eb, ev);
        uint32_t c = code_ll[s];
        uint32_t cl = c >> 16;
        put((c & 0xffff) | (ev << cl), cl + eb); /* <= 15 + 5 */
        dist_symbol(dist, s, eb, ev);
        c = code_d[s];
        cl = c >> 16;
        put((c & 0xffff) | (ev << cl), cl + eb); /* <= 15 + 13 */
    }
};

/* ---- warp-parallel code construction --------------------------------------------------------- */
constexpr int DF_KSLOTS = 9; /* 9 * 32 = 288 symbols per warp pass */

/* Length-limited prefix-code lengths for `n` (<=288) symbols, max `M` bits. One full warp.
 * Output: complete code (Kraft sum exactly 1) with >= 2 coded symbols, as zlib's inflate requires
 * of dynamic blocks. lens[i] = 0 for unused symbols. */
__device__ inline void warp_build_lengths(const uint32_t *hist, int n, int M, uint8_t *lens) {
    const unsigned lane = lane_id();
    uint32_t c[DF_KSLOTS];
    float ideal[DF_KSLOTS];
    uint32_t used = 0, total = 0, first_used = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < DF_KSLOTS; k++) {
        int i = k * 32 + (int)lane;
        c[k] = (i < n) ? hist[i] : 0u;
        if (c[k]) {
            used++;
            if (first_used == 0xffffffffu) first_used = (uint32_t)i;
        }
    }
    used = __reduce_add_sync(MZ_FULL_MASK, used);
    first_used = __reduce_min_sync(MZ_FULL_MASK, first_used);
    if (used < 2) { /* force two coded symbols */
        uint32_t d0 = (used == 0) ? 0u : (first_used == 0 ? 1u : 0u);
        uint32_t d1 = (used == 0) ? 1u : d0;
        if (lane == d0 && c[0] == 0) c[0] = 1;
        if (lane == d1 && c[0] == 0) c[0] = 1;
    }
#pragma unroll
    for (int k = 0; k < DF_KSLOTS; k++) total += c[k];
    total = __reduce_add_sync(MZ_FULL_MASK, total);
    const float lt = __log2f((float)total);
#pragma unroll
    for (int k = 0; k < DF_KSLOTS; k++) ideal[k] = c[k] ? lt - __log2f((float)c[k]) : 0.f;

    const uint32_t one = 1u << M;
    float lo = -3.0f, hi = 1.0f;
    for (int it = 0; it < 10; it++) {
        float mid = 0.5f * (lo + hi);
        uint32_t kr = 0;
#pragma unroll
        for (int k = 0; k < DF_KSLOTS; k++)
            if (c[k]) {
                int l = (int)ceilf(ideal[k] - mid);
                l = l < 1 ? 1 : (l > M ? M : l);
                kr += 1u << (M - l);
            }
        kr = __reduce_add_sync(MZ_FULL_MASK, kr);
        if (kr <= one) lo = mid; else hi = mid;
    }
    int L[DF_KSLOTS];
    uint32_t kr = 0;
#pragma unroll
    for (int k = 0; k < DF_KSLOTS; k++) {
        L[k] = 0;
        if (c[k]) {
            int l = (int)ceilf(ideal[k] - lo);
            L[k] = l < 1 ? 1 : (l > M ? M : l);
            kr += 1u << (M - L[k]);
        }
    }
    kr = __reduce_add_sync(MZ_FULL_MASK, kr);
    uint32_t slack = one - kr; /* kr <= one by construction (lo always feasible, -3 is) */
    /* exact completion: shorten codes, short ones first, until the Kraft sum is exactly 1 */
    for (int pass = 0; pass < 32 && slack > 0; pass++) {
        for (int len = 2; len <= M && slack > 0; len++) {
            uint32_t w = 1u << (M - len);
            uint32_t can = slack >> (M - len); /* how many symbols of this length may be shortened */
            if (can == 0) continue;
            uint32_t base = 0;
#pragma unroll
            for (int k = 0; k < DF_KSLOTS; k++) {
                unsigned b = __ballot_sync(MZ_FULL_MASK, L[k] == len);
                uint32_t rank = base + (uint32_t)__popc(b & ((1u << lane) - 1));
                if (L[k] == len && rank < can) L[k] = len - 1;
                base += (uint32_t)__popc(b);
            }
            uint32_t took = base < can ? base : can;
            slack -= took * w;
        }
    }
#pragma unroll
    for (int k = 0; k < DF_KSLOTS; k++) {
        int i = k * 32 + (int)lane;
        if (i < n) lens[i] = (uint8_t)L[k];
    }
    __syncwarp();
}

/* Canonical codes (RFC1951 3.2.2), bit-reversed for LSB-first packing. One full warp.
 * codes[i] = reversed_code | len << 16.  `scratch` = 16 words of shared memory. */
__device__ inline void warp_assign_codes(const uint8_t *lens, int n, uint32_t *codes, uint32_t *scratch) {
    const unsigned lane = lane_id();
    if (lane < 16) scratch[lane] = 0;
    __syncwarp();
    for (int i = (int)lane; i < n; i += 32)
        if (lens[i]) atomicAdd(&scratch[lens[i]], 1u);
    __syncwarp();
    uint32_t next = 0;
    if (lane >= 1 && lane < 16) {
        for (unsigned b = 1; b <= lane; b++) next = (next + scratch[b - 1]) << 1; /* scratch[0] == 0 */
    }
    __syncwarp();
    if (lane < 16) scratch[lane] = next;
    __syncwarp();
    for (int base = 0; base < n; base += 32) {
        int i = base + (int)lane;
        uint32_t l = (i < n) ? lens[i] : 0u;
        unsigned m = __match_any_sync(MZ_FULL_MASK, l);
        uint32_t rank = (uint32_t)__popc(m & ((1u << lane) - 1));
        if (l) {
            uint32_t code = scratch[l] + rank;
            codes[i] = (__brev(code) >> (32 - l)) | (l << 16);
        } else if (i < n) {
            codes[i] = 0;
        }
        __syncwarp();
        if (l && rank == 0) scratch[l] += (uint32_t)__popc(m);
        __syncwarp();
    }
}

/* Build the dynamic block header bit string into hdr[] (starting at bit 0); returns bit count.
 * One full warp. lens_ll[288], lens_d[32] final; hist_cl/codes_cl/lens_cl/scratch shared scratch. */
__device__ inline uint32_t warp_build_header(const uint8_t *lens_ll, const uint8_t *lens_d, uint32_t bfinal, uint32_t *hist_cl,
                                             uint8_t *lens_cl, uint32_t *codes_cl, uint32_t *scratch, uint32_t *hdr) {
    const unsigned lane = lane_id();
    /* HLIT / HDIST: trailing zero lengths are not sent */
    uint32_t last_ll = 0, last_d = 0;
    for (int i = (int)lane; i < 286; i += 32)
        if (lens_ll[i]) last_ll = (uint32_t)i;
    if (lane < 30 && lens_d[lane]) last_d = lane;
    last_ll = __reduce_max_sync(MZ_FULL_MASK, last_ll);
    last_d = __reduce_max_sync(MZ_FULL_MASK, last_d);
    const uint32_t nlit = last_ll + 1 < 257 ? 257 : last_ll + 1;
    const uint32_t ndist = last_d + 1;
    const uint32_t nseq = nlit + ndist;
    if (lane < 19) hist_cl[lane] = 0;
    for (int j = (int)lane; j < DF_HDR_WORDS; j += 32) hdr[j] = 0;
    __syncwarp();
    for (uint32_t j = lane; j < nseq; j += 32) {
        uint32_t v = j < nlit ? lens_ll[j] : lens_d[j - nlit];
        atomicAdd(&hist_cl[v], 1u);
    }
    __syncwarp();
    warp_build_lengths(hist_cl, 19, 7, lens_cl);
    warp_assign_codes(lens_cl, 19, codes_cl, scratch);
    __syncwarp();
    uint32_t pos = 0;
    if (lane == 0) {
        stage_put(hdr, 0, bfinal | (2u << 1), 3);
        stage_put(hdr, 3, nlit - 257, 5);
        stage_put(hdr, 8, ndist - 1, 5);
        stage_put(hdr, 13, 19 - 4, 4);
    }
    pos = 17;
    if (lane < 19) {
        /* order of code-length code lengths, RFC1951 3.2.7 */
        const uint32_t ord_lo = 0x0A060908u, ord_mid = 0x030C040Bu; /* unused packing helpers */
        (void)ord_lo; (void)ord_mid;
        uint32_t sym;
        switch (lane) {
            case 0: sym = 16; break; case 1: sym = 17; break; case 2: sym = 18; break; case 3: sym = 0; break;
            case 4: sym = 8; break; case 5: sym = 7; break; case 6: sym = 9; break; case 7: sym = 6; break;
            case 8: sym = 10; break; case 9: sym = 5; break; case 10: sym = 11; break; case 11: sym = 4; break;
            case 12: sym = 12; break; case 13: sym = 3; break; case 14: sym = 13; break; case 15: sym = 2; break;
            case 16: sym = 14; break; case 17: sym = 1; break; default: sym = 15; break;
        }
        stage_put(hdr, pos + 3 * lane, lens_cl[sym], 3);
    }
    pos += 57;
    __syncwarp();
    for (uint32_t base = 0; base < nseq; base += 32) {
        uint32_t j = base + lane;
        uint32_t code = 0, nb = 0;
        if (j < nseq) {
            uint32_t v = j < nlit ? lens_ll[j] : lens_d[j - nlit];
            code = codes_cl[v] & 0xffff;
            nb = codes_cl[v] >> 16;
        }
        uint32_t incl = warp_incl_sum(nb);
        stage_put(hdr, pos + incl - nb, code, nb);
        pos += __shfl_sync(MZ_FULL_MASK, incl, 31);
    }
    __syncwarp();
    return pos;
}

/* block-wide exclusive scans over one value per thread; `scan` = 64 words of shared scratch.
 * Contains __syncthreads: every thread of the CTA must call. */
__device__ inline uint32_t block_excl_sum(uint32_t v, uint32_t *scan, uint32_t &total) {
    uint32_t incl = warp_incl_sum(v);
    if (lane_id() == 31) scan[warp_id()] = incl;
    __syncthreads();
    if (warp_id() == 0) {
        uint32_t w = scan[lane_id()];
        uint32_t wi = warp_incl_sum(w);
        scan[32 + lane_id()] = wi - w;
        if (lane_id() == 31) scan[31] = wi; /* grand total parked in slot 31 after use */
    }
    __syncthreads();
    uint32_t res = scan[32 + warp_id()] + incl - v;
    total = scan[31];
    __syncthreads();
    return res;
}
__device__ inline uint32_t block_excl_max(uint32_t v, uint32_t identity, uint32_t *scan) {
    uint32_t incl = warp_incl_max(v);
    if (lane_id() == 31) scan[warp_id()] = incl;
    __syncthreads();
    if (warp_id() == 0) {
        uint32_t w = scan[lane_id()];
        uint32_t wi = warp_incl_max(w);
        uint32_t ex = __shfl_up_sync(MZ_FULL_MASK, wi, 1);
        scan[32 + lane_id()] = lane_id() == 0 ? identity : ex;
    }
    __syncthreads();
    uint32_t prev = __shfl_up_sync(MZ_FULL_MASK, incl, 1);
    uint32_t wbase = scan[32 + warp_id()];
    uint32_t res = lane_id() == 0 ? wbase : (prev > wbase ? prev : wbase);
    __syncthreads();
    return res;
}

/* ---- the kernel ------------------------------------------------------------------------------- */
struct MatchCfg {
    int ways;  /* candidates per hash bucket: 1, 2 or 4 */
    int lazy;  /* one-step lazy evaluation */
};
__host__ __device__ inline MatchCfg match_cfg_for_level(int level) {
    MatchCfg m;
    m.ways = level <= 1 ? 1 : (level <= 3 ? 2 : 4);
    m.lazy = level >= 6;
    return m;
}

/* longest match of in[p..] against in[cand..], both inside the chunk, at most maxlen bytes;
 * first 4 bytes already known equal */
__device__ __forceinline__ uint32_t extend_match(const uint8_t *s_in, uint32_t cand, uint32_t p, uint32_t maxlen) {
    uint32_t len = 4;
    while (len < maxlen) {
        uint32_t a = load32u(s_in, p + len), b = load32u(s_in, cand + len);
        uint32_t x = a ^ b;
        if (x) {
            len += (uint32_t)(__ffs((int)x) - 1) >> 3;
            break;
        }
        len += 4;
    }
    return len < maxlen ? len : maxlen;
}

__device__ __forceinline__ uint32_t hash4(uint32_t v, int bits) { return (v * 2654435761u) >> (32 - bits); }

/* find the best match at p among the bucket's candidates and insert p; returns len (0 = none) */
__device__ __forceinline__ uint32_t find_match(const uint8_t *s_in, uint16_t *s_hash, uint32_t p, uint32_t limit, int ways,
                                               uint32_t &best_dist) {
    uint32_t v = load32u(s_in, p);
    uint32_t maxlen = limit - p;
    if (maxlen > 258) maxlen = 258;
    uint32_t best = 0;
    best_dist = 0;
    if (ways == 1) {
        uint32_t h = hash4(v, 14);
        uint32_t cand = s_hash[h];
        s_hash[h] = (uint16_t)p;
        if (cand < p && p - cand <= 32768u && load32u(s_in, cand) == v) {
            best = extend_match(s_in, cand, p, maxlen);
            best_dist = p - cand;
        }
    } else {
        const int hb = ways == 2 ? 13 : 12;
        uint32_t h = hash4(v, hb) * (uint32_t)ways;
        uint32_t prev_slot = p;
        for (int wy = 0; wy < ways; wy++) {
            uint32_t cand = s_hash[h + wy];
            s_hash[h + wy] = (uint16_t)prev_slot; /* FIFO: newest first */
            prev_slot = cand;
            if (cand < p && p - cand <= 32768u && load32u(s_in, cand) == v) {
                uint32_t l = extend_match(s_in, cand, p, maxlen);
                if (l > best) { best = l; best_dist = p - cand; }
            }
        }
    }
    return best;
}

__global__ void __launch_bounds__(DF_THREADS, 1) deflate_chunks_kernel(DeflateParams P) {
    MZ_DYN_SMEM(smem);
    uint8_t *s_in = smem + DF_OFF_IN;
    uint16_t *s_hash = (uint16_t *)(smem + DF_OFF_HASH);
    uint32_t *s_rec = (uint32_t *)(smem + DF_OFF_REC);
    uint32_t *s_stage = (uint32_t *)(smem + DF_OFF_STAGE);
    uint32_t *s_hdr = (uint32_t *)(smem + DF_OFF_HDR);
    uint32_t *s_hist_ll = (uint32_t *)(smem + DF_OFF_HIST);
    uint32_t *s_hist_d = s_hist_ll + 288;
    uint32_t *s_hist_cl = s_hist_d + 32;
    uint32_t *s_code_ll = (uint32_t *)(smem + DF_OFF_CODE);
    uint32_t *s_code_d = s_code_ll + 288;
    uint32_t *s_code_cl = s_code_d + 32;
    uint8_t *s_lens_ll = smem + DF_OFF_LENS;
    uint8_t *s_lens_d = s_lens_ll + 288;
    uint8_t *s_lens_cl = s_lens_d + 32;
    uint32_t *s_scan = (uint32_t *)(smem + DF_OFF_SCAN);
    uint32_t *s_misc = (uint32_t *)(smem + DF_OFF_MISC);
#ifndef MZ_EMU
    uint64_t *s_bar = (uint64_t *)(smem + DF_OFF_MISC + 32 * 4);
    uint32_t bar_phase = 0;
    if (threadIdx.x == 0) {
        mbar_init(s_bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
#endif
    const uint32_t tid = threadIdx.x;
    const MatchCfg mcfg = match_cfg_for_level(P.level);

    for (uint32_t chunk = blockIdx.x; chunk < P.nchunks; chunk += gridDim.x) {
        /* ---- locate the chunk -------------------------------------------------------------- */
        uint64_t off;
        uint32_t len, flags;
        if (P.in_off) {
            off = P.in_off[chunk];
            len = P.in_len[chunk];
            flags = P.flags ? P.flags[chunk] : 0u;
        } else {
            off = (uint64_t)chunk * P.chunk_size;
            uint64_t rem = P.total_len - off;
            len = rem < P.chunk_size ? (uint32_t)rem : P.chunk_size;
            flags = P.flags ? P.flags[chunk] : (chunk == P.nchunks - 1 ? P.last_flags : 0u);
        }
        const uint8_t *gin = P.in + off;
        uint8_t *gout = P.out + (uint64_t)chunk * P.slot_stride;

        /* ---- load input into shared memory, reset tables ---------------------------------------- */
        const uint32_t len16 = len & ~15u;
        bool bulk = false;
#ifndef MZ_EMU
        bulk = (((uintptr_t)gin) & 15) == 0 && len16 > 0;
        if (bulk && tid == 0) {
            fence_proxy_async(); /* earlier generic-proxy reads of s_in are done (trailing __syncthreads) */
            mbar_expect_tx(s_bar, len16);
            tma_load_1d(s_in, gin, len16, s_bar);
        }
#endif
        if (!bulk) {
            if ((((uintptr_t)gin) & 15) == 0) {
                for (uint32_t i = tid * 16; i < len16; i += DF_THREADS * 16) *(uint4 *)(s_in + i) = ldg_stream((const uint4 *)(gin + i));
            } else {
                for (uint32_t i = tid; i < len16; i += DF_THREADS) s_in[i] = gin[i];
            }
        }
        for (uint32_t i = len16 + tid; i < len; i += DF_THREADS) s_in[i] = gin[i];
        for (uint32_t i = len + tid; i < ((len + 63) & ~15u) + 16 && i < DF_CHUNK_MAX + 64; i += DF_THREADS) s_in[i] = 0; /* zero pad */
        for (uint32_t i = tid; i < DF_HASH_ENTRIES / 8; i += DF_THREADS) ((uint4 *)s_hash)[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
        for (uint32_t i = tid; i < DF_STAGE_WORDS; i += DF_THREADS) s_stage[i] = 0;
#ifndef MZ_EMU
        if (bulk) {
            mbar_wait(s_bar, bar_phase);
            bar_phase ^= 1;
        }
#endif
        __syncthreads();

        uint32_t flushed = 0; /* bytes of this chunk already in global memory (multiple of 16) */
        uint32_t bitpos = 0;  /* valid bits in the staging buffer; uniform across the CTA */
        const uint32_t nsb = (len + DF_SB - 1) / DF_SB;

        if (len == 0 && (flags & DF_FLAG_FINAL)) {
            /* zlib's answer for an empty stream: one fixed-Huffman block holding only EOB = 03 00 */
            if (tid == 0) stage_put(s_stage, 0, 1u | (1u << 1), 3);
            bitpos = 10;
            __syncthreads();
        }

        for (uint32_t sb = 0; sb < nsb; sb++) {
            const uint32_t sb_start = sb * DF_SB;
            const uint32_t sb_end = (sb_start + DF_SB < len) ? sb_start + DF_SB : len;
            const uint32_t sb_len = sb_end - sb_start;
            const uint32_t bfinal = (sb == nsb - 1 && (flags & DF_FLAG_FINAL)) ? 1u : 0u;
            const uint32_t seg_start = sb_start + tid * DF_SEG < sb_end ? sb_start + tid * DF_SEG : sb_end;
            const uint32_t seg_end = seg_start + DF_SEG < sb_end ? seg_start + DF_SEG : sb_end;
            bool stored = (P.level == 0);
            uint32_t nrec = 0, cover = 0;

            if (!stored) {
                /* ---- A: parse ------------------------------------------------------------------ */
                uint32_t p = seg_start;
                while (p < seg_end) {
                    uint32_t mlen = 0, mdist = 0;
                    if (p + DF_MINMATCH <= sb_end) {
                        mlen = find_match(s_in, s_hash, p, sb_end, mcfg.ways, mdist);
                        if (mcfg.lazy && mlen >= DF_MINMATCH && mlen < 32 && p + 1 < seg_end && p + 1 + DF_MINMATCH <= sb_end) {
                            uint32_t d2, l2 = find_match(s_in, s_hash, p + 1, sb_end, mcfg.ways, d2);
                            if (l2 > mlen) { /* literal now, better match next */
                                p += 1;
                                mlen = l2;
                                mdist = d2;
                            }
                        }
                    }
                    if (mlen >= DF_MINMATCH && nrec < DF_MAXREC) {
                        s_rec[nrec * DF_THREADS + tid] = (p - seg_start) | ((mlen - 3) << 5) | ((mdist - 1) << 13);
                        nrec++;
                        p += mlen;
                    } else {
                        p += 1;
                    }
                }
                /* ---- B: cover = where earlier threads' matches end ------------------------------ */
                cover = block_excl_max(p, sb_start, s_scan);
                for (uint32_t i = tid; i < 288 + 32; i += DF_THREADS) s_hist_ll[i] = 0;
                __syncthreads();
                /* ---- C: histograms ---------------------------------------------------------------- */
                {
                    HistVisitor hv;
                    hv.hist_ll = s_hist_ll;
                    hv.hist_d = s_hist_d;
                    if (cover < seg_end || nrec) walk_tokens(s_in, s_rec, tid, nrec, seg_start, seg_end, cover, hv);
                    if (tid == 0) s_hist_ll[256] = 1;
                }
                __syncthreads();
                /* ---- D: codes + header -------------------------------------------------------- */
                if (warp_id() == 0) {
                    warp_build_lengths(s_hist_ll, 286, 15, s_lens_ll);
                    if (lane_id() < 2) s_lens_ll[286 + lane_id()] = 0;
                    warp_assign_codes(s_lens_ll, 286, s_code_ll, s_misc + MISC_BLCNT);
                } else if (warp_id() == 1) {
                    warp_build_lengths(s_hist_d, 30, 15, s_lens_d);
                    if (lane_id() < 2) s_lens_d[30 + lane_id()] = 0;
                    warp_assign_codes(s_lens_d, 30, s_code_d, s_scan + 40); /* scan scratch idle here */
                }
                __syncthreads();
                if (warp_id() == 0) {
                    uint32_t hb = warp_build_header(s_lens_ll, s_lens_d, bfinal, s_hist_cl, s_lens_cl, s_code_cl,
                                                    s_misc + MISC_BLCNT, s_hdr);
                    if (lane_id() == 0) s_misc[MISC_HDRBITS] = hb;
                }
                /* ---- E: bit counts (other warps proceed; codes are final) ------------------------ */
                CountVisitor cv;
                cv.code_ll = s_code_ll;
                cv.code_d = s_code_d;
                cv.bits = 0;
                if (cover < seg_end || nrec) walk_tokens(s_in, s_rec, tid, nrec, seg_start, seg_end, cover, cv);
                uint32_t tokbits;
                uint32_t myoff = block_excl_sum(cv.bits, s_scan, tokbits);
                const uint32_t hdrbits = s_misc[MISC_HDRBITS];
                const uint32_t eob = s_code_ll[256];
                const uint32_t dyn_bits = hdrbits + tokbits + (eob >> 16);
                const uint32_t stored_bits = (((bitpos + 3 + 7) & ~7u) - bitpos) + 32 + sb_len * 8;
                if (dyn_bits >= stored_bits) {
                    stored = true;
                } else {
                    /* ---- F: emit --------------------------------------------------------------- */
                    const uint32_t base = bitpos + hdrbits;
                    if (warp_id() == 0) {
                        for (uint32_t j = lane_id(); j * 32 < hdrbits; j += 32) {
                            uint32_t n = hdrbits - j * 32;
                            stage_put(s_stage, bitpos + j * 32, s_hdr[j], n > 32 ? 32 : n);
                        }
                    }
                    if (cv.bits) {
                        EmitVisitor ev;
                        ev.code_ll = s_code_ll;
                        ev.code_d = s_code_d;
                        ev.stage = s_stage;
                        ev.init(base + myoff);
                        walk_tokens(s_in, s_rec, tid, nrec, seg_start, seg_end, cover, ev);
                        ev.finish();
                    }
                    if (tid == DF_THREADS - 1) stage_put(s_stage, base + tokbits, eob & 0xffff, eob >> 16);
                    bitpos += dyn_bits;
                }
            }
            if (stored) {
                /* stored block: header, pad to byte, LEN, ~LEN, raw bytes */
                const uint32_t p0 = (bitpos + 3 + 7) >> 3; /* byte index of LEN */
                if (tid == 0) {
                    stage_put(s_stage, bitpos, bfinal, 3);
                    stage_put(s_stage, p0 * 8, sb_len, 16);
                    stage_put(s_stage, p0 * 8 + 16, sb_len ^ 0xffffu, 16);
                }
                for (uint32_t i = tid * 4; i < sb_len; i += DF_THREADS * 4) {
                    uint32_t v = load32u(s_in, sb_start + i);
                    uint32_t n = sb_len - i;
                    stage_put(s_stage, (p0 + 4 + i) * 8, v, n >= 4 ? 32 : n * 8);
                }
                bitpos = (p0 + 4 + sb_len) * 8;
            }
            __syncthreads();
            /* ---- G: flush whole 16-byte units, carry the tail ------------------------------------ */
            {
                const uint32_t n16 = bitpos >> 7;
                const uint32_t used_words = (bitpos + 31) >> 5;
                for (uint32_t i = tid; i < n16; i += DF_THREADS) ((uint4 *)(gout + flushed))[i] = ((const uint4 *)s_stage)[i];
                uint32_t carry = 0;
                if (tid < 4 && n16 * 4 + tid < used_words) carry = s_stage[n16 * 4 + tid];
                __syncthreads();
                for (uint32_t i = tid; i < used_words; i += DF_THREADS) s_stage[i] = 0;
                __syncthreads();
                if (tid < 4) s_stage[tid] = carry;
                bitpos -= n16 * 128;
                flushed += n16 * 16;
                __syncthreads();
            }
        }

        /* ---- chunk trailer ---------------------------------------------------------------------- */
        {
            if (flags & DF_FLAG_FINAL) {
                bitpos = (bitpos + 7) & ~7u;
            } else {
                /* empty stored block: BFINAL=0 BTYPE=00, pad, LEN=0000 NLEN=FFFF (sync-flush marker) */
                uint32_t p0 = (bitpos + 3 + 7) >> 3;
                if (tid == 0) stage_put(s_stage, p0 * 8 + 16, 0xffffu, 16);
                bitpos = (p0 + 4) * 8;
            }
            __syncthreads();
            const uint32_t bytes = bitpos >> 3;
            const uint32_t n16 = (bytes + 15) >> 4; /* slot has >= 16 bytes of slack */
            for (uint32_t i = tid; i < n16; i += DF_THREADS) ((uint4 *)(gout + flushed))[i] = ((const uint4 *)s_stage)[i];
            if (tid == 0) P.out_len[chunk] = flushed + bytes;
            __syncthreads();
        }
    }
}

} // namespace mzc
#endif
