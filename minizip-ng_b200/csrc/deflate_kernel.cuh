/* deflate_kernel.cuh -- K2+K3: per-chunk LZ77 match + dynamic-Huffman RFC1951 encode (sm_100a).
 *
 * Replaces, on the write path, what zlib's deflate() does behind mz_stream_zlib_write /
 * mz_stream_zlib_close (mz_strm_zlib.c:203-240, :243-264, :280-305).  The compressed bytes are NOT
 * zlib's; parity = the reference inflate (mz_stream_zlib_read) reproduces the input bit-exactly.
 *
 * Work decomposition (one CTA of 1024 threads per chunk; persistent grid-stride loop over chunks):
 *   chunk      <= 64 KiB of input, independent LZ77 history, resident in shared memory (TMA bulk load)
 *   sub-block  32 KiB = 1024 threads x 32-byte segments; normally one DEFLATE block per sub-block
 *   A parse    every thread runs a greedy hash-table parser over its own 32-byte segment; matches may
 *              overrun the segment (up to 258 B, clamped at the sub-block end); the hash table
 *              (shared memory, 16-bit chunk-relative positions) is racy by design: every candidate is
 *              validated by comparing bytes, and any earlier position is a legal LZ77 source
 *   B cover    exclusive prefix-max over the threads' parse end positions: a thread drops / trims the
 *              tokens that an earlier thread's overrunning match already covers
 *   T tokens   every thread turns what it keeps into entries of ONE block-wide ordered token list
 *              (16-bit: position | match flag), placed by a block exclusive scan of token counts
 *   C hist     threads stride over the list: literal/length + distance histograms (shared-memory
 *              atomics); match records are rewritten in place to their symbol form
 *   D codes    block-parallel length-limited prefix codes: one thread per symbol, two 16-candidate
 *              sweeps of a global rounding offset on the ideal -log2 p lengths (Kraft sums by warp
 *              reduce + atomics), exact Kraft completion and canonical code ranks via per-warp
 *              match_any counts; code-length code + header by one warp, overlapped with E
 *   E count    every thread owns an equal, contiguous run of tokens: bit totals -> block scan
 *   F emit     every thread packs its run at its bit offset into the staging buffer
 *   G flush    staging -> global in 16-byte units; partial tail carried into the next block
 * Blocks that would not shrink are emitted as stored blocks. If a sub-block has more tokens than the
 * list holds it is coded as two blocks (thread halves). A non-final chunk ends with an empty stored
 * block (00 00 FF FF after bit padding) so chunks join byte-wise; the final chunk carries BFINAL.
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
constexpr int DF_HDR_WORDS = 96;   /* dynamic header <= 17 + 57 + 316*7 bits = 2286 bits = 72 words */
constexpr int DF_TOK_CAP = 24576;  /* token list entries (u16) */
constexpr uint32_t DF_TOK_MATCH = 0x8000u; /* entry is a code word of a match: bits 0-12 = record slot */
constexpr uint32_t DF_TOK_DIST = 0x4000u;  /* ... its distance code word (else its length code word) */

constexpr uint32_t DF_FLAG_FINAL = 1u; /* chunk ends the stream: BFINAL on its last block, no sync marker */

/* shared-memory carve-up (bytes) */
constexpr int DF_OFF_IN = 0;
constexpr int DF_OFF_HASH = DF_OFF_IN + DF_CHUNK_MAX + 64;
constexpr int DF_OFF_REC = DF_OFF_HASH + DF_HASH_ENTRIES * 2;
constexpr int DF_OFF_STAGE = DF_OFF_REC + DF_MAXREC * DF_THREADS * 4;
constexpr int DF_OFF_TOK = DF_OFF_STAGE + DF_STAGE_WORDS * 4;
constexpr int DF_OFF_HDR = DF_OFF_TOK + DF_TOK_CAP * 2;
constexpr int DF_OFF_HIST = DF_OFF_HDR + DF_HDR_WORDS * 4;      /* u32[288 + 32 + 32] */
constexpr int DF_OFF_CODE = DF_OFF_HIST + (288 + 32 + 32) * 4;  /* u32[288 + 32 + 32] code | len<<16 */
constexpr int DF_OFF_LENS = DF_OFF_CODE + (288 + 32 + 32) * 4;  /* u8[288 + 32 + 32] */
constexpr int DF_OFF_SCAN = DF_OFF_LENS + (288 + 32 + 32);      /* u32[64] */
constexpr int DF_OFF_BB = DF_OFF_SCAN + 64 * 4;                 /* u32[512] block code-builder scratch */
constexpr int DF_OFF_MISC = DF_OFF_BB + 512 * 4;                /* u32[32] + mbarrier */
constexpr int DF_SMEM_BYTES = DF_OFF_MISC + 32 * 4 + 16;
static_assert(DF_SMEM_BYTES <= 227 * 1024, "shared memory budget");

enum { MISC_HDRBITS = 1, MISC_E0 = 2, MISC_CHUNK = 3, MISC_BLCNT = 8 /* 16 words */ };

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
    uint32_t *work_counter;  /* zeroed before launch: CTAs take chunks dynamically (NULL = static striding) */
};

__host__ __device__ inline uint64_t deflate_slot_bound(uint32_t chunk_size) {
    /* stored worst case: 5 bytes per block, up to 2 blocks per 32 KiB sub-block, + sync marker + slack */
    return (((uint64_t)chunk_size + 12ull * ((chunk_size + DF_SB - 1) / DF_SB + 1) + 64 + 15) & ~15ull);
}

/* ---- symbol mapping (RFC1951 3.2.5) without tables ------------------------------------------- */
__device__ __forceinline__ void length_symbol(uint32_t len, uint32_t &sym, uint32_t &ebits, uint32_t &eval) {
    uint32_t l = len - 3; /* sym is 0..28 (add 257 for the alphabet index) */
    if (l < 8) {
        sym = l; ebits = 0; eval = 0;
    } else if (len == 258) {
        sym = 28; ebits = 0; eval = 0;
    } else {
        uint32_t msb = 31 - __clz((int)l); /* 3..7 */
        ebits = msb - 2;
        sym = 4 * (ebits + 1) + ((l >> ebits) & 3);
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
__device__ __forceinline__ uint32_t len_extra_bits(uint32_t lsym) { return (lsym < 8 || lsym == 28) ? 0u : (lsym >> 2) - 1; }
__device__ __forceinline__ uint32_t dist_extra_bits(uint32_t dsym) { return dsym < 4 ? 0u : (dsym >> 1) - 1; }

/* match record forms (32 bits each, 8 per thread, [r][tid] layout):
 *   parsed : off(5) | (len-3)(8) << 5 | (dist-1)(15) << 13
 *   symbol : lsym(5) | lextra(5) << 5 | dsym(5) << 10 | dextra(13) << 15     (after phase C) */

/* One token-list entry = one code word. Branch-free decode shared by the count and emit passes:
 * idx = index into the concatenated code table (288 literal/length codes, then 32 distance codes),
 * ev / eb = extra-bits value / count. Literal entries carry the byte; for them the record load reads an
 * unrelated but valid word and the selects discard it. */
__device__ __forceinline__ void token_code(uint32_t e, const Smem &sm, uint32_t &idx, uint32_t &ev, uint32_t &eb) {
    const uint32_t rec = sm.ld32(DF_OFF_REC + (e & 0x1fffu) * 4);
    const bool isrec = (e & DF_TOK_MATCH) != 0, isdist = (e & DF_TOK_DIST) != 0;
    const uint32_t ls = rec & 31, ds = (rec >> 10) & 31;
    const uint32_t sym = isdist ? 288u + ds : 257u + ls;
    const uint32_t xv = isdist ? rec >> 15 : (rec >> 5) & 31u;
    const uint32_t xb = isdist ? dist_extra_bits(ds) : len_extra_bits(ls);
    idx = isrec ? sym : e;
    ev = isrec ? xv : 0u;
    eb = isrec ? xb : 0u;
}

/* OR `n` (<=32) bits of v into the staging bit string at bit position pos */
__device__ __forceinline__ void stage_put(uint32_t *stage, uint32_t pos, uint32_t v, uint32_t n) {
    if (n == 0) return;
    if (n < 32) v &= (1u << n) - 1;
    uint32_t w = pos >> 5, s = pos & 31;
    atomicOr(&stage[w], v << s);
    if (s + n > 32) atomicOr(&stage[w + 1], v >> (32 - s));
}

/* per-thread bit packer: first word of the range by atomicOr (shared with the previous thread), words it
 * fills completely by plain store, the trailing partial word by atomicOr */
struct BitWriter {
    Smem sm;
    uint64_t acc;
    uint32_t nb, wa; /* wa = byte offset of the current staging word */
    __device__ __forceinline__ void init(const Smem &s, uint32_t bitoff) {
        sm = s; wa = DF_OFF_STAGE + (bitoff >> 5) * 4; nb = bitoff & 31; acc = 0;
    }
    __device__ __forceinline__ void put(uint32_t v, uint32_t n) {
        acc |= (uint64_t)v << nb;
        nb += n;
        if (nb >= 32) { /* short body: predicated, no divergence */
            sm.red_or32(wa, (uint32_t)acc);
            acc >>= 32; nb -= 32; wa += 4;
        }
    }
    __device__ __forceinline__ void finish() {
        if (nb > 0) sm.red_or32(wa, (uint32_t)acc);
    }
};

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

/* ---- D: block-parallel literal/length + distance codes ------------------------------------------------
 * Thread i < 288 owns literal/length symbol i (286 used), thread 288 + j owns distance symbol j (30 used);
 * both alphabets are warp aligned (warps 0-8 and warp 9). Called by threads 0..319 ONLY; they synchronise
 * among themselves on named barrier 1 so the other 22 warps are not dragged through a dozen barriers.
 * bb = 512 words of shared scratch. Outputs lens (u8) and codes (reversed code | len << 16). */
constexpr int DF_BB_THREADS = 320;
enum { BB_STAT = 0 /* [2][4]: used,total,first */, BB_KRAFT = 8 /* [2 sweeps][2 alph][16] */, BB_K = 72 /* [2][16] */,
       BB_NEXT = 104 /* [2][16] */, BB_SLACK = 136 /* [2] */, BB_CNTW = 144 /* [2 bufs][10 warps][16] */ };

__device__ __forceinline__ void bb_count_lengths(uint32_t *cntw_row, uint32_t L, unsigned lane, unsigned &m) {
    if (lane < 16) cntw_row[lane] = 0;
    __syncwarp();
    m = __match_any_sync(MZ_FULL_MASK, L);
    if (L > 0 && lane == (unsigned)(__ffs((int)m) - 1)) cntw_row[L] = (uint32_t)__popc(m);
    __syncwarp();
}

__device__ inline void block_build_codes(const uint32_t *hist_ll, const uint32_t *hist_d, uint8_t *lens_ll, uint8_t *lens_d,
                                         uint32_t *code_ll, uint32_t *code_d, uint32_t *bb) {
    const uint32_t tid = threadIdx.x;
    const unsigned lane = lane_id(), w = warp_id();
    const int a = tid < 288 ? 0 : (tid < 320 ? 1 : 2);
    const uint32_t sym = a == 0 ? tid : tid - 288;
    const uint32_t nsym = a == 0 ? 286u : 30u;
    const int M = 15;
    const uint32_t one = 1u << M;
    const uint32_t w0 = a == 0 ? 0u : 9u; /* first warp of my alphabet */
    uint32_t c = (a < 2 && sym < nsym) ? (a == 0 ? hist_ll[sym] : hist_d[sym]) : 0u;

    if (tid < 144) bb[tid] = (tid == BB_STAT + 2 || tid == BB_STAT + 6) ? 0xffffffffu : 0u;
    bar_sync(1, DF_BB_THREADS);
    if (a < 2) {
        uint32_t used_w = __reduce_add_sync(MZ_FULL_MASK, c ? 1u : 0u);
        uint32_t tot_w = __reduce_add_sync(MZ_FULL_MASK, c);
        uint32_t first_w = __reduce_min_sync(MZ_FULL_MASK, c ? sym : 0xffffffffu);
        if (lane == 0) {
            atomicAdd(&bb[BB_STAT + a * 4 + 0], used_w);
            atomicAdd(&bb[BB_STAT + a * 4 + 1], tot_w);
            atomicMin(&bb[BB_STAT + a * 4 + 2], first_w);
        }
    }
    bar_sync(1, DF_BB_THREADS);
    float ideal = 0.f;
    if (a < 2) {
        uint32_t used = bb[BB_STAT + a * 4 + 0], total = bb[BB_STAT + a * 4 + 1], fu = bb[BB_STAT + a * 4 + 2];
        if (used < 2) { /* force two coded symbols (a complete code needs them; zlib's inflate insists) */
            uint32_t d0 = used == 0 ? 0u : (fu == 0 ? 1u : 0u);
            uint32_t d1 = used == 0 ? 1u : d0;
            if ((sym == d0 || sym == d1) && c == 0) c = 1;
            total += 2 - used;
        }
        if (c) ideal = __log2f((float)total) - __log2f((float)c);
    }
    /* two sweeps of 16 candidate rounding offsets: coarse step 1/4 over [-3, 1), then step 1/64 */
    float lo = -3.0f, step = 0.25f;
    uint32_t slack[2] = {0, 0};
    for (int sweep = 0; sweep < 2; sweep++) {
        if (a < 2) {
#pragma unroll
            for (int cnd = 0; cnd < 16; cnd++) {
                uint32_t kr = 0;
                if (c) {
                    int l = (int)ceilf(ideal - (lo + step * (float)cnd));
                    l = l < 1 ? 1 : (l > M ? M : l);
                    kr = 1u << (M - l);
                }
                kr = __reduce_add_sync(MZ_FULL_MASK, kr);
                if (lane == 0) atomicAdd(&bb[BB_KRAFT + sweep * 32 + a * 16 + cnd], kr);
            }
        }
        bar_sync(1, DF_BB_THREADS);
        /* everyone derives both alphabets' choices (uniform): the largest feasible candidate */
        float lo_a = lo;
        for (int aa = 0; aa < 2; aa++) {
            int best = 0;
            for (int cnd = 1; cnd < 16; cnd++)
                if (bb[BB_KRAFT + sweep * 32 + aa * 16 + cnd] <= one) best = cnd;
            slack[aa] = one - bb[BB_KRAFT + sweep * 32 + aa * 16 + best];
            if (aa == a) lo_a = lo + step * (float)best;
        }
        /* lo is per alphabet from here on; the loop variable is private to the thread */
        lo = lo_a;
        step *= 0.0625f;
    }
    uint32_t L = 0;
    if (c) {
        int l = (int)ceilf(ideal - lo);
        L = (uint32_t)(l < 1 ? 1 : (l > M ? M : l));
    }
    /* exact Kraft completion: shorten codes (short ones first) until the sum is exactly 1 */
    unsigned m = 0;
    int buf = 0;
    for (int pass = 0; pass < 40; pass++) {
        if (slack[0] == 0 && slack[1] == 0) break;
        uint32_t *cntw = bb + BB_CNTW + buf * 160;
        if (a < 2) bb_count_lengths(cntw + w * 16, L, lane, m);
        bar_sync(1, DF_BB_THREADS);
        if (w == 0 || w == 9) { /* one warp per alphabet: lane = code length; the recurrence over lengths is uniform */
            const int aa = w == 0 ? 0 : 1;
            const uint32_t wa = aa == 0 ? 0u : 9u, wb = aa == 0 ? 9u : 10u;
            uint32_t tot = 0;
            if (lane >= 2 && lane <= (unsigned)M)
                for (uint32_t ww = wa; ww < wb; ww++) tot += cntw[ww * 16 + lane];
            uint32_t sl = slack[aa], myk = 0;
            for (int len = 2; len <= M; len++) {
                uint32_t t = __shfl_sync(MZ_FULL_MASK, tot, len);
                uint32_t can = sl >> (M - len);
                uint32_t k = t < can ? t : can;
                if (lane == (unsigned)len) myk = k;
                sl -= k << (M - len);
            }
            if (lane >= 2 && lane <= (unsigned)M) bb[BB_K + aa * 16 + lane] = myk;
            if (lane == 0) bb[BB_SLACK + aa] = sl;
        }
        bar_sync(1, DF_BB_THREADS);
        if (a < 2 && L >= 2) {
            uint32_t rank = (uint32_t)__popc(m & ((1u << lane) - 1));
            for (uint32_t ww = w0; ww < w; ww++) rank += cntw[ww * 16 + L];
            if (rank < bb[BB_K + a * 16 + L]) L -= 1;
        }
        slack[0] = bb[BB_SLACK + 0];
        slack[1] = bb[BB_SLACK + 1];
        buf ^= 1;
    }
    /* canonical codes: code = first code of its length + rank among equal lengths in symbol order */
    uint32_t *cntw = bb + BB_CNTW + buf * 160;
    if (a < 2) bb_count_lengths(cntw + w * 16, L, lane, m);
    bar_sync(1, DF_BB_THREADS);
    if (w == 0 || w == 9) {
        const int aa = w == 0 ? 0 : 1;
        const uint32_t wa = aa == 0 ? 0u : 9u, wb = aa == 0 ? 9u : 10u;
        uint32_t tot = 0;
        if (lane >= 1 && lane <= (unsigned)M)
            for (uint32_t ww = wa; ww < wb; ww++) tot += cntw[ww * 16 + lane];
        uint32_t code = 0, prev = 0, mycode = 0;
        for (int len = 1; len <= M; len++) {
            code = (code + prev) << 1;
            if (lane == (unsigned)len) mycode = code;
            prev = __shfl_sync(MZ_FULL_MASK, tot, len);
        }
        if (lane >= 1 && lane <= (unsigned)M) bb[BB_NEXT + aa * 16 + lane] = mycode;
    }
    bar_sync(1, DF_BB_THREADS);
    if (a < 2) {
        uint32_t cw = 0;
        if (L) {
            uint32_t rank = (uint32_t)__popc(m & ((1u << lane) - 1));
            for (uint32_t ww = w0; ww < w; ww++) rank += cntw[ww * 16 + L];
            uint32_t code = bb[BB_NEXT + a * 16 + L] + rank;
            cw = (__brev(code) >> (32 - L)) | (L << 16);
        }
        if (a == 0) {
            lens_ll[sym] = (uint8_t)L; /* sym up to 287: entries 286, 287 get 0 */
            code_ll[sym] = cw;
        } else {
            lens_d[sym] = (uint8_t)L;  /* sym up to 31 */
            code_d[sym] = cw;
        }
    }
}

/* Dynamic block header with a FIXED code-length code (all sixteen length values 0..15 coded in 4 bits:
 * a complete prefix code, so zlib accepts it) and no run-length symbols: constant size, every field at a
 * known bit position, written by 317 threads in parallel. Costs ~30-60 bytes per block against an optimal
 * code-length code; buys the removal of a serial warp-level code construction from the critical path. */
constexpr uint32_t DF_HDR_BITS = 17 + 19 * 3 + 4 * (286 + 30);
__device__ __forceinline__ void emit_block_header(uint32_t *stage, uint32_t pos, uint32_t bfinal, const uint8_t *lens_ll,
                                                  const uint8_t *lens_d, uint32_t tid) {
    if (tid == 0) stage_put(stage, pos, bfinal | (2u << 1) | (29u << 3) | (29u << 8) | (15u << 13), 17);
    /* code-length code lengths in the order 16,17,18,0,8,7,...: three zeros, then sixteen 4s */
    if (tid >= 1 && tid <= 16) stage_put(stage, pos + 17 + 9 + 3 * (tid - 1), 4u, 3);
    if (tid < 286 + 30) {
        uint32_t v = tid < 286 ? lens_ll[tid] : lens_d[tid - 286];
        stage_put(stage, pos + 17 + 57 + 4 * tid, __brev(v) >> 28, 4); /* canonical 4-bit code of value v, MSB first */
    }
}

/* ---- A: match finding ---------------------------------------------------------------------------------- */
/* longest match of in[p..] against in[cand..], both inside the chunk, at most maxlen bytes;
 * first 4 bytes already known equal */
__device__ __forceinline__ uint32_t extend_match(const Smem &sm, uint32_t cand, uint32_t p, uint32_t maxlen) {
    uint32_t len = 4;
    while (len < maxlen) {
        uint32_t x = sm.ld32u(DF_OFF_IN, p + len) ^ sm.ld32u(DF_OFF_IN, cand + len);
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
template <int WAYS>
__device__ __forceinline__ uint32_t find_match(const Smem &sm, uint32_t p, uint32_t limit, uint32_t &best_dist) {
    uint32_t v = sm.ld32u(DF_OFF_IN, p);
    uint32_t maxlen = limit - p;
    if (maxlen > 258) maxlen = 258;
    uint32_t best = 0;
    best_dist = 0;
    if (WAYS == 1) {
        uint32_t ha = DF_OFF_HASH + hash4(v, 14) * 2;
        uint32_t cand = sm.ld16(ha);
        sm.st16(ha, p);
        if (cand < p && p - cand <= 32768u && sm.ld32u(DF_OFF_IN, cand) == v) {
            best = extend_match(sm, cand, p, maxlen);
            best_dist = p - cand;
        }
    } else {
        constexpr int hb = WAYS == 2 ? 13 : 12;
        uint32_t ha = DF_OFF_HASH + hash4(v, hb) * (uint32_t)(WAYS * 2);
        uint32_t prev_slot = p;
#pragma unroll
        for (int wy = 0; wy < WAYS; wy++) {
            uint32_t cand = sm.ld16(ha + 2 * wy);
            sm.st16(ha + 2 * wy, prev_slot); /* FIFO: newest first */
            prev_slot = cand;
            if (cand < p && p - cand <= 32768u && sm.ld32u(DF_OFF_IN, cand) == v) {
                uint32_t l = extend_match(sm, cand, p, maxlen);
                if (l > best) { best = l; best_dist = p - cand; }
            }
        }
    }
    return best;
}

__device__ __forceinline__ uint32_t bit_range(uint32_t a, uint32_t b) { /* bits [a, b), 0 <= a, b <= 32 */
    uint32_t hi = b >= 32 ? 0xffffffffu : (1u << b) - 1;
    uint32_t lo = a >= 32 ? 0xffffffffu : (1u << a) - 1;
    return hi & ~lo;
}

/* ---- the kernel ------------------------------------------------------------------------------- */
template <int WAYS, bool LAZY>
__global__ void __launch_bounds__(DF_THREADS, 1) deflate_chunks_kernel(DeflateParams P) {
    MZ_DYN_SMEM(smem);
    uint8_t *s_in = smem + DF_OFF_IN;
    uint16_t *s_hash = (uint16_t *)(smem + DF_OFF_HASH);
    uint32_t *s_rec = (uint32_t *)(smem + DF_OFF_REC);
    uint32_t *s_stage = (uint32_t *)(smem + DF_OFF_STAGE);
    uint16_t *s_tok = (uint16_t *)(smem + DF_OFF_TOK);
    uint32_t *s_hist_ll = (uint32_t *)(smem + DF_OFF_HIST);
    uint32_t *s_hist_d = s_hist_ll + 288;
    uint32_t *s_code_ll = (uint32_t *)(smem + DF_OFF_CODE);
    uint32_t *s_code_d = s_code_ll + 288;
    uint8_t *s_lens_ll = smem + DF_OFF_LENS;
    uint8_t *s_lens_d = s_lens_ll + 288;
    uint32_t *s_scan = (uint32_t *)(smem + DF_OFF_SCAN);
    uint32_t *s_bb = (uint32_t *)(smem + DF_OFF_BB);
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
    Smem sm;
    sm.init(smem);

    /* Chunks are handed out by an atomic counter: if another kernel (e.g. an NCCL all-gather overlapped with this
     * one) holds some SMs, the resident CTAs simply take more chunks instead of leaving a tail to late CTAs. */
    for (uint32_t chunk_static = blockIdx.x;; chunk_static += gridDim.x) {
        uint32_t chunk = chunk_static;
        if (P.work_counter) {
            if (tid == 0) s_misc[MISC_CHUNK] = atomicAdd(P.work_counter, 1u);
            __syncthreads();
            chunk = s_misc[MISC_CHUNK];
        }
        if (chunk >= P.nchunks) break;
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
            const uint32_t seg_start = sb_start + tid * DF_SEG < sb_end ? sb_start + tid * DF_SEG : sb_end;
            const uint32_t seg_end = seg_start + DF_SEG < sb_end ? seg_start + DF_SEG : sb_end;
            const bool last_sb = sb == nsb - 1;
            uint32_t nhalf = 1, n_tok = 0, ntok_all = 0, tok_excl = 0, e0 = sb_end;
            uint32_t litmask = 0, keptrec = 0, recmask = 0, strad_cnt = 0, strad_pos = 0, strad_tok = 0;
            bool strad_match = false;

            if (P.level != 0) {
                /* ---- A: parse ------------------------------------------------------------------ */
                uint32_t nrec = 0;
                uint32_t p = seg_start;
                while (p < seg_end) {
                    uint32_t mlen = 0, mdist = 0;
                    if (p + DF_MINMATCH <= sb_end) {
                        mlen = find_match<WAYS>(sm, p, sb_end, mdist);
                        if (LAZY && mlen >= DF_MINMATCH && mlen < 32 && p + 1 < seg_end && p + 1 + DF_MINMATCH <= sb_end) {
                            uint32_t d2, l2 = find_match<WAYS>(sm, p + 1, sb_end, d2);
                            if (l2 > mlen) { /* literal now, better match next */
                                p += 1;
                                mlen = l2;
                                mdist = d2;
                            }
                        }
                    }
                    /* at most 8 matches of >= 4 bytes start inside a 32-byte span, so nrec cannot overflow */
                    const bool ism = mlen >= DF_MINMATCH;
                    if (ism) sm.st32(DF_OFF_REC + (nrec * DF_THREADS + tid) * 4, (p - seg_start) | ((mlen - 3) << 5) | ((mdist - 1) << 13));
                    nrec += ism ? 1u : 0u;
                    p += ism ? mlen : 1u;
                }
                /* ---- B: cover = where earlier threads' matches end ------------------------------ */
                const uint32_t cover = block_excl_max(p, sb_start, s_scan);
                if (tid == DF_THREADS / 2) s_misc[MISC_E0] = cover; /* where the second thread half starts */
                /* ---- T: classify what this thread keeps ------------------------------------------- */
                const uint32_t c_rel = cover > seg_start ? cover - seg_start : 0u; /* may exceed 32 */
                uint32_t covmask = 0;
                recmask = 0;
                for (uint32_t r = 0; r < nrec; r++) {
                    uint32_t rec = s_rec[r * DF_THREADS + tid];
                    uint32_t roff = rec & 31, rlen = ((rec >> 5) & 255) + 3;
                    uint32_t rend = roff + rlen;
                    recmask |= 1u << roff;
                    covmask |= bit_range(roff + 1, rend < 32 ? rend : 32);
                    if (rend <= c_rel) continue; /* covered entirely by an earlier thread's match */
                    if (roff >= c_rel) {
                        keptrec |= 1u << roff;
                    } else {
                        uint32_t rem = rend - c_rel; /* straddles: trim the front */
                        if (rem >= 3) {
                            strad_match = true;
                            strad_tok = DF_TOK_MATCH | (r * DF_THREADS + tid);
                            s_rec[r * DF_THREADS + tid] = roff | ((rem - 3) << 5) | (rec & ~0x1fffu);
                        } else {
                            strad_cnt = rem; /* 1-2 bytes left: plain literals */
                            strad_pos = seg_start + c_rel;
                        }
                    }
                }
                const uint32_t valid = seg_end - seg_start;
                litmask = ~covmask & ~recmask & ~bit_range(0, c_rel < 32 ? c_rel : 32) & bit_range(0, valid);
                n_tok = (uint32_t)__popc(litmask) + 2u * ((uint32_t)__popc(keptrec) + (strad_match ? 1u : 0u)) + strad_cnt; /* a match = 2 code words */
                tok_excl = block_excl_sum(n_tok, s_scan, ntok_all);
                e0 = s_misc[MISC_E0];
                nhalf = ntok_all > (uint32_t)DF_TOK_CAP ? 2u : 1u;
            }

            for (uint32_t half = 0; half < nhalf; half++) {
                /* byte range this block codes: the kept tokens of the participating threads tile it exactly */
                const uint32_t blk_start = (nhalf == 2 && half == 1) ? e0 : sb_start;
                const uint32_t blk_end = (nhalf == 2 && half == 0) ? e0 : sb_end;
                const uint32_t blk_len = blk_end - blk_start;
                const uint32_t bfinal = (last_sb && half == nhalf - 1 && (flags & DF_FLAG_FINAL)) ? 1u : 0u;
                bool stored = (P.level == 0);

                if (!stored) {
                    uint32_t ntok = ntok_all, excl = tok_excl;
                    const bool mine = nhalf == 1 || (tid >> 9) == half;
                    if (nhalf == 2) excl = block_excl_sum(mine ? n_tok : 0u, s_scan, ntok);
                    /* ---- T: write the ordered token list ------------------------------------------ */
                    for (uint32_t i = tid; i < 288 + 32; i += DF_THREADS) s_hist_ll[i] = 0;
                    __syncthreads();
                    /* literal entry = the byte itself; match entry = flag | record slot (r * 1024 + owner) */
                    if (mine) {
                        uint32_t o = excl;
                        if (strad_match) {
                            s_tok[o++] = (uint16_t)strad_tok;
                            s_tok[o++] = (uint16_t)(strad_tok | DF_TOK_DIST);
                        }
                        for (uint32_t q = 0; q < strad_cnt; q++) s_tok[o++] = s_in[strad_pos + q];
                        uint32_t mm = litmask | keptrec;
                        uint32_t oa = DF_OFF_TOK + o * 2;
                        while (mm) {
                            uint32_t b = (uint32_t)__ffs((int)mm) - 1;
                            mm &= mm - 1;
                            uint32_t e = sm.ld8(DF_OFF_IN + seg_start + b);
                            if ((keptrec >> b) & 1u) {
                                e = DF_TOK_MATCH | ((uint32_t)__popc(recmask & ((1u << b) - 1)) * DF_THREADS + tid);
                                sm.st16(oa, e); /* length code word, then the distance code word */
                                oa += 2;
                                e |= DF_TOK_DIST;
                            }
                            sm.st16(oa, e);
                            oa += 2;
                        }
                        /* ---- C (matches): this thread's kept records -> symbol form + histograms ------------- */
                        uint32_t mk = keptrec;
                        bool do_strad = strad_match;
                        while (mk || do_strad) {
                            uint32_t slot;
                            if (do_strad) {
                                slot = strad_tok & 0x1fffu;
                                do_strad = false;
                            } else {
                                uint32_t b = (uint32_t)__ffs((int)mk) - 1;
                                mk &= mk - 1;
                                slot = (uint32_t)__popc(recmask & ((1u << b) - 1)) * DF_THREADS + tid;
                            }
                            uint32_t rec = s_rec[slot];
                            uint32_t ls, lb, lv, ds, db, dv;
                            length_symbol(((rec >> 5) & 255) + 3, ls, lb, lv);
                            dist_symbol((rec >> 13) + 1, ds, db, dv);
                            s_rec[slot] = ls | (lv << 5) | (ds << 10) | (dv << 15);
                            atomicAdd(&s_hist_ll[257 + ls], 1u);
                            atomicAdd(&s_hist_d[ds], 1u);
                        }
                    }
                    __syncthreads();
                    /* ---- C (literals): threads stride over the list ------------------------------------------ */
                    for (uint32_t k = tid; k < ntok; k += DF_THREADS) {
                        uint32_t e = sm.ld16(DF_OFF_TOK + k * 2);
                        if (!(e & DF_TOK_MATCH)) sm.red_add32(DF_OFF_HIST + e * 4, 1u);
                    }
                    if (tid == 0) s_hist_ll[256] = 1;
                    __syncthreads();
                    /* ---- D: codes (10 warps on a named barrier; the rest wait here) ------------------------- */
                    if (tid < DF_BB_THREADS) block_build_codes(s_hist_ll, s_hist_d, s_lens_ll, s_lens_d, s_code_ll, s_code_d, s_bb);
                    __syncthreads();
                    /* ---- E: bit counts over equal contiguous token runs (codes are final) ---------------- */
                    const uint32_t run = ((ntok + DF_THREADS - 1) / DF_THREADS) | 1u; /* odd stride: no bank conflicts */
                    const uint32_t k0 = tid * run < ntok ? tid * run : ntok;
                    const uint32_t k1 = k0 + run < ntok ? k0 + run : ntok;
                    uint32_t mybits = 0;
                    for (uint32_t k = k0; k < k1; k++) {
                        uint32_t idx, ev, eb;
                        token_code(sm.ld16(DF_OFF_TOK + k * 2), sm, idx, ev, eb);
                        mybits += (sm.ld32(DF_OFF_CODE + idx * 4) >> 16) + eb;
                    }
                    uint32_t tokbits;
                    uint32_t myoff = block_excl_sum(mybits, s_scan, tokbits);
                    const uint32_t hdrbits = DF_HDR_BITS;
                    const uint32_t eob = s_code_ll[256];
                    const uint32_t dyn_bits = hdrbits + tokbits + (eob >> 16);
                    const uint32_t stored_bits = (((bitpos + 3 + 7) & ~7u) - bitpos) + 32 + blk_len * 8;
                    if (dyn_bits >= stored_bits) {
                        stored = true;
                    } else {
                        /* ---- F: emit --------------------------------------------------------------- */
                        const uint32_t base = bitpos + hdrbits;
                        emit_block_header(s_stage, bitpos, bfinal, s_lens_ll, s_lens_d, tid);
                        if (mybits) {
                            BitWriter bw;
                            bw.init(sm, base + myoff);
                            for (uint32_t k = k0; k < k1; k++) {
                                uint32_t idx, ev, eb;
                                token_code(sm.ld16(DF_OFF_TOK + k * 2), sm, idx, ev, eb);
                                uint32_t cw = sm.ld32(DF_OFF_CODE + idx * 4); /* distance codes follow the 288 literal/length codes */
                                uint32_t cl = cw >> 16;
                                bw.put((cw & 0xffff) | (ev << cl), cl + eb); /* <= 15 + 13 bits */
                            }
                            bw.finish();
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
                        stage_put(s_stage, p0 * 8, blk_len, 16);
                        stage_put(s_stage, p0 * 8 + 16, blk_len ^ 0xffffu, 16);
                    }
                    for (uint32_t i = tid * 4; i < blk_len; i += DF_THREADS * 4) {
                        uint32_t v = load32u(s_in, blk_start + i);
                        uint32_t n = blk_len - i;
                        stage_put(s_stage, (p0 + 4 + i) * 8, v, n >= 4 ? 32 : n * 8);
                    }
                    bitpos = (p0 + 4 + blk_len) * 8;
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

/* level -> (bucket ways, lazy) */
__host__ __device__ inline int deflate_ways_for_level(int level) { return level <= 1 ? 1 : (level <= 3 ? 2 : 4); }
__host__ __device__ inline bool deflate_lazy_for_level(int level) { return level >= 6; }

} // namespace mzc
#endif
