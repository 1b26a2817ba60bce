/* inflate_spec_kernel.cuh -- K6: one LONG foreign DEFLATE stream decoded by thousands of warps (sm_100a).
 *
 * What it accelerates: mz_stream_zlib_read over one big member (mz_strm_zlib.c:116-193, config C3: a 4 GiB
 * .gz written by somebody else, no sync points). A DEFLATE stream has no index, so block starts are GUESSED and
 * the guesses are proven afterwards; nothing unproven ever reaches the caller:
 *
 *   K6a find      the compressed window is cut into segments; warp k scans the bit offsets of segment k for the
 *                 first position where a non-final dynamic block header parses into complete Huffman codes
 *                 (RFC1951 3.2.7 + zlib's completeness rules -- the same inf_dynamic_header the decoder uses).
 *                 Segment 0 starts at the known block boundary where the serial decoder stands.
 *   K6b scan      warp k decodes from its start until the first block boundary at or after the next segment's
 *                 start. The 32 KiB of history before its start are unknown, so the output is SYMBOLIC: a ring of
 *                 16-bit symbols, < 256 = literal byte, 0x8000|i = "byte i of my unknown window" (copies of
 *                 unknown bytes stay references). Nothing but the ring, the byte count and the end bit is kept.
 *   K6c chain     one thread walks the segments from the anchored one: segment j joins the chain only if the
 *                 previous chain member ended EXACTLY on j's guessed start (and the output still fits). A wrong
 *                 guess, an error, missing input -- the chain simply ends there and the serial decoder carries on.
 *   K6d resolve   each chain member's symbolic last-32-KiB becomes bytes (member i's window = member i-1's resolved
 *                 final window): windows are maps and maps compose, so the chain is folded group-wise in parallel,
 *                 linked across groups serially (96 steps), then expanded group-wise in parallel.
 *   K6e emit      warp per chain member decodes its segment AGAIN, now with a resolved window and a known output
 *                 offset, straight into the caller-visible output window; it must reproduce K6b's end bit and
 *                 byte count or the whole round is discarded.
 *
 * Decoding twice costs instructions but needs no symbolic copy of the whole output and no per-segment output
 * capacity guess (deflate can expand 1032:1). Traffic per output byte: K6b ~2 B ring (L2-resident), K6e 1 B
 * written + the compressed bytes read twice.
 */
#ifndef MZ_INFLATE_SPEC_KERNEL_CUH
#define MZ_INFLATE_SPEC_KERNEL_CUH

#include "inflate_kernel.cuh"

namespace mzc {

constexpr uint64_t SPEC_NONE = ~0ull;
constexpr uint32_t SPEC_RING = 65536;  /* symbols per segment ring */
constexpr int SPEC_RESOLVE_THREADS = 1024;
enum { SPEC_FLAG_MISMATCH = 1 };

struct SpecSeg {
    uint64_t start_bit; /* absolute bit of the (guessed) block start, SPEC_NONE if none was found */
    uint64_t end_bit;   /* absolute bit where the scan stopped */
    uint64_t out_count; /* bytes the segment produces */
    uint64_t out_off;   /* chain: bytes produced by the chain members before this one */
    int32_t status, why;
    uint32_t blocks, pad;
};

struct SpecSummary {
    uint64_t end_bit;    /* absolute bit position after the last chain member */
    uint64_t total_out;  /* bytes emitted by the round */
    uint32_t nchain;     /* segments accepted */
    int32_t status;      /* INF_ST_RUN or INF_ST_END */
    uint32_t blocks;
    uint32_t flags;      /* SPEC_FLAG_* : the round must be discarded */
    uint32_t candidates; /* segments in which a block start was found */
    uint32_t pad;
};

struct SpecParams {
    const uint8_t *in;   /* compressed window, 4-byte aligned, >= 16 readable bytes past in_avail */
    uint64_t in_base;    /* absolute stream byte of in[0] */
    uint64_t in_avail;
    uint64_t start_bit;  /* absolute bit of the block boundary where the serial decoder stands */
    uint64_t seg_bits;   /* segment length in bits */
    uint8_t *out;        /* output window; out[0] is absolute output byte out_base */
    uint64_t out_base;
    uint64_t out_pos;    /* absolute output position where this round starts (history before it is in `out`) */
    uint64_t out_end;    /* absolute end of the writable output window */
    uint32_t in_final;
    uint32_t nseg;
    SpecSeg *seg;        /* [nseg] */
    InflateState *states; /* [nseg] */
    uint16_t *rings;     /* [nseg][SPEC_RING] */
    uint8_t *wins;       /* [nseg][32768] */
    uint32_t *chain;     /* [nseg] pairs {segment, bytes produced mod 2^32}, chain order */
    SpecSummary *summary;
    uint16_t *gmaps;     /* [SPEC_GROUPS][32768] group maps (K6d compose) */
    uint8_t *gwins;      /* [SPEC_GROUPS][32768] real window after each group (K6d link) */
    uint32_t *work;      /* [4] work counters (zeroed per round): [0] scan, [1] emit */
};

/* ---- K6a ------------------------------------------------------------------------------------------------------ */
constexpr int SPEC_FIND_SMEM = INF_SMEM_BYTES + 256 * 2; /* decode tables + a queue of 256 surviving bit offsets */

/* first 17 bits of a would-be header: BFINAL=0, BTYPE=2, HLIT <= 29, HDIST <= 29 */
__device__ __forceinline__ bool spec_head_ok(uint32_t h) {
    return (h & 7u) == 4u && ((h >> 3) & 31u) <= 29u && ((h >> 8) & 31u) <= 29u;
}
/* the code-length code announced at this offset must be COMPLETE (zlib rejects anything else, inftrees.c CODES) */
__device__ __forceinline__ bool spec_kraft_ok(const uint32_t *w, uint64_t relbit) {
    const uint64_t i = relbit >> 5;
    const uint32_t s = (uint32_t)relbit & 31u;
    const uint32_t w0 = w[i], w1 = w[i + 1], w2 = w[i + 2], w3 = w[i + 3];
    const uint32_t v0 = __funnelshift_r(w0, w1, s), v1 = __funnelshift_r(w1, w2, s), v2 = __funnelshift_r(w2, w3, s);
    const uint32_t hc = ((v0 >> 13) & 15u) + 4u;
    uint64_t x = ((((uint64_t)v1 << 32) | v0) >> 17) | ((uint64_t)v2 << 47);
    uint32_t sum = 0;
#pragma unroll
    for (uint32_t k = 0; k < 19; k++) {
        const uint32_t l = (uint32_t)x & 7u;
        x >>= 3;
        if (k < hc && l) sum += 128u >> l;
    }
    return sum == 128u;
}

__global__ void __launch_bounds__(INF_THREADS) inflate_spec_find_kernel(SpecParams P) {
    MZ_DYN_SMEM(smem);
    InfTables &T = *reinterpret_cast<InfTables *>(smem);
    uint16_t *queue = reinterpret_cast<uint16_t *>(smem + INF_SMEM_BYTES);
    const unsigned lane = lane_id();
    for (uint32_t k = blockIdx.x; k < P.nseg; k += gridDim.x) {
        uint64_t found = SPEC_NONE;
        if (k == 0) {
            found = P.start_bit;
        } else {
            const uint64_t base_abs = P.in_base * 8;
            const uint64_t avail_bits = P.in_avail * 8;
            uint64_t lo = P.start_bit + (uint64_t)k * P.seg_bits - base_abs; /* relative to in[0] */
            uint64_t hi = lo + P.seg_bits;
            const uint64_t limit = avail_bits > 160 ? avail_bits - 160 : 0; /* a header plus its first symbols must lie inside the window */
            if (hi > limit) hi = limit;
            const uint32_t *w = (const uint32_t *)P.in;
            /* 256 bit offsets per step: each lane screens 8 consecutive ones on their first 17 bits; the survivors
             * (~11 %) are queued in offset order so that the 19-field Kraft test runs on full warps */
            for (uint64_t p0 = lo; p0 < hi && found == SPEC_NONE; p0 += 256) {
                const uint64_t q0 = p0 + lane * 8u;
                uint32_t mask = 0;
                if (q0 < hi) {
                    const uint64_t i = q0 >> 5;
                    const uint32_t v = __funnelshift_r(w[i], w[i + 1], (uint32_t)q0 & 31u);
#pragma unroll
                    for (uint32_t j = 0; j < 8; j++)
                        if (spec_head_ok(v >> j) && q0 + j < hi) mask |= 1u << j;
                }
                const uint32_t cnt = (uint32_t)__popc(mask);
                const uint32_t incl = warp_incl_sum(cnt);
                const uint32_t total = __shfl_sync(MZ_FULL_MASK, incl, 31);
                uint32_t at = incl - cnt;
                for (uint32_t mm = mask; mm; mm &= mm - 1) queue[at++] = (uint16_t)(lane * 8u + (uint32_t)__ffs((int)mm) - 1u);
                __syncwarp();
                for (uint32_t r0 = 0; r0 < total && found == SPEC_NONE; r0 += 32) {
                    const uint32_t qi = r0 + lane;
                    const bool cand = qi < total && spec_kraft_ok(w, p0 + queue[qi]);
                    unsigned m = __ballot_sync(MZ_FULL_MASK, cand);
                    while (m && found == SPEC_NONE) { /* the real header parse, lowest offset first */
                        const uint32_t l = (uint32_t)__ffs((int)m) - 1;
                        m &= m - 1;
                        const uint64_t p = p0 + queue[r0 + l];
                        InfBits b;
                        b.init(P.in, P.in_avail, p + 3);
                        uint32_t nlit, ndist;
                        const int err = inf_dynamic_header(b, T, nlit, ndist);
                        __syncwarp();
                        if (!err) found = base_abs + p;
                    }
                }
                __syncwarp();
            }
        }
        if (lane == 0) P.seg[k].start_bit = found;
    }
}

/* ---- K6b ------------------------------------------------------------------------------------------------------ */
/* Where a scanning warp stops: at a block boundary that IS some later segment's guess (the chain can link there).
 * A boundary that is nobody's guess (a segment whose only guess was a false positive, or one without any) is
 * walked through: this warp simply decodes the next block as well. Past the last guess of the window it stops at
 * the first boundary beyond its own segment; what lies further is the next round's business. */
struct SpecStop {
    const SpecSeg *seg;
    uint64_t start0, seg_bits, first_later;
    uint32_t nseg, k;
    __device__ __forceinline__ bool operator()(uint64_t pos) const {
        if (first_later != SPEC_NONE && pos < first_later) return false;
        uint64_t j = (pos - start0) / seg_bits;
        if (j <= k) j = k + 1;
        for (; j < nseg; j++) {
            const uint64_t g = seg[j].start_bit;
            if (g != SPEC_NONE && g >= pos) return g == pos;
        }
        return pos >= start0 + (uint64_t)(k + 1) * seg_bits;
    }
};

__global__ void __launch_bounds__(INF_THREADS) inflate_spec_scan_kernel(SpecParams P) {
    MZ_DYN_SMEM(smem);
    InfTables &T = *reinterpret_cast<InfTables *>(smem);
    const unsigned lane = lane_id();
    for (;;) { /* segments differ a lot in cost: pull them from a counter */
        const uint32_t k = inf_next_work(&P.work[0]);
        if (k >= P.nseg) break;
        SpecSeg *sg = &P.seg[k];
        const uint64_t start = sg->start_bit;
        if (start == SPEC_NONE) {
            if (lane == 0) { sg->status = INF_ST_RUN; sg->why = INF_WHY_NONE; sg->out_count = 0; sg->end_bit = 0; sg->blocks = 0; }
            continue;
        }
        /* the first guess after this segment: nothing can be linked before it */
        uint64_t first_later = SPEC_NONE;
        for (uint32_t j0 = k + 1; j0 < P.nseg && first_later == SPEC_NONE; j0 += 32) {
            const uint32_t j = j0 + lane;
            const uint64_t s = j < P.nseg ? P.seg[j].start_bit : SPEC_NONE;
            const unsigned m = __ballot_sync(MZ_FULL_MASK, s != SPEC_NONE);
            if (m) first_later = __shfl_sync(MZ_FULL_MASK, s, __ffs((int)m) - 1);
        }
        SpecStop stop;
        stop.seg = P.seg;
        stop.nseg = P.nseg;
        stop.k = k;
        stop.start0 = P.start_bit;
        stop.seg_bits = P.seg_bits;
        stop.first_later = first_later;
        uint16_t *ring = P.rings + (size_t)k * SPEC_RING;
        for (uint32_t i = lane; i < 32768; i += 32) ring[32768 + i] = (uint16_t)(0x8000u | i); /* positions -32768..-1 */
        InflateState *st = &P.states[k];
        if (lane == 0) {
            st->in_bitpos = start;
            st->out_pos = 0;
            st->status = INF_ST_RUN;
            st->why = INF_WHY_NONE;
            st->phase = INF_PH_HEADER;
            st->last_block = 0;
            st->stored_remaining = 0;
            st->nlit = st->ndist = 0;
            st->blocks = 0;
        }
        __syncwarp();
        InflateJob job;
        job.in = P.in;
        job.in_base = P.in_base;
        job.in_avail = P.in_avail;
        job.out = nullptr;
        job.out_base = 0;
        job.out_cap = 1ull << 62; /* the ring never fills */
        job.in_final = P.in_final;
        job.flags = 0;
        OutSymRing o;
        o.ring = ring;
        inf_decode_window(job, st, T, o, stop);
        if (lane == 0) {
            sg->end_bit = st->in_bitpos;
            sg->out_count = st->out_pos;
            sg->status = st->status;
            sg->why = st->why;
            sg->blocks = st->blocks;
        }
        __syncwarp();
    }
}

/* ---- K6c ------------------------------------------------------------------------------------------------------ */
/* One CTA. Parallel part: every segment decides whether its scan ended on a block boundary and which later segment
 * (if any) starts exactly there; the verdicts go to shared memory. Serial part: thread 0 follows the links from the
 * anchored segment -- a walk over shared memory, a few nanoseconds per member -- adding up output offsets until a
 * link is missing, the output window is full, or the stream ends. chain[] holds {segment, byte count mod 2^32} pairs. */
constexpr int SPEC_CHAIN_THREADS = 1024;
constexpr uint32_t SPEC_LINK_NONE = 0xffffffffu, SPEC_LINK_END = 0xfffffffeu, SPEC_LINK_BROKEN = 0xfffffffdu;
constexpr uint32_t SPEC_MAX_SEGMENTS = 12288; /* 16 bytes of shared memory per segment in K6c */
__global__ void __launch_bounds__(SPEC_CHAIN_THREADS) inflate_spec_chain_kernel(SpecParams P) {
    MZ_DYN_SMEM(smem);
    uint64_t *s_cnt = reinterpret_cast<uint64_t *>(smem);
    uint32_t *s_link = reinterpret_cast<uint32_t *>(smem + (size_t)P.nseg * 8);
    uint32_t *s_blocks = s_link + P.nseg;
    __shared__ uint32_t s_cand;
    if (threadIdx.x == 0) s_cand = 0;
    __syncthreads();
    uint32_t cand = 0;
    for (uint32_t k = threadIdx.x; k < P.nseg; k += blockDim.x) {
        const SpecSeg sg = P.seg[k];
        cand += sg.start_bit != SPEC_NONE;
        uint32_t link = SPEC_LINK_BROKEN;
        if (sg.start_bit != SPEC_NONE) {
            if (sg.status == INF_ST_END) link = SPEC_LINK_END;
            else if (sg.status == INF_ST_RUN && sg.why == INF_WHY_BOUNDARY) {
                link = SPEC_LINK_NONE; /* whole, but maybe nobody starts where it ends */
                uint32_t j = k + 1;    /* guesses the scan walked through (false positives) lie before its end bit */
                while (j < P.nseg) {
                    const uint64_t g = P.seg[j].start_bit;
                    if (g != SPEC_NONE && g >= sg.end_bit) break;
                    j++;
                }
                if (j < P.nseg && P.seg[j].start_bit == sg.end_bit) link = j;
            }
        }
        s_link[k] = link;
        s_cnt[k] = sg.out_count;
        s_blocks[k] = sg.blocks;
    }
    if (cand) atomicAdd(&s_cand, cand);
    __syncthreads();
    if (threadIdx.x != 0) return;
    SpecSummary s;
    s.end_bit = P.start_bit;
    s.total_out = 0;
    s.nchain = 0;
    s.status = INF_ST_RUN;
    s.blocks = 0;
    s.flags = 0;
    s.candidates = s_cand;
    s.pad = 0;
    uint32_t cur = 0, lastk = SPEC_LINK_NONE;
    while (cur < P.nseg) {
        const uint32_t link = s_link[cur];
        if (link == SPEC_LINK_BROKEN) break;
        const uint64_t cnt = s_cnt[cur];
        if (P.out_pos + s.total_out + cnt > P.out_end) break;
        P.seg[cur].out_off = s.total_out;
        P.chain[2 * s.nchain] = cur;
        P.chain[2 * s.nchain + 1] = (uint32_t)cnt;
        s.nchain++;
        s.total_out += cnt;
        s.blocks += s_blocks[cur];
        lastk = cur;
        if (link == SPEC_LINK_END) { s.status = INF_ST_END; break; }
        if (link == SPEC_LINK_NONE) break; /* nobody started where the stream really continues */
        cur = link;
    }
    if (lastk != SPEC_LINK_NONE) s.end_bit = P.seg[lastk].end_bit;
    *P.summary = s;
}

/* ---- K6d ------------------------------------------------------------------------------------------------------ */
/* Turning symbolic windows into bytes is a chain: member i's window needs member i-1's. A window is a MAP over the
 * window before it (each of its 32768 symbols is a literal or an index into the earlier window), and maps compose,
 * so the chain is cut into SPEC_GROUPS groups and resolved in three short steps instead of one long serial one:
 *   compose  (one CTA per group, parallel): fold the group's members into ONE map over the window before the group,
 *   link     (one CTA): walk the groups -- SPEC_GROUPS steps -- turning each group map into the real bytes of the
 *            window after that group,
 *   resolve  (one CTA per group, parallel): walk the group's members again, now from real bytes, writing every
 *            member's initial window for K6e.
 * Per step: the member's last 32768 ring symbols are staged into shared memory with aligned 16-byte loads (the
 * slice starts anywhere in the ring); the loads for step i+1 are issued BEFORE step i's lookups and land in
 * registers, so the global round trip hides behind the arithmetic of the current step; each thread then handles
 * 8 x 4 symbols against the previous window held in shared memory. */
constexpr uint32_t SPEC_GROUPS = 96;
constexpr int SPEC_SLICE_VECS = 32768 / 8 + 1; /* 16-byte vectors per slice */
constexpr int SPEC_RESOLVE_SMEM = 32768 * 2 + (32768 + 16) * 2;   /* two byte windows + slice */
constexpr int SPEC_COMPOSE_SMEM = 65536 * 2 + (32768 + 16) * 2;   /* two symbol windows + slice */

__device__ __forceinline__ void spec_group_range(uint32_t n, uint32_t g, uint32_t &lo, uint32_t &hi) {
    const uint32_t m = (n + SPEC_GROUPS - 1) / SPEC_GROUPS;
    lo = g * m;
    hi = lo + m < n ? lo + m : n;
    if (lo > n) lo = n;
}
__device__ __forceinline__ void spec_load_slice(const SpecParams &P, uint32_t k, uint32_t cnt, uint4 (&r)[5]) {
    const uint16_t *ring = P.rings + (size_t)k * SPEC_RING;
    const uint32_t tail8 = (cnt - 32768u) & ~7u;
#pragma unroll
    for (int t = 0; t < 5; t++) {
        const uint32_t v = threadIdx.x + (uint32_t)t * SPEC_RESOLVE_THREADS;
        if (v < SPEC_SLICE_VECS) r[t] = *reinterpret_cast<const uint4 *>(ring + ((tail8 + 8u * v) & (SPEC_RING - 1)));
    }
}
__device__ __forceinline__ void spec_store_slice(uint16_t *slice, const uint4 (&r)[5]) {
#pragma unroll
    for (int t = 0; t < 5; t++) {
        const uint32_t v = threadIdx.x + (uint32_t)t * SPEC_RESOLVE_THREADS;
        if (v < SPEC_SLICE_VECS) reinterpret_cast<uint4 *>(slice)[v] = r[t];
    }
}
/* the real window before absolute output position P.out_pos (zeros where the stream has no history yet) */
__device__ __forceinline__ uint8_t spec_history_byte(const SpecParams &P, uint32_t j) {
    const uint64_t back = 32768 - j;
    return back <= P.out_pos - P.out_base && back <= P.out_pos ? P.out[P.out_pos - back - P.out_base] : (uint8_t)0;
}

/* compose: the group's members folded into one symbolic map (16-bit symbols over the window before the group) */
__global__ void __launch_bounds__(SPEC_RESOLVE_THREADS) inflate_spec_compose_kernel(SpecParams P) {
    MZ_DYN_SMEM(smem);
    uint16_t *ca = reinterpret_cast<uint16_t *>(smem), *cb = ca + 32768;
    uint16_t *slice = cb + 32768;
    const uint32_t n = P.summary->nchain;
    uint32_t lo, hi;
    spec_group_range(n, blockIdx.x, lo, hi);
    if (lo >= hi) return;
    uint32_t k = P.chain[2 * lo], cnt = P.chain[2 * lo + 1];
    uint4 r[5];
    spec_load_slice(P, k, cnt, r);
    for (uint32_t j = threadIdx.x; j < 32768; j += blockDim.x) ca[j] = (uint16_t)(0x8000u | j); /* identity */
    for (uint32_t i = lo; i < hi; i++) {
        spec_store_slice(slice, r);
        const uint32_t a = (cnt - 32768u) & 7u;
        __syncthreads();
        if (i + 1 < hi) {
            k = P.chain[2 * (i + 1)];
            cnt = P.chain[2 * (i + 1) + 1];
            spec_load_slice(P, k, cnt, r);
        }
        uint32_t *sd = reinterpret_cast<uint32_t *>(cb);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const uint32_t e0 = threadIdx.x * 4u + 4096u * q;
            uint32_t w01, w23;
            {
                const uint32_t s0 = slice[a + e0], s1 = slice[a + e0 + 1], s2 = slice[a + e0 + 2], s3 = slice[a + e0 + 3];
                const uint32_t v0 = s0 & 0x8000u ? ca[s0 & 0x7fffu] : s0, v1 = s1 & 0x8000u ? ca[s1 & 0x7fffu] : s1;
                const uint32_t v2 = s2 & 0x8000u ? ca[s2 & 0x7fffu] : s2, v3 = s3 & 0x8000u ? ca[s3 & 0x7fffu] : s3;
                w01 = v0 | (v1 << 16);
                w23 = v2 | (v3 << 16);
            }
            sd[e0 >> 1] = w01;
            sd[(e0 >> 1) + 1] = w23;
        }
        __syncthreads();
        uint16_t *t = ca; ca = cb; cb = t;
    }
    uint4 *dst = reinterpret_cast<uint4 *>(P.gmaps + (size_t)blockIdx.x * 32768);
    for (uint32_t v = threadIdx.x; v < 32768 / 8; v += blockDim.x) dst[v] = reinterpret_cast<const uint4 *>(ca)[v];
}

/* link: real bytes of the window after every group, in group order */
__global__ void __launch_bounds__(SPEC_RESOLVE_THREADS) inflate_spec_link_kernel(SpecParams P) {
    MZ_DYN_SMEM(smem);
    uint8_t *wa = smem, *wb = smem + 32768;
    const uint32_t n = P.summary->nchain;
    if (n == 0) return;
    for (uint32_t j = threadIdx.x; j < 32768; j += blockDim.x) wa[j] = spec_history_byte(P, j);
    __syncthreads();
    for (uint32_t g = 0; g < SPEC_GROUPS; g++) {
        uint32_t lo, hi;
        spec_group_range(n, g, lo, hi);
        if (lo >= hi) break;
        const uint16_t *map = P.gmaps + (size_t)g * 32768;
        uint32_t *gd = reinterpret_cast<uint32_t *>(P.gwins + (size_t)g * 32768);
        uint32_t *sd = reinterpret_cast<uint32_t *>(wb);
        uint2 m[8];
#pragma unroll
        for (int q = 0; q < 8; q++) m[q] = reinterpret_cast<const uint2 *>(map)[threadIdx.x + 1024u * q];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const uint32_t s0 = m[q].x & 0xffffu, s1 = m[q].x >> 16, s2 = m[q].y & 0xffffu, s3 = m[q].y >> 16;
            const uint32_t v0 = s0 & 0x8000u ? wa[s0 & 0x7fffu] : s0 & 0xffu, v1 = s1 & 0x8000u ? wa[s1 & 0x7fffu] : s1 & 0xffu;
            const uint32_t v2 = s2 & 0x8000u ? wa[s2 & 0x7fffu] : s2 & 0xffu, v3 = s3 & 0x8000u ? wa[s3 & 0x7fffu] : s3 & 0xffu;
            const uint32_t acc = v0 | (v1 << 8) | (v2 << 16) | (v3 << 24);
            gd[threadIdx.x + 1024u * q] = acc;
            sd[threadIdx.x + 1024u * q] = acc;
        }
        __syncthreads();
        uint8_t *t = wa; wa = wb; wb = t;
    }
}

/* resolve: every member's initial window, group by group in parallel */
__global__ void __launch_bounds__(SPEC_RESOLVE_THREADS) inflate_spec_resolve_kernel(SpecParams P) {
    MZ_DYN_SMEM(smem);
    uint8_t *wa = smem, *wb = smem + 32768;
    uint16_t *slice = reinterpret_cast<uint16_t *>(smem + 65536); /* 32768 + 8 symbols, starts at an 8-aligned ring index */
    const uint32_t n = P.summary->nchain;
    uint32_t lo, hi;
    spec_group_range(n, blockIdx.x, lo, hi);
    if (lo >= hi) return;
    uint32_t k = P.chain[2 * lo], cnt = P.chain[2 * lo + 1];
    uint32_t knext = lo + 1 < hi ? P.chain[2 * (lo + 1)] : 0, cntnext = lo + 1 < hi ? P.chain[2 * (lo + 1) + 1] : 0;
    uint4 r[5];
    if (lo + 1 < hi) spec_load_slice(P, k, cnt, r);
    /* the window before the group's first member: real history for group 0, the link step's result otherwise */
    const uint8_t *init = blockIdx.x ? P.gwins + (size_t)(blockIdx.x - 1) * 32768 : nullptr;
    for (uint32_t j = threadIdx.x; j < 32768; j += blockDim.x) {
        const uint8_t v = init ? init[j] : spec_history_byte(P, j);
        wa[j] = v;
        P.wins[(size_t)k * 32768 + j] = v;
    }
    for (uint32_t i = lo; i + 1 < hi; i++) {
        spec_store_slice(slice, r); /* member i's slice */
        const uint32_t a = (cnt - 32768u) & 7u;
        __syncthreads();
        /* in flight during the lookups: member i+1's slice and the chain entry after it */
        uint32_t k2 = 0, cnt2 = 0;
        if (i + 2 < hi) {
            k2 = P.chain[2 * (i + 2)];
            cnt2 = P.chain[2 * (i + 2) + 1];
            spec_load_slice(P, knext, cntnext, r);
        }
        uint32_t *gd = reinterpret_cast<uint32_t *>(P.wins + (size_t)knext * 32768);
        uint32_t *sd = reinterpret_cast<uint32_t *>(wb);
#pragma unroll
        for (int q = 0; q < 8; q++) { /* thread t owns symbols 4t..4t+3 of every 4096-symbol row: conflict-light shared accesses */
            const uint32_t e0 = threadIdx.x * 4u + 4096u * q;
            uint32_t acc = 0;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t s0 = slice[a + e0 + u];
                acc |= (s0 & 0x8000u ? (uint32_t)wa[s0 & 0x7fffu] : (s0 & 0xffu)) << (8 * u);
            }
            gd[e0 >> 2] = acc;
            sd[e0 >> 2] = acc;
        }
        __syncthreads();
        uint8_t *t = wa; wa = wb; wb = t;
        k = knext; cnt = cntnext;
        knext = k2; cntnext = cnt2;
    }
}

/* ---- K6e ------------------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(INF_THREADS) inflate_spec_emit_kernel(SpecParams P) {
    MZ_DYN_SMEM(smem);
    InfTables &T = *reinterpret_cast<InfTables *>(smem);
    const unsigned lane = lane_id();
    const uint32_t n = P.summary->nchain;
    for (;;) {
        const uint32_t i = inf_next_work(&P.work[1]);
        if (i >= n) break;
        const uint32_t k = P.chain[2 * i];
        const SpecSeg sg = P.seg[k];
        const uint64_t first = P.out_pos + sg.out_off;
        InflateState *st = &P.states[k];
        __syncwarp();
        if (lane == 0) {
            st->in_bitpos = sg.start_bit;
            st->out_pos = first;
            st->status = INF_ST_RUN;
            st->why = INF_WHY_NONE;
            st->phase = INF_PH_HEADER;
            st->last_block = 0;
            st->stored_remaining = 0;
            st->nlit = st->ndist = 0;
            st->blocks = 0;
        }
        __syncwarp();
        InflateJob job;
        job.in = P.in;
        job.in_base = P.in_base;
        job.in_avail = P.in_avail;
        job.out = P.out;
        job.out_base = P.out_base;
        job.out_cap = first + sg.out_count - P.out_base; /* exactly this member's bytes */
        job.in_final = P.in_final;
        job.flags = 0;
        OutBytesWin o;
        o.base = P.out - P.out_base;
        o.win = P.wins + (size_t)k * 32768;
        o.floor = first;
        inf_decode_window(job, st, T, o, StopAtBit{sg.end_bit});
        if (lane == 0) {
            const bool same = st->in_bitpos == sg.end_bit && st->out_pos == first + sg.out_count && st->status == sg.status &&
                              (sg.status == INF_ST_END || st->why == INF_WHY_BOUNDARY);
            if (!same) atomicOr(&P.summary->flags, (uint32_t)SPEC_FLAG_MISMATCH);
        }
        __syncwarp();
    }
}

} // namespace mzc
#endif
