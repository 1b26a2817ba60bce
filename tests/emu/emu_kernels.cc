/* emu_kernels.cc -- runs the product's kernel SOURCES on the CPU emulator (TEST INFRASTRUCTURE ONLY).
 * Built as tests/emu/libmzemu.so by tests/emu/Makefile; used by tests/test_emu_*.py to debug kernel
 * logic without a GPU. Never part of the product library. */
#define MZ_EMU 1
#include "../../minizip-ng_b200/csrc/concat_kernel.cuh"
#include "../../minizip-ng_b200/csrc/crc32_kernel.cuh"
#include "../../minizip-ng_b200/csrc/deflate_kernel.cuh"
#include "../../minizip-ng_b200/csrc/inflate_kernel.cuh"
#include "../../minizip-ng_b200/csrc/inflate_spec_kernel.cuh"
#include "../../minizip-ng_b200/csrc/sha256_kernel.cuh"

using namespace mzc;

static CrcConsts g_consts;
static bool g_ready;
static void ensure() {
    if (!g_ready) {
        crc_consts_init(g_consts);
        g_ready = true;
    }
}

static uint32_t emu_work_counter2[2];
#define emu_work_counter emu_work_counter2[0]

extern "C" {

uint64_t emu_deflate_slot_bound(uint32_t chunk) { return deflate_slot_bound(chunk); }

/* uniform partition; returns total bytes written to dst (joined stream) or <0 */
int64_t emu_deflate(const uint8_t *in, uint64_t len, uint32_t chunk_size, int level, uint32_t last_flags, uint8_t *dst, uint64_t dst_cap,
                    uint32_t grid, uint32_t *out_len_opt) {
    uint32_t nchunks = len == 0 ? 1 : (uint32_t)((len + chunk_size - 1) / chunk_size);
    uint64_t stride = deflate_slot_bound(chunk_size);
    std::vector<uint8_t> slots((size_t)(stride * nchunks + 64));
    uint8_t *sl = (uint8_t *)(((uintptr_t)slots.data() + 15) & ~(uintptr_t)15);
    std::vector<uint32_t> out_len(nchunks);
    std::vector<uint64_t> offs(nchunks + 1);
    /* input must be readable as uint4 when 16-byte aligned; copy into an aligned padded buffer */
    std::vector<uint8_t> inbuf((size_t)len + 64);
    uint8_t *ia = (uint8_t *)(((uintptr_t)inbuf.data() + 15) & ~(uintptr_t)15);
    memcpy(ia, in, (size_t)len);
    DeflateParams P;
    memset(&P, 0, sizeof(P));
    P.in = ia;
    P.total_len = len;
    P.chunk_size = chunk_size;
    P.nchunks = nchunks;
    P.last_flags = last_flags;
    P.level = level;
    P.out = sl;
    P.slot_stride = stride;
    P.out_len = out_len.data();
    static uint32_t work_counter[2];
    work_counter[0] = work_counter[1] = 0;
    P.work_counter = (grid & 0x80000000u) ? nullptr : work_counter; /* high bit of grid selects static striding */
    grid &= 0x7fffffffu;
    if (grid == 0 || grid > nchunks) grid = nchunks;
    if (deflate_stride_for_level(level) == 2) MZ_LAUNCH((deflate_chunks_kernel<2, false>), dim3(grid), dim3(DF_THREADS), DF_SMEM_BYTES, 0, P);
    else if (!deflate_lazy_for_level(level)) MZ_LAUNCH((deflate_chunks_kernel<1, false>), dim3(grid), dim3(DF_THREADS), DF_SMEM_BYTES, 0, P);
    else if (!deflate_hist_for_level(level)) MZ_LAUNCH((deflate_chunks_kernel<1, true>), dim3(grid), dim3(DF_THREADS), DF_SMEM_BYTES, 0, P);
    else MZ_LAUNCH((deflate_chunks_kernel<1, true, true>), dim3(grid), dim3(DF_THREADS), DFH_SMEM_BYTES, 0, P);
    MZ_LAUNCH(scan_lengths_kernel, dim3(1), dim3(SCAN_THREADS), 0, 0, (const uint32_t *)out_len.data(), nchunks, (uint64_t)0, offs.data());
    if (offs[nchunks] > dst_cap) return -5;
    MZ_LAUNCH(gather_slots_kernel, dim3(nchunks < 8 ? nchunks : 8), dim3(GATHER_THREADS), 0, 0, (const uint8_t *)sl, stride,
              (const uint32_t *)out_len.data(), (const uint64_t *)offs.data(), nchunks, dst);
    if (out_len_opt) memcpy(out_len_opt, out_len.data(), nchunks * 4);
    return (int64_t)offs[nchunks];
}

/* per-segment CRCs + fold; misalign shifts the start address to exercise the head path */
uint32_t emu_crc32(const uint8_t *in, uint64_t len, uint64_t seg_size, uint32_t misalign, uint32_t *seg_crcs_opt) {
    ensure();
    std::vector<uint8_t> buf((size_t)len + 96);
    uint8_t *a = (uint8_t *)(((uintptr_t)buf.data() + 15) & ~(uintptr_t)15) + (misalign & 15);
    memcpy(a, in, (size_t)len);
    uint32_t nseg = len == 0 ? 0 : (uint32_t)((len + seg_size - 1) / seg_size);
    std::vector<uint32_t> res(nseg + 1), crcs(nseg + 1);
    uint32_t out2[2] = {0, 0};
    if (nseg) {
        CrcParams P;
        memset(&P, 0, sizeof(P));
        P.in = a;
        P.total_len = len;
        P.seg_size = seg_size;
        P.nseg = nseg;
        P.consts = &g_consts;
        P.out_residue = res.data();
        P.out_crc = crcs.data();
        MZ_LAUNCH(crc32_segments_kernel, dim3(2), dim3(CRC_THREADS), CRC_SMEM_BYTES, 0, P);
    }
    MZ_LAUNCH(crc32_fold_kernel, dim3(1), dim3(CRCF_THREADS), 0, 0, (const uint32_t *)res.data(), nseg, seg_size, len, (const CrcConsts *)&g_consts, out2);
    if (seg_crcs_opt) memcpy(seg_crcs_opt, crcs.data(), nseg * 4);
    return out2[1];
}

uint32_t emu_crc32_combine(uint32_t a, uint32_t b, uint64_t len_b) {
    ensure();
    return gf2_mulmod(a, gf2_xpow(g_consts.x2n, 8ull * len_b)) ^ b;
}

/* Decode one raw stream, feeding it through bounded windows like the vtbl read path would.
 * in_window / out_window = 0 means "everything at once". Returns status; *consumed, *produced filled. */
int32_t emu_inflate(const uint8_t *in, uint64_t in_len, uint8_t *out, uint64_t out_cap, uint64_t in_window, uint64_t out_window,
                    uint64_t *consumed, uint64_t *produced, uint32_t *blocks) {
    InflateState st;
    memset(&st, 0, sizeof(st));
    std::vector<uint8_t> inbuf;
    uint64_t fed = in_window ? (in_window < in_len ? in_window : in_len) : in_len;
    uint64_t out_limit = out_window ? (out_window < out_cap ? out_window : out_cap) : out_cap;
    for (int iter = 0; iter < 1000000; iter++) {
        /* window of input starting at the byte holding the current bit position */
        uint64_t base = st.in_bitpos >> 3;
        inbuf.assign((size_t)(fed - base) + 32, 0);
        memcpy(inbuf.data(), in + base, (size_t)(fed - base));
        InflateJob job;
        memset(&job, 0, sizeof(job));
        job.in = inbuf.data();
        job.in_base = base;
        job.in_avail = fed - base;
        job.out = out;
        job.out_base = 0;
        job.out_cap = out_limit;
        job.in_final = fed == in_len;
        emu_work_counter2[0] = emu_work_counter2[1] = 0;
        MZ_LAUNCH(inflate_streams_kernel, dim3(1), dim3(INF_THREADS), INF_SMEM_BYTES, 0, (const InflateJob *)&job, &st, 1u, &emu_work_counter);
        if (st.status != INF_ST_RUN) break;
        if (st.why == INF_WHY_INPUT) {
            if (fed == in_len) { st.status = -99; break; }
            fed = fed + in_window < in_len ? fed + in_window : in_len;
        } else if (st.why == INF_WHY_OUTPUT) {
            if (out_limit == out_cap) { st.status = INF_ST_BUF_ERROR; break; }
            out_limit = out_limit + out_window < out_cap ? out_limit + out_window : out_cap;
        } else {
            st.status = -98;
            break;
        }
    }
    *consumed = (st.in_bitpos + 7) >> 3;
    *produced = st.out_pos;
    if (blocks) *blocks = st.blocks;
    return st.status;
}

/* Decode one raw stream with the segment-speculative rounds (K6), falling back to one serially decoded block
 * whenever a round makes no progress -- the same policy as the vtbl read path. stats: [0] rounds, [1] chain members,
 * [2] serial launches, [3] candidates found, [4] rounds discarded by the cross-check. */
int32_t emu_inflate_spec(const uint8_t *in, uint64_t in_len, uint8_t *out, uint64_t out_cap, uint64_t seg_bytes, uint32_t max_seg,
                         uint64_t window, uint64_t *consumed, uint64_t *produced, uint32_t *stats) {
    InflateState st;
    memset(&st, 0, sizeof(st));
    std::vector<SpecSeg> seg(max_seg);
    std::vector<InflateState> states(max_seg);
    std::vector<uint32_t> chain(2 * (size_t)max_seg);
    std::vector<uint16_t> rings((size_t)max_seg * SPEC_RING);
    std::vector<uint8_t> wins((size_t)max_seg * 32768);
    std::vector<uint16_t> gmaps((size_t)SPEC_GROUPS * 32768);
    std::vector<uint8_t> gwins((size_t)SPEC_GROUPS * 32768);
    std::vector<uint32_t> inbuf;
    SpecSummary sum;
    for (int i = 0; i < 5; i++) stats[i] = 0;
    for (int iter = 0; iter < 1000000 && st.status == INF_ST_RUN; iter++) {
        if (st.phase == INF_PH_HEADER) {
            const uint64_t base = (st.in_bitpos >> 3) & ~3ull;
            uint64_t avail = in_len - base;
            if (window && avail > window) avail = window;
            inbuf.assign((size_t)(avail + 64) / 4 + 1, 0);
            memcpy(inbuf.data(), in + base, (size_t)avail);
            SpecParams P;
            memset(&P, 0, sizeof(P));
            P.in = (const uint8_t *)inbuf.data();
            P.in_base = base;
            P.in_avail = avail;
            P.start_bit = st.in_bitpos;
            P.seg_bits = seg_bytes * 8;
            P.out = out;
            P.out_base = 0;
            P.out_pos = st.out_pos;
            P.out_end = out_cap;
            P.in_final = base + avail == in_len;
            uint64_t span = avail * 8 - (st.in_bitpos - base * 8);
            uint64_t ns = (span + P.seg_bits - 1) / P.seg_bits;
            P.nseg = (uint32_t)(ns > max_seg ? max_seg : (ns ? ns : 1));
            P.seg = seg.data();
            P.states = states.data();
            P.rings = rings.data();
            P.wins = wins.data();
            P.gmaps = gmaps.data();
            P.gwins = gwins.data();
            uint32_t work[4] = {0, 0, 0, 0};
            P.work = work;
            P.chain = chain.data();
            P.summary = &sum;
            MZ_LAUNCH(inflate_spec_find_kernel, dim3(P.nseg), dim3(INF_THREADS), SPEC_FIND_SMEM, 0, P);
            MZ_LAUNCH(inflate_spec_scan_kernel, dim3(P.nseg), dim3(INF_THREADS), INF_SMEM_BYTES, 0, P);
            MZ_LAUNCH(inflate_spec_chain_kernel, dim3(1), dim3(SPEC_CHAIN_THREADS), (size_t)P.nseg * 16, 0, P);
            MZ_LAUNCH(inflate_spec_compose_kernel, dim3(SPEC_GROUPS), dim3(SPEC_RESOLVE_THREADS), SPEC_COMPOSE_SMEM, 0, P);
            MZ_LAUNCH(inflate_spec_link_kernel, dim3(1), dim3(SPEC_RESOLVE_THREADS), 65536, 0, P);
            MZ_LAUNCH(inflate_spec_resolve_kernel, dim3(SPEC_GROUPS), dim3(SPEC_RESOLVE_THREADS), SPEC_RESOLVE_SMEM, 0, P);
            MZ_LAUNCH(inflate_spec_emit_kernel, dim3(P.nseg), dim3(INF_THREADS), INF_SMEM_BYTES, 0, P);
            stats[0]++;
            stats[3] += sum.candidates;
            if (sum.flags) stats[4]++;
            if (!sum.flags && sum.nchain > 0 && (sum.end_bit > st.in_bitpos || sum.status == INF_ST_END)) {
                stats[1] += sum.nchain;
                st.in_bitpos = sum.end_bit;
                st.out_pos += sum.total_out;
                st.blocks += sum.blocks;
                st.status = sum.status;
                continue;
            }
        }
        /* serial: to the next block boundary */
        std::vector<uint8_t> all((size_t)in_len + 64, 0);
        memcpy(all.data(), in, (size_t)in_len);
        InflateJob job;
        memset(&job, 0, sizeof(job));
        job.in = all.data();
        job.in_avail = in_len;
        job.out = out;
        job.out_cap = out_cap;
        job.in_final = 1;
        job.flags = INF_JOB_STOP_AT_BOUNDARY;
        emu_work_counter2[0] = emu_work_counter2[1] = 0;
        MZ_LAUNCH(inflate_streams_kernel, dim3(1), dim3(INF_THREADS), INF_SMEM_BYTES, 0, (const InflateJob *)&job, &st, 1u, &emu_work_counter);
        stats[2]++;
        if (st.status == INF_ST_RUN && st.why != INF_WHY_BOUNDARY) { st.status = st.why == INF_WHY_OUTPUT ? INF_ST_BUF_ERROR : -99; break; }
    }
    *consumed = (st.in_bitpos + 7) >> 3;
    *produced = st.out_pos;
    return st.status;
}


/* n messages packed at the given offsets of one buffer -> n x 32 digest bytes */
void emu_sha256(const uint8_t *in, const uint64_t *off, const uint64_t *len, uint32_t n, uint8_t *digest) {
    Sha256Params P;
    P.in = in;
    P.off = off;
    P.len = len;
    P.n = n;
    P.digest = digest;
    MZ_LAUNCH(sha256_batch_kernel, dim3(2), dim3(SHA_THREADS), 0, 0, P);
}
}
