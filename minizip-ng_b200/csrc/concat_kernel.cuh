/* concat_kernel.cuh -- K4: join per-chunk bitstreams (byte-aligned by construction) into one stream.
 *
 * Every chunk slot ends on a byte boundary (sync-flush marker or final padding, see deflate_kernel.cuh),
 * so concatenation is an exclusive scan over out_len[] followed by a gather-copy. The result is one
 * valid RFC1951 stream: what mz_stream_zlib_write + close hand to the base stream (mz_strm_zlib.c:196-201).
 */
#ifndef MZ_CONCAT_KERNEL_CUH
#define MZ_CONCAT_KERNEL_CUH

#include "mzcuda_common.cuh"

namespace mzc {

constexpr int SCAN_THREADS = 1024;

/* exclusive scan of n 32-bit lengths into 64-bit offsets; offsets[n] = total. Single CTA. */
__global__ void __launch_bounds__(SCAN_THREADS, 1) scan_lengths_kernel(const uint32_t *len, uint32_t n, uint64_t base, uint64_t *offsets) {
    __shared__ uint64_t s_part[SCAN_THREADS];
    const uint32_t tid = threadIdx.x;
    const uint32_t per = (n + SCAN_THREADS - 1) / SCAN_THREADS;
    const uint64_t lo = (uint64_t)tid * per;
    uint64_t sum = 0;
    for (uint32_t k = 0; k < per; k++)
        if (lo + k < n) sum += len[lo + k];
    s_part[tid] = sum;
    __syncthreads();
    /* Hillis-Steele inclusive scan over the 1024 partial sums */
    for (uint32_t d = 1; d < SCAN_THREADS; d <<= 1) {
        uint64_t v = tid >= d ? s_part[tid - d] : 0;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    uint64_t run = base + s_part[tid] - sum;
    for (uint32_t k = 0; k < per; k++)
        if (lo + k < n) {
            offsets[lo + k] = run;
            run += len[lo + k];
        }
    if (tid == SCAN_THREADS - 1) offsets[n] = base + s_part[tid];
}

constexpr int GATHER_THREADS = 256;

/* copy slot i (16-byte aligned source) to dst + offsets[i] (arbitrary alignment) */
__global__ void __launch_bounds__(GATHER_THREADS) gather_slots_kernel(const uint8_t *slots, uint64_t slot_stride, const uint32_t *len,
                                                                      const uint64_t *offsets, uint32_t n, uint8_t *dst) {
    for (uint32_t c = blockIdx.x; c < n; c += gridDim.x) {
        const uint8_t *src = slots + (uint64_t)c * slot_stride;
        uint8_t *d = dst + offsets[c];
        const uint32_t nbytes = len[c];
        uint32_t head = (uint32_t)((4 - ((uintptr_t)d & 3)) & 3);
        if (head > nbytes) head = nbytes;
        if (threadIdx.x < head) d[threadIdx.x] = src[threadIdx.x];
        const uint32_t nwords = (nbytes - head) >> 2;
        uint32_t *dw = (uint32_t *)(d + head);
        if ((((uintptr_t)(d + head)) & 15) == 0 && head == 0) {
            /* fully aligned: 16-byte copies */
            const uint32_t n16 = nwords >> 2;
            for (uint32_t i = threadIdx.x; i < n16; i += GATHER_THREADS) ((uint4 *)dw)[i] = ((const uint4 *)src)[i];
            for (uint32_t i = (n16 << 2) + threadIdx.x; i < nwords; i += GATHER_THREADS) dw[i] = ((const uint32_t *)src)[i];
        } else {
            const uint32_t *sw = (const uint32_t *)src; /* slot is padded: reading one word past is safe */
            const uint32_t sh = (head & 3) * 8, wo = head >> 2;
            for (uint32_t i = threadIdx.x; i < nwords; i += GATHER_THREADS) dw[i] = __funnelshift_r(sw[wo + i], sw[wo + i + 1], sh);
        }
        const uint32_t done = head + (nwords << 2);
        if (threadIdx.x < nbytes - done) d[done + threadIdx.x] = src[done + threadIdx.x];
    }
}

/* copy n small blobs (zip local headers) from a packed source to dst + dst_off[i]: one warp per blob, byte copies */
__global__ void __launch_bounds__(256) scatter_blobs_kernel(const uint8_t *blob, const uint32_t *blob_off, const uint64_t *dst_off, uint32_t n, uint8_t *dst) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = warp; i < n; i += nwarps) {
        const uint32_t a = blob_off[i], b = blob_off[i + 1];
        uint8_t *d = dst + dst_off[i];
        for (uint32_t k = a + lane; k < b; k += 32) d[k - a] = blob[k];
    }
}

/* per-chunk rows {crc32, in_len, out_len} for the multi-GPU table (one thread per chunk) */
__global__ void __launch_bounds__(256) pack_rows_kernel(const uint32_t *chunk_crc, const uint32_t *out_len, uint32_t nchunks, uint64_t total_len,
                                                        uint32_t chunk_size, uint32_t *rows) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const uint64_t off = (uint64_t)c * chunk_size;
    const uint64_t rem = total_len - off;
    rows[3 * c] = chunk_crc[c];
    rows[3 * c + 1] = rem < chunk_size ? (uint32_t)rem : chunk_size;
    rows[3 * c + 2] = out_len[c];
}

} // namespace mzc
#endif
