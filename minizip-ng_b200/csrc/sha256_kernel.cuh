/* sha256_kernel.cuh -- K7: SHA-256 of many independent buffers (scope row f3: the per-entry hash the reference's zip
 * writer stores in the MZ_ZIP_EXTENSION_HASH extra field, mz_zip_rw.c:1339-1354 (begin), :1441-1444 (update), :1365-1420
 * (digest -> extra field), and its reader verifies, :410-450).
 *
 * SHA-256 is a serial chain over 64-byte blocks, so the parallelism is across messages: one THREAD per message, all 64
 * rounds unrolled with the 16-word message schedule kept in registers; the 32 lanes of a warp walk 32 different
 * messages in lockstep (zip entries of one archive are mostly of similar size). Each lane reads its message with 16-byte
 * loads (messages start 16-byte aligned in the batch buffers); FIPS 180-4 arithmetic, written from the standard. */
#ifndef MZ_SHA256_KERNEL_CUH
#define MZ_SHA256_KERNEL_CUH

#include "mzcuda_common.cuh"

namespace mzc {

constexpr int SHA_THREADS = 128;

struct Sha256Params {
    const uint8_t *in;       /* base of the batch buffer */
    const uint64_t *off;     /* per message: byte offset into `in` (any alignment; 16-byte aligned offsets take the fast path) */
    const uint64_t *len;     /* per message: bytes */
    uint32_t n;
    uint8_t *digest;         /* n x 32 bytes, big-endian words as the standard prints them */
};

__device__ __forceinline__ uint32_t sha_rotr(uint32_t x, int r) { return __funnelshift_r(x, x, r); }
__device__ __forceinline__ uint32_t sha_bswap(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

__device__ __forceinline__ uint32_t sha_k(int i) {
    /* first 32 bits of the fractional parts of the cube roots of the first 64 primes */
    constexpr uint32_t K[64] = {
        0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, 0xd807aa98u, 0x12835b01u, 0x243185beu,
        0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u, 0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau,
        0x5cb0a9dcu, 0x76f988dau, 0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u, 0x27b70a85u,
        0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u, 0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u,
        0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u, 0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu,
        0x682e6ff3u, 0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
    return K[i];
}

/* one compression: w[16] = the block as big-endian words (destroyed), h[8] updated */
__device__ __forceinline__ void sha256_block(uint32_t (&h)[8], uint32_t (&w)[16]) {
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        if (i >= 16) {
            const uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            const uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
            const uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
            w[i & 15] += s0 + w[(i + 9) & 15] + s1;
        }
        const uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
        const uint32_t ch = (e & f) ^ (~e & g);
        const uint32_t t1 = hh + S1 + ch + sha_k(i) + w[i & 15];
        const uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
        const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        const uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

__global__ void __launch_bounds__(SHA_THREADS) sha256_batch_kernel(Sha256Params P) {
    for (uint32_t m = blockIdx.x * blockDim.x + threadIdx.x; m < P.n; m += gridDim.x * blockDim.x) {
        const uint8_t *p = P.in + P.off[m];
        const uint64_t len = P.len[m];
        uint32_t h[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
        uint32_t w[16];
        const bool al = (((uintptr_t)p) & 15) == 0;
        const uint64_t nfull = len >> 6;
        for (uint64_t blk = 0; blk < nfull; blk++) {
            if (al) {
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint4 v = ((const uint4 *)p)[q];
                    w[4 * q] = sha_bswap(v.x); w[4 * q + 1] = sha_bswap(v.y); w[4 * q + 2] = sha_bswap(v.z); w[4 * q + 3] = sha_bswap(v.w);
                }
            } else {
#pragma unroll
                for (int q = 0; q < 16; q++) w[q] = ((uint32_t)p[4 * q] << 24) | ((uint32_t)p[4 * q + 1] << 16) | ((uint32_t)p[4 * q + 2] << 8) | p[4 * q + 3];
            }
            sha256_block(h, w);
            p += 64;
        }
        /* padding: 0x80, zeros, the bit length as a 64-bit big-endian number; one or two more blocks */
        const uint32_t r = (uint32_t)(len & 63);
#pragma unroll
        for (int q = 0; q < 16; q++) w[q] = 0;
        for (uint32_t i = 0; i < r; i++) w[i >> 2] |= (uint32_t)p[i] << (24 - 8 * (i & 3));
        w[r >> 2] |= 0x80u << (24 - 8 * (r & 3));
        if (r >= 56) {
            sha256_block(h, w);
#pragma unroll
            for (int q = 0; q < 16; q++) w[q] = 0;
        }
        const uint64_t bits = len << 3;
        w[14] = (uint32_t)(bits >> 32);
        w[15] = (uint32_t)bits;
        sha256_block(h, w);
        uint32_t *out = (uint32_t *)(P.digest + (uint64_t)m * 32);
#pragma unroll
        for (int q = 0; q < 8; q++) out[q] = sha_bswap(h[q]);
    }
}

} // namespace mzc
#endif
