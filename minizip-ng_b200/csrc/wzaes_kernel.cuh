/* wzaes_kernel.cuh -- K8: the arithmetic of WinZip AES entries for a whole batch of zip entries (scope row f4, the part of it
 * that follows the codec): what mz_strm_wzaes.c does per entry with the host crypto provider,
 *   key derivation   PBKDF2-HMAC-SHA1, 1000 iterations, 2 * keylen + 2 bytes        (mz_strm_wzaes.c:95-97, mz_crypt.c:94-160)
 *   encryption       AES in counter mode, the counter a little-endian number in the first 8 nonce bytes, first block = 1
 *                    (mz_strm_wzaes.c:147-171)
 *   authentication   HMAC-SHA1 over the ciphertext, first 10 bytes stored           (mz_strm_wzaes.c:110-113, :236-256)
 * done by three kernels over per-entry tables. Every entry has its own salt, hence its own keys.
 *   wzaes_derive_kernel  one THREAD per (entry, 20-byte block of derived key): 2000 SHA-1 compressions each, no memory traffic
 *   wzaes_ctr_kernel     one CTA per (entry, 64 KiB part): round keys expanded into shared memory once per CTA, the T-table in
 *                        shared memory, one 16-byte block per thread and step; HBM: reads and writes the data once
 *   wzaes_hmac_kernel    one THREAD per entry (SHA-1 is a serial chain; like K7 the parallelism is across entries)
 * FIPS 197 / FIPS 180-4 / RFC 2104 / RFC 2898 arithmetic written from the standards; the S-box is generated (GF(2^8) inverse +
 * affine map) on the host at start-up, not pasted. */
#ifndef MZ_WZAES_KERNEL_CUH
#define MZ_WZAES_KERNEL_CUH

#include "mzcuda_common.cuh"

namespace mzc {

constexpr int WZ_KEYREC = 80;      /* per entry: encryption key (<= 32) | authentication key (<= 32) at +32 | verifier (2) at +64 */
constexpr int WZ_CTR_THREADS = 256;
constexpr int WZ_CTR_PART = 65536; /* bytes per CTA */

/* ---- SHA-1 ------------------------------------------------------------------------------------------------------- */
__device__ __forceinline__ uint32_t wz_rotl(uint32_t x, int r) { return __funnelshift_l(x, x, r); }
__device__ __forceinline__ uint32_t wz_bswap(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

/* one compression: w[16] = the block as big-endian words (destroyed), h[5] updated */
__device__ __forceinline__ void sha1_block(uint32_t (&h)[5], uint32_t (&w)[16]) {
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4];
#pragma unroll
    for (int i = 0; i < 80; i++) {
        if (i >= 16) w[i & 15] = wz_rotl(w[(i + 13) & 15] ^ w[(i + 8) & 15] ^ w[(i + 2) & 15] ^ w[i & 15], 1);
        uint32_t f, k;
        if (i < 20) { f = (b & c) | (~b & d); k = 0x5a827999u; }
        else if (i < 40) { f = b ^ c ^ d; k = 0x6ed9eba1u; }
        else if (i < 60) { f = (b & c) | (b & d) | (c & d); k = 0x8f1bbcdcu; }
        else { f = b ^ c ^ d; k = 0xca62c1d6u; }
        const uint32_t t = wz_rotl(a, 5) + f + e + k + w[i & 15];
        e = d; d = c; c = wz_rotl(b, 30); b = a; a = t;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e;
}
__device__ __forceinline__ void sha1_init(uint32_t (&h)[5]) {
    h[0] = 0x67452301u; h[1] = 0xefcdab89u; h[2] = 0x98badcfeu; h[3] = 0x10325476u; h[4] = 0xc3d2e1f0u;
}

/* HMAC key (<= 64 bytes after hashing a longer one) -> the two chaining values every HMAC with this key starts from */
__device__ inline void hmac_sha1_states(const uint8_t *key, uint32_t klen, uint32_t (&hi)[5], uint32_t (&ho)[5]) {
    uint32_t k[16], w[16];
#pragma unroll
    for (int q = 0; q < 16; q++) k[q] = 0;
    if (klen > 64) { /* key = SHA1(key) */
        uint32_t h[5];
        sha1_init(h);
        uint32_t done = 0;
        for (; done + 64 <= klen; done += 64) {
#pragma unroll
            for (int q = 0; q < 16; q++)
                w[q] = ((uint32_t)key[done + 4 * q] << 24) | ((uint32_t)key[done + 4 * q + 1] << 16) | ((uint32_t)key[done + 4 * q + 2] << 8) | key[done + 4 * q + 3];
            sha1_block(h, w);
        }
        const uint32_t r = klen - done;
#pragma unroll
        for (int q = 0; q < 16; q++) w[q] = 0;
        for (uint32_t i = 0; i < r; i++) w[i >> 2] |= (uint32_t)key[done + i] << (24 - 8 * (i & 3));
        w[r >> 2] |= 0x80u << (24 - 8 * (r & 3));
        if (r >= 56) {
            sha1_block(h, w);
#pragma unroll
            for (int q = 0; q < 16; q++) w[q] = 0;
        }
        w[15] = klen * 8;
        sha1_block(h, w);
#pragma unroll
        for (int q = 0; q < 5; q++) k[q] = h[q];
    } else {
        for (uint32_t i = 0; i < klen; i++) k[i >> 2] |= (uint32_t)key[i] << (24 - 8 * (i & 3));
    }
    sha1_init(hi);
#pragma unroll
    for (int q = 0; q < 16; q++) w[q] = k[q] ^ 0x36363636u;
    sha1_block(hi, w);
    sha1_init(ho);
#pragma unroll
    for (int q = 0; q < 16; q++) w[q] = k[q] ^ 0x5c5c5c5cu;
    sha1_block(ho, w);
}

/* outer hash of an HMAC whose inner hash is `in`: SHA1(opad block || in) */
__device__ __forceinline__ void hmac_sha1_outer(const uint32_t (&ho)[5], const uint32_t (&in)[5], uint32_t (&out)[5]) {
    uint32_t w[16];
#pragma unroll
    for (int q = 0; q < 5; q++) w[q] = in[q];
    w[5] = 0x80000000u;
#pragma unroll
    for (int q = 6; q < 15; q++) w[q] = 0;
    w[15] = (64 + 20) * 8;
#pragma unroll
    for (int q = 0; q < 5; q++) out[q] = ho[q];
    sha1_block(out, w);
}

struct WzDeriveParams {
    const uint8_t *password; /* device copy */
    uint32_t pw_len;         /* <= 128 (MZ_AES_PW_LENGTH_MAX) */
    const uint8_t *salts;    /* n x 16 bytes, the first salt_len of each used */
    uint32_t salt_len;       /* 8 / 12 / 16 */
    uint32_t key_len;        /* 16 / 24 / 32 */
    uint32_t iterations;     /* 1000 (MZ_AES_KEYING_ITERATIONS) */
    uint32_t n;
    uint8_t *keys;           /* n x WZ_KEYREC */
};

__global__ void __launch_bounds__(128) wzaes_derive_kernel(WzDeriveParams P) {
    const uint32_t nblk = (2 * P.key_len + 2 + 19) / 20; /* 20-byte blocks of derived key: 2 / 3 / 4 */
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < P.n * nblk; t += gridDim.x * blockDim.x) {
        const uint32_t e = t / nblk, bi = t - e * nblk;
        uint32_t hi[5], ho[5];
        hmac_sha1_states(P.password, P.pw_len, hi, ho);
        /* U1 = HMAC(P, salt || INT_32_BE(bi + 1)) */
        uint32_t w[16], in[5], u[5], acc[5];
#pragma unroll
        for (int q = 0; q < 16; q++) w[q] = 0;
        const uint8_t *salt = P.salts + (uint64_t)e * 16;
        for (uint32_t i = 0; i < P.salt_len; i++) w[i >> 2] |= (uint32_t)salt[i] << (24 - 8 * (i & 3));
        const uint32_t m = P.salt_len + 4; /* the counter's bytes: 0 0 0 (bi + 1) */
        w[(m - 1) >> 2] |= (bi + 1) << (24 - 8 * ((m - 1) & 3));
        w[m >> 2] |= 0x80u << (24 - 8 * (m & 3));
        w[15] = (64 + m) * 8;
#pragma unroll
        for (int q = 0; q < 5; q++) in[q] = hi[q];
        sha1_block(in, w);
        hmac_sha1_outer(ho, in, u);
#pragma unroll
        for (int q = 0; q < 5; q++) acc[q] = u[q];
        for (uint32_t it = 1; it < P.iterations; it++) { /* Uj = HMAC(P, Uj-1): two compressions */
#pragma unroll
            for (int q = 0; q < 5; q++) w[q] = u[q];
            w[5] = 0x80000000u;
#pragma unroll
            for (int q = 6; q < 15; q++) w[q] = 0;
            w[15] = (64 + 20) * 8;
#pragma unroll
            for (int q = 0; q < 5; q++) in[q] = hi[q];
            sha1_block(in, w);
            hmac_sha1_outer(ho, in, u);
#pragma unroll
            for (int q = 0; q < 5; q++) acc[q] ^= u[q];
        }
        /* derived bytes [20 bi, 20 bi + 20) -> encryption key | authentication key | verifier */
        uint8_t *rec = P.keys + (uint64_t)e * WZ_KEYREC;
        for (uint32_t i = 0; i < 20; i++) {
            const uint32_t pos = bi * 20 + i;
            const uint8_t v = (uint8_t)(acc[i >> 2] >> (24 - 8 * (i & 3)));
            if (pos < P.key_len) rec[pos] = v;
            else if (pos < 2 * P.key_len) rec[32 + pos - P.key_len] = v;
            else if (pos < 2 * P.key_len + 2) rec[64 + pos - 2 * P.key_len] = v;
        }
    }
}

/* ---- AES in counter mode ------------------------------------------------------------------------------------------ */
struct WzCtrParams {
    uint8_t *data;            /* base; entry e's stream occupies [off[e], off[e] + len[e]) and is XOR-ed in place */
    const uint64_t *off, *len;
    uint32_t n;
    uint32_t parts;           /* grid.y: 64 KiB parts of the longest entry */
    const uint8_t *keys;      /* n x WZ_KEYREC */
    uint32_t key_len;
    const uint32_t *te0;      /* 256 words: (2 S[x], S[x], S[x], 3 S[x]) big-endian */
    const uint8_t *sbox;      /* 256 bytes */
};

__device__ __forceinline__ uint32_t wz_ror(uint32_t x, int r) { return __funnelshift_r(x, x, r); }

__global__ void __launch_bounds__(WZ_CTR_THREADS) wzaes_ctr_kernel(WzCtrParams P) {
    __shared__ uint32_t s_te[256];
    __shared__ uint32_t s_sb[256];
    __shared__ uint32_t s_rk[60];
    const uint32_t tid = threadIdx.x;
    s_te[tid] = P.te0[tid]; /* (WZ_CTR_THREADS == 256) */
    s_sb[tid] = P.sbox[tid];
    for (uint32_t e = blockIdx.x; e < P.n; e += gridDim.x) {
        const uint64_t len = P.len[e];
        const uint64_t part0 = (uint64_t)blockIdx.y * WZ_CTR_PART;
        if (part0 >= len) continue; /* (uniform) */
        __syncthreads(); /* tables loaded; the previous entry's round keys no longer in use */
        if (tid == 0) { /* key expansion (FIPS 197 5.2), big-endian words */
            const uint8_t *k = P.keys + (uint64_t)e * WZ_KEYREC;
            const uint32_t nk = P.key_len / 4, nr = nk + 6;
            for (uint32_t i = 0; i < nk; i++) s_rk[i] = ((uint32_t)k[4 * i] << 24) | ((uint32_t)k[4 * i + 1] << 16) | ((uint32_t)k[4 * i + 2] << 8) | k[4 * i + 3];
            uint32_t rcon = 1;
            for (uint32_t i = nk; i < 4 * (nr + 1); i++) {
                uint32_t t = s_rk[i - 1];
                if (i % nk == 0) {
                    t = (t << 8) | (t >> 24);
                    t = (s_sb[t >> 24] << 24) | (s_sb[(t >> 16) & 255] << 16) | (s_sb[(t >> 8) & 255] << 8) | s_sb[t & 255];
                    t ^= rcon << 24;
                    rcon = (rcon << 1) ^ ((rcon & 0x80) ? 0x11b : 0);
                } else if (nk > 6 && i % nk == 4) {
                    t = (s_sb[t >> 24] << 24) | (s_sb[(t >> 16) & 255] << 16) | (s_sb[(t >> 8) & 255] << 8) | s_sb[t & 255];
                }
                s_rk[i] = s_rk[i - nk] ^ t;
            }
        }
        __syncthreads();
        const uint32_t nr = P.key_len / 4 + 6;
        uint8_t *base = P.data + P.off[e];
        const uint64_t part_end = part0 + WZ_CTR_PART < len ? part0 + WZ_CTR_PART : len;
        for (uint64_t o = part0 + (uint64_t)tid * 16; o < part_end; o += (uint64_t)WZ_CTR_THREADS * 16) {
            const uint64_t ctr = (o >> 4) + 1; /* little-endian in nonce bytes 0..7 (mz_strm_wzaes.c:158-160); bytes 8..15 stay zero */
            uint32_t s0 = wz_bswap((uint32_t)ctr) ^ s_rk[0], s1 = wz_bswap((uint32_t)(ctr >> 32)) ^ s_rk[1], s2 = s_rk[2], s3 = s_rk[3];
            for (uint32_t r = 1; r < nr; r++) {
                const uint32_t t0 = s_te[s0 >> 24] ^ wz_ror(s_te[(s1 >> 16) & 255], 8) ^ wz_ror(s_te[(s2 >> 8) & 255], 16) ^ wz_ror(s_te[s3 & 255], 24) ^ s_rk[4 * r];
                const uint32_t t1 = s_te[s1 >> 24] ^ wz_ror(s_te[(s2 >> 16) & 255], 8) ^ wz_ror(s_te[(s3 >> 8) & 255], 16) ^ wz_ror(s_te[s0 & 255], 24) ^ s_rk[4 * r + 1];
                const uint32_t t2 = s_te[s2 >> 24] ^ wz_ror(s_te[(s3 >> 16) & 255], 8) ^ wz_ror(s_te[(s0 >> 8) & 255], 16) ^ wz_ror(s_te[s1 & 255], 24) ^ s_rk[4 * r + 2];
                const uint32_t t3 = s_te[s3 >> 24] ^ wz_ror(s_te[(s0 >> 16) & 255], 8) ^ wz_ror(s_te[(s1 >> 8) & 255], 16) ^ wz_ror(s_te[s2 & 255], 24) ^ s_rk[4 * r + 3];
                s0 = t0; s1 = t1; s2 = t2; s3 = t3;
            }
            uint32_t ks[4];
            ks[0] = ((s_sb[s0 >> 24] << 24) | (s_sb[(s1 >> 16) & 255] << 16) | (s_sb[(s2 >> 8) & 255] << 8) | s_sb[s3 & 255]) ^ s_rk[4 * nr];
            ks[1] = ((s_sb[s1 >> 24] << 24) | (s_sb[(s2 >> 16) & 255] << 16) | (s_sb[(s3 >> 8) & 255] << 8) | s_sb[s0 & 255]) ^ s_rk[4 * nr + 1];
            ks[2] = ((s_sb[s2 >> 24] << 24) | (s_sb[(s3 >> 16) & 255] << 16) | (s_sb[(s0 >> 8) & 255] << 8) | s_sb[s1 & 255]) ^ s_rk[4 * nr + 2];
            ks[3] = ((s_sb[s3 >> 24] << 24) | (s_sb[(s0 >> 16) & 255] << 16) | (s_sb[(s1 >> 8) & 255] << 8) | s_sb[s2 & 255]) ^ s_rk[4 * nr + 3];
            uint8_t *p = base + o;
            const uint32_t nb = part_end - o >= 16 ? 16u : (uint32_t)(part_end - o);
            if (nb == 16 && (((uintptr_t)p) & 3) == 0) {
                uint32_t *q = (uint32_t *)p;
#pragma unroll
                for (int j = 0; j < 4; j++) q[j] ^= wz_bswap(ks[j]);
            } else {
                for (uint32_t j = 0; j < nb; j++) p[j] ^= (uint8_t)(ks[j >> 2] >> (24 - 8 * (j & 3)));
            }
        }
    }
}

/* ---- HMAC-SHA1 of every entry's ciphertext ------------------------------------------------------------------------ */
struct WzHmacParams {
    const uint8_t *data;
    const uint64_t *off, *len;
    uint32_t n;
    const uint8_t *keys;   /* n x WZ_KEYREC; the authentication key at +32 */
    uint32_t key_len;
    uint8_t *mac;          /* n x 20 bytes */
};

__global__ void __launch_bounds__(128) wzaes_hmac_kernel(WzHmacParams P) {
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < P.n; e += gridDim.x * blockDim.x) {
        uint32_t hi[5], ho[5], w[16];
        hmac_sha1_states(P.keys + (uint64_t)e * WZ_KEYREC + 32, P.key_len, hi, ho);
        const uint8_t *p = P.data + P.off[e];
        const uint64_t len = P.len[e];
        const bool al = (((uintptr_t)p) & 3) == 0;
        for (uint64_t blk = 0; blk < (len >> 6); blk++) {
            if (al) {
#pragma unroll
                for (int q = 0; q < 16; q++) w[q] = wz_bswap(((const uint32_t *)p)[q]);
            } else {
#pragma unroll
                for (int q = 0; q < 16; q++) w[q] = ((uint32_t)p[4 * q] << 24) | ((uint32_t)p[4 * q + 1] << 16) | ((uint32_t)p[4 * q + 2] << 8) | p[4 * q + 3];
            }
            sha1_block(hi, w);
            p += 64;
        }
        const uint32_t r = (uint32_t)(len & 63);
#pragma unroll
        for (int q = 0; q < 16; q++) w[q] = 0;
        for (uint32_t i = 0; i < r; i++) w[i >> 2] |= (uint32_t)p[i] << (24 - 8 * (i & 3));
        w[r >> 2] |= 0x80u << (24 - 8 * (r & 3));
        if (r >= 56) {
            sha1_block(hi, w);
#pragma unroll
            for (int q = 0; q < 16; q++) w[q] = 0;
        }
        const uint64_t bits = (len + 64) << 3;
        w[14] = (uint32_t)(bits >> 32);
        w[15] = (uint32_t)bits;
        sha1_block(hi, w);
        uint32_t out[5];
        hmac_sha1_outer(ho, hi, out);
        uint8_t *m = P.mac + (uint64_t)e * 20;
        for (int i = 0; i < 20; i++) m[i] = (uint8_t)(out[i >> 2] >> (24 - 8 * (i & 3)));
    }
}

} // namespace mzc
#endif
