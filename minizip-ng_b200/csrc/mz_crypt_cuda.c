/* mz_crypt_cuda.c -- replacement for mz_crypt_crc32_update (mz_crypt.c:35-92).
 *
 * Same contract: `value` is the running CRC-32 (starts at 0, chains), inversion happens inside
 * (mz_crypt.c:81,90); size 0 returns value unchanged (mz_os.c:340 relies on it). Like the reference
 * function it is re-entrant: any number of threads may call it at once with different buffers.
 *
 * Large buffers go to the K1 kernel (include/mz_cuda_batch.h) in 8 MiB pieces: the copy of piece k+1 into
 * pinned staging overlaps the upload and the kernels of piece k, the pieces' CRCs are folded on the host
 * (mz_cuda_crc32_combine). The staging (two pinned + two device buffers, a stream, two events, per DEVICE)
 * is used by one caller at a time (mutex held over the whole call).
 *
 * Tiny buffers do not go to the GPU: the reference calls this once per byte from the PKWARE key schedule
 * (mz_strm_pkcrypt.c:79,86) and once per <=64 KiB from the zip entry loop (mz_zip.c:2049,2064); a PCIe
 * round trip per call would be absurd, so calls below MZ_CUDA_CRC_MIN_BYTES (default 1 MiB) are answered
 * on the host: by carry-less-multiplication folding where the CPU has PCLMULQDQ (64 bytes and more: ~17 GB/s per
 * core on 64 KiB calls, zlib 1.3's crc32: 3.8), else and for the tails by a slice-by-16 table loop (the table-driven
 * form of mz_crypt.c:81-90). This is the only host arithmetic in the product and it is stated here, in DESIGN.md
 * and in include/mz_strm_cuda.h.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mz_abi.h"
#include "mz_cuda_batch.h"
#include "mz_strm_cuda.h"

/* ---- host path: slice-by-16 ------------------------------------------------------------------------------ */
static uint32_t g_tab[16][256];
static pthread_once_t g_tab_once = PTHREAD_ONCE_INIT;

static void tab_init(void) {
    for (uint32_t n = 0; n < 256; n++) {
        uint32_t c = n;
        for (int k = 0; k < 8; k++)
            c = (c & 1) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
        g_tab[0][n] = c;
    }
    for (uint32_t n = 0; n < 256; n++)
        for (int k = 1; k < 16; k++)
            g_tab[k][n] = (g_tab[k - 1][n] >> 8) ^ g_tab[0][g_tab[k - 1][n] & 0xFF];
}

static inline uint32_t ld32le(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

/* ---- host path on x86-64 with carry-less multiplication: fold 64 bytes per step --------------------------------
 * (V. Gopal et al., "Fast CRC Computation for Generic Polynomials Using PCLMULQDQ Instruction", Intel 2009, the bit-reflected
 * variant.) The message is a polynomial over GF(2); 128-bit pieces A, B of it that lie T bits apart satisfy
 * A x^T + B = (A_hi x^(T+64) mod P) + (A_lo x^T mod P) + B (mod P), two PCLMULQDQ per 16 bytes instead of 16 table lookups.
 * The constants are k(n) = bitreflect32(x^n mod P) << 1 for P = 0x104C11DB7 (the shift makes up for the reflected product being
 * one bit short): fold across 512 bits k(4*128+32), k(4*128-32); across 128 bits k(128+32), k(128-32); 128 -> 64 bits k(64);
 * Barrett reduction with mu' = bitreflect33(floor(x^64 / P)) and P' = bitreflect33(P). tests/test_abi_cpu.py re-derives them.
 * Used for calls of 64 bytes and more when the CPU has PCLMULQDQ + SSE4.1 (zlib 1.3's braided loop: ~3.8 GB/s per core here;
 * the table loop below: ~3.0; this: the memory stream). */
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define MZ_HAVE_CLMUL_PATH 1
static int g_clmul_ok;
static pthread_once_t g_cpu_once = PTHREAD_ONCE_INIT;
static void cpu_init(void) {
    __builtin_cpu_init();
    const char *off = getenv("MZ_CUDA_CRC_NO_CLMUL"); /* tests: keep the table loop covered on machines that have the instruction */
    g_clmul_ok = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1") && !(off && off[0] == '1');
}

#define CL_FOLD(x, k, d)                                             \
    do {                                                             \
        const __m128i lo_ = _mm_clmulepi64_si128((x), (k), 0x00);    \
        (x) = _mm_clmulepi64_si128((x), (k), 0x11);                  \
        (x) = _mm_xor_si128(_mm_xor_si128((x), lo_), (d));           \
    } while (0)

/* state in, state out (the inverted running value); len >= 64 and a multiple of 16 */
__attribute__((target("pclmul,sse4.1"))) static uint32_t crc_clmul(uint32_t c, const uint8_t *buf, size_t len) {
    const __m128i k512 = _mm_set_epi64x(0x01c6e41596ll, 0x0154442bd4ll); /* {k(4*128+32), k(4*128-32)} */
    const __m128i k128 = _mm_set_epi64x(0x00ccaa009ell, 0x01751997d0ll); /* {k(128+32), k(128-32)} */
    const __m128i k64 = _mm_set_epi64x(0, 0x0163cd6124ll);               /* k(64) */
    const __m128i pmu = _mm_set_epi64x(0x01f7011641ll, 0x01db710641ll);  /* {P', mu'} */
    const __m128i mask32 = _mm_setr_epi32(~0, 0, ~0, 0);
    __m128i a = _mm_loadu_si128((const __m128i *)(buf + 0)), b = _mm_loadu_si128((const __m128i *)(buf + 16));
    __m128i d = _mm_loadu_si128((const __m128i *)(buf + 32)), e = _mm_loadu_si128((const __m128i *)(buf + 48));
    a = _mm_xor_si128(a, _mm_cvtsi32_si128((int)c));
    buf += 64;
    len -= 64;
    while (len >= 64) { /* four independent chains, 512 bits apart */
        CL_FOLD(a, k512, _mm_loadu_si128((const __m128i *)(buf + 0)));
        CL_FOLD(b, k512, _mm_loadu_si128((const __m128i *)(buf + 16)));
        CL_FOLD(d, k512, _mm_loadu_si128((const __m128i *)(buf + 32)));
        CL_FOLD(e, k512, _mm_loadu_si128((const __m128i *)(buf + 48)));
        buf += 64;
        len -= 64;
    }
    CL_FOLD(a, k128, b); /* the four chains into one, 128 bits apart each */
    CL_FOLD(a, k128, d);
    CL_FOLD(a, k128, e);
    while (len >= 16) {
        CL_FOLD(a, k128, _mm_loadu_si128((const __m128i *)buf));
        buf += 16;
        len -= 16;
    }
    /* 128 -> 64 bits: the low half times x^(64+32), then the low 32 bits of what is left times x^64 */
    __m128i t = _mm_clmulepi64_si128(a, k128, 0x10);
    a = _mm_xor_si128(_mm_srli_si128(a, 8), t);
    t = _mm_srli_si128(a, 4);
    a = _mm_xor_si128(_mm_clmulepi64_si128(_mm_and_si128(a, mask32), k64, 0x00), t);
    /* Barrett: 64 -> 32 bits */
    t = _mm_and_si128(_mm_clmulepi64_si128(_mm_and_si128(a, mask32), pmu, 0x10), mask32);
    a = _mm_xor_si128(a, _mm_clmulepi64_si128(t, pmu, 0x00));
    return (uint32_t)_mm_extract_epi32(a, 1);
}
#endif

static uint32_t crc_small(uint32_t value, const uint8_t *buf, size_t size) {
    pthread_once(&g_tab_once, tab_init);
    uint32_t c = ~value;
#ifdef MZ_HAVE_CLMUL_PATH
    pthread_once(&g_cpu_once, cpu_init);
    if (g_clmul_ok && size >= 64) {
        const size_t n = size & ~(size_t)15;
        c = crc_clmul(c, buf, n);
        buf += n;
        size -= n;
    }
#endif
    while (size >= 16) {
        const uint32_t a = ld32le(buf) ^ c, b = ld32le(buf + 4), d = ld32le(buf + 8), e = ld32le(buf + 12);
        c = g_tab[15][a & 0xFF] ^ g_tab[14][(a >> 8) & 0xFF] ^ g_tab[13][(a >> 16) & 0xFF] ^ g_tab[12][a >> 24] ^
            g_tab[11][b & 0xFF] ^ g_tab[10][(b >> 8) & 0xFF] ^ g_tab[9][(b >> 16) & 0xFF] ^ g_tab[8][b >> 24] ^
            g_tab[7][d & 0xFF] ^ g_tab[6][(d >> 8) & 0xFF] ^ g_tab[5][(d >> 16) & 0xFF] ^ g_tab[4][d >> 24] ^
            g_tab[3][e & 0xFF] ^ g_tab[2][(e >> 8) & 0xFF] ^ g_tab[1][(e >> 16) & 0xFF] ^ g_tab[0][e >> 24];
        buf += 16;
        size -= 16;
    }
    while (size > 0) {
        c = (c >> 8) ^ g_tab[0][(c ^ *buf) & 0xFF];
        buf += 1;
        size -= 1;
    }
    return ~c;
}

/* ---- device path: per-device staging, one caller at a time ---------------------------------------------------- */
#define CRC_PIECE ((size_t)8 << 20)
#define CRC_SEG ((uint64_t)16384)
#define CRC_MAX_DEV 16
#define CRC_MAX_PIECES 256 /* 2 GiB / 8 MiB */

typedef struct crc_dev {
    pthread_mutex_t mu;
    int ready;
    void *stream;
    void *ev[2];
    uint8_t *dev[2];
    uint8_t *pin[2];
    uint32_t *d_res[2]; /* per slot: residues of the piece's segments + 2 words of fold output */
    uint32_t *h_out;    /* pinned: 2 words per piece */
} crc_dev;

static crc_dev g_devs[CRC_MAX_DEV];
static pthread_once_t g_dev_once = PTHREAD_ONCE_INIT;
static int64_t g_min_bytes = -1;

static void dev_init_all(void) {
    for (int i = 0; i < CRC_MAX_DEV; i++)
        pthread_mutex_init(&g_devs[i].mu, NULL);
    const char *v = getenv("MZ_CUDA_CRC_MIN_BYTES");
    int64_t m = (v && *v) ? atoll(v) : (1 << 20);
    g_min_bytes = m < 0 ? (1 << 20) : m;
}

static void die(const char *what) {
    fprintf(stderr, "mz_crypt_crc32_update: %s: %s\n", what, mz_cuda_last_error());
    abort(); /* there is no error return in this contract, and a wrong CRC would be worse than a loud stop */
}

static void dev_prepare(crc_dev *d) { /* called with d->mu held */
    if (d->ready)
        return;
    const size_t nres = (size_t)(CRC_PIECE / CRC_SEG) + 8;
    d->stream = mz_cuda_stream_create();
    for (int s = 0; s < 2; s++) {
        d->ev[s] = mz_cuda_event_create();
        d->dev[s] = (uint8_t *)mz_cuda_malloc(CRC_PIECE + 64);
        d->pin[s] = (uint8_t *)mz_cuda_host_alloc(CRC_PIECE);
        d->d_res[s] = (uint32_t *)mz_cuda_malloc(nres * 4);
    }
    d->h_out = (uint32_t *)mz_cuda_host_alloc(CRC_MAX_PIECES * 8);
    if (!d->stream || !d->ev[0] || !d->ev[1] || !d->dev[0] || !d->dev[1] || !d->pin[0] || !d->pin[1] || !d->d_res[0] || !d->d_res[1] || !d->h_out)
        die("cannot allocate the staging buffers");
    d->ready = 1;
}

uint32_t mz_crypt_crc32_update(uint32_t value, const uint8_t *buf, int32_t size) {
    if (size <= 0 || !buf)
        return value;
    pthread_once(&g_dev_once, dev_init_all);
    if ((int64_t)size < g_min_bytes)
        return crc_small(value, buf, (size_t)size);
    if (mz_cuda_init() != MZ_OK) {
        fprintf(stderr, "mz_crypt_crc32_update: no usable sm_100 GPU (%s); refusing to fall back on the CPU for %d bytes\n",
                mz_cuda_last_error(), size);
        abort();
    }
    const int ord = mz_cuda_get_device();
    if (ord < 0 || ord >= CRC_MAX_DEV)
        die("unexpected device ordinal");
    crc_dev *d = &g_devs[ord];
    pthread_mutex_lock(&d->mu);
    dev_prepare(d);
    const int pinned = mz_cuda_host_is_pinned(buf);
    const size_t total = (size_t)size;
    const uint32_t npieces = (uint32_t)((total + CRC_PIECE - 1) / CRC_PIECE);
    for (uint32_t i = 0; i < npieces; i++) {
        const int s = (int)(i & 1);
        const size_t off = (size_t)i * CRC_PIECE;
        const size_t n = total - off < CRC_PIECE ? total - off : CRC_PIECE;
        const uint32_t nseg = (uint32_t)((n + CRC_SEG - 1) / CRC_SEG);
        const uint8_t *src = buf + off;
        if (i >= 2 && mz_cuda_event_sync(d->ev[s]) != MZ_OK) /* the slot's previous piece has left the staging */
            die("event sync");
        if (!pinned) {
            memcpy(d->pin[s], src, n);
            src = d->pin[s];
        }
        if (mz_cuda_memcpy_h2d(d->dev[s], src, n, d->stream) != MZ_OK ||
            mz_cuda_crc32_segments(d->dev[s], n, CRC_SEG, NULL, NULL, nseg, d->d_res[s], NULL, d->stream) != MZ_OK ||
            mz_cuda_crc32_fold(d->d_res[s], nseg, CRC_SEG, n, d->d_res[s] + nseg, d->stream) != MZ_OK ||
            mz_cuda_memcpy_d2h(d->h_out + 2 * i, d->d_res[s] + nseg, 8, d->stream) != MZ_OK ||
            mz_cuda_event_record(d->ev[s], d->stream) != MZ_OK)
            die("CUDA failure");
    }
    if (mz_cuda_stream_sync(d->stream) != MZ_OK)
        die("stream sync");
    for (uint32_t i = 0; i < npieces; i++) {
        const size_t off = (size_t)i * CRC_PIECE;
        const size_t n = total - off < CRC_PIECE ? total - off : CRC_PIECE;
        value = mz_cuda_crc32_combine(value, d->h_out[2 * i + 1], n); /* crc(v, A) (+) crc(0, B) -> crc(v, A || B) */
    }
    pthread_mutex_unlock(&d->mu);
    return value;
}
