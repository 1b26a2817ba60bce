/* mz_crypt_cuda.c -- replacement for mz_crypt_crc32_update (mz_crypt.c:35-92).
 *
 * Same contract: `value` is the running CRC-32 (starts at 0, chains), inversion happens inside
 * (mz_crypt.c:81,90); size 0 returns value unchanged (mz_os.c:340 relies on it).
 *
 * Large buffers go to the K1 kernel (include/mz_cuda_batch.h). Tiny ones do not: the reference calls
 * this once per byte from the PKWARE key schedule (mz_strm_pkcrypt.c:79,86) and once per <=64 KiB from
 * the zip entry loop (mz_zip.c:2049,2064); a PCIe round trip per call would be absurd, so calls below
 * MZ_CUDA_CRC_MIN_BYTES (default 1 MiB) are answered by the same ten-line table loop the reference
 * keeps for builds without zlib (mz_crypt.c:81-90). That loop is the only host arithmetic in the
 * product and it is stated here, in DESIGN.md and in include/mz_strm_cuda.h.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mz_abi.h"
#include "mz_cuda_batch.h"
#include "mz_strm_cuda.h"

static uint32_t g_tab[256];
static int g_tab_ready;

static void tab_init(void) {
    for (uint32_t n = 0; n < 256; n++) {
        uint32_t c = n;
        for (int k = 0; k < 8; k++)
            c = (c & 1) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
        g_tab[n] = c;
    }
    g_tab_ready = 1;
}

static uint32_t crc_small(uint32_t value, const uint8_t *buf, int32_t size) {
    if (!g_tab_ready)
        tab_init();
    value = ~value;
    while (size > 0) {
        value = (value >> 8) ^ g_tab[(value ^ *buf) & 0xFF];
        buf += 1;
        size -= 1;
    }
    return ~value;
}

static int64_t g_min_bytes = -1;
static uint8_t *g_dev;    /* device staging, grown on demand */
static uint8_t *g_pinned; /* pinned bounce buffer for pageable callers */
static size_t g_cap;
#define CRC_PIECE ((size_t)64 << 20)

uint32_t mz_crypt_crc32_update(uint32_t value, const uint8_t *buf, int32_t size) {
    if (size <= 0 || !buf)
        return value;
    if (g_min_bytes < 0) {
        const char *v = getenv("MZ_CUDA_CRC_MIN_BYTES");
        g_min_bytes = (v && *v) ? atoll(v) : (1 << 20);
        if (g_min_bytes < 0)
            g_min_bytes = 1 << 20;
    }
    if ((int64_t)size < g_min_bytes)
        return crc_small(value, buf, size);
    if (mz_cuda_init() != MZ_OK) {
        fprintf(stderr, "mz_crypt_crc32_update: no usable sm_100 GPU (%s); refusing to fall back on the CPU for %d bytes\n",
                mz_cuda_last_error(), size);
        abort();
    }
    size_t need = (size_t)size < CRC_PIECE ? (size_t)size : CRC_PIECE;
    if (g_cap < need) {
        mz_cuda_free(g_dev);
        mz_cuda_host_free(g_pinned);
        g_dev = (uint8_t *)mz_cuda_malloc(need + 64);
        g_pinned = (uint8_t *)mz_cuda_host_alloc(need);
        g_cap = (g_dev && g_pinned) ? need : 0;
        if (!g_cap) {
            fprintf(stderr, "mz_crypt_crc32_update: cannot allocate %zu bytes of staging\n", need);
            abort();
        }
    }
    int pinned = mz_cuda_host_is_pinned(buf);
    size_t pos = 0;
    while (pos < (size_t)size) {
        size_t n = (size_t)size - pos < g_cap ? (size_t)size - pos : g_cap;
        const uint8_t *src = buf + pos;
        if (!pinned) {
            memcpy(g_pinned, src, n);
            src = g_pinned;
        }
        if (mz_cuda_memcpy_h2d(g_dev, src, n, NULL) != MZ_OK || mz_cuda_crc32_device(g_dev, n, value, &value) != MZ_OK) {
            fprintf(stderr, "mz_crypt_crc32_update: CUDA failure: %s\n", mz_cuda_last_error());
            abort();
        }
        pos += n;
    }
    return value;
}
