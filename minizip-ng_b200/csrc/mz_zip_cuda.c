/* mz_zip_cuda.c -- batch zip-entry writer on the reference's raw-entry seam (include/mz_zip_cuda.h).
 *
 * What it replaces, per entry, in the reference: mz_zip_entry_write_open(raw = 0) creating a mz_stream_zlib
 * (mz_zip.c:1771-1773), the caller's write loop feeding it (mz_zip_rw.c:1424-1480), the CRC accumulation
 * (mz_zip.c:2062-2064) and mz_zip_entry_close (mz_zip.c:2116-2160). Here the codec work of MANY entries is one
 * K2+K3 launch + one K1 launch + one K4 launch per round, and the container calls run with raw = 1.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "mz_abi.h"
#include "mz_cuda_batch.h"
#include "mz_zip_cuda.h"

#define ZC_CHUNK 65536u

/* mz_zip.h:25-51 (mz_zip_file): the raw seam takes this struct by pointer and copies it (mz_zip.c:1939) */
typedef struct zc_file_info_s {
    uint16_t version_madeby;
    uint16_t version_needed;
    uint16_t flag;
    uint16_t compression_method;
    time_t modified_date;
    time_t accessed_date;
    time_t creation_date;
    uint32_t crc;
    int64_t compressed_size;
    int64_t uncompressed_size;
    uint16_t filename_size;
    uint16_t extrafield_size;
    uint16_t comment_size;
    uint32_t disk_number;
    int64_t disk_offset;
    uint16_t internal_fa;
    uint32_t external_fa;
    const char *filename;
    const uint8_t *extrafield;
    const char *comment;
    const char *linkname;
    uint16_t zip64;
    uint16_t aes_version;
    uint8_t aes_strength;
    uint16_t pk_verify;
} zc_file_info;

/* the container stays the reference's: these come from the host program (mz_zip.c:1915, 2056, 2272) */
extern int32_t mz_zip_entry_write_open(void *handle, const void *file_info, int16_t compress_level, uint8_t raw, const char *password)
    __attribute__((weak));
extern int32_t mz_zip_entry_write(void *handle, const void *buf, int32_t len) __attribute__((weak));
extern int32_t mz_zip_entry_close_raw(void *handle, int64_t uncompressed_size, uint32_t crc32) __attribute__((weak));

uint32_t mz_zip_cuda_abi_file_info_size(void) {
    return (uint32_t)sizeof(zc_file_info);
}

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec * 1e3 + (double)ts.tv_nsec * 1e-6;
}

typedef struct zc_bufs_s {
    size_t round_bytes;
    uint32_t max_chunks;
    uint64_t stride;
    uint8_t *h_in, *d_in, *d_slots, *d_out, *h_out;
    uint64_t *h_off, *d_off, *d_joined_off, *h_joined_off;
    uint32_t *h_len, *d_len, *d_out_len, *d_residue, *d_crc, *h_crc;
    uint8_t *h_flags, *d_flags;
} zc_bufs;

static void zc_free(zc_bufs *b) {
    mz_cuda_host_free(b->h_in);
    mz_cuda_free(b->d_in);
    mz_cuda_free(b->d_slots);
    mz_cuda_free(b->d_out);
    mz_cuda_host_free(b->h_out);
    mz_cuda_host_free(b->h_off);
    mz_cuda_free(b->d_off);
    mz_cuda_free(b->d_joined_off);
    mz_cuda_host_free(b->h_joined_off);
    mz_cuda_host_free(b->h_len);
    mz_cuda_free(b->d_len);
    mz_cuda_free(b->d_out_len);
    mz_cuda_free(b->d_residue);
    mz_cuda_free(b->d_crc);
    mz_cuda_host_free(b->h_crc);
    mz_cuda_host_free(b->h_flags);
    mz_cuda_free(b->d_flags);
    memset(b, 0, sizeof(*b));
}

static int zc_alloc(zc_bufs *b, size_t round_bytes, uint32_t max_chunks) {
    memset(b, 0, sizeof(*b));
    b->round_bytes = round_bytes;
    b->max_chunks = max_chunks;
    b->stride = mz_cuda_deflate_slot_bound(ZC_CHUNK);
    const size_t slots = (size_t)max_chunks * b->stride;
    b->h_in = (uint8_t *)mz_cuda_host_alloc(round_bytes + 64);
    b->d_in = (uint8_t *)mz_cuda_malloc(round_bytes + 64);
    b->d_slots = (uint8_t *)mz_cuda_malloc(slots);
    b->d_out = (uint8_t *)mz_cuda_malloc(slots);
    b->h_out = (uint8_t *)mz_cuda_host_alloc(slots);
    b->h_off = (uint64_t *)mz_cuda_host_alloc((size_t)max_chunks * 8);
    b->d_off = (uint64_t *)mz_cuda_malloc((size_t)max_chunks * 8);
    b->d_joined_off = (uint64_t *)mz_cuda_malloc(((size_t)max_chunks + 1) * 8);
    b->h_joined_off = (uint64_t *)mz_cuda_host_alloc(((size_t)max_chunks + 1) * 8);
    b->h_len = (uint32_t *)mz_cuda_host_alloc((size_t)max_chunks * 4);
    b->d_len = (uint32_t *)mz_cuda_malloc((size_t)max_chunks * 4);
    b->d_out_len = (uint32_t *)mz_cuda_malloc((size_t)max_chunks * 4);
    b->d_residue = (uint32_t *)mz_cuda_malloc((size_t)max_chunks * 4);
    b->d_crc = (uint32_t *)mz_cuda_malloc((size_t)max_chunks * 4);
    b->h_crc = (uint32_t *)mz_cuda_host_alloc((size_t)max_chunks * 4);
    b->h_flags = (uint8_t *)mz_cuda_host_alloc(max_chunks);
    b->d_flags = (uint8_t *)mz_cuda_malloc(max_chunks);
    if (!b->h_in || !b->d_in || !b->d_slots || !b->d_out || !b->h_out || !b->h_off || !b->d_off || !b->d_joined_off || !b->h_joined_off ||
        !b->h_len || !b->d_len || !b->d_out_len || !b->d_residue || !b->d_crc || !b->h_crc || !b->h_flags || !b->d_flags) {
        zc_free(b);
        return 0;
    }
    return 1;
}

static uint32_t chunks_of(int64_t size) {
    return size <= 0 ? 1u : (uint32_t)(((uint64_t)size + ZC_CHUNK - 1) / ZC_CHUNK);
}

/* one round = consecutive entries that fit the staging buffers; prepared (packed, compressed, downloaded) by a
 * worker thread while the caller's thread hands the previous round to the container */
typedef struct zc_round_s {
    zc_bufs b;
    const mz_cuda_zip_item *items;
    uint32_t count, first, last, nch;
    int16_t level;
    int32_t device, err;
    double pack_ms, gpu_ms;
} zc_round;

static void *zc_prepare(void *arg) {
    zc_round *r = (zc_round *)arg;
    zc_bufs *b = &r->b;
    int32_t err = mz_cuda_set_device(r->device);
    double t0 = now_ms();
    uint32_t last = r->first, nch = 0;
    size_t pos = 0;
    while (last < r->count) {
        const size_t need = ((size_t)r->items[last].size + 15) & ~(size_t)15;
        const uint32_t c = chunks_of(r->items[last].size);
        if (last > r->first && (pos + need > b->round_bytes || nch + c > b->max_chunks))
            break;
        if (r->items[last].size > 0)
            memcpy(b->h_in + pos, r->items[last].data, (size_t)r->items[last].size);
        int64_t left = r->items[last].size;
        for (uint32_t k = 0; k < c; k++) {
            const uint32_t n = left > (int64_t)ZC_CHUNK ? ZC_CHUNK : (uint32_t)left;
            b->h_off[nch] = pos + (uint64_t)k * ZC_CHUNK;
            b->h_len[nch] = n;
            b->h_flags[nch] = k + 1 == c ? 1u : 0u; /* BFINAL on the entry's last chunk */
            left -= n;
            nch++;
        }
        pos += need;
        last++;
    }
    r->last = last;
    r->nch = nch;
    if (!err) err = mz_cuda_memcpy_h2d(b->d_in, b->h_in, pos + 16, NULL);
    if (!err) err = mz_cuda_memcpy_h2d(b->d_off, b->h_off, (size_t)nch * 8, NULL);
    if (!err) err = mz_cuda_memcpy_h2d(b->d_len, b->h_len, (size_t)nch * 4, NULL);
    if (!err) err = mz_cuda_memcpy_h2d(b->d_flags, b->h_flags, nch, NULL);
    double t1 = now_ms();
    /* device: compress, checksum, join; download */
    if (!err) err = mz_cuda_deflate_chunks(b->d_in, 0, 0, b->d_off, b->d_len, b->d_flags, nch, 0, r->level, b->d_slots, b->stride, b->d_out_len, NULL);
    if (!err) err = mz_cuda_crc32_segments(b->d_in, 0, 0, b->d_off, b->d_len, nch, b->d_residue, b->d_crc, NULL);
    if (!err) err = mz_cuda_concat(b->d_slots, b->stride, b->d_out_len, nch, b->d_joined_off, b->d_out, NULL);
    if (!err) err = mz_cuda_memcpy_d2h(b->h_joined_off, b->d_joined_off, ((size_t)nch + 1) * 8, NULL);
    if (!err) err = mz_cuda_memcpy_d2h(b->h_crc, b->d_crc, (size_t)nch * 4, NULL);
    if (!err) err = mz_cuda_stream_sync(NULL);
    if (!err) err = mz_cuda_memcpy_d2h(b->h_out, b->d_out, (size_t)b->h_joined_off[nch], NULL);
    if (!err) err = mz_cuda_stream_sync(NULL);
    r->pack_ms = t1 - t0;
    r->gpu_ms = now_ms() - t1;
    r->err = err;
    return NULL;
}

int32_t mz_zip_cuda_add_buffers(void *zip_handle, const mz_cuda_zip_item *items, uint32_t count, int16_t level, mz_cuda_zip_stats *stats) {
    zc_round rd[2];
    int32_t err = MZ_OK;
    mz_cuda_zip_stats st;
    memset(&st, 0, sizeof(st));
    memset(rd, 0, sizeof(rd));
    if (!mz_zip_entry_write_open || !mz_zip_entry_write || !mz_zip_entry_close_raw)
        return MZ_SUPPORT_ERROR; /* no zip container in this process */
    if (!zip_handle || (!items && count))
        return MZ_PARAM_ERROR;
    if (level == MZ_COMPRESS_LEVEL_DEFAULT)
        level = 6;
    if (level < 0 || level > 9)
        return MZ_PARAM_ERROR;
    if (mz_cuda_init() != MZ_OK) {
        fprintf(stderr, "mz_zip_cuda: no usable sm_100 GPU (%s); there is no CPU fallback\n", mz_cuda_last_error());
        return MZ_SUPPORT_ERROR;
    }
    for (uint32_t i = 0; i < count; i++)
        if (!items[i].filename || items[i].size < 0 || (items[i].size > 0 && !items[i].data))
            return MZ_PARAM_ERROR;
    if (count == 0) {
        if (stats)
            *stats = st;
        return MZ_OK;
    }
    size_t round_bytes = 256u << 20;
    {
        const char *v = getenv("MZ_CUDA_ZIP_ROUND_MB");
        if (v && atoll(v) > 0)
            round_bytes = (size_t)atoll(v) << 20;
    }
    /* an entry is never split over rounds: the round must hold the largest one */
    for (uint32_t i = 0; i < count; i++)
        if ((size_t)items[i].size + 16 > round_bytes)
            round_bytes = ((size_t)items[i].size + 16 + 65535) & ~(size_t)65535;
    /* chunks per round: full chunks by size plus one (partial or empty) chunk per entry; size the tables for the
     * densest window of the actual input (entries are packed at 16-byte aligned offsets) */
    uint32_t max_chunks = (uint32_t)(round_bytes / ZC_CHUNK) + 1;
    {
        uint32_t j = 0, ch = 0, best = 0;
        size_t bytes = 0;
        for (uint32_t i = 0; i < count; i++) {
            bytes += ((size_t)items[i].size + 15) & ~(size_t)15;
            ch += chunks_of(items[i].size);
            while (bytes > round_bytes) {
                bytes -= ((size_t)items[j].size + 15) & ~(size_t)15;
                ch -= chunks_of(items[j].size);
                j++;
            }
            if (ch > best)
                best = ch;
        }
        if (best + 1 > max_chunks)
            max_chunks = best + 1;
    }
    const int32_t device = mz_cuda_get_device();
    for (int k = 0; k < 2; k++) {
        if (!zc_alloc(&rd[k].b, round_bytes, max_chunks)) {
            zc_free(&rd[0].b);
            return MZ_MEM_ERROR;
        }
        rd[k].items = items;
        rd[k].count = count;
        rd[k].level = level;
        rd[k].device = device < 0 ? 0 : device;
    }
    /* round r is prepared by a worker while this thread feeds round r-1 to the container */
    int cur = 0;
    pthread_t th;
    rd[cur].first = 0;
    zc_prepare(&rd[cur]);
    while (err == MZ_OK) {
        zc_round *r = &rd[cur];
        err = r->err;
        if (err)
            break;
        int have_next = r->last < count, started = 0;
        if (have_next) {
            rd[cur ^ 1].first = r->last;
            started = pthread_create(&th, NULL, zc_prepare, &rd[cur ^ 1]) == 0;
            if (!started)
                zc_prepare(&rd[cur ^ 1]);
        }
        /* ---- container: the reference writes headers and copies the finished streams --------------------------- */
        double t2 = now_ms();
        const zc_bufs *b = &r->b;
        uint32_t c0 = 0;
        for (uint32_t i = r->first; i < r->last && err == MZ_OK; i++) {
            const uint32_t c = chunks_of(items[i].size);
            uint32_t crc = b->h_crc[c0];
            for (uint32_t k = 1; k < c; k++)
                crc = mz_cuda_crc32_combine(crc, b->h_crc[c0 + k], b->h_len[c0 + k]);
            if (items[i].size == 0)
                crc = 0;
            const uint8_t *comp = b->h_out + b->h_joined_off[c0];
            const uint64_t csize = b->h_joined_off[c0 + c] - b->h_joined_off[c0];
            zc_file_info fi;
            memset(&fi, 0, sizeof(fi));
            fi.version_madeby = (3u << 8) | 45u; /* MZ_HOST_SYSTEM_UNIX (mz.h:104), zip 4.5 as MZ_VERSION_MADEBY without extra codecs (mz_os.h:27-42) */
            /* MZ_ZIP_FLAG_UTF8 (mz.h:84) | MZ_ZIP_FLAG_DATA_DESCRIPTOR (mz.h:83): like the reference's own compressed
             * entries (mz_zip.c:1992) the sizes follow the data, so close_raw never seeks back to patch the header */
            fi.flag = (1u << 11) | (1u << 3);
            fi.compression_method = 8;           /* MZ_COMPRESS_METHOD_DEFLATE, mz.h:64 */
            fi.modified_date = items[i].modified_date ? (time_t)items[i].modified_date : time(NULL);
            fi.crc = crc;
            fi.compressed_size = (int64_t)csize;
            fi.uncompressed_size = items[i].size;
            fi.external_fa = items[i].external_fa ? items[i].external_fa : (0100644u << 16);
            fi.filename = items[i].filename;
            err = mz_zip_entry_write_open(zip_handle, &fi, level, 1, NULL);
            uint64_t done = 0;
            while (err == MZ_OK && done < csize) {
                const int32_t piece = csize - done > (1u << 30) ? (int32_t)(1u << 30) : (int32_t)(csize - done);
                const int32_t w = mz_zip_entry_write(zip_handle, comp + done, piece);
                if (w != piece)
                    err = w < 0 ? w : MZ_WRITE_ERROR;
                done += (uint64_t)piece;
            }
            if (err == MZ_OK)
                err = mz_zip_entry_close_raw(zip_handle, items[i].size, crc);
            st.bytes_in += (uint64_t)items[i].size;
            st.bytes_out += csize;
            c0 += c;
        }
        st.container_ms += now_ms() - t2;
        st.pack_ms += r->pack_ms;
        st.gpu_ms += r->gpu_ms;
        st.entries += r->last - r->first;
        st.rounds++;
        if (started)
            pthread_join(th, NULL);
        if (!have_next)
            break;
        cur ^= 1;
    }
    zc_free(&rd[0].b);
    zc_free(&rd[1].b);
    if (stats)
        *stats = st;
    return err;
}
