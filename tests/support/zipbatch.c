/* zipbatch.c -- config C4 driver (TEST / BENCH INFRASTRUCTURE): N in-memory entries into one zip archive,
 * either through the product's batch writer (mz_zip_cuda_add_buffers, GPU codec + the reference's raw-entry seam)
 * or through the unmodified reference path (mz_zip_entry_write_open raw=0 -> mz_stream_zlib, one entry at a time).
 * Compiled against the reference's own headers and objects by oracle/Makefile (-> oracle/_ref/zipbatch_cuda); the
 * container code in both modes is the reference's.
 *
 *   zipbatch_cuda <out.zip> <entries> <entry_bytes> <level> <cuda|cuda_sha|native|native_sha|native_aes|native_all|ref> [dump_dir dump_every]
 *   zipbatch_cuda <in.zip>  <entries> <entry_bytes> <level> <extract|extract_ref|extract_aes|extract_ref_aes>   (batch extractor / the reference's loop; _aes: password $ZIPBATCH_PASSWORD or "secret")
 *
 * Entry i: i%10 < 7 text-like, < 9 binary records, else incompressible; name e/%06d.
 * With dump_dir, every dump_every-th entry's plain bytes are also written to dump_dir/%06d for comparison.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "mz.h"
#include "mz_strm.h"
#include "mz_os.h"
#include "mz_strm_buf.h"
#include "mz_strm_os.h"
#include "mz_zip.h"

#include "mz_zip_cuda.h"

#ifndef ZIPBATCH_NO_CUDA
#include "mz_cuda_batch.h"
static void zb_cuda_warm(void) { mz_cuda_init(); }
#else
static void zb_cuda_warm(void) {}
#endif
#ifdef ZIPBATCH_NO_CUDA
/* reference-only build (oracle/_ref/zipbatch_ref: the reference arm of bench.py, which must not map the product library):
 * the modes `ref` and `extract_ref` work, the product's entry points are absent */
int32_t mz_zip_cuda_add_buffers_ex(void *z, const mz_cuda_zip_item *it, uint32_t n, int16_t level, uint32_t flags, mz_cuda_zip_stats *st) {
    (void)z; (void)it; (void)n; (void)level; (void)flags; (void)st;
    return MZ_SUPPORT_ERROR;
}
int32_t mz_zip_cuda_extract_all(void *z, mz_cuda_zip_entry_cb cb, void *u, mz_cuda_zip_stats *st) {
    (void)z; (void)cb; (void)u; (void)st;
    return MZ_SUPPORT_ERROR;
}
int32_t mz_zip_cuda_extract_all_aes(void *z, const char *pw, mz_cuda_zip_entry_cb cb, void *u, mz_cuda_zip_stats *st) {
    (void)z; (void)pw; (void)cb; (void)u; (void)st;
    return MZ_SUPPORT_ERROR;
}
int32_t mz_zip_cuda_write_archive(void *b, const mz_cuda_zip_item *it, uint32_t n, int16_t level, uint32_t flags, mz_cuda_zip_stats *st) {
    (void)b; (void)it; (void)n; (void)level; (void)flags; (void)st;
    return MZ_SUPPORT_ERROR;
}
int32_t mz_zip_cuda_write_archive_aes(void *b, const mz_cuda_zip_item *it, uint32_t n, int16_t level, uint32_t flags, const char *pw, uint8_t strength,
                                      mz_cuda_zip_stats *st) {
    (void)b; (void)it; (void)n; (void)level; (void)flags; (void)pw; (void)strength; (void)st;
    return MZ_SUPPORT_ERROR;
}
uint32_t mz_zip_cuda_abi_file_info_size(void) { return (uint32_t)sizeof(mz_zip_file); }
#endif

static uint64_t rng_state;
static inline uint64_t rng(void) { /* splitmix64 */
    uint64_t z = (rng_state += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

static char vocab[4096][12];
static int vocab_len[4096];
static void make_vocab(void) {
    rng_state = 1234;
    for (int i = 0; i < 4096; i++) {
        int n = 2 + (int)(rng() % 9);
        for (int k = 0; k < n; k++) vocab[i][k] = (char)('a' + rng() % 26);
        vocab_len[i] = n;
    }
}

static void gen_entry(uint8_t *p, size_t n, uint32_t idx) {
    rng_state = 0xC4C4ull * 1000003ull + idx;
    const uint32_t kind = idx % 10;
    size_t o = 0;
    if (kind < 7) { /* words with a skewed rank distribution, some punctuation and digits */
        while (o < n) {
            uint64_t r = rng();
            uint32_t rank = (uint32_t)((r & 0xfff) * ((r >> 12) & 0xfff) >> 12); /* product of two uniforms: skewed to small ranks */
            int w = (int)rank & 4095, L = vocab_len[w];
            for (int k = 0; k < L && o < n; k++) p[o++] = (uint8_t)vocab[w][k];
            if (o < n) p[o++] = (r >> 40) % 17 == 0 ? '\n' : ((r >> 44) % 23 == 0 ? (uint8_t)('0' + (r >> 50) % 10) : ' ');
        }
    } else if (kind < 9) { /* 48-byte records: counter, a few random fields, constant padding */
        uint32_t ctr = idx * 7919u;
        while (o < n) {
            uint8_t rec[48];
            memset(rec, 0x20, sizeof(rec));
            memcpy(rec, &ctr, 4);
            uint64_t r = rng();
            memcpy(rec + 8, &r, 3);
            rec[16] = (uint8_t)(r >> 32) & 3;
            ctr++;
            size_t k = n - o < sizeof(rec) ? n - o : sizeof(rec);
            memcpy(p + o, rec, k);
            o += k;
        }
    } else {
        while (o + 8 <= n) { uint64_t r = rng(); memcpy(p + o, &r, 8); o += 8; }
        while (o < n) p[o++] = (uint8_t)rng();
    }
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + (double)ts.tv_nsec * 1e-9;
}

/* ---- extraction modes ------------------------------------------------------------------------------------------ */
typedef struct xstate_s {
    uint64_t entries, bytes, verified, mismatches;
    size_t esz;
    uint8_t *scratch;
} xstate;

static int32_t on_entry(void *ud, const char *name, const void *data, int64_t size, uint32_t crc) {
    xstate *x = (xstate *)ud;
    (void)crc;
    x->entries++;
    x->bytes += (uint64_t)size;
    unsigned idx = 0;
    if (sscanf(name, "e/%u", &idx) == 1 && idx % 101 == 0 && (size_t)size <= x->esz) { /* regenerate and compare a sample */
        gen_entry(x->scratch, (size_t)size, idx);
        x->verified++;
        if (size && memcmp(x->scratch, data, (size_t)size) != 0) x->mismatches++;
    }
    return MZ_OK;
}

static int extract_main(const char *path, size_t esz, int use_cuda, const char *password) {
    void *file_stream = mz_stream_os_create();
    void *stream = mz_stream_buffered_create();
    void *zip = mz_zip_create();
    mz_stream_set_base(stream, file_stream);
    int32_t err = mz_stream_open(stream, path, MZ_OPEN_MODE_READ);
    if (err == MZ_OK) err = mz_zip_open(zip, stream, MZ_OPEN_MODE_READ);
    if (err != MZ_OK) { fprintf(stderr, "open failed %d\n", err); return 5; }
    xstate x;
    memset(&x, 0, sizeof(x));
    x.esz = esz;
    x.scratch = (uint8_t *)malloc(esz + 16);
    make_vocab();
    mz_cuda_zip_stats st;
    memset(&st, 0, sizeof(st));
    double t0 = now_s();
    if (use_cuda) {
        err = password ? mz_zip_cuda_extract_all_aes(zip, password, on_entry, &x, &st) : mz_zip_cuda_extract_all(zip, on_entry, &x, &st);
    } else { /* the reference's own loop: one mz_stream_zlib per entry, CRC checked by mz_zip_entry_close (mz_zip_rw.c:818-909) */
        uint8_t *buf = (uint8_t *)malloc(esz + 65536);
        err = mz_zip_goto_first_entry(zip);
        while (err == MZ_OK) {
            mz_zip_file *fi = NULL;
            err = mz_zip_entry_get_info(zip, &fi);
            if (err != MZ_OK) break;
            err = mz_zip_entry_read_open(zip, 0, password);
            if (err != MZ_OK) break;
            int64_t got = 0;
            for (;;) {
                int32_t r = mz_zip_entry_read(zip, buf + got, 65536);
                if (r < 0) { err = r; break; }
                if (r == 0) break;
                got += r;
                if ((size_t)got > esz) { err = MZ_BUF_ERROR; break; }
            }
            if (err == MZ_OK) on_entry(&x, fi->filename, buf, got, fi->crc);
            int32_t cerr = mz_zip_entry_close(zip);
            if (err == MZ_OK) err = cerr;
            if (err != MZ_OK) break;
            err = mz_zip_goto_next_entry(zip);
        }
        if (err == MZ_END_OF_LIST) err = MZ_OK;
        free(buf);
    }
    double dt = now_s() - t0;
    mz_zip_close(zip);
    mz_stream_close(stream);
    mz_zip_delete(&zip);
    mz_stream_buffered_delete(&stream);
    mz_stream_os_delete(&file_stream);
    printf("{\"mode\": \"%s\", \"err\": %d, \"entries\": %llu, \"bytes_out\": %llu, \"verified\": %llu, \"mismatches\": %llu, \"s\": %.4f, "
           "\"entries_per_s\": %.0f, \"GiB_per_s\": %.3f, \"read_ms\": %.1f, \"gpu_ms\": %.1f, \"deliver_ms\": %.1f, \"rounds\": %u}\n",
           use_cuda ? "extract_cuda" : "extract_ref", err, (unsigned long long)x.entries, (unsigned long long)x.bytes, (unsigned long long)x.verified,
           (unsigned long long)x.mismatches, dt, x.entries / dt, (double)x.bytes / (1ull << 30) / dt, st.pack_ms, st.gpu_ms, st.container_ms, st.rounds);
    free(x.scratch);
    return err == MZ_OK && x.mismatches == 0 ? 0 : 1;
}

int main(int argc, char **argv) {
    if (argc >= 6 && strncmp(argv[5], "extract", 7) == 0) { /* extract | extract_ref | extract_aes | extract_ref_aes */
        if (sizeof(mz_zip_file) != mz_zip_cuda_abi_file_info_size()) return 3;
        const char *pw = strstr(argv[5], "aes") ? (getenv("ZIPBATCH_PASSWORD") ? getenv("ZIPBATCH_PASSWORD") : "secret") : NULL;
        return extract_main(argv[1], (size_t)atoll(argv[3]), strstr(argv[5], "ref") == NULL, pw);
    }
    if (argc < 6) {
        fprintf(stderr, "usage: %s out.zip entries entry_bytes level cuda|ref [dump_dir dump_every]\n", argv[0]);
        return 2;
    }
    const char *path = argv[1];
    const uint32_t n = (uint32_t)atoi(argv[2]);
    const size_t esz = (size_t)atoll(argv[3]);
    const int16_t level = (int16_t)atoi(argv[4]);
    const int use_sha = strcmp(argv[5], "cuda_sha") == 0; /* + SHA-256 extra field per entry (scope row f3) */
    /* native / native_sha / native_all: the product writes the WHOLE archive itself (mz_zip_cuda_write_archive) to the file stream */
    const int use_native = strncmp(argv[5], "native", 6) == 0;
    const uint32_t native_flags = (strstr(argv[5], "sha") ? MZ_ZIP_CUDA_HASH_SHA256 : 0u) | (strstr(argv[5], "all") ? MZ_ZIP_CUDA_ALL_DEVICES : 0u) |
                                  (strstr(argv[5], "aes") ? MZ_ZIP_CUDA_AES : 0u);
    const char *password = getenv("ZIPBATCH_PASSWORD") ? getenv("ZIPBATCH_PASSWORD") : "secret";
    const int use_cuda = strcmp(argv[5], "cuda") == 0 || use_sha;
    const char *dump_dir = argc > 7 ? argv[6] : NULL;
    const uint32_t dump_every = argc > 7 ? (uint32_t)atoi(argv[7]) : 0;
    if (sizeof(mz_zip_file) != mz_zip_cuda_abi_file_info_size()) {
        fprintf(stderr, "mz_zip_file layout mismatch: %zu vs %u\n", sizeof(mz_zip_file), mz_zip_cuda_abi_file_info_size());
        return 3;
    }
    make_vocab();
    uint8_t *data = (uint8_t *)malloc((size_t)n * esz + 16);
    char *names = (char *)malloc((size_t)n * 16);
    mz_cuda_zip_item *items = (mz_cuda_zip_item *)calloc(n ? n : 1, sizeof(*items));
    if (!data || !names || !items) return 4;
    double t0 = now_s();
    for (uint32_t i = 0; i < n; i++) {
        /* a few entries deviate in size (empty, tiny, ragged) so the batch is not perfectly uniform */
        size_t sz = esz;
        if (i % 997 == 5) sz = 0;
        else if (i % 997 == 6) sz = 1;
        else if (i % 251 == 7) sz = esz - esz / 3 - 1;
        gen_entry(data + (size_t)i * esz, sz, i);
        snprintf(names + (size_t)i * 16, 16, "e/%06u", i);
        items[i].filename = names + (size_t)i * 16;
        items[i].data = data + (size_t)i * esz;
        items[i].size = (int64_t)sz;
        items[i].modified_date = 1700000000;
        if (dump_dir && dump_every && i % dump_every == 0) {
            char f[1200];
            snprintf(f, sizeof(f), "%s/%06u", dump_dir, i);
            FILE *fp = fopen(f, "wb");
            if (fp) { fwrite(items[i].data, 1, sz, fp); fclose(fp); }
        }
    }
    double t_gen = now_s() - t0;

    /* ZIPBATCH_REPEAT=N: write the archive N times in this process, one JSON line each (a long-lived writer: the library keeps its
     * staging between calls; bench.py's warm-up and timed steps) */
    const int repeat = getenv("ZIPBATCH_REPEAT") && atoi(getenv("ZIPBATCH_REPEAT")) > 0 ? atoi(getenv("ZIPBATCH_REPEAT")) : 1;
    int rc = 0;
    for (int rep = 0; rep < repeat && rc == 0; rep++) {
        /* file <- buffered stream <- zip, as the reference's own writer stacks them (mz_zip_rw.c:1205-1222) */
        void *file_stream = mz_stream_os_create();
        void *stream = mz_stream_buffered_create();
        void *zip = mz_zip_create();
        mz_stream_set_base(stream, file_stream);
        if (use_native) { /* the native writer hands over whole rounds (tens of MiB per call): no 32 KiB buffering layer in between */
            mz_stream_buffered_delete(&stream);
            stream = file_stream;
        }
        int32_t err = mz_stream_open(stream, path, MZ_OPEN_MODE_CREATE | MZ_OPEN_MODE_WRITE);
        if (err == MZ_OK && !use_native) err = mz_zip_open(zip, stream, MZ_OPEN_MODE_WRITE);
        if (err != MZ_OK) { fprintf(stderr, "open failed %d\n", err); return 5; }
        mz_cuda_zip_stats st;
        memset(&st, 0, sizeof(st));
        uint64_t bytes_in = 0;
        /* CUDA context creation (a few hundred ms, once per process) is reported on its own, not inside the throughput figure */
        t0 = now_s();
        if (use_native || use_cuda) zb_cuda_warm();
        const double t_init = now_s() - t0;
        t0 = now_s();
        if (use_native) {
            err = (native_flags & MZ_ZIP_CUDA_AES) ? mz_zip_cuda_write_archive_aes(stream, items, n, level, native_flags, password, 0, &st)
                                                   : mz_zip_cuda_write_archive(stream, items, n, level, native_flags, &st);
            bytes_in = st.bytes_in;
        } else if (use_cuda) {
            err = mz_zip_cuda_add_buffers_ex(zip, items, n, level, use_sha ? MZ_ZIP_CUDA_HASH_SHA256 : 0u, &st);
            bytes_in = st.bytes_in;
        } else {
            for (uint32_t i = 0; i < n && err == MZ_OK; i++) {
                mz_zip_file fi;
                memset(&fi, 0, sizeof(fi));
                fi.version_madeby = MZ_VERSION_MADEBY;
                fi.flag = MZ_ZIP_FLAG_UTF8;
                fi.compression_method = MZ_COMPRESS_METHOD_DEFLATE;
                fi.modified_date = 1700000000;
                fi.filename = items[i].filename;
                fi.uncompressed_size = items[i].size;
                err = mz_zip_entry_write_open(zip, &fi, level, 0, NULL);
                int64_t done = 0;
                while (err == MZ_OK && done < items[i].size) { /* the reference's writers feed <= 64 KiB at a time (mz_zip_rw.c:1424) */
                    int32_t piece = items[i].size - done > 65536 ? 65536 : (int32_t)(items[i].size - done);
                    int32_t w = mz_zip_entry_write(zip, (const uint8_t *)items[i].data + done, piece);
                    if (w != piece) err = w < 0 ? w : MZ_WRITE_ERROR;
                    done += piece;
                }
                if (err == MZ_OK) err = mz_zip_entry_close(zip);
                bytes_in += (uint64_t)items[i].size;
            }
        }
        double t_add = now_s() - t0;
        t0 = now_s();
        int32_t cerr = use_native ? MZ_OK : mz_zip_close(zip);
        mz_stream_close(stream);
        double t_close = now_s() - t0;
        mz_zip_delete(&zip);
        if (stream != file_stream)
            mz_stream_buffered_delete(&stream);
        mz_stream_os_delete(&file_stream);
        printf("{\"mode\": \"%s\", \"entries\": %u, \"entry_bytes\": %zu, \"level\": %d, \"err\": %d, \"close_err\": %d, \"bytes_in\": %llu, "
               "\"bytes_out\": %llu, \"gen_s\": %.3f, \"add_s\": %.4f, \"close_s\": %.4f, \"entries_per_s\": %.0f, \"GiB_per_s\": %.3f, "
               "\"pack_ms\": %.1f, \"gpu_ms\": %.1f, \"container_ms\": %.1f, \"setup_ms\": %.1f, \"cuda_init_s\": %.3f, \"rounds\": %u}\n",
               use_native ? argv[5] : (use_cuda ? "cuda" : "ref"), n, esz, level, err, cerr, (unsigned long long)bytes_in, (unsigned long long)st.bytes_out, t_gen, t_add,
               t_close, n / (t_add + t_close), (double)bytes_in / (1ull << 30) / (t_add + t_close), st.pack_ms, st.gpu_ms, st.container_ms, st.setup_ms, t_init, st.rounds);
        fflush(stdout);
        rc = err == MZ_OK && cerr == MZ_OK ? 0 : 1;
    }
    free(data);
    free(names);
    free(items);
    return rc;
}
