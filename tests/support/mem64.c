/* mem64.c -- a 64-bit in-memory base stream for tests and bench (TEST/BENCH SUPPORT, not product).
 *
 * The reference's mz_stream_mem keeps int32 sizes and positions (mz_strm_mem.c:40-48), so it cannot sit
 * under a codec stream for the 4 GiB / 16 GiB configurations. This is the ~60-line
 * vtbl implementation the survey calls for: a growable sink / fixed source with 64-bit offsets,
 * speaking the same plug-in ABI (include/mz_abi.h).
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/mz_abi.h"

/* ---- optional multi-threaded copy for large sink writes ------------------------------------------------------
 * One host thread memcpy's ~10 GB/s; a consumer that absorbs the compressed stream at memory bandwidth (several
 * copy engines / threads, what the reference arm of bench.py gets for free from its one-sink-per-thread layout) is
 * modelled by mz_stream_mem64_set_copy_threads(sink, n). Default n = 1: a plain memcpy like mz_stream_mem. */
#define PM_MAX 16
static struct {
    pthread_t th[PM_MAX];
    pthread_mutex_t mu, op;
    pthread_cond_t go, fin;
    uint8_t *dst;
    const uint8_t *src;
    size_t n;
    unsigned gen, done, nthreads, started;
} pm = {.mu = PTHREAD_MUTEX_INITIALIZER, .op = PTHREAD_MUTEX_INITIALIZER, .go = PTHREAD_COND_INITIALIZER, .fin = PTHREAD_COND_INITIALIZER};

static void *pm_worker(void *arg) {
    const unsigned idx = (unsigned)(uintptr_t)arg;
    unsigned seen = 0;
    for (;;) {
        pthread_mutex_lock(&pm.mu);
        while (pm.gen == seen) pthread_cond_wait(&pm.go, &pm.mu);
        seen = pm.gen;
        uint8_t *dst = pm.dst;
        const uint8_t *src = pm.src;
        const size_t n = pm.n;
        const unsigned nt = pm.nthreads;
        pthread_mutex_unlock(&pm.mu);
        if (idx < nt) {
            const size_t per = (n / nt + 63) & ~(size_t)63, lo = (size_t)idx * per;
            if (lo < n) memcpy(dst + lo, src + lo, lo + per > n ? n - lo : per);
        }
        pthread_mutex_lock(&pm.mu);
        if (++pm.done == pm.started) pthread_cond_signal(&pm.fin);
        pthread_mutex_unlock(&pm.mu);
    }
    return NULL;
}

static void par_memcpy(uint8_t *dst, const uint8_t *src, size_t n, unsigned nthreads) {
    if (nthreads <= 1 || n < (4u << 20)) {
        memcpy(dst, src, n);
        return;
    }
    if (nthreads > PM_MAX) nthreads = PM_MAX;
    pthread_mutex_lock(&pm.op); /* one parallel copy at a time */
    pthread_mutex_lock(&pm.mu);
    while (pm.started < nthreads) {
        if (pthread_create(&pm.th[pm.started], NULL, pm_worker, (void *)(uintptr_t)pm.started) != 0) break;
        pm.started++;
    }
    if (pm.started == 0) {
        pthread_mutex_unlock(&pm.mu);
        pthread_mutex_unlock(&pm.op);
        memcpy(dst, src, n);
        return;
    }
    pm.dst = dst; pm.src = src; pm.n = n;
    pm.nthreads = nthreads < pm.started ? nthreads : pm.started;
    pm.done = 0;
    pm.gen++;
    pthread_cond_broadcast(&pm.go);
    while (pm.done < pm.started) pthread_cond_wait(&pm.fin, &pm.mu);
    pthread_mutex_unlock(&pm.mu);
    pthread_mutex_unlock(&pm.op);
}

typedef struct mem64_s {
    mz_stream stream;
    uint8_t *buf;
    int64_t size;     /* valid bytes */
    int64_t cap;
    int64_t pos;
    int32_t owns;
    int32_t open;
    int32_t discard;  /* sink that only counts */
    int32_t copy_threads;
} mem64;

static int32_t m_open(void *s, const char *path, int32_t mode) { (void)path; (void)mode; ((mem64 *)s)->open = 1; return MZ_OK; }
static int32_t m_is_open(void *s) { return ((mem64 *)s)->open ? MZ_OK : MZ_OPEN_ERROR; }
static int32_t m_read(void *s, void *buf, int32_t size) {
    mem64 *m = (mem64 *)s;
    int64_t left = m->size - m->pos;
    if (size > left) size = (int32_t)left;
    if (size <= 0) return 0;
    memcpy(buf, m->buf + m->pos, (size_t)size);
    m->pos += size;
    return size;
}
static int32_t m_write(void *s, const void *buf, int32_t size) {
    mem64 *m = (mem64 *)s;
    if (size <= 0) return 0;
    if (!m->discard) {
        if (m->pos + size > m->cap) {
            if (!m->owns && m->buf) return MZ_WRITE_ERROR;
            int64_t ncap = m->cap ? m->cap * 2 : (1 << 20);
            while (ncap < m->pos + size) ncap *= 2;
            uint8_t *nb = (uint8_t *)realloc(m->buf, (size_t)ncap);
            if (!nb) return MZ_WRITE_ERROR;
            m->buf = nb; m->cap = ncap; m->owns = 1;
        }
        par_memcpy(m->buf + m->pos, (const uint8_t *)buf, (size_t)size, (unsigned)m->copy_threads);
    }
    m->pos += size;
    if (m->pos > m->size) m->size = m->pos;
    return size;
}
static int64_t m_tell(void *s) { return ((mem64 *)s)->pos; }
static int32_t m_seek(void *s, int64_t off, int32_t origin) {
    mem64 *m = (mem64 *)s;
    int64_t np = origin == MZ_SEEK_SET ? off : (origin == MZ_SEEK_CUR ? m->pos + off : m->size + off);
    if (np < 0 || np > m->size) return MZ_SEEK_ERROR;
    m->pos = np;
    return MZ_OK;
}
static int32_t m_close(void *s) { ((mem64 *)s)->open = 0; return MZ_OK; }
static int32_t m_error(void *s) { (void)s; return MZ_OK; }
static int32_t m_get_prop(void *s, int32_t prop, int64_t *v) { (void)s; (void)prop; (void)v; return MZ_EXIST_ERROR; }
static int32_t m_set_prop(void *s, int32_t prop, int64_t v) { (void)s; (void)prop; (void)v; return MZ_EXIST_ERROR; }
void *mz_stream_mem64_create(void);
void mz_stream_mem64_delete(void **stream);

static mz_stream_vtbl mem64_vtbl = {m_open, m_is_open, m_read, m_write, m_tell, m_seek, m_close, m_error,
                                    mz_stream_mem64_create, mz_stream_mem64_delete, m_get_prop, m_set_prop};

void *mz_stream_mem64_create(void) {
    mem64 *m = (mem64 *)calloc(1, sizeof(mem64));
    if (m) { m->stream.vtbl = &mem64_vtbl; m->open = 1; }
    return m;
}
void mz_stream_mem64_delete(void **stream) {
    if (!stream || !*stream) return;
    mem64 *m = (mem64 *)*stream;
    if (m->owns) free(m->buf);
    free(m);
    *stream = NULL;
}
/* read source over caller memory (not copied) */
void mz_stream_mem64_set_buffer(void *s, void *buf, int64_t size) {
    mem64 *m = (mem64 *)s;
    if (m->owns) free(m->buf);
    m->buf = (uint8_t *)buf; m->size = m->cap = size; m->pos = 0; m->owns = 0;
}
/* sink into caller memory of fixed capacity (not copied), e.g. pinned */
void mz_stream_mem64_set_sink(void *s, void *buf, int64_t cap) {
    mem64 *m = (mem64 *)s;
    if (m->owns) free(m->buf);
    m->buf = (uint8_t *)buf; m->cap = cap; m->size = 0; m->pos = 0; m->owns = 0;
}
void mz_stream_mem64_set_discard(void *s, int32_t on) { ((mem64 *)s)->discard = on; }
void mz_stream_mem64_set_copy_threads(void *s, int32_t n) { ((mem64 *)s)->copy_threads = n; }
int64_t mz_stream_mem64_get_buffer(void *s, const void **buf) {
    mem64 *m = (mem64 *)s;
    if (buf) *buf = m->buf;
    return m->size;
}
/* ---- generic dispatch helpers so Python drives ANY stream without the reference library present ---- */
int32_t mzt_open(void *s, const char *p, int32_t mode) { return ((mz_stream *)s)->vtbl->open(s, p, mode); }
int32_t mzt_is_open(void *s) { return ((mz_stream *)s)->vtbl->is_open(s); }
int32_t mzt_read(void *s, void *b, int32_t n) { return mz_abi_base_read(s, b, n); }
int32_t mzt_write(void *s, const void *b, int32_t n) { return mz_abi_base_write(s, b, n); }
int64_t mzt_tell(void *s) { return ((mz_stream *)s)->vtbl->tell(s); }
int32_t mzt_seek(void *s, int64_t o, int32_t w) { return ((mz_stream *)s)->vtbl->seek(s, o, w); }
int32_t mzt_close(void *s) { return ((mz_stream *)s)->vtbl->close(s); }
int32_t mzt_error(void *s) { return ((mz_stream *)s)->vtbl->error(s); }
int32_t mzt_get_prop(void *s, int32_t p, int64_t *v) { return ((mz_stream *)s)->vtbl->get_prop_int64(s, p, v); }
int32_t mzt_set_prop(void *s, int32_t p, int64_t v) { return ((mz_stream *)s)->vtbl->set_prop_int64(s, p, v); }
void mzt_set_base(void *s, void *base) { ((mz_stream *)s)->base = (mz_stream *)base; }
void mzt_delete(void **s) { if (s && *s) ((mz_stream *)*s)->vtbl->destroy(s); }
/* write a large host buffer through the stream in calls of at most `piece` bytes (callers use 16 KiB / 64 KiB / 1 GiB) */
int64_t mzt_write_all(void *s, const uint8_t *buf, int64_t len, int32_t piece) {
    int64_t pos = 0;
    while (pos < len) {
        int32_t n = len - pos > piece ? piece : (int32_t)(len - pos);
        int32_t r = mz_abi_base_write(s, buf + pos, n);
        if (r != n) return r < 0 ? r : MZ_WRITE_ERROR;
        pos += n;
    }
    return pos;
}
/* read until end of stream into buf (capacity cap); returns bytes read or negative error */
int64_t mzt_read_all(void *s, uint8_t *buf, int64_t cap, int32_t piece) {
    int64_t pos = 0;
    for (;;) {
        int32_t want = cap - pos > piece ? piece : (int32_t)(cap - pos);
        if (want <= 0) {
            uint8_t tmp[1];
            int32_t r = mz_abi_base_read(s, tmp, 1);
            return r == 0 ? pos : (r < 0 ? r : MZ_BUF_ERROR);
        }
        int32_t r = mz_abi_base_read(s, buf + pos, want);
        if (r < 0) return r;
        if (r == 0) return pos;
        pos += r;
    }
}
