/* textgen_host.c -- host side of the synthetic text generator (tables + multi-threaded fill). TEST / BENCH SUPPORT. */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <unistd.h>

#include "textgen.h"

static uint8_t *g_words;
static uint32_t g_off[MZT_NWORDS + 1];
static uint32_t g_zipf[MZT_ZIPF_LUT + 1];
static uint32_t g_nbytes;
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void build(void) {
    /* 50k pseudo-words of 2..10 lowercase letters, fixed seed, skewed letter frequencies so words share n-grams */
    static const char alpha[] = "etaoinshrdlcumwfgypbvkjxqz";
    uint64_t s = 1234;
    g_words = (uint8_t *)malloc((size_t)MZT_NWORDS * 10);
    uint32_t n = 0;
    for (uint32_t i = 0; i < MZT_NWORDS; i++) {
        g_off[i] = n;
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        const uint32_t len = 2 + (uint32_t)((s >> 33) % 9);
        for (uint32_t k = 0; k < len; k++) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            const uint32_t r = (uint32_t)(s >> 40) % 100;
            const uint32_t idx = r < 60 ? r % 8 : (r < 90 ? 8 + r % 10 : 18 + r % 8);
            g_words[n++] = (uint8_t)alpha[idx];
        }
    }
    g_off[MZT_NWORDS] = n;
    g_nbytes = n;
    for (int k = 0; k <= MZT_ZIPF_LUT; k++) {
        double v = 65536.0 * pow((double)MZT_NWORDS, (double)k / MZT_ZIPF_LUT);
        g_zipf[k] = v >= 4294967295.0 ? 4294967295u : (uint32_t)v;
    }
}

const uint8_t *mzt_vocab_words(uint32_t *nbytes) {
    pthread_once(&g_once, build);
    if (nbytes) *nbytes = g_nbytes;
    return g_words;
}
const uint32_t *mzt_vocab_offsets(void) { pthread_once(&g_once, build); return g_off; }
const uint32_t *mzt_zipf_lut(void) { pthread_once(&g_once, build); return g_zipf; }

typedef struct { uint8_t *out; uint64_t nbytes, seed, p0, p1; } job_t;
static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    for (uint64_t piece = j->p0; piece < j->p1; piece++) {
        const uint64_t off = piece * MZT_TEXT_PIECE;
        const uint32_t room = (uint32_t)(j->nbytes - off < MZT_TEXT_PIECE ? j->nbytes - off : MZT_TEXT_PIECE);
        mzt_piece(j->out + off, room, j->seed, piece, g_words, g_off, g_zipf);
    }
    return NULL;
}

void mzt_textgen_host(uint8_t *out, uint64_t nbytes, uint64_t seed, int threads) {
    pthread_once(&g_once, build);
    if (threads <= 0) threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (threads > 64) threads = 64;
    if (threads < 1) threads = 1;
    const uint64_t npieces = (nbytes + MZT_TEXT_PIECE - 1) / MZT_TEXT_PIECE;
    pthread_t th[64];
    job_t jobs[64];
    for (int t = 0; t < threads; t++) {
        jobs[t].out = out; jobs[t].nbytes = nbytes; jobs[t].seed = seed;
        jobs[t].p0 = npieces * (uint64_t)t / (uint64_t)threads;
        jobs[t].p1 = npieces * (uint64_t)(t + 1) / (uint64_t)threads;
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
}
