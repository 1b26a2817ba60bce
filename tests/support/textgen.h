/* textgen.h -- synthetic "enwik-style" text: TEST / BENCH SUPPORT, not product code.
 *
 * Zipf(1)-like draws from a fixed vocabulary of 50 000 pseudo-words joined by spaces, ~2 % markup / newline / digit
 * tokens. One 256-byte piece is a function of (seed, piece index) alone, so the text is position-independent, and the
 * SAME bytes come out of the host generator (textgen_host.c, used by the reference arm of bench.py, which must not touch
 * the GPU library) and of the device generator (textgen_dev.cu): everything is integer arithmetic on tables that are
 * built once on the host (no device exp2f). */
#ifndef MZT_TEXTGEN_H
#define MZT_TEXTGEN_H
#include <stdint.h>

#define MZT_TEXT_PIECE 256
#define MZT_NWORDS 50000u
#define MZT_ZIPF_LUT 4096 /* entries + 1: floor(65536 * N^(k/4096)) */

#if defined(__CUDACC__)
#define MZT_HD __host__ __device__ __forceinline__
#else
#define MZT_HD static inline
#endif

MZT_HD uint32_t mzt_next(uint64_t *s) { /* splitmix-style */
    *s += 0x9E3779B97F4A7C15ull;
    uint64_t z = *s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((z ^ (z >> 31)) >> 32);
}

/* fill out[0..room) with piece number `piece` of the text of `seed` */
MZT_HD void mzt_piece(uint8_t *p, uint32_t room, uint64_t seed, uint64_t piece, const uint8_t *words, const uint32_t *word_off, const uint32_t *zipf) {
    uint64_t s = seed * 0xD1342543DE82EF95ull + piece * 0x2545F4914F6CDD1Dull + 1;
    uint32_t pos = 0;
    while (pos < room) {
        const uint32_t r = mzt_next(&s);
        const uint32_t u = r >> 8;                       /* 24-bit uniform */
        const uint32_t k = u >> 12, f = u & 4095u;
        const uint64_t v = (uint64_t)zipf[k] * (4096u - f) + (uint64_t)zipf[k + 1] * f; /* 16.16 fixed point of N^u, x 4096 */
        uint32_t rank = (uint32_t)(v >> 28);
        if (rank >= MZT_NWORDS) rank = MZT_NWORDS - 1;
        const uint32_t a = word_off[rank], b = word_off[rank + 1];
        for (uint32_t i = a; i < b && pos < room; i++) p[pos++] = words[i];
        const uint32_t m = r & 255u;
        if (m < 3 && pos < room) p[pos++] = '.';
        if (m == 0 && pos < room) p[pos++] = '\n';
        else if (m == 1) {
            uint32_t d = mzt_next(&s);
            if (pos < room) p[pos++] = ' ';
            if (pos < room) p[pos++] = '<';
            for (int q = 0; q < 4 && pos < room; q++) { p[pos++] = (uint8_t)('0' + d % 10); d /= 10; }
            if (pos < room) p[pos++] = '>';
        }
        if (pos < room) p[pos++] = ' ';
    }
}

/* host-side table construction (textgen_host.c): words + offsets (nwords + 1) + zipf LUT (MZT_ZIPF_LUT + 1) */
#ifdef __cplusplus
extern "C" {
#endif
const uint8_t *mzt_vocab_words(uint32_t *nbytes);
const uint32_t *mzt_vocab_offsets(void);
const uint32_t *mzt_zipf_lut(void);
/* host generator: `threads` worker threads (0 = one per online CPU) */
void mzt_textgen_host(uint8_t *out, uint64_t nbytes, uint64_t seed, int threads);
#ifdef __cplusplus
}
#endif
#endif
