/* lzsim.c -- CPU model of candidate parse strategies for the deflate kernel (DESIGN TOOL, not product, not oracle).
 * Generates the bench text (same generator as tests/support textgen), parses 64 KiB chunks under a chosen
 * model and reports the exact RFC1951 cost (optimal-ish length-limited Huffman + full dynamic header).
 * build: gcc -O2 -o lzsim lzsim.c -lz -lm */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

/* ---------------- text generator (port of textgen_kernel + make_vocab) ---------------- */
static uint8_t *g_words; static uint32_t *g_off; static uint32_t g_nw = 50000;
static void make_vocab(void) {
    uint64_t s = 1234; size_t cap = 1 << 20, n = 0;
    g_words = malloc(cap); g_off = malloc((g_nw + 1) * 4);
    for (uint32_t i = 0; i < g_nw; i++) {
        g_off[i] = (uint32_t)n;
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        uint32_t len = 2 + (uint32_t)((s >> 33) % 9);
        for (uint32_t k = 0; k < len; k++) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            uint32_t r = (uint32_t)(s >> 40) % 100;
            static const char alpha[] = "etaoinshrdlcumwfgypbvkjxqz";
            uint32_t idx = r < 60 ? r % 8 : (r < 90 ? 8 + r % 10 : 18 + r % 8);
            g_words[n++] = (uint8_t)alpha[idx];
        }
    }
    g_off[g_nw] = (uint32_t)n;
}
static uint32_t tg_next(uint64_t *s) { *s = *s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(*s >> 32); }
#define TEXT_PIECE 4096
static void textgen(uint8_t *out, uint64_t nbytes, uint64_t seed) {
    float lnw = log2f((float)g_nw);
    uint64_t npieces = (nbytes + TEXT_PIECE - 1) / TEXT_PIECE;
    for (uint64_t piece = 0; piece < npieces; piece++) {
        uint64_t s = seed * 0xD1342543DE82EF95ull + piece * 0x2545F4914F6CDD1Dull + 1;
        uint8_t *p = out + piece * TEXT_PIECE;
        uint32_t room = (uint32_t)(nbytes - piece * TEXT_PIECE < TEXT_PIECE ? nbytes - piece * TEXT_PIECE : TEXT_PIECE);
        uint32_t pos = 0;
        while (pos < room) {
            uint32_t r = tg_next(&s);
            float u = (float)(r >> 8) * (1.0f / 16777216.0f);
            uint32_t rank = (uint32_t)exp2f(u * lnw);
            if (rank >= g_nw) rank = g_nw - 1;
            uint32_t a = g_off[rank], b = g_off[rank + 1];
            for (uint32_t i = a; i < b && pos < room; i++) p[pos++] = g_words[i];
            uint32_t m = r & 255;
            if (m < 3 && pos < room) p[pos++] = '.';
            if (m == 0 && pos < room) p[pos++] = '\n';
            else if (m == 1) {
                uint32_t d = tg_next(&s);
                if (pos < room) p[pos++] = ' ';
                if (pos < room) p[pos++] = '<';
                for (int k = 0; k < 4 && pos < room; k++) { p[pos++] = (uint8_t)('0' + d % 10); d /= 10; }
                if (pos < room) p[pos++] = '>';
            }
            if (pos < room) p[pos++] = ' ';
        }
    }
}

/* ---------------- Huffman cost ---------------- */
static void huff_lengths(const uint32_t *freq, int n, int maxbits, uint8_t *len) {
    /* simple O(n^2)-ish Huffman via sorted merge; then limit by the kraft heuristic */
    int idx[320], m = 0;
    for (int i = 0; i < n; i++) { len[i] = 0; if (freq[i]) idx[m++] = i; }
    if (m == 0) return;
    if (m == 1) { len[idx[0]] = 1; return; }
    /* nodes */
    static uint64_t w[640]; static int parent[640];
    int nn = m;
    for (int i = 0; i < m; i++) { w[i] = freq[idx[i]]; parent[i] = -1; }
    int alive[640], na = m;
    for (int i = 0; i < m; i++) alive[i] = i;
    while (na > 1) {
        int a = 0, b = 1;
        if (w[alive[b]] < w[alive[a]]) { a = 1; b = 0; }
        for (int k = 2; k < na; k++) {
            if (w[alive[k]] < w[alive[a]]) { b = a; a = k; }
            else if (w[alive[k]] < w[alive[b]]) b = k;
        }
        w[nn] = w[alive[a]] + w[alive[b]]; parent[nn] = -1;
        parent[alive[a]] = nn; parent[alive[b]] = nn;
        int hi = a > b ? a : b, lo = a > b ? b : a;
        alive[lo] = nn; alive[hi] = alive[na - 1]; na--; nn++;
    }
    int over = 0;
    for (int i = 0; i < m; i++) { int d = 0, x = i; while (parent[x] >= 0) { x = parent[x]; d++; } if (d > maxbits) { d = maxbits; over = 1; } len[idx[i]] = (uint8_t)d; }
    if (over) {
        /* fix kraft: while sum > 1, lengthen the longest code shorter than maxbits with smallest freq */
        for (;;) {
            uint64_t k = 0; for (int i = 0; i < m; i++) k += 1ull << (maxbits - len[idx[i]]);
            if (k <= (1ull << maxbits)) break;
            int best = -1;
            for (int i = 0; i < m; i++) if (len[idx[i]] < maxbits && (best < 0 || len[idx[i]] > len[idx[best]] || (len[idx[i]] == len[idx[best]] && freq[idx[i]] < freq[idx[best]]))) best = i;
            len[idx[best]]++;
        }
    }
}
static const uint8_t lext[29] = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
static const uint8_t dext[30] = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};
static int lsym(int len) { int l = len - 3; if (l < 8) return l; if (len == 258) return 28; int msb = 31 - __builtin_clz(l); int eb = msb - 2; return 4 * (eb + 1) + ((l >> eb) & 3); }
static int dsym(int dist) { int d = dist - 1; if (d < 4) return d; int msb = 31 - __builtin_clz(d); int eb = msb - 1; return 2 * msb + ((d >> eb) & 1); }

static int g_hdrmode = 0; /* 0 = optimal RLE header, 1 = fixed 1338-bit header (v2/v3.1 kernel), 3 = run-length symbols with the kernel's STATIC code-length code */
static uint64_t g_cf[19]; /* code-length symbol counts over all blocks (printed with cf=1: the input for choosing the static code) */
static uint8_t g_static_cl[19] = {4, 6, 6, 5, 5, 4, 4, 3, 3, 3, 4, 4, 5, 5, 6, 6, 5, 5, 4};
static uint64_t block_cost(const uint32_t *fl, const uint32_t *fd, uint64_t *hdr_out) {
    uint8_t ll[288], dl[32];
    uint32_t f2[288]; memcpy(f2, fl, sizeof(f2)); f2[256] = 1;
    uint32_t d2[32]; memcpy(d2, fd, sizeof(d2));
    int nd = 0; for (int i = 0; i < 30; i++) nd += d2[i] != 0;
    if (nd == 0) { d2[0] = 1; d2[1] = 1; } else if (nd == 1) { if (d2[0]) d2[1] = 1; else d2[0] = 1; }
    huff_lengths(f2, 286, 15, ll); huff_lengths(d2, 30, 15, dl);
    uint64_t bits = 0;
    for (int i = 0; i < 286; i++) bits += (uint64_t)f2[i] * ll[i] + (i >= 257 ? (uint64_t)f2[i] * lext[i - 257] : 0);
    for (int i = 0; i < 30; i++) bits += (uint64_t)fd[i] * (dl[i] + dext[i]);
    uint64_t hdr;
    if (g_hdrmode == 1) hdr = 1338;
    else {
        int hlit = 286; while (hlit > 257 && ll[hlit - 1] == 0) hlit--;
        int hdist = 30; while (hdist > 1 && dl[hdist - 1] == 0) hdist--;
        uint8_t seq[320]; int n = 0;
        for (int i = 0; i < hlit; i++) seq[n++] = ll[i];
        for (int i = 0; i < hdist; i++) seq[n++] = dl[i];
        uint32_t cf[19] = {0}; uint64_t extra = 0;
        for (int i = 0; i < n;) {
            int j = i; while (j < n && seq[j] == seq[i]) j++;
            int run = j - i;
            if (seq[i] == 0) { while (run >= 11) { int r = run > 138 ? 138 : run; cf[18]++; extra += 7; run -= r; } if (run >= 3) { cf[17]++; extra += 3; run = 0; } cf[0] += run; }
            else { cf[seq[i]]++; run--; while (run >= 3) { int r = run > 6 ? 6 : run; cf[16]++; extra += 2; run -= r; } cf[seq[i]] += run; }
            i = j;
        }
        uint8_t cl[19];
        for (int i = 0; i < 19; i++) g_cf[i] += cf[i];
        if (g_hdrmode == 3) memcpy(cl, g_static_cl, 19);
        else huff_lengths(cf, 19, 7, cl);
        static const int ord[19] = {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15};
        int hclen = 19; while (hclen > 4 && cl[ord[hclen - 1]] == 0) hclen--;
        hdr = 17 + 3 * hclen + extra; for (int i = 0; i < 19; i++) hdr += (uint64_t)cf[i] * cl[i];
    }
    if (hdr_out) *hdr_out = hdr;
    return bits + ll[256] + hdr;
}

/* ---------------- parse models ---------------- */
typedef struct { int hashmode, hbits, hbytes, minmatch, span, lazy, batch, cap, sbsize, dict, toofar, ways, lstride, chunk; } Model;
static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint32_t hashv(uint32_t v, int hbytes, int bits) { if (hbytes == 3) v &= 0xffffff; return (v * 2654435761u) >> (32 - bits); }
static int mlen(const uint8_t *a, const uint8_t *b, int max) { int l = 0; while (l < max && a[l] == b[l]) l++; return l; }

static uint64_t g_tok_lit, g_tok_match, g_matchbytes, g_hdrbits;
/* compress one chunk [in, in+n) with optional history of `hist` bytes before in; returns bits */
static uint64_t chunk_bits(const uint8_t *in, int n, int hist, const Model *M) {
    int W = hist + n;                       /* window positions 0..W, chunk starts at hist */
    const uint8_t *w = in - hist;
    int *cand = malloc(sizeof(int) * (W + 8));
    int *cand2 = malloc(sizeof(int) * (W + 8));
    int HS = 1 << M->hbits;
    int *tab = malloc(sizeof(int) * HS * (M->ways > 1 ? M->ways : 1));
    for (int i = 0; i < W; i++) { cand[i] = -1; cand2[i] = -1; }
    int lastp = W - (M->hbytes);            /* last position with a full hash */
    if (M->hashmode == 0) {                 /* sequential nearest previous, all positions inserted, ways-deep bucket */
        int ways = M->ways > 1 ? M->ways : 1;
        for (int i = 0; i < HS * ways; i++) tab[i] = -1;
        for (int p = 0; p <= lastp; p++) {
            uint32_t h = hashv(rd32(w + p), M->hbytes, M->hbits);
            cand[p] = tab[h * ways];
            if (ways > 1) cand2[p] = tab[h * ways + 1];
            for (int k = ways - 1; k > 0; k--) tab[h * ways + k] = tab[h * ways + k - 1];
            tab[h * ways] = p;
        }
    } else if (M->hashmode == 1) {          /* first occurrence in window */
        for (int i = 0; i < HS; i++) tab[i] = -1;
        for (int p = 0; p <= lastp; p++) { uint32_t h = hashv(rd32(w + p), M->hbytes, M->hbits); if (tab[h] < 0) tab[h] = p; }
        for (int p = 0; p <= lastp; p++) { uint32_t h = hashv(rd32(w + p), M->hbytes, M->hbits); if (tab[h] < p) cand[p] = tab[h]; }
    } else if (M->hashmode == 2) {          /* batches: old = first occurrence in latest earlier batch holding the hash; new = first in own batch */
        for (int i = 0; i < HS; i++) tab[i] = -1;
        int B = M->batch;
        for (int b0 = 0; b0 <= lastp; b0 += B) {
            int b1 = b0 + B - 1 < lastp ? b0 + B - 1 : lastp;
            for (int p = b0; p <= b1; p++) { uint32_t h = hashv(rd32(w + p), M->hbytes, M->hbits); cand2[p] = tab[h]; }
            /* atomicMin of (~batch, pos): entry from this batch wins over older; within batch the smallest pos */
            for (int p = b0; p <= b1; p++) { uint32_t h = hashv(rd32(w + p), M->hbytes, M->hbits); if (tab[h] < b0) tab[h] = p; }
            for (int p = b0; p <= b1; p++) { uint32_t h = hashv(rd32(w + p), M->hbytes, M->hbits); if (tab[h] < p) cand[p] = tab[h]; else { cand[p] = cand2[p]; cand2[p] = -1; } }
        }
        if (M->ways < 2) for (int p = 0; p < W; p++) cand2[p] = -1;
    } else if (M->hashmode == 3) {          /* batches: lookup before insert, latest in batch wins (atomicMax); no intra-batch */
        for (int i = 0; i < HS; i++) tab[i] = -1;
        int B = M->batch;
        for (int b0 = 0; b0 <= lastp; b0 += B) {
            int b1 = b0 + B - 1 < lastp ? b0 + B - 1 : lastp;
            for (int p = b0; p <= b1; p++) { uint32_t h = hashv(rd32(w + p), M->hbytes, M->hbits); cand[p] = tab[h]; }
            for (int p = b0; p <= b1; p++) { uint32_t h = hashv(rd32(w + p), M->hbytes, M->hbits); tab[h] = p; }
        }
    }
    if (M->lstride > 1) for (int p = 0; p < W; p++) if (p % M->lstride) { cand[p] = -1; cand2[p] = -1; }
    /* per-position verified lengths (exact) */
    int *len = calloc(W + 8, sizeof(int)); int *dist = calloc(W + 8, sizeof(int));
    uint64_t bits = 0;
    for (int sb = hist; sb < W; sb += M->sbsize) {
        int sbe = sb + M->sbsize < W ? sb + M->sbsize : W;
        for (int p = sb; p < sbe; p++) {
            len[p] = 0; dist[p] = 0;
            for (int k = 0; k < 2; k++) {
                int c = k == 0 ? cand[p] : cand2[p];
                if (c < 0 || p - c > 32768) continue;
                int max = sbe - p < 258 ? sbe - p : 258;
                int l = mlen(w + p, w + c, max);
                if (l >= M->minmatch && !(l == 3 && p - c > M->toofar) && l > len[p]) { len[p] = l; dist[p] = p - c; }
            }
        }
        /* span-greedy with cover */
        uint32_t fl[288] = {0}, fd[32] = {0};
        int cover = sb;
        for (int s0 = sb; s0 < sbe; s0 += M->span) {
            int s1 = s0 + M->span < sbe ? s0 + M->span : sbe;
            int p = s0;
            while (p < s1) {
                int l = len[p], d = dist[p], lc = l > M->cap ? M->cap : l;
                if (M->lazy && l >= M->minmatch && p + 1 < s1) {
                    int l2 = len[p + 1] > M->cap ? M->cap : len[p + 1];
                    if (l2 > lc) l = 0;
                }
                int tstart = p, tlen = l >= M->minmatch ? l : 1;
                p += tlen;
                /* cover trimming */
                int tend = tstart + tlen;
                if (tend <= cover) continue;
                if (tlen == 1) { if (tstart >= cover) { fl[w[tstart]]++; g_tok_lit++; cover = tend; } continue; }
                if (tstart < cover) { int rem = tend - cover; if (rem >= 3) { fl[257 + lsym(rem)]++; fd[dsym(d)]++; g_tok_match++; g_matchbytes += rem; } else { for (int q = cover; q < tend; q++) { fl[w[q]]++; g_tok_lit++; } } cover = tend; continue; }
                fl[257 + lsym(tlen)]++; fd[dsym(d)]++; g_tok_match++; g_matchbytes += tlen; cover = tend;
            }
        }
        uint64_t hb; uint64_t c = block_cost(fl, fd, &hb); g_hdrbits += hb;
        uint64_t stored = (uint64_t)(sbe - sb) * 8 + 40;
        bits += c < stored ? c : stored;
    }
    free(cand); free(cand2); free(tab); free(len); free(dist);
    return bits + 3 + 7 + 32; /* sync marker approx */
}

static double zlib_ratio(const uint8_t *in, size_t n, int level, int chunk) {
    uint64_t out = 0; size_t cap = compressBound(chunk ? chunk : n) + 64; uint8_t *buf = malloc(cap);
    if (!chunk) chunk = (int)n;
    for (size_t o = 0; o < n; o += chunk) {
        size_t m = n - o < (size_t)chunk ? n - o : chunk;
        z_stream z; memset(&z, 0, sizeof z); deflateInit2(&z, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
        z.next_in = (uint8_t *)in + o; z.avail_in = m; z.next_out = buf; z.avail_out = cap; deflate(&z, Z_FINISH); out += z.total_out; deflateEnd(&z);
    }
    free(buf); return (double)out / n;
}

int main(int argc, char **argv) {
    size_t n = 32u << 20; const char *file = NULL; int zl = 0;
    Model M = {0, 14, 4, 4, 32, 0, 4096, 258, 32768, 0, 4096, 1, 1, 65536};
    for (int i = 1; i < argc; i++) {
        if (!strncmp(argv[i], "n=", 2)) n = (size_t)atol(argv[i] + 2) << 20;
        else if (!strncmp(argv[i], "file=", 5)) file = argv[i] + 5;
        else if (!strncmp(argv[i], "mode=", 5)) M.hashmode = atoi(argv[i] + 5);
        else if (!strncmp(argv[i], "hbits=", 6)) M.hbits = atoi(argv[i] + 6);
        else if (!strncmp(argv[i], "hbytes=", 7)) M.hbytes = atoi(argv[i] + 7);
        else if (!strncmp(argv[i], "min=", 4)) M.minmatch = atoi(argv[i] + 4);
        else if (!strncmp(argv[i], "span=", 5)) M.span = atoi(argv[i] + 5);
        else if (!strncmp(argv[i], "lazy=", 5)) M.lazy = atoi(argv[i] + 5);
        else if (!strncmp(argv[i], "batch=", 6)) M.batch = atoi(argv[i] + 6);
        else if (!strncmp(argv[i], "cap=", 4)) M.cap = atoi(argv[i] + 4);
        else if (!strncmp(argv[i], "sb=", 3)) M.sbsize = atoi(argv[i] + 3);
        else if (!strncmp(argv[i], "dict=", 5)) M.dict = atoi(argv[i] + 5);
        else if (!strncmp(argv[i], "toofar=", 7)) M.toofar = atoi(argv[i] + 7);
        else if (!strncmp(argv[i], "ways=", 5)) M.ways = atoi(argv[i] + 5);
        else if (!strncmp(argv[i], "lstride=", 8)) M.lstride = atoi(argv[i] + 8);
        else if (!strncmp(argv[i], "chunk=", 6)) M.chunk = atoi(argv[i] + 6);
        else if (!strncmp(argv[i], "hdr=", 4)) g_hdrmode = atoi(argv[i] + 4);
        else if (!strcmp(argv[i], "zlib")) zl = 1;
        else if (!strncmp(argv[i], "cl=", 3)) { const char *q = argv[i] + 3; for (int k = 0; k < 19 && *q; k++) { g_static_cl[k] = (uint8_t)strtol(q, (char **)&q, 10); if (*q == ',') q++; } }
    }
    uint8_t *buf;
    if (file) { FILE *f = fopen(file, "rb"); fseek(f, 0, SEEK_END); n = ftell(f); fseek(f, 0, SEEK_SET); buf = malloc(n + 64); if (fread(buf, 1, n, f) != n) return 1; fclose(f); }
    else { make_vocab(); buf = malloc(n + 64); textgen(buf, n, 1); }
    memset(buf + n, 0, 64);
    if (zl) {
        printf("zlib whole: L1 %.4f L6 %.4f L9 %.4f | 64K chunks: L1 %.4f L6 %.4f L9 %.4f\n", zlib_ratio(buf, n, 1, 0), zlib_ratio(buf, n, 6, 0), zlib_ratio(buf, n, 9, 0),
               zlib_ratio(buf, n, 1, 65536), zlib_ratio(buf, n, 6, 65536), zlib_ratio(buf, n, 9, 65536));
        return 0;
    }
    uint64_t bits = 0;
    for (size_t o = 0; o < n; o += M.chunk) {
        int m = (int)(n - o < (size_t)M.chunk ? n - o : (size_t)M.chunk);
        int hist = M.dict && o >= 32768 ? 32768 : 0;
        bits += chunk_bits(buf + o, m, hist, &M);
    }
    printf("lstride=%d chunk=%d ", M.lstride, M.chunk); printf("mode=%d hbits=%d hbytes=%d min=%d span=%d lazy=%d batch=%d cap=%d sb=%d dict=%d ways=%d hdr=%d : ratio %.4f  (lit/B %.3f match/B %.4f avgmatch %.2f hdr %.4f)\n", M.hashmode, M.hbits, M.hbytes, M.minmatch, M.span,
           M.lazy, M.batch, M.cap, M.sbsize, M.dict, M.ways, g_hdrmode, bits / 8.0 / n, (double)g_tok_lit / n, (double)g_tok_match / n, (double)g_matchbytes / (g_tok_match ? g_tok_match : 1), g_hdrbits / 8.0 / n);
    printf("cf:"); for (int i = 0; i < 19; i++) printf(" %llu", (unsigned long long)g_cf[i]); printf("\n");
    return 0;
}
