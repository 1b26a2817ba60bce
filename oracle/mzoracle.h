/* mzoracle.h -- CPU oracle for the DEFLATE + CRC-32 hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This directory is the checker, never the product: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it. Nothing under
 * minizip-ng_b200/ links, imports or calls it.
 *
 * What it restates. minizip-ng itself holds no compression arithmetic: mz_strm_zlib.c is a
 * buffering shim over zlib-ng (third-party, NOT vendored under /root/reference: fetched at
 * configure time as branch `stable`, CMakeLists.txt:202) and mz_crypt.c:47 forwards CRC-32 to
 * it. The published algorithms behind those calls are RFC 1951 (DEFLATE), RFC 1952 (gzip
 * member framing), RFC 1950 (zlib framing) and CRC-32/ISO-HDLC; the only arithmetic present
 * in the reference tree is the byte-table CRC loop at mz_crypt.c:81-90. This file restates
 * those, and the call-site semantics of mz_strm_zlib.c (window_bits selection :87/:97,
 * consumed/produced accounting :139-182).
 *
 * Pinning. The restatement is pinned (tests/test_oracle.py) against
 *   - the reference's golden fixtures for the path (test/fuzz/unzip_fuzzer_seed_corpus/*.zip
 *     DEFLATE entries with header CRC + size; test/random.bin CRC a85d40dc), committed as
 *     tests/golden/ vectors by tests/golden/make_golden.py;
 *   - the reference itself run here: oracle/_ref/libmzref.so = the reference's own
 *     mz_strm_zlib.c / mz_crypt.c / mz_strm*.c compiled where they lie + system zlib 1.3
 *     (see oracle/Makefile). zlib-ng proper cannot be built offline; zlib 1.3 in
 *     ZLIB_COMPAT mode is the configuration minizip-ng documents (mz_strm_zlib.c:15-29).
 */
#ifndef MZORACLE_H
#define MZORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Framing selected exactly like zlib's windowBits at mz_strm_zlib.c:87/:97 */
#define ORC_WRAP_RAW  0 /* window_bits < 0   : bare RFC1951                    */
#define ORC_WRAP_ZLIB 1 /* window_bits 8..15 : RFC1950 2-byte header + Adler32 */
#define ORC_WRAP_GZIP 2 /* window_bits 24..31: RFC1952 header + CRC32 + ISIZE  */

/* Error codes mirror mz.h:21-26 (zlib-compatible) */
#define ORC_OK          0
#define ORC_DATA_ERROR (-3)
#define ORC_MEM_ERROR  (-4)
#define ORC_BUF_ERROR  (-5)

/* CRC-32 running update; same contract as mz_crypt_crc32_update (mz_crypt.c:35):
 * value starts at 0 and chains; pre/post inversion happens inside (mz_crypt.c:81,90). */
uint32_t orc_crc32_update(uint32_t value, const uint8_t *buf, size_t size);

/* crc(A||B) from crc(A), crc(B), len(B). Not present in the reference; oracle for the
 * multi-chunk combine (system zlib's crc32_combine is the independent cross-check). */
uint32_t orc_crc32_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b);

uint32_t orc_adler32_update(uint32_t value, const uint8_t *buf, size_t size);

/* One-shot decoder for a single member/stream.
 *   in/in_len    compressed bytes (may have trailing garbage, cf. zip over-read, mz_zip.c:2100)
 *   out/out_cap  destination; ORC_BUF_ERROR if the stream would produce more
 *   *consumed    compressed bytes used, including framing (TOTAL_IN semantics, mz_strm_zlib.c:168-175)
 *   *produced    bytes written
 * Returns ORC_OK at a clean end of stream, ORC_DATA_ERROR for invalid streams / trailer
 * mismatch, ORC_BUF_ERROR for truncated input or short output. */
int orc_inflate(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap, int wrap,
                size_t *consumed, size_t *produced);

/* Walk the stream and report each DEFLATE block: type (0,1,2), start bit, output bytes.
 * Used by tests to check block structure of product streams. Returns number of blocks or <0. */
typedef struct {
    uint64_t start_bit;
    uint64_t out_bytes;
    int32_t type;
    int32_t final;
} orc_block_info;
int64_t orc_inflate_blocks(const uint8_t *in, size_t in_len, int wrap, orc_block_info *blocks, size_t max_blocks,
                           size_t *produced);

/* Reference-free encoder restating RFC1951 with a greedy hash-chain parse and one dynamic
 * block per 64 KiB of tokens. level 0 = stored blocks. Used only as a "port" CPU baseline
 * and as a second producer of streams for the decoder tests. Returns bytes written or <0. */
int64_t orc_deflate(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap, int level, int wrap);

size_t orc_deflate_bound(size_t in_len);

#ifdef __cplusplus
}
#endif
#endif
