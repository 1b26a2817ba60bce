/* mz_strm_cuda.h -- drop-in DEFLATE codec stream + CRC-32 for minizip-ng, computed on a B200 (C ABI).
 *
 * This is the boundary. It replaces, one for one:
 *
 *   mz_stream_cuda_open            <- mz_stream_zlib_open            mz_strm_zlib.c:65-107   (mz_strm_zlib.h:20)
 *   mz_stream_cuda_is_open         <- mz_stream_zlib_is_open         mz_strm_zlib.c:109-114  (mz_strm_zlib.h:21)
 *   mz_stream_cuda_read            <- mz_stream_zlib_read            mz_strm_zlib.c:116-193  (mz_strm_zlib.h:22)
 *   mz_stream_cuda_write           <- mz_stream_zlib_write           mz_strm_zlib.c:243-264  (mz_strm_zlib.h:23)
 *   mz_stream_cuda_tell            <- mz_stream_zlib_tell            mz_strm_zlib.c:266-270  (mz_strm_zlib.h:24)
 *   mz_stream_cuda_seek            <- mz_stream_zlib_seek            mz_strm_zlib.c:272-278  (mz_strm_zlib.h:25)
 *   mz_stream_cuda_close           <- mz_stream_zlib_close           mz_strm_zlib.c:280-305  (mz_strm_zlib.h:26)
 *   mz_stream_cuda_error           <- mz_stream_zlib_error           mz_strm_zlib.c:307-310  (mz_strm_zlib.h:27)
 *   mz_stream_cuda_get_prop_int64  <- mz_stream_zlib_get_prop_int64  mz_strm_zlib.c:312-334  (mz_strm_zlib.h:29)
 *   mz_stream_cuda_set_prop_int64  <- mz_stream_zlib_set_prop_int64  mz_strm_zlib.c:336-355  (mz_strm_zlib.h:30)
 *   mz_stream_cuda_create          <- mz_stream_zlib_create          mz_strm_zlib.c:357-365  (mz_strm_zlib.h:32)
 *   mz_stream_cuda_delete          <- mz_stream_zlib_delete          mz_strm_zlib.c:367-374  (mz_strm_zlib.h:33)
 *   mz_stream_cuda_get_interface   <- mz_stream_zlib_get_interface   mz_strm_zlib.c:376-378  (mz_strm_zlib.h:35)
 *   mz_crypt_crc32_update          <- mz_crypt_crc32_update          mz_crypt.c:35-92        (mz_crypt.h:20)
 *
 * Same argument meaning, return values and error codes as the functions they replace (see the
 * conventions in SURVEY.md section 8b). The object returned by create() starts with the
 * reference's `mz_stream { vtbl, base }` header (mz_strm.h:69-72), so mz_stream_set_base() and the
 * generic mz_stream_* dispatchers (mz_strm.c:20-130, :375-379) work on it unchanged, and it can be
 * chained above any reference stream (mem / buffered / split / os / crypt).
 *
 * Compressed bytes differ from zlib's; every stream written is valid RFC1951 (raw, window_bits<0),
 * RFC1950 (8..15) or RFC1952 (24..31) and is inflated bit-exactly by the reference's own reader.
 * There is no CPU codec behind this interface: without a usable sm_100 GPU open() returns
 * MZ_SUPPORT_ERROR.
 */
#ifndef MZ_STREAM_CUDA_H
#define MZ_STREAM_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int32_t mz_stream_cuda_open(void *stream, const char *path, int32_t mode);
int32_t mz_stream_cuda_is_open(void *stream);
int32_t mz_stream_cuda_read(void *stream, void *buf, int32_t size);
int32_t mz_stream_cuda_write(void *stream, const void *buf, int32_t size);
int64_t mz_stream_cuda_tell(void *stream);
int32_t mz_stream_cuda_seek(void *stream, int64_t offset, int32_t origin);
int32_t mz_stream_cuda_close(void *stream);
int32_t mz_stream_cuda_error(void *stream);

int32_t mz_stream_cuda_get_prop_int64(void *stream, int32_t prop, int64_t *value);
int32_t mz_stream_cuda_set_prop_int64(void *stream, int32_t prop, int64_t value);

void *mz_stream_cuda_create(void);
void mz_stream_cuda_delete(void **stream);

void *mz_stream_cuda_get_interface(void);

/* Replacement for mz_crypt.c:35. Calls below MZ_CUDA_CRC_MIN_BYTES (default 1 MiB, environment
 * override) are answered on the host (PCLMULQDQ folding on x86-64, a table loop elsewhere): the reference calls this with 1-byte
 * (mz_strm_pkcrypt.c:79,86) and <=64 KiB (mz_zip.c:2049,2064) buffers for which a PCIe round trip is
 * absurd; anything larger goes to the GPU kernel. */
uint32_t mz_crypt_crc32_update(uint32_t value, const uint8_t *buf, int32_t size);

#ifdef __cplusplus
}
#endif
#endif
