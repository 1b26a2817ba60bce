/* mz_abi.h -- the slice of minizip-ng's plug-in ABI that mz_strm_cuda is written against.
 *
 * The stream ABI is a struct layout plus integer constants; a plug-in cannot choose them, it has to
 * repeat them. When the real reference headers are on the include path (-DMZ_CUDA_USE_REFERENCE_HEADERS)
 * they are used instead and this file only checks nothing. Sources: vtbl slot order and the
 * {vtbl, base} object header -- mz_strm.h:53-72; property ids -- mz_strm.h:20-30; error codes, open
 * modes, seek origins, compression level sentinel -- mz.h:20-74.
 */
#ifndef MZ_ABI_H
#define MZ_ABI_H

#include <stdint.h>

#ifdef MZ_CUDA_USE_REFERENCE_HEADERS
#include "mz.h"
#include "mz_strm.h"
#else

/* mz.h:21-47 */
enum {
    MZ_OK = 0,
    MZ_STREAM_ERROR = -1,
    MZ_DATA_ERROR = -3,
    MZ_MEM_ERROR = -4,
    MZ_BUF_ERROR = -5,
    MZ_VERSION_ERROR = -6,
    MZ_END_OF_LIST = -100,
    MZ_END_OF_STREAM = -101,
    MZ_PARAM_ERROR = -102,
    MZ_FORMAT_ERROR = -103,
    MZ_INTERNAL_ERROR = -104,
    MZ_CRC_ERROR = -105,
    MZ_CRYPT_ERROR = -106,
    MZ_EXIST_ERROR = -107,
    MZ_PASSWORD_ERROR = -108,
    MZ_SUPPORT_ERROR = -109,
    MZ_HASH_ERROR = -110,
    MZ_OPEN_ERROR = -111,
    MZ_CLOSE_ERROR = -112,
    MZ_SEEK_ERROR = -113,
    MZ_TELL_ERROR = -114,
    MZ_READ_ERROR = -115,
    MZ_WRITE_ERROR = -116,
    MZ_SIGN_ERROR = -117,
    MZ_SYMLINK_ERROR = -118
};

/* mz.h:50-60 */
enum { MZ_OPEN_MODE_READ = 0x01, MZ_OPEN_MODE_WRITE = 0x02, MZ_OPEN_MODE_APPEND = 0x04, MZ_OPEN_MODE_CREATE = 0x08 };
enum { MZ_SEEK_SET = 0, MZ_SEEK_CUR = 1, MZ_SEEK_END = 2 };

/* mz.h:71-74 */
enum { MZ_COMPRESS_LEVEL_DEFAULT = -1 };

/* mz_strm.h:20-30 */
enum {
    MZ_STREAM_PROP_TOTAL_IN = 1,
    MZ_STREAM_PROP_TOTAL_IN_MAX = 2,
    MZ_STREAM_PROP_TOTAL_OUT = 3,
    MZ_STREAM_PROP_TOTAL_OUT_MAX = 4,
    MZ_STREAM_PROP_HEADER_SIZE = 5,
    MZ_STREAM_PROP_FOOTER_SIZE = 6,
    MZ_STREAM_PROP_DISK_SIZE = 7,
    MZ_STREAM_PROP_DISK_NUMBER = 8,
    MZ_STREAM_PROP_COMPRESS_LEVEL = 9,
    MZ_STREAM_PROP_COMPRESS_METHOD = 10,
    MZ_STREAM_PROP_COMPRESS_WINDOW = 11
};

/* mz_strm.h:53-72: twelve slots in this order, then the object header every stream starts with */
typedef struct mz_stream_vtbl_s {
    int32_t (*open)(void *stream, const char *path, int32_t mode);
    int32_t (*is_open)(void *stream);
    int32_t (*read)(void *stream, void *buf, int32_t size);
    int32_t (*write)(void *stream, const void *buf, int32_t size);
    int64_t (*tell)(void *stream);
    int32_t (*seek)(void *stream, int64_t offset, int32_t origin);
    int32_t (*close)(void *stream);
    int32_t (*error)(void *stream);
    void *(*create)(void);
    void (*destroy)(void **stream);
    int32_t (*get_prop_int64)(void *stream, int32_t prop, int64_t *value);
    int32_t (*set_prop_int64)(void *stream, int32_t prop, int64_t value);
} mz_stream_vtbl;

typedef struct mz_stream_s {
    mz_stream_vtbl *vtbl;
    struct mz_stream_s *base;
} mz_stream;

#endif /* MZ_CUDA_USE_REFERENCE_HEADERS */

/* Calls into the base stream the way mz_strm.c:34-41 / :101-110 dispatch them: null-check, is_open,
 * then the slot. Written out here because the plug-in must not depend on which library provides
 * mz_stream_read()/mz_stream_write(). */
static inline int32_t mz_abi_base_read(void *base, void *buf, int32_t size) {
    mz_stream *s = (mz_stream *)base;
    if (!s || !s->vtbl || !s->vtbl->read)
        return MZ_PARAM_ERROR;
    if (s->vtbl->is_open && s->vtbl->is_open(s) != MZ_OK)
        return MZ_STREAM_ERROR;
    return s->vtbl->read(s, buf, size);
}

static inline int32_t mz_abi_base_write(void *base, const void *buf, int32_t size) {
    mz_stream *s = (mz_stream *)base;
    if (size == 0)
        return size;
    if (!s || !s->vtbl || !s->vtbl->write)
        return MZ_PARAM_ERROR;
    if (s->vtbl->is_open && s->vtbl->is_open(s) != MZ_OK)
        return MZ_STREAM_ERROR;
    return s->vtbl->write(s, buf, size);
}

#endif
