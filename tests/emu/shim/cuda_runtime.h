/* cuda_runtime.h -- the few CUDA runtime entry points csrc/mz_cuda_api.cu uses, implemented on the host so that the
 * REAL API shim (and with it the real host-side stream code, mz_strm_cuda.c / mz_zip_cuda.c) can be linked against the
 * CPU execution-model emulator: tests/emu/libmz_strm_emu.so. "Device" memory is host memory, streams are synchronous,
 * events are timestamps. TEST INFRASTRUCTURE ONLY -- never part of the product library. */
#ifndef MZ_EMU_CUDA_RUNTIME_SHIM_H
#define MZ_EMU_CUDA_RUNTIME_SHIM_H

#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorNotReady = 600 };
typedef struct emu_stream_s *cudaStream_t;
typedef struct emu_event_s { double t; } *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes { enum cudaMemoryType type; };
struct cudaDeviceProp { int major, minor, multiProcessorCount; };

static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : (e == cudaErrorMemoryAllocation ? "out of memory" : "emulated error"); }
static inline cudaError_t cudaGetLastError(void) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(struct cudaDeviceProp *p, int dev) {
    (void)dev;
    p->major = 10; p->minor = 0;
    p->multiProcessorCount = 2; /* keeps emulated grids small */
    return cudaSuccess;
}
/* page-aligned like cudaMalloc (>= 256-byte alignment is relied upon) */
static inline cudaError_t cudaMalloc(void **p, size_t n) { return posix_memalign(p, 4096, n ? n : 16) == 0 ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned flags) { (void)flags; return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, enum cudaMemcpyKind k) { (void)k; memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, enum cudaMemcpyKind k, cudaStream_t st) { (void)k; (void)st; memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t st) { (void)st; memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags) { (void)flags; *s = (cudaStream_t)malloc(8); return *s ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t s) { (void)s; return cudaSuccess; }
static inline double emu_now_ms(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t)malloc(sizeof(**e)); if (!*e) return cudaErrorMemoryAllocation; (*e)->t = 0; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) { (void)s; e->t = emu_now_ms(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t e) { (void)e; return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t e) { (void)e; return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F f, enum cudaFuncAttribute a, int v) { (void)f; (void)a; (void)v; return cudaSuccess; }
static inline cudaError_t cudaPointerGetAttributes(struct cudaPointerAttributes *a, const void *p) { (void)p; a->type = cudaMemoryTypeUnregistered; return cudaSuccess; }

/* multi-device / IPC entry points: the emulator has one "device" whose memory is the host's, so several shards can be run on
 * ordinal 0 (tests exercise the piece pipeline, the region layout and the CRC fold of mz_cuda_deflate_sharded that way) */
enum { cudaErrorPeerAccessAlreadyEnabled = 704, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1 };
typedef struct { char reserved[64]; } cudaIpcMemHandle_t;
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { memset(h, 0, sizeof(*h)); memcpy(h->reserved, &p, sizeof(p)); return cudaSuccess; }
static inline cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned flags) { (void)flags; memcpy(p, h.reserved, sizeof(*p)); return cudaSuccess; }
static inline cudaError_t cudaIpcCloseMemHandle(void *p) { (void)p; return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int peer, unsigned flags) { (void)peer; (void)flags; return cudaSuccess; }
static inline cudaError_t cudaMemcpyPeerAsync(void *d, int dd, const void *s, int sd, size_t n, cudaStream_t st) { (void)dd; (void)sd; (void)st; memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags) { (void)s; (void)e; (void)flags; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags) { (void)flags; return cudaEventCreate(e); }

#endif
