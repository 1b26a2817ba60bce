/* textgen_dev.cu -- device side of the synthetic text generator: TEST / BENCH SUPPORT (libmztextgen.so), not product code.
 * Same bytes as textgen_host.c for the same (nbytes, seed). Tables are uploaded once per device. */
#include <cuda_runtime.h>
#include <mutex>

#include "textgen.h"

namespace {
struct Tables { uint8_t *words = nullptr; uint32_t *off = nullptr; uint32_t *zipf = nullptr; };
Tables g_t[16];
std::mutex g_mu;

__global__ void __launch_bounds__(256) mzt_kernel(uint8_t *out, uint64_t nbytes, uint64_t seed, const uint8_t *words, const uint32_t *off, const uint32_t *zipf) {
    const uint64_t npieces = (nbytes + MZT_TEXT_PIECE - 1) / MZT_TEXT_PIECE;
    for (uint64_t piece = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; piece < npieces; piece += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t o = piece * MZT_TEXT_PIECE;
        const uint32_t room = (uint32_t)(nbytes - o < MZT_TEXT_PIECE ? nbytes - o : MZT_TEXT_PIECE);
        mzt_piece(out + o, room, seed, piece, words, off, zipf);
    }
}
} // namespace

extern "C" int mzt_textgen_device(void *d_out, uint64_t nbytes, uint64_t seed, void *stream) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return -1;
    Tables &t = g_t[dev];
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (!t.words) {
            uint32_t nb = 0;
            const uint8_t *w = mzt_vocab_words(&nb);
            if (cudaMalloc(&t.words, nb) != cudaSuccess || cudaMalloc(&t.off, (MZT_NWORDS + 1) * 4) != cudaSuccess ||
                cudaMalloc(&t.zipf, (MZT_ZIPF_LUT + 1) * 4) != cudaSuccess)
                return -2;
            cudaMemcpy(t.words, w, nb, cudaMemcpyHostToDevice);
            cudaMemcpy(t.off, mzt_vocab_offsets(), (MZT_NWORDS + 1) * 4, cudaMemcpyHostToDevice);
            cudaMemcpy(t.zipf, mzt_zipf_lut(), (MZT_ZIPF_LUT + 1) * 4, cudaMemcpyHostToDevice);
        }
    }
    if (nbytes == 0) return 0;
    const uint64_t pieces = (nbytes + MZT_TEXT_PIECE - 1) / MZT_TEXT_PIECE;
    const uint64_t blocks = (pieces + 255) / 256;
    const unsigned grid = blocks < 148ull * 32 ? (unsigned)blocks : 148u * 32u;
    mzt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((uint8_t *)d_out, nbytes, seed, t.words, t.off, t.zipf);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
