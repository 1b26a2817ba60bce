/* cuda_emu.cc -- fiber scheduler behind cuda_emu.h (TEST INFRASTRUCTURE ONLY). */
#define MZ_EMU 1
#include "cuda_emu.h"

#include <pthread.h>
#include <sys/mman.h>

emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
uint8_t *emu_dyn_smem;

__asm__(
    ".text\n"
    ".globl emu_switch\n"
    ".type emu_switch,@function\n"
    "emu_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n");

static const size_t kStack = 256 * 1024;
static std::vector<emu_fiber> g_fibers;
static std::vector<emu_warp_state> g_warps;
static void *g_sched_sp;
static int g_cur = -1;
static unsigned g_nthreads, g_live;
static unsigned g_bar_count, g_bar_gen;
static unsigned long g_progress;
static void (*g_entry)(void *);
static void *g_arg;
static std::vector<uint8_t> g_dyn;

emu_warp_state *emu_cur_warp(void) { return &g_warps[g_cur >> 5]; }
int emu_lane(void) { return g_cur & 31; }

void emu_yield(void) {
    emu_fiber *f = &g_fibers[g_cur];
    emu_switch(&f->sp, g_sched_sp);
}

static void release_check_block(void) {
    if (g_live > 0 && g_bar_count >= g_live) {
        g_bar_count = 0;
        g_bar_gen++;
        g_progress++;
    }
}

void emu_block_barrier(void) {
    unsigned gen = g_bar_gen;
    g_bar_count++;
    release_check_block();
    while (g_bar_gen == gen) emu_yield();
}

static unsigned g_nb_count[16], g_nb_gen[16];
void emu_named_barrier(int id, int nthreads) {
    unsigned gen = g_nb_gen[id];
    g_nb_count[id]++;
    if (g_nb_count[id] >= (unsigned)nthreads) {
        g_nb_count[id] = 0;
        g_nb_gen[id]++;
        g_progress++;
        return;
    }
    while (g_nb_gen[id] == gen) emu_yield();
}

void emu_warp_barrier(unsigned mask) {
    emu_warp_state *w = emu_cur_warp();
    unsigned need = (unsigned)__builtin_popcount(mask & w->live_mask);
    unsigned gen = w->gen;
    w->arrived++;
    if ((unsigned)w->arrived >= need) {
        w->arrived = 0;
        w->gen++;
        g_progress++;
        return;
    }
    while (w->gen == gen) emu_yield();
}

static void trampoline(void) {
    g_entry(g_arg);
    emu_fiber *f = &g_fibers[g_cur];
    f->done = 1;
    g_live--;
    emu_warp_state *w = emu_cur_warp();
    w->live_mask &= ~(1u << emu_lane());
    g_progress++;
    /* an exiting thread may complete a barrier others wait on */
    release_check_block();
    if (w->arrived > 0 && (unsigned)w->arrived >= (unsigned)__builtin_popcount(w->live_mask)) {
        w->arrived = 0;
        w->gen++;
    }
    emu_switch(&f->sp, g_sched_sp);
    abort();
}

static pthread_mutex_t g_launch_mu = PTHREAD_MUTEX_INITIALIZER;
void emu_launch_lock(void) { pthread_mutex_lock(&g_launch_mu); }
void emu_launch_unlock(void) { pthread_mutex_unlock(&g_launch_mu); }

void emu_run_block(unsigned nthreads, void (*entry)(void *), void *arg, size_t dyn_smem) {
    if (g_fibers.size() < nthreads) {
        size_t old = g_fibers.size();
        g_fibers.resize(nthreads);
        for (size_t i = old; i < nthreads; i++) {
            g_fibers[i].stack = mmap(NULL, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (g_fibers[i].stack == MAP_FAILED) { perror("mmap"); abort(); }
        }
    }
    g_warps.assign((nthreads + 31) / 32, emu_warp_state());
    g_dyn.assign(dyn_smem + 64, 0xCD); /* garbage like the device */
    emu_dyn_smem = (uint8_t *)(((uintptr_t)g_dyn.data() + 15) & ~(uintptr_t)15);
    g_entry = entry;
    g_arg = arg;
    g_nthreads = g_live = nthreads;
    g_bar_count = 0;
    for (int i = 0; i < 16; i++) g_nb_count[i] = 0;
    for (unsigned i = 0; i < nthreads; i++) {
        emu_fiber *f = &g_fibers[i];
        f->done = 0;
        f->tid = (int)i;
        uintptr_t top = ((uintptr_t)f->stack + kStack) & ~(uintptr_t)15;
        void **sp = (void **)top;
        *--sp = NULL;                 /* keeps rsp%16==8 at trampoline entry */
        *--sp = (void *)trampoline;   /* return address for emu_switch's ret */
        for (int k = 0; k < 6; k++) *--sp = NULL; /* r15 r14 r13 r12 rbx rbp */
        f->sp = sp;
        g_warps[i >> 5].live_mask |= 1u << (i & 31);
    }
    while (g_live > 0) {
        unsigned long before = g_progress;
        for (unsigned i = 0; i < nthreads; i++) {
            if (g_fibers[i].done) continue;
            g_cur = (int)i;
            threadIdx = emu_dim3(i % blockDim.x, (i / blockDim.x) % blockDim.y, i / (blockDim.x * blockDim.y));
            emu_switch(&g_sched_sp, g_fibers[i].sp);
        }
        if (g_progress == before && g_live > 0) {
            fprintf(stderr, "cuda_emu: deadlock in block (%u,%u,%u): %u threads stuck (barrier count %u)\n", blockIdx.x,
                    blockIdx.y, blockIdx.z, g_live, g_bar_count);
            abort();
        }
    }
    g_cur = -1;
}
