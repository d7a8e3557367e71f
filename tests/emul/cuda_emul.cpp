// cuda_emul.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emul.h).  Fiber scheduler of the SIMT emulator.
#include "cuda_emul.h"

#include <time.h>
#include <mutex>
#include <vector>

namespace emul {

Idx g_threadIdx, g_blockIdx;
dim3 g_blockDim, g_gridDim;
unsigned char* g_dyn_smem = nullptr;

enum State { RUNNABLE, WAIT_CTA, WAIT_WARP, DONE };

// Minimal x86-64 SysV context switch (callee-saved registers + stack pointer); ~20x cheaper than swapcontext,
// which makes two sigprocmask system calls per switch.
extern "C" void emul_switch(void** from_sp, void* to_sp);
asm(R"(
.text
.globl emul_switch
.type emul_switch,@function
emul_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emul_switch,.-emul_switch
)");

struct Fiber {
    void* sp;
    State st;
    Idx tid;
    int warp;
};

static const size_t STACK = 128 * 1024;
static std::vector<Fiber> g_fibers;
static std::vector<unsigned char> g_stacks;
static void* g_sched_sp = nullptr;
static int g_cur = -1;
static const std::function<void()>* g_body = nullptr;
static int g_cta_arrived = 0, g_cta_live = 0;
static std::vector<int> g_warp_arrived, g_warp_live;
static std::vector<uint64_t> g_xchg;      // per thread exchange slot
int g_vote_slot[3] = {0, 0, 0};
static std::vector<int> g_vote_calls;
int vote_enter() { return g_vote_calls[g_cur]++; }

static void yield_to_sched() { emul_switch(&g_fibers[g_cur].sp, g_sched_sp); }

static void fiber_main() {
    (*g_body)();
    Fiber& f = g_fibers[g_cur];
    f.st = DONE;
    g_cta_live--;
    g_warp_live[f.warp]--;
    emul_switch(&f.sp, g_sched_sp);
    abort();     // a finished fiber is never resumed
}

void sync_cta() {
    g_fibers[g_cur].st = WAIT_CTA;
    g_cta_arrived++;
    yield_to_sched();
}

void sync_warp() {
    Fiber& f = g_fibers[g_cur];
    f.st = WAIT_WARP;
    g_warp_arrived[f.warp]++;
    yield_to_sched();
}

uint64_t warp_exchange(uint64_t v, int src_lane) {
    int base = g_fibers[g_cur].warp * 32;
    g_xchg[g_cur] = v;
    sync_warp();
    int src = base + src_lane;
    uint64_t r = (src < (int)g_fibers.size()) ? g_xchg[src] : v;
    sync_warp();
    return r;
}

uint32_t warp_ballot(int pred) {
    int w = g_fibers[g_cur].warp, base = w * 32;
    g_xchg[g_cur] = pred ? 1 : 0;
    // lanes that already exited must not contribute stale values
    sync_warp();
    uint32_t r = 0;
    for (int l = 0; l < 32; l++) {
        int t = base + l;
        if (t < (int)g_fibers.size() && g_fibers[t].st != DONE && g_xchg[t]) r |= 1u << l;
    }
    sync_warp();
    return r;
}

static void release_barriers() {
    if (g_cta_live > 0 && g_cta_arrived == g_cta_live) {
        g_cta_arrived = 0;
        for (auto& f : g_fibers) if (f.st == WAIT_CTA) f.st = RUNNABLE;
    }
    for (size_t w = 0; w < g_warp_live.size(); w++) {
        if (g_warp_live[w] > 0 && g_warp_arrived[w] == g_warp_live[w]) {
            g_warp_arrived[w] = 0;
            for (size_t t = w * 32; t < (w + 1) * 32 && t < g_fibers.size(); t++)
                if (g_fibers[t].st == WAIT_WARP) g_fibers[t].st = RUNNABLE;
        }
    }
}

static void run_cta(dim3 block) {
    size_t n = (size_t)block.x * block.y * block.z;
    g_fibers.resize(n);
    if (g_stacks.size() < n * STACK) g_stacks.resize(n * STACK);
    g_xchg.assign(n, 0);
    g_vote_calls.assign(n, 0);
    g_vote_slot[0] = g_vote_slot[1] = g_vote_slot[2] = 0;
    size_t nw = (n + 31) / 32;
    g_warp_arrived.assign(nw, 0);
    g_warp_live.assign(nw, 0);
    g_cta_arrived = 0;
    g_cta_live = (int)n;
    for (size_t t = 0; t < n; t++) {
        Fiber& f = g_fibers[t];
        f.st = RUNNABLE;
        f.tid.x = (unsigned)(t % block.x);
        f.tid.y = (unsigned)((t / block.x) % block.y);
        f.tid.z = (unsigned)(t / ((size_t)block.x * block.y));
        f.warp = (int)(t / 32);
        g_warp_live[f.warp]++;
        uintptr_t top = ((uintptr_t)(g_stacks.data() + (t + 1) * STACK)) & ~(uintptr_t)15;
        void** sp = (void**)top;
        *--sp = nullptr;                       // alignment slot: rsp == 8 (mod 16) when fiber_main starts
        *--sp = (void*)fiber_main;             // "return address" of the first switch
        for (int k = 0; k < 6; k++) *--sp = nullptr;   // rbp rbx r12 r13 r14 r15
        f.sp = (void*)sp;
    }
    size_t done = 0;
    while (done < n) {
        bool progressed = false;
        for (size_t t = 0; t < n; t++) {
            Fiber& f = g_fibers[t];
            if (f.st != RUNNABLE) continue;
            g_cur = (int)t;
            g_threadIdx = f.tid;
            emul_switch(&g_sched_sp, f.sp);
            progressed = true;
            if (f.st == DONE) done++;
            release_barriers();
        }
        if (!progressed) {
            release_barriers();
            bool any = false;
            for (auto& f : g_fibers) if (f.st == RUNNABLE) any = true;
            if (!any && done < n) {
                fprintf(stderr, "cuda_emul: DEADLOCK in CTA (%u,%u): %zu/%zu threads done, cta barrier %d/%d\n",
                        g_blockIdx.x, g_blockIdx.y, done, n, g_cta_arrived, g_cta_live);
                abort();
            }
        }
    }
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    static std::mutex mu;                       // the emulator state is global: one kernel at a time
    std::lock_guard<std::mutex> lock(mu);
    static std::vector<unsigned char> dyn;
    if (dyn.size() < smem + 16) dyn.resize(smem + 16);
    g_dyn_smem = dyn.data();
    g_body = &body;
    g_gridDim = grid;
    g_blockDim = block;
    for (unsigned z = 0; z < grid.z; z++)
        for (unsigned y = 0; y < grid.y; y++)
            for (unsigned x = 0; x < grid.x; x++) {
                g_blockIdx.x = x; g_blockIdx.y = y; g_blockIdx.z = z;
                run_cta(block);
            }
    g_body = nullptr;
}

}  // namespace emul

static double now_ms() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec * 1e-6; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emul_event{0}; return 0; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = now_ms(); return 0; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return 0; }

// CTA-wide vote barriers.  Each fiber counts its own vote calls; call k accumulates in slot k%3 and clears slot
// (k+1)%3, which no fiber can still be using (barriers keep fibers within one call of each other).
int __syncthreads_count(int pred) {
    int k = emul::vote_enter();
    emul::g_vote_slot[(k + 1) % 3] = 0;
    if (pred) emul::g_vote_slot[k % 3] += 1;
    emul::sync_cta();
    return emul::g_vote_slot[k % 3];
}
int __syncthreads_or(int pred) { return __syncthreads_count(pred) != 0; }
int __syncthreads_and(int pred) { return __syncthreads_count(!pred) == 0; }
