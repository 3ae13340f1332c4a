// DEBUGGING AID, tests only: runs the device source of the solver core (pip_solver.h) on the
// CPU.  A warp is 32 cooperative fibers; a warp collective is a rendezvous of all 32.  The lane
// order of every scheduling round can be reversed or shuffled so that a missing __syncwarp in
// the device source shows up as a wrong answer here rather than as a heisenbug on the GPU.
// Never linked into the product library, never a fallback: the product has no CPU path.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef PIPEMU_TRACE
#include <execinfo.h>
#endif

namespace pipemu {

enum { NL = 32, STACK = 256 * 1024 };

static void *lane_sp[NL];
static void *sched_sp;
static char *stacks[NL];
static bool finished[NL];
static int cur = 0;
static int arrived = 0;
static unsigned gen = 0;
static long long xch[NL];
static int order_mode = 0;          // 0 forward, 1 reverse, 2 pseudo-random
static unsigned rng = 12345;
static void (*entry_fn)(void *, int);
static void *entry_arg;

extern "C" void pipemu_switch(void **save_sp, void *new_sp);
asm(R"(
.text
.globl pipemu_switch
.type pipemu_switch,@function
pipemu_switch:
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
)");

static void yield_to_sched() { pipemu_switch(&lane_sp[cur], sched_sp); }

static void trampoline()
{
  entry_fn(entry_arg, cur);
  finished[cur] = true;
  yield_to_sched();
  abort();
}

int lane() { return cur; }

static int tags[NL];
#ifdef PIPEMU_TRACE
static void *bt[NL][12];
static int btn[NL];
#endif
static void barrier_tag(int tag)
{
#ifdef PIPEMU_TRACE
  btn[cur] = backtrace(bt[cur], 12);
#endif
  // every lane must be at the same kind of collective: a mismatch is divergent control flow
  // around a collective in the device source (undefined behaviour on the GPU)
  tags[cur] = tag;
  unsigned g = gen;
  if (++arrived == NL) {
    for (int i = 0; i < NL; i++)
      if (tags[i] != tag) {
        fprintf(stderr, "pipemu: divergent collective: lane %d at kind %d, lane %d at kind %d\n", cur, tag, i, tags[i]);
#ifdef PIPEMU_TRACE
        fprintf(stderr, "--- lane %d\n", cur); backtrace_symbols_fd(bt[cur], btn[cur], 2);
        fprintf(stderr, "--- lane %d\n", i); backtrace_symbols_fd(bt[i], btn[i], 2);
#endif
        abort();
      }
    arrived = 0; gen++;
  }
  else while (gen == g) yield_to_sched();
}
void barrier() { barrier_tag(1); }

unsigned ballot(bool p)
{
  xch[cur] = p ? 1 : 0;
  barrier_tag(2);
  unsigned m = 0;
  for (int i = 0; i < NL; i++) if (xch[i]) m |= 1u << i;
  barrier_tag(2);
  return m;
}
long long shfl64(long long v, int src)
{
  xch[cur] = v;
  barrier_tag(3);
  long long r = xch[src & 31];
  barrier_tag(3);
  return r;
}
unsigned redmin(unsigned v)
{
  xch[cur] = v;
  barrier_tag(4);
  unsigned m = 0xffffffffu;
  for (int i = 0; i < NL; i++) if ((unsigned)xch[i] < m) m = (unsigned)xch[i];
  barrier_tag(4);
  return m;
}
unsigned redmax(unsigned v)
{
  xch[cur] = v;
  barrier_tag(5);
  unsigned m = 0;
  for (int i = 0; i < NL; i++) if ((unsigned)xch[i] > m) m = (unsigned)xch[i];
  barrier_tag(5);
  return m;
}
unsigned atomic_add(unsigned *p, unsigned v) { unsigned o = *p; *p = o + v; return o; }

void set_order(int mode) { order_mode = mode; }

// run fn(arg, lane) on 32 fibers to completion
void run_warp(void (*fn)(void *, int), void *arg)
{
  entry_fn = fn; entry_arg = arg;
  arrived = 0;
  for (int i = 0; i < NL; i++) {
    if (!stacks[i]) stacks[i] = (char *)aligned_alloc(64, STACK);
    finished[i] = false;
    // initial frame: six callee-saved registers then the return address
    uintptr_t top = ((uintptr_t)(stacks[i] + STACK)) & ~(uintptr_t)15;
    void **sp = (void **)(top - 8);          // keep (rsp+8) % 16 == 0 at function entry
    *--sp = (void *)trampoline;
    for (int r = 0; r < 6; r++) *--sp = 0;
    lane_sp[i] = sp;
  }
  int perm[NL];
  for (;;) {
    int live = 0;
    for (int i = 0; i < NL; i++) perm[i] = order_mode == 1 ? NL - 1 - i : i;
    if (order_mode == 2)
      for (int i = NL - 1; i > 0; i--) {
        rng = rng * 1664525u + 1013904223u;
        int j = (rng >> 8) % (i + 1);
        int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
      }
    for (int q = 0; q < NL; q++) {
      int i = perm[q];
      if (finished[i]) continue;
      live++;
      cur = i;
      pipemu_switch(&sched_sp, lane_sp[i]);
    }
    if (!live) break;
  }
}

}  // namespace pipemu
