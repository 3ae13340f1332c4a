// DEBUGGING AID, tests only (see emu_runtime.cpp): the device source of the solver compiled
// for the host with -DPIP_EMU and driven one warp at a time.
#include <stdlib.h>
#include <string.h>

#include "../../piplib_b200/csrc/pip_warp_main.h"

namespace pipemu {
void run_warp(void (*fn)(void *, int), void *arg);
void set_order(int mode);
}

struct EmuArgs { PipLaunch L; pip_i64 *arena; };

static void warp_entry(void *a, int) { EmuArgs *e = (EmuArgs *)a; pip_warp_main(e->L, 0, e->arena); }

extern "C" int pipemu_solve_batch(const PipProblem *prob, int nprob, const pip_i64 *pool, PipResult *res,
                                  PipCell *cells, long long cells_cap, int work_words,
                                  long long stack_words, int slack_level, int order_mode,
                                  int sol_size, int maxcol)
{
  EmuArgs e;
  unsigned queue[2] = {0, 0};
  memset(&e, 0, sizeof e);
  e.L.prob = prob; e.L.pool = pool; e.L.pool_elem_log2 = 3; e.L.order = 0; e.L.nprob = nprob; e.L.res = res;
  e.L.cells = cells; e.L.cells_per_warp = cells_cap;
  e.L.stack = (pip_i64 *)malloc(sizeof(pip_i64) * stack_words);
  e.L.stack_words_per_warp = stack_words;
  e.L.gwork = 0; e.L.work_words = work_words; e.L.queue = queue;
  e.L.sol_size = sol_size > 0 ? sol_size : PIP_SOL_SIZE;
  e.L.maxcol = maxcol > 0 ? maxcol : PIP_MAXCOL;
  e.L.maxparm = PIP_MAXPARM;
  e.L.slack_level = slack_level;
  e.arena = (pip_i64 *)malloc(sizeof(pip_i64) * (size_t)work_words);
  memset(e.arena, 0x5a, sizeof(pip_i64) * (size_t)work_words);   // poison: nothing may rely on zeros
  for (int i = 0; i < nprob; i++) res[i].status = PIP_ST_PENDING;
  pipemu::set_order(order_mode);
  pipemu::run_warp(warp_entry, &e);
  free(e.arena);
  free(e.L.stack);
  return 0;
}
