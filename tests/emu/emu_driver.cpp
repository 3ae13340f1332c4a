// DEBUGGING AID, tests only (see emu_runtime.cpp): the device source of the solver compiled
// for the host with -DPIP_EMU and driven one warp at a time.
#include <stdlib.h>
#include <string.h>

#include "../../piplib_b200/csrc/pip_warp_main.h"

namespace pipemu {
void run_warp(void (*fn)(void *, int), void *arg);
void set_order(int mode);
}

struct EmuArgs { PipLaunch L; pip_i64 *arena; int narrow; int steal; };

static void warp_entry(void *a, int)
{
  EmuArgs *e = (EmuArgs *)a;
  if (e->steal && e->narrow) pip_warp_main<int, false, true>(e->L, 0, e->arena);
  else if (e->steal) pip_warp_main<pip_i64, false, true>(e->L, 0, e->arena);
  else if (e->narrow == 2) pip_warp_main<pip_i64, true>(e->L, 0, e->arena, nullptr);   // the global-memory code path (classes G / M)
  else if (e->narrow) pip_warp_main<int>(e->L, 0, e->arena);
  else pip_warp_main<pip_i64>(e->L, 0, e->arena);
}

extern "C" int pipemu_solve_batch(const PipProblem *prob, int nprob, const pip_i64 *pool, PipResult *res,
                                  PipCell *cells, long long cells_cap, int work_words,
                                  long long stack_words, int slack_level, int order_mode,
                                  int sol_size, int maxcol, int narrow, int emit_words, unsigned long long *hash_out)
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
  e.L.emit_words = emit_words;
  (void)hash_out;
  e.narrow = narrow;
  e.arena = (pip_i64 *)malloc(sizeof(pip_i64) * (size_t)work_words);
  memset(e.arena, 0x5a, sizeof(pip_i64) * (size_t)work_words);   // poison: nothing may rely on zeros
  for (int i = 0; i < nprob; i++) res[i].status = PIP_ST_PENDING;
  pipemu::set_order(order_mode);
  pipemu::run_warp(warp_entry, &e);
  free(e.arena);
  free(e.L.stack);
  return 0;
}

// ---- a uniform (dense) batch as the engine runs it: arena layout carved once, arena images built ahead of the
// solve by the solver's own loader, word mode
template <class V> struct EmuImage { const PipProblem *prob; const pip_i64 *pool; int n; PipLayout L; pip_i64 *images; int words; };
template <class V> static void image_entry(void *a, int)
{
  EmuImage<V> *e = (EmuImage<V> *)a;
  for (int p = 0; p < e->n; p++) {
    PipLayout L = e->L;
    L.m.ni = e->prob[p].ni;
    PipSolver<V>::pip_load_problem(e->prob[p], e->pool, 3, e->images + (size_t)p * e->words, L.m, L.ctx, L.cstride);
  }
}
extern "C" int pipemu_solve_uniform(const PipProblem *prob, int nprob, const pip_i64 *pool, PipResult *res,
                                    PipCell *cells, long long cells_cap, int work_words, long long stack_words,
                                    int slack_level, int order_mode, int narrow, int max_ni, int max_nc)
{
  EmuArgs e;
  unsigned queue[2] = {0, 0};
  memset(&e, 0, sizeof e);
  e.L.prob = prob; e.L.pool = pool; e.L.pool_elem_log2 = 3; e.L.order = 0; e.L.nprob = nprob; e.L.res = res;
  e.L.cells = cells; e.L.cells_per_warp = cells_cap;
  e.L.stack = (pip_i64 *)malloc(sizeof(pip_i64) * stack_words);
  e.L.stack_words_per_warp = stack_words;
  e.L.gwork = 0; e.L.work_words = work_words; e.L.queue = queue;
  e.L.sol_size = PIP_SOL_SIZE; e.L.maxcol = PIP_MAXCOL; e.L.maxparm = PIP_MAXPARM;
  e.L.slack_level = slack_level;
  e.L.emit_words = 1;
  const int vb = narrow ? 4 : 8;
  int level = slack_level;
  while (!pip_layout(prob[0].nvar, prob[0].nparm, max_ni, max_nc, prob[0].flags, level, work_words, vb, e.L.layout)) level--;
  e.L.layout.s.ni = max_nc + 1;
  e.L.have_layout = 1;
  const PipLayout &Y = e.L.layout;
  const int w1 = (Y.m.data - Y.m.den) + (int)(((long long)max_ni * Y.m.stride * vb + 7) / 8);
  const int w2 = (int)(((long long)max_nc * Y.cstride * vb + 7) / 8);
  pip_i64 *images = (pip_i64 *)malloc(sizeof(pip_i64) * (size_t)nprob * (w1 + w2) + 64);
  memset(images, 0x6b, sizeof(pip_i64) * (size_t)nprob * (w1 + w2));
  PipLayout R = Y;
  R.m.fl -= R.m.den; R.tmp -= R.m.den; R.m.data -= R.m.den; R.m.den = 0; R.ctx = w1;
  pipemu::set_order(order_mode);
  if (narrow) { EmuImage<int> im = {prob, pool, nprob, R, images, w1 + w2}; pipemu::run_warp(image_entry<int>, &im); }
  else { EmuImage<pip_i64> im = {prob, pool, nprob, R, images, w1 + w2}; pipemu::run_warp(image_entry<pip_i64>, &im); }
  e.L.images = images; e.L.image_words = w1 + w2; e.L.image_w1 = w1;
  e.narrow = narrow;
  e.arena = (pip_i64 *)malloc(sizeof(pip_i64) * (size_t)work_words);
  memset(e.arena, 0x5a, sizeof(pip_i64) * (size_t)work_words);
  for (int i = 0; i < nprob; i++) res[i].status = PIP_ST_PENDING;
  pipemu::run_warp(warp_entry, &e);
  free(e.arena); free(e.L.stack); free(images);
  return 0;
}

// ---- subtree donation in emulation (PipSteal mode 2): one warp offers the ELSE branch of every outermost open
// split, finishes its own part, then solves the offered subtrees itself as separate segments; the segment walk
// of the copy kernel (pip_segments.h) resolves the verdict and the words are concatenated in pre-order
#include "../../piplib_b200/csrc/pip_segments.h"
extern "C" int pipemu_solve_batch_steal(const PipProblem *prob, int nprob, const pip_i64 *pool, PipResult *res,
                                        PipCell *cells, long long cells_cap, int work_words, long long stack_words,
                                        int slack_level, int order_mode, int sol_size, int narrow,
                                        long long *words_out, long long words_cap, long long *words_off, int *nsegs,
                                        unsigned budget, int *handed, unsigned handed_max)
{
  EmuArgs e;
  unsigned queue[4] = {0, 0, 0, 0};
  memset(&e, 0, sizeof e);
  e.L.prob = prob; e.L.pool = pool; e.L.pool_elem_log2 = 3; e.L.order = 0; e.L.nprob = nprob; e.L.res = res;
  e.L.cells = cells; e.L.cells_per_warp = cells_cap;
  e.L.stack = (pip_i64 *)malloc(sizeof(pip_i64) * stack_words);
  e.L.stack_words_per_warp = stack_words;
  e.L.gwork = 0; e.L.work_words = work_words; e.L.queue = queue;
  e.L.sol_size = sol_size > 0 ? sol_size : PIP_SOL_SIZE;
  e.L.maxcol = PIP_MAXCOL; e.L.maxparm = PIP_MAXPARM;
  e.L.slack_level = slack_level;
  e.L.emit_words = 1;
  PipSteal &S = e.L.steal;
  e.steal = 1;
  S.mode = 2; S.cap = 1 << 16;
  S.offers = (PipOffer *)calloc(S.cap, sizeof(PipOffer));
  unsigned ctl[PIP_STL_NCTL] = {0};
  ctl[PIP_STL_TOTAL] = 1;
  S.ctl = ctl;
  S.segs = (PipResult *)calloc(S.cap, sizeof(PipResult));
  S.seg_next = (int *)malloc(sizeof(int) * S.cap);
  S.seg_hwm = (int *)calloc(S.cap, sizeof(int));
  S.head_next = (int *)malloc(sizeof(int) * (nprob + 1));
  S.head_hwm = (int *)calloc(nprob + 1, sizeof(int));
  for (int i = 0; i < S.cap; i++) S.seg_next[i] = -1;
  for (int i = 0; i < nprob; i++) S.head_next[i] = -1;
  e.narrow = narrow;
  e.arena = (pip_i64 *)malloc(sizeof(pip_i64) * (size_t)work_words);
  memset(e.arena, 0x5a, sizeof(pip_i64) * (size_t)work_words);
  for (int i = 0; i < nprob; i++) res[i].status = PIP_ST_PENDING;
  pipemu::set_order(order_mode);
  /* a frame stack per problem would be the GPU's (one per warp, reused): here the single warp's stack is reused
   * by the next problem while offers still point into it, so every problem is run with its offers drained:
   * the queue hands out one problem at a time */
  long long at = 0;
  int *heavy = (int *)malloc(sizeof(int) * (nprob + 1));
  if (budget) {
    /* heavy-problem hand-over as the engine runs it: a first launch without donation in which a problem past
     * `budget` pivots stops at its next split and is listed, then the donation launch over the list, writing
     * behind the first launch's windows */
    PipLaunch first = e.L;
    e.L.steal.mode = 0; e.steal = 0;
    e.L.budget = budget; e.L.heavy = heavy; e.L.heavy_max = handed_max; e.L.cells_per_warp = cells_cap / 2;
    pipemu::run_warp(warp_entry, &e);
    *handed = (int)queue[2];
    e.L = first; e.steal = 1;
    e.L.heavy = heavy; e.L.from_heavy = 1; e.L.heavy_warps = 8; e.L.cell_base = cells_cap / 2; e.L.cells_per_warp = cells_cap / 2;
  }
  const unsigned nheavy = queue[2];
  for (int i = 0; i < nprob; i++) {
    if (!budget) {
      queue[0] = (unsigned)i;
      e.L.nprob = i + 1;
      pipemu::run_warp(warp_entry, &e);
    } else {
      /* (every run of the emulated warp starts its window afresh: a listed problem is solved right before
       * its words are collected) */
      for (unsigned j = 0; j < nheavy; j++)
        if (heavy[j] == i) { queue[2] = j + 1; queue[3] = j; pipemu::run_warp(warp_entry, &e); }
    }
    PipResolved R;
    pip_resolve_segments(res[i], S.head_next[i], S, e.L.sol_size, R);
    words_off[i] = at;
    nsegs[i] = R.nseg;
    res[i].status = R.status;
    res[i].pivots = (unsigned)R.pivots; res[i].cuts = (unsigned)R.cuts; res[i].subsolves = (unsigned)R.subsolves;
    res[i].splits = (unsigned)R.splits; res[i].ncells = (int)R.cells;
    if (R.status == PIP_ST_OK || R.status == PIP_ST_VOID) {
      int seg = -1;
      for (;;) {
        const PipResult sr = seg < 0 ? res[i] : S.segs[seg];
        const long long n = seg < 0 && S.head_next[i] >= 0 ? (long long)sr.ser_words : (long long)sr.ser_words;
        const void *src = (const void *)(cells + sr.cell_off);
        for (long long k = 0; k < n && at < words_cap; k++) words_out[at++] = narrow == 1 ? (long long)((const int *)src)[k] : ((const long long *)src)[k];
        seg = seg < 0 ? S.head_next[i] : S.seg_next[seg];
        if (seg < 0) break;
      }
    }
  }
  words_off[nprob] = at;
  free(heavy);
  free(e.arena); free(e.L.stack); free(S.offers); free(S.segs); free(S.seg_next); free(S.seg_hwm); free(S.head_next); free(S.head_hwm);
  return 0;
}

// ---- large-tableau solver in emulation: one CTA of one warp ---------------------------------
#include "../../piplib_b200/csrc/pip_large.h"

struct EmuLarge { PipLarge L; long long align_; int red[PIPL_RED_INTS]; pip_i64 *stage; };
static void large_entry(void *a, int) { EmuLarge *e = (EmuLarge *)a; pipl_solve(e->L, e->red, e->stage); }

static int g_large_staged = 0;
extern "C" void pipemu_large_staged(int on) { g_large_staged = on; }

extern "C" int pipemu_solve_large(int nvar, int ni, int nq, const pip_i64 *tab, int cut_rows, int sol_size,
                                  int maxcol, int *status, PipCell *cells, int *ncells, long long *info,
                                  int order_mode)
{
  EmuLarge e;
  PipLarge &L = e.L;
  memset(&e, 0, sizeof e);
  const int ncol = nvar + 1;
  L.nvar = nvar; L.ni = ni; L.flags = nq ? PIP_F_INT : 0;
  L.stride = (ncol + 1) & ~1;
  L.rcap = ni + cut_rows; L.pcap = nvar + L.rcap;
  L.sol_size = sol_size > 0 ? sol_size : PIP_SOL_SIZE;
  L.maxcol = maxcol > 0 ? maxcol : PIP_MAXCOL;
  L.data = (pip_i64 *)calloc((size_t)L.rcap * L.stride + 2, 8);
  for (int r = 0; r < ni; r++) memcpy(L.data + (size_t)r * L.stride, tab + (size_t)r * ncol, 8 * (size_t)ncol);
  L.den = (pip_i64 *)calloc(L.pcap + 1, 8);
  L.fl = (int *)calloc(L.pcap + 1, 4);
  L.csign = (signed char *)calloc(L.pcap + 1, 1);
  L.colpos = (int *)calloc(nvar + 4, 4);
  L.sbits = (unsigned *)calloc((L.pcap + 31) / 32 + 4, 4);
  L.active = (int *)calloc(L.pcap + 4, 4);
  L.cand = (int *)calloc((L.pcap > nvar ? L.pcap : nvar) + 4, 4);
  L.member = (unsigned char *)calloc(nvar + 16, 1);
  L.cut = (pip_i64 *)calloc(L.stride + 2, 8);
  int ctl[PIPL_NCTL] = {0};
  pip_i64 ctl64[8] = {0};
  ctl[PIPL_NI] = ni; ctl[PIPL_LDET] = 1; ctl64[2] = 1;
  L.ctl = ctl; L.ctl64 = ctl64;
  L.cells = cells;
  unsigned long long prof[16] = {0};
  L.prof = prof;
  // the staged (TMA) form of the row update or the direct one (pipemu_large_staged)
  e.stage = g_large_staged ? (pip_i64 *)calloc((size_t)(PIPL_NG + 1) * L.stride, 8) : nullptr;
  L.staged = e.stage ? 1 : 0;
  pipemu::set_order(order_mode);
  pipemu::run_warp(large_entry, &e);
  free(e.stage);
  *status = ctl[PIPL_STATUS];
  *ncells = ctl[PIPL_NCELL];
  if (info) { info[0] = ctl[PIPL_PIVOTS]; info[1] = ctl[PIPL_CUTS]; info[2] = (unsigned)ctl[PIPL_SKIPPED_LO]; info[3] = ctl[PIPL_NI]; }
  free(L.data); free(L.den); free(L.fl); free(L.csign); free(L.cand); free(L.member); free(L.cut); free(L.colpos); free(L.sbits); free(L.active);
  return 0;
}

// ---- warp-per-problem decoder in emulation, next to the thread decoder it must agree with ----
#include "../../piplib_b200/csrc/pip_decode_warp.h"

struct EmuDecode {
  const PipCell *cells; int n, bg, urs, flags;
  pip_i64 *tile; pip_i64 *out; long long cap; int narrow;
  long long len; pip_u64 h; unsigned wide; int ok;
};
static void decode_entry(void *a, int)
{
  EmuDecode *e = (EmuDecode *)a;
  PipWarpSer s;
  s.tile = e->tile; s.out = e->out; s.cap = e->cap; s.len = 0; s.fill = 0; s.h = 0;
  s.narrow_out = e->narrow; s.wide = 0;
  PipRawCells c = {e->cells};
  const bool ok = pip_wser_cells(s, c, e->n, e->bg, e->urs, e->flags);
  pip_wser_flush(s);
  const pip_u64 h = pip_wser_hash(s);
  if (W::lane() == 0) { e->len = s.len; e->h = h; e->wide = s.wide; e->ok = ok ? 1 : 0; }
}

// which = 0: pip_ser_cells (one thread, the host/device reference decoder), 1: the warp decoder
extern "C" int pipemu_decode(int which, const PipCell *cells, int n, int bg, int urs, int flags, long long *out,
                             long long cap, int narrow, long long *len, unsigned long long *hash, unsigned *wide,
                             int order_mode)
{
  if (which == 0) {
    PipSer s;
    s.out = out; s.cap = cap; s.len = 0; s.h = PIP_HASH_INIT; s.hashing = 1; s.narrow_out = narrow; s.wide = 0;
    PipRawCells c = {cells};
    const bool ok = pip_ser_cells(s, c, n, bg, urs, flags);
    *len = s.len; *hash = s.h; *wide = s.wide;
    return ok ? 1 : 0;
  }
  EmuDecode e;
  memset(&e, 0, sizeof e);
  e.cells = cells; e.n = n; e.bg = bg; e.urs = urs; e.flags = flags; e.out = out; e.cap = cap; e.narrow = narrow;
  e.tile = (pip_i64 *)malloc(sizeof(pip_i64) * PIP_WS_TILE);
  memset(e.tile, 0x5a, sizeof(pip_i64) * PIP_WS_TILE);
  pipemu::set_order(order_mode);
  pipemu::run_warp(decode_entry, &e);
  free(e.tile);
  *len = e.len; *hash = e.h; *wide = e.wide;
  return e.ok;
}
