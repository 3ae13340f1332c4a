/* Persistent warp loop: pull problem indices from a device-side queue, solve, append the
 * solution cells to this warp's window of the cell pool, write the per-problem record.
 * A warp retires when its window can no longer hold a worst-case solution (SOL_SIZE cells);
 * problems nobody solved keep status PIP_ST_PENDING and the host launches again for them. */
#ifndef PIP_WARP_MAIN_H
#define PIP_WARP_MAIN_H

#include "pip_solver.h"

/* one unit of work: a problem (offer < 0) or a donated subtree of one (offer = index of the claimed PipOffer);
 * returns the window space consumed, in cells */
template <class V, bool TEAM, bool STEAL, bool WORDS = false>
PIP_DEV pip_i64 pip_warp_unit(const PipLaunch &L, int warp_id, pip_i64 *arena, PipTeam *tm, PipCell *window, pip_i64 used,
                              pip_i64 *stk, int p, int offer, pip_i64 cpw)
{
  /* (test mode 2 slices the warp's frame stack per unit, see pip_warp_main) */
  const pip_i64 stk_cap = (STEAL && L.steal.mode == 2) ? L.stack_words_per_warp / 64 : L.stack_words_per_warp;
  const int lane = W::lane();
  const PipProblem P = L.prob[p];
  PipStats st;
  st.pivots = st.cuts = st.subsolves = st.splits = st.max_rows = st.max_cols = 0;
  st.elem_updates = 0;
  st.wrapped = 0;
#ifdef PIP_PROFILE
  for (int k = 0; k < PIP_NPHASE; k++) st.cyc[k] = 0;
  st.lap = clock64();
#endif
  int status = PIP_ST_OK, ncell = 0, hwm = 0;
  unsigned rflags = 0, nwords = 0;
  /* word mode (PipLaunch::emit_words): the solver writes the serialised quast itself into the window */
  const bool wordmode = WORDS || (L.emit_words && (P.flags & PIP_F_SIMPLE_SER));   /* (WORDS: the host checked every problem) */
  const PipSteal *stl = (STEAL && wordmode && L.steal.mode) ? &L.steal : nullptr;
  PipSolver<V, TEAM, STEAL, WORDS>::pip_solve_one(P, L.pool, L.pool_elem_log2, arena, L.work_words, L.slack_level, window + used, stk,
                stk_cap, L.sol_size, L.maxcol, L.maxparm, status, ncell, rflags, st, tm, &nwords,
                wordmode, L.have_layout ? &L.layout : nullptr, stl, p, offer,
                (STEAL && offer >= 0) ? L.steal.offers[offer].frame : nullptr, STEAL ? &hwm : nullptr,
                (L.images && L.have_layout) ? L.images + (pip_i64)p * L.image_words : nullptr, L.image_w1,
                (STEAL || offer >= 0) ? 0u : L.budget, &L.queue[2], L.heavy_max);
  if (!STEAL && status == PIP_ST_PENDING) {
    /* handed over: listed for the launch that follows, the record stays PENDING, the window is not consumed */
    if (lane == 0) L.heavy[W::atomic_add(&L.queue[2], 1u)] = p;      /* (heavy[] holds nprob entries) */
    W::sync();
    return 0;
  }
  if (lane == 0) {
    PipResult r;
    r.status = status; r.ncells = ncell;
    r.cell_off = L.cell_base + (pip_i64)warp_id * cpw + used;
    r.pivots = st.pivots; r.cuts = st.cuts; r.subsolves = st.subsolves; r.splits = st.splits;
    r.max_rows = st.max_rows; r.max_cols = st.max_cols; r.ser_words = 0;
    if (P.flags & PIP_F_SIMPLE_SER) {
      /* int32 storage keeps every value inside 31 bits, so the quast words fit int32 too */
      r.ser_words = nwords;
      rflags |= PIP_RES_SIZED | (PipVal<V>::narrow ? PIP_RES_SER32 : 0u);
    }
    if (wordmode) {
      /* in word mode PIP_RES_WIDE means "some word left int32": the complement is PIP_RES_SER32 */
      rflags = (rflags & ~(PIP_RES_WIDE | PIP_RES_SER32)) | PIP_RES_SIZED | PIP_RES_WORDS |
               ((rflags & PIP_RES_WIDE) ? 0u : PIP_RES_SER32) | (PipVal<V>::narrow ? PIP_RES_SRC32 : 0u);
    }
    r.elem_updates_lo = (unsigned)(st.elem_updates & 0xffffffffull);
    r.elem_updates_hi = (unsigned)(st.elem_updates >> 32);
    r.rflags = rflags;
    if (STEAL && offer >= 0) { L.steal.segs[offer] = r; L.steal.seg_hwm[offer] = hwm; }
    else {
      L.res[p] = r;
      if (STEAL && stl) L.steal.head_hwm[p] = hwm;
    }
#ifdef PIP_PROFILE
    if (L.prof) for (int k = 0; k < PIP_NPHASE; k++) atomicAdd(&L.prof[k], st.cyc[k]);
#endif
  }
  W::sync();
  /* window space consumed, in cells: the cells themselves, or the words written over them */
  if (wordmode) return ((pip_i64)((status == PIP_ST_OK || status == PIP_ST_VOID) ? nwords : 0u) * (pip_i64)sizeof(V) + (pip_i64)sizeof(PipCell) - 1) / (pip_i64)sizeof(PipCell);
  return ncell;
}

/* one out-of-line copy of the unit for the donation instantiation: it is entered from three places (the problem
 * loop, the test-mode drain, the idle loop), and inlining the solver three times made that kernel 15 000
 * instructions long -- in a kernel bound by the instruction cache */
template <class V, bool TEAM, bool STEAL, bool WORDS>
PIP_DEVNI pip_i64 pip_warp_unit_shared(const PipLaunch &L, int warp_id, pip_i64 *arena, PipTeam *tm, PipCell *window, pip_i64 used,
                                       pip_i64 *stk, int p, int offer, pip_i64 cpw)
{
  return pip_warp_unit<V, TEAM, STEAL, WORDS>(L, warp_id, arena, tm, window, used, stk, p, offer, cpw);
}

/* a claimed offer becomes segment `idx` of its problem: link it right after the donor's segment (later, inner
 * donations of the same donor thus come before earlier, outer ones: pre-order), then solve the subtree */
template <class V, bool WORDS = false>
PIP_DEV pip_i64 pip_warp_steal(const PipLaunch &L, int warp_id, pip_i64 *arena, PipCell *window, pip_i64 used, pip_i64 *stk, int idx,
                               pip_i64 cpw)
{
  const PipSteal &S = L.steal;
  int p = 0;
  if (W::lane() == 0) {
    W::fence();
    const PipOffer o = S.offers[idx];
    p = o.problem;
    int *cell = o.parent_seg < 0 ? &S.head_next[p] : &S.seg_next[o.parent_seg];
    S.seg_next[idx] = W::atomic_exch(cell, idx);
    W::atomic_add(&S.ctl[PIP_STL_CLAIMS], 1u);
  }
  p = W::shfl(p, 0);
  W::sync();
  return pip_warp_unit_shared<V, false, true, WORDS>(L, warp_id, arena, nullptr, window, used, stk, p, idx, cpw);
}

template <class V, bool TEAM = false, bool STEAL = false, bool WORDS = false>
PIP_DEV void pip_warp_main(const PipLaunch &L, int warp_id, pip_i64 *arena, PipTeam *tm = nullptr)
{
  const int lane = W::lane();
  pip_i64 cpw = L.cells_per_warp;       /* cells of this warp's window */
  pip_i64 *stk = L.stack + (pip_i64)warp_id * L.stack_words_per_warp;
  pip_i64 used = 0;
  /* the hand-over list of the previous launch (PipLaunch::from_heavy): its length is known on the device only;
   * a short list keeps most warps out of the way (other launches may share the machine) */
  unsigned nprob = (unsigned)L.nprob;
  unsigned *cursor = &L.queue[0];
  const int *order = L.order;
  if (STEAL && L.from_heavy) {
    nprob = W::load_volatile(&L.queue[2]);
    cursor = &L.queue[3];
    order = L.heavy;
    unsigned act = nprob * (unsigned)L.heavy_warps;
    if (act < 128u) act = 128u;
    if (nprob == 0 || (unsigned)warp_id >= act) return;
    /* the launch's cell region is shared out among the warps that stay: few listed problems, large windows
     * (a listed problem's stream is long, and a warp whose window is full cannot take subtrees either) */
    if ((pip_i64)act * cpw < L.heavy_region) cpw = L.heavy_region / (pip_i64)act;
  }
  PipCell *window = L.cells + L.cell_base + (pip_i64)warp_id * cpw;
  /* subtree donation: warps that have started (a CTA the hardware has not scheduled yet must not be waited for) */
  if (STEAL && L.steal.mode == 1 && L.emit_words && lane == 0) W::atomic_add(&L.steal.ctl[PIP_STL_TOTAL], 1u);
  for (;;) {
    if (cpw - used < (pip_i64)L.sol_size) break;
    unsigned q = 0;
    if (lane == 0) q = W::atomic_add(cursor, 1u);
    q = (unsigned)W::shfl((int)q, 0);
    if (q >= nprob) break;
    const int p = order ? order[q] : (int)q;
    if (STEAL) used += pip_warp_unit_shared<V, TEAM, STEAL, WORDS>(L, warp_id, arena, tm, window, used, stk, p, -1, cpw);
    else used += pip_warp_unit<V, TEAM, STEAL, WORDS>(L, warp_id, arena, tm, window, used, stk, p, -1, cpw);
  }
  if (!STEAL || !L.steal.mode || !L.emit_words) return;
  const PipSteal &S = L.steal;
  if (S.mode == 2) {
    /* test mode (one emulated warp): every offer was "claimed" at publication; solve them here, in order of
     * publication, including the ones the donated subtrees publish themselves */
    for (;;) {
      unsigned c = 0, n = 0;
      if (lane == 0) { c = W::load_volatile(&S.ctl[PIP_STL_CURSOR]); n = W::load_volatile(&S.ctl[PIP_STL_OFFERS]); }
      c = (unsigned)W::shfl((int)c, 0); n = (unsigned)W::shfl((int)n, 0);
      if (c >= n || c >= (unsigned)S.cap) break;
      if (cpw - used < (pip_i64)L.sol_size) break;
      if (lane == 0) W::atomic_add(&S.ctl[PIP_STL_CURSOR], 1u);
      /* (one warp plays donor and thief: every unit gets its own slice of the frame stack, so that the frames
       * still on offer are not overwritten by the subtree being solved) */
      const pip_i64 slice = L.stack_words_per_warp / 64;
      used += pip_warp_steal<V, WORDS>(L, warp_id, arena, window, used, stk + (1 + (c % 63)) * slice, (int)c, cpw);
    }
    return;
  }
  /* idle: look for offered subtrees until every warp that has started is idle (a donor is not idle, so nobody
   * leaves while an offer can still appear; an offer nobody takes is reclaimed by its owner, so leaving early
   * only loses parallelism, never an answer) */
  if (lane == 0) W::atomic_add(&S.ctl[PIP_STL_IDLE], 1u);
  unsigned scan = 0;                    /* every offer below was seen closed (states only move away from OPEN) */
  for (;;) {
    int got = -1;
    const bool room = cpw - used >= (pip_i64)L.sol_size;
    if (lane == 0 && room) {
      unsigned n = W::load_volatile(&S.ctl[PIP_STL_OFFERS]);
      if (n > (unsigned)S.cap) n = (unsigned)S.cap;
      bool all_closed = true;
      for (unsigned i = scan; i < n; i++) {
        const int stt = W::load_volatile(&S.offers[i].state);
        if (stt == PIP_OFFER_OPEN) {
          if (W::atomic_cas(&S.offers[i].state, PIP_OFFER_OPEN, PIP_OFFER_CLAIMED) == PIP_OFFER_OPEN) { got = (int)i; break; }
        }
        if (stt == PIP_OFFER_NONE) all_closed = false;          /* being published */
        if (all_closed && stt != PIP_OFFER_NONE) scan = i + 1;
      }
    }
    got = W::shfl(got, 0);
    if (got >= 0) {
      if (lane == 0) W::atomic_add(&S.ctl[PIP_STL_IDLE], 0xffffffffu);      /* idle-- */
      used += pip_warp_steal<V, WORDS>(L, warp_id, arena, window, used, stk, got, cpw);
      if (lane == 0) W::atomic_add(&S.ctl[PIP_STL_IDLE], 1u);
      continue;
    }
    unsigned idle = 0, total = 0;
    if (lane == 0) { idle = W::load_volatile(&S.ctl[PIP_STL_IDLE]); total = W::load_volatile(&S.ctl[PIP_STL_TOTAL]); }
    idle = (unsigned)W::shfl((int)idle, 0); total = (unsigned)W::shfl((int)total, 0);
    if (idle >= total) break;
    W::nap();
  }
}

/* helper warps of a team (class M): serve the leader's update commands until it retires */
template <class V>
PIP_DEV void pip_team_helper(PipTeam *tm, int tid)
{
  for (;;) {
    pip_team_barrier(tm->nthreads);
    const int cmd = tm->cmd;
    if (cmd == PIP_TEAM_EXIT) break;
    if (cmd == PIP_TEAM_FIRST) PipSolver<V, true>::pip_team_first_body(tm, tid);
    else if (cmd == PIP_TEAM_EXAM) PipSolver<V, true>::pip_team_exam_body(tm, tid);
    else {
      unsigned ovf = 0;
      bool fault = false;
      PipSolver<V, true>::pip_update_rows(tm->B, tm->T, tm->pivi, tm->pivj, (V)tm->pivot, (V)tm->dpiv, tid, tm->nthreads,
                                          ovf, fault);
      if (fault) tm->fault = 1;
      if (ovf) tm->ovf = 1;
    }
    pip_team_barrier(tm->nthreads);
  }
}

#endif
