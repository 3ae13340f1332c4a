/* Persistent warp loop: pull problem indices from a device-side queue, solve, append the
 * solution cells to this warp's window of the cell pool, write the per-problem record.
 * A warp retires when its window can no longer hold a worst-case solution (SOL_SIZE cells);
 * problems nobody solved keep status PIP_ST_PENDING and the host launches again for them. */
#ifndef PIP_WARP_MAIN_H
#define PIP_WARP_MAIN_H

#include "pip_solver.h"

template <class V, bool TEAM = false>
PIP_DEV void pip_warp_main(const PipLaunch &L, int warp_id, pip_i64 *arena, PipTeam *tm = nullptr)
{
  const int lane = W::lane();
  PipCell *window = L.cells + (pip_i64)warp_id * L.cells_per_warp;
  pip_i64 *stk = L.stack + (pip_i64)warp_id * L.stack_words_per_warp;
  pip_i64 used = 0;
  for (;;) {
    if (L.cells_per_warp - used < (pip_i64)L.sol_size) break;
    unsigned q = 0;
    if (lane == 0) q = W::atomic_add(&L.queue[0], 1u);
    q = (unsigned)W::shfl((int)q, 0);
    if (q >= (unsigned)L.nprob) break;
    const int p = L.order ? L.order[q] : (int)q;
    const PipProblem P = L.prob[p];
    PipStats st;
    st.pivots = st.cuts = st.subsolves = st.splits = st.max_rows = st.max_cols = 0;
    st.elem_updates = 0;
    st.wrapped = 0;
#ifdef PIP_PROFILE
    for (int k = 0; k < PIP_NPHASE; k++) st.cyc[k] = 0;
    st.lap = clock64();
#endif
    int status = PIP_ST_OK, ncell = 0;
    unsigned rflags = 0, nwords = 0;
    /* word mode (PipLaunch::emit_words): the solver writes the serialised quast itself into the window */
    const bool wordmode = L.emit_words && (P.flags & PIP_F_SIMPLE_SER);
    PipSolver<V, TEAM>::pip_solve_one(P, L.pool, L.pool_elem_log2, arena, L.work_words, L.slack_level, window + used, stk,
                  L.stack_words_per_warp, L.sol_size, L.maxcol, L.maxparm, status, ncell, rflags, st, tm, &nwords,
                  wordmode, L.have_layout ? &L.layout : nullptr);
    if (lane == 0) {
      PipResult r;
      r.status = status; r.ncells = ncell;
      r.cell_off = (pip_i64)warp_id * L.cells_per_warp + used;
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
      L.res[p] = r;
#ifdef PIP_PROFILE
      if (L.prof) for (int k = 0; k < PIP_NPHASE; k++) atomicAdd(&L.prof[k], st.cyc[k]);
#endif
    }
    /* window space consumed, in cells: the cells themselves, or the words written over them */
    if (wordmode) used += ((pip_i64)((status == PIP_ST_OK || status == PIP_ST_VOID) ? nwords : 0u) * (pip_i64)sizeof(V) + (pip_i64)sizeof(PipCell) - 1) / (pip_i64)sizeof(PipCell);
    else used += ncell;
    W::sync();
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
