/* Grid-per-problem solver for one large non-parametric tableau (BASELINE config 4: ~4096 x 4097
 * int64, 134 MB, HBM-bound).  The whole grid works on one tableau that lives in global memory;
 * a pivot is two grid-wide phases separated by cooperative grid syncs:
 *
 *   phase AB (CTA 0)   row pick in position order (chercher + exam_coef, source/traiter.c:669-680),
 *                      cut generation when no row is negative (integrer case (d), source/integrer.c:
 *                      342-481), lexicographic pivot column (choisir_piv, source/traiter.c:297-341)
 *                      as a candidate-set filtering pass over the positions, determinant / overflow
 *                      bookkeeping (source/traiter.c:394-447)
 *   phase C (all CTAs) rank-1 update of every stored row against the pivot row with the per-row gcd
 *                      normalisation (source/traiter.c:467-502); re-flag and constant-sign bookkeeping
 *                      fused in.  Every CTA scans its own stripe of positions for the rows whose update
 *                      is not the identity, stages the pivot row and then one row per group of warps
 *                      in shared memory with cp.async.bulk (TMA engine, mbarrier completion) and
 *                      writes each row back once with 16-byte stores; rows beyond PIPL_KEEP per CTA
 *                      are shared through an overflow queue.  Short rows (<= 64 words) take a warp each.
 *
 * nparm = 0 only (no context, no splits): that is what "one large dense ILP tableau" is.  The
 * position/slot model, flags and arithmetic are those of pip_solver.h.  The code is written against
 * the tiny G:: abstraction so that tests/emu can run it on the CPU as one 32-thread CTA.
 */
#ifndef PIP_LARGE_H
#define PIP_LARGE_H

#include "pip_arith.h"
#include "pip_types.h"
#include "simt.h"

#define PIPL_INF 0x7fffffff
#if defined(PIPL_WALK_STATS) && defined(__CUDACC__)
__device__ unsigned long long pipl_dbg[8];   /* diagnostic build only: walks, candidates, window steps / stops (CTA, warp) */
#define PIPL_DBG(i, v) do { if (G::tid() == 0) pipl_dbg[i] += (unsigned long long)(v); } while (0)
__device__ unsigned long long pipl_cdbg[8];  /* update-phase laps of CTA 0 */
#define PIPL_CLAP(i) do { if (G::cta() == 0 && G::tid() == 0) { const long long n_ = pip_clock(); pipl_cdbg[i] += (unsigned long long)(n_ - clap); clap = n_; } } while (0)
#define PIPL_CLAP_BEGIN long long clap = pip_clock()
__device__ unsigned long long pipl_adbg[16];  /* choice-phase laps of CTA 0 */
#define PIPL_ALAP(i) do { if (G::tid() == 0) { const long long n_ = pip_clock(); pipl_adbg[i] += (unsigned long long)(n_ - alap); alap = n_; } } while (0)
#define PIPL_ALAP_BEGIN long long alap = pip_clock()
#else
#define PIPL_ALAP(i) do { } while (0)
#define PIPL_ALAP_BEGIN do { } while (0)
#define PIPL_CLAP(i) do { } while (0)
#define PIPL_CLAP_BEGIN do { } while (0)
#define PIPL_DBG(i, v) do { } while (0)
#endif
#define PIPL_K 8            /* candidates per thread kept in registers by the column walk */
#ifndef PIPL_NG
#define PIPL_NG 4            /* groups of warps per CTA in the update phase (one row per group at a time) */
#endif
#ifndef PIPL_KEEP
#define PIPL_KEEP (2 * PIPL_NG)   /* rows a CTA updates itself before it shares through the overflow queue */
#endif
#define PIPL_RP 16           /* positions per thread per round of the row-pick sweep */
#define PIPL_LCAP 128        /* local list capacity of the update phase */
#define PIPL_RED_INTS (128 + 6 * PIPL_LCAP + 64 * PIPL_NG + 2 * (PIPL_NG + 1))   /* shared scratch of the kernel, in ints */
#define PIPL_AL 8           /* pivot-row entries per thread per round of the candidate scan */
/* sub-phase timers of CTA 0 (thread 0): prof[2..7] = swap, row pick, column choice, determinant,
 * (the single-CTA active-row list of earlier versions: now ~0), spare */
#define PIPL_T(i) do { if (tid == 0) { const long long n_ = pip_clock(); L.prof[i] += (unsigned long long)(n_ - tlap); tlap = n_; } } while (0)

struct PipLarge {
  /* problem */
  int nvar, ni, flags;            /* nparm = 0 */
  int staged;                     /* rows go through shared memory (cp.async.bulk), see pipl_update_row_staged */
  int stride;                     /* words per slot (even) */
  int pcap, rcap;                 /* position / slot capacity */
  int sol_size, maxcol;
  pip_i64 *data;                  /* rcap x stride */
  pip_i64 *den;                   /* [pcap] */
  int *fl;                        /* [pcap] flag | link << 8 */
  signed char *csign;             /* [pcap] sign of the constant column of each stored row */
  int *colpos;                    /* [nvar] position of the Unit row that owns column j */
  unsigned *sbits;                /* [(pcap+31)/32] bitmap of the stored (non-Unit) positions */
  int *active;                    /* [pcap] positions whose update is not the identity, this pivot */
  /* scratch of phase AB */
  int *cand;                      /* [nvar] compact candidate list */
  unsigned char *member;          /* [nvar] */
  pip_i64 *cut;                   /* [stride] */
  /* control block (global, written by CTA 0, read by everyone after a grid sync) */
  int *ctl;                       /* see PIPL_* below */
  pip_i64 *ctl64;                 /* [8]: pivot, dpiv, det[4] */
  PipCell *cells;
  unsigned long long *prof;       /* [8] cycle counters of CTA 0: AB, C, sync */
};
enum { PIPL_ACTION = 0, PIPL_PIVI, PIPL_PIVJ, PIPL_STATUS, PIPL_NCELL, PIPL_NI, PIPL_LDET, PIPL_PIVOTS,
       PIPL_CUTS, PIPL_SKIPPED_LO, PIPL_SKIPPED_HI, PIPL_NACTIVE, PIPL_NEXT, PIPL_PUSHED, PIPL_NCTL = 16 };
enum { PIPL_GO = 0, PIPL_STOP = 1 };

/* ---- CTA-level helpers (shared scratch: int red[64]) ------------------------------------- */
PIP_DEV int pipl_cta_min(int v, int *red)
{
  v = (int)W::redmin((unsigned)v + 0x80000000u) - (int)0x80000000u;
  const int lane = W::lane(), wid = G::tid() >> 5, nw = (G::T() + 31) >> 5;
  G::cta_sync();
  if (lane == 0) red[wid] = v;
  G::cta_sync();
  int r = PIPL_INF;
  for (int i = 0; i < nw; i++) r = red[i] < r ? red[i] : r;
  return r;
}
PIP_DEV int pipl_cta_max(int v, int *red) { return -pipl_cta_min(-v, red); }
PIP_DEV int pipl_cta_sum(int v, int *red)
{
  const int lane = W::lane(), wid = G::tid() >> 5, nw = (G::T() + 31) >> 5;
  for (int o = 16; o > 0; o >>= 1) { int y = W::shfl_down(v, o); if (lane + o < 32) v += y; }
  G::cta_sync();
  if (lane == 0) red[wid] = v;
  G::cta_sync();
  int r = 0;
  for (int i = 0; i < nw; i++) r += red[i];
  return r;
}

PIP_DEV pip_i64 *pipl_row(const PipLarge &L, int slot) { return L.data + (pip_i64)slot * L.stride; }

/* a/b < c/d for b, d > 0, exactly (128-bit products) */
PIP_DEV int pipl_ratio_cmp(pip_i64 a, pip_i64 b, pip_i64 c, pip_i64 d)
{
  const pip_i64 lh = pip_mulhi(a, d), rh = pip_mulhi(c, b);
  const pip_u64 ll = (pip_u64)a * (pip_u64)d, rl = (pip_u64)c * (pip_u64)b;
  if (lh != rh) return lh < rh ? -1 : 1;
  if (ll != rl) return ll < rl ? -1 : 1;
  return 0;
}

/* the next PIPL_WIN stored (non-Unit) positions at or after k, in order (PIPL_INF-padded); every
 * thread computes the same list from the bitmap */
#ifndef PIPL_WIN
#define PIPL_WIN 4            /* measured on B200, 4096 x 4097: 4 rows per step 41.6 us per pivot, 3: 43.3, 8: 44.0, 2: 46.3, 16: 49.9 */
#endif
PIP_DEV void pipl_window(const unsigned *sb, int k, int nl, int nwords, int *wp)
{
  int w = k >> 5;
  unsigned bits = w < nwords ? (sb[w] & (~0u << (k & 31))) : 0u;
  #pragma unroll
  for (int b = 0; b < PIPL_WIN; b++) {
    while (!bits && w + 1 < nwords) { w++; bits = sb[w]; }
    if (bits) {
      const int pp = (w << 5) + pip_ffs(bits) - 1;
      bits &= bits - 1;
      wp[b] = pp < nl ? pp : PIPL_INF;
    } else wp[b] = PIPL_INF;
  }
}

/* rebuild the compact candidate list from member[] (ordered by column), returns its length: every
 * thread owns a contiguous run of columns, one exclusive scan across the CTA */
PIP_DEV int pipl_compact(const PipLarge &L, int *red, int *sh_base)
{
  const int tid = G::tid(), T = G::T();
  const int per = (L.nvar + T - 1) / T;
  const int j0 = tid * per, j1 = j0 + per < L.nvar ? j0 + per : L.nvar;
  int m = 0;
  for (int j = j0; j < j1; j++) m += L.member[j] ? 1 : 0;
  int x = m;
  const int lane = W::lane(), wid = tid >> 5, nw = (T + 31) >> 5;
  for (int o = 1; o < 32; o <<= 1) { int y = W::shfl_up(x, o); if (lane >= o) x += y; }
  G::cta_sync();
  if (lane == 31) red[wid] = x;
  G::cta_sync();
  int before = 0, all = 0;
  for (int i = 0; i < nw; i++) { if (i < wid) before += red[i]; all += red[i]; }
  int at = before + x - m;
  for (int j = j0; j < j1; j++) if (L.member[j]) L.cand[at++] = j;
  G::cta_sync();
  (void)sh_base;
  return all;
}

/* ---- phase AB: CTA 0 ------------------------------------------------------------------------ */
PIP_DEV void pipl_put(const PipLarge &L, int idx, int kind, pip_i64 p1, pip_i64 p2)
{
  L.cells[idx].kind = kind; L.cells[idx].pad = 0; L.cells[idx].p1 = p1; L.cells[idx].p2 = p2;
}

/* finish with a status (and, for OK, the solution list); CTA 0 only */
PIP_DEV void pipl_finish(const PipLarge &L, int status, int kind /*0 none, 1 nil, 2 solution*/)
{
  const int tid = G::tid(), T = G::T();
  int ncell = 0;
  if (status == PIP_ST_OK && kind == 1) {
    if (1 >= L.sol_size) status = PIP_ST_FATAL + 26;
    else { if (tid == 0) pipl_put(L, 0, PIP_C_NIL, 0, 0); ncell = 1; }
  } else if (status == PIP_ST_OK && kind == 2) {
    const int total = 1 + 2 * L.nvar;                 /* solution_xx with nparm = 0 */
    if (total >= L.sol_size) status = PIP_ST_FATAL + 26;
    else {
      for (int c = tid; c < total; c += T) {
        if (c == 0) { pipl_put(L, 0, PIP_C_LIST, L.nvar, 0); continue; }
        const int i = (c - 1) >> 1;
        if ((c - 1) & 1) {
          const int f = L.fl[i];
          const pip_i64 d = L.den[i];
          const pip_i64 v = (f & PIP_UNIT) ? (PIP_LINK(f) == L.nvar ? d : 0) : pipl_row(L, PIP_LINK(f))[L.nvar];
          pipl_put(L, c, PIP_C_VAL, v, d);
        } else pipl_put(L, c, PIP_C_FORM, 1, 0);
      }
      ncell = total;
    }
  }
  G::cta_sync();
  if (tid == 0) {
    L.ctl[PIPL_STATUS] = status;
    L.ctl[PIPL_NCELL] = status == PIP_ST_OK ? ncell : 0;
    L.ctl[PIPL_ACTION] = PIPL_STOP;
#if defined(PIPL_WALK_STATS) && defined(__CUDACC__)
    printf("walk stats: walks %llu, candidates %llu, CTA path: window steps %llu stops %llu (candidates left after stops %llu); warp path: window steps %llu stops %llu\n",
           pipl_dbg[0], pipl_dbg[1], pipl_dbg[2], pipl_dbg[3], pipl_dbg[6], pipl_dbg[4], pipl_dbg[5]);
    printf("update phase, CTA 0, us per pivot: scan %.2f publish %.2f own rows %.2f wait %.2f overflow %.2f\n",
           pipl_cdbg[0] / 1965.0 / L.ctl[PIPL_PIVOTS], pipl_cdbg[1] / 1965.0 / L.ctl[PIPL_PIVOTS], pipl_cdbg[2] / 1965.0 / L.ctl[PIPL_PIVOTS],
           pipl_cdbg[3] / 1965.0 / L.ctl[PIPL_PIVOTS], pipl_cdbg[4] / 1965.0 / L.ctl[PIPL_PIVOTS]);
    printf("choice phase, us per pivot: sweep %.2f reduce %.2f reflag %.2f | prow scan %.2f registers %.2f CTA walk %.2f warp finish %.2f (other %.2f)\n",
           pipl_adbg[0] / 1965.0 / L.ctl[PIPL_PIVOTS], pipl_adbg[1] / 1965.0 / L.ctl[PIPL_PIVOTS], pipl_adbg[2] / 1965.0 / L.ctl[PIPL_PIVOTS],
           pipl_adbg[4] / 1965.0 / L.ctl[PIPL_PIVOTS], pipl_adbg[5] / 1965.0 / L.ctl[PIPL_PIVOTS], pipl_adbg[6] / 1965.0 / L.ctl[PIPL_PIVOTS],
           pipl_adbg[7] / 1965.0 / L.ctl[PIPL_PIVOTS], pipl_adbg[3] / 1965.0 / L.ctl[PIPL_PIVOTS]);
    for (int i = 0; i < 8; i++) { pipl_dbg[i] = 0; pipl_cdbg[i] = 0; }
    for (int i = 0; i < 16; i++) pipl_adbg[i] = 0;
#endif
  }
}

/* One AB phase.  Leaves ctl[ACTION] = GO with (pivi, pivj, pivot, dpiv) or STOP with a status. */
PIP_DEV void pipl_phase_ab(const PipLarge &L, int *red, bool first_call, pip_i64 *stage)
{
  const int tid = G::tid(), T = G::T();
  const int nvar = L.nvar, ncol = nvar + 1;
  int ni = L.ctl[PIPL_NI];
  int nl = nvar + ni;
  long long tlap = pip_clock();

  /* the swap of the previous pivot (source/traiter.c:503-516) */
  if (!first_call) {
    const int pivi = L.ctl[PIPL_PIVI], pivj = L.ctl[PIPL_PIVJ];
    const pip_i64 pivot = L.ctl64[0], dpiv = L.ctl64[1];
    const int pslot = PIP_LINK(L.fl[pivi]);
    pip_i64 *prow = pipl_row(L, pslot);
    const int ku = L.colpos[pivj];          /* the Unit position that owns the pivot column */
    for (int base = 0; base < ncol; base += PIPL_AL * T) {        /* loads of a round first: one round trip */
      pip_i64 v[PIPL_AL];
      #pragma unroll
      for (int i = 0; i < PIPL_AL; i++) { const int j = base + i * T + tid; v[i] = j < ncol ? prow[j] : 0; }
      #pragma unroll
      for (int i = 0; i < PIPL_AL; i++) { const int j = base + i * T + tid; if (j < ncol) prow[j] = (j == pivj) ? dpiv : -v[i]; }
    }
    G::cta_sync();
    if (tid == 0) {
      const pip_i64 c = prow[nvar];
      L.fl[ku] = PIP_MKFL(PIP_PLUS, pslot); L.den[ku] = pivot;
      L.csign[ku] = c < 0 ? -1 : c > 0 ? 1 : 0;
      L.fl[pivi] = PIP_MKFL(PIP_UNIT | PIP_ZERO, pivj); L.den[pivi] = 1;
      L.colpos[pivj] = pivi;
      L.sbits[pivi >> 5] &= ~(1u << (pivi & 31));
      L.sbits[ku >> 5] |= 1u << (ku & 31);
    }
    G::cta_sync();
  }

  /* the bitmap of the stored positions is walked word by word, each word a dependent load: CTA 0 keeps
   * a copy in shared memory (the staging buffers of the update phase are idle during this phase) */
  const bool sbs = stage && (size_t)((L.pcap + 31) >> 5) * sizeof(unsigned) <= (size_t)(PIPL_NG + 1) * L.stride * sizeof(pip_i64);
  unsigned *sbw = sbs ? (unsigned *)stage : L.sbits;
  const unsigned *sb = sbw;
  if (sbs) for (int w = tid; w < ((L.pcap + 31) >> 5); w += T) sbw[w] = L.sbits[w];
  PIPL_T(2);
  PIPL_ALAP_BEGIN;
  int pivi = PIPL_INF;
  for (;;) {
    /* chercher(Minus), and in the same sweep the first Unknown row with a negative constant (what
     * exam_coef with nparm = 0 would stop at: an Unknown row takes the sign of its constant, rows
     * after the first negative one stay Unknown) */
    int c = PIPL_INF, c2 = PIPL_INF;
    /* PIPL_RP positions per thread per round, every load of a round issued before the first use (the
     * sweep is one L2 round trip per round, not one per position) */
    int f[PIPL_RP];
    signed char sg[PIPL_RP];
    for (int base = 0; base < nl; base += PIPL_RP * T) {
      #pragma unroll
      for (int i = 0; i < PIPL_RP; i++) {
        const int k = base + i * T + tid;
        f[i] = k < nl ? L.fl[k] : 0;
        sg[i] = k < nl ? L.csign[k] : (signed char)0;
      }
      #pragma unroll
      for (int i = PIPL_RP - 1; i >= 0; i--) {
        const int k = base + i * T + tid;
        if (k >= nl) continue;
        if ((f[i] & PIP_MINUS) && k < c) c = k;
        if (PIP_FLAG(f[i]) == PIP_UNKNOWN && sg[i] < 0 && k < c2) c2 = k;
      }
    }
    PIPL_ALAP(0);
    {
      c = (int)W::redmin((unsigned)c);
      c2 = (int)W::redmin((unsigned)c2);
      const int lane = W::lane(), wid = tid >> 5, nw = (T + 31) >> 5;
      G::cta_sync();
      if (lane == 0) { red[wid] = c; red[32 + wid] = c2; }
      G::cta_sync();
      c = PIPL_INF; c2 = PIPL_INF;
      for (int i = 0; i < nw; i++) { c = red[i] < c ? red[i] : c; c2 = red[32 + i] < c2 ? red[32 + i] : c2; }
    }
    pivi = c;
    PIPL_ALAP(1);
    if (pivi < nl) break;
    const int firstneg = c2;
    if (nl <= PIPL_RP * T) {
      /* one round: the flags and signs of the sweep are still in registers */
      #pragma unroll
      for (int i = 0; i < PIPL_RP; i++) {
        const int k = i * T + tid;
        if (k >= nl || PIP_FLAG(f[i]) != PIP_UNKNOWN || k > firstneg) continue;
        L.fl[k] = PIP_MKFL(sg[i] < 0 ? PIP_MINUS : sg[i] > 0 ? PIP_PLUS : PIP_ZERO, PIP_LINK(f[i]));
      }
    } else {
      for (int k = tid; k < nl; k += T) {
        const int fk = L.fl[k];
        if (PIP_FLAG(fk) != PIP_UNKNOWN || k > firstneg) continue;
        const int s = L.csign[k];
        L.fl[k] = PIP_MKFL(s < 0 ? PIP_MINUS : s > 0 ? PIP_PLUS : PIP_ZERO, PIP_LINK(fk));
      }
    }
    G::cta_sync();
    PIPL_ALAP(2);
    if (firstneg < nl) { pivi = firstneg; break; }
    /* all rows non-negative */
    if (!(L.flags & PIP_F_INT)) { pipl_finish(L, PIP_ST_OK, 2); return; }
    /* integrer, nparm = 0: cases (a), (b), (d) */
    if (ncol >= L.maxcol) { pipl_finish(L, PIP_ST_FATAL + 3, 0); return; }
    int cand = PIPL_INF;
    for (int i = tid; i < nvar; i += T) {
      const int f = L.fl[i];
      if (!(f & PIP_UNIT) && L.den[i] != 1 && i < cand) cand = i;
    }
    int verdict = 0, row_i = pipl_cta_min(cand, red);
    while (row_i < nvar) {
      const pip_i64 D = L.den[row_i];
      if (D == 0) { pipl_finish(L, PIP_ST_FAULT, 0); return; }
      const pip_i64 *row = pipl_row(L, PIP_LINK(L.fl[row_i]));
      int okv = 0, okc = 0;
      for (int j = tid; j < ncol; j += T) {
        pip_i64 x;
        if (j < nvar) { x = pip_mod(row[j], D); okv |= x > 0; }
        else { x = -pip_mod(-row[j], D); okc |= x != 0; }
        L.cut[j] = x;
      }
      okv = pipl_cta_max(okv, red);
      okc = pipl_cta_max(okc, red);
      if (okc) { verdict = okv ? 1 : -1; break; }
      /* case (a): integral row, look for the next candidate */
      cand = PIPL_INF;
      for (int i = row_i + 1 + tid; i < nvar; i += T) {
        const int f = L.fl[i];
        if (!(f & PIP_UNIT) && L.den[i] != 1 && i < cand) cand = i;
      }
      row_i = pipl_cta_min(cand, red);
    }
    if (verdict == 0) { pipl_finish(L, PIP_ST_OK, 2); return; }
    if (verdict < 0) { pipl_finish(L, PIP_ST_OK, 1); return; }
    if (ni >= L.rcap || nl >= L.pcap) { pipl_finish(L, PIP_ST_CAPACITY, 0); return; }
    {
      pip_i64 *nr = pipl_row(L, ni);
      for (int j = tid; j < ncol; j += T) nr[j] = L.cut[j];
      G::cta_sync();
      if (tid == 0) {
        L.fl[nl] = PIP_MKFL(PIP_MINUS, ni);
        L.den[nl] = L.den[row_i];
        L.csign[nl] = L.cut[nvar] < 0 ? -1 : L.cut[nvar] > 0 ? 1 : 0;
        L.ctl[PIPL_NI] = ni + 1;
        L.sbits[nl >> 5] |= 1u << (nl & 31);
        if (sbs) sbw[nl >> 5] |= 1u << (nl & 31);
        L.ctl[PIPL_CUTS] = L.ctl[PIPL_CUTS] + 1;
      }
      G::cta_sync();
      pivi = nl;
      ni++; nl++;
      break;
    }
  }

  /* ---- choisir_piv (source/traiter.c:297-341) as candidate-set filtering in position order ----
   * S = columns with a positive pivot-row entry.  Walking the positions in order, a Unit position
   * strikes its own column off S (its ratio is 1/prow > 0 against 0 for everybody else) and a
   * stored row keeps only the columns with the minimal ratio; the last column standing wins.
   * Unit positions are never touched one by one: between two stored rows every member whose Unit
   * position colpos[j] lies in the gap is struck at once (if that is all of S, the one with the
   * largest colpos survives).  The loop therefore runs once per *stored* row met before the
   * decision -- typically once or twice. */
  PIPL_T(3);
  PIPL_ALAP(3);
  const pip_i64 *prow = pipl_row(L, PIP_LINK(L.fl[pivi]));
  const int nwords = (nl + 31) >> 5;
  int pivj = PIPL_INF;
  bool decided = false;
  /* Usual case (about 100 positive entries in the pivot row of a 4096-column tableau): at most one
   * candidate per thread.  The candidates are collected in shared memory, live in registers (column,
   * pivot-row entry, Unit position, alive) for the whole walk, and the survivor comes out of one
   * reduction -- no member[] / cand[] round trips through global memory. */
  {
    int *scnt = red + 126, *scj = red + 128;
    const int cap = T < 6 * PIPL_LCAP ? T : 6 * PIPL_LCAP;
    if (tid == 0) *scnt = 0;
    G::cta_sync();
    for (int base = 0; base < nvar; base += PIPL_AL * T) {
      pip_i64 pv[PIPL_AL];
      #pragma unroll
      for (int i = 0; i < PIPL_AL; i++) { const int j = base + i * T + tid; pv[i] = j < nvar ? prow[j] : 0; }
      #pragma unroll
      for (int i = 0; i < PIPL_AL; i++)
        if (pv[i] > 0) { const int at = (int)G::atomic_add_u((unsigned *)scnt, 1u); if (at < cap) scj[at] = base + i * T + tid; }
    }
    G::cta_sync();
    PIPL_ALAP(4);
    const int n0 = *scnt;
    if (n0 == 0) { pipl_finish(L, PIP_ST_OK, 1); return; }
    if (n0 <= cap) {
      decided = true;
      PIPL_DBG(0, 1); PIPL_DBG(1, n0);
      bool alive = tid < n0;
      int cj = alive ? scj[tid] : 0;
      const pip_i64 cp = alive ? prow[cj] : 1;
      int cu = alive ? L.colpos[cj] : PIPL_INF;
      int ncand = n0, k = 0;
      pip_i64 *red64 = (pip_i64 *)red;
      G::cta_sync();
      while (ncand > 32) {
        int pst = PIPL_INF;
        for (;;) {
          int wp[PIPL_WIN];
          pipl_window(sb, k, nl, nwords, wp);
          if (wp[0] == PIPL_INF) break;
          PIPL_DBG(2, 1);
          int first = PIPL_WIN;
          if (alive) {
            pip_i64 v[PIPL_WIN];
            #pragma unroll
            for (int b = 0; b < PIPL_WIN; b++) v[b] = wp[b] < cu ? pipl_row(L, PIP_LINK(L.fl[wp[b]]))[cj] : 0;
            #pragma unroll
            for (int b = PIPL_WIN - 1; b >= 0; b--) if (v[b] != 0) first = b;
          }
          /* one exchange for both questions of the step: the first row with a non-zero entry, and how many
           * candidates are still in play beyond this window (cu > PIPL_INF never holds) */
          int inplay = pip_popc(W::ballot(alive && cu > wp[PIPL_WIN - 1]));
          first = (int)W::redmin((unsigned)first);
          {
            const int lane = W::lane(), wid = tid >> 5, nw = (T + 31) >> 5;
            G::cta_sync();
            if (lane == 0) { red[wid] = first; red[32 + wid] = inplay; }
            G::cta_sync();
            first = PIPL_WIN; inplay = 0;
            for (int i = 0; i < nw; i++) { first = red[i] < first ? red[i] : first; inplay += red[32 + i]; }
          }
          if (first < PIPL_WIN) { pst = wp[first]; break; }
          if (wp[PIPL_WIN - 1] == PIPL_INF) break;
          if (inplay <= 1) break;
          k = wp[PIPL_WIN - 1] + 1;
        }
        const bool struck = alive && cu < pst;
        const int nel = pipl_cta_sum(struck ? 1 : 0, red);
        if (nel >= ncand) {
          const int umax = pipl_cta_max(struck ? cu : -1, red);
          if (alive && cu != umax) alive = false;
          ncand = 1;
          break;
        }
        if (struck) alive = false;
        ncand -= nel;
        if (pst >= nl || ncand <= 1) break;
        const pip_i64 *row = pipl_row(L, PIP_LINK(L.fl[pst]));
        const pip_i64 va = alive ? row[cj] : 0;
        pip_i64 ba = va, bp = cp;
        int valid = alive ? 1 : 0;
        for (int o = 16; o > 0; o >>= 1) {
          const pip_i64 oa = W::shfl_xor64(ba, o), op = W::shfl_xor64(bp, o);
          const int ov = W::shfl_xor(valid, o);
          if (ov && (!valid || pipl_ratio_cmp(oa, op, ba, bp) < 0)) { ba = oa; bp = op; valid = 1; }
        }
        {
          const int lane = W::lane(), wid = tid >> 5, nw = (T + 31) >> 5;
          G::cta_sync();
          if (lane == 0) { red64[2 * wid] = ba; red64[2 * wid + 1] = bp; red[96 + wid] = valid; }
          G::cta_sync();
          valid = 0;
          for (int i = 0; i < nw; i++)
            if (red[96 + i] && (!valid || pipl_ratio_cmp(red64[2 * i], red64[2 * i + 1], ba, bp) < 0)) {
              ba = red64[2 * i]; bp = red64[2 * i + 1]; valid = 1;
            }
        }
        const bool out = alive && pipl_ratio_cmp(va, cp, ba, bp) != 0;
        if (out) alive = false;
        ncand -= pipl_cta_sum(out ? 1 : 0, red);
        k = pst + 1;
        PIPL_DBG(3, 1); PIPL_DBG(6, ncand);
      }
      if (ncand > 1) {
        /* at most 32 left: warp 0 finishes the walk on its own, no CTA barriers */
        G::cta_sync();
        if (tid == 0) *scnt = 0;
        G::cta_sync();
        if (alive) scj[G::atomic_add_u((unsigned *)scnt, 1u)] = cj;
        G::cta_sync();
        if (tid < 32) {
          const int lane = tid;
          int n = *scnt;
          alive = lane < n;
          cj = alive ? scj[lane] : 0;
          const pip_i64 pj = alive ? prow[cj] : 1;
          const int u = alive ? L.colpos[cj] : PIPL_INF;
          while (n > 1) {
            int pst = PIPL_INF;
            for (;;) {
              int wp[PIPL_WIN];
              pipl_window(sb, k, nl, nwords, wp);
              if (wp[0] == PIPL_INF) break;
              PIPL_DBG(4, 1);
              pip_i64 v[PIPL_WIN];
              #pragma unroll
              for (int b = 0; b < PIPL_WIN; b++)
                v[b] = (alive && wp[b] < u) ? pipl_row(L, PIP_LINK(L.fl[wp[b]]))[cj] : 0;
              int first = PIPL_WIN;
              #pragma unroll
              for (int b = PIPL_WIN - 1; b >= 0; b--) if (v[b] != 0) first = b;
              first = (int)W::redmin((unsigned)first);
              if (first < PIPL_WIN) { pst = wp[first]; break; }
              if (wp[PIPL_WIN - 1] == PIPL_INF) break;
              if (pip_popc(W::ballot(alive && u > wp[PIPL_WIN - 1])) <= 1) break;
              k = wp[PIPL_WIN - 1] + 1;
            }
            const unsigned mel = W::ballot(alive && u < pst);
            const int nel = pip_popc(mel);
            if (nel >= n) {
              const int umax = (int)W::redmax(alive ? (unsigned)u : 0u);
              alive = alive && u == umax;
              n = 1;
              break;
            }
            if (alive && u < pst) alive = false;
            n -= nel;
            if (pst >= nl || n <= 1) break;
            const pip_i64 *row = pipl_row(L, PIP_LINK(L.fl[pst]));
            const pip_i64 a = alive ? row[cj] : 0;
            pip_i64 ba = a, bp = pj;
            int valid = alive ? 1 : 0;
            for (int o = 16; o > 0; o >>= 1) {
              const pip_i64 oa = W::shfl_xor64(ba, o), op = W::shfl_xor64(bp, o);
              const int ov = W::shfl_xor(valid, o);
              if (ov && (!valid || pipl_ratio_cmp(oa, op, ba, bp) < 0)) { ba = oa; bp = op; valid = 1; }
            }
            if (alive && pipl_ratio_cmp(a, pj, ba, bp) != 0) alive = false;
            n = pip_popc(W::ballot(alive));
            k = pst + 1;
            PIPL_DBG(5, 1);
          }
          const int best = (int)W::redmin(alive ? (unsigned)cj : (unsigned)PIPL_INF);
          if (lane == 0) red[125] = best;
        }
        G::cta_sync();
        pivj = red[125];
      } else pivj = pipl_cta_min(alive ? cj : PIPL_INF, red);
    }
  }
  if (!decided) {
    /* more positive entries than threads: candidate set in global memory (member[], cand[]) */
    for (int j = tid; j < nvar; j += T) L.member[j] = prow[j] > 0 ? 1 : 0;
    G::cta_sync();
    const int ncand0 = pipl_compact(L, red, 0);
    int ncand = ncand0;
    if (ncand == 0) { pipl_finish(L, PIP_ST_OK, 1); return; }
    PIPL_DBG(0, 1); PIPL_DBG(1, ncand0);
    int k = 0;
    while (ncand > 1) {
      if (ncand <= 32) {
        /* few candidates left: one warp finishes the walk with the candidates in its lanes
         * (column, pivot-row entry and Unit position in registers), no CTA barriers */
        ncand = pipl_compact(L, red, 0);
        if (tid < 32) {
          const int lane = tid;
          bool alive = lane < ncand;
          const int j = alive ? L.cand[lane] : 0;
          const pip_i64 pj = alive ? prow[j] : 1;
          const int u = alive ? L.colpos[j] : PIPL_INF;
          const bool was = alive;
          int n = ncand;
          while (n > 1) {
            /* skip the rows that are zero in every column still in play, PIPL_WIN rows per step */
            int pst = PIPL_INF;
            for (;;) {
              int wp[PIPL_WIN];
              pipl_window(sb, k, nl, nwords, wp);
              if (wp[0] == PIPL_INF) break;
              PIPL_DBG(4, 1);
              pip_i64 v[PIPL_WIN];
              #pragma unroll
              for (int b = 0; b < PIPL_WIN; b++)
                v[b] = (alive && wp[b] < u) ? pipl_row(L, PIP_LINK(L.fl[wp[b]]))[j] : 0;
              int first = PIPL_WIN;
              #pragma unroll
              for (int b = PIPL_WIN - 1; b >= 0; b--) if (v[b] != 0) first = b;
              first = (int)W::redmin((unsigned)first);
              if (first < PIPL_WIN) { pst = wp[first]; break; }
              if (wp[PIPL_WIN - 1] == PIPL_INF) break;
              if (pip_popc(W::ballot(alive && u > wp[PIPL_WIN - 1])) <= 1) break;     /* decided by the Unit positions */
              k = wp[PIPL_WIN - 1] + 1;
            }
            const unsigned mel = W::ballot(alive && u < pst);
            const int nel = pip_popc(mel);
            if (nel >= n) {
              const int umax = (int)W::redmax(alive ? (unsigned)u : 0u);
              alive = alive && u == umax;
              n = 1;
              break;
            }
            if (alive && u < pst) alive = false;
            n -= nel;
            if (pst >= nl || n <= 1) break;
            const pip_i64 *row = pipl_row(L, PIP_LINK(L.fl[pst]));
            const pip_i64 a = alive ? row[j] : 0;
            pip_i64 ba = a, bp = pj;
            int valid = alive ? 1 : 0;
            for (int o = 16; o > 0; o >>= 1) {
              const pip_i64 oa = W::shfl_xor64(ba, o), op = W::shfl_xor64(bp, o);
              const int ov = W::shfl_xor(valid, o);
              if (ov && (!valid || pipl_ratio_cmp(oa, op, ba, bp) < 0)) { ba = oa; bp = op; valid = 1; }
            }
            if (alive && pipl_ratio_cmp(a, pj, ba, bp) != 0) alive = false;
  #ifdef PIPL_WALK_STATS
            { const int n2 = pip_popc(W::ballot(alive));
              if (lane == 0) L.prof[7] += 1ull + ((ba == 0) ? (1ull << 20) : 0ull) + ((ba == 0 && n2 == n) ? (1ull << 40) : 0ull); }
            PIPL_DBG(5, 1);
  #endif
            n = pip_popc(W::ballot(alive));
            k = pst + 1;
          }
          if (was && !alive) L.member[j] = 0;
        }
        G::cta_sync();
        ncand = 1;
        break;
      }
      /* many candidates: the whole CTA walks, each thread keeping up to PIPL_K candidates (column,
       * pivot-row entry, Unit position) in registers so that one walk step costs one gather of the
       * stored row plus a few barriers */
      {
        int cj[PIPL_K], cu[PIPL_K];
        pip_i64 cp[PIPL_K];
        unsigned alive = 0;
        #pragma unroll
        for (int i = 0; i < PIPL_K; i++) {
          const int m = tid + i * T;
          cj[i] = 0; cu[i] = PIPL_INF; cp[i] = 1;
          if (m < ncand0 && L.member[L.cand[m]]) { cj[i] = L.cand[m]; cu[i] = L.colpos[cj[i]]; cp[i] = prow[cj[i]]; alive |= 1u << i; }
        }
        const bool fits = ncand0 <= PIPL_K * T;
        pip_i64 *red64 = (pip_i64 *)red;
        while (fits && ncand > 32) {
          /* Skip ahead to the first stored row that can discriminate: a row whose entries are zero in
           * every column still in play (97 % of the rows met on consecutive-ones tableaus) leaves the
           * candidate set alone, so the walk examines PIPL_WIN rows per step -- all gathers of a window
           * in flight together, one reduction -- and only stops at a row with a non-zero entry.
           * Candidates whose Unit position lies before a row are out of play at that row; they are
           * struck (lazily) by the cu < pst test below. */
          int pst = PIPL_INF;
          for (;;) {
            int wp[PIPL_WIN];
            pipl_window(sb, k, nl, nwords, wp);
            if (wp[0] == PIPL_INF) break;
            PIPL_DBG(2, 1);
            const pip_i64 *wr[PIPL_WIN];
            #pragma unroll
            for (int b = 0; b < PIPL_WIN; b++) wr[b] = wp[b] != PIPL_INF ? pipl_row(L, PIP_LINK(L.fl[wp[b]])) : prow;
            int first = PIPL_WIN;
            #pragma unroll
            for (int i = 0; i < PIPL_K; i++) {
              if (!((alive >> i) & 1u)) continue;
              pip_i64 v[PIPL_WIN];
              #pragma unroll
              for (int b = 0; b < PIPL_WIN; b++) v[b] = wp[b] < cu[i] ? wr[b][cj[i]] : 0;      /* PIPL_INF < cu never holds */
              #pragma unroll
              for (int b = PIPL_WIN - 1; b >= 0; b--) if (v[b] != 0 && b < first) first = b;
            }
            first = pipl_cta_min(first, red);
            if (first < PIPL_WIN) { pst = wp[first]; break; }
            if (wp[PIPL_WIN - 1] == PIPL_INF) break;               /* no stored row left */
            /* the Unit positions passed so far may already have decided the walk: with at most one
             * candidate still in play beyond this window the survivor is the one with the last Unit
             * position, which the pst = PIPL_INF case below picks */
            int inplay = 0;
            #pragma unroll
            for (int i = 0; i < PIPL_K; i++) if (((alive >> i) & 1u) && cu[i] > wp[PIPL_WIN - 1]) inplay++;
            inplay = pipl_cta_sum(inplay, red);
            if (inplay <= 1) break;
            k = wp[PIPL_WIN - 1] + 1;
          }
          int nel = 0, umax = -1;
          #pragma unroll
          for (int i = 0; i < PIPL_K; i++)
            if (((alive >> i) & 1u) && cu[i] < pst) { nel++; if (cu[i] > umax) umax = cu[i]; }
          nel = pipl_cta_sum(nel, red);
          if (nel >= ncand) {
            umax = pipl_cta_max(umax, red);
            #pragma unroll
            for (int i = 0; i < PIPL_K; i++) if (((alive >> i) & 1u) && cu[i] != umax) alive &= ~(1u << i);
            ncand = 1;
            break;
          }
          #pragma unroll
          for (int i = 0; i < PIPL_K; i++) if (((alive >> i) & 1u) && cu[i] < pst) alive &= ~(1u << i);
          ncand -= nel;
          if (pst >= nl || ncand <= 1) break;
          const pip_i64 *row = pipl_row(L, PIP_LINK(L.fl[pst]));
          pip_i64 va[PIPL_K];
          #pragma unroll
          for (int i = 0; i < PIPL_K; i++) va[i] = ((alive >> i) & 1u) ? row[cj[i]] : 0;
          pip_i64 ba = 0, bp = 1;
          int valid = 0;
          #pragma unroll
          for (int i = 0; i < PIPL_K; i++)
            if (((alive >> i) & 1u) && (!valid || pipl_ratio_cmp(va[i], cp[i], ba, bp) < 0)) { ba = va[i]; bp = cp[i]; valid = 1; }
          for (int o = 16; o > 0; o >>= 1) {
            const pip_i64 oa = W::shfl_xor64(ba, o), op = W::shfl_xor64(bp, o);
            const int ov = W::shfl_xor(valid, o);
            if (ov && (!valid || pipl_ratio_cmp(oa, op, ba, bp) < 0)) { ba = oa; bp = op; valid = 1; }
          }
          {
            const int lane = W::lane(), wid = tid >> 5, nw = (T + 31) >> 5;
            G::cta_sync();
            if (lane == 0) { red64[2 * wid] = ba; red64[2 * wid + 1] = bp; red[96 + wid] = valid; }
            G::cta_sync();
            valid = 0;
            for (int i = 0; i < nw; i++)
              if (red[96 + i] && (!valid || pipl_ratio_cmp(red64[2 * i], red64[2 * i + 1], ba, bp) < 0 || !valid)) {
                if (!valid || pipl_ratio_cmp(red64[2 * i], red64[2 * i + 1], ba, bp) < 0) { ba = red64[2 * i]; bp = red64[2 * i + 1]; }
                valid = 1;
              }
          }
          int removed = 0;
          #pragma unroll
          for (int i = 0; i < PIPL_K; i++)
            if (((alive >> i) & 1u) && pipl_ratio_cmp(va[i], cp[i], ba, bp) != 0) { alive &= ~(1u << i); removed++; }
          removed = pipl_cta_sum(removed, red);
          ncand -= removed;
          k = pst + 1;
  #ifdef PIPL_WALK_STATS
          if (tid == 0) L.prof[7] += 1ull + ((ba == 0) ? (1ull << 20) : 0ull) + ((ba == 0 && removed == 0) ? (1ull << 40) : 0ull);
          PIPL_DBG(3, 1); PIPL_DBG(6, ncand);
  #endif
        }
        /* publish the survivors */
        #pragma unroll
        for (int i = 0; i < PIPL_K; i++) {
          const int m = tid + i * T;
          if (m < ncand0 && L.member[L.cand[m]] && !((alive >> i) & 1u) && fits) L.member[L.cand[m]] = 0;
        }
        G::cta_sync();
        if (fits) continue;                       /* <= 32 left (warp path) or decided */
      }
      /* fallback for more than PIPL_K * T candidates: everything through global memory */
      /* next stored position >= k */
      int c = PIPL_INF;
      for (int w = (k >> 5) + tid; w < nwords; w += T) {
        unsigned bits = sb[w];
        if (w == (k >> 5)) bits &= ~0u << (k & 31);
        if (bits) { const int pp = (w << 5) + pip_ffs(bits) - 1; if (pp < nl && pp < c) c = pp; break; }
      }
      const int pst = pipl_cta_min(c, red);
      /* strike the members whose Unit position lies in [k, pst) */
      int nel = 0, umax = -1;
      for (int m = tid; m < ncand0; m += T) {
        const int j = L.cand[m];
        if (!L.member[j]) continue;
        const int u = L.colpos[j];
        if (u < pst) { nel++; if (u > umax) umax = u; }
      }
      nel = pipl_cta_sum(nel, red);
      if (nel >= ncand) {
        umax = pipl_cta_max(umax, red);
        for (int m = tid; m < ncand0; m += T) {
          const int j = L.cand[m];
          if (L.member[j] && L.colpos[j] != umax) L.member[j] = 0;
        }
        G::cta_sync();
        ncand = 1;
        break;
      }
      if (nel) {
        for (int m = tid; m < ncand0; m += T) {
          const int j = L.cand[m];
          if (L.member[j] && L.colpos[j] < pst) L.member[j] = 0;
        }
        G::cta_sync();
        ncand -= nel;
      }
      if (pst >= nl || ncand <= 1) break;
      /* keep the members with the minimal ratio at the stored row pst */
      const pip_i64 *row = pipl_row(L, PIP_LINK(L.fl[pst]));
      int best = -1;
      for (int m = tid; m < ncand0; m += T) {
        const int j = L.cand[m];
        if (!L.member[j]) continue;
        if (best < 0 || pipl_ratio_cmp(row[j], prow[j], row[best], prow[best]) < 0) best = j;
      }
      int winner = best;
      {
        const int lane = W::lane(), wid = tid >> 5, nw = (T + 31) >> 5;
        for (int o = 16; o > 0; o >>= 1) {
          const int other = W::shfl_down(winner, o);
          if (lane + o < 32 && other >= 0 && (winner < 0 || pipl_ratio_cmp(row[other], prow[other], row[winner], prow[winner]) < 0))
            winner = other;
        }
        G::cta_sync();
        if (lane == 0) red[wid] = winner;
        G::cta_sync();
        winner = -1;
        for (int i = 0; i < nw; i++) {
          const int o = red[i];
          if (o >= 0 && (winner < 0 || pipl_ratio_cmp(row[o], prow[o], row[winner], prow[winner]) < 0)) winner = o;
        }
      }
      int removed = 0;
      for (int m = tid; m < ncand0; m += T) {
        const int j = L.cand[m];
        if (L.member[j] && pipl_ratio_cmp(row[j], prow[j], row[winner], prow[winner]) != 0) { L.member[j] = 0; removed++; }
      }
      removed = pipl_cta_sum(removed, red);
      G::cta_sync();
      ncand -= removed;
      k = pst + 1;
    }
    /* the survivor (smallest column if, against the theory, several remain) */
    int pj = PIPL_INF;
    for (int j = tid; j < nvar; j += T) if (L.member[j] && j < pj) pj = j;
    pivj = pipl_cta_min(pj, red);

  }
  PIPL_ALAP(7);
  if (pivj >= nvar) { pipl_finish(L, PIP_ST_FAULT, 0); return; }

  PIPL_T(4);
  /* ---- determinant bookkeeping, source/traiter.c:394-447 (one thread) ------------------------- */
  if (tid == 0) {
    const pip_i64 pivot = prow[pivj], dpiv = L.den[pivi];
    int status = PIP_ST_OK;
    /* unit pivots on integer rows are the rule: gcd(x, +-1) = 1 and x / 1 = x need no 64-bit gcd or division */
    const bool unit = dpiv == 1 || pivot == 1 || pivot == -1;
    pip_i64 d = unit ? 1 : pip_gcd(pivot, dpiv);
    if (d == 0) status = PIP_ST_FAULT;
    else {
      pip_i64 ppivot = d == 1 ? pivot : pip_div(pivot, d), dppiv = d == 1 ? dpiv : pip_div(dpiv, d);
      int ldet = L.ctl[PIPL_LDET];
      pip_i64 *det = L.ctl64 + 2;
      for (int i = 0; dppiv != 1 && i < ldet && status == PIP_ST_OK; i++) {
        const pip_i64 g = pip_gcd(det[i], dppiv);
        if (g == 0) { status = PIP_ST_FAULT; break; }
        det[i] = pip_div(det[i], g);
        dppiv = pip_div(dppiv, g);
      }
      if (status == PIP_ST_OK && dppiv != 1) status = PIP_ST_FATAL + 1;
      if (status == PIP_ST_OK) {
        int i = 0;
        const int bp = pip_bitlen(ppivot);
        for (; i < ldet; i++)
          if (pip_bitlen(det[i]) + bp < 64) { det[i] = (pip_i64)((pip_u64)det[i] * (pip_u64)ppivot); break; }
        if (i >= ldet) {
          ldet++;
          if (ldet >= PIP_MAX_DET) status = PIP_ST_FATAL + 1;
          else det[i] = ppivot;
        }
        L.ctl[PIPL_LDET] = ldet;
      }
    }
    L.ctl64[0] = pivot; L.ctl64[1] = dpiv;
    L.ctl[PIPL_PIVI] = pivi; L.ctl[PIPL_PIVJ] = pivj;
    L.ctl[PIPL_STATUS] = status;
    L.ctl[PIPL_ACTION] = status == PIP_ST_OK ? PIPL_GO : PIPL_STOP;
    if (status == PIP_ST_OK) L.ctl[PIPL_PIVOTS] = L.ctl[PIPL_PIVOTS] + 1;
    L.ctl[PIPL_NACTIVE] = 0;
    L.ctl[PIPL_NEXT] = 0;
    L.ctl[PIPL_PUSHED] = 0;
  }
  G::cta_sync();
  PIPL_T(5);
  PIPL_T(6);
}

/* ---- phase C: rank-1 update of the stored rows (source/traiter.c:461-502), all CTAs --------------
 * Every CTA looks at its own stripe of positions (cta, cta + ncta, ...), one position per thread, and
 * keeps the rows whose update is not the identity (pivot-column entry != 0 or a denominator to
 * normalise) in a shared-memory list together with what the scan already loaded.  It updates up to
 * PIPL_KEEP of them itself, the whole CTA on one row; what is left over goes to a grid-wide overflow
 * queue that the CTAs drain once everybody has published (ctl[PUSHED] == ncta).  No single-CTA list
 * phase, no atomic per row on the common path. */

/* a group of warps of the CTA working on one row: group-local thread id, size, named barrier, scratch */
struct PiplGroup { int tid, T, bar; int *red; };
PIP_DEV void pipl_gsync(const PiplGroup &g) { G::group_sync(g.bar, g.T); }

PIP_DEV void pipl_update_row(const PipLarge &L, const PiplGroup &grp, int k, int f, pip_i64 foo, pip_i64 dk, const pip_i64 *prow,
                             pip_i64 pivot, pip_i64 dpiv, int pivj)
{
  const int tid = grp.tid, T = grp.T;
  int *red = grp.red;
  const int nvar = L.nvar, ncol = nvar + 1;
  pip_i64 *row = pipl_row(L, PIP_LINK(f));
  pip_i64 lpiv = pivot;
  if (foo == 0) lpiv = 1;
  else if (pivot != 1 && foo != 1 && foo != -1) {
    const pip_i64 d = pip_gcd(pivot, foo);
    if (d != 1) { lpiv = pip_div(pivot, d); foo = pip_div(foo, d); }
  }
  const pip_i64 newden = (pip_i64)((pip_u64)lpiv * (pip_u64)dk);
  const pip_i64 zp = (pip_i64)((pip_u64)dpiv * (pip_u64)foo);
  pip_u64 orz = 0;
  /* pass 1: z = row*lpiv - prow*foo, 16-byte accesses, the whole group on one row */
  const int pairs = ncol >> 1;
  pip_i64x2 *row2 = (pip_i64x2 *)row;
  const pip_i64x2 *prow2 = (const pip_i64x2 *)prow;
  #pragma unroll 4
  for (int q = tid; q < pairs; q += T) {
    const pip_i64x2 a = row2[q], b = prow2[q];
    pip_i64x2 z;
    z.x = (pip_i64)((pip_u64)a.x * (pip_u64)lpiv - (pip_u64)b.x * (pip_u64)foo);
    z.y = (pip_i64)((pip_u64)a.y * (pip_u64)lpiv - (pip_u64)b.y * (pip_u64)foo);
    if (2 * q == pivj) z.x = zp;
    if (2 * q + 1 == pivj) z.y = zp;
    row2[q] = z;
    orz |= (pip_u64)z.x | (pip_u64)z.y;
  }
  if ((ncol & 1) && tid == 0) {
    const int j = ncol - 1;
    pip_i64 z = (pip_i64)((pip_u64)row[j] * (pip_u64)lpiv - (pip_u64)prow[j] * (pip_u64)foo);
    if (j == pivj) z = zp;
    row[j] = z;
    orz |= (pip_u64)z;
  }
  pip_i64 g = newden;
  if (g != 1) {                        /* uniform over the group */
    if ((g & (g - 1)) == 0 && g > 0) {
      orz |= (pip_u64)g;
      unsigned lo = W::redor((unsigned)orz), hi = W::redor((unsigned)(orz >> 32));
      const int lane = W::lane(), wid = tid >> 5, nw = (T + 31) >> 5;
      pipl_gsync(grp);
      if (lane == 0) { red[2 * wid] = (int)lo; red[2 * wid + 1] = (int)hi; }
      pipl_gsync(grp);
      lo = 0; hi = 0;
      for (int i = 0; i < nw; i++) { lo |= (unsigned)red[2 * i]; hi |= (unsigned)red[2 * i + 1]; }
      const pip_u64 all = ((pip_u64)hi << 32) | lo;
      g = (pip_i64)(all & (0ull - all));
    } else {
      pipl_gsync(grp);                 /* pass 1 stores visible */
      for (int j = tid; j < ncol && g != 1; j += T) g = pip_gcd(g, row[j]);
      for (int o = 16; o > 0; o >>= 1) g = pip_gcd(g, W::shfl_xor64(g, o));
      const int lane = W::lane(), wid = tid >> 5, nw = (T + 31) >> 5;
      pip_i64 *red64 = (pip_i64 *)red;
      pipl_gsync(grp);
      if (lane == 0) red64[wid] = g;
      pipl_gsync(grp);
      g = red64[0];
      for (int i = 1; i < nw; i++) g = pip_gcd(g, red64[i]);
    }
  }
  pip_i64 nd = newden;
  if (g != 1 && g != 0) {
    pipl_gsync(grp);
    const PipExactDiv e = pip_exact_prepare(g);
    for (int j = tid; j < ncol; j += T) row[j] = pip_exact_apply(row[j], e);
    nd = pip_exact_apply(newden, e);
  }
  pipl_gsync(grp);
  if (tid == 0) {
    L.den[k] = nd;
    const pip_i64 c = row[nvar];
    L.csign[k] = c < 0 ? -1 : c > 0 ? 1 : 0;
    int ff = PIP_FLAG(f);
    const int fff = zp < 0 ? PIP_MINUS : zp == 0 ? PIP_ZERO : PIP_PLUS;
    if (fff != PIP_ZERO && fff != ff) {
      if (ff == PIP_ZERO) ff = (fff == PIP_MINUS ? PIP_UNKNOWN : fff);
      else ff = PIP_UNKNOWN;
      L.fl[k] = PIP_MKFL(ff, PIP_LINK(f));
    }
    if (g == 0) L.ctl[PIPL_STATUS] = PIP_ST_FAULT;
  }
}

/* The same update with the row staged in shared memory by the TMA engine (cp.async.bulk, one
 * instruction for the whole 32 KB row: every byte of the row is in flight at once, which no number of
 * per-thread 16-byte loads reaches at 124 registers per thread).  `srow` receives the row, `sprow` holds
 * the pivot row (staged once per pivot and CTA).  A row that needs no normalisation is written
 * straight to global memory; otherwise z stays in shared memory until its gcd is known, so every row
 * is read once and written once. */
PIP_DEV void pipl_update_row_staged(const PipLarge &L, const PiplGroup &grp, int k, int f, pip_i64 foo, pip_i64 dk,
                                    const pip_i64 *sprow, pip_i64 *srow, unsigned long long *mbar, unsigned &parity,
                                    pip_i64 pivot, pip_i64 dpiv, int pivj)
{
  const int tid = grp.tid, T = grp.T;
  int *red = grp.red;
  const int nvar = L.nvar, ncol = nvar + 1;
  pip_i64 *row = pipl_row(L, PIP_LINK(f));
  if (tid == 0) { G::proxy_fence(); G::bulk_g2s(srow, row, (unsigned)(L.stride * sizeof(pip_i64)), mbar); }
  pip_i64 lpiv = pivot;
  if (foo == 0) lpiv = 1;
  else if (pivot != 1 && foo != 1 && foo != -1) {
    const pip_i64 d = pip_gcd(pivot, foo);
    if (d != 1) { lpiv = pip_div(pivot, d); foo = pip_div(foo, d); }
  }
  const pip_i64 newden = (pip_i64)((pip_u64)lpiv * (pip_u64)dk);
  const pip_i64 zp = (pip_i64)((pip_u64)dpiv * (pip_u64)foo);
  const bool direct = newden == 1;             /* nothing to normalise: z goes straight to global memory */
  G::mbar_wait(mbar, parity);
  parity ^= 1u;
  pip_u64 orz = 0;
  const int pairs = L.stride >> 1;             /* the padding word of an odd row is computed too (0 * x - 0 * y) */
  pip_i64x2 *row2 = (pip_i64x2 *)row, *srow2 = (pip_i64x2 *)srow;
  const pip_i64x2 *sprow2 = (const pip_i64x2 *)sprow;
  #pragma unroll 4
  for (int q = tid; q < pairs; q += T) {
    const pip_i64x2 a = srow2[q], b = sprow2[q];
    pip_i64x2 z;
    z.x = (pip_i64)((pip_u64)a.x * (pip_u64)lpiv - (pip_u64)b.x * (pip_u64)foo);
    z.y = (pip_i64)((pip_u64)a.y * (pip_u64)lpiv - (pip_u64)b.y * (pip_u64)foo);
    if (2 * q == pivj) z.x = zp;
    if (2 * q + 1 == pivj) z.y = zp;
    if (2 * q >= ncol) z.x = a.x;
    if (2 * q + 1 >= ncol) z.y = a.y;
    if (direct) row2[q] = z; else srow2[q] = z;
    if ((nvar >> 1) == q) { const pip_i64 c = (nvar & 1) ? z.y : z.x; red[60] = c < 0 ? -1 : c > 0 ? 1 : 0; }
    if (2 * q < ncol) orz |= (pip_u64)z.x;
    if (2 * q + 1 < ncol) orz |= (pip_u64)z.y;
  }
  pip_i64 g = newden, nd = newden;
  if (!direct) {
    if ((g & (g - 1)) == 0 && g > 0) {
      orz |= (pip_u64)g;
      unsigned lo = W::redor((unsigned)orz), hi = W::redor((unsigned)(orz >> 32));
      const int lane = W::lane(), wid = tid >> 5, nw = (T + 31) >> 5;
      pipl_gsync(grp);
      if (lane == 0) { red[2 * wid] = (int)lo; red[2 * wid + 1] = (int)hi; }
      pipl_gsync(grp);
      lo = 0; hi = 0;
      for (int i = 0; i < nw; i++) { lo |= (unsigned)red[2 * i]; hi |= (unsigned)red[2 * i + 1]; }
      const pip_u64 all = ((pip_u64)hi << 32) | lo;
      g = (pip_i64)(all & (0ull - all));
    } else {
      pipl_gsync(grp);                 /* pass 1 stores visible */
      for (int j = tid; j < ncol && g != 1; j += T) g = pip_gcd(g, srow[j]);
      for (int o = 16; o > 0; o >>= 1) g = pip_gcd(g, W::shfl_xor64(g, o));
      const int lane = W::lane(), wid = tid >> 5, nw = (T + 31) >> 5;
      pip_i64 *red64 = (pip_i64 *)red;
      pipl_gsync(grp);
      if (lane == 0) red64[wid] = g;
      pipl_gsync(grp);
      g = red64[0];
      for (int i = 1; i < nw; i++) g = pip_gcd(g, red64[i]);
    }
    if (g != 1 && g != 0) {
      const PipExactDiv e = pip_exact_prepare(g);
      for (int q = tid; q < pairs; q += T) {
        pip_i64x2 z = srow2[q];
        if (2 * q < ncol) z.x = pip_exact_apply(z.x, e);
        if (2 * q + 1 < ncol) z.y = pip_exact_apply(z.y, e);
        row2[q] = z;
      }
      nd = pip_exact_apply(newden, e);
    } else {
      for (int q = tid; q < pairs; q += T) row2[q] = srow2[q];
    }
  }
  pipl_gsync(grp);                     /* red[60] and every read of the staging buffer are done */
  if (tid == 0) {
    L.den[k] = nd;
    L.csign[k] = (signed char)red[60];  /* the sign of the constant does not change with the division by g > 0 */
    int ff = PIP_FLAG(f);
    const int fff = zp < 0 ? PIP_MINUS : zp == 0 ? PIP_ZERO : PIP_PLUS;
    if (fff != PIP_ZERO && fff != ff) {
      if (ff == PIP_ZERO) ff = (fff == PIP_MINUS ? PIP_UNKNOWN : fff);
      else ff = PIP_UNKNOWN;
      L.fl[k] = PIP_MKFL(ff, PIP_LINK(f));
    }
    if (g == 0) L.ctl[PIPL_STATUS] = PIP_ST_FAULT;
  }
}

/* Short rows (at most 64 words: the tall, narrow tableaus that cut chains produce -- thousands of rows
 * of a few dozen columns): one WARP per row, each lane owns columns lane and lane + 32, the pivot row's
 * two entries stay in registers for the whole phase, the row gcd is a shuffle reduction.  No shared
 * memory, no barrier. */
PIP_DEV void pipl_update_row_warp(const PipLarge &L, int k, int f, pip_i64 foo, pip_i64 dk, pip_i64 b0, pip_i64 b1,
                                  pip_i64 pivot, pip_i64 dpiv, int pivj)
{
  const int lane = W::lane();
  const int nvar = L.nvar, ncol = nvar + 1;
  const int j0 = lane, j1 = lane + 32;
  pip_i64 *row = pipl_row(L, PIP_LINK(f));
  const pip_i64 a0 = j0 < ncol ? row[j0] : 0, a1 = j1 < ncol ? row[j1] : 0;
  pip_i64 lpiv = pivot;
  if (foo == 0) lpiv = 1;
  else if (pivot != 1 && foo != 1 && foo != -1) {
    const pip_i64 d = pip_gcd(pivot, foo);
    if (d != 1) { lpiv = pip_div(pivot, d); foo = pip_div(foo, d); }
  }
  const pip_i64 newden = (pip_i64)((pip_u64)lpiv * (pip_u64)dk);
  const pip_i64 zp = (pip_i64)((pip_u64)dpiv * (pip_u64)foo);
  pip_i64 z0 = (pip_i64)((pip_u64)a0 * (pip_u64)lpiv - (pip_u64)b0 * (pip_u64)foo);
  pip_i64 z1 = (pip_i64)((pip_u64)a1 * (pip_u64)lpiv - (pip_u64)b1 * (pip_u64)foo);
  if (j0 == pivj) z0 = zp;
  if (j1 == pivj) z1 = zp;
  if (j0 >= ncol) z0 = 0;
  if (j1 >= ncol) z1 = 0;
  pip_i64 g = newden, nd = newden;
  if (g != 1) {
    if ((g & (g - 1)) == 0 && g > 0) {
      const pip_u64 orz = (pip_u64)z0 | (pip_u64)z1 | (pip_u64)g;
      const pip_u64 all = ((pip_u64)W::redor((unsigned)(orz >> 32)) << 32) | W::redor((unsigned)orz);
      g = (pip_i64)(all & (0ull - all));
    } else {
      if (g != 1) g = pip_gcd(g, z0);
      if (g != 1) g = pip_gcd(g, z1);
      for (int o = 16; o > 0; o >>= 1) g = pip_gcd(g, W::shfl_xor64(g, o));
    }
    if (g != 1 && g != 0) {
      const PipExactDiv e = pip_exact_prepare(g);
      z0 = pip_exact_apply(z0, e); z1 = pip_exact_apply(z1, e);
      nd = pip_exact_apply(newden, e);
    }
  }
  if (j0 < ncol) row[j0] = z0;
  if (j1 < ncol) row[j1] = z1;
  const pip_i64 c = W::shfl64(nvar < 32 ? z0 : z1, nvar & 31);
  if (lane == 0) {
    L.den[k] = nd;
    L.csign[k] = c < 0 ? -1 : c > 0 ? 1 : 0;
    int ff = PIP_FLAG(f);
    const int fff = zp < 0 ? PIP_MINUS : zp == 0 ? PIP_ZERO : PIP_PLUS;
    if (fff != PIP_ZERO && fff != ff) {
      if (ff == PIP_ZERO) ff = (fff == PIP_MINUS ? PIP_UNKNOWN : fff);
      else ff = PIP_UNKNOWN;
      L.fl[k] = PIP_MKFL(ff, PIP_LINK(f));
    }
    if (g == 0) L.ctl[PIPL_STATUS] = PIP_ST_FAULT;
  }
}

PIP_DEV void pipl_phase_c(const PipLarge &L, int *red, pip_i64 *stage, unsigned &par_prow, unsigned &par_row)
{
  const int tid = G::tid(), T = G::T(), cta = G::cta(), ncta = G::ncta();
  const int nl = L.nvar + L.ctl[PIPL_NI];
  const int pivi = L.ctl[PIPL_PIVI], pivj = L.ctl[PIPL_PIVJ];
  const pip_i64 pivot = L.ctl64[0], dpiv = L.ctl64[1];
  const pip_i64 *prow = pipl_row(L, PIP_LINK(L.fl[pivi]));
  int *lcnt = red + 127, *lk = red + 128, *lf = red + 128 + PIPL_LCAP;
  pip_i64 *lfoo = (pip_i64 *)(red + 128 + 2 * PIPL_LCAP), *lden = lfoo + PIPL_LCAP;
  PIPL_CLAP_BEGIN;
  /* staged mode: mbarriers behind the group scratch, buffers = pivot row + one row per group */
  unsigned long long *mbar = (unsigned long long *)(red + 128 + 6 * PIPL_LCAP + 64 * PIPL_NG);
  if (tid == 0) {
    *lcnt = 0;
    if (stage) { G::proxy_fence(); G::bulk_g2s(stage, prow, (unsigned)(L.stride * sizeof(pip_i64)), mbar); }
  }
  G::cta_sync();
  /* scan the stripe: every load of a position is issued by its own thread, one round trip in all */
  int nskip = 0;
  for (int p = cta + tid * ncta; p < nl; p += T * ncta) {
    const int f = L.fl[p];
    if (p == pivi || (f & PIP_UNIT)) continue;
    const pip_i64 foo = pipl_row(L, PIP_LINK(f))[pivj];
    const pip_i64 dk = L.den[p];
    if (foo == 0 && dk == 1) { nskip++; continue; }
    const int at = (int)G::atomic_add_u((unsigned *)lcnt, 1u);
    if (at < PIPL_LCAP) { lk[at] = p; lf[at] = f; lfoo[at] = foo; lden[at] = dk; }
    else L.active[G::atomic_add_u((unsigned *)&L.ctl[PIPL_NACTIVE], 1u)] = p;     /* tall tableaus only: > 128 active rows per stripe */
  }
  nskip = pipl_cta_sum(nskip, red);                /* also orders the list stores before the reads */
  PIPL_CLAP(0);
  int nloc = *lcnt < PIPL_LCAP ? *lcnt : PIPL_LCAP;
  const bool skinny = L.stride <= 64;               /* short rows: a warp per row, nothing to share (below) */
  const int keep = (skinny || nloc < PIPL_KEEP) ? nloc : PIPL_KEEP;
  /* publish the left-overs, then tell the grid this CTA has nothing more to add */
  if (nloc > keep) {
    if (tid == 0) red[63] = (int)G::atomic_add_u((unsigned *)&L.ctl[PIPL_NACTIVE], (unsigned)(nloc - keep));
    G::cta_sync();
    const int at = red[63];
    for (int i = keep + tid; i < nloc; i += T) L.active[at + i - keep] = lk[i];
  }
  G::cta_sync();
  if (tid == 0) {
    if (nskip) G::atomic_add_u((unsigned *)&L.ctl[PIPL_SKIPPED_LO], (unsigned)nskip);
    G::fence();
    G::atomic_add_u((unsigned *)&L.ctl[PIPL_PUSHED], 1u);
  }
  PIPL_CLAP(1);
  if (skinny) {
    const int lane = W::lane(), wid = tid >> 5, nw = (T + 31) >> 5;
    const int ncol = L.nvar + 1;
    const pip_i64 b0 = lane < ncol ? prow[lane] : 0, b1 = lane + 32 < ncol ? prow[lane + 32] : 0;
    if (stage) { G::mbar_wait(mbar, par_prow); par_prow ^= 1u; }      /* the staging of the pivot row was issued: drain it */
    for (int i = wid; i < keep; i += nw) pipl_update_row_warp(L, lk[i], lf[i], lfoo[i], lden[i], b0, b1, pivot, dpiv, pivj);
    /* rows beyond the local list's capacity (stripes with more than PIPL_LCAP active rows) */
    int nactive = 0;
    if (lane == 0) {
      while ((int)G::atomic_add_u((unsigned *)&L.ctl[PIPL_PUSHED], 0u) < ncta) G::relax();
      nactive = (int)G::atomic_add_u((unsigned *)&L.ctl[PIPL_NACTIVE], 0u);
    }
    nactive = W::shfl(nactive, 0);
    G::fence();
    while (nactive > 0) {
      int idx = 0;
      if (lane == 0) idx = (int)G::atomic_add_u((unsigned *)&L.ctl[PIPL_NEXT], 1u);
      idx = W::shfl(idx, 0);
      if (idx >= nactive) break;
      const int k = G::load_int(&L.active[idx]);
      const int f = L.fl[k];
      const pip_i64 foo = pipl_row(L, PIP_LINK(f))[pivj];
      const pip_i64 dk = L.den[k];
      W::sync();
      pipl_update_row_warp(L, k, f, foo, dk, b0, b1, pivot, dpiv, pivj);
    }
    G::cta_sync();
    PIPL_CLAP(4);
    return;
  }
  /* the CTA splits into PIPL_NG groups of warps, one row per group at a time (a row is 32 KB: its
   * update is one load round trip, so rows in flight are what fills the memory pipes) */
  const int ng = T >= 32 * PIPL_NG ? PIPL_NG : 1;
  PiplGroup grp;
  grp.T = T / ng;
  const int gid = tid / grp.T;
  grp.tid = tid - gid * grp.T; grp.bar = 1 + gid; grp.red = red + 128 + 6 * PIPL_LCAP + 64 * gid;
  pip_i64 *srow = stage ? stage + (pip_i64)(1 + gid) * L.stride : nullptr;
  if (stage) { G::mbar_wait(mbar, par_prow); par_prow ^= 1u; }
  /* own rows */
  for (int i = gid; i < keep; i += ng) {
    if (stage) pipl_update_row_staged(L, grp, lk[i], lf[i], lfoo[i], lden[i], stage, srow, mbar + 1 + gid, par_row, pivot, dpiv, pivj);
    else pipl_update_row(L, grp, lk[i], lf[i], lfoo[i], lden[i], prow, pivot, dpiv, pivj);
    pipl_gsync(grp);
  }
  PIPL_CLAP(2);
  /* overflow queue, once every CTA has published */
  if (grp.tid == 0) {
    while ((int)G::atomic_add_u((unsigned *)&L.ctl[PIPL_PUSHED], 0u) < ncta) G::relax();
    grp.red[62] = (int)G::atomic_add_u((unsigned *)&L.ctl[PIPL_NACTIVE], 0u);
  }
  pipl_gsync(grp);
  G::fence();
  PIPL_CLAP(3);
  const int nactive = grp.red[62];
  for (;;) {
    pipl_gsync(grp);
    if (grp.tid == 0) grp.red[63] = (int)G::atomic_add_u((unsigned *)&L.ctl[PIPL_NEXT], 1u);
    pipl_gsync(grp);
    const int idx = grp.red[63];
    if (idx >= nactive) break;
    const int k = G::load_int(&L.active[idx]);
    const int f = L.fl[k];
    const pip_i64 foo = pipl_row(L, PIP_LINK(f))[pivj];
    const pip_i64 dk = L.den[k];
    pipl_gsync(grp);                    /* every thread has read foo before pass 1 overwrites row[pivj] */
    if (stage) pipl_update_row_staged(L, grp, k, f, foo, dk, stage, srow, mbar + 1 + gid, par_row, pivot, dpiv, pivj);
    else pipl_update_row(L, grp, k, f, foo, dk, prow, pivot, dpiv, pivj);
  }
  G::cta_sync();
  PIPL_CLAP(4);
}

/* ---- the whole solve (called by every thread of the cooperative grid) ------------------------ */
PIP_DEV void pipl_init_rows(const PipLarge &L, float *sz);
PIP_DEV void pipl_sort(const PipLarge &L, float *sz, int *red);

PIP_DEV void pipl_solve(const PipLarge &L, int *red, pip_i64 *stage = nullptr)
{
  unsigned par_prow = 0, par_row = 0;
  bool first = true;
  float *sz = (float *)L.cand;                 /* scratch: pcap floats fit (cand has pcap ints) */
  if (stage) {
    unsigned long long *mbar = (unsigned long long *)(red + 128 + 6 * PIPL_LCAP + 64 * PIPL_NG);
    if (G::tid() == 0) { for (int i = 0; i <= PIPL_NG; i++) G::mbar_init(mbar + i, 1); G::proxy_fence(); }
    G::cta_sync();
  }
  pipl_init_rows(L, sz);
  G::grid_sync();
  if (G::cta() == 0) pipl_sort(L, sz, red);
  long long t0 = pip_clock();
  for (;;) {
    if (G::cta() == 0) pipl_phase_ab(L, red, first, stage);
    first = false;
    G::grid_sync();
    const long long t1 = pip_clock();
    if (L.ctl[PIPL_ACTION] == PIPL_STOP) break;
    pipl_phase_c(L, red, stage, par_prow, par_row);
    G::grid_sync();
    const long long t2 = pip_clock();
    if (G::cta() == 0 && G::tid() == 0) { L.prof[0] += (unsigned long long)(t1 - t0); L.prof[1] += (unsigned long long)(t2 - t1); }
    t0 = t2;
    if (L.ctl[PIPL_STATUS] != PIP_ST_OK) break;
  }
}

/* entry of traiter for the large problem: flags, tab_simplify (source/tab.c:396-427) and the
 * row "size" of tab_sort_rows, one warp per row across the whole grid */
PIP_DEV void pipl_init_rows(const PipLarge &L, float *sz)
{
  const int lane = W::lane();
  const int wpc = G::T() >> 5;
  const int gw = G::cta() * wpc + (G::tid() >> 5), nw = G::ncta() * wpc;
  const int nvar = L.nvar, ni = L.ctl[PIPL_NI], ncol = nvar + 1;
  for (int k = gw; k < nvar + ni; k += nw) {
    if (k < nvar) { if (lane == 0) { L.fl[k] = PIP_MKFL(PIP_UNIT, k); L.den[k] = 1; L.csign[k] = 0; sz[k] = 0.f; L.colpos[k] = k; } continue; }
    pip_i64 *row = pipl_row(L, k - nvar);
    if (L.flags & PIP_F_INT) {
      pip_i64 g = 0;
      for (int j = lane; j < nvar && g != 1; j += 32) g = pip_gcd(g, row[j]);
      for (int o = 16; o > 0; o >>= 1) g = pip_gcd(g, W::shfl_xor64(g, o));
      if (g != 0 && g != 1) {
        for (int j = lane; j < ncol; j += 32) row[j] = (j == nvar) ? pip_floor_q(row[j], g) : pip_div(row[j], g);
        W::sync();
      }
    }
    unsigned s = 0;
    for (int j = lane; j < nvar; j += 32) {
      const pip_u64 u = pip_uabs(row[j]);
      if (u < 2147483648ull && (unsigned)u > s) s = (unsigned)u;
    }
    s = W::redmax(s);
    if (lane == 0) {
      const pip_i64 c = row[nvar];
      L.fl[k] = PIP_MKFL(PIP_UNKNOWN, k - nvar); L.den[k] = 1;
      L.csign[k] = c < 0 ? -1 : c > 0 ? 1 : 0;
      sz[k] = (float)(double)s;
    }
  }
}

/* tab_sort_rows_xx (source/traiter.c:591-614) by CTA 0: selection sort of the position records by
 * first minimum strictly below the maximum; nothing to do when no size is below the maximum */
PIP_DEV void pipl_sort(const PipLarge &L, float *sz, int *red)
{
  const int tid = G::tid(), T = G::T();
  const int nvar = L.nvar, nl = nvar + L.ctl[PIPL_NI];
  int mx = 0;
  for (int k = nvar + tid; k < nl; k += T) { const int b = (int)pip_f2u(sz[k]); if (b > mx) mx = b; }
  const int smax_bits = pipl_cta_max(mx, red);       /* sizes are non-negative floats: bit order = value order */
  const double smax = (double)pip_u2f((unsigned)smax_bits);
  /* the stored-position bitmap (the sort only permutes stored rows among stored positions) */
  for (int w = tid; w < (L.pcap + 31) / 32; w += T) {
    unsigned bits = 0;
    for (int b = 0; b < 32; b++) { const int pp = w * 32 + b; if (pp >= nvar && pp < nl) bits |= 1u << b; }
    L.sbits[w] = bits;
  }
  G::cta_sync();
  for (int i = nvar; i < nl; i++) {
    int best = PIPL_INF, bestk = PIPL_INF;
    for (int k = i + tid; k < nl; k += T) {
      const float s = sz[k];
      if ((double)s < smax) {
        const int b = (int)pip_f2u(s);
        if (b < best) { best = b; bestk = k; }        /* ascending k per thread keeps the first */
      }
    }
    const int m = pipl_cta_min(best, red);
    if (m == PIPL_INF) break;                          /* no candidate left for any later i */
    const int src = pipl_cta_min(best == m ? bestk : PIPL_INF, red);
    if (src != i && tid == 0) {
      const int f = L.fl[i]; L.fl[i] = L.fl[src]; L.fl[src] = f;
      const pip_i64 d = L.den[i]; L.den[i] = L.den[src]; L.den[src] = d;
      const signed char c = L.csign[i]; L.csign[i] = L.csign[src]; L.csign[src] = c;
      const float s = sz[i]; sz[i] = sz[src]; sz[src] = s;
    }
    G::cta_sync();
  }
}

#endif
