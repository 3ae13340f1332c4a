/* Warp-per-problem parametric dual simplex + Gomory cuts (the PipLib solver core), sm_100a.
 *
 * One warp owns one problem.  The problem's whole working set -- tableau, context, the tableau
 * of the current compatibility sub-solve, cut scratch -- lives in a per-warp arena (shared
 * memory for size class S, global memory for class G).  Lanes map to row *positions*: each
 * lane updates its own row during a pivot (the running-gcd early-out of the reference stays a
 * per-lane scalar loop), and every order-dependent scan (first negative row, lexicographic
 * column choice, sign classification) is one ballot + find-first-set per 32 positions, so the
 * reference's position order is preserved bit for bit.
 *
 * The reference's recursion (source/traiter.c:628-791: traiter -> compa_test -> traiter and
 * traiter -> traiter at a split) is flattened into one state machine:
 *   - a compatibility sub-solve is a "call" that switches the current tableau descriptor to
 *     the sub tableau and returns to one of three sites (context check, +test, -test);
 *   - a split pushes the ELSE continuation (tableau + context snapshot) on a per-warp frame
 *     stack in global memory and carries on with the THEN branch in place; a finished branch
 *     pops the stack.  The solution cells come out in the reference's pre-order.
 *
 * Tableau model: position k has a packed word fl[k] = flag | link << 8,
 * link = owned column for a Unit position, storage slot otherwise; den[k] is the row's common
 * denominator.  Exactly ni slots are live at any time ([0, ni)); a pivot hands the pivot row's
 * slot to the Unit position that owned the pivot column (source/traiter.c:503-516).
 *
 * Every function cites the reference lines it restates; none of it is derived from the
 * reference's code structure (no malloc'd Tableau, no pointer-swapped rows, no recursion).
 */
#ifndef PIP_SOLVER_H
#define PIP_SOLVER_H

#include "pip_arith.h"
#include "pip_decode.h"
#include "pip_types.h"
#include "simt.h"


struct PipStats {
  unsigned pivots, cuts, subsolves, splits, max_rows, max_cols;
  unsigned wrapped;              /* int64 classes: some exact product left 64 bits (per lane, OR-ed) */
  unsigned long long elem_updates;
#ifdef PIP_PROFILE
  long long lap;                 /* clock64 of the last phase boundary */
  unsigned long long cyc[PIP_NPHASE];
#endif
};

/* phase accounting (only in the -DPIP_PROFILE build, libpiplib_dp_prof.so): a lap timer, every
 * PIP_LAP(st, phase) charges the cycles since the previous lap to `phase` */
#ifdef PIP_PROFILE
#define PIP_LAP(st, ph) do { long long n_ = clock64(); (st).cyc[ph] += (unsigned long long)(n_ - (st).lap); (st).lap = n_; } while (0)
/* sub-phases of the update as seen by one thread (charged on top of PIP_PH_UPDATE) */
#define PIP_ULAP(stp, ph) do { if (stp) { long long n_ = clock64(); (stp)->cyc[ph] += (unsigned long long)(n_ - ulap); ulap = n_; } } while (0)
#define PIP_ULAP_BEGIN(stp) long long ulap = clock64()
#else
#define PIP_LAP(st, ph) do { } while (0)
#define PIP_ULAP(stp, ph) do { } while (0)
#define PIP_ULAP_BEGIN(stp) do { } while (0)
#endif

/* capacity slack per level: {new parameters, main cut rows, extra context rows, sub cut rows} */
#ifndef PIP_WIDE_DP
#define PIP_WIDE_DP 6
#define PIP_WIDE_DR 32
#define PIP_WIDE_DX 24
#define PIP_WIDE_DS 24
#endif
PIP_HD void pip_slack(int level, int &dp, int &dr, int &dx, int &ds)
{
  if (level == PIP_LEVEL_S_WIDE) { dp = PIP_WIDE_DP; dr = PIP_WIDE_DR; dx = PIP_WIDE_DX; ds = PIP_WIDE_DS; }   /* int32 shared class: spare shared memory */
  else if (level >= 3) {
    const int k = level - 3 > 5 ? 5 : level - 3;      /* classes G3..G8 grow geometrically */
    dp = 6 << k; dr = 64 << (2 * k); dx = 24 << (2 * k); ds = 24 << (2 * k);
  }
  else if (level == 2) { dp = 3; dr = 12; dx = 10; ds = 10; }
  else if (level == 1) { dp = 2; dr = 6; dx = 6; ds = 6; }
  else { dp = 1; dr = 2; dx = 3; ds = 3; }
}

/* Carve the arena for one problem.  Returns false when even this slack level does not fit. */
PIP_HD bool pip_layout(int nvar, int nparm, int ni, int nc, int flags, int level, int words, int vbytes, PipLayout &L)
{
#define PIP_W(n) ((int)(((long long)(n) * vbytes + 7) / 8))   /* words holding n stored values */
  int dp, dr, dx, ds;
  pip_slack(level, dp, dr, dx, ds);
  const bool integer = (flags & PIP_F_INT) != 0;
  const bool ctxful = nparm > 0 || nc > 0;
  if (!(integer && nparm > 0)) dp = 0;
  if (!integer) dr = 0;
  if (!ctxful) { dx = 0; ds = 0; }
  int C = (nvar + 1 + nparm + dp) | 1;
  int R = ni + dr, Pm = nvar + R;
  int XC = (nparm + dp + 1) | 1;
  int XR = ctxful ? nc + dx + 2 * dp : 0;
  int SR = ctxful ? XR + 1 + ds : 0;
  int SP = ctxful ? nparm + dp + SR : 0;
  int o = 0;
  L.m.det = o; o += 4;
  L.s.det = o; o += 4;
  L.cut = o; o += PIP_W(C + 3);
  L.m.den = o; o += PIP_W(Pm);
  L.m.fl = o; o += (Pm + 1) / 2;
  L.tmp = o; o += ((Pm > SP ? Pm : SP) + 1) / 2;
  L.m.data = o; o += PIP_W(R * C);
  L.ctx = o; o += PIP_W(XR * XC);
  L.s.den = o; o += PIP_W(SP);
  L.s.fl = o; o += (SP + 1) / 2;
  L.s.data = o; o += PIP_W(SR * XC);
  if (flags & PIP_F_DUAL) o += (R + 1) / 2;     /* Compute_dual: pos[] of the entry sort, the last (R+1)/2 words */
  L.total = o;
#undef PIP_W
  L.m.stride = C; L.m.pcap = Pm; L.m.rcap = R;
  L.m.nvar = nvar; L.m.nparm = nparm; L.m.ni = ni; L.m.ldet = 1;
  L.s.stride = XC; L.s.pcap = SP; L.s.rcap = SR;
  L.s.nvar = nparm; L.s.nparm = 0; L.s.ni = 0; L.s.ldet = 1;
  L.cstride = XC; L.crcap = XR;
  return o <= words;
}

/* value-type traits: the solver is instantiated for int64 (the reference's arithmetic, wrapping)
 * and for int32 storage with exact 64-bit intermediates (typical polyhedral problems never leave
 * 31 bits; a value that would is flagged and the problem is re-run by the int64 instantiation, so
 * the answers are identical by construction) */
template <class V> struct PipVal;
template <> struct PipVal<pip_i64> {
  enum { bytes = 8, narrow = 0 };
  /* The value is the reference's: the low 64 bits of the products (source/traiter.c:483-485 wraps silently;
   * its verdict comes later from the determinant check).  `wrapped` additionally records whether the exact
   * 128-bit result left int64 -- the immediate "true overflow" flag of SURVEY.md 8 P3, reported per problem
   * (PIP_RES_WRAPPED) next to the reference's own verdict, never instead of it. */
  PIP_DM static pip_i64 mulsub(pip_i64 a, pip_i64 l, pip_i64 b, pip_i64 f, unsigned &wrapped)
  {
    const pip_u64 lo1 = (pip_u64)a * (pip_u64)l, lo2 = (pip_u64)b * (pip_u64)f;
    const pip_u64 dlo = lo1 - lo2;
    const pip_i64 dhi = pip_mulhi(a, l) - pip_mulhi(b, f) - (pip_i64)(lo1 < lo2);
    wrapped |= (unsigned)(dhi != ((pip_i64)dlo >> 63));
    return (pip_i64)dlo;
  }
  PIP_DM static pip_i64 mul(pip_i64 a, pip_i64 b, unsigned &wrapped)
  {
    const pip_i64 lo = (pip_i64)((pip_u64)a * (pip_u64)b);
    wrapped |= (unsigned)(pip_mulhi(a, b) != (lo >> 63));
    return lo;
  }
  PIP_HDM static pip_i64 cross(pip_i64 p, pip_i64 a, pip_i64 b, pip_i64 f)
  { return (pip_i64)((pip_u64)p * (pip_u64)a - (pip_u64)b * (pip_u64)f); }
  PIP_HDM static pip_i64 store(pip_i64 v, unsigned &) { return v; }
};
template <> struct PipVal<int> {
  enum { bytes = 4, narrow = 1 };
  /* values are kept inside (-2^31+16, 2^31-16) so that the +-1 adjustments of the algorithm stay exact */
  PIP_HDM static int store(pip_i64 z, unsigned &ovf)
  { ovf |= (unsigned)(((pip_u64)(z + 0x7ffffff0LL)) > 0xffffffe0ULL); return (int)z; }
  PIP_HDM static int mulsub(int a, int l, int b, int f, unsigned &ovf)
  { return store((pip_i64)a * l - (pip_i64)b * f, ovf); }
  PIP_HDM static int mul(int a, int b, unsigned &ovf) { return store((pip_i64)a * b, ovf); }
  PIP_HDM static pip_i64 cross(int p, int a, int b, int f) { return (pip_i64)p * a - (pip_i64)b * f; }
};

/* Team mode (size class M, tall tableaus): one CTA per problem.  Warp 0 runs the solver's state
 * machine exactly as in the warp-per-problem classes; for the rank-1 update of a pivot -- the only
 * phase whose cost grows with rows x columns -- it posts the pivot in this block and every warp of
 * the CTA updates its share of the rows (thread = row position), between two named barriers. */
struct PipTeam {
  int cmd;                       /* PIP_TEAM_* */
  int nthreads;
  int pivi, pivj;
  pip_i64 pivot, dpiv;
  PipTab T;
  pip_i64 *B;
  unsigned ovf;
  int fault;
  /* position scans (PIP_TEAM_FIRST / PIP_TEAM_EXAM): predicate kind + argument, range, result */
  int op_kind, op_arg, op_from, op_n;
  int result;
};
enum { PIP_TEAM_UPDATE = 0, PIP_TEAM_EXIT = 1, PIP_TEAM_FIRST = 2, PIP_TEAM_EXAM = 3,
       PIP_TEAM_MIN_ROWS = 64,      /* fewer positions: the leader warp updates alone */
       PIP_TEAM_MIN_SCAN = 128 };   /* fewer positions: the leader warp scans alone */
#if defined(__CUDACC__) && !defined(PIP_EMU)
PIP_DEV void pip_team_barrier(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }
PIP_DEV void pip_team_min(int *p, int v) { atomicMin(p, v); }
#else
PIP_DEV void pip_team_barrier(int) {}
PIP_DEV void pip_team_min(int *p, int v) { if (v < *p) *p = v; }
#endif

/* TEAM = the global-memory code path (classes G and M): blocked row walks, and -- when a PipTeam
 * block is passed -- CTA-wide update and scans.  false = shared-memory classes (and the emulator). */
/* STEAL = the instantiation with subtree donation compiled in (PipSteal): a separate kernel, because every
 * instruction added to the bulk kernel costs it throughput (instruction cache, section 4 of DESIGN.md) */
/* WORDS = every problem of the launch writes its serialised quast itself (PipLaunch::emit_words, word mode): the
 * cell emitters are not compiled in, which makes the executed code denser (section 4 of DESIGN.md) */
template <class V, bool TEAM = false, bool STEAL = false, bool WORDS = false>
struct PipSolver {
/* ---- small accessors ------------------------------------------------------------------- */
PIP_SDEV int *pip_fl(pip_i64 *B, const PipTab &T) { return (int *)(B + T.fl); }
PIP_SDEV V *pip_den(pip_i64 *B, const PipTab &T) { return (V *)(B + T.den); }
PIP_SDEV V *pip_row(pip_i64 *B, const PipTab &T, int slot) { return (V *)(B + T.data) + slot * T.stride; }

/* chercher_xx, source/traiter.c:39-44: first position in [from, n) whose flag meets the mask */
PIP_SDEVNI int pip_first_flag_impl(const int *fl, int mask, int from, int n)
{
  const int lane = W::lane();
  #pragma unroll 1
  for (int base = from & ~31; base < n; base += 32) {
    int k = base + lane;
    bool p = (k >= from && k < n) && ((fl[k] & mask) != 0);
    unsigned m = W::ballot(p);
    if (m) return base + pip_ffs(m) - 1;
  }
  return n;
}

PIP_SDEV int pip_first_flag(pip_i64 *B, const PipTab &T, int mask, int from, int n)
{
  return pip_first_flag_impl(pip_fl(B, T), mask, from, n);
}


/* warp-cooperative 2-D copy (rows x cols words) between arbitrary strides; one out-of-line copy
 * of this loop serves problem load, sub-tableau construction and the frame stack */
PIP_SDEVNI void pip_copy2d(V *dst, int dstride, const V *src, int sstride, int rows, int cols)
{
  if (cols <= 0) return;
  /* lane walks the (row, column) grid in steps of 32 elements: one division up front, then a
   * constant-time carry per step */
  const int lane = W::lane();
  int r = lane / cols, j = lane - r * cols;
  const int dr = 32 / cols, dj = 32 - dr * cols;
  #pragma unroll 1
  while (r < rows) {
    dst[r * dstride + j] = src[r * sstride + j];
    r += dr; j += dj;
    if (j >= cols) { j -= cols; r++; }
  }
}

/* warp-cooperative copy of n 64-bit words (the frame stack, arena images, the sub-tableau).  (Telling the
 * compiler which side is shared memory -- __builtin_assume(__isShared(..)) -- shortens the loop from 15 to 7-12
 * instructions per word but needs one copy of the function per direction: measured neutral, not kept.) */
PIP_SDEVNI void pip_copy_words(pip_i64 *dst, const pip_i64 *src, int n)
{
  #pragma unroll 1
  for (int k = W::lane(); k < n; k += 32) dst[k] = src[k];
}
/* ... out of a frame in global memory, past the L1: the frame may have been written by another SM (donation) */
PIP_SDEVNI void pip_copy_frame_in(pip_i64 *dst, const pip_i64 *src, int n)
{
  #pragma unroll 1
  for (int k = W::lane(); k < n; k += 32) dst[k] = W::load_cg(src + k);
}

/* a claimed frame is read by the thief straight out of this warp's stack: before the donor leaves the frame
 * behind for good (next problem, next stolen subtree on the same stack) the thief must have copied it */
PIP_SDEVNI void pip_offer_wait_copied(const PipSteal &S, int idx)
{
  if (S.mode != 1) return;
  for (;;) {
    int stt = 0;
    if (W::lane() == 0) stt = W::load_volatile(&S.offers[idx].state);
    if (W::shfl(stt, 0) != PIP_OFFER_CLAIMED) return;
    W::nap();
  }
}

/* subtree donation, donor side (PipSteal, pip_types.h): publish the bottom frame of the stack -- the ELSE
 * branch of the outermost open split -- when some warp is idle; one outstanding offer per segment */
PIP_SDEVNI void pip_offer_bottom(const PipSteal &S, int problem, int seg, pip_i64 *stk, pip_i64 top, pip_i64 &top_base,
                                 int &my_offer)
{
  if (!S.mode) return;
  const int lane = W::lane();
  if (my_offer >= 0) {
    /* an offer that was claimed: its frame left this stack for good, the next one up is the new bottom */
    int stt = 0;
    if (lane == 0) stt = W::load_volatile(&S.offers[my_offer].state);
    stt = W::shfl(stt, 0);
    if (stt == PIP_OFFER_OPEN) return;
    pip_offer_wait_copied(S, my_offer);
    top_base += stk[top_base + 6];
    my_offer = -1;
  }
  if (top <= top_base) return;
  if (S.mode == 1) {
    unsigned idle = 0;
    if (lane == 0) idle = W::load_volatile(&S.ctl[PIP_STL_IDLE]);
    if (!W::shfl((int)idle, 0)) return;
  }
  unsigned idx = 0;
  if (lane == 0) idx = W::atomic_add(&S.ctl[PIP_STL_OFFERS], 1u);
  idx = (unsigned)W::shfl((int)idx, 0);
  if (idx >= (unsigned)S.cap) return;
  W::fence();                       /* every lane's share of the frame is visible before the offer is */
  W::sync();
  if (lane == 0) {
    PipOffer &o = S.offers[idx];
    o.problem = problem; o.parent_seg = seg; o.frame = stk + top_base; o.pad = 0;
    S.seg_next[idx] = -1;
    W::fence();
    W::atomic_exch(&o.state, S.mode == 2 ? PIP_OFFER_CLAIMED : PIP_OFFER_OPEN);
  }
  W::sync();
  my_offer = (int)idx;
  if (S.mode == 2) { top_base += stk[top_base + 6]; my_offer = -1; }     /* test mode: given away at once */
}

/* problem load: like pip_copy2d but the source elements are int8 / int32 / int64 (the host ships
 * the narrowest width that holds the whole batch) */
PIP_SDEVNI unsigned pip_load2d(V *dst, int dstride, const void *src, int elem_log2, pip_i64 src_off, int rows, int cols)
{
  unsigned ovf = 0;
  if (cols <= 0) return 0;
  const int lane = W::lane();
  int r = lane / cols, j = lane - r * cols;
  const int dr = 32 / cols, dj = 32 - dr * cols;
  #pragma unroll 1
  while (r < rows) {
    const pip_i64 k = src_off + (pip_i64)r * cols + j;
    pip_i64 v;
    if (elem_log2 == 3) v = ((const pip_i64 *)src)[k];
    else if (elem_log2 == 2) v = ((const int *)src)[k];
    else v = ((const signed char *)src)[k];
    dst[r * dstride + j] = PipVal<V>::store(v, ovf);
    r += dr; j += dj;
    if (j >= cols) { j -= cols; r++; }
  }
  return ovf;
}

/* one term of the row "size" with a non-unit denominator: |(int)(double(v)/double(d))| or 0
 * when the conversion is out of range (source/traiter.c:580-584) */
PIP_SDEVNI unsigned pip_size_term(pip_i64 v, pip_i64 d)
{
  double t = pip_ll2d(v) / pip_ll2d(d);
  double a = t < 0 ? -t : t;
  if (a < 2147483648.0) return (unsigned)(int)a;
  return 0u;
}

/* tab_simplify_xx, source/tab.c:396-427 on `rows` rows of `width` words (lane = row) */
PIP_SDEVNI bool pip_simplify_rows(V *base, int rows, int stride, int width, int cst)
{
  bool fault = false;
  #pragma unroll 1
  for (int r = W::lane(); r < rows; r += 32) {
    V *row = base + r * stride;
    V g = 0;
    #pragma unroll 1
    for (int j = 0; j < width; j++) {
      if (j == cst) continue;
      g = pip_gcd(g, row[j]);
      if (g == 1) break;
    }
    if (g == 0 || g == 1) continue;
    #pragma unroll 1
    for (int j = 0; j < width; j++)
      row[j] = (j == cst) ? pip_floor_q(row[j], g) : pip_div(row[j], g);
  }
  return fault;
}

/* tab_sort_rows_xx, source/traiter.c:556-623.  size = max_j |(int)(double(T[i][j])/double(d))|
 * stored as float; an out-of-range (int) conversion is INT_MIN on the reference's x86-64 and
 * never raises the maximum.  Selection sort by first minimum strictly below smax. */
PIP_SDEV void pip_sort_rows(pip_i64 *B, const PipTab &T, int tmpoff)
{
  const int lane = W::lane();
  const int nl = T.nvar + T.ni;
  int *fl = pip_fl(B, T);
  V *den = pip_den(B, T);
  float *sz = (float *)(B + tmpoff);
  unsigned smax_u = 0;

  #pragma unroll 1
  for (int k = T.nvar + lane; k < nl; k += 32) {
    int f = fl[k];
    if (f & PIP_UNIT) continue;
    const V *row = pip_row(B, T, PIP_LINK(f));
    V d = den[k];
    unsigned s = 0;
    if (d == 1) {
      #pragma unroll 1
      for (int j = 0; j < T.nvar; j++) {
        pip_u64 u = pip_uabs(row[j]);
        if (u < 2147483648ull && (unsigned)u > s) s = (unsigned)u;
      }
    } else {
      #pragma unroll 1
      for (int j = 0; j < T.nvar; j++) {
        unsigned v = pip_size_term(row[j], d);
        if (v > s) s = v;
      }
    }
    sz[k] = (float)(double)s;
    if (s > smax_u) smax_u = s;
  }
  smax_u = W::redmax(smax_u);
  const double smax = (double)smax_u;
  W::sync();
  if (T.ni <= 32) {
    /* register-resident selection sort: lane r holds the record of position nvar + r; a swap
     * is a pair of shuffles, no shared-memory traffic and no barriers */
    const int k = T.nvar + lane;
    const bool in = lane < T.ni;
    int f = in ? fl[k] : PIP_UNIT;
    V d = in ? den[k] : 0;
    unsigned sb = in ? pip_f2u(sz[k]) : 0u;
    const bool movable = in && !(f & PIP_UNIT);
    /* nothing moves when the movable rows are already in non-decreasing order */
    {
      unsigned pm = movable ? sb : 0u;                    /* inclusive prefix max over movable rows */
      #pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned y = (unsigned)W::shfl_up((int)pm, o);
        if (lane >= o && y > pm) pm = y;
      }
      unsigned before = (unsigned)W::shfl_up((int)pm, 1);
      if (lane == 0) before = 0u;
      if (!W::any(movable && sb < before)) return;
    }
    bool moved = false;
    #pragma unroll 1
    for (int r = 0; r < T.ni; r++) {
      const bool unit_r = (W::shfl(f, r) & PIP_UNIT) != 0;
      if (unit_r) continue;
      unsigned key = 0xffffffffu;
      if (movable && lane >= r && (double)pip_u2f(sb) < smax) key = sb;
      const unsigned m = W::redmin(key);
      if (m == 0xffffffffu) break;                         /* no candidate left for any later r */
      const int src = pip_ffs(W::ballot(key == m)) - 1;
      if (src != r) {
        const int partner = lane == r ? src : lane == src ? r : lane;
        f = W::shfl(f, partner);
        d = W::shfl64(d, partner);
        sb = (unsigned)W::shfl((int)sb, partner);
        moved = true;
      }
    }
    if (moved && in) { fl[k] = f; den[k] = d; }
    W::sync();
    return;
  }
  #pragma unroll 1
  for (int i = T.nvar; i < nl; i++) {
    if (fl[i] & PIP_UNIT) continue;            /* uniform read */
    unsigned best = 0xffffffffu;
    int bestk = i;
    #pragma unroll 1
    for (int base = i & ~31; base < nl; base += 32) {
      int k = base + lane;
      unsigned key = 0xffffffffu;
      if (k >= i && k < nl && !(fl[k] & PIP_UNIT)) {
        float s = sz[k];
        if ((double)s < smax) key = pip_f2u(s);
      }
      unsigned m = W::redmin(key);
      if (m < best) {
        best = m;
        bestk = base + pip_ffs(W::ballot(key == m)) - 1;
      }
    }
    if (best != 0xffffffffu && bestk != i) {
      W::sync();
      if (lane == 0) {
        int f = fl[i]; fl[i] = fl[bestk]; fl[bestk] = f;
        V d = den[i]; den[i] = den[bestk]; den[bestk] = d;
        float s = sz[i]; sz[i] = sz[bestk]; sz[bestk] = s;
      }
      W::sync();
    }
  }
  W::sync();
}

/* ---- cold paths (PipOptions the polyhedral tools rarely set), out of line so that the hot code of
 * the instruction-supply-bound kernel keeps its layout ---------------------------------------- */

/* tab_sort_rows_xx with TRAITER_DUAL, source/traiter.c:556-623: the same selection sort, one step at
 * a time on lane 0, recording pos[r] = position of input row r after the sort (no Unit row exists in
 * the constraint range at the entry of a split-free solve) */
PIP_SDEVNI void pip_sort_rows_dual(pip_i64 *B, const PipTab &T, int tmpoff, int *pos)
{
  const int lane = W::lane();
  const int nl = T.nvar + T.ni;
  int *fl = pip_fl(B, T);
  V *den = pip_den(B, T);
  float *sz = (float *)(B + tmpoff);
  unsigned smax_u = 0;
  #pragma unroll 1
  for (int k = T.nvar + lane; k < nl; k += 32) {
    const V *row = pip_row(B, T, PIP_LINK(fl[k]));
    const V d = den[k];
    unsigned s = 0;
    #pragma unroll 1
    for (int j = 0; j < T.nvar; j++) {
      unsigned v;
      if (d == 1) { const pip_u64 u = pip_uabs(row[j]); v = u < 2147483648ull ? (unsigned)u : 0u; }
      else v = pip_size_term(row[j], d);
      if (v > s) s = v;
    }
    sz[k] = (float)(double)s;
    pos[k - T.nvar] = k;
    if (s > smax_u) smax_u = s;
  }
  smax_u = W::redmax(smax_u);
  const double smax = (double)smax_u;
  W::sync();
  if (lane == 0) {
    #pragma unroll 1
    for (int i = T.nvar; i < nl; i++) {
      double s = smax;
      int pivi = i;
      #pragma unroll 1
      for (int j = i; j < nl; j++) if ((double)sz[j] < s) { s = (double)sz[j]; pivi = j; }
      if (pivi == i) continue;
      const int f = fl[i]; fl[i] = fl[pivi]; fl[pivi] = f;
      const V d = den[i]; den[i] = den[pivi]; den[pivi] = d;
      const float t = sz[i]; sz[i] = sz[pivi]; sz[pivi] = t;
      int ri = -1, rb = -1;                            /* the two input rows trade positions */
      #pragma unroll 1
      for (int r = 0; r < T.ni; r++) { if (pos[r] == i) ri = r; else if (pos[r] == pivi) rb = r; }
      if (ri >= 0) pos[ri] = pivi;
      if (rb >= 0) pos[rb] = i;
    }
  }
  W::sync();
}

/* solution_dual_xx, source/traiter.c:274-294: one form per input row; a row that left the basis (its
 * position is Unit) reports the entry of position 0 in the column it owns.  Without cuts the tableau
 * height is nvar + ni, so the list has ni forms (1 + 2 ni cells at out[at..]). */
PIP_SDEVNI bool pip_emit_dual(pip_i64 *B, const PipTab &T, const int *pos, PipCell *out, int at)
{
  bool wide = false;
  const int dtotal = 1 + 2 * T.ni;
  const int *fl = pip_fl(B, T);
  const V *den = pip_den(B, T);
  const int f0 = fl[0];
  const V d0 = den[0];
  #pragma unroll 1
  for (int c = W::lane(); c < dtotal; c += 32) {
    if (c == 0) { pip_put(out, at, PIP_C_LIST, T.ni, 0); continue; }
    const int i = (c - 1) >> 1;
    if (((c - 1) & 1) == 0) { pip_put(out, at + c, PIP_C_FORM, 1, 0); continue; }
    const int fp = fl[pos[i]];
    if (fp & PIP_UNIT) wide = pip_put(out, at + c, PIP_C_VAL, pip_entry(B, T, f0, d0, PIP_LINK(fp)), d0) || wide;
    else pip_put(out, at + c, PIP_C_VAL, 0, 1);
  }
  return wide;
}

/* Gondran's deepest cut, source/integrer.c:417-438 (constant cuts only): multiply the cut by the unit
 * lambda of Z/D that maximises its depth; lambda is scalar work for lane 0.  Returns false on a
 * division by zero (the reference would die of SIGFPE). */
PIP_SDEVNI bool pip_deepest_cut(V *cut, int nvar, V D, unsigned &ovf)
{
  const int lane = W::lane();
  pip_i64 lambda = 0;
  int bad = 0;
  const pip_i64 D64 = (pip_i64)D;
  if (lane == 0) {
    const pip_i64 tt = -(pip_i64)cut[nvar];
    const pip_i64 delta = pip_gcd(tt, D64);
    if (delta == 0) bad = 1;
    else {
      const pip_i64 tau = pip_div(tt, delta), dd = pip_div(D64, delta);
      lambda = pip_bezout(dd - 1, tau, dd);
      int guard = 0;
      while (pip_gcd(lambda, D64) != 1) {
        lambda = (pip_i64)((pip_u64)lambda + (pip_u64)dd);
        if (++guard > (1 << 22)) { bad = 1; break; }
      }
    }
  }
  if (W::shfl(bad, 0)) return false;
  lambda = W::shfl64(lambda, 0);
  #pragma unroll 1
  for (int j = lane; j < nvar; j += 32)
    cut[j] = PipVal<V>::store(pip_mod((pip_i64)((pip_u64)lambda * (pip_u64)(pip_i64)cut[j]), D64), ovf);
  if (lane == 0) {
    const pip_i64 tt = pip_mod((pip_i64)((pip_u64)(pip_i64)cut[nvar] * (pip_u64)lambda), D64);
    cut[nvar] = PipVal<V>::store(-(D64 - tt), ovf);
  }
  W::sync();
  return true;
}

/* exam_coef_xx, source/traiter.c:101-159.  Returns the first row proved negative or nl. */
PIP_SDEV int pip_exam_coef(pip_i64 *B, const PipTab &T, int bigparm)
{
  const int lane = W::lane();
  const int nl = T.nvar + T.ni, ncol = T.nvar + T.nparm + 1;
  int *fl = pip_fl(B, T);
  if (bigparm >= 0) {
    #pragma unroll 1
    for (int base = 0; base < nl; base += 32) {
      int k = base + lane;
      int f = k < nl ? fl[k] : 0;
      bool unk = PIP_FLAG(f) == PIP_UNKNOWN;
      V v = unk ? pip_row(B, T, PIP_LINK(f))[bigparm] : 0;
      unsigned mneg = W::ballot(unk && v < 0);
      int first = mneg ? pip_ffs(mneg) - 1 : 32;
      if (unk && v > 0 && lane < first) fl[k] = PIP_MKFL(PIP_PLUS, PIP_LINK(f));
      if (mneg) {
        if (lane == first) fl[k] = PIP_MKFL(PIP_MINUS, PIP_LINK(f));
        W::sync();
        return base + first;
      }
    }
    W::sync();
  }
  #pragma unroll 1
  for (int base = 0; base < nl; base += 32) {
    int k = base + lane;
    int f = k < nl ? fl[k] : 0;
    bool unk = PIP_FLAG(f) == PIP_UNKNOWN;
    int ff = PIP_ZERO;
    if (unk) {
      const V *row = pip_row(B, T, PIP_LINK(f));
      #pragma unroll 1
      for (int j = T.nvar + 1; j < ncol; j++) {
        V v = row[j];
        int fff = v < 0 ? PIP_MINUS : v > 0 ? PIP_PLUS : PIP_ZERO;
        if (fff != PIP_ZERO && fff != ff) {
          if (ff == PIP_ZERO) ff = fff;
          else { ff = PIP_UNKNOWN; break; }
        }
      }
      V c = row[T.nvar];
      int fff = c < 0 ? PIP_MINUS : c > 0 ? PIP_PLUS : PIP_ZERO;
      if (ff == PIP_PLUS) { if (fff == PIP_MINUS) ff = PIP_UNKNOWN; }
      else if (ff == PIP_ZERO) ff = fff;
      else if (ff == PIP_MINUS) { if (fff != PIP_MINUS) ff = PIP_UNKNOWN; }
    }
    unsigned mneg = W::ballot(unk && ff == PIP_MINUS);
    int first = mneg ? pip_ffs(mneg) - 1 : 32;
    if (unk && lane <= first) fl[k] = PIP_MKFL(ff, PIP_LINK(f));
    if (mneg) { W::sync(); return base + first; }
  }
  W::sync();
  return nl;
}

/* sign class of an Unknown row from its parametric part and constant (source/traiter.c:118-153) */
PIP_SDEV int pip_classify_row(const V *row, int nvar, int ncol)
{
  int ff = PIP_ZERO;
  #pragma unroll 1
  for (int j = nvar + 1; j < ncol; j++) {
    const V v = row[j];
    const int fff = v < 0 ? PIP_MINUS : v > 0 ? PIP_PLUS : PIP_ZERO;
    if (fff != PIP_ZERO && fff != ff) {
      if (ff == PIP_ZERO) ff = fff;
      else { ff = PIP_UNKNOWN; break; }
    }
  }
  const V c = row[nvar];
  const int fff = c < 0 ? PIP_MINUS : c > 0 ? PIP_PLUS : PIP_ZERO;
  if (ff == PIP_PLUS) { if (fff == PIP_MINUS) ff = PIP_UNKNOWN; }
  else if (ff == PIP_ZERO) ff = fff;
  else if (ff == PIP_MINUS) { if (fff != PIP_MINUS) ff = PIP_UNKNOWN; }
  return ff;
}

/* ---- team-wide position scans (class M): every thread of the CTA runs the *_body between the two
 * barriers of a command; thread = position modulo the team size, first hit = team minimum -------- */
PIP_SDEV void pip_team_first_body(PipTeam *tm, int tid)
{
  const int *fl = (const int *)(tm->B + tm->T.fl);
  const int n = tm->op_n, kind = tm->op_kind, arg = tm->op_arg;
  int best = n;
  #pragma unroll 1
  for (int k = tm->op_from + tid; k < n; k += tm->nthreads) {
    const int f = fl[k];
    const bool hit = kind == 0 ? (f & arg) != 0 : ((f & PIP_UNIT) != 0 && PIP_LINK(f) == arg);
    if (hit) { best = k; break; }
  }
  best = (int)W::redmin((unsigned)best);
  if (W::lane() == 0 && best < n) pip_team_min(&tm->result, best);
}

/* exam_coef_xx without a big parameter (source/traiter.c:118-157): classify the Unknown rows up to
 * and including the first one proved negative */
PIP_SDEV void pip_team_exam_body(PipTeam *tm, int tid)
{
  const PipTab T = tm->T;
  pip_i64 *B = tm->B;
  const int nl = T.nvar + T.ni, ncol = T.nvar + T.nparm + 1;
  int *fl = pip_fl(B, T);
  int best = nl;
  #pragma unroll 1
  for (int k = tid; k < nl; k += tm->nthreads) {
    const int f = fl[k];
    if (PIP_FLAG(f) != PIP_UNKNOWN) continue;
    if (pip_classify_row(pip_row(B, T, PIP_LINK(f)), T.nvar, ncol) == PIP_MINUS) { best = k; break; }
  }
  best = (int)W::redmin((unsigned)best);
  if (W::lane() == 0 && best < nl) pip_team_min(&tm->result, best);
  pip_team_barrier(tm->nthreads);
  const int first = tm->result;
  #pragma unroll 1
  for (int k = tid; k < nl && k <= first; k += tm->nthreads) {
    const int f = fl[k];
    if (PIP_FLAG(f) != PIP_UNKNOWN) continue;
    fl[k] = PIP_MKFL(pip_classify_row(pip_row(B, T, PIP_LINK(f)), T.nvar, ncol), PIP_LINK(f));
  }
}

/* leader side: post a scan command, take part in it, return the team's answer */
PIP_SDEV int pip_team_scan(PipTeam *tm, pip_i64 *B, const PipTab &T, int cmd, int kind, int arg, int from, int n)
{
  const int lane = W::lane();
  W::sync();
  if (lane == 0) {
    tm->cmd = cmd; tm->T = T; tm->B = B; tm->op_kind = kind; tm->op_arg = arg; tm->op_from = from; tm->op_n = n;
    tm->result = n;
  }
  W::sync();
  pip_team_barrier(tm->nthreads);
  if (cmd == PIP_TEAM_FIRST) pip_team_first_body(tm, lane);
  else pip_team_exam_body(tm, lane);
  pip_team_barrier(tm->nthreads);
  const int r = tm->result;
  W::sync();
  return r;
}

/* chercher(Minus) + exam_coef for a tableau of at most 32 positions: one register-resident pass
 * (source/traiter.c:669-680); returns the pivot row or nl.  (-DPIP_NO_SCAN32 builds without it: 160
 * instructions less of executed code, but measured no faster once the layout luck of a build is averaged
 * out -- profiles/r2_build_variants.log.) */
PIP_SDEV int pip_scan32(pip_i64 *B, const PipTab &T, int bigparm)
{
  const int lane = W::lane();
  const int nl = T.nvar + T.ni, ncol = T.nvar + T.nparm + 1;
  int *fl = pip_fl(B, T);
  int f = lane < nl ? fl[lane] : 0;
  unsigned m = W::ballot((f & PIP_MINUS) != 0);
  if (m) return pip_ffs(m) - 1;
  const V *row = pip_row(B, T, PIP_LINK(f));
  bool dirty = false;
  if (bigparm >= 0) {
    const bool unk = PIP_FLAG(f) == PIP_UNKNOWN;
    const V v = unk ? row[bigparm] : 0;
    m = W::ballot(unk && v < 0);
    const int first = m ? pip_ffs(m) - 1 : 32;
    if (unk && v > 0 && lane < first) { f = PIP_MKFL(PIP_PLUS, PIP_LINK(f)); fl[lane] = f; }
    if (m) {
      if (lane == first) fl[lane] = PIP_MKFL(PIP_MINUS, PIP_LINK(f));
      W::sync();
      return first;
    }
  }
  const bool unk = PIP_FLAG(f) == PIP_UNKNOWN;
  int ff = PIP_ZERO;
  if (unk) {
    #pragma unroll 1
    for (int j = T.nvar + 1; j < ncol; j++) {
      const V v = row[j];
      const int fff = v < 0 ? PIP_MINUS : v > 0 ? PIP_PLUS : PIP_ZERO;
      if (fff != PIP_ZERO && fff != ff) {
        if (ff == PIP_ZERO) ff = fff;
        else { ff = PIP_UNKNOWN; break; }
      }
    }
    const V c = row[T.nvar];
    const int fff = c < 0 ? PIP_MINUS : c > 0 ? PIP_PLUS : PIP_ZERO;
    if (ff == PIP_PLUS) { if (fff == PIP_MINUS) ff = PIP_UNKNOWN; }
    else if (ff == PIP_ZERO) ff = fff;
    else if (ff == PIP_MINUS) { if (fff != PIP_MINUS) ff = PIP_UNKNOWN; }
    dirty = true;
  }
  m = W::ballot(unk && ff == PIP_MINUS);
  const int first = m ? pip_ffs(m) - 1 : 32;
  if (dirty && lane <= first) fl[lane] = PIP_MKFL(ff, PIP_LINK(f));
  W::sync();
  return m ? first : nl;
}

/* valeur_xx, source/traiter.c:246-252 */
PIP_SDEV V pip_entry(pip_i64 *B, const PipTab &T, int f, V d, int j)
{
  if (f & PIP_UNIT) return PIP_LINK(f) == j ? d : 0;
  return pip_row(B, T, PIP_LINK(f))[j];
}

/* choisir_piv_xx, source/traiter.c:297-341: lexicographic pivot column.  For each candidate
 * column the difference x_k = pivot*val(k,j) - val(k,pivj)*foo is evaluated for 32 positions at
 * a time; the first non-zero x in position order decides. */
PIP_SDEV int pip_choose_column(pip_i64 *B, const PipTab &T, int pivi, V &pivot_out)
{
  const int lane = W::lane();
  const int nl = T.nvar + T.ni;
  const int *fl = pip_fl(B, T);
  const V *den = pip_den(B, T);
  const V *prow = pip_row(B, T, PIP_LINK(fl[pivi]));
  int pivj = -1;
  V pivot = 0;
  #pragma unroll 1
  for (int cb = 0; cb < T.nvar; cb += 32) {
    int jc = cb + lane;
    unsigned cand = W::ballot(jc < T.nvar && prow[jc] > 0);
    while (cand) {
      int j = cb + pip_ffs(cand) - 1;
      cand &= cand - 1;
      V foo = prow[j];
      if (pivj < 0) { pivj = j; pivot = foo; continue; }
      bool neg = false;
      #pragma unroll 1
      for (int base = 0; base < nl; base += 32) {
        int k = base + lane;
        pip_i64 x = 0;
        if (k < nl) {
          int f = fl[k];
          V a, b;
          if (f & PIP_UNIT) {
            int u = PIP_LINK(f);
            V d = den[k];
            a = (u == j) ? d : 0;
            b = (u == pivj) ? d : 0;
          } else {
            const V *row = pip_row(B, T, PIP_LINK(f));
            a = row[j]; b = row[pivj];
          }
          x = PipVal<V>::cross(pivot, a, b, foo);
        }
        unsigned nz = W::ballot(x != 0);
        if (nz) {
          unsigned ng = W::ballot(x < 0);
          neg = (ng >> (pip_ffs(nz) - 1)) & 1u;
          break;
        }
      }
      if (neg) { pivj = j; pivot = foo; }
    }
  }
  pivot_out = pivot;
  return pivj;
}

/* rank-1 update of every stored row but the pivot row (source/traiter.c:467-502), fused with
 * the re-flagging from the sign of the new pivot-column entry (source/traiter.c:518-529).
 * The caller's thread handles positions first, first + step, ... */
PIP_SDEV void pip_update_rows(pip_i64 *B, const PipTab &T, int pivi, int pivj, V pivot, V dpiv, int first, int step,
                              unsigned &ovf, bool &fault, PipStats *stp = nullptr)
{
  PIP_ULAP_BEGIN(stp);
  const int nl = T.nvar + T.ni, ncol = T.nvar + T.nparm + 1;
  int *fl = pip_fl(B, T);
  V *den = pip_den(B, T);
  const V *prow = pip_row(B, T, PIP_LINK(fl[pivi]));
  #pragma unroll 1
  for (int k = first; k < nl; k += step) {
    if (k == pivi) continue;
    const int f = fl[k];
    if (f & PIP_UNIT) continue;
    V *row = pip_row(B, T, PIP_LINK(f));
    V foo = row[pivj];
    const V dk = den[k];
    if (foo == 0 && dk == 1) continue;        /* identity update, g stays 1, sign Zero: no re-flag */
    V lpiv = pivot;
    if (foo == 0) lpiv = 1;                    /* gcd(pivot,0) = pivot */
    else if (pivot != 1 && foo != 1 && foo != -1) {
      V d = pip_gcd(pivot, foo);
      if (d != 1) { lpiv = pip_div(pivot, d); foo = pip_div(foo, d); }
    }
    const V newden = PipVal<V>::mul(lpiv, dk, ovf);
    V g = newden;
    V zp;
    PIP_ULAP(stp, PIP_PH_U_HEAD);
    if (TEAM) {
      /* arena in global memory (classes G / M): the row is walked in blocks of PIP_UB columns with all
       * loads of a block issued before the first use, so a pass costs one L2 round trip per block
       * instead of one per column (the loop is latency-bound, not bandwidth-bound) */
      enum { PIP_UB = 8 };
      zp = PipVal<V>::mul(dpiv, foo, ovf);
      pip_u64 orz = (pip_u64)(pip_i64)zp;
      #pragma unroll 1
      for (int j0 = 0; j0 < ncol; j0 += PIP_UB) {
        V a[PIP_UB], b[PIP_UB];
        #pragma unroll
        for (int u = 0; u < PIP_UB; u++) { const bool in = j0 + u < ncol; a[u] = in ? row[j0 + u] : 0; b[u] = in ? prow[j0 + u] : 0; }
        #pragma unroll
        for (int u = 0; u < PIP_UB; u++) {
          V z = PipVal<V>::mulsub(a[u], lpiv, b[u], foo, ovf);
          if (j0 + u == pivj) z = zp;
          if (j0 + u < ncol) row[j0 + u] = z;
          orz |= (pip_u64)(pip_i64)z;
        }
      }
      PIP_ULAP(stp, PIP_PH_U_PASS1);
      if (g != 1) {
        if ((g & (g - 1)) == 0 && g > 0) { orz |= (pip_u64)(pip_i64)g; g = (V)(pip_i64)(orz & (0ull - orz)); }
        else {
          #pragma unroll 1
          for (int j0 = 0; j0 < ncol && g != 1; j0 += PIP_UB) {
            V a[PIP_UB];
            #pragma unroll
            for (int u = 0; u < PIP_UB; u++) a[u] = j0 + u < ncol ? row[j0 + u] : 0;
            #pragma unroll 1
            for (int u = 0; u < PIP_UB && g != 1; u++) g = pip_gcd(g, a[u]);      /* gcd(g, 0) = g */
          }
        }
      }
      PIP_ULAP(stp, PIP_PH_U_GCD);
      if (g != 1) {
        if (g == 0) { fault = true; continue; }
        const PipExactDiv e = pip_exact_prepare((pip_i64)g);
        #pragma unroll 1
        for (int j0 = 0; j0 < ncol; j0 += PIP_UB) {
          V a[PIP_UB];
          #pragma unroll
          for (int u = 0; u < PIP_UB; u++) a[u] = j0 + u < ncol ? row[j0 + u] : 0;
          #pragma unroll
          for (int u = 0; u < PIP_UB; u++) if (j0 + u < ncol) row[j0 + u] = (V)pip_exact_apply((pip_i64)a[u], e);
        }
        den[k] = (V)pip_exact_apply((pip_i64)newden, e);
      } else den[k] = newden;
      PIP_ULAP(stp, PIP_PH_U_DIV);
    } else {
    /* pass 1: pure arithmetic.  The generic formula gives 0 in column pivj (foo*lpiv == pivot*foo'),
     * the real value dpiv*foo' is patched in afterwards, so the loop body has no special case */
    pip_u64 orz = 0;
#ifndef PIP_NO_UPD2
    if (PipVal<V>::narrow) {
      /* int32 storage: the range check is made once per row on the OR of the magnitudes instead of per entry
       * (3 instructions per entry instead of 5.5).  The OR only proves |z| < 2^30 for all entries, so a row
       * with an entry in [2^30, 2^31) now sends the problem to the int64 class although it would have fitted */
      unsigned acc_hi = 0, acc_lo = 0, orl = 0;
      const V *pr = prow;
      V *r = row;
      V *const rend = row + ncol;
      /* (not unrolled: with unroll 2 / 4 the same build ran the 10^6 job in 124.3 / 128.8 ms instead of 119.9) */
#if defined(PIP_UPD_UNROLL2)
      #pragma unroll 2
#elif defined(PIP_UPD_UNROLL4)
      #pragma unroll 4
#else
      #pragma unroll 1
#endif
      for (; r != rend; r++, pr++) {
        const pip_i64 z = PipVal<V>::cross(*r, lpiv, *pr, foo);
        const int lo = (int)z, hi = (int)(z >> 32), sg = lo >> 31;
        *r = (V)lo;
        orl |= (unsigned)lo; acc_hi |= (unsigned)(hi ^ sg); acc_lo |= (unsigned)(lo ^ sg);
      }
      ovf |= acc_hi | (acc_lo >> 30);
      orz = orl;
    } else
#endif
    {
    #pragma unroll 2
    for (int j = 0; j < ncol; j++) {
      const V z = PipVal<V>::mulsub(row[j], lpiv, prow[j], foo, ovf);
      row[j] = z;
      orz |= (pip_u64)(pip_i64)z;
    }
    }
    zp = PipVal<V>::mul(dpiv, foo, ovf);
    row[pivj] = zp;
    orz |= (pip_u64)(pip_i64)zp;
    /* pass 2 (only when the row has a common factor to shed): g = gcd(newden, z_0 .. z_n).
     * A power-of-two g folds into the OR: gcd(2^a, z..) = lowest set bit of (2^a | z_0 | ..) */
    if (g != 1) {
      if ((g & (g - 1)) == 0 && g > 0) { orz |= (pip_u64)(pip_i64)g; g = (V)(pip_i64)(orz & (0ull - orz)); }
      else {
        #pragma unroll 1
        for (int j = 0; j < ncol && g != 1; j++) g = pip_gcd(g, row[j]);
      }
    }
    if (g != 1) {
      if (g == 0) { fault = true; continue; }
      if ((g & (g - 1)) == 0) {
        int sh = 0;
        while (((pip_u64)(pip_i64)g >> sh) != 1ull) sh++;
        #pragma unroll 2
        for (int j = 0; j < ncol; j++) row[j] = row[j] >> sh;
        den[k] = newden >> sh;
      } else if (PipVal<V>::narrow) {
        /* int32 storage: the shared 32-bit division (the modular-inverse route is 64-bit code the int32
         * kernel otherwise never runs) */
        #pragma unroll 1
        for (int j = 0; j < ncol; j++) row[j] = pip_div(row[j], g);
        den[k] = pip_div(newden, g);
      } else {
        PipExactDiv e = pip_exact_prepare((pip_i64)g);
        #pragma unroll 2
        for (int j = 0; j < ncol; j++) row[j] = (V)pip_exact_apply((pip_i64)row[j], e);
        den[k] = (V)pip_exact_apply((pip_i64)newden, e);
      }
    } else den[k] = newden;
    }
    /* sign of the new entry in column pivj = sign of zp (g > 0) */
    int ff = PIP_FLAG(f);
    const int fff = zp < 0 ? PIP_MINUS : zp == 0 ? PIP_ZERO : PIP_PLUS;
    if (fff != PIP_ZERO && fff != ff) {
      if (ff == PIP_ZERO) ff = (fff == PIP_MINUS ? PIP_UNKNOWN : fff);
      else ff = PIP_UNKNOWN;
      fl[k] = PIP_MKFL(ff, PIP_LINK(f));
    }
  }
}

/* pivoter_xx, source/traiter.c:345-548.
 * returns 0 done, -1 no positive coefficient (infeasible), or a PIP_ST_* fatal status */
PIP_SDEV int pip_pivot(pip_i64 *B, PipTab &T, int pivi, PipStats &st, PipTeam *tm)
{
  const int lane = W::lane();
  const int nl = T.nvar + T.ni, ncol = T.nvar + T.nparm + 1;
  int *fl = pip_fl(B, T);
  V *den = pip_den(B, T);
  V pivot;
  const int pivj = pip_choose_column(B, T, pivi, pivot);
  PIP_LAP(st, PIP_PH_CHOOSE);
  if (pivj < 0) return -1;

  const int pslot = PIP_LINK(fl[pivi]);
  V *prow = pip_row(B, T, pslot);
  const V dpiv = den[pivi];
  /* determinant bookkeeping = the overflow verdict, source/traiter.c:394-447, always in 64 bits
   * (the factors grow up to 2^63).  The common case -- integer pivot row (dpiv == 1), unit pivot,
   * first factor far from full -- changes nothing and is recognised up front; everything else is
   * scalar work done by lane 0 on the factors in the arena. */
  {
    pip_i64 *det = B + T.det;
    int verdict = 0, ldet = T.ldet;
    const pip_i64 det0 = det[0];
    const bool trivial = dpiv == 1 && pivot == 1 && det0 > -(1ll << 61) && det0 < (1ll << 61);
    if (!trivial) {
      if (lane == 0) {
        pip_i64 d = (dpiv == 1) ? 1 : pip_gcd((pip_i64)pivot, (pip_i64)dpiv);
        if (d == 0) verdict = PIP_ST_FAULT;
        else {
          pip_i64 ppivot = pivot, dppiv = dpiv;
          if (d != 1) { ppivot = pip_div((pip_i64)pivot, d); dppiv = pip_div((pip_i64)dpiv, d); }
          #pragma unroll 1
          for (int i = 0; i < ldet && dppiv != 1; i++) {
            const pip_i64 g = pip_gcd(det[i], dppiv);
            if (g == 0) { verdict = PIP_ST_FAULT; break; }
            if (g != 1) { det[i] = pip_div(det[i], g); dppiv = pip_div(dppiv, g); }
          }
          if (!verdict && dppiv != 1) verdict = PIP_ST_FATAL + 1;        /* "Integer overflow" */
          if (!verdict) {
            int i = 0;
            const int bp = pip_bitlen(ppivot);
            #pragma unroll 1
            for (; i < ldet; i++)
              if (pip_bitlen(det[i]) + bp < 64) { det[i] = (pip_i64)((pip_u64)det[i] * (pip_u64)ppivot); break; }
            if (i >= ldet) {
              ldet++;
              if (ldet >= PIP_MAX_DET) verdict = PIP_ST_FATAL + 1;       /* "Integer overflow : 4" */
              else det[i] = ppivot;
            }
          }
        }
      }
      verdict = W::shfl(verdict, 0);
      T.ldet = W::shfl(ldet, 0);
      if (verdict) return verdict;
    }
  }
  st.pivots++;
  if ((unsigned)nl > st.max_rows) st.max_rows = nl;
  if ((unsigned)ncol > st.max_cols) st.max_cols = ncol;
  st.elem_updates += (unsigned long long)(T.ni - 1) * ncol;

  /* rank-1 update + re-flag (pip_update_rows): by this warp, or in team mode by the whole CTA */
  bool fault = false;
  unsigned ovf = 0;
  if (TEAM && tm != nullptr && nl >= PIP_TEAM_MIN_ROWS) {
    if (lane == 0) {
      tm->cmd = PIP_TEAM_UPDATE; tm->pivi = pivi; tm->pivj = pivj; tm->pivot = (pip_i64)pivot; tm->dpiv = (pip_i64)dpiv;
      tm->T = T; tm->B = B; tm->fault = 0; tm->ovf = 0;
    }
    W::sync();
    pip_team_barrier(tm->nthreads);
    pip_update_rows(B, T, pivi, pivj, pivot, dpiv, lane, tm->nthreads, ovf, fault, lane == 0 ? &st : nullptr);
    if (fault) tm->fault = 1;
    if (ovf) tm->ovf = 1;
#ifdef PIP_PROFILE
    const long long w0_ = clock64();
#endif
    pip_team_barrier(tm->nthreads);
#ifdef PIP_PROFILE
    if (lane == 0) st.cyc[PIP_PH_U_WAIT] += (unsigned long long)(clock64() - w0_);
#endif
    fault = tm->fault != 0;
    ovf |= tm->ovf;
  } else pip_update_rows(B, T, pivi, pivj, pivot, dpiv, lane, 32, ovf, fault);
  if (PipVal<V>::narrow && W::any(ovf != 0)) return PIP_ST_WIDEN;
  if (!PipVal<V>::narrow) st.wrapped |= ovf;
  if (W::any(fault)) return PIP_ST_FAULT;
  W::sync();
  PIP_LAP(st, PIP_PH_UPDATE);
  /* the Unit position owning pivj takes the pivot row's slot, source/traiter.c:503-516 */
  int ku = nl;
  if (TEAM && tm != nullptr && nl >= PIP_TEAM_MIN_SCAN) ku = pip_team_scan(tm, B, T, PIP_TEAM_FIRST, 1, pivj, 0, nl);
  else {
    #pragma unroll 1
    for (int base = 0; base < nl; base += 32) {
      int k = base + lane;
      int f = k < nl ? fl[k] : 0;
      unsigned m = W::ballot((f & PIP_UNIT) && PIP_LINK(f) == pivj && k < nl);
      if (m) { ku = base + pip_ffs(m) - 1; break; }
    }
  }
  if (ku >= nl) return PIP_ST_FAULT;
  #pragma unroll 1
  for (int j = lane; j < ncol; j += 32) prow[j] = (j == pivj) ? dpiv : -prow[j];
  W::sync();
  if (lane == 0) {
    fl[ku] = PIP_MKFL(PIP_PLUS, pslot); den[ku] = pivot;
    fl[pivi] = PIP_MKFL(PIP_UNIT | PIP_ZERO, pivj); den[pivi] = 1;
  }
  W::sync();
  /* (the new row at ku keeps Plus: its pivot-column entry is dpiv > 0; every other row was
   * re-flagged inside the update loop) */
  PIP_LAP(st, PIP_PH_SWAP);
  return 0;
}

/* append one cell; returns true when it does not fit the packed wire format */
PIP_SDEV bool pip_put(PipCell *out, int idx, int kind, V p1, V p2)
{
  out[idx].kind = kind; out[idx].pad = 0; out[idx].p1 = p1; out[idx].p2 = p2;
  return !PIP_CELL_FITS(p1, p2);
}

/* solution_xx, source/traiter.c:255-271: 1 + nvar*(2+nparm) cells, lane-parallel */
PIP_SDEVNI bool pip_emit_solution(pip_i64 *B, const PipTab &T, PipCell *out, int at)
{
  bool wide = false;
  const int per = T.nparm + 2, total = 1 + T.nvar * per;
  const int *fl = pip_fl(B, T);
  const V *den = pip_den(B, T);
  #pragma unroll 1
  for (int c = W::lane(); c < total; c += 32) {
    if (c == 0) { pip_put(out, at, PIP_C_LIST, T.nvar, 0); continue; }
    int i = (c - 1) / per, r = (c - 1) % per;
    if (r == 0) { pip_put(out, at + c, PIP_C_FORM, T.nparm + 1, 0); continue; }
    int j = (r == per - 1) ? T.nvar : T.nvar + r;
    int f = fl[i];
    V d = den[i];
    wide = pip_put(out, at + c, PIP_C_VAL, pip_entry(B, T, f, d, j), d) || wide;
  }
  return wide;
}

/* ---- word mode: the solver writes the SERIALISED quast itself (the word stream of pip_decode.h) instead
 * of solution cells for a decode kernel to parse.  Only for problems whose decode needs no column surgery
 * (PIP_F_SIMPLE_SER), i.e. every bulk workload: the cells were written (24 B each), read back and re-parsed
 * by a second kernel that cost a fifth of the solve.  The grammar is pre-order like the cells; the only
 * forward reference is the new-parameter count that opens a node, kept as a reserved slot (node_at) and
 * patched when the node's kind is known.  (The stream's hash is taken by the copy kernel that follows:
 * hashing here, 20 instructions at every one of two dozen emission sites, cost the instruction-supply-bound
 * solve kernel more than the whole decode kernel it replaced.) */
struct PipWordOut {
  V *w;                    /* the warp's window, as words of the stored type */
  unsigned pos;            /* words emitted so far (warp-uniform) */
  int node_at;             /* slot of the open node's new-parameter count, -1 = no node open */
  unsigned nnew;
  bool wide;               /* some word left int32 (int64 classes) */
};
PIP_SDEV void pip_wout(PipWordOut &o, unsigned idx, pip_i64 v)
{
  o.w[idx] = (V)v;
  if (!PipVal<V>::narrow) o.wide = o.wide || (v != (pip_i64)(int)v);
}
PIP_SDEV void pip_wnode_open(PipWordOut &o)
{
  if (o.node_at < 0) { o.node_at = (int)o.pos; o.pos++; o.nnew = 0; }
}
/* the node's kind is about to be written: patch the new-parameter count (lane 0) */
PIP_SDEV void pip_wnode_kind(PipWordOut &o, int kind)
{
  pip_wnode_open(o);
  if (W::lane() == 0) { pip_wout(o, (unsigned)o.node_at, (pip_i64)o.nnew); pip_wout(o, o.pos, kind); }
  o.pos++;
  o.node_at = -1;
}
/* solution_xx as words: `1 nvar { 1 VEC }* 0`, VEC = nparm+1 { num den }*, every value reduced by its gcd
 * with the row denominator exactly as sol_vector_edit_xx does (source/sol.c:435-512) */
PIP_SDEVNI void pip_emit_solution_words(pip_i64 *B, const PipTab &T, PipWordOut &o)
{
  const int lane = W::lane();
  const int np1 = T.nparm + 1, per = 2 + 2 * np1;
  pip_wnode_kind(o, 1);
  const unsigned base = o.pos;
  if (T.nvar == 0) {
    if (lane == 0) { pip_wout(o, base, 1); pip_wout(o, base + 1, 0); pip_wout(o, base + 2, 0); }
    o.pos += 3;
    return;
  }
  const int *fl = pip_fl(B, T);
  const V *den = pip_den(B, T);
  if (lane == 0) { pip_wout(o, base, T.nvar); pip_wout(o, base + 1 + (unsigned)T.nvar * per, 0); }
  #pragma unroll 1
  for (int i = lane; i < T.nvar; i += 32) { pip_wout(o, base + 1 + i * per, 1); pip_wout(o, base + 2 + i * per, np1); }
  const int pairs = T.nvar * np1;
  #pragma unroll 1
  for (int q = lane; q < pairs; q += 32) {
    const int i = q / np1, r = q - i * np1;
    const int j = (r == np1 - 1) ? T.nvar : T.nvar + 1 + r;
    const V D = den[i];
    const V N = pip_entry(B, T, fl[i], D, j);
    V num = N, dd = 1;
    if (D != 1) {
      const V d = pip_gcd(N, D);
      if (d != 1) num = d ? pip_div(N, d) : (V)0;
      if (d != D) dd = d ? pip_div(D, d) : (V)0;
    }
    pip_wout(o, base + 3 + i * per + 2 * r, (pip_i64)num);
    pip_wout(o, base + 4 + i * per + 2 * r, (pip_i64)dd);
  }
  o.pos += 2 + (unsigned)T.nvar * per;
}

/* has_cut_xx, source/integrer.c:230-254 (serial, one lane) */
PIP_SDEV bool pip_has_cut(const V *ctx, int cstride, int nr, int nparm, int p, const V *cut)
{
  #pragma unroll 1
  for (int row = 0; row < nr; row++) {
    const V *r = ctx + row * cstride;
    if (r[p] != cut[1 + nparm]) continue;
    if (r[nparm] != cut[0]) continue;
    int col;
    #pragma unroll 1
    for (col = p + 1; col < nparm; col++) if (r[col] != 0) break;
    if (col < nparm) continue;
    #pragma unroll 1
    for (col = 0; col < p; col++) if (r[col] != cut[1 + col]) break;
    if (col < p) continue;
    return true;
  }
  return false;
}

/* find_parm_xx, source/integrer.c:258-291 (serial, one lane; cut = const, params, denominator) */
PIP_SDEV int pip_find_parm(const V *ctx, int cstride, int nr, int nparm, V *cut)
{
  if (cut[1 + nparm - 1] != 0) return -1;
  cut[0] = cut[0] + cut[1 + nparm] - 1;
  #pragma unroll 1
  for (int p = nparm - 1; p >= 0; p--) {
    if (cut[1 + p] != 0) break;
    if (!pip_has_cut(ctx, cstride, nr, nparm, p, cut)) continue;
    cut[0] = cut[0] + 1 - cut[1 + nparm];
    #pragma unroll 1
    for (int c = 0; c < nparm + 2; c++) cut[c] = -cut[c];
    bool found = pip_has_cut(ctx, cstride, nr, nparm, p, cut);
    #pragma unroll 1
    for (int c = 0; c < nparm + 2; c++) cut[c] = -cut[c];
    if (found) return p;
    cut[0] = cut[0] + cut[1 + nparm] - 1;
  }
  cut[0] = cut[0] + 1 - cut[1 + nparm];
  return -1;
}

/* ---- register-resident feasibility solve (compa_test_xx / the context check) ----------------------
 * Nine out of ten traiter_xx activations of a parametric problem are the two integer feasibility solves
 * compa_test_xx runs per tested row (source/traiter.c:191-220) on a tableau of nparm Unit positions plus
 * the nc context rows plus the tested row, nparm + 1 columns wide: a handful of rows of a handful of
 * words.  Solved through the general path every one of them pays a copy into the arena, full-warp scans
 * and barriers for 8 busy lanes.  Here the whole sub-tableau lives in registers: lane = one record
 * (a Unit position, or a stored row whose PIP_SUBREG_NC words sit in the lane's registers), labelled with
 * its *position*; every order-dependent scan of the reference (chercher, exam_coef, the k loop of
 * choisir_piv, the i loop of integrer) is a minimum over position labels, the sort and the slot swap of a
 * pivot only permute labels, the pivot row reaches the other lanes by shuffles.  Nothing is read from or
 * written to memory after the load, so there is no barrier in the loop.
 *
 * Same arithmetic, same decisions, same pivot sequence as the general path (and the reference): the
 * emulator tests run both and compare cells, statuses and pivot counts.  Returns 1 feasible, 0 infeasible,
 * -1 when the solve does not fit the register form (more than 32 positions, more than PIP_SUBREG_NC
 * columns: the caller falls back to the general path, nothing has been counted), or a PIP_ST_* status. */
#ifndef PIP_SUBREG_NC
#define PIP_SUBREG_NC 8
#endif

PIP_SDEV V pip_bcast(V v, int src)
{
  if (sizeof(V) == 8) return (V)W::shfl64((pip_i64)v, src);
  return (V)W::shfl((int)v, src);
}

PIP_SDEVNI int pip_subsolve_regs(const V *ctx, int cstride, int nc, int np, const V *trow, int mnvar, int mode, int critic,
                                 PipStats &st)
{
  enum { NC = PIP_SUBREG_NC, INF = 0x7fffffff };
  const int lane = W::lane();
  const int ncol = np + 1;
  const int extra = mode == 0 ? 0 : 1;
  int nl = np + nc + extra;
  if (ncol > NC || nl > 32) return -1;
  unsigned ovf = 0;
  /* ---- load: expanser_xx(context, nparm, nc, nparm+1, nparm, extra, 0), source/traiter.c:191-218 ---- */
  V row[NC];
  #pragma unroll
  for (int j = 0; j < NC; j++) row[j] = 0;
  V den = 1;
  int pos = lane < nl ? lane : INF;
  int flag = lane < np ? PIP_UNIT : lane < nl ? PIP_UNKNOWN : 0;
  int unit = lane;
  if (lane >= np && lane < np + nc) {
    const V *r = ctx + (lane - np) * cstride;
    #pragma unroll
    for (int j = 0; j < NC; j++) if (j < ncol) row[j] = r[j];
  } else if (extra && lane == np + nc) {
    #pragma unroll
    for (int j = 0; j < NC; j++) if (j < ncol) {
      V v = (j < np) ? trow[mnvar + 1 + j] : trow[mnvar];
      if (mode == 1) { if (j == np && !critic) v -= 1; }
      else { v = -v; if (j == np) v -= 1; }
      row[j] = v;
    }
  }
  pip_i64 det[PIP_MAX_DET];
  #pragma unroll
  for (int k = 0; k < PIP_MAX_DET; k++) det[k] = 0;
  det[0] = 1;
  int ldet = 1;
  unsigned pivots = 0, cuts = 0, max_rows = 0;
  unsigned long long elem = 0;
  /* counters are committed on every exit but the fall-back one (the general path then counts the solve) */
#define PIP_SUBREG_COMMIT() do { st.pivots += pivots; st.cuts += cuts; st.elem_updates += elem; \
    if (max_rows > st.max_rows) st.max_rows = max_rows; \
    if (pivots && (unsigned)ncol > st.max_cols) st.max_cols = ncol; } while (0)

  /* ---- tab_sort_rows_xx, source/traiter.c:556-623: at entry lane == position and no row position is Unit */
  {
    const bool isrow = lane >= np && lane < nl;
    unsigned sz = 0;
    if (isrow) {
      #pragma unroll
      for (int j = 0; j < NC - 1; j++) if (j < np) {
        const pip_u64 u = pip_uabs((pip_i64)row[j]);
        if (u < 2147483648ull && (unsigned)u > sz) sz = (unsigned)u;        /* den == 1 */
      }
    }
    const unsigned sb = pip_f2u((float)(double)sz);          /* size is stored as float */
    const unsigned smax_u = W::redmax(isrow ? sz : 0u);
    const double smax = (double)smax_u;
    /* nothing moves when the rows are already in non-decreasing order of size */
    unsigned pm = isrow ? sb : 0u;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = (unsigned)W::shfl_up((int)pm, o);
      if (lane >= o && y > pm) pm = y;
    }
    unsigned before = (unsigned)W::shfl_up((int)pm, 1);
    if (lane == 0) before = 0u;
    if (W::any(isrow && sb < before)) {
      const bool movable = isrow && (double)pip_u2f(sb) < smax;
      #pragma unroll 1
      for (int i = np; i < nl; i++) {
        /* first minimum strictly below smax among the positions >= i */
        const unsigned key = (movable && pos >= i) ? sb : 0xffffffffu;
        const unsigned m = W::redmin(key);
        if (m == 0xffffffffu) break;
        const int bestk = (int)W::redmin(key == m ? (unsigned)pos : 0xffffffffu);
        if (bestk != i) {                                   /* the two records trade positions */
          if (pos == i) pos = bestk;
          else if (pos == bestk) pos = i;
        }
      }
    }
  }

  #pragma unroll 1
  for (;;) {
    const bool isunit = (flag & PIP_UNIT) != 0;
    /* ---- chercher(Minus), then exam_coef without parameters: the sign of the constant, rows in position
     * order up to and including the first negative one (source/traiter.c:669-680, 118-157) ---- */
    int pivi = (int)W::redmin((flag & PIP_MINUS) ? (unsigned)pos : (unsigned)INF);
    if (pivi == INF) {
      const bool unk = flag == PIP_UNKNOWN;
      V cst = 0;
      #pragma unroll
      for (int j = 0; j < NC; j++) if (j == np) cst = row[j];
      const int ff = cst < 0 ? PIP_MINUS : cst > 0 ? PIP_PLUS : PIP_ZERO;
      const int first = (int)W::redmin((unk && ff == PIP_MINUS) ? (unsigned)pos : (unsigned)INF);
      if (unk && pos <= first) flag = ff;
      pivi = first;
    }
    if (pivi == INF) {
      /* ---- integrer_xx without parameters, source/integrer.c:305-534: first position below nvar whose
       * stored row needs a cut ---- */
      int verdict = 0, from = 0;
      #pragma unroll 1
      for (;;) {
        const bool cand = !isunit && pos < np && pos >= from && den != 1;
        const int i = (int)W::redmin(cand ? (unsigned)pos : (unsigned)INF);
        if (i == INF) break;
        const int Lc = pip_ffs(W::ballot(pos == i)) - 1;
        const V D = pip_bcast(den, Lc);
        if (D == 0) { PIP_SUBREG_COMMIT(); return PIP_ST_FAULT; }
        /* the residues, computed by the owner of the row, handed to every lane */
        V cut[NC];
        bool okv = false, okc = false;
        #pragma unroll
        for (int j = 0; j < NC; j++) {
          V x = 0;
          if (lane == Lc && j < ncol) {
            if (j < np) { x = pip_mod(row[j], D); okv = okv || x > 0; }
            else { x = -pip_mod(-row[j], D); okc = x != 0; }
          }
          cut[j] = pip_bcast(x, Lc);
        }
        const bool ok_var = W::any(okv), ok_const = W::any(okc);
        if (!ok_const) { from = i + 1; continue; }           /* case (a): this row is integral */
        if (!ok_var) { verdict = -1; break; }                /* case (b): no solution */
        if (nl >= 32) return -1;                             /* no lane left for the cut: general path */
        if (lane == nl) {
          #pragma unroll
          for (int j = 0; j < NC; j++) row[j] = cut[j];
          den = D; flag = PIP_MINUS; pos = nl; unit = 0;
        }
        verdict = nl + 1;                                    /* > 0: the position of the cut, plus one */
        nl++;
        cuts++;
        break;
      }
      if (verdict <= 0) {
        PIP_SUBREG_COMMIT();
        if (PipVal<V>::narrow && W::any(ovf != 0)) return PIP_ST_WIDEN;
        return verdict == 0 ? 1 : 0;
      }
      pivi = verdict - 1;
    }

    /* ---- pivoter_xx, source/traiter.c:345-548 ---- */
    const bool isu = (flag & PIP_UNIT) != 0;                /* (the cut lane changed its flag above) */
    const int Lp = pip_ffs(W::ballot(pos == pivi)) - 1;
    V prow[NC];
    #pragma unroll
    for (int j = 0; j < NC; j++) prow[j] = pip_bcast(row[j], Lp);
    const V dpiv = pip_bcast(den, Lp);
    /* choisir_piv_xx, source/traiter.c:297-341: candidates in increasing column order, the incumbent's
     * column entry of every record kept in a register */
    int pivj = -1;
    V pivot = 0, bcol = 0;
    #pragma unroll
    for (int j = 0; j < NC - 1; j++) {
      if (j < np && prow[j] > 0) {                           /* warp-uniform */
        const V a = isu ? ((unit == j) ? den : (V)0) : row[j];
        if (pivj < 0) { pivj = j; pivot = prow[j]; bcol = a; }
        else {
          const pip_i64 x = (pos < nl) ? PipVal<V>::cross(pivot, a, bcol, prow[j]) : 0;
          const int k = (int)W::redmin(x != 0 ? (unsigned)pos : (unsigned)INF);
          if (k != INF && W::any(x < 0 && pos == k)) { pivj = j; pivot = prow[j]; bcol = a; }
        }
      }
    }
    if (pivj < 0) {                                          /* no positive coefficient: infeasible */
      PIP_SUBREG_COMMIT();
      if (PipVal<V>::narrow && W::any(ovf != 0)) return PIP_ST_WIDEN;
      return 0;
    }
    /* determinant bookkeeping = the overflow verdict, source/traiter.c:394-447 (uniform, every lane) */
    {
      const bool trivial = dpiv == 1 && pivot == 1 && det[0] > -(1ll << 61) && det[0] < (1ll << 61);
      if (!trivial) {
        const pip_i64 d = (dpiv == 1) ? 1 : pip_gcd((pip_i64)pivot, (pip_i64)dpiv);
        if (d == 0) { PIP_SUBREG_COMMIT(); return PIP_ST_FAULT; }
        pip_i64 ppivot = pivot, dppiv = dpiv;
        if (d != 1) { ppivot = pip_div((pip_i64)pivot, d); dppiv = pip_div((pip_i64)dpiv, d); }
        #pragma unroll
        for (int i = 0; i < PIP_MAX_DET; i++) {
          if (i < ldet && dppiv != 1) {
            const pip_i64 g = pip_gcd(det[i], dppiv);
            if (g == 0) { PIP_SUBREG_COMMIT(); return PIP_ST_FAULT; }
            if (g != 1) { det[i] = pip_div(det[i], g); dppiv = pip_div(dppiv, g); }
          }
        }
        if (dppiv != 1) { PIP_SUBREG_COMMIT(); return PIP_ST_FATAL + 1; }   /* "Integer overflow" */
        const int bp = pip_bitlen(ppivot);
        bool placed = false;
        #pragma unroll
        for (int i = 0; i < PIP_MAX_DET; i++) {
          if (!placed && i < ldet && pip_bitlen(det[i]) + bp < 64) { det[i] = (pip_i64)((pip_u64)det[i] * (pip_u64)ppivot); placed = true; }
        }
        if (!placed) {
          ldet++;
          if (ldet >= PIP_MAX_DET) { PIP_SUBREG_COMMIT(); return PIP_ST_FATAL + 1; }   /* "Integer overflow : 4" */
          #pragma unroll
          for (int i = 0; i < PIP_MAX_DET; i++) if (i == ldet - 1) det[i] = ppivot;
        }
      }
    }
    pivots++;
    if ((unsigned)nl > max_rows) max_rows = nl;
    elem += (unsigned long long)(nl - np - 1) * ncol;
    /* the Unit record that owns column pivj (source/traiter.c:503-516) */
    const unsigned um = W::ballot(isu && unit == pivj && pos < nl);
    if (um == 0) { PIP_SUBREG_COMMIT(); return PIP_ST_FAULT; }
    const int Lu = pip_ffs(um) - 1;
    const int ku = W::shfl(pos, Lu);
    /* rank-1 update of every stored row but the pivot row + re-flag (source/traiter.c:467-502, 518-529) */
    bool fault = false;
    if (!isu && lane != Lp && pos < nl) {
      V foo = bcol;
      if (!(foo == 0 && den == 1)) {
        V lpiv = pivot;
        if (foo == 0) lpiv = 1;
        else if (pivot != 1 && foo != 1 && foo != -1) {
          const V d = pip_gcd(pivot, foo);
          if (d != 1) { lpiv = pip_div(pivot, d); foo = pip_div(foo, d); }
        }
        const V newden = PipVal<V>::mul(lpiv, den, ovf);
        V g = newden;
        const V zp = PipVal<V>::mul(dpiv, foo, ovf);
        pip_u64 orz = 0;
        #pragma unroll
        for (int j = 0; j < NC; j++) {
          V z = PipVal<V>::mulsub(row[j], lpiv, prow[j], foo, ovf);
          if (j == pivj) z = zp;
          row[j] = z;
          orz |= (pip_u64)(pip_i64)z;
        }
        if (g != 1) {
          if ((g & (g - 1)) == 0 && g > 0) { orz |= (pip_u64)(pip_i64)g; g = (V)(pip_i64)(orz & (0ull - orz)); }
          else {
            #pragma unroll
            for (int j = 0; j < NC; j++) if (j < ncol && g != 1) g = pip_gcd(g, row[j]);
          }
        }
        if (g != 1) {
          if (g == 0) fault = true;
          else if ((g & (g - 1)) == 0) {
            int sh = 0;
            while (((pip_u64)(pip_i64)g >> sh) != 1ull) sh++;
            #pragma unroll
            for (int j = 0; j < NC; j++) row[j] = row[j] >> sh;
            den = newden >> sh;
          } else {
            const PipExactDiv e = pip_exact_prepare((pip_i64)g);
            #pragma unroll
            for (int j = 0; j < NC; j++) row[j] = (V)pip_exact_apply((pip_i64)row[j], e);
            den = (V)pip_exact_apply((pip_i64)newden, e);
          }
        } else den = newden;
        if (!fault) {
          int ff = PIP_FLAG(flag);
          const int fff = zp < 0 ? PIP_MINUS : zp == 0 ? PIP_ZERO : PIP_PLUS;
          if (fff != PIP_ZERO && fff != ff) {
            if (ff == PIP_ZERO) ff = (fff == PIP_MINUS ? PIP_UNKNOWN : fff);
            else ff = PIP_UNKNOWN;
            flag = ff;
          }
        }
      }
    }
    if (W::any(fault)) { PIP_SUBREG_COMMIT(); return PIP_ST_FAULT; }
    if (PipVal<V>::narrow && W::any(ovf != 0)) return PIP_ST_WIDEN;
    /* the pivot row's storage goes to position ku, position pivi becomes Unit for pivj */
    if (lane == Lp) {
      #pragma unroll
      for (int j = 0; j < NC; j++) row[j] = (j == pivj) ? dpiv : (V)(-prow[j]);
      den = pivot; flag = PIP_PLUS; pos = ku;
    } else if (lane == Lu) {
      flag = PIP_UNIT | PIP_ZERO; unit = pivj; den = 1; pos = pivi;
    }
  }
#undef PIP_SUBREG_COMMIT
}

/* ---- problem load: source/tab.c:222-248 (tab_get) + tab_simplify when an integer solution is wanted ----
 * Fills the arena regions a fresh problem consists of: den | fl of the nvar + ni positions, the ni stored rows,
 * the nc context rows.  Runs inside the solver (general path) or, for dense batches, ahead of it in
 * pip_image_kernel with B = the problem's ARENA IMAGE in global memory: the solver then starts from two block
 * copies and none of this code is in its instruction stream.  Returns the int32 range flag of the loads. */
PIP_SDEVNI unsigned pip_load_problem(const PipProblem &P, const void *pool, int elem_log2, pip_i64 *B, const PipTab T,
                                     int ctx_off, int cstride)
{
  const int lane = W::lane();
  const int ncol = P.nvar + P.nparm + 1;
  int *fl = pip_fl(B, T);
  V *den = pip_den(B, T);
  V *ctx = (V *)(B + ctx_off);
  unsigned ovf = 0;
  #pragma unroll 1
  for (int k = lane; k < P.nvar + P.ni; k += 32) {
    if (k < P.nvar) { fl[k] = PIP_MKFL(PIP_UNIT, k); den[k] = 1; }
    else { fl[k] = PIP_MKFL(PIP_UNKNOWN, k - P.nvar); den[k] = 1; }
  }
  ovf |= pip_load2d((V *)(B + T.data), T.stride, pool, elem_log2, P.off, P.ni, ncol);
  ovf |= pip_load2d(ctx, cstride, pool, elem_log2, P.off + (pip_i64)P.ni * ncol, P.nc, P.nparm + 1);
  W::sync();
  if (PipVal<V>::narrow && W::any(ovf != 0)) return 1u;
  if (P.flags & PIP_F_INT) {
    pip_simplify_rows((V *)(B + T.data), P.ni, T.stride, ncol, P.nvar);
    pip_simplify_rows(ctx, P.nc, cstride, P.nparm + 1, P.nparm);
    W::sync();
  }
  return 0u;
}

/* The solver for one problem.  `B` is the warp's working arena (`words` words), `out` the
 * warp's cell window (at least sol_size cells free), `stk` the warp's frame stack. */
PIP_SDEV void pip_solve_one(const PipProblem &P, const void *pool, int elem_log2, pip_i64 *B, int words, int slack_level,
                           PipCell *out, pip_i64 *stk, pip_i64 stk_cap,
                           int sol_size, int maxcol, int maxparm,
                           int &status_out, int &ncell_out, unsigned &rflags_out, PipStats &st, PipTeam *tm = nullptr,
                           unsigned *nwords_out = nullptr, bool wordmode_arg = false, const PipLayout *pre = nullptr,
                           const PipSteal *stl = nullptr, int stl_problem = 0, int stl_seg = -1,
                           const pip_i64 *resume = nullptr, int *hwm_out = nullptr,
                           const pip_i64 *image = nullptr, int image_w1 = 0, unsigned budget = 0,
                           const unsigned *handed = nullptr, unsigned handed_max = 0)
{
  const int lane = W::lane();
  const bool integer = (P.flags & PIP_F_INT) != 0;
  const bool wordmode = WORDS || wordmode_arg;
  PipWordOut wo;
  wo.w = (V *)out; wo.pos = 0; wo.node_at = -1; wo.nnew = 0; wo.wide = false;
  PipLayout L;
  /* dense batches: one shape for the whole launch, the arena was carved once on the host (`pre`, for the
   * largest row counts of the batch; capacities only ever decide CAPACITY, never an answer).  (Copied into
   * registers: reading it in place through a pointer put every access into local memory -- 130 -> 153 ms.) */
  if (pre && pre->m.nvar == P.nvar && pre->m.nparm == P.nparm && P.ni <= pre->m.ni && P.nc + 1 <= pre->s.ni) {
    L = *pre;
    L.m.ni = P.ni;
    L.s.ni = 0;
  } else {
    int level_try = slack_level;
    while (!pip_layout(P.nvar, P.nparm, P.ni, P.nc, P.flags, level_try, words, (int)sizeof(V), L)) {
      level_try = level_try == PIP_LEVEL_S_WIDE ? 2 : level_try - 1;
      if (level_try < 0) { status_out = PIP_ST_CAPACITY; ncell_out = 0; return; }
    }
  }
  /* device-converted input that does not fit the int32 pool (pip_convert.h): only the int64 pool, built
   * on demand by the host, holds this problem */
  if ((P.flags & PIP_F_WIDE_INPUT) && elem_log2 < 3) { status_out = PIP_ST_WIDEN; ncell_out = 0; return; }
  /* Compute_dual with parameters: the reference re-sorts the copy made at a split with Unit rows in
   * the constraint range and reads ineq[] entries it never wrote (source/traiter.c:585 vs 616-617),
   * so its answer is undefined; only the split-free case is implemented */
  if ((P.flags & PIP_F_DUAL) && (P.nparm > 0 || P.nc > 0)) { status_out = PIP_ST_UNSUPPORTED; ncell_out = 0; return; }
  /* a big-parameter column outside the tableau (test/challenges/pipFile_1: column 12 of 12): the reference reads
   * past the row (source/traiter.c:111) and its answer depends on the heap; refused instead of answered at random */
  if (P.bigparm >= P.nvar + P.nparm + 1) { status_out = PIP_ST_UNSUPPORTED; ncell_out = 0; return; }
  /* the rarely-set options live in the global-memory instantiations only (TEAM), so that the code of
   * the instruction-supply-bound shared-memory kernels is not touched by them; a problem that asks
   * for one here is handed to the next class like any other that does not fit */
  if (!TEAM && (P.flags & (PIP_F_DUAL | PIP_F_DEEPEST))) { status_out = PIP_ST_CAPACITY; ncell_out = 0; return; }
  const bool dual = TEAM && (P.flags & PIP_F_DUAL) != 0 && !integer;

  PipTab T = L.m, M = L.m;       /* current tableau, saved main tableau while in a sub-solve */
  int level = 0;                  /* 0 = main problem, 1 = compatibility / context sub-solve */
  int nc = P.nc;
  int ncell = 0, status = PIP_ST_OK;
  /* words of the serialised quast (pip_decode.h) when no column surgery applies (PIP_F_SIMPLE_SER):
   * a node costs 2 (newparm count, kind), a vector of n forms 1 + 2n */
  unsigned nwords = 0;
  int ret_site = 0, ci = 0, cplus = 0, critic = 0, pivi = 0, depth = 0;
  bool feasible = false, wide = false;
  rflags_out = 0;
  pip_i64 top = 0;
  unsigned ovf = 0;                /* int32 instantiation: some value left the 31-bit range */
  /* subtree donation (PipSteal, pip_types.h) */
  int hwm = 0;                     /* largest cell count this segment checked against SOL_SIZE */
  pip_i64 top_base = 0;            /* frames below this stack offset were given away */
  int my_offer = -1;               /* the one outstanding offer of this segment */
  const pip_i64 *Fr = nullptr;     /* the frame being restored */
/* the SOL_SIZE check of sol_alloc (source/sol.c:96-100); a segment of a donated subtree does not know how many
 * cells precede it in pre-order, so it also records the largest count it checked (resolved by the copy kernel) */
#define PIP_NEED(x) do { const int need_ = ncell + (x); if (STEAL) hwm = need_ > hwm ? need_ : hwm; \
                         if (need_ >= sol_size) { status = PIP_ST_FATAL + 26; goto DONE; } } while (0)
  V *ctx = (V *)(B + L.ctx);
  V *cut = (V *)(B + L.cut);
  const int cstride = L.cstride;

  /* a donated subtree: no load, the state comes out of the donor's frame */
  if (STEAL && resume) { Fr = resume; goto RESTORE; }

  /* ---- load: the arena image prepared by pip_image_kernel (two block copies), or the general loader */
  if (image) {
    pip_copy_words(B + T.den, image, (T.data - T.den) + (int)(((pip_i64)P.ni * T.stride * (pip_i64)sizeof(V) + 7) / 8));
    pip_copy_words(B + L.ctx, image + image_w1, (int)(((pip_i64)P.nc * cstride * (pip_i64)sizeof(V) + 7) / 8));
    if (lane == 0) { B[L.m.det] = 1; B[L.s.det] = 1; }
    W::sync();
  } else {
    if (pip_load_problem(P, pool, elem_log2, B, T, L.ctx, cstride)) { status = PIP_ST_WIDEN; goto DONE; }
    if (lane == 0) { B[L.m.det] = 1; B[L.s.det] = 1; }
    W::sync();
  }

  PIP_LAP(st, PIP_PH_LOAD);
  /* context emptiness check, source/piplib.c:818-826 / source/maind.c:199-204 */
  if (nc > 0) { ret_site = 0; goto BUILD_SUB; }
  goto ENTRY;

BUILD_SUB:
  /* expanser_xx(context, nparm, nc, nparm+1, nparm, extra, 0) (source/traiter.c:191-199,
   * 211-218): nparm Unit positions, the nc context rows, and for ret_site 1/2 the tested row */
  {
    M = T;
    PipTab S = L.s;
    const int np = M.nparm;
    const int extra = ret_site == 0 ? 0 : 1;
#ifdef PIP_USE_SUBREG
    {
      /* the register-resident form first (pip_subsolve_regs); -1 = does not fit, take the general path */
      const V *trow = extra ? pip_row(B, M, PIP_LINK(pip_fl(B, M)[ci])) : (const V *)nullptr;
      /* (Deepest_cut rewrites the constant cuts of the sub-solves too, source/integrer.c:417-438: general path) */
      const int r = (P.flags & PIP_F_DEEPEST) ? -1 : pip_subsolve_regs(ctx, cstride, nc, np, trow, M.nvar, ret_site, critic, st);
      if (r >= 0) {
        st.subsolves++;
        PIP_LAP(st, PIP_PH_BUILDSUB);
        if (r > 1) { status = r; goto DONE; }
        feasible = r == 1;
        T = S; T.nvar = np; T.nparm = 0;           /* SUB_DONE sizes the transient cells from the sub tableau */
        level = 1;
        goto SUB_DONE;
      }
    }
#endif
    S.nvar = np; S.nparm = 0; S.ni = nc + extra; S.ldet = 1;
    if (np + nc + extra > S.pcap || nc + extra > S.rcap || np + 1 > S.stride) { status = PIP_ST_CAPACITY; goto DONE; }
    int *sfl = pip_fl(B, S);
    V *sden = pip_den(B, S);
    #pragma unroll 1
    for (int k = lane; k < np + nc + extra; k += 32) {
      sfl[k] = k < np ? PIP_MKFL(PIP_UNIT, k) : PIP_MKFL(PIP_UNKNOWN, k - np);
      sden[k] = 1;
    }
    /* the sub tableau's rows have the context's stride (pip_layout: both XC), so the nc context rows are one
     * block of arena words (an odd count of int32 values spills one value into the slot of row nc, which is
     * written next or never read) */
    pip_copy_words(B + S.data, B + L.ctx, (int)(((pip_i64)nc * cstride * (pip_i64)sizeof(V) + 7) / 8));
    if (sizeof(V) < 8) W::sync();             /* (the spilled value and the tested row's first entry share a word) */
    if (extra) {
      const int f = pip_fl(B, M)[ci];
      const V *row = pip_row(B, M, PIP_LINK(f));
      V *nr = pip_row(B, S, nc);
      #pragma unroll 1
      for (int j = lane; j <= np; j += 32) {
        V v = (j < np) ? row[M.nvar + 1 + j] : row[M.nvar];
        if (ret_site == 1) { if (j == np && !critic) v -= 1; }
        else { v = -v; if (j == np) v -= 1; }
        nr[j] = v;
      }
    }
    if (lane == 0) B[S.det] = 1;
    T = S;
    level = 1;
    W::sync();
    PIP_LAP(st, PIP_PH_BUILDSUB);
  }

ENTRY:
  /* traiter_xx entry, source/traiter.c:643-656 (the private context copy is implicit) */
  if (level) st.subsolves++;
  if (dual && level == 0) pip_sort_rows_dual(B, T, L.tmp, (int *)(B + L.total - (L.m.rcap + 1) / 2));
  else pip_sort_rows(B, T, L.tmp);
  PIP_LAP(st, PIP_PH_SORT);

LOOP:
  {
    const int nl = T.nvar + T.ni;
#ifndef PIP_NO_SCAN32
    if (nl <= 32) pivi = pip_scan32(B, T, level ? -1 : P.bigparm);
    else
#endif
    {
      const int bg = level ? -1 : P.bigparm;
      if (TEAM && tm != nullptr && nl >= PIP_TEAM_MIN_SCAN) {
        pivi = pip_team_scan(tm, B, T, PIP_TEAM_FIRST, 0, PIP_MINUS, 0, nl);
        if (pivi >= nl) pivi = bg >= 0 ? pip_exam_coef(B, T, bg) : pip_team_scan(tm, B, T, PIP_TEAM_EXAM, 0, 0, 0, nl);
      } else {
        pivi = pip_first_flag(B, T, PIP_MINUS, 0, nl);
        if (pivi >= nl) pivi = pip_exam_coef(B, T, bg);
      }
    }
    PIP_LAP(st, PIP_PH_SCAN);
    if (pivi < nl) goto PIVOT;
    if (T.nparm == 0) goto NONNEG;
    /* compa_test_xx, source/traiter.c:162-243 */
    if (T.nparm >= maxparm) { status = PIP_ST_FATAL + 1; goto DONE; }
    ci = 0;
  }

COMPA_NEXT:
  {
    const int nl = T.nvar + T.ni;
    ci = pip_first_flag(B, T, PIP_CRITIC | PIP_UNKNOWN, ci, nl);
    if (ci >= nl) goto AFTER_COMPA;
    const V *row = pip_row(B, T, PIP_LINK(pip_fl(B, T)[ci]));
    bool pos = false;
    #pragma unroll 1
    for (int j = lane; j < T.nvar; j += 32) pos = pos || row[j] > 0;
    critic = W::any(pos) ? 0 : 1;
    ret_site = 1;
    goto BUILD_SUB;
  }

SUB_DONE:
  {
    /* the sub-solve's transient cells count against SOL_SIZE (source/sol.c:96-100) */
    const int need = feasible ? 1 + 2 * T.nvar : 1;
    PIP_NEED(need);
    T = M;
    level = 0;
    if (ret_site == 0) {
      if (!feasible) { status = PIP_ST_VOID; goto DONE; }
      goto ENTRY;
    }
    if (ret_site == 1) { cplus = feasible; ret_site = 2; goto BUILD_SUB; }
    {
      int *fl = pip_fl(B, T);
      const int f = fl[ci];
      int nf;
      const bool cminus = feasible;
      if (cplus && cminus) nf = critic ? PIP_CRITIC : PIP_UNKNOWN;
      else if (cminus) nf = PIP_MINUS;
      else nf = cplus ? PIP_PLUS : PIP_ZERO;
      W::sync();
      if (lane == 0) fl[ci] = PIP_MKFL(nf, PIP_LINK(f));
      W::sync();
      if (nf == PIP_MINUS) goto AFTER_COMPA;
      ci++;
      goto COMPA_NEXT;
    }
  }

AFTER_COMPA:
  {
    const int nl = T.nvar + T.ni;
    pivi = pip_first_flag(B, T, PIP_MINUS, 0, nl);
    if (pivi < nl) goto PIVOT;
    pivi = pip_first_flag(B, T, PIP_CRITIC, 0, nl);
    if (pivi >= nl) pivi = pip_first_flag(B, T, PIP_UNKNOWN, 0, nl);
    if (pivi >= nl) goto NONNEG;
  }
  /* split, source/traiter.c:695-759 */
  {
    const int np = T.nparm;
    if (nc >= L.crcap) { status = PIP_ST_CAPACITY; goto DONE; }
    if (np >= maxparm) { status = PIP_ST_FATAL + 2; goto DONE; }
    /* heavy-problem hand-over (PipLaunch::budget): a big tree is better solved by many warps */
    if (!STEAL && budget && st.pivots > budget) {
      /* ... unless the whole batch is like that (handed_max): then the plain launch is the right one */
      unsigned h = 0;
      if (lane == 0) h = W::load_volatile(handed);
      if ((unsigned)W::shfl((int)h, 0) < handed_max) { status = PIP_ST_PENDING; goto DONE; }
      budget = 0;
    }
    PIP_NEED(np + 3);
    int *fl = pip_fl(B, T);
    const V *row = pip_row(B, T, PIP_LINK(fl[pivi]));
    V g = 0;
    #pragma unroll 1
    for (int j = 0; j < np; j++) g = pip_gcd(g, row[T.nvar + 1 + j]);
    if (!integer) g = pip_gcd(g, row[T.nvar]);
    if (g == 0) { status = PIP_ST_FAULT; goto DONE; }
    V *crow = ctx + nc * cstride;
    #pragma unroll 1
    for (int j = lane; j <= np; j += 32) {
      V v;
      if (j < np) v = pip_div(row[T.nvar + 1 + j], g);
      else v = integer ? pip_floor_q(row[T.nvar], g) : pip_div(row[T.nvar], g);
      crow[j] = v;
      if (!wordmode) wide = pip_put(out, ncell + 2 + j, PIP_C_VAL, v, 1) || wide;
    }
    if (wordmode) {
      /* `nnew 2 VEC(condition)`: the vector is np + 1 values over the denominator 1 */
      pip_wnode_kind(wo, 2);
      if (lane == 0) pip_wout(wo, wo.pos, np + 1);
      #pragma unroll 1
      for (int j = lane; j <= np; j += 32) { pip_wout(wo, wo.pos + 1 + 2 * j, (pip_i64)crow[j]); pip_wout(wo, wo.pos + 2 + 2 * j, 1); }
      wo.pos += 1 + 2 * (np + 1);
    } else if (lane == 0) {
      pip_put(out, ncell, PIP_C_IF, 0, 0);
      pip_put(out, ncell + 1, PIP_C_FORM, np + 1, 0);
    }
    ncell += np + 3;
    nwords += 2 * np + 5;
    W::sync();
    /* push the ELSE continuation: 12 header words, then two block copies of arena words -- the main tableau
     * as it lies in the arena (den | fl | scratch | the ni stored rows with their stride: one contiguous
     * region, pip_layout) and the nc+1 context rows -- then the frame size.  (A compact frame was four index
     * loops here and four in the pop: 300 instructions of a kernel whose executed code has to fit the 32 KB
     * instruction cache; two word-copy loops cost twice the frame bytes and a fifth of the code.) */
    {
#define PIP_VW(n) (((pip_i64)(n) * (pip_i64)sizeof(V) + 7) / 8)
      const int w1 = (T.data - T.den) + (int)PIP_VW((pip_i64)T.ni * T.stride);
      const int w2 = (int)PIP_VW((pip_i64)(nc + 1) * cstride);
      const pip_i64 fsize = 12 + (pip_i64)w1 + w2 + 1;
      if (top + fsize > stk_cap) { status = PIP_ST_CAPACITY; goto DONE; }
      pip_i64 *F = stk + top;
      if (lane == 0) {
        F[0] = T.nvar; F[1] = np; F[2] = T.ni; F[3] = nc; F[4] = pivi; F[5] = T.ldet; F[6] = fsize;
        #pragma unroll 1
        for (int k = 0; k < PIP_MAX_DET; k++) F[8 + k] = B[T.det + k];
        F[fsize - 1] = fsize;
      }
      pip_copy_words(F + 12, B + T.den, w1);
      pip_copy_words(F + 12 + w1, B + L.ctx, w2);
      top += fsize;
    }
    if (STEAL && stl) pip_offer_bottom(*stl, stl_problem, stl_seg, stk, top, top_base, my_offer);
    W::sync();
    if (lane == 0) fl[pivi] = PIP_MKFL(PIP_PLUS, PIP_LINK(fl[pivi]));
    W::sync();
    nc++;
    depth++;
    st.splits++;
    PIP_LAP(st, PIP_PH_FRAME);
    goto ENTRY;
  }

NONNEG:
  if (level == 0 && !integer) {
    const int total = 1 + T.nvar * (T.nparm + 2);
    PIP_NEED(total);
    if (wordmode) pip_emit_solution_words(B, T, wo);
    else wide = pip_emit_solution(B, T, out, ncell) || wide;
    ncell += total;
    nwords += T.nvar ? 4 + T.nvar * (2 * T.nparm + 4) : 5;
    if (dual) {
      const int dtotal = 1 + 2 * T.ni;
      PIP_NEED(dtotal);
      wide = pip_emit_dual(B, T, (const int *)(B + L.total - (L.m.rcap + 1) / 2), out, ncell) || wide;
      ncell += dtotal;
    }
    goto LEAF;
  }
  /* integrer_xx, source/integrer.c:305-534 */
  {
    const int nvar = T.nvar, np = T.nparm, ncol = nvar + np + 1, nl = nvar + T.ni;
    if (ncol >= maxcol) { status = PIP_ST_FATAL + 3; goto DONE; }
    int *fl = pip_fl(B, T);
    V *den = pip_den(B, T);
    int verdict = 0;     /* 0 integral, -1 none, >0 cut row */
    #pragma unroll 1
    for (int i = 0; i < nvar; i++) {
      const V D = den[i];
      const int f = fl[i];
      if (D == 1) continue;
      if (f & PIP_UNIT) continue;
      if (D == 0) { status = PIP_ST_FAULT; goto DONE; }
      const V *row = pip_row(B, T, PIP_LINK(f));
      bool okv = false, okc = false, okp = false;
      W::sync();
      #pragma unroll 1
      for (int j = lane; j < ncol; j += 32) {
        V v = row[j], x;
        if (j < nvar) { x = pip_mod(v, D); okv = okv || x > 0; }
        else if (j == nvar) { x = -pip_mod(-v, D); okc = okc || x != 0; }
        else if (!level && j == P.bigparm) x = 0;
        else { x = -pip_mod(-v, D); okp = okp || x != 0; }
        cut[j] = x;
      }
      if (lane == 0) cut[ncol] = D;
      const bool ok_var = W::any(okv), ok_const = W::any(okc), ok_parm = W::any(okp);
      W::sync();
      if (!ok_parm && !ok_const) continue;                  /* case (a) */
      if (!ok_parm && !ok_var) { verdict = -1; break; }     /* case (b) */
      if (T.ni >= T.rcap || nl >= T.pcap) { status = PIP_ST_CAPACITY; goto DONE; }
      if (TEAM && (P.flags & PIP_F_DEEPEST) && !ok_parm && !pip_deepest_cut(cut, nvar, D, ovf)) { status = PIP_ST_FAULT; goto DONE; }
      int parm = -1;
      if (ok_parm) {                                         /* case (e), source/integrer.c:493-520 */
        if (lane == 0) parm = pip_find_parm(ctx, cstride, nc, np, cut + nvar);
        W::sync();
        parm = W::shfl(parm, 0);
        if (parm == -1) {
          /* add_parm_xx, source/integrer.c:156-227 */
          if (nc + 2 > L.crcap || np + 2 > cstride || ncol + 1 > T.stride) { status = PIP_ST_CAPACITY; goto DONE; }
          PIP_NEED(np + 5);
          const V *c = cut + nvar;
          if (wordmode) {
            /* `rank deno VEC`: the new parameter's rank, its divisor, the np + 1 values of the form over 1 */
            pip_wnode_open(wo);
            if (lane == 0) {
              const unsigned b0 = wo.pos;
              pip_wout(wo, b0, np); pip_wout(wo, b0 + 1, (pip_i64)c[1 + np]); pip_wout(wo, b0 + 2, np + 1);
              #pragma unroll 1
              for (int j = 0; j < np; j++) { pip_wout(wo, b0 + 3 + 2 * j, -(pip_i64)c[1 + j]); pip_wout(wo, b0 + 4 + 2 * j, 1); }
              pip_wout(wo, b0 + 3 + 2 * np, -(pip_i64)c[0]); pip_wout(wo, b0 + 4 + 2 * np, 1);
            }
            wo.pos += 2 * np + 5;
            wo.nnew++;
          }
          if (lane == 0) {
            if (!wordmode) {
            pip_put(out, ncell, PIP_C_NEW, np, 0);
            pip_put(out, ncell + 1, PIP_C_DIV, 0, 0);
            pip_put(out, ncell + 2, PIP_C_FORM, np + 1, 0);
            #pragma unroll 1
            for (int j = 0; j < np; j++) wide = pip_put(out, ncell + 3 + j, PIP_C_VAL, -c[1 + j], 1) || wide;
            wide = pip_put(out, ncell + 3 + np, PIP_C_VAL, -c[0], 1) || wide;
            wide = pip_put(out, ncell + 4 + np, PIP_C_VAL, c[1 + np], 1) || wide;
            }
            #pragma unroll 1
            for (int k = 0; k < nc; k++) { V *r = ctx + k * cstride; r[np + 1] = r[np]; r[np] = 0; }
            V *r0 = ctx + nc * cstride, *r1 = r0 + cstride;
            #pragma unroll 1
            for (int j = 0; j < np; j++) { r0[j] = -c[1 + j]; r1[j] = c[1 + j]; }
            r0[np] = -c[1 + np]; r1[np] = c[1 + np];
            r0[np + 1] = -c[0]; r1[np + 1] = PipVal<V>::store((pip_i64)c[0] - 1 + (pip_i64)c[1 + np], ovf);
          }
          /* the new parameter's tableau column starts at zero in every stored row */
          #pragma unroll 1
          for (int r = lane; r < T.ni; r += 32) pip_row(B, T, r)[ncol] = 0;
          ncell += np + 5;
          nwords += 2 * np + 5;
          parm = np;
          T.nparm = np + 1;
          nc += 2;
          W::sync();
        }
        if (!ok_var) { status = PIP_ST_FATAL + 134; goto DONE; }   /* assert(ok_var) */
      }
      {
        V *nr = pip_row(B, T, T.ni);
        #pragma unroll 1
        for (int j = lane; j < ncol; j += 32) nr[j] = cut[j];
        W::sync();
        if (lane == 0) {
          if (ok_parm) {
            if (parm == np) nr[ncol] = cut[ncol];           /* fresh column: 0 + D */
            else nr[nvar + 1 + parm] = PipVal<V>::store((pip_i64)nr[nvar + 1 + parm] + (pip_i64)cut[ncol], ovf);
          }
          fl[nl] = PIP_MKFL(PIP_MINUS, T.ni);
          den[nl] = D;
        }
        T.ni++;
        st.cuts++;
        verdict = nl;
        W::sync();
        if (PipVal<V>::narrow && W::any(ovf != 0)) { status = PIP_ST_WIDEN; goto DONE; }
      }
      break;
    }
    PIP_LAP(st, PIP_PH_CUT);
    if (verdict > 0) { pivi = verdict; goto PIVOT; }
    if (level) { feasible = (verdict == 0); goto SUB_DONE; }
    if (verdict == 0) {
      const int total = 1 + T.nvar * (T.nparm + 2);
      PIP_NEED(total);
      if (wordmode) pip_emit_solution_words(B, T, wo);
      else wide = pip_emit_solution(B, T, out, ncell) || wide;
      ncell += total;
      nwords += T.nvar ? 4 + T.nvar * (2 * T.nparm + 4) : 5;
      PIP_LAP(st, PIP_PH_EMIT);
    } else {
      PIP_NEED(1);
      if (wordmode) pip_wnode_kind(wo, 0);
      else if (lane == 0) pip_put(out, ncell, PIP_C_NIL, 0, 0);
      ncell += 1;
      nwords += 2;
    }
    goto LEAF;
  }

PIVOT:
  {
    PIP_LAP(st, PIP_PH_OTHER);
    int rc = pip_pivot(B, T, pivi, st, tm);
    if (rc == 0) goto LOOP;
    if (rc > 0) { status = rc; goto DONE; }
    if (level) { feasible = false; goto SUB_DONE; }
    PIP_NEED(1);
    if (wordmode) pip_wnode_kind(wo, 0);
    else if (lane == 0) pip_put(out, ncell, PIP_C_NIL, 0, 0);
    ncell += 1;
    nwords += 2;
  }

LEAF:
  /* a branch of the main problem is finished: resume the innermost pending ELSE branch
   * (source/traiter.c:747-758) or stop */
  if (STEAL ? top == top_base : depth == 0) goto DONE;
  {
    W::sync();
    const pip_i64 fsize = stk[top - 1];
    if (STEAL && my_offer >= 0 && top - fsize == top_base) {
      /* the frame on offer: take it back, unless an idle warp took it -- then that subtree is its segment */
      int old = PIP_OFFER_OPEN;
      if (lane == 0) old = W::atomic_cas(&stl->offers[my_offer].state, PIP_OFFER_OPEN, PIP_OFFER_RECLAIMED);
      old = W::shfl(old, 0);
      const int idx_ = my_offer;
      my_offer = -1;
      if (old != PIP_OFFER_OPEN) { pip_offer_wait_copied(*stl, idx_); goto DONE; }
    }
    Fr = stk + (top - fsize);
    top -= fsize;
    depth--;
  }

RESTORE:
  {
    const pip_i64 *F = Fr;
#define PIP_FW(i) (STEAL ? W::load_cg(F + (i)) : F[i])      /* a donated frame was written by another SM */
    T.nvar = (int)PIP_FW(0); T.nparm = (int)PIP_FW(1); T.ni = (int)PIP_FW(2); nc = (int)PIP_FW(3); pivi = (int)PIP_FW(4);
    T.ldet = (int)PIP_FW(5);
    const int np = T.nparm;
    int *fl = pip_fl(B, T);
    if (lane == 0) for (int k = 0; k < PIP_MAX_DET; k++) B[T.det + k] = PIP_FW(8 + k);
#undef PIP_FW
    const int w1 = (T.data - T.den) + (int)PIP_VW((pip_i64)T.ni * T.stride);
    if (STEAL) {
      pip_copy_frame_in(B + T.den, F + 12, w1);
      pip_copy_frame_in(B + L.ctx, F + 12 + w1, (int)PIP_VW((pip_i64)(nc + 1) * cstride));
    } else {
      pip_copy_words(B + T.den, F + 12, w1);
      pip_copy_words(B + L.ctx, F + 12 + w1, (int)PIP_VW((pip_i64)(nc + 1) * cstride));
    }
    W::sync();
    #pragma unroll 1
    for (int j = lane; j <= np; j += 32) {                 /* the negated condition */
      V v = ctx[nc * cstride + j];
      ctx[nc * cstride + j] = (j < np) ? -v : -(v + 1);
    }
    W::sync();
    if (lane == 0) fl[pivi] = PIP_MKFL(PIP_MINUS, PIP_LINK(fl[pivi]));
    W::sync();
    if (STEAL && resume && Fr == resume && stl && lane == 0) {     /* a donated frame: the donor may have its stack back */
      W::fence();
      W::atomic_exch(&stl->offers[stl_seg].state, PIP_OFFER_COPIED);
    }
    nc++;
    PIP_LAP(st, PIP_PH_FRAME);
    goto PIVOT;
  }

DONE:
  if (STEAL && my_offer >= 0) {
    int old = PIP_OFFER_OPEN;
    if (lane == 0) old = W::atomic_cas(&stl->offers[my_offer].state, PIP_OFFER_OPEN, PIP_OFFER_RECLAIMED);
    if (W::shfl(old, 0) != PIP_OFFER_OPEN) pip_offer_wait_copied(*stl, my_offer);
  }
  if (hwm_out) *hwm_out = hwm;
  W::sync();
  PIP_LAP(st, PIP_PH_OTHER);
  status_out = status;
  ncell_out = (status == PIP_ST_OK) ? ncell : 0;
  if (nwords_out) *nwords_out = status == PIP_ST_OK ? nwords : status == PIP_ST_VOID ? 1u : 0u;
  if (wordmode) {
    /* an empty context is the one word -1 */
    if (status == PIP_ST_VOID) { wo.wide = false; if (lane == 0) pip_wout(wo, 0, -1); }
    /* the solver's own count (nwords) and the words written must agree: a mismatch is a bug, not a verdict */
    if (status == PIP_ST_OK && wo.pos != nwords) status_out = PIP_ST_FAULT + 2;
    wide = !PipVal<V>::narrow && wo.wide;
  }
  rflags_out = (W::any(wide) ? PIP_RES_WIDE : 0u) | ((!PipVal<V>::narrow && W::any(st.wrapped != 0)) ? PIP_RES_WRAPPED : 0u);
}

};

#endif
