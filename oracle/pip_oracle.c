/* TEST INFRASTRUCTURE ONLY -- see pip_oracle.h.
 *
 * A from-scratch CPU restatement of the PipLib 1.4.0 solver core in plain C: recursive, like
 * the reference, but re-entrant (no globals), with errors as status codes instead of exit(),
 * and with its own tableau model: every row *position* carries (flag, denominator, link) where
 * link is the owned column for a Unit position and a storage-row index otherwise.
 * Arithmetic is wrapping int64 (build with -fwrapv), exactly what the reference computes when
 * built the same way.  Each function cites the reference lines it follows (paths relative to
 * /root/reference).
 */
#include "pip_oracle.h"

#include <setjmp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef pio_int I;

enum { F_UNIT = 1, F_PLUS = 2, F_MINUS = 4, F_ZERO = 8, F_CRITIC = 16, F_UNKNOWN = 32 }; /* source/tab.h:58-63 */
enum { K_FREE, K_NIL, K_IF, K_LIST, K_FORM, K_NEW, K_DIV, K_VAL, K_ERROR };               /* source/sol.c:42-50 */
enum { T_INT = 1, T_DUAL = 2 };                                                            /* source/funcall.h:34-35 */
enum { S_SHIFT = 1, S_NEGATE = 2, S_REMOVE = 4, S_DUAL = 8 };                              /* source/sol.h:35-48 */
#define MAX_DET 4   /* source/tab.h:70 */

typedef struct { int rows, cols; const I *p; } Mat;      /* a PolyLib matrix view */
typedef struct {
  int height, width;     /* row positions, columns */
  int *flag, *link;      /* link: unit column (Unit position) or storage row */
  I *den;
  float *size;
  I *val;                /* storage rows, row-major, `width` wide */
  I det[MAX_DET];
  int ldet;
} Tab;

typedef struct {
  void **blk; int nblk, capblk;       /* LIFO allocation stack (stands in for source/tab.c:54-220) */
  int *cf; I *c1, *c2; int ncell;     /* solution cells (source/sol.c:37-102) */
  int sol_size, maxcol, maxparm;      /* source/type.h:39,49,50 */
  int deepest_cut;
  int in_dual;               /* serialiser: inside the dual list of a leaf */
  const Mat *dual_dom;       /* Compute_dual with equalities: the domain (flags the equality rows) */
  int depth;
  jmp_buf env;
  pio_stats *st;
} Ctx;

#define AT(t, i, j) ((t)->val[(size_t)(t)->link[i] * (t)->width + (j)])

static void fatal(Ctx *c, int code) { longjmp(c->env, code); }

/* exploration aid for workload design (never used in parity runs): give up after this many
 * pivots with status 3001; 0 = unlimited like the reference */
static long long pio_pivot_cap = 0;
void piporacle_set_pivot_cap(long long cap) { pio_pivot_cap = cap; }

/* ---- integer helpers: source/integrer.c:41-89, include/piplib/piplib.h:128-169 ---------- */
static I gcd_abs(I a, I b)
{
  while (b) { I r = a % b; a = b; b = r; }
  return a < 0 ? -a : a;
}
static I mod_pos(Ctx *c, I a, I b)
{
  I m;
  if (b == 0) fatal(c, PIO_FAULT);
  m = a % b;
  if (m < 0) m += (b < 0 ? -b : b);
  return m;
}
static I fdiv_q(Ctx *c, I a, I b) { return (a - mod_pos(c, a, b)) / b; }
static I xdiv(Ctx *c, I a, I b) { if (b == 0) fatal(c, PIO_FAULT); return a / b; }
static int bitlen(I x)
{
  int n = 0;
  unsigned long long u = x < 0 ? 0ULL - (unsigned long long)x : (unsigned long long)x;
  while (u) { u >>= 1; n++; }
  return n ? n : 1;
}
static void note_mul(Ctx *c, I a, I b)
{
  __int128 p = (__int128)a * b;
  if (p != (__int128)(I)p) c->st->wrapped++;
}

/* ---- allocation ------------------------------------------------------------------------ */
static void *grab(Ctx *c, size_t n)
{
  void *p = calloc(1, n ? n : 1);
  if (!p) fatal(c, PIO_FATAL + 23);
  if (c->nblk == c->capblk) {
    c->capblk = c->capblk ? 2 * c->capblk : 64;
    c->blk = realloc(c->blk, sizeof(void *) * c->capblk);
  }
  c->blk[c->nblk++] = p;
  return p;
}
static int mark(Ctx *c) { return c->nblk; }
static void release(Ctx *c, int m) { while (c->nblk > m) free(c->blk[--c->nblk]); }

/* tab_alloc_xx, source/tab.c:158-220: h stored rows after n virtual Unit rows */
static Tab *tab_new(Ctx *c, int h, int w, int n)
{
  Tab *t = grab(c, sizeof(Tab));
  int i;
  t->height = h + n; t->width = w;
  t->flag = grab(c, sizeof(int) * (h + n));
  t->link = grab(c, sizeof(int) * (h + n));
  t->den = grab(c, sizeof(I) * (h + n));
  t->size = grab(c, sizeof(float) * (h + n));
  t->val = grab(c, sizeof(I) * (size_t)h * w);
  t->det[0] = 1; t->ldet = 1;
  for (i = 0; i < n; i++) { t->flag[i] = F_UNIT; t->link[i] = i; t->den[i] = 1; }
  for (i = n; i < h + n; i++) { t->flag[i] = 0; t->link[i] = i - n; t->den[i] = 0; }
  return t;
}

/* expanser_xx, source/traiter.c:55-88 */
static Tab *expand(Ctx *c, Tab *tp, int virt, int reel, int ncol, int off, int dh, int dw)
{
  Tab *rp; int i, j, next = 0;
  if (!tp) return NULL;
  rp = tab_new(c, reel + dh, ncol + dw, virt);
  rp->ldet = tp->ldet;
  for (i = 0; i < tp->ldet; i++) rp->det[i] = tp->det[i];
  for (i = off; i < virt + reel; i++) {
    int ff = rp->flag[i] = tp->flag[i - off];
    rp->den[i] = tp->den[i - off];
    if (ff & F_UNIT) rp->link[i] = tp->link[i - off];
    else {
      rp->link[i] = next++;
      for (j = 0; j < ncol; j++) AT(rp, i, j) = AT(tp, i - off, j);
    }
  }
  return rp;
}

/* ---- solution cells: source/sol.c:89-226 ------------------------------------------------- */
static void cell(Ctx *c, int kind, I p1, I p2)
{
  int i = c->ncell;
  c->cf[i] = kind; c->c1[i] = p1; c->c2[i] = p2;
  c->ncell++;
  if (c->ncell >= c->sol_size) fatal(c, PIO_FATAL + 26);
}

/* chercher_xx, source/traiter.c:39-44 */
static int first_flag(Tab *t, int mask, int n)
{
  int i;
  for (i = 0; i < n; i++) if (t->flag[i] & mask) break;
  return i;
}

/* exam_coef_xx, source/traiter.c:101-159 */
static int exam_coef(Tab *t, int nvar, int ncol, int bigparm)
{
  int i, j, ff, fff;
  if (bigparm >= 0)
    for (i = 0; i < t->height; i++) {
      if (t->flag[i] != F_UNKNOWN) continue;
      if (AT(t, i, bigparm) < 0) { t->flag[i] = F_MINUS; return i; }
      else if (AT(t, i, bigparm) > 0) t->flag[i] = F_PLUS;
    }
  for (i = 0; i < t->height; i++) {
    ff = t->flag[i];
    if (ff == 0) break;
    if (ff != F_UNKNOWN) continue;
    ff = F_ZERO;
    for (j = nvar + 1; j < ncol; j++) {
      I v = AT(t, i, j);
      fff = v < 0 ? F_MINUS : v > 0 ? F_PLUS : F_ZERO;
      if (fff != F_ZERO && fff != ff) {
        if (ff == F_ZERO) ff = fff;
        else { ff = F_UNKNOWN; break; }
      }
    }
    { I v = AT(t, i, nvar); fff = v < 0 ? F_MINUS : v > 0 ? F_PLUS : F_ZERO; }
    if (ff == F_PLUS) { if (fff == F_MINUS) ff = F_UNKNOWN; }
    else if (ff == F_ZERO) ff = fff;
    else if (ff == F_MINUS) { if (fff != F_MINUS) ff = F_UNKNOWN; }
    t->flag[i] = ff;
    if (ff == F_MINUS) return i;
  }
  return i;
}

static void solve(Ctx *c, Tab *tp, Tab *ctxt, int nvar, int nparm, int ni, int nc, int bigparm, int flags);

/* compa_test_xx, source/traiter.c:162-243 */
static void compa_test(Ctx *c, Tab *tp, Tab *context, int ni, int nvar, int nparm, int nc)
{
  int i, j, m;
  if (nparm == 0) return;
  if (nparm >= c->maxparm) fatal(c, PIO_FATAL + 1);
  m = mark(c);
  for (i = 0; i < ni + nvar; i++) {
    int critic = 1, p, cplus, cminus;
    Tab *t;
    if (!(tp->flag[i] & (F_CRITIC | F_UNKNOWN))) continue;
    for (j = 0; j < nvar; j++) if (AT(tp, i, j) > 0) { critic = 0; break; }
    c->st->compa_rows++;
    t = expand(c, context, nparm, nc, nparm + 1, nparm, 1, 0);
    t->flag[nparm + nc] = F_UNKNOWN;
    for (j = 0; j < nparm; j++) AT(t, nparm + nc, j) = AT(tp, i, j + nvar + 1);
    AT(t, nparm + nc, nparm) = AT(tp, i, nvar);
    if (!critic) AT(t, nparm + nc, nparm) -= 1;
    t->den[nparm + nc] = 1;
    p = c->ncell;
    solve(c, t, NULL, nparm, 0, nc + 1, 0, -1, T_INT);
    cplus = c->cf[p] != K_NIL;
    c->ncell = p;
    t = expand(c, context, nparm, nc, nparm + 1, nparm, 1, 0);
    t->flag[nparm + nc] = F_UNKNOWN;
    for (j = 0; j < nparm; j++) AT(t, nparm + nc, j) = -AT(tp, i, j + nvar + 1);
    AT(t, nparm + nc, nparm) = -AT(tp, i, nvar) - 1;
    t->den[nparm + nc] = 1;
    solve(c, t, NULL, nparm, 0, nc + 1, 0, -1, T_INT);
    cminus = c->cf[p] != K_NIL;
    c->ncell = p;
    if (cplus && cminus) tp->flag[i] = critic ? F_CRITIC : F_UNKNOWN;
    else if (cminus) { tp->flag[i] = F_MINUS; break; }
    else tp->flag[i] = cplus ? F_PLUS : F_ZERO;
  }
  release(c, m);
}

/* valeur_xx, source/traiter.c:246-252 */
static I entry(Tab *t, int i, int j)
{
  if (t->flag[i] & F_UNIT) return t->link[i] == j ? t->den[i] : 0;
  return AT(t, i, j);
}

/* solution_xx, source/traiter.c:255-271 */
static void emit_solution(Ctx *c, Tab *t, int nvar, int nparm)
{
  int i, j, ncol = nvar + nparm + 1;
  cell(c, K_LIST, nvar, 0);
  for (i = 0; i < nvar; i++) {
    cell(c, K_FORM, nparm + 1, 0);
    for (j = nvar + 1; j < ncol; j++) cell(c, K_VAL, entry(t, i, j), t->den[i]);
    cell(c, K_VAL, entry(t, i, nvar), t->den[i]);
  }
}

/* solution_dual_xx, source/traiter.c:274-294 */
static void emit_dual(Ctx *c, Tab *t, int nvar, int *pos)
{
  int i;
  cell(c, K_LIST, t->height - nvar, 0);
  for (i = 0; i < t->height - nvar; i++) {
    cell(c, K_FORM, 1, 0);
    if (t->flag[pos[i]] & F_UNIT) cell(c, K_VAL, entry(t, 0, t->link[pos[i]]), t->den[0]);
    else cell(c, K_VAL, 0, 1);
  }
}

/* choisir_piv_xx, source/traiter.c:297-341 */
static int choose_column(Tab *t, int pivi, int nvar, int nligne)
{
  int j, k, pivj = -1;
  I pivot = 0, x = 0;
  for (j = 0; j < nvar; j++) {
    I foo = AT(t, pivi, j);
    if (foo <= 0) continue;
    if (pivj < 0) { pivj = j; pivot = foo; continue; }
    for (k = 0; k < nligne; k++) {
      x = pivot * entry(t, k, j) - entry(t, k, pivj) * foo;
      if (x) break;
    }
    if (x < 0) { pivj = j; pivot = foo; }
  }
  return pivj;
}

/* pivoter_xx, source/traiter.c:345-548 */
static int pivot_step(Ctx *c, Tab *t, int pivi, int nvar, int nparm, int ni)
{
  int ncol = nvar + nparm + 1, nligne = nvar + ni, i, j, k, pivj, ff, fff, slot;
  I pivot, dpiv, d, ppivot, dppiv;
  I *prow;
  long long rows = 0;

  pivj = choose_column(t, pivi, nvar, nligne);
  if (pivj < 0) return -1;
  pivot = AT(t, pivi, pivj);
  dpiv = t->den[pivi];
  d = gcd_abs(pivot, dpiv);
  ppivot = xdiv(c, pivot, d);
  dppiv = xdiv(c, dpiv, d);
  /* determinant bookkeeping = the overflow verdict, source/traiter.c:412-447 */
  for (i = 0; i < t->ldet; i++) {
    d = gcd_abs(t->det[i], dppiv);
    t->det[i] = xdiv(c, t->det[i], d);
    dppiv = xdiv(c, dppiv, d);
  }
  if (dppiv != 1) fatal(c, PIO_FATAL + 1);
  for (i = 0; i < t->ldet; i++)
    if (bitlen(t->det[i]) + bitlen(ppivot) < 64) { t->det[i] *= ppivot; break; }
  if (i >= t->ldet) {
    t->ldet++;
    if (t->ldet >= MAX_DET) fatal(c, PIO_FATAL + 1);
    t->det[i] = ppivot;
  }
  c->st->pivots++;
  if (pio_pivot_cap && c->st->pivots > pio_pivot_cap) fatal(c, 3001);

  prow = &AT(t, pivi, 0);
  for (k = 0; k < nligne; k++) {
    I foo, lpiv, g, z, *p;
    if (t->flag[k] & F_UNIT) continue;
    if (k == pivi) continue;
    rows++;
    p = &AT(t, k, 0);
    foo = p[pivj];
    d = gcd_abs(pivot, foo);
    lpiv = xdiv(c, pivot, d);
    foo = xdiv(c, foo, d);
    note_mul(c, lpiv, t->den[k]);
    g = lpiv * t->den[k];
    t->den[k] = g;
    for (j = 0; j < ncol; j++) {
      if (j == pivj) { note_mul(c, dpiv, foo); z = dpiv * foo; }
      else { note_mul(c, p[j], lpiv); note_mul(c, prow[j], foo); z = p[j] * lpiv - prow[j] * foo; }
      p[j] = z;
      if (g != 1) g = gcd_abs(g, z);
    }
    if (g != 1) {
      for (j = 0; j < ncol; j++) p[j] = xdiv(c, p[j], g);
      t->den[k] = xdiv(c, t->den[k], g);
    }
  }
  c->st->elem_updates += rows * ncol;
  /* the Unit position that owned column pivj takes over the pivot row's storage, 503-516 */
  for (k = 0; k < nligne; k++)
    if ((t->flag[k] & F_UNIT) && t->link[k] == pivj) break;
  slot = t->link[pivi];
  for (j = 0; j < ncol; j++) prow[j] = (j == pivj) ? dpiv : -prow[j];
  t->flag[k] = F_PLUS; t->link[k] = slot; t->den[k] = pivot;
  t->flag[pivi] = F_UNIT | F_ZERO; t->den[pivi] = 1; t->link[pivi] = pivj;
  /* re-flag from the sign of the pivot-column entry, 518-529 */
  for (k = 0; k < nligne; k++) {
    I v;
    ff = t->flag[k];
    if (ff & F_UNIT) continue;
    v = AT(t, k, pivj);
    fff = v < 0 ? F_MINUS : v == 0 ? F_ZERO : F_PLUS;
    if (fff != F_ZERO && fff != ff) {
      if (ff == F_ZERO) ff = (fff == F_MINUS ? F_UNKNOWN : fff);
      else ff = F_UNKNOWN;
    }
    t->flag[k] = ff;
  }
  if (nligne > c->st->max_rows) c->st->max_rows = nligne;
  if (ncol > c->st->max_cols) c->st->max_cols = ncol;
  return 0;
}

/* (int)t of an out-of-range or NaN double is INT_MIN on x86-64 (cvttsd2si); abs(INT_MIN) stays
 * INT_MIN; so such a value never raises the running maximum.  source/traiter.c:582-583 */
static double size_term(double t)
{
  if (!(t > -2147483649.0 && t < 2147483648.0)) return -2147483648.0;
  { int v = (int)t; if (v == (-2147483647 - 1)) return -2147483648.0; return (double)(v < 0 ? -v : v); }
}

/* tab_sort_rows_xx, source/traiter.c:556-623 */
static int *sort_rows(Ctx *c, Tab *t, int nvar, int nligne, int flags)
{
  int i, j, pivi, *pos = NULL, *ineq = NULL;
  double s, d, smax = 0;
  if (flags & T_DUAL) {
    ineq = grab(c, sizeof(int) * t->height);
    pos = grab(c, sizeof(int) * (t->height - nvar + 1));
  }
  for (i = nvar; i < nligne; i++) {
    if (t->flag[i] & F_UNIT) continue;
    s = 0;
    d = (double)t->den[i];
    for (j = 0; j < nvar; j++) {
      double v = size_term((double)AT(t, i, j) / d);
      if (v > s) s = v;
    }
    t->size[i] = (float)s;
    if (s > smax) smax = s;
    if (flags & T_DUAL) ineq[i] = i - nvar;
  }
  for (i = nvar; i < nligne; i++) {
    if (t->flag[i] & F_UNIT) continue;
    s = smax; pivi = i;
    for (j = i; j < nligne; j++) {
      if (t->flag[j] & F_UNIT) continue;
      if (t->size[j] < s) { s = t->size[j]; pivi = j; }
    }
    if (pivi != i) {
      int fi = t->flag[i], li = t->link[i]; I di = t->den[i]; float si = t->size[i];
      t->flag[i] = t->flag[pivi]; t->link[i] = t->link[pivi]; t->den[i] = t->den[pivi]; t->size[i] = t->size[pivi];
      t->flag[pivi] = fi; t->link[pivi] = li; t->den[pivi] = di; t->size[pivi] = si;
      if (flags & T_DUAL) { j = ineq[i]; ineq[i] = ineq[pivi]; ineq[pivi] = j; }
    }
  }
  if (flags & T_DUAL) for (i = nvar; i < nligne; i++) pos[ineq[i]] = i;
  return pos;
}

/* bezout_xx, source/integrer.c:98-150: z with z*y = x (mod delta) */
static I bezout(Ctx *cx, I x, I y, I delta)
{
  I a = 1, b = 0, c = 0, d = 1, u = y, v = delta;
  for (;;) {
    I q = fdiv_q(cx, u, v), r = mod_pos(cx, u, v), e, f;
    if (r == 0) break;
    u = v; v = r;
    e = a - q * c; f = b - q * d;
    a = c; b = d; c = e; d = f;
  }
  if (v != 1) return 0;
  return mod_pos(cx, c * x, delta);
}

/* has_cut_xx, source/integrer.c:230-254 */
static int has_cut(Tab *ctx, int nr, int nparm, int p, I *cut)
{
  int row, col;
  for (row = 0; row < nr; row++) {
    if (AT(ctx, row, p) != cut[1 + nparm]) continue;
    if (AT(ctx, row, nparm) != cut[0]) continue;
    for (col = p + 1; col < nparm; col++) if (AT(ctx, row, col) != 0) break;
    if (col < nparm) continue;
    for (col = 0; col < p; col++) if (AT(ctx, row, col) != cut[1 + col]) break;
    if (col < p) continue;
    return 1;
  }
  return 0;
}

/* find_parm_xx, source/integrer.c:258-291 (cut = constant, parameters, denominator) */
static int find_parm(Tab *ctx, int nr, int nparm, I *cut)
{
  int p, col, found;
  if (cut[1 + nparm - 1] != 0) return -1;
  cut[0] = cut[0] + cut[1 + nparm] - 1;
  for (p = nparm - 1; p >= 0; p--) {
    if (cut[1 + p] != 0) break;
    if (!has_cut(ctx, nr, nparm, p, cut)) continue;
    cut[0] = cut[0] + 1 - cut[1 + nparm];
    for (col = 0; col < 1 + nparm + 1; col++) cut[col] = -cut[col];
    found = has_cut(ctx, nr, nparm, p, cut);
    for (col = 0; col < 1 + nparm + 1; col++) cut[col] = -cut[col];
    if (found) return p;
    cut[0] = cut[0] + cut[1 + nparm] - 1;
  }
  cut[0] = cut[0] + 1 - cut[1 + nparm];
  return -1;
}

/* add_parm_xx, source/integrer.c:156-227 */
static void add_parm(Ctx *c, Tab **pctx, int nr, int *pnparm, int *pni, int *pnc, I *cut)
{
  int nparm = *pnparm, j, k;
  Tab *x;
  cell(c, K_NEW, nparm, 0);
  cell(c, K_DIV, 0, 0);
  cell(c, K_FORM, nparm + 1, 0);
  for (j = 0; j < nparm; j++) cell(c, K_VAL, -cut[1 + j], 1);
  cell(c, K_VAL, -cut[0], 1);
  cell(c, K_VAL, cut[1 + nparm], 1);
  if (nr + 2 > (*pctx)->height || nparm + 1 + 1 > (*pctx)->width) {
    int dcw = bitlen(cut[1 + nparm]);
    *pctx = expand(c, *pctx, 0, nr, nparm + 1, 0, 2 * dcw + *pni, dcw);
  }
  x = *pctx;
  for (k = 0; k < nr; k++) { AT(x, k, nparm + 1) = AT(x, k, nparm); AT(x, k, nparm) = 0; }
  for (j = 0; j < nparm; j++) { AT(x, nr, j) = -cut[1 + j]; AT(x, nr + 1, j) = cut[1 + j]; }
  AT(x, nr, nparm) = -cut[1 + nparm];
  AT(x, nr + 1, nparm) = cut[1 + nparm];
  AT(x, nr, nparm + 1) = -cut[0];
  AT(x, nr + 1, nparm + 1) = cut[0] - 1 + cut[1 + nparm];
  x->flag[nr] = x->flag[nr + 1] = F_UNKNOWN;
  x->den[nr] = x->den[nr + 1] = 1;
  (*pnparm)++;
  (*pnc) += 2;
}

/* integrer_xx, source/integrer.c:305-534: returns the new cut row, 0 (integral) or -1 */
static int make_cut(Ctx *c, Tab **ptp, Tab **pctx, int *pnvar, int *pnparm, int *pni, int *pnc, int bigparm)
{
  int nvar = *pnvar, nparm = *pnparm, ni = *pni, nc = *pnc;
  int ncol = nvar + nparm + 1, nligne = nvar + ni, i, j, parm, m;
  I *cut;
  if (ncol >= c->maxcol) fatal(c, PIO_FATAL + 3);
  m = mark(c); (void)m;
  cut = grab(c, sizeof(I) * (ncol + 2));
  for (i = 0; i < nvar; i++) {
    Tab *t = *ptp;
    I D = t->den[i], x;
    int ok_var = 0, ok_const, ok_parm = 0;
    if (D == 1) continue;
    if (t->flag[i] & F_UNIT) continue;
    for (j = 0; j < nvar; j++) {
      x = mod_pos(c, AT(t, i, j), D);
      cut[j] = x;
      if (x > 0) ok_var = 1;
    }
    x = -mod_pos(c, -AT(t, i, nvar), D);
    cut[nvar] = x;
    ok_const = (x != 0);
    for (j = nvar + 1; j < ncol; j++) {
      if (j == bigparm) { cut[j] = 0; continue; }
      cut[j] = -mod_pos(c, -AT(t, i, j), D);
      if (cut[j] != 0) ok_parm = 1;
    }
    cut[ncol] = D;
    if (!ok_parm && !ok_const) continue;                       /* case (a) */
    if (!ok_parm) {
      if (!ok_var) return -1;                                    /* case (b) */
      if (nligne >= t->height) {                                 /* case (d) */
        int dth = bitlen(D);
        *ptp = t = expand(c, t, nvar, ni, ncol, 0, dth, 0);
      }
      if (c->deepest_cut) {                                      /* source/integrer.c:417-438 */
        I tt = -cut[nvar], delta = gcd_abs(tt, D), tau = xdiv(c, tt, delta), dd = xdiv(c, D, delta), lambda;
        tt = dd - 1;
        lambda = bezout(c, tt, tau, dd);
        tt = gcd_abs(lambda, D);
        while (tt != 1) { lambda += dd; tt = gcd_abs(lambda, D); }
        for (j = 0; j < nvar; j++) cut[j] = mod_pos(c, lambda * cut[j], D);
        tt = mod_pos(c, cut[nvar] * lambda, D);
        cut[nvar] = -(D - tt);
      }
      t->flag[nligne] = F_MINUS;
      t->den[nligne] = D;
      for (j = 0; j < ncol; j++) AT(t, nligne, j) = cut[j];
      (*pni)++;
      c->st->cuts_const++;
      return nligne;
    }
    /* case (e): parametric cut, source/integrer.c:493-520 */
    parm = find_parm(*pctx, nc, nparm, cut + nvar);
    if (parm == -1) {
      add_parm(c, pctx, nc, pnparm, pni, pnc, cut + nvar);
      parm = nparm;
    }
    if (!ok_var) fatal(c, PIO_FATAL + 134);                      /* assert(ok_var) -> abort */
    if (nligne >= t->height || ncol >= t->width) {
      int d = bitlen(D);
      *ptp = t = expand(c, t, nvar, ni, ncol, 0, d + ni, d);
    }
    t->flag[nligne] = F_MINUS;
    t->den[nligne] = D;
    for (j = 0; j < ncol; j++) AT(t, nligne, j) = cut[j];
    AT(t, nligne, nvar + 1 + parm) += cut[ncol];
    (*pni)++;
    c->st->cuts_parm++;
    return nligne;
  }
  return 0;
}

static int det_bits(Tab *t)
{
  int i, n = 0;
  for (i = 0; i < t->ldet; i++) n += bitlen(t->det[i]);
  return n;
}

/* traiter_xx, source/traiter.c:628-791 */
static void solve(Ctx *c, Tab *tp, Tab *ctxt, int nvar, int nparm, int ni, int nc, int bigparm, int flags)
{
  int j, pivi, nligne, ncol, dcw, dch, *pos, x;
  Tab *context;

  c->st->traiter_calls++;
  c->depth++;
  if (c->depth > c->st->max_depth) c->st->max_depth = c->depth;
  dcw = det_bits(tp);
  dch = 2 * dcw + 1;
  x = mark(c);
  nligne = nvar + ni;
  context = expand(c, ctxt, 0, nc, nparm + 1, 0, dch, dcw);
  pos = sort_rows(c, tp, nvar, nligne, flags);

  for (;;) {
    nligne = nvar + ni; ncol = nvar + nparm + 1;
    if (nc > c->st->max_ctx_rows) c->st->max_ctx_rows = nc;
    pivi = first_flag(tp, F_MINUS, nligne);
    if (pivi < nligne) goto pirouette;
    pivi = exam_coef(tp, nvar, ncol, bigparm);
    if (pivi < nligne) goto pirouette;
    compa_test(c, tp, context, ni, nvar, nparm, nc);
    pivi = first_flag(tp, F_MINUS, nligne);
    if (pivi < nligne) goto pirouette;
    pivi = first_flag(tp, F_CRITIC, nligne);
    if (pivi >= nligne) pivi = first_flag(tp, F_UNKNOWN, nligne);
    if (pivi < nligne) {                                          /* split, source/traiter.c:695-759 */
      Tab *ntp; I g = 0; int q;
      if (nc >= context->height) {
        dcw = det_bits(tp);
        context = expand(c, context, 0, nc, nparm + 1, 0, 2 * dcw + 1, dcw);
      }
      if (nparm >= c->maxparm) fatal(c, PIO_FATAL + 2);
      q = mark(c);
      ntp = expand(c, tp, nvar, ni, ncol, 0, 0, 0);
      c->st->splits++;
      cell(c, K_IF, 0, 0);
      cell(c, K_FORM, nparm + 1, 0);
      for (j = 0; j < nparm; j++) g = gcd_abs(g, AT(tp, pivi, j + nvar + 1));
      if (!(flags & T_INT)) g = gcd_abs(g, AT(tp, pivi, nvar));
      for (j = 0; j < nparm; j++) {
        AT(context, nc, j) = xdiv(c, AT(tp, pivi, j + nvar + 1), g);
        cell(c, K_VAL, AT(context, nc, j), 1);
      }
      if (!(flags & T_INT)) AT(context, nc, nparm) = xdiv(c, AT(tp, pivi, nvar), g);
      else AT(context, nc, nparm) = fdiv_q(c, AT(tp, pivi, nvar), g);
      cell(c, K_VAL, AT(context, nc, nparm), 1);
      context->flag[nc] = F_UNKNOWN;
      context->den[nc] = 1;
      ntp->flag[pivi] = F_PLUS;
      solve(c, ntp, context, nvar, nparm, ni, nc + 1, bigparm, flags);
      release(c, q);
      for (j = 0; j < nparm; j++) AT(context, nc, j) = -AT(context, nc, j);
      AT(context, nc, nparm) = -(AT(context, nc, nparm) + 1);
      tp->flag[pivi] = F_MINUS;
      context->den[nc] = 1;
      nc++;
      goto pirouette;
    }
    if (!(flags & T_INT)) {
      emit_solution(c, tp, nvar, nparm);
      if (flags & T_DUAL) emit_dual(c, tp, nvar, pos);
      break;
    }
    pivi = make_cut(c, &tp, &context, &nvar, &nparm, &ni, &nc, bigparm);
    if (pivi > 0) goto pirouette;
    if (pivi == 0) emit_solution(c, tp, nvar, nparm);
    else cell(c, K_NIL, 0, 0);
    break;
pirouette:
    if (pivot_step(c, tp, pivi, nvar, nparm, ni) < 0) { cell(c, K_NIL, 0, 0); break; }
  }
  release(c, x);
  c->depth--;
}

/* ---- boundary: source/tab.c:292-427 ------------------------------------------------------- */
#define M(m, i, j) ((m)->p[(size_t)(i) * (m)->cols + (j)])

/* tab_Matrix2Tableau_xx, source/tab.c:292-393 */
static Tab *matrix_to_tab(Ctx *c, const Mat *mx, int nineq, int nv, int n, int shift, int bg, int urs)
{
  Tab *p; int i, j, k, cur, decal = 0, isnew, ctx, cst, ncolm;
  I big = 0;
  ctx = (n == -1);
  if (ctx) n = 0;
  ncolm = mx->cols - 1;
  isnew = shift && (bg + ctx > 0) && ((unsigned)(bg + ctx) > (unsigned)(mx->cols - 2));
  if (isnew) ncolm++;
  if (ctx) { shift = 0; cst = nv + urs; } else cst = nv;
  p = tab_new(c, nineq, ncolm + urs, n);
  for (i = 0; i < mx->rows; i++) {
    int ineq;
    cur = i + n + decal;
    p->flag[cur] = F_UNKNOWN;
    p->den[cur] = 1;
    if (shift) big = 0;
    ineq = (M(mx, i, 0) != 0);
    for (j = 0; j < nv; j++) {
      if (isnew && j == bg) continue;
      if (shift) big += M(mx, i, 1 + j);
      AT(p, cur, j) = shift > 0 ? -M(mx, i, 1 + j) : M(mx, i, 1 + j);
    }
    for (k = j = nv + 1; j < ncolm; j++) {
      if (isnew && j == bg) continue;
      AT(p, cur, j) = M(mx, i, k);
      k++;
    }
    for (j = 0; j < urs; j++) {
      int pos_n = ncolm - ctx + j, pos = pos_n - urs;
      if (pos <= bg) --pos;
      AT(p, cur, pos_n) = -AT(p, cur, pos);
    }
    AT(p, cur, cst) = M(mx, i, mx->cols - 1);
    if (shift) {
      if (shift < 0) big = -big;
      if (isnew) AT(p, cur, bg) = big; else AT(p, cur, bg) += big;
    }
    if (!ineq) {
      decal++;
      p->flag[cur + 1] = F_UNKNOWN;
      p->den[cur + 1] = 1;
      for (j = 0; j < ncolm + urs; j++) AT(p, cur + 1, j) = -AT(p, cur, j);
    }
  }
  return p;
}

/* tab_simplify_xx, source/tab.c:396-427 */
static void simplify(Ctx *c, Tab *t, int cst)
{
  int i, j;
  for (i = 0; i < t->height; i++) {
    I g = 0;
    if (t->flag[i] & F_UNIT) continue;
    for (j = 0; j < t->width; j++) {
      if (j == cst) continue;
      g = gcd_abs(g, AT(t, i, j));
      if (g == 1) break;
    }
    if (g == 0 || g == 1) continue;
    for (j = 0; j < t->width; j++)
      AT(t, i, j) = (j == cst) ? fdiv_q(c, AT(t, i, j), g) : xdiv(c, AT(t, i, j), g);
  }
}

/* skip_xx / sol_simplify_xx, source/sol.c:236-288 */
static int skip(Ctx *c, int i);
static int skip_new(Ctx *c, int i) { if (c->cf[i] != K_NEW) return i; return skip(c, i + 1); }
static int skip(Ctx *c, int i)
{
  int n;
  while (c->cf[i] == K_FREE || c->cf[i] == K_ERROR) i++;
  switch (c->cf[i]) {
  case K_NIL: case K_VAL: i++; break;
  case K_NEW: i = skip_new(c, i); break;
  case K_IF: i = skip(c, i + 1); i = skip(c, i); i = skip(c, i); break;
  case K_LIST: case K_FORM: n = (int)c->c1[i]; i++; while (n--) i = skip(c, i); break;
  case K_DIV: i = skip(c, i + 1); i = skip(c, i); break;
  }
  return skip_new(c, i);
}
static void sol_simplify(Ctx *c, int i)
{
  int j, k, l;
  if (c->cf[i] != K_IF) return;
  j = skip(c, i + 1);
  k = skip(c, j);
  sol_simplify(c, k);
  sol_simplify(c, j);
  if (c->cf[j] == K_NIL && c->cf[k] == K_NIL) {
    c->cf[i] = K_NIL;
    if (k >= c->ncell - 1) c->ncell = i + 1;
    else for (l = i + 1; l <= k; l++) c->cf[l] = K_FREE;
  }
}

/* ---- cells -> serialised quast (the format of oracle/ref_harness.c), following the decoder
 * sol_quast_edit_xx & co, source/sol.c:435-734 --------------------------------------------- */
typedef struct { I *out; long cap, len; } Ser;
static void sput(Ser *s, I v) { if (s->len < s->cap) s->out[s->len] = v; s->len++; }

static void ser_vector(Ctx *c, Ser *s, int *i, int bg, int urs, int flags)
{
  int j, k, n = (int)c->c1[*i], unbounded = 0, first_urs;
  long at;
  if (flags & S_REMOVE) --n;
  n -= urs;
  first_urs = urs + (bg >= 0);
  sput(s, n);
  at = s->len;
  for (j = 0, k = 0; k < n; j++) {
    I N, D, d;
    (*i)++;
    N = c->c1[*i]; D = c->c2[*i];
    d = gcd_abs(N, D);
    if ((flags & S_SHIFT) && j == bg) { N -= D; if (N != 0) unbounded = 1; }
    if ((flags & S_REMOVE) && j == bg) continue;
    if (first_urs <= j && j < first_urs + urs) continue;
    N = d ? N / d : 0;
    if (flags & S_NEGATE) N = -N;
    sput(s, N);
    sput(s, d == D ? 1 : (d ? D / d : 0));
    k++;
  }
  if (unbounded)
    for (k = 0; k < n; k++) if (at + 2 * k + 1 < s->cap) s->out[at + 2 * k + 1] = 0;
  (*i)++;
}

static void ser_node(Ctx *c, Ser *s, int *i, int bg, int urs, int flags)
{
  int n = 0, k, kind;
  while (c->cf[*i] == K_FREE) (*i)++;
  /* count the chain of newparm definitions first (sol_newparm_edit_xx, source/sol.c:525-577) */
  if (c->cf[*i] == K_NEW) {
    int t = *i;
    while (c->cf[t] == K_NEW) {
      int m = (int)c->c1[t + 2];       /* Form length */
      n++;
      t = t + 2 + 1 + m + 1;           /* New Div Form Val*m Val(deno) */
    }
  }
  sput(s, n);
  for (k = 0; k < n; k++) {
    int newcell = *i, rank;
    (*i) += 2;
    rank = (int)c->c1[newcell];
    if (flags & S_REMOVE) rank--;
    rank -= urs;
    sput(s, rank);
    sput(s, c->c1[*i + (int)c->c1[*i] + 1]);     /* the divisor cell follows the form */
    ser_vector(c, s, i, bg, urs, flags & S_REMOVE);
    (*i)++;
  }
  kind = c->cf[*i];
  (*i)++;
  if (kind == K_LIST) {
    int ne = (int)c->c1[*i - 1];
    sput(s, 1);
    if (ne == 0) { sput(s, 1); sput(s, 0); }     /* one list element with a NULL vector */
    else if (c->in_dual && c->dual_dom) {
      /* pip_quast_equalities_dual_xx, source/piplib.c:651-690: an equality row was solved as two
       * inequalities; keep the first dual value when it is non-zero, else minus the second */
      const Mat *dm = c->dual_dom;
      int r, kept = ne, e = 0;
      for (r = 0; r < dm->rows; r++) if (M(dm, r, 0) == 0) kept--;
      sput(s, kept);
      for (r = 0; r < dm->rows && e < ne; r++) {
        Ser drop = {NULL, 0, 0};
        if (M(dm, r, 0) != 0) { sput(s, 1); ser_vector(c, s, i, bg, urs, flags); e++; continue; }
        if (c->c1[*i + 1] != 0) {                  /* Form 1, Val: the value cell follows the form */
          sput(s, 1); ser_vector(c, s, i, bg, urs, flags);
          ser_vector(c, &drop, i, bg, urs, flags);
        } else {
          long at;
          ser_vector(c, &drop, i, bg, urs, flags);
          sput(s, 1);
          at = s->len;
          ser_vector(c, s, i, bg, urs, flags);
          if (at + 1 < s->cap) s->out[at + 1] = -s->out[at + 1];
        }
        e += 2;
      }
    } else {
      sput(s, ne);
      for (k = 0; k < ne; k++) { sput(s, 1); ser_vector(c, s, i, bg, urs, flags); }
    }
    if (flags & S_DUAL) { sput(s, 1); c->in_dual = 1; ser_node(c, s, i, bg, urs, 0); c->in_dual = 0; }
    else sput(s, 0);
  } else if (kind == K_NIL) {
    sput(s, 0);
  } else if (kind == K_IF) {
    sput(s, 2);
    ser_vector(c, s, i, bg, urs, flags & S_REMOVE);
    ser_node(c, s, i, bg, urs, flags);
    ser_node(c, s, i, bg, urs, flags);
  } else fatal(c, PIO_FATAL + 1);
}

static void ctx_init(Ctx *c, pio_stats *st, int sol_size, int maxcol)
{
  memset(c, 0, sizeof(*c));
  c->sol_size = sol_size > 0 ? sol_size : 4096;
  c->maxcol = maxcol > 0 ? maxcol : 512;
  c->maxparm = 50;
  c->cf = malloc(sizeof(int) * (c->sol_size + 1));
  c->c1 = malloc(sizeof(I) * (c->sol_size + 1));
  c->c2 = malloc(sizeof(I) * (c->sol_size + 1));
  c->st = st;
}
static void ctx_done(Ctx *c)
{
  release(c, 0);
  free(c->blk); free(c->cf); free(c->c1); free(c->c2);
}

/* the CLI path, source/maind.c:150-232, on an already-lexed problem */
int piporacle_traiter(int nvar, int nparm, int ni, int nc, int bigparm, int nq,
                      const I *tab, const I *ctx,
                      int *cell_flags, I *cell_p1, I *cell_p2, int cap, int *ncells,
                      pio_stats *st, int sol_size, int maxcol)
{
  Ctx c; pio_stats local; Tab *ineq, *context; int i, j, rc, ncol = nvar + nparm + 1, nonvoid = 1;
  if (!st) { st = &local; }
  memset(st, 0, sizeof(*st));
  ctx_init(&c, st, sol_size, maxcol);
  *ncells = 0;
  /* nq: bit 0 = integer solution wanted (the .dat field); test-only extensions: bit 1 = TRAITER_DUAL
   * (rational problems), bit 2 = deepest cuts */
  c.deepest_cut = (nq >> 2) & 1;
  nq &= 3;
  rc = setjmp(c.env);
  if (rc) { ctx_done(&c); return rc; }
  ineq = tab_new(&c, ni, ncol, nvar);
  for (i = 0; i < ni; i++) {
    ineq->flag[nvar + i] = F_UNKNOWN; ineq->den[nvar + i] = 1;
    for (j = 0; j < ncol; j++) AT(ineq, nvar + i, j) = tab[(size_t)i * ncol + j];
  }
  if (nq & 1) simplify(&c, ineq, nvar);
  context = tab_new(&c, nc, nparm + 1, 0);
  for (i = 0; i < nc; i++) {
    context->flag[i] = F_UNKNOWN; context->den[i] = 1;
    for (j = 0; j < nparm + 1; j++) AT(context, i, j) = ctx[(size_t)i * (nparm + 1) + j];
  }
  if (nq & 1) simplify(&c, context, nparm);
  if (nc) {
    Tab *t = expand(&c, context, nparm, nc, nparm + 1, nparm, 0, 0);
    solve(&c, t, NULL, nparm, 0, nc, 0, -1, T_INT);
    nonvoid = c.cf[0] != K_NIL;
    c.ncell = 0;
  }
  if (nonvoid) {
    solve(&c, ineq, context, nvar, nparm, ni, nc, bigparm, (nq & 1) ? T_INT : (nq & 2) ? T_DUAL : 0);
    for (i = 0; i < c.ncell && i < cap; i++) { cell_flags[i] = c.cf[i]; cell_p1[i] = c.c1[i]; cell_p2[i] = c.c2[i]; }
    *ncells = c.ncell;
  }
  ctx_done(&c);
  return nonvoid ? PIO_OK : PIO_VOID;
}

/* pip_solve_xx, source/piplib.c:722-880.  Returns status; *voidp = 1 when the answer is NULL */
static int lib_solve(Ctx *c, const Mat *dom, const Mat *par, int bg, const int *opts, Ser *s)
{
  int np, nn, nl, nm = 0, i, shift = 0, urs = 0, sol_flags = 0, nonvoid = 1, flags = 0, xq = 0;
  Tab *context, *ineq;
  int nq = opts[0], simp = opts[2], maxi = opts[4], urs_p = opts[5], urs_u = opts[6], dual = opts[7];
  c->deepest_cut = opts[3];
  np = par ? par->cols - 2 : 0;
  nn = dom->cols - np - 2;
  nl = dom->rows;
  for (i = 0; i < dom->rows; i++) if (M(dom, i, 0) == 0) nl++;
  if (maxi) { sol_flags |= S_SHIFT | S_NEGATE; shift = 1; }
  else if (urs_u) { sol_flags |= S_SHIFT; shift = -1; }
  if (urs_p) { urs = np - (bg >= 0); np += urs; }
  if (maxi || urs_u) if (bg < 0) { bg = dom->cols - 1; np++; sol_flags |= S_REMOVE; }
  if (par) {
    nm = par->rows;
    for (i = 0; i < par->rows; i++) if (M(par, i, 0) == 0) nm++;
    context = matrix_to_tab(c, par, nm, np - urs, -1, shift, bg - nn - 1, urs);
    if (nq) simplify(c, context, np);
    if (nm) {
      Tab *t = expand(c, context, np, nm, np + 1, np, 0, 0);
      solve(c, t, NULL, np, 0, nm, 0, -1, T_INT);
      nonvoid = c->cf[0] != K_NIL;
      c->ncell = 0;
    }
  } else {
    Mat empty = {0, 2, NULL};
    context = matrix_to_tab(c, &empty, 0, np - urs, -1, shift, bg - nn - 1, urs);
  }
  if (!nonvoid) { sput(s, -1); return PIO_OK; }
  ineq = matrix_to_tab(c, dom, nl, nn, nn, shift, bg, urs);
  if (nq) simplify(c, ineq, nn);
  if (nq) flags |= T_INT;
  else if (dual) { flags |= T_DUAL; sol_flags |= S_DUAL; }
  solve(c, ineq, context, nn, np, nl, nm, bg, flags);
  if (simp) sol_simplify(c, 0);
  c->dual_dom = ((sol_flags & S_DUAL) && nl > dom->rows) ? dom : NULL;     /* source/piplib.c:867-868 */
  ser_node(c, s, &xq, bg - nn - 1, urs, sol_flags);

  return PIO_OK;
}

int piporacle_solve_ser(int dom_rows, int dom_cols, const I *dom,
                        int has_ctx, int ctx_rows, int ctx_cols, const I *ctx,
                        int bg, const int *opts, I *ser, long cap, long *ser_n, pio_stats *st)
{
  Ctx c; pio_stats local; Mat D = {dom_rows, dom_cols, dom}, P = {ctx_rows, ctx_cols, ctx}; Ser s = {ser, cap, 0};
  int rc;
  if (!st) st = &local;
  memset(st, 0, sizeof(*st));
  ctx_init(&c, st, 0, 0);
  rc = setjmp(c.env);
  if (rc) { ctx_done(&c); if (ser_n) *ser_n = 0; return rc; }
  rc = lib_solve(&c, &D, has_ctx ? &P : NULL, bg, opts, &s);
  if (ser_n) *ser_n = s.len;
  ctx_done(&c);
  return rc;
}

/* hash of a serialised quast: start value + sum over the words of the splitmix64 finaliser of
 * (word + (index + 1) * golden ratio): the function the library applies (pip_decode.h, pip_hash_word) */
static unsigned long long hash_word(unsigned long long v, unsigned long long k)
{
  unsigned long long x = v + (k + 1ULL) * 0x9E3779B97F4A7C15ULL;
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
  x ^= x >> 27; x *= 0x94D049BB133111EBULL;
  x ^= x >> 31;
  return x;
}

/* timed loop over a dense batch; hashes are hash_word sums over the serialised quast words, the same
 * function oracle/ref_harness.c applies to the reference's PipQuast */
double piporacle_bench_dense(long first, long count, int dom_rows, int dom_cols, const I *dom,
                             int has_ctx, int ctx_rows, int ctx_cols, const I *ctx, int bg,
                             const int *opts, int *status, unsigned long long *hashes, pio_stats *tot)
{
  long i, k; double total = 0; struct timespec t0, t1;
  long cap = 1 << 16; I *buf = malloc(sizeof(I) * cap);
  pio_stats st;
  if (tot) memset(tot, 0, sizeof(*tot));
  for (i = first; i < first + count; i++) {
    long n = 0; int rc;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    rc = piporacle_solve_ser(dom_rows, dom_cols, dom + (size_t)i * dom_rows * dom_cols, has_ctx, ctx_rows, ctx_cols,
                             has_ctx ? ctx + (size_t)i * ctx_rows * ctx_cols : NULL, bg, opts, buf, cap, &n, &st);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    total += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    if (status) status[i - first] = rc;
    if (hashes) {
      unsigned long long h = 0xcbf29ce484222325ULL;
      if (rc == PIO_OK) { for (k = 0; k < n && k < cap; k++) h += hash_word((unsigned long long)buf[k], (unsigned long long)k); } else h = 0;
      hashes[i - first] = h;
    }
    if (tot) {
      tot->pivots += st.pivots; tot->cuts_const += st.cuts_const; tot->cuts_parm += st.cuts_parm;
      tot->traiter_calls += st.traiter_calls; tot->compa_rows += st.compa_rows; tot->splits += st.splits;
      tot->elem_updates += st.elem_updates; tot->wrapped += st.wrapped;
      if (st.max_rows > tot->max_rows) tot->max_rows = st.max_rows;
      if (st.max_cols > tot->max_cols) tot->max_cols = st.max_cols;
      if (st.max_depth > tot->max_depth) tot->max_depth = st.max_depth;
      if (st.max_ctx_rows > tot->max_ctx_rows) tot->max_ctx_rows = st.max_ctx_rows;
    }
  }
  free(buf);
  return total;
}
