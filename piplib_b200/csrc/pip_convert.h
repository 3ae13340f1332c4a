/* PolyLib matrix rows -> tableau rows: tab_Matrix2Tableau_xx (source/tab.c:292-393) restated as a
 * per-row function that the host (pip_host.cpp: pageable caller buffers, narrowing while converting)
 * and the device (pip_convert_kernel: caller buffers in pinned memory, raw rows uploaded by DMA) both
 * compile -- one source of truth for the column surgery (Shift / big parameter / Urs_parms copies).
 *
 * A problem of a dense batch is `dr` domain rows of `dc` words [eq(0)/ineq(1) | unknowns | parameters |
 * constant] and `cr` context rows of `cc` words.  Everything that does not depend on the data is derived
 * once per batch on the host (PipConvertShape, source/piplib.c:758-797); only the number of equality
 * rows (each becomes two tableau rows, source/tab.c:327-337) is per problem.
 */
#ifndef PIP_CONVERT_H
#define PIP_CONVERT_H

#include "pip_types.h"
#include "simt.h"

/* what pip_solve derives from (dimensions, Bg, options) before any tableau exists */
typedef struct {
  int dr, dc, cr, cc, has_ctx;     /* dense input dimensions */
  int Nn, Np;                      /* unknowns, parameters of the tableau (after Urs / big-parameter synthesis) */
  int Bg, Shift, Urs;              /* tableau column of the big parameter (or < 0), +1 / -1 / 0, Urs_parms copies */
  int pflags;                      /* PIP_F_* of every problem */
  int width, cwidth;               /* tableau row width Nn + Np + 1, context row width Np + 1 */
} PipConvertShape;

/* one PolyLib row -> one tableau row of `width` words (source/tab.c:307-326, 338-376).  `ctx` = the row
 * belongs to the context (n == -1 in the reference).  Bits lost by narrowing to T are OR-ed into `lost`.
 * Returns true for an inequality, false for an equality (the caller appends the negated row). */
template <class T>
PIP_HDM bool pip_convert_row(const pip_i64 *in, int cols, T *r, int width, int Nv, int ctx, int Shift, int Bg, int Urs,
                             pip_i64 &lost)
{
#define PIP_CV(dst, val) do { const pip_i64 v_ = (val); const T t_ = (T)v_; (dst) = t_; lost |= ((pip_i64)t_ ^ v_); } while (0)
  int ncolm = cols - 1;
  const bool isnew = Shift && (Bg + ctx > 0) && ((unsigned)(Bg + ctx) > (unsigned)(cols - 2));
  if (isnew) ncolm++;
  int cst;
  if (ctx) { Shift = 0; cst = Nv + Urs; } else cst = Nv;
  for (int j = 0; j < width; j++) r[j] = 0;
  pip_i64 big = 0;
  int j;
  for (j = 0; j < Nv; j++) {
    if (isnew && j == Bg) continue;
    if (Shift) big += in[1 + j];
    PIP_CV(r[j], Shift > 0 ? -in[1 + j] : in[1 + j]);
  }
  int k = Nv + 1;
  for (j = Nv + 1; j < ncolm; j++) {
    if (isnew && j == Bg) continue;
    PIP_CV(r[j], in[k]);
    k++;
  }
  for (j = 0; j < Urs; j++) {
    int pos_n = ncolm - ctx + j, pos = pos_n - Urs;
    if (pos <= Bg) --pos;
    PIP_CV(r[pos_n], -(pip_i64)r[pos]);
  }
  PIP_CV(r[cst], in[cols - 1]);
  if (Shift) {
    if (Shift < 0) big = -big;
    if (isnew) PIP_CV(r[Bg], big); else PIP_CV(r[Bg], (pip_i64)r[Bg] + big);
  }
#undef PIP_CV
  return in[0] != 0;
}

/* the negated copy of an equality's row (source/tab.c:327-337) */
template <class T>
PIP_HDM void pip_convert_negate(const T *r, T *r2, int width, pip_i64 &lost)
{
  for (int j = 0; j < width; j++) {
    const pip_i64 v = -(pip_i64)r[j];
    const T t = (T)v;
    r2[j] = t;
    lost |= ((pip_i64)t ^ v);
  }
}

/* device-side conversion of a chunk (pip_kernels.cu) */
typedef struct {
  PipConvertShape s;
  const pip_i64 *dom, *ctx;        /* raw rows in device memory: [n][dr][dc], [n][cr][cc] */
  long long n;
  long long stride;                /* pool elements reserved per problem: 2*dr*width + 2*cr*cwidth (all equalities) */
  void *pool;                      /* int32 or int64 elements */
  PipProblem *prob;
  int *dims;                       /* [0] max tableau rows, [1] max context rows, [2] problems whose input left int32 */
} PipConvertArgs;

#endif
