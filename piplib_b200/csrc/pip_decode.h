/* Solution cells -> serialised quast words: one implementation for the host (pip_host.cpp) and
 * the device (pip_serialize kernels).  Restates the decoder of the reference,
 * sol_quast_edit_xx / sol_newparm_edit_xx / sol_list_edit_xx / sol_vector_edit_xx
 * (source/sol.c:435-734), but emits the pre-order word stream of include/piplib_b200.h instead of
 * a malloc'd tree.  The cell stream is itself pre-order (If, then-subtree, else-subtree), so a
 * flat scan is enough: no recursion, which is what lets one GPU thread decode one problem.
 *
 *   node  := NNEWPARM { rank deno VEC }*  KIND ...
 *   KIND  := 0 (leaf "()")  | 1 LIST | 2 VEC(condition) node(then) node(else)
 *   LIST  := nvec { present(0/1) [VEC] }*  has_dual(0)
 *   VEC   := n { num den }*
 */
#ifndef PIP_DECODE_H
#define PIP_DECODE_H

#include "pip_arith.h"
#include "pip_types.h"

enum { PIP_SOL_SHIFT = 1, PIP_SOL_NEGATE = 2, PIP_SOL_REMOVE = 4, PIP_SOL_DUAL = 8 };   /* source/sol.h:35-48 */

struct PipSer {
  pip_i64 *out;            /* may be NULL: count / hash only */
  long long cap, len;
  pip_u64 h;
  int hashing;
  int narrow_out;          /* write 32-bit words (two per 64-bit slot) instead of 64-bit ones */
  unsigned wide;           /* set when some word does not fit 32 bits */
};
#define PIP_HASH_INIT 0xcbf29ce484222325ULL

/* hash of a serialised quast = PIP_HASH_INIT + sum over the words of a 64-bit mix of (word, index)
 * (the splitmix64 finaliser).  A sum, so that the lanes of a warp hash their own words and add up: the
 * chained form this replaces (h = (h ^ w) * K, word after word) was the longest dependency chain of the
 * decode kernel.  The checker libraries apply the same function to the reference's trees. */
PIP_HD pip_u64 pip_hash_word(pip_u64 v, pip_u64 k)
{
  pip_u64 x = v + (k + 1ull) * 0x9E3779B97F4A7C15ULL;
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
  x ^= x >> 27; x *= 0x94D049BB133111EBULL;
  x ^= x >> 31;
  return x;
}

PIP_HD void pip_sput(PipSer &s, pip_i64 v)
{
  if (s.hashing) s.h += pip_hash_word((pip_u64)v, (pip_u64)s.len);
  s.wide |= (unsigned)(v != (pip_i64)(int)v);
  if (s.out && s.len < s.cap) {
    if (s.narrow_out) ((int *)s.out)[s.len] = (int)v;
    else s.out[s.len] = v;
  }
  s.len++;
}

/* raw {kind, p1, p2} cells */
struct PipRawCells {
  const PipCell *c;
  PIP_HDM int kind(int i) const { return c[i].kind; }
  PIP_HDM pip_i64 p1(int i) const { return c[i].p1; }
  PIP_HDM pip_i64 p2(int i) const { return c[i].p2; }
};

/* sol_vector_edit_xx, source/sol.c:435-512 */
template <class C>
PIP_HD void pip_ser_vector(PipSer &s, const C &c, int &i, int Bg, int Urs_p, int flags)
{
  int n = (int)c.p1(i), unbounded = 0;
  if (flags & PIP_SOL_REMOVE) --n;
  n -= Urs_p;
  const int first_urs = Urs_p + (Bg >= 0);
  /* the unbounded marker rewrites every denominator, so it must be known before emitting */
  if (flags & PIP_SOL_SHIFT) {
    int t = i;
    for (int j = 0, k = 0; k < n; j++) {
      t++;
      if (j == Bg && c.p1(t) - c.p2(t) != 0) unbounded = 1;
      if ((flags & PIP_SOL_REMOVE) && j == Bg) continue;
      if (first_urs <= j && j < first_urs + Urs_p) continue;
      k++;
    }
  }
  pip_sput(s, n);
  for (int j = 0, k = 0; k < n; j++) {
    i++;
    pip_i64 N = c.p1(i);
    const pip_i64 D = c.p2(i);
    const pip_i64 d = (D == 1) ? 1 : pip_gcd(N, D);
    if ((flags & PIP_SOL_SHIFT) && j == Bg) N -= D;
    if ((flags & PIP_SOL_REMOVE) && j == Bg) continue;
    if (first_urs <= j && j < first_urs + Urs_p) continue;
    pip_i64 num = d ? pip_div(N, d) : 0;
    if (flags & PIP_SOL_NEGATE) num = -num;
    pip_sput(s, num);
    pip_sput(s, unbounded ? 0 : ((d == D) ? 1 : (d ? pip_div(D, d) : 0)));
    k++;
  }
  i++;
}

/* the whole stream of one problem (n cells); returns false on a malformed stream */
template <class C>
PIP_HD bool pip_ser_cells(PipSer &s, const C &c, int n, int Bg, int Urs_p, int flags)
{
  int i = 0;
  bool dual_node = false;          /* the node being decoded is the dual list of the previous leaf */
  const int all_flags = flags;
  while (i < n) {
    while (i < n && c.kind(i) == PIP_C_FREE) i++;
    if (i >= n) break;
    flags = dual_node ? 0 : all_flags;                 /* sol_quast_edit_xx(i, solution, Bg, Urs_p, 0), source/sol.c:706-708 */
    int nnew = 0;
    for (int t = i; t < n && c.kind(t) == PIP_C_NEW; t += (int)c.p1(t + 2) + 4) nnew++;   /* New Div Form Val*m Val */
    pip_sput(s, nnew);
    for (int k = 0; k < nnew; k++) {
      const int newcell = i;
      i += 2;
      int rank = (int)c.p1(newcell);
      if (flags & PIP_SOL_REMOVE) rank--;
      rank -= Urs_p;
      pip_sput(s, rank);
      pip_sput(s, c.p1(i + (int)c.p1(i) + 1));          /* the divisor follows the form */
      pip_ser_vector(s, c, i, Bg, Urs_p, flags & PIP_SOL_REMOVE);
      i++;
    }
    const int kind = c.kind(i);
    const int nb = (int)c.p1(i);
    i++;
    if (kind == PIP_C_LIST) {
      pip_sput(s, 1);
      if (nb == 0) { pip_sput(s, 1); pip_sput(s, 0); }
      else {
        pip_sput(s, nb);
        for (int e = 0; e < nb; e++) { pip_sput(s, 1); pip_ser_vector(s, c, i, Bg, Urs_p, flags); }
      }
      if (!dual_node && (all_flags & PIP_SOL_DUAL)) { pip_sput(s, 1); dual_node = true; }   /* the dual list follows */
      else { pip_sput(s, 0); dual_node = false; }
    } else if (kind == PIP_C_NIL) {
      pip_sput(s, 0);
      dual_node = false;
    } else if (kind == PIP_C_IF) {
      pip_sput(s, 2);
      pip_ser_vector(s, c, i, Bg, Urs_p, flags & PIP_SOL_REMOVE);
      dual_node = false;
    } else return false;
  }
  return true;
}

#endif
