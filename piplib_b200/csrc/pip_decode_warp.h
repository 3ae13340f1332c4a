/* Solution cells -> serialised quast words, one WARP per problem (device, pass 1 of the decode).
 *
 * Same grammar and same arithmetic as pip_decode.h (which restates sol_quast_edit_xx /
 * sol_newparm_edit_xx / sol_list_edit_xx / sol_vector_edit_xx, source/sol.c:435-734); what changes
 * is the mapping to the machine.  With one thread per problem every lane of a warp walks a
 * different tree, so the warp diverges completely and every dependent load of every lane is paid
 * one after the other (5.4 ms per 131072 problems, 40x off the HBM time of the bytes moved).  Here
 * the control flow is warp-uniform: the cells of the problem are staged in shared memory with one
 * coalesced copy, all lanes parse the same stream, the entries of a vector (9 of 10 cells are
 * `Val`s) are reduced lane-parallel, and the words leave through a shared-memory tile that is
 * hashed (every lane its own words, the hash is a sum) and written out coalesced.
 */
#ifndef PIP_DECODE_WARP_H
#define PIP_DECODE_WARP_H

#include "pip_decode.h"
#include "simt.h"

#define PIP_WS_TILE 256        /* words per staging tile (per warp) */
#define PIP_WS_CELLS 288       /* cells staged in shared memory (per warp); longer streams are read in place */

struct PipWarpSer {
  pip_i64 *tile;          /* PIP_WS_TILE words of shared memory owned by the warp */
  pip_i64 *out;           /* destination (64-bit slots), may be NULL */
  long long cap, len;     /* words the destination holds / words emitted so far (warp-uniform) */
  int fill;               /* words waiting in the tile (warp-uniform) */
  pip_u64 h;              /* this LANE's share of the hash sum (pip_hash_word); pip_wser_hash adds the lanes up */
  int narrow_out;
  unsigned wide;
};

/* hash the tile (the chain of pip_sput, in order), note words that leave 32 bits, write it out.
 * Out of line and by value: the flush is reached from every place a word is put, and inlining it
 * there made the kernel 11 k instructions long -- an instruction-cache problem, not a decoder. */
struct PipFlushOut { pip_u64 h; unsigned wide; };
PIP_DEVNI PipFlushOut pip_wser_flush_tile(const pip_i64 *tile, int n, pip_u64 h, pip_i64 *out, long long base, long long cap,
                                        int narrow_out)
{
  W::sync();
  const int lane = W::lane();
  bool w = false;
  for (int k = lane; k < n; k += 32) {
    const pip_i64 v = tile[k];
    h += pip_hash_word((pip_u64)v, (pip_u64)(base + k));
    w = w || (v != (pip_i64)(int)v);
    if (out && base + k < cap) {
      if (narrow_out) ((int *)out)[base + k] = (int)v;
      else out[base + k] = v;
    }
  }
  PipFlushOut r;
  r.h = h;
  r.wide = W::any(w) ? 1u : 0u;
  W::sync();
  return r;
}

PIP_DEV void pip_wser_flush(PipWarpSer &s)
{
  const PipFlushOut r = pip_wser_flush_tile(s.tile, s.fill, s.h, s.out, s.len - s.fill, s.cap, s.narrow_out);
  s.h = r.h;
  s.wide |= r.wide;
  s.fill = 0;
}

/* the hash of the whole stream once everything is flushed: start value + the lanes' shares */
PIP_DEV pip_u64 pip_wser_hash(const PipWarpSer &s)
{
  pip_u64 h = s.h;
  for (int o = 16; o > 0; o >>= 1) h += (pip_u64)W::shfl_xor64((long long)h, o);
  return h + PIP_HASH_INIT;
}

PIP_DEV void pip_wput(PipWarpSer &s, pip_i64 v)
{
  if (W::lane() == 0) s.tile[s.fill] = v;
  s.fill++; s.len++;
  if (s.fill == PIP_WS_TILE) pip_wser_flush(s);
}

/* sol_vector_edit_xx, source/sol.c:435-512; i is the index of the Form cell */
template <class C>
PIP_DEV void pip_wser_vector(PipWarpSer &s, const C &c, int &i, int Bg, int Urs_p, int flags)
{
  int n = (int)c.p1(i);
  if (flags & PIP_SOL_REMOVE) --n;
  n -= Urs_p;
  const int first_urs = Urs_p + (Bg >= 0);
  const bool rm = (flags & PIP_SOL_REMOVE) != 0;
  /* J = Val cells the entry loop of the reference consumes: n kept ones plus the skipped ones among them */
  int J = n > 0 ? n : 0;
  if (n > 0 && (rm || Urs_p > 0)) {
    J = 0;
    for (int k = 0; k < n; J++) {
      const bool skip = (rm && J == Bg) || (first_urs <= J && J < first_urs + Urs_p);
      if (!skip) k++;
    }
  }
  int unbounded = 0;
  if ((flags & PIP_SOL_SHIFT) && Bg >= 0 && Bg < J) unbounded = (c.p1(i + 1 + Bg) - c.p2(i + 1 + Bg) != 0);
  const int words = 1 + 2 * J;                     /* upper bound of what this vector emits */
  if (n <= 0 || words > PIP_WS_TILE) {
    /* degenerate or very long vector: word by word, exactly the loop of pip_ser_vector */
    pip_wput(s, n);
    for (int j = 0, k = 0; k < n; j++) {
      pip_i64 N = c.p1(i + 1 + j);
      const pip_i64 D = c.p2(i + 1 + j);
      const pip_i64 d = (D == 1) ? 1 : pip_gcd(N, D);
      if ((flags & PIP_SOL_SHIFT) && j == Bg) N -= D;
      if (rm && j == Bg) continue;
      if (first_urs <= j && j < first_urs + Urs_p) continue;
      pip_i64 num = d ? pip_div(N, d) : 0;
      if (flags & PIP_SOL_NEGATE) num = -num;
      pip_wput(s, num);
      pip_wput(s, unbounded ? 0 : ((d == D) ? 1 : (d ? pip_div(D, d) : 0)));
      k++;
    }
  } else {
    if (s.fill + words > PIP_WS_TILE) pip_wser_flush(s);
    const int lane = W::lane(), base = s.fill;
    if (lane == 0) s.tile[base] = n;
    int kbase = 0;
    for (int j0 = 0; j0 < J; j0 += 32) {
      const int j = j0 + lane;
      const bool in = j < J;
      const bool keep = in && !((rm && j == Bg) || (first_urs <= j && j < first_urs + Urs_p));
      const unsigned km = W::ballot(keep);
      if (keep) {
        const int k = kbase + pip_popc(km & ((1u << lane) - 1u));
        pip_i64 N = c.p1(i + 1 + j);
        const pip_i64 D = c.p2(i + 1 + j);
        const pip_i64 d = (D == 1) ? 1 : pip_gcd(N, D);
        if ((flags & PIP_SOL_SHIFT) && j == Bg) N -= D;
        pip_i64 num = d == 1 ? N : (d ? pip_div(N, d) : 0);       /* the 64-bit division is a call: not for d = 1 */
        if (flags & PIP_SOL_NEGATE) num = -num;
        s.tile[base + 1 + 2 * k] = num;
        s.tile[base + 2 + 2 * k] = unbounded ? 0 : ((d == D) ? 1 : (d ? pip_div(D, d) : 0));
      }
      kbase += pip_popc(km);
    }
    s.fill += 1 + 2 * n; s.len += 1 + 2 * n;
    if (s.fill == PIP_WS_TILE) pip_wser_flush(s);
  }
  i += J + 1;
}

/* the whole stream of one problem (n cells), all lanes in lock step; false on a malformed stream */
template <class C>
PIP_DEV bool pip_wser_cells(PipWarpSer &s, const C &c, int n, int Bg, int Urs_p, int flags)
{
  int i = 0;
  bool dual_node = false;
  const int all_flags = flags;
  while (i < n) {
    while (i < n && c.kind(i) == PIP_C_FREE) i++;
    if (i >= n) break;
    flags = dual_node ? 0 : all_flags;
    int nnew = 0;
    for (int t = i; t < n && c.kind(t) == PIP_C_NEW; t += (int)c.p1(t + 2) + 4) nnew++;
    pip_wput(s, nnew);
    for (int k = 0; k < nnew; k++) {
      const int newcell = i;
      i += 2;
      int rank = (int)c.p1(newcell);
      if (flags & PIP_SOL_REMOVE) rank--;
      rank -= Urs_p;
      pip_wput(s, rank);
      pip_wput(s, c.p1(i + (int)c.p1(i) + 1));
      pip_wser_vector(s, c, i, Bg, Urs_p, flags & PIP_SOL_REMOVE);
      i++;
    }
    const int kind = c.kind(i);
    const int nb = (int)c.p1(i);
    i++;
    if (kind == PIP_C_LIST) {
      pip_wput(s, 1);
      if (nb == 0) { pip_wput(s, 1); pip_wput(s, 0); }
      else {
        pip_wput(s, nb);
        for (int e = 0; e < nb; e++) { pip_wput(s, 1); pip_wser_vector(s, c, i, Bg, Urs_p, flags); }
      }
      if (!dual_node && (all_flags & PIP_SOL_DUAL)) { pip_wput(s, 1); dual_node = true; }
      else { pip_wput(s, 0); dual_node = false; }
    } else if (kind == PIP_C_NIL) {
      pip_wput(s, 0);
      dual_node = false;
    } else if (kind == PIP_C_IF) {
      pip_wput(s, 2);
      pip_wser_vector(s, c, i, Bg, Urs_p, flags & PIP_SOL_REMOVE);
      dual_node = false;
    } else return false;
  }
  return true;
}

#endif
