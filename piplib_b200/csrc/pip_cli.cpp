/* pip64 -- the command-line front end of PipLib over the B200 batch solver.
 *
 * Drop-in for the reference's `pip` (source/maind.c:76-250): same arguments
 *     pip64 [-s | -v[v..]] [-d] [-z] [input [output]]
 * same input language (doc/piplib.texi:570-659), same output text (the .ll files of test/), same
 * messages and exit codes for the fatal verdicts.  What differs is the schedule: the reference reads,
 * solves and prints one problem at a time; this front end lexes the WHOLE input first, hands every
 * well-formed problem to ONE pip_traiter_batch_dp call (one GPU launch sequence for the file) and
 * prints the answers in input order.
 *
 * The lexer restates the character-level behaviour of dgetc_xx / dscanf_xx (source/piplib.c:76-163),
 * balance_xx / escape_xx (source/maind.c:49-74) and tab_get_xx (source/tab.c:222-248) on an in-memory
 * copy of the input, including tab_get's habit of skipping to the next ']' after the last row of a
 * tableau (which makes a problem with an empty context "()" swallow text of the next one).  One
 * liberty: the reference refills its line buffer 255 characters at a time inside dscanf and can split
 * a number that straddles such a boundary; here numbers are never split.
 *
 * The printer restates sol_edit_xx (source/sol.c:291-422).
 */
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "piplib_b200.h"

namespace {

struct Lexer {
  std::string s;
  size_t at = 0;
  int getc_() { return at < s.size() ? (unsigned char)s[at++] : EOF; }
  /* dscanf_xx: skip blanks (space, tab, newline), read an optionally signed decimal */
  int scan(long long *v)
  {
    while (at < s.size() && (s[at] == ' ' || s[at] == '\n' || s[at] == '\t')) at++;
    if (at >= s.size()) return EOF;
    char *end = nullptr;
    const char *p = s.c_str() + at;
    /* sscanf("%lld") accepts leading blanks (none left), a sign, digits */
    if (!(isdigit((unsigned char)*p) || ((*p == '-' || *p == '+') && isdigit((unsigned char)p[1])))) return -1;
    *v = strtoll(p, &end, 10);
    /* the reference then steps over '-' and digits only */
    while (at < s.size() && (s[at] == '-' || isdigit((unsigned char)s[at]))) at++;
    (void)end;
    return 0;
  }
};

struct Problem {
  std::string comment;              /* echo of balance_xx */
  int nvar = 0, nparm = 0, ni = 0, nc = 0, bigparm = -1, nq = 0;
  std::vector<long long> tab, ctx;
  bool syntax_error = false;        /* escape_xx was taken ... */
  bool escape_closed = false;       /* ... and found its closing parenthesis: "\nSyntax error\n)\n" is printed */
};

/* balance_xx, source/maind.c:49-61 */
void balance(Lexer &in, std::string &out)
{
  int level = 0, c;
  while ((c = in.getc_()) != EOF) {
    if (c == '(') level++;
    else if (c == ')' && --level == 0) return;
    out.push_back((char)c);
  }
}
/* escape_xx, source/maind.c:63-74: true when the closing parenthesis was found */
bool escape(Lexer &in, int level)
{
  int c;
  while ((c = in.getc_()) != EOF) {
    if (c == '(') level++;
    else if (c == ')' && --level == 0) return true;
  }
  return false;
}
/* tab_get_xx, source/tab.c:222-248 */
bool tab_get(Lexer &in, int h, int w, std::vector<long long> &out)
{
  int c;
  out.assign((size_t)(h > 0 ? h : 0) * (size_t)(w > 0 ? w : 0), 0);
  while ((c = in.getc_()) != EOF) if (c == '(') break;
  for (int i = 0; i < h; i++) {
    while ((c = in.getc_()) != EOF) if (c == '[') break;
    for (int j = 0; j < w; j++) {
      long long x;
      if (in.scan(&x) < 0) return false;
      out[(size_t)i * w + j] = x;
    }
  }
  while ((c = in.getc_()) != EOF) if (c == ']') break;
  return true;
}

/* sol_edit_xx, source/sol.c:291-422 on the cells of one problem; returns the next cell index */
long long gcd_ll(long long a, long long b)
{
  unsigned long long x = a < 0 ? 0ull - (unsigned long long)a : (unsigned long long)a;
  unsigned long long y = b < 0 ? 0ull - (unsigned long long)b : (unsigned long long)b;
  while (y) { unsigned long long r = x % y; x = y; y = r; }
  return (long long)x;
}
void print_val(FILE *f, long long N, long long D)
{
  const long long d = gcd_ll(N, D);
  if (d == D) { fprintf(f, " %lld", d ? N / d : N); return; }
  fprintf(f, " %lld/%lld", d ? N / d : N, d ? D / d : D);
}
int sol_edit(FILE *f, const PipCell_dp *c, int n, int i)
{
  for (;;) {
    if (i >= n) return n;
    if (c[i].kind == 0) { i++; continue; }                          /* Free */
    if (c[i].kind == 5) {                                           /* New */
      fprintf(f, "(newparm %d ", (int)c[i].p1);
      i = sol_edit(f, c, n, i + 1);
      fprintf(f, ")\n");
      continue;
    }
    break;
  }
  switch (c[i].kind) {
  case 1: fprintf(f, "()\n"); i++; break;                           /* Nil */
  case 8: fprintf(f, "Error %d\n", (int)c[i].p1); i++; break;       /* Error */
  case 2:                                                           /* If */
    fprintf(f, "(if ");
    i = sol_edit(f, c, n, i + 1);
    i = sol_edit(f, c, n, i);
    i = sol_edit(f, c, n, i);
    fprintf(f, ")\n");
    break;
  case 3: {                                                         /* List */
    fprintf(f, "(list ");
    int k = (int)c[i].p1;
    i++;
    while (k--) i = sol_edit(f, c, n, i);
    fprintf(f, ")\n");
    break;
  }
  case 4: {                                                         /* Form */
    fprintf(f, "#[");
    const int k = (int)c[i].p1;
    for (int j = 0; j < k; j++) { i++; print_val(f, c[i].p1, c[i].p2); }
    fprintf(f, "]\n");
    i++;
    break;
  }
  case 6:                                                           /* Div */
    fprintf(f, "(div ");
    i = sol_edit(f, c, n, i + 1);
    i = sol_edit(f, c, n, i);
    fprintf(f, ")\n");
    break;
  case 7: print_val(f, c[i].p1, c[i].p2); i++; break;               /* Val */
  default: fprintf(f, "Inconnu : sol\n");
  }
  return i;
}

/* the reference's last words (source/traiter.c:424,442,711; source/integrer.c:325; source/sol.c:98)
 * and exit code for a fatal verdict.  The counts it prints with two of them are not known to the
 * batch API; the header values are printed instead. */
int report_fatal(int status, const Problem &P)
{
  switch (status) {
  case 1001: fprintf(stderr, "Integer overflow\n"); return 1;
  case 1002: fprintf(stdout, "Too much parameters : %d\n", P.nparm); return 2;
  case 1003: fprintf(stderr, "Too many variables: %d\n", P.nvar + P.nparm + 1); return 3;
  case 1026: fprintf(stderr, "The solution is too complex! : sol\n"); return 26;
  case 2000: fprintf(stderr, "Floating point exception\n"); return 136;
  }
  fprintf(stderr, "pip64: the problem exceeds the device size classes (status %d)\n", status);
  return 1;
}

}  // namespace

int main(int argc, char **argv)
{
  int p = 1, silent = 0, deepest = 0, simple = 0, lex_only = 0;
  FILE *in = stdin, *out = stdout;
  if (argc > 1 && strcmp(argv[1], "--lex-only") == 0) {   /* test hook: dump the lexed problems, no device */
    lex_only = 1; silent = 1;
    argv++; argc--;
  }
  if (argc > 1 && !lex_only) {
    if (strcmp(argv[1], "-s") == 0) { silent = 1; p = 2; }
    else if (strncmp(argv[1], "-v", 2) == 0) p = 2;                  /* the dump file is not produced */
    if (argc > p && strcmp(argv[p], "-d") == 0) { deepest = 1; p++; }
  }
  if (!silent) fprintf(stderr, "Version %s\n", pip_b200_version());
  if (argc > p) {
    if (strcmp(argv[p], "-z") == 0) { simple = 1; p++; }
    in = fopen(argv[p], "r");
    if (!in) { fprintf(stderr, "%s unaccessible\n", argv[p]); return 1; }
  }
  p++;
  if (argc > p) {
    out = fopen(argv[p], "w");
    if (!out) { fprintf(stderr, "%s unaccessible\n", argv[p]); return 2; }
  }

  /* ---- lex the whole input ------------------------------------------------------------------ */
  Lexer lx;
  {
    char buf[1 << 16];
    size_t k;
    while ((k = fread(buf, 1, sizeof buf, in)) > 0) lx.s.append(buf, k);
  }
  std::vector<Problem> probs;
  int c;
  while ((c = lx.getc_()) != EOF) {
    if (c != '(') continue;
    Problem P;
    balance(lx, P.comment);
    long long h[6];
    bool ok = true;
    for (int k = 0; k < 6 && ok; k++) if (lx.scan(&h[k]) < 0) ok = false;
    if (!ok) { P.syntax_error = true; P.escape_closed = escape(lx, 1); probs.push_back(P); continue; }
    P.nvar = (int)h[0]; P.nparm = (int)h[1]; P.ni = (int)h[2]; P.nc = (int)h[3]; P.bigparm = (int)h[4]; P.nq = (int)h[5];
    if (!tab_get(lx, P.ni, P.nvar + P.nparm + 1, P.tab)) { P.syntax_error = true; P.escape_closed = escape(lx, 2); probs.push_back(P); continue; }
    if (!tab_get(lx, P.nc, P.nparm + 1, P.ctx)) { P.syntax_error = true; P.escape_closed = escape(lx, 2); probs.push_back(P); continue; }
    probs.push_back(P);
  }

  if (lex_only) {
    for (const Problem &P : probs) {
      fprintf(out, "problem error=%d comment=%zu %d %d %d %d %d %d\n", (int)P.syntax_error, P.comment.size(), P.nvar,
              P.nparm, P.ni, P.nc, P.bigparm, P.nq);
      if (P.syntax_error) continue;
      for (long long v : P.tab) fprintf(out, " %lld", v);
      fprintf(out, "\n");
      for (long long v : P.ctx) fprintf(out, " %lld", v);
      fprintf(out, "\n");
    }
    return 0;
  }

  /* ---- one batch for the file ----------------------------------------------------------------- */
  std::vector<int> live;
  for (size_t i = 0; i < probs.size(); i++) if (!probs[i].syntax_error) live.push_back((int)i);
  const int n = (int)live.size();
  std::vector<PipTableauHeader_dp> hdr(n);
  std::vector<const long long *> tabs(n), ctxs(n);
  std::vector<int> status(n, 0), ncells(n, 0);
  std::vector<long long> off(n, 0);
  static const long long zero = 0;
  for (int q = 0; q < n; q++) {
    const Problem &P = probs[live[q]];
    hdr[q].nvar = P.nvar; hdr[q].nparm = P.nparm; hdr[q].ni = P.ni; hdr[q].nc = P.nc; hdr[q].bigparm = P.bigparm;
    hdr[q].nq = (P.nq ? 1 : 0) | (deepest ? 4 : 0);
    tabs[q] = P.tab.empty() ? &zero : P.tab.data();
    ctxs[q] = P.ctx.empty() ? &zero : P.ctx.data();
  }
  std::vector<PipCell_dp> cells;
  if (n > 0) {
    long long cap = 4096ll * (n < 64 ? n : 64) + 4096, need = 0;
    for (;;) {
      cells.assign((size_t)cap, PipCell_dp());
      const int rc = pip_traiter_batch_dp(n, hdr.data(), tabs.data(), ctxs.data(), status.data(), cells.data(), cap,
                                          off.data(), ncells.data(), &need);
      if (rc == -2) { cap = need + 16; continue; }
      if (rc != 0) return 1;
      break;
    }
  }

  /* ---- print in input order (source/maind.c:171-236) ------------------------------------------- */
  int q = 0;
  for (size_t i = 0; i < probs.size(); i++) {
    const Problem &P = probs[i];
    fprintf(out, "(%s", P.comment.c_str());
    if (P.syntax_error) { if (P.escape_closed) fprintf(out, "\nSyntax error\n)\n"); continue; }
    const int st = status[q];
    if (st == 1) fprintf(out, "void\n");                              /* empty context */
    else if (st != 0) {
      /* the reference dies inside traiter_xx: everything printed so far stays, then message + exit */
      fflush(out);
      return report_fatal(st, P);
    } else {
      fputs(")\n", out);
      PipCell_dp *pc = cells.data() + off[q];
      int nc = ncells[q];
      if (simple) pip_cells_simplify_dp(pc, &nc);
      int at = 0;
      while (at < nc) at = sol_edit(out, pc, nc, at);
    }
    fprintf(out, ")\n");
    fflush(out);
    if (!silent) fprintf(stderr, "cross : (%ld), compa : (%d)\n\r", 0L, 0);
    q++;
  }
  return 0;
}
