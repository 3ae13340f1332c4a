/* TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * Harness around the UNMODIFIED reference (compiled from /root/reference/source by
 * oracle/Makefile into oracle/_ref/libpipref.so).  The reference reports every error with
 * fprintf + exit(code) (source/traiter.c:424-427, source/sol.c:97-100, ...).  The Makefile
 * compiles the reference with -Dexit=pipref_exit_hook; the hook below turns the exit into a
 * longjmp so one process can evaluate millions of problems including the fatal verdicts.
 *
 * Three drivers:
 *   pipref_traiter      the CLI path of source/maind.c:150-232 on an already-lexed problem
 *                       (tableau rows in .dat order), returning the raw solution cells
 *   pipref_solve_ser    pip_solve_dp (source/piplib.c:722) on PolyLib matrices, returning the
 *                       PipQuast serialised to an int64 stream (format below) and/or the text
 *                       of pip_quast_print_dp
 *   pipref_bench_dense  a timed loop of pip_solve_dp over a dense batch (CPU baseline)
 */
#define _GNU_SOURCE
#include <setjmp.h>
#include <signal.h>
#include <sys/time.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <fcntl.h>

#include "pip.h"

/* sol.c keeps the cell type private; this mirrors source/sol.c:37-40 */
struct ref_cell { int flags; long long param1, param2; };
extern struct ref_cell *sol_space_dp;
extern int verbose_dp, deepest_cut_dp;

static sigjmp_buf pipref_env;
static volatile int pipref_armed = 0;
static int pipref_timeout_ms = 0;
#define setjmp(e) sigsetjmp(e, 1)

void pipref_exit_hook(int code)
{
  if (pipref_armed) siglongjmp(pipref_env, 1000 + code);
  _exit(code);
}

/* arithmetic faults of the reference (division by zero after a silent wrap) and run-away
 * problems become statuses too: 2000 = SIGFPE, 2001 = SIGSEGV, 3000 = time limit */
static void on_signal(int sig)
{
  if (!pipref_armed) _exit(128 + sig);
  siglongjmp(pipref_env, sig == SIGFPE ? 2000 : sig == SIGSEGV ? 2001 : 3000);
}
void pipref_set_timeout_ms(int ms) { pipref_timeout_ms = ms; }
static void guard_on(void)
{
  struct itimerval it;
  static int installed = 0;
  if (!installed) { signal(SIGFPE, on_signal); signal(SIGSEGV, on_signal); signal(SIGALRM, on_signal); installed = 1; }
  if (pipref_timeout_ms <= 0) return;
  memset(&it, 0, sizeof it);
  it.it_value.tv_sec = pipref_timeout_ms / 1000;
  it.it_value.tv_usec = (pipref_timeout_ms % 1000) * 1000;
  setitimer(ITIMER_REAL, &it, NULL);
}
static void guard_off(void)
{
  struct itimerval it;
  if (pipref_timeout_ms <= 0) return;
  memset(&it, 0, sizeof it);
  setitimer(ITIMER_REAL, &it, NULL);
}

static int saved_stderr = -1, saved_stdout = -1;
static void quiet_begin(void)
{
  fflush(stdout); fflush(stderr);
  static int devnull = -1;
  if (devnull < 0) devnull = open("/dev/null", O_WRONLY);
  saved_stderr = dup(2); saved_stdout = dup(1);
  if (devnull >= 0) { dup2(devnull, 2); dup2(devnull, 1); }
}
static void quiet_end(void)
{
  fflush(stdout); fflush(stderr);
  if (saved_stderr >= 0) { dup2(saved_stderr, 2); close(saved_stderr); saved_stderr = -1; }
  if (saved_stdout >= 0) { dup2(saved_stdout, 1); close(saved_stdout); saved_stdout = -1; }
}

static void after_fatal(void)
{
  /* the arenas are in an undefined state after a longjmp out of traiter: drop them */
  pip_close_dp();
}

/* ---- CLI path (maind.c) ------------------------------------------------------------ */
/* returns 0 = solved (cells filled), 1 = void context, 1000+code = fatal exit(code) */
int pipref_traiter(int nvar, int nparm, int ni, int nc, int bigparm, int nq,
                   const long long *tab, const long long *ctx,
                   int *cell_flags, long long *cell_p1, long long *cell_p2,
                   int cap, int *ncells)
{
  Tableau_dp *ineq, *context, *ctxt;
  struct high_water_mark_dp hq;
  int i, j, p, q, xq, non_vide, rc;
  int ncol = nvar + nparm + 1;

  *ncells = 0;
  pip_init_dp();
  verbose_dp = -1;
  deepest_cut_dp = 0;
  pipref_armed = 1;
  quiet_begin();
  rc = setjmp(pipref_env);
  if (rc) {
    pipref_armed = 0;
    guard_off();
    quiet_end();
    after_fatal();
    return rc;
  }
  guard_on();
  hq = tab_hwm_dp();
  ineq = tab_alloc_dp(ni, ncol, nvar);              /* as tab_get_dp, source/tab.c:231-242 */
  for (i = 0; i < ni; i++) {
    ineq->row[nvar + i].flags = Unknown;
    ineq->row[nvar + i].d = 1;
    for (j = 0; j < ncol; j++) ineq->row[nvar + i].objet.val[j] = tab[(size_t)i * ncol + j];
  }
  if (nq) tab_simplify_dp(ineq, nvar);
  context = tab_alloc_dp(nc, nparm + 1, 0);
  for (i = 0; i < nc; i++) {
    context->row[i].flags = Unknown;
    context->row[i].d = 1;
    for (j = 0; j < nparm + 1; j++) context->row[i].objet.val[j] = ctx[(size_t)i * (nparm + 1) + j];
  }
  if (nq) tab_simplify_dp(context, nparm);
  xq = p = sol_hwm_dp();
  if (nc) {
    ctxt = expanser_dp(context, nparm, nc, nparm + 1, nparm, 0, 0);
    traiter_dp(ctxt, NULL, nparm, 0, nc, 0, -1, TRAITER_INT);
    non_vide = is_not_Nil_dp(p);
    sol_reset_dp(p);
  } else non_vide = 1;
  if (non_vide) {
    traiter_dp(ineq, context, nvar, nparm, ni, nc, bigparm, nq ? TRAITER_INT : 0);
    q = sol_hwm_dp();
    for (i = xq; i < q && (i - xq) < cap; i++) {
      cell_flags[i - xq] = sol_space_dp[i].flags;
      cell_p1[i - xq] = sol_space_dp[i].param1;
      cell_p2[i - xq] = sol_space_dp[i].param2;
    }
    *ncells = q - xq;
    sol_reset_dp(p);
  }
  tab_reset_dp(hq);
  pipref_armed = 0;
  guard_off();
  quiet_end();
  return non_vide ? 0 : 1;
}

/* ---- PipQuast serialisation ---------------------------------------------------------
 * pre-order int64 stream; identical code exists in the product (pip_quast_serialize_dp) and
 * in the tests (python) so trees of both libraries can be compared word for word:
 *   node  := NNEWPARM { rank deno VEC }*  KIND ...
 *   KIND  := 0 (leaf "()")  | 1 LIST | 2 VEC(condition) node(then) node(else)
 *   LIST  := nvec { present(0/1) [VEC] }*  has_dual(0/1) [node]
 *   VEC   := n { num den }*
 * a NULL quast (void) is the single word -1.
 */
static long long *ser_out; static long ser_cap, ser_len;
static void put(long long v) { if (ser_len < ser_cap) ser_out[ser_len] = v; ser_len++; }
static void ser_vec(PipVector_dp *v)
{
  int i;
  put(v->nb_elements);
  for (i = 0; i < v->nb_elements; i++) { put(v->the_vector[i]); put(v->the_deno[i]); }
}
static void ser_quast(PipQuast_dp *s)
{
  PipNewparm_dp *np; PipList_dp *l; long n = 0;
  if (!s) { put(-1); return; }
  for (np = s->newparm; np; np = np->next) n++;
  put(n);
  for (np = s->newparm; np; np = np->next) { put(np->rank); put(np->deno); ser_vec(np->vector); }
  if (s->condition) {
    put(2); ser_vec(s->condition); ser_quast(s->next_then); ser_quast(s->next_else);
  } else if (s->list) {
    put(1);
    n = 0; for (l = s->list; l; l = l->next) n++;
    put(n);
    for (l = s->list; l; l = l->next) { put(l->vector != NULL); if (l->vector) ser_vec(l->vector); }
    put(s->next_then != NULL);
    if (s->next_then) ser_quast(s->next_then);
  } else put(0);
}

static PipMatrix_dp *mk_matrix(int rows, int cols, const long long *data)
{
  PipMatrix_dp *m = pip_matrix_alloc_dp(rows, cols);
  if (rows > 0 && cols > 0) memcpy(m->p_Init, data, sizeof(long long) * (size_t)rows * cols);
  return m;
}

/* opts = {Nq, Verbose, Simplify, Deepest_cut, Maximize, Urs_parms, Urs_unknowns, Compute_dual}
 * returns 0 ok (quast may be void), 1000+code fatal.  ser may be NULL; text may be NULL. */
int pipref_solve_ser(int dom_rows, int dom_cols, const long long *dom,
                     int has_ctx, int ctx_rows, int ctx_cols, const long long *ctx,
                     int bg, const int *opts,
                     long long *ser, long cap, long *ser_n,
                     char *text, long text_cap)
{
  PipMatrix_dp *D, *C = NULL; PipOptions_dp *o; PipQuast_dp *q; int rc;
  pipref_armed = 1;
  quiet_begin();
  rc = setjmp(pipref_env);
  if (rc) {
    pipref_armed = 0; guard_off(); quiet_end(); after_fatal();
    if (ser_n) *ser_n = 0;
    return rc;                       /* matrices leak on the fatal path: test harness only */
  }
  guard_on();
  D = mk_matrix(dom_rows, dom_cols, dom);
  if (has_ctx) C = mk_matrix(ctx_rows, ctx_cols, ctx);
  o = pip_options_init_dp();
  o->Nq = opts[0]; o->Verbose = -1; o->Simplify = opts[2]; o->Deepest_cut = opts[3];
  o->Maximize = opts[4]; o->Urs_parms = opts[5]; o->Urs_unknowns = opts[6]; o->Compute_dual = opts[7];
  q = pip_solve_dp(D, C, bg, o);
  pipref_armed = 0;
  guard_off();
  quiet_end();
  if (ser) { ser_out = ser; ser_cap = cap; ser_len = 0; ser_quast(q); if (ser_n) *ser_n = ser_len; }
  if (text && text_cap > 0) {
    char *buf = NULL; size_t len = 0; FILE *f = open_memstream(&buf, &len);
    pip_quast_print_dp(f, q, 0); fclose(f);
    if ((long)len >= text_cap) len = text_cap - 1;
    memcpy(text, buf, len); text[len] = 0; free(buf);
  }
  pip_quast_free_dp(q);
  pip_options_free_dp(o);
  pip_matrix_free_dp(D);
  if (C) pip_matrix_free_dp(C);
  return 0;
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
static unsigned long long hash_h, hash_k;
static void hput(long long v) { hash_h += hash_word((unsigned long long)v, hash_k++); }
static void hash_vec(PipVector_dp *v)
{ int i; hput(v->nb_elements); for (i = 0; i < v->nb_elements; i++) { hput(v->the_vector[i]); hput(v->the_deno[i]); } }
static void hash_quast(PipQuast_dp *s)
{
  PipNewparm_dp *np; PipList_dp *l; long n = 0;
  if (!s) { hput(-1); return; }
  for (np = s->newparm; np; np = np->next) n++;
  hput(n);
  for (np = s->newparm; np; np = np->next) { hput(np->rank); hput(np->deno); hash_vec(np->vector); }
  if (s->condition) { hput(2); hash_vec(s->condition); hash_quast(s->next_then); hash_quast(s->next_else); }
  else if (s->list) {
    hput(1); n = 0; for (l = s->list; l; l = l->next) n++; hput(n);
    for (l = s->list; l; l = l->next) { hput(l->vector != NULL); if (l->vector) hash_vec(l->vector); }
    hput(s->next_then != NULL); if (s->next_then) hash_quast(s->next_then);
  } else hput(0);
}

/* Timed loop of pip_solve_dp over problems [first, first+count) of a dense batch (all problems
 * share one shape).  hashes[i] = hash_word sum of the serialised quast (0 for fatal), status[i] as
 * above.  Returns seconds spent inside the solve loop (matrix setup included, as a real caller
 * pays it; hashing and freeing excluded by a second clock). */
double pipref_bench_dense(long first, long count,
                          int dom_rows, int dom_cols, const long long *dom,
                          int has_ctx, int ctx_rows, int ctx_cols, const long long *ctx,
                          int bg, const int *opts,
                          int *status, unsigned long long *hashes)
{
  volatile long i; volatile double total = 0; struct timespec t0, t1;
  PipOptions_dp *o = pip_options_init_dp();
  PipMatrix_dp *D = pip_matrix_alloc_dp(dom_rows, dom_cols);
  PipMatrix_dp *C = has_ctx ? pip_matrix_alloc_dp(ctx_rows, ctx_cols) : NULL;
  o->Nq = opts[0]; o->Verbose = -1; o->Simplify = opts[2]; o->Deepest_cut = opts[3];
  o->Maximize = opts[4]; o->Urs_parms = opts[5]; o->Urs_unknowns = opts[6]; o->Compute_dual = opts[7];
  quiet_begin();
  for (i = first; i < first + count; i++) {
    PipQuast_dp *q = NULL; int rc;
    size_t dsz = (size_t)dom_rows * dom_cols, csz = (size_t)ctx_rows * ctx_cols;
    pipref_armed = 1;
    guard_on();                          /* the time limit (if any) is per problem */
    rc = setjmp(pipref_env);
    if (rc) {
      pipref_armed = 0; after_fatal();
      if (status) status[i - first] = rc;
      if (hashes) hashes[i - first] = 0;
      continue;
    }
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (dsz) memcpy(D->p_Init, dom + (size_t)i * dsz, sizeof(long long) * dsz);
    if (C && csz) memcpy(C->p_Init, ctx + (size_t)i * csz, sizeof(long long) * csz);
    q = pip_solve_dp(D, C, bg, o);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    guard_off();
    pipref_armed = 0;
    total += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    if (status) status[i - first] = 0;
    if (hashes) { hash_h = 0xcbf29ce484222325ULL; hash_k = 0; hash_quast(q); hashes[i - first] = hash_h; }
    pip_quast_free_dp(q);
  }
  quiet_end();
  pip_options_free_dp(o);
  pip_matrix_free_dp(D);
  if (C) pip_matrix_free_dp(C);
  return total;
}
