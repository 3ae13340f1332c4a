/* TEST INFRASTRUCTURE ONLY -- CPU restatement of PipLib's parametric dual simplex + Gomory cuts.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this; the product library never does (it has no CPU path at all).
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against every golden
 * vector of the reference (the 33 test .dat/.ll pairs and 15 example .pip/.ll pairs, committed under
 * tests/golden/) and, in this container, against the unmodified reference compiled into
 * oracle/_ref/libpipref.so on the fixtures without goldens and on seeded random problems.
 */
#ifndef PIP_ORACLE_H
#define PIP_ORACLE_H

typedef long long pio_int;

/* status codes shared by the oracle, the reference harness and the product:
 *   0        solved, cells valid
 *   1        context empty ("void", source/piplib.c:872-873, source/maind.c:228)
 *   1000+c   the reference would have printed a message and called exit(c):
 *            1001 "Integer overflow" (source/traiter.c:424-427,441-444) or
 *                 "Too much parameters" (source/traiter.c:174-177)
 *            1002 "Too much parameters" at a split (source/traiter.c:710-713)
 *            1003 "Too many variables" (source/integrer.c:324-327)
 *            1026 "The solution is too complex! : sol" (source/sol.c:97-100)
 *   2000     arithmetic fault the reference would die of (division by zero, SIGFPE)
 */
#define PIO_OK 0
#define PIO_VOID 1
#define PIO_FATAL 1000
#define PIO_FAULT 2000

typedef struct {
  long long pivots;        /* successful passes through the pivot (source/traiter.c:394-548) */
  long long cuts_const;    /* case (d) cuts, source/integrer.c:409-481 */
  long long cuts_parm;     /* case (e) cuts, source/integrer.c:493-520 */
  long long traiter_calls; /* activations of the solver loop incl. sub-solves */
  long long compa_rows;    /* rows sign-tested by the compatibility test */
  long long splits;
  long long max_rows, max_cols, max_depth;
  long long elem_updates;  /* sum over pivots of (updated rows x columns) */
  long long max_ctx_rows;
  long long wrapped;       /* # of 64-bit products whose exact value did not fit (informative) */
} pio_stats;

#endif
