/* piplib-b200: PipLib-compatible public header (int64 "dp" width only).
 *
 * Drop-in for the reference's <piplib/piplib.h> when PIPLIB_INT_DP is selected (the reference
 * selects the width at include/piplib/piplib.h:42-108; <piplib/piplib64.h> -> piplib_dp.h:31-47).
 * Structures have the reference's field order and LP64 layout (include/piplib/piplib.h:194-329),
 * functions the reference's names with the _dp suffix (include/piplib/piplib.h:331-402) and the
 * un-suffixed legacy aliases (include/piplib/piplib.h:534-573).  The solver behind them is the
 * sm_100a CUDA implementation; there is no CPU path.
 *
 * New in this library: pip_solve_batch_dp and the tableau-level batch entry points declared in
 * <piplib_b200.h>.
 */
#ifndef PIPLIB_B200_PIPLIB_H
#define PIPLIB_B200_PIPLIB_H

#include <stdio.h>

#if defined(PIPLIB_INT_SP) || defined(PIPLIB_INT_GMP)
#error "piplib-b200 implements the 64-bit (PIPLIB_INT_DP / piplib64.h) width only"
#endif
#ifndef PIPLIB_INT_DP
#define PIPLIB_INT_DP 1
#endif

#define PIPLIB_NAME(name) name##_dp

typedef long long int piplib_int_t_dp;
#define piplib_int_format "%lld"

#if defined(__cplusplus)
extern "C" {
#endif

/* PolyLib-style constraint matrix: row = [eq(0)/ineq(1) | unknowns | parameters | constant] */
struct pipmatrix_dp {
  unsigned int NbRows, NbColumns;
  piplib_int_t_dp **p;
  piplib_int_t_dp *p_Init;
  int p_Init_size;
};
typedef struct pipmatrix_dp PipMatrix_dp;

struct pipvector_dp {
  int nb_elements;
  piplib_int_t_dp *the_vector;   /* numerators */
  piplib_int_t_dp *the_deno;     /* denominators */
};
typedef struct pipvector_dp PipVector_dp;

struct pipnewparm_dp {
  int rank;
  PipVector_dp *vector;
  piplib_int_t_dp deno;
  struct pipnewparm_dp *next;
};
typedef struct pipnewparm_dp PipNewparm_dp;

struct piplist_dp {
  PipVector_dp *vector;
  struct piplist_dp *next;
};
typedef struct piplist_dp PipList_dp;

struct pipquast_dp {
  PipNewparm_dp *newparm;
  PipList_dp *list;
  PipVector_dp *condition;
  struct pipquast_dp *next_then;
  struct pipquast_dp *next_else;
  struct pipquast_dp *father;
};
typedef struct pipquast_dp PipQuast_dp;

struct pipoptions_dp {
  int Nq;            /* 1: integer solution, 0: rational */
  int Verbose;       /* accepted, ignored (no dump file) */
  int Simplify;      /* remove (if p () ()) */
  int Deepest_cut;
  int Maximize;
  int Urs_parms;
  int Urs_unknowns;
  int Compute_dual;
};
typedef struct pipoptions_dp PipOptions_dp;

void pip_options_print_dp(FILE *, PipOptions_dp *);
void pip_matrix_print_dp(FILE *, PipMatrix_dp *);
void pip_vector_print_dp(FILE *, PipVector_dp *);
void pip_newparm_print_dp(FILE *, PipNewparm_dp *, int);
void pip_list_print_dp(FILE *, PipList_dp *, int);
void pip_quast_print_dp(FILE *, PipQuast_dp *, int);

void pip_matrix_free_dp(PipMatrix_dp *);
void pip_vector_free_dp(PipVector_dp *);
void pip_newparm_free_dp(PipNewparm_dp *);
void pip_list_free_dp(PipList_dp *);
void pip_quast_free_dp(PipQuast_dp *);
void pip_options_free_dp(PipOptions_dp *);

PipMatrix_dp *pip_matrix_alloc_dp(unsigned int, unsigned int);
PipMatrix_dp *pip_matrix_read_dp(FILE *);
PipOptions_dp *pip_options_init_dp(void);

void pip_init_dp(void);
void pip_close_dp(void);

PipQuast_dp *pip_solve_dp(PipMatrix_dp *domain, PipMatrix_dp *parameters, int bignum, PipOptions_dp *options);

/* reference include/piplib/piplib.h:398-402 (source/sol.c:664-734): decode the quast starting at cell
 * *i of the solution space.  The reference reads its global sol_space (the cells of the last solve);
 * here that is the cells of the last pip_solve_dp on the calling thread (Bg = bignum - Nn - 1, flags
 * as in source/sol.h:35-48), or whatever pip_cells_bind_dp (piplib_b200.h) bound. */
PipQuast_dp *sol_quast_edit_dp(int *i, PipQuast_dp *father, int Bg, int Urs_p, int flags);

#if defined(__cplusplus)
}
#endif

/* legacy un-suffixed names */
#define piplib_int_t piplib_int_t_dp
#define Entier piplib_int_t_dp
#define PipMatrix PipMatrix_dp
#define PipVector PipVector_dp
#define PipNewparm PipNewparm_dp
#define PipList PipList_dp
#define PipQuast PipQuast_dp
#define PipOptions PipOptions_dp
#define pip_options_print pip_options_print_dp
#define pip_matrix_print pip_matrix_print_dp
#define pip_vector_print pip_vector_print_dp
#define pip_newparm_print pip_newparm_print_dp
#define pip_list_print pip_list_print_dp
#define pip_quast_print pip_quast_print_dp
#define pip_matrix_free pip_matrix_free_dp
#define pip_vector_free pip_vector_free_dp
#define pip_newparm_free pip_newparm_free_dp
#define pip_list_free pip_list_free_dp
#define pip_quast_free pip_quast_free_dp
#define pip_options_free pip_options_free_dp
#define pip_matrix_alloc pip_matrix_alloc_dp
#define pip_matrix_read pip_matrix_read_dp
#define pip_options_init pip_options_init_dp
#define pip_init pip_init_dp
#define pip_close pip_close_dp
#define pip_solve pip_solve_dp
#define sol_quast_edit sol_quast_edit_dp
#define pip_solve_batch pip_solve_batch_dp      /* the batched entry point (piplib_b200.h) */

#endif
