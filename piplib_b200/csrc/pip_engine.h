/* Host-side batch engine: owns the device buffers, drives the size-class ladder
 * (S: shared-memory arenas, G3..G8: global-memory arenas of growing capacity), compacts and
 * fetches the solution cells.  Internal to the library (the C-ABI is include/piplib_b200.h). */
#ifndef PIP_ENGINE_H
#define PIP_ENGINE_H

#include <cuda_runtime.h>

#include <vector>

#include "pip_types.h"

struct PipBatchIn {
  size_t n = 0;
  const PipProblem *h_prob = nullptr;   /* host descriptors: always required (class planning) */
  const void *h_pool = nullptr;         /* host pool, uploaded unless d_pool is given */
  size_t pool_words = 0;                /* elements in the pool */
  int elem_log2 = 3;                    /* log2(bytes per pool element): 0, 2 or 3 */
  const PipProblem *d_prob = nullptr;   /* optional device-resident copies */
  const void *d_pool = nullptr;
  bool fetch_cells = true;              /* copy the compacted cells back to the host */
  const PipDecodeParm *h_decode = nullptr; /* if set: decode on the device, ship serialised quast words
                                              (+ hashes) instead of cells */
  int sol_size = PIP_SOL_SIZE, maxcol = PIP_MAXCOL;
};

struct PipBatchTimes {
  double h2d = 0, kernel = 0, d2h = 0, total = 0;   /* seconds, wall clock around synchronised phases */
  float device_ms = 0;                               /* CUDA-event time first launch -> last kernel */
  int launches = 0, rounds = 0;
  size_t h2d_bytes = 0, d2h_bytes = 0;
  unsigned long long phase_cycles[PIP_NPHASE] = {0};   /* profile build only */
  float round_s[8] = {0};                              /* wall seconds of each solve round */
  int round_n[8] = {0};                                /* problems in each round */
};

/* read-only view of one problem's cells in the wire format (pip_types.h) */
struct PipCellView {
  const pip_u64 *w;
  bool wide;
  int n;
  int kind(int i) const { return wide ? (int)(w[3 * i] & 0xffffffffull) : PIP_CELL_KIND(w[i]); }
  pip_i64 p1(int i) const { return wide ? (pip_i64)w[3 * i + 1] : PIP_CELL_P1(w[i]); }
  pip_i64 p2(int i) const { return wide ? (pip_i64)w[3 * i + 2] : PIP_CELL_P2(w[i]); }
};

struct PipBatchOut {
  std::vector<PipResult> res;           /* cell_off = word offset into the round's host chunk */
  std::vector<const pip_u64 *> base;    /* per problem: base pointer of its round's host chunk */
  std::vector<pip_u64> hashes;          /* device-decode mode: hash of each serialised quast */
  PipBatchTimes times;
  PipCellView cells_of(size_t i) const
  {
    PipCellView v;
    v.w = base[i] ? base[i] + res[i].cell_off : nullptr;
    v.wide = (res[i].rflags & PIP_RES_WIDE) != 0;
    v.n = res[i].ncells;
    return v;
  }
};

class PipEngine {
 public:
  static PipEngine &get() { return lane(0); }
  /* independent engines (own stream + buffers) so that batches can be pipelined: while one lane
   * waits for its kernels, another converts inputs or decodes results */
  enum { MAX_LANES = 8 };
  static PipEngine &lane(int i);
  /* engine-owned pinned staging for the input pool (valid until the next call on this lane) */
  void *pinned_input(size_t bytes);
  /* Runs the whole ladder.  Host cell storage referenced by `out` is engine-owned pinned memory,
   * valid until the next run() call.  Throws std::runtime_error on CUDA errors. */
  void run(const PipBatchIn &in, PipBatchOut &out);
  int set_device(int dev);
  int sm_count();
  cudaStream_t stream();
  struct Impl;

 private:
  PipEngine();
  Impl *impl_;
};

void pip_cuda_check(cudaError_t e, const char *what);
int pip_engine_device();

#endif
