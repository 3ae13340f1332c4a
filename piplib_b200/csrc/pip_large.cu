/* Cooperative grid-wide kernel for one large tableau (pip_large.h) + its host API. */
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <stdexcept>
#include <vector>

#include "../../include/piplib_b200.h"
#include "pip_engine.h"
#include "pip_large.h"

namespace cg = cooperative_groups;

__device__ void G::grid_sync() { cg::this_grid().sync(); }

#define PIPL_THREADS 512

__global__ void __launch_bounds__(PIPL_THREADS) pip_large_kernel(const PipLarge L)
{
  __shared__ __align__(16) int red[PIPL_RED_INTS];
  extern __shared__ __align__(128) unsigned char pipl_dyn[];
  pipl_solve(L, red, L.staged ? (pip_i64 *)pipl_dyn : nullptr);
}

/* restore the working tableau from the pristine copy and reset the control block */
__global__ void pip_large_reset_kernel(pip_i64 *dst, const pip_i64 *src, size_t words, int *ctl, pip_i64 *ctl64, int ni,
                                       unsigned long long *prof)
{
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
  for (size_t i = i0; i < words; i += step) dst[i] = src[i];
  if (i0 == 0) {
    for (int k = 0; k < PIPL_NCTL; k++) ctl[k] = 0;
    ctl[PIPL_NI] = ni; ctl[PIPL_LDET] = 1; ctl[PIPL_STATUS] = PIP_ST_OK;
    for (int k = 0; k < 8; k++) ctl64[k] = 0;
    ctl64[2] = 1;
    for (int k = 0; k < 8; k++) prof[k] = 0;
  }
}

struct pip_large_problem {
  PipLarge L;
  pip_i64 *pristine = nullptr;
  size_t words = 0;
  int ni0 = 0;
  int grid = 0;
  size_t dyn = 0;                 /* dynamic shared memory: staging buffers of the update phase */
  cudaStream_t stream = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  std::vector<void *> allocs;
};

#define CKL(x) pip_cuda_check((x), #x)

/* the dynamic shared-memory limit of the kernel is process-wide: only ever raised, under a lock
 * (several problems may be created / run from different threads) */
static std::mutex g_large_attr_mu;
static size_t g_large_smem = 0;
static void pip_large_raise_smem(size_t bytes)
{
  std::lock_guard<std::mutex> g(g_large_attr_mu);
  if (bytes <= g_large_smem) return;
  CKL(cudaFuncSetAttribute(pip_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  g_large_smem = bytes;
}

extern "C" {

pip_large_problem *pip_large_create_dp(int nvar, int ni, int nq, const long long *tab, int cut_rows,
                                       int sol_size, int maxcol)
{
  pip_large_problem *P = nullptr;
  try {
    /* no PipEngine call here: the batch engine runs this from inside PipEngine::run (lock held) */
    const int dev = pip_engine_device();
    CKL(cudaSetDevice(dev));
    int sms = 0, major = 0;
    CKL(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CKL(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major < 10) throw std::runtime_error("piplib-b200: this library is built for sm_100a (B200) only");
    P = new pip_large_problem;
    PipLarge &L = P->L;
    memset(&L, 0, sizeof L);
    const int ncol = nvar + 1;
    L.nvar = nvar; L.ni = ni; L.flags = nq ? PIP_F_INT : 0;
    L.stride = (ncol + 1) & ~1;
    L.rcap = ni + (cut_rows > 0 ? cut_rows : 0);
    L.pcap = nvar + L.rcap;
    L.sol_size = sol_size > 0 ? sol_size : PIP_SOL_SIZE;
    L.maxcol = maxcol > 0 ? maxcol : PIP_MAXCOL;
    auto dalloc = [&](size_t bytes) { void *p = nullptr; CKL(cudaMalloc(&p, bytes ? bytes : 8)); P->allocs.push_back(p); return p; };
    P->words = (size_t)L.rcap * L.stride;
    L.data = (pip_i64 *)dalloc(P->words * 8);
    P->pristine = (pip_i64 *)dalloc((size_t)ni * L.stride * 8);
    L.den = (pip_i64 *)dalloc((size_t)L.pcap * 8);
    L.fl = (int *)dalloc((size_t)L.pcap * 4);
    L.csign = (signed char *)dalloc((size_t)L.pcap);
    L.colpos = (int *)dalloc((size_t)nvar * 4 + 16);
    L.sbits = (unsigned *)dalloc((size_t)((L.pcap + 31) / 32) * 4 + 16);
    L.active = (int *)dalloc((size_t)L.pcap * 4 + 16);
    L.cand = (int *)dalloc((size_t)(L.pcap > nvar ? L.pcap : nvar) * 4 + 16);
    L.member = (unsigned char *)dalloc((size_t)nvar + 16);
    L.cut = (pip_i64 *)dalloc((size_t)L.stride * 8);
    L.ctl = (int *)dalloc(PIPL_NCTL * 4);
    L.ctl64 = (pip_i64 *)dalloc(8 * 8);
    L.cells = (PipCell *)dalloc((size_t)L.sol_size * sizeof(PipCell));
    L.prof = (unsigned long long *)dalloc(8 * 8);
    P->ni0 = ni;
    /* upload with the row stride */
    std::vector<pip_i64> host((size_t)ni * L.stride, 0);
    for (int r = 0; r < ni; r++) memcpy(&host[(size_t)r * L.stride], tab + (size_t)r * ncol, sizeof(pip_i64) * ncol);
    CKL(cudaMemcpy(P->pristine, host.data(), host.size() * 8, cudaMemcpyHostToDevice));
    CKL(cudaStreamCreateWithFlags(&P->stream, cudaStreamNonBlocking));
    CKL(cudaEventCreate(&P->e0));
    CKL(cudaEventCreate(&P->e1));
    /* staged update (cp.async.bulk): the pivot row + one row per group of warps in shared memory */
    P->dyn = (size_t)(PIPL_NG + 1) * L.stride * sizeof(pip_i64);
    if (P->dyn > 200 * 1024 || getenv("PIPLIB_B200_NO_TMA")) P->dyn = 0;
    L.staged = P->dyn ? 1 : 0;
    if (P->dyn) pip_large_raise_smem(P->dyn);
    int per_sm = 0;
    CKL(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pip_large_kernel, PIPL_THREADS, P->dyn));
    if (per_sm < 1) per_sm = 1;
    P->grid = sms * per_sm;
    return P;
  } catch (const std::exception &e) {
    fprintf(stderr, "%s\n", e.what());
    pip_large_destroy_dp(P);          /* whatever was allocated before the failure */
    return nullptr;
  }
}

/* one solve: restore the tableau (untimed), then the cooperative kernel (timed with CUDA events) */
int pip_large_run_dp(pip_large_problem *P, float *kernel_ms)
{
  try {
    const size_t used = (size_t)P->ni0 * P->L.stride;
    pip_large_reset_kernel<<<1024, 256, 0, P->stream>>>(P->L.data, P->pristine, used, P->L.ctl, P->L.ctl64, P->ni0, P->L.prof);
    CKL(cudaGetLastError());
    void *args[] = {(void *)&P->L};
    CKL(cudaEventRecord(P->e0, P->stream));
    if (P->dyn) pip_large_raise_smem(P->dyn);
    CKL(cudaLaunchCooperativeKernel((void *)pip_large_kernel, dim3(P->grid), dim3(PIPL_THREADS), args, P->dyn, P->stream));
    CKL(cudaEventRecord(P->e1, P->stream));
    CKL(cudaStreamSynchronize(P->stream));
    if (kernel_ms) CKL(cudaEventElapsedTime(kernel_ms, P->e0, P->e1));
  } catch (const std::exception &e) {
    fprintf(stderr, "%s\n", e.what());
    return -1;
  }
  return 0;
}

/* results of the last run.  info = {pivots, cuts, rows skipped (identity updates), ni at the end} */
int pip_large_fetch_dp(pip_large_problem *P, int *status, PipCell_dp *cells, int cell_cap, int *ncells, long long *info)
{
  try {
    int ctl[PIPL_NCTL];
    CKL(cudaMemcpy(ctl, P->L.ctl, sizeof ctl, cudaMemcpyDeviceToHost));
    *status = ctl[PIPL_STATUS];
    *ncells = ctl[PIPL_NCELL];
    if (info) {
      unsigned long long prof[8];
      CKL(cudaMemcpy(prof, P->L.prof, sizeof prof, cudaMemcpyDeviceToHost));
      info[0] = ctl[PIPL_PIVOTS]; info[1] = ctl[PIPL_CUTS]; info[2] = (unsigned)ctl[PIPL_SKIPPED_LO]; info[3] = ctl[PIPL_NI];
      info[4] = (long long)prof[0]; info[5] = (long long)prof[1];
      for (int k = 2; k < 8; k++) info[4 + k] = (long long)prof[k];
    }
    if (cells && *ncells > 0 && *ncells <= cell_cap)
      CKL(cudaMemcpy(cells, P->L.cells, sizeof(PipCell) * (size_t)*ncells, cudaMemcpyDeviceToHost));
  } catch (const std::exception &e) {
    fprintf(stderr, "%s\n", e.what());
    return -1;
  }
  return 0;
}

void pip_large_destroy_dp(pip_large_problem *P)
{
  if (!P) return;
  for (void *p : P->allocs) cudaFree(p);
  if (P->stream) cudaStreamDestroy(P->stream);
  if (P->e0) cudaEventDestroy(P->e0);
  if (P->e1) cudaEventDestroy(P->e1);
  delete P;
}

}  // extern "C"
