/* Warp-level primitives used by the solver core.
 *
 * On the device they are the sm_100a intrinsics.  With -DPIP_EMU (tests/emu only, a debugging
 * aid that runs the *device source* on the CPU with 32 cooperative fibers per warp) they are
 * provided by tests/emu/emu_runtime.cpp.  The product library is never built with PIP_EMU.
 */
#ifndef PIP_SIMT_H
#define PIP_SIMT_H

#include <stdint.h>

#if defined(PIP_EMU)

#define PIP_DEV static inline
#define PIP_DEVNI static __attribute__((noinline))
#define PIP_SDEV static inline
#define PIP_SDEVNI static __attribute__((noinline))
#define PIP_HD static inline
#define PIP_HDM inline
#define PIP_DM inline
#define PIP_HDNI static __attribute__((noinline))
#define PIP_ASSUME_SHARED(p) ((void)0)

namespace pipemu {
int lane();
void barrier();
unsigned ballot(bool p);
long long shfl64(long long v, int src);
unsigned redmin(unsigned v);
unsigned redmax(unsigned v);
unsigned atomic_add(unsigned *p, unsigned v);
}  // namespace pipemu

struct W {
  static inline int lane() { return pipemu::lane(); }
  static inline void sync() { pipemu::barrier(); }
  static inline unsigned ballot(bool p) { return pipemu::ballot(p); }
  static inline bool any(bool p) { return pipemu::ballot(p) != 0; }
  static inline int shfl(int v, int src) { return (int)pipemu::shfl64(v, src); }
  static inline long long shfl64(long long v, int src) { return pipemu::shfl64(v, src); }
  static inline int shfl_up(int v, int d) { int l = pipemu::lane(); int r = (int)pipemu::shfl64(v, l >= d ? l - d : l); return r; }
  static inline unsigned redmin(unsigned v) { return pipemu::redmin(v); }
  static inline unsigned redmax(unsigned v) { return pipemu::redmax(v); }
  static inline unsigned redor(unsigned v) { unsigned r = 0; for (int b = 0; b < 32; b++) if (pipemu::ballot((v >> b) & 1u)) r |= 1u << b; return r; }
  static inline int shfl_down(int v, int d) { int l = pipemu::lane(); return (int)pipemu::shfl64(v, l + d < 32 ? l + d : l); }
  static inline long long shfl_xor64(long long v, int m) { return pipemu::shfl64(v, pipemu::lane() ^ m); }
  static inline int shfl_xor(int v, int m) { return (int)pipemu::shfl64(v, pipemu::lane() ^ m); }
  static inline unsigned atomic_add(unsigned *p, unsigned v) { return pipemu::atomic_add(p, v); }
  /* (called by one lane at a time: the emulated warp is cooperative fibers, nothing runs in between) */
  static inline int atomic_cas(int *p, int expect, int v) { const int o = *p; if (o == expect) *p = v; return o; }
  static inline int atomic_exch(int *p, int v) { const int o = *p; *p = v; return o; }
  static inline int load_volatile(const int *p) { return *p; }
  static inline unsigned load_volatile(const unsigned *p) { return *p; }
  static inline long long load_cg(const long long *p) { return *p; }
  static inline void fence() {}
  static inline void nap() {}
};

/* CTA / grid abstraction of the large-tableau kernel: in emulation one CTA of one warp */
struct G {
  static inline int tid() { return pipemu::lane(); }
  static inline int T() { return 32; }
  static inline int cta() { return 0; }
  static inline int ncta() { return 1; }
  static inline void cta_sync() { pipemu::barrier(); }
  static inline void grid_sync() { pipemu::barrier(); }
  static inline unsigned atomic_add_u(unsigned *p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
  static inline void group_sync(int, int) { pipemu::barrier(); }
  /* bulk (TMA) copy global -> shared signalled on an mbarrier: in emulation a memcpy by the issuing
   * fiber, the wait is a rendezvous of the warp */
  static inline void mbar_init(unsigned long long *, int) {}
  static inline void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *) { __builtin_memcpy(dst, src, bytes); }
  static inline void mbar_wait(unsigned long long *, unsigned) { pipemu::barrier(); }
  static inline void proxy_fence() {}
  static inline void fence() {}
  static inline void relax() {}
  static inline int load_int(const int *p) { return *p; }
};
struct pip_i64x2 { long long x, y; };

static inline int pip_ffs(unsigned m) { return __builtin_ffs((int)m); }
static inline int pip_popc(unsigned m) { return __builtin_popcount(m); }
static inline int pip_clzll(unsigned long long v) { return v ? __builtin_clzll(v) : 64; }
static inline int pip_ctzll(unsigned long long v) { return v ? __builtin_ctzll(v) : 64; }
static inline long long pip_mulhi(long long a, long long b) { return (long long)(((__int128)a * b) >> 64); }
static inline double pip_ll2d(long long v) { return (double)v; }
static inline long long pip_clock() { return 0; }
static inline unsigned pip_f2u(float f) { unsigned u; __builtin_memcpy(&u, &f, 4); return u; }
static inline float pip_u2f(unsigned u) { float f; __builtin_memcpy(&f, &u, 4); return f; }

#elif !defined(__CUDACC__)  /* plain host translation unit (g++): attributes only */

#define PIP_DEV static inline
#define PIP_DEVNI static __attribute__((noinline))
#define PIP_SDEV static inline
#define PIP_SDEVNI static __attribute__((noinline))
#define PIP_HD static inline
#define PIP_HDM inline
#define PIP_DM inline
#define PIP_HDNI static __attribute__((noinline, unused))
static inline int pip_clzll(unsigned long long v) { return v ? __builtin_clzll(v) : 64; }

#else  /* device */

#define PIP_DEV __device__ __forceinline__
#define PIP_DEVNI __device__ __noinline__
#define PIP_SDEV static __device__ __forceinline__
#define PIP_SDEVNI static __device__ __noinline__
#define PIP_HD __host__ __device__ __forceinline__
#define PIP_HDM __host__ __device__ __forceinline__
#define PIP_DM __device__ __forceinline__          /* device-only member function */
#define PIP_HDNI static __host__ __device__ __noinline__
#define PIP_ASSUME_SHARED(p) __builtin_assume(__isShared(p))

struct W {
  static __device__ __forceinline__ int lane() { return (int)(threadIdx.x & 31u); }
  static __device__ __forceinline__ void sync() { __syncwarp(); }
  static __device__ __forceinline__ unsigned ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
  static __device__ __forceinline__ bool any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
  static __device__ __forceinline__ int shfl(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
  static __device__ __forceinline__ long long shfl64(long long v, int src) { return __shfl_sync(0xffffffffu, v, src); }
  static __device__ __forceinline__ int shfl_up(int v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
  static __device__ __forceinline__ unsigned redmin(unsigned v) { return __reduce_min_sync(0xffffffffu, v); }
  static __device__ __forceinline__ unsigned redmax(unsigned v) { return __reduce_max_sync(0xffffffffu, v); }
  static __device__ __forceinline__ unsigned redor(unsigned v) { return __reduce_or_sync(0xffffffffu, v); }
  static __device__ __forceinline__ int shfl_down(int v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
  static __device__ __forceinline__ long long shfl_xor64(long long v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
  static __device__ __forceinline__ int shfl_xor(int v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
  static __device__ __forceinline__ unsigned atomic_add(unsigned *p, unsigned v) { return atomicAdd(p, v); }
  static __device__ __forceinline__ int atomic_cas(int *p, int expect, int v) { return atomicCAS(p, expect, v); }
  static __device__ __forceinline__ int atomic_exch(int *p, int v) { return atomicExch(p, v); }
  static __device__ __forceinline__ int load_volatile(const int *p) { return *(const volatile int *)p; }
  static __device__ __forceinline__ unsigned load_volatile(const unsigned *p) { return *(const volatile unsigned *)p; }
  static __device__ __forceinline__ long long load_cg(const long long *p) { return __ldcg(p); }   /* past the L1: written by another SM */
  static __device__ __forceinline__ void fence() { __threadfence(); }
  static __device__ __forceinline__ void nap() { __nanosleep(200); }
};

struct G {
  static __device__ __forceinline__ int tid() { return (int)threadIdx.x; }
  static __device__ __forceinline__ int T() { return (int)blockDim.x; }
  static __device__ __forceinline__ int cta() { return (int)blockIdx.x; }
  static __device__ __forceinline__ int ncta() { return (int)gridDim.x; }
  static __device__ __forceinline__ void cta_sync() { __syncthreads(); }
  static __device__ void grid_sync();      /* cooperative groups, defined in pip_large.cu */
  static __device__ __forceinline__ unsigned atomic_add_u(unsigned *p, unsigned v) { return atomicAdd(p, v); }
  static __device__ __forceinline__ void group_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
  static __device__ __forceinline__ void fence() { __threadfence(); }
  /* cp.async.bulk (the TMA engine's 1-D copy) global -> shared, completion counted in bytes on an mbarrier */
  static __device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
  }
  static __device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
  {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar), d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(src), "r"(bytes), "r"(b) : "memory");
  }
  static __device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
  {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    unsigned ok;
    do {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok) : "r"(b), "r"(parity) : "memory");
    } while (!ok);
  }
  static __device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
  static __device__ __forceinline__ void relax() { __nanosleep(32); }
  static __device__ __forceinline__ int load_int(const int *p) { return __ldcg(p); }
};
struct __align__(16) pip_i64x2 { long long x, y; };

static __device__ __forceinline__ int pip_ffs(unsigned m) { return __ffs((int)m); }
static __device__ __forceinline__ int pip_popc(unsigned m) { return __popc(m); }
static __host__ __device__ __forceinline__ int pip_clzll(unsigned long long v)
{
#if defined(__CUDA_ARCH__)
  return __clzll((long long)v);
#else
  return v ? __builtin_clzll(v) : 64;
#endif
}
static __device__ __forceinline__ int pip_ctzll(unsigned long long v) { return v ? __ffsll((long long)v) - 1 : 64; }
static __device__ __forceinline__ long long pip_mulhi(long long a, long long b) { return __mul64hi(a, b); }
static __device__ __forceinline__ double pip_ll2d(long long v) { return __ll2double_rn(v); }
static __device__ __forceinline__ long long pip_clock() { return clock64(); }
static __device__ __forceinline__ unsigned pip_f2u(float f) { return __float_as_uint(f); }
static __device__ __forceinline__ float pip_u2f(unsigned u) { return __uint_as_float(u); }

#endif

#endif
