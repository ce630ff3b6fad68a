/*
 *  cuda_runtime.h (tests/emu) -- TEST INFRASTRUCTURE, never part of the product.
 *
 *  A stand-in for the CUDA headers that lets g++ compile the .cu files of fiasco_b200/csrc as plain
 *  C++ and run one thread block at a time on the CPU: every CUDA thread is a fibre on one
 *  OS thread, __syncthreads() and the warp collectives are rendezvous points of the fibre
 *  scheduler (tests/emu/emu_runtime.cpp).  It exists so that the control flow of the device
 *  code -- the same source file the product compiles with nvcc for sm_100a -- can be
 *  exercised against the oracle in the CPU-only test suite, where no GPU is present.  Only
 *  tests/ builds or loads the resulting libfiasco_b200_emu.so; the product library
 *  (fiasco_b200/lib/libfiasco_b200.so) has no CPU path and fails with FB200_ENODEVICE without
 *  a device.  Nothing measured or shipped goes through this file.
 */
#ifndef FB200_EMU_CUDA_RUNTIME_H
#define FB200_EMU_CUDA_RUNTIME_H

#ifndef FB200_EMU
#error "tests/emu/cuda_runtime.h is only for the -DFB200_EMU test build"
#endif

#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <functional>

/* ---- qualifiers ---- */
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__ ((noinline))
#define __shared__ static	/* one block runs at a time */
#define __constant__ static
#define __align__(n) __attribute__ ((aligned (n)))
#define __launch_bounds__(...)

/* ---- built-in variables and vector types ---- */
struct emu_dim3
{
   unsigned x, y, z;
   emu_dim3 (unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x (x_), y (y_), z (z_) {}
};
typedef emu_dim3 dim3;
extern emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

struct __attribute__ ((aligned (16))) float4 { float x, y, z, w; };
struct __attribute__ ((aligned (16))) uint4  { unsigned x, y, z, w; };

/* ---- scheduler entry points (emu_runtime.cpp) ---- */
void	 emu_launch (emu_dim3 grid, emu_dim3 block, size_t smem, const std::function<void ()> &body);
void	 emu_launch_cluster (emu_dim3 grid, emu_dim3 block, unsigned cluster, size_t smem,
			     const std::function<void ()> &body);
void	 emu_syncthreads (void);
void	 emu_cluster_sync (void);		/* barrier.cluster arrive + wait */
unsigned emu_cluster_rank (void);
unsigned emu_cluster_size (void);
void	*emu_map_shared_rank (const void *p, unsigned rank);	/* mapa */
unsigned emu_warp_exchange (unsigned value, int kind, int arg);	/* 0 shfl, 1 shfl_up, 2 ballot, 3 sync, 4 shfl_xor */
unsigned char *emu_dyn_smem (void);

static inline void __syncthreads (void) { emu_syncthreads (); }
static inline void __syncwarp (unsigned = 0xffffffffu) { emu_warp_exchange (0, 3, 0); }
static inline void __threadfence (void) {}

template <typename T> static inline T
__shfl_sync (unsigned, T v, int src)
{
   static_assert (sizeof (T) == 4, "32-bit shuffles only");
   unsigned u;
   memcpy (&u, &v, 4);
   u = emu_warp_exchange (u, 0, src);
   memcpy (&v, &u, 4);
   return v;
}
template <typename T> static inline T
__shfl_up_sync (unsigned, T v, unsigned delta)
{
   static_assert (sizeof (T) == 4, "32-bit shuffles only");
   unsigned u;
   memcpy (&u, &v, 4);
   u = emu_warp_exchange (u, 1, (int) delta);
   memcpy (&v, &u, 4);
   return v;
}
template <typename T> static inline T
__shfl_xor_sync (unsigned, T v, int mask)
{
   static_assert (sizeof (T) == 4, "32-bit shuffles only");
   unsigned u;
   memcpy (&u, &v, 4);
   u = emu_warp_exchange (u, 4, mask);
   memcpy (&v, &u, 4);
   return v;
}
static inline unsigned __ballot_sync (unsigned, int pred) { return emu_warp_exchange (pred != 0, 2, 0); }

/* ---- arithmetic intrinsics ---- */
static inline int min (int a, int b) { return a < b ? a : b; }
static inline int max (int a, int b) { return a > b ? a : b; }
static inline unsigned __float_as_uint (float f) { unsigned u; memcpy (&u, &f, 4); return u; }
static inline float    __uint_as_float (unsigned u) { float f; memcpy (&f, &u, 4); return f; }
static inline float    __int_as_float (int i) { float f; memcpy (&f, &i, 4); return f; }
static inline int      __float_as_int (float f) { int i; memcpy (&i, &f, 4); return i; }
static inline int      __clz (int x) { return x ? __builtin_clz ((unsigned) x) : 32; }
static inline int      __popc (unsigned x) { return __builtin_popcount (x); }
static inline int      __ffs (int x) { return __builtin_ffs (x); }
static inline unsigned
__fns (unsigned mask, unsigned base, int offset)	/* offset-th set bit at or above base */
{
   for (unsigned b = base; b < 32; b++)
      if ((mask >> b) & 1u)
	 if (--offset <= 0)
	    return b;
   return 0xffffffffu;
}
static inline bool	   __isGlobal (const void *) { return true; }
#define __builtin_assume(x) ((void) 0)
static inline size_t	   __cvta_generic_to_shared (const void *p) { return (size_t) p; }
long long clock64 (void);

static inline int atomicCAS (int *p, int cmp, int val) { int old = *p; if (old == cmp) *p = val; return old; }
static inline int atomicExch (int *p, int val) { int old = *p; *p = val; return old; }

/* ---- the slice of the runtime API the host side of csrc/ uses ---- */
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorNoDevice = 100, cudaErrorInsufficientDriver = 35 };
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };
enum { cudaStreamNonBlocking = 1 };
enum { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };

static inline const char *cudaGetErrorName (cudaError_t) { return "emu"; }
static inline const char *cudaGetErrorString (cudaError_t) { return "emulated runtime"; }
static inline cudaError_t cudaGetLastError (void) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount (int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice (int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice (int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t
cudaDeviceGetAttribute (int *v, int attr, int)
{
   /* few "SMs": the launcher then picks the 128-thread shape for batches, and a workspace
      per tile */
   *v = attr == cudaDevAttrMultiProcessorCount ? 2 : 227 * 1024;
   return cudaSuccess;
}
template <typename T> static inline cudaError_t
cudaMalloc (T **p, size_t n) { *p = (T *) calloc (n ? n : 1, 1); return *p ? cudaSuccess : 2; }
template <typename T> static inline cudaError_t
cudaMallocHost (T **p, size_t n) { *p = (T *) calloc (n ? n : 1, 1); return *p ? cudaSuccess : 2; }
static inline cudaError_t cudaFree (void *p) { free (p); return cudaSuccess; }
static inline cudaError_t cudaFreeHost (void *p) { free (p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy (void *d, const void *s, size_t n, int) { memcpy (d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync (void *d, const void *s, size_t n, int, cudaStream_t = 0) { memcpy (d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset (void *d, int v, size_t n) { memset (d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync (void *d, int v, size_t n, cudaStream_t = 0) { memset (d, v, n); return cudaSuccess; }
template <typename T, size_t N> static inline cudaError_t
cudaMemcpyToSymbol (T (&sym) [N], const void *s, size_t n) { memcpy (sym, s, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags (cudaStream_t *s, int) { *s = (void *) 1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy (cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize (cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize (void) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate (cudaEvent_t *e) { *e = (void *) 1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy (cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord (cudaEvent_t, cudaStream_t = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize (cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime (float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0; return cudaSuccess; }
template <typename K> static inline cudaError_t cudaFuncSetAttribute (K, int, int) { return cudaSuccess; }
template <typename K> static inline cudaError_t
cudaOccupancyMaxActiveBlocksPerMultiprocessor (int *n, K, int, size_t) { *n = 1 << 20; return cudaSuccess; }

#endif
