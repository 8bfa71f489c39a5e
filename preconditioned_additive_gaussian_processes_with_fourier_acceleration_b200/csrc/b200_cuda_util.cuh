// Shared helpers of the .cu translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>

cudaStream_t b200_stream();
void b200_count_launch();
void b200_prof_begin(int cls);
void b200_prof_end();
// kernel classes of the profiling facility (keep in sync with bench.py)
enum { PC_PREP = 0, PC_SPREAD = 1, PC_FFT_FWD = 2, PC_FFT_INV = 3, PC_GATHER = 4, PC_COMBINE = 5, PC_VEC = 6, PC_NYS = 7, PC_CSR = 8, PC_GEMV = 9, PC_OTHER = 10, PC_COUNT = 11 };
struct B200ProfScope { B200ProfScope(int c) { b200_prof_begin(c); } ~B200ProfScope() { b200_prof_end(); } };

#define B200_CHECK(call)                                                                                  \
   do                                                                                                     \
   {                                                                                                      \
      cudaError_t e__ = (call);                                                                           \
      if (e__ != cudaSuccess)                                                                             \
      {                                                                                                   \
         fprintf(stderr, "nfft4gp_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); \
         return -1;                                                                                       \
      }                                                                                                   \
   } while (0)

// call after every kernel launch: counts it and surfaces launch-configuration errors
#define B200_LAUNCHED()                                                                                   \
   do                                                                                                     \
   {                                                                                                      \
      b200_count_launch();                                                                                \
      cudaError_t e__ = cudaGetLastError();                                                               \
      if (e__ != cudaSuccess)                                                                             \
      {                                                                                                   \
         fprintf(stderr, "nfft4gp_b200: kernel launch failed: %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); \
         return -1;                                                                                       \
      }                                                                                                   \
   } while (0)

static inline int b200_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ double b200_warp_sum(double v)
{
#pragma unroll
   for (int o = 16; o > 0; o >>= 1)
      v += __shfl_xor_sync(0xffffffffu, v, o);
   return v;
}
