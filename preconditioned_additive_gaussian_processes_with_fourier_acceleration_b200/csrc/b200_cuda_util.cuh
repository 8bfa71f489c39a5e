// Shared helpers of the .cu translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>

cudaStream_t b200_stream();
void b200_count_launch();

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
