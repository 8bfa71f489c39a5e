// Device runtime of libnfft4gp_b200: device selection, memory, one library stream, events, launch counter.
// There is deliberately NO CPU fallback: every entry point fails loudly when no CUDA device is usable.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <atomic>

#include "b200_dev.h"
#include "b200_cuda_util.cuh"

static int g_device = -1;
static cudaStream_t g_stream = nullptr;
std::atomic<long long> g_launches{0};

cudaStream_t b200_stream() { return g_stream; }
void b200_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- per-kernel-class event timing (bench.py roofline): events are recorded on the library stream around every
// launch-layer call while profiling is on; elapsed times are summed per class after a synchronize.
#define B200_PROF_MAX 16384
static bool g_prof_on = false;
static int g_prof_n = 0;
static cudaEvent_t g_prof_a[B200_PROF_MAX], g_prof_b[B200_PROF_MAX];
static int g_prof_cls[B200_PROF_MAX];
static int g_prof_created = 0;

void b200_prof_begin(int cls)
{
   if (!g_prof_on || g_prof_n >= B200_PROF_MAX)
      return;
   if (g_prof_n >= g_prof_created)
   {
      cudaEventCreate(&g_prof_a[g_prof_created]);
      cudaEventCreate(&g_prof_b[g_prof_created]);
      g_prof_created++;
   }
   g_prof_cls[g_prof_n] = cls;
   cudaEventRecord(g_prof_a[g_prof_n], g_stream);
}

void b200_prof_end()
{
   if (!g_prof_on || g_prof_n >= B200_PROF_MAX)
      return;
   cudaEventRecord(g_prof_b[g_prof_n], g_stream);
   g_prof_n++;
}

extern "C" {

void b200_prof_start(void)
{
   g_prof_n = 0;
   g_prof_on = true;
}

// sums[c] = total ms of class c, counts[c] = number of timed regions; returns number of regions
int b200_prof_stop(double *sums, long long *counts, int ncls)
{
   g_prof_on = false;
   cudaStreamSynchronize(g_stream);
   for (int c = 0; c < ncls; c++)
   {
      sums[c] = 0.0;
      counts[c] = 0;
   }
   for (int i = 0; i < g_prof_n; i++)
   {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, g_prof_a[i], g_prof_b[i]) == cudaSuccess && g_prof_cls[i] < ncls)
      {
         sums[g_prof_cls[i]] += ms;
         counts[g_prof_cls[i]]++;
      }
   }
   int n = g_prof_n;
   g_prof_n = 0;
   return n;
}

int b200_prof_active(void) { return g_prof_on ? 1 : 0; }

// ---- CUDA graphs: the launch sequence of one matvec (11 kernels + 1 memset for three 3-D windows) is captured once per
// (x, y, alpha, beta) signature and replayed; at small per-GPU sizes (multi-GPU strong scaling) the matvec is otherwise
// bound by the host's launch rate, not by the GPU.
int b200_graph_begin(void)
{
   if (cudaStreamBeginCapture(g_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
   {
      cudaGetLastError();
      return -1;
   }
   return 0;
}

// ends the capture; returns an executable graph or NULL (capture invalidated: the caller falls back to plain launches)
void *b200_graph_end(void)
{
   cudaGraph_t g = nullptr;
   cudaGraphExec_t ex = nullptr;
   if (cudaStreamEndCapture(g_stream, &g) != cudaSuccess || !g)
   {
      cudaGetLastError();
      return nullptr;
   }
   if (cudaGraphInstantiate(&ex, g, 0) != cudaSuccess)
   {
      cudaGetLastError();
      ex = nullptr;
   }
   cudaGraphDestroy(g);
   return (void *)ex;
}

// set while a solver captures a whole iteration: operators that keep their own graphs enqueue their kernels directly
static int g_outer_capture = 0;
void b200_outer_capture_set(int on) { g_outer_capture = on; }
int b200_outer_capturing(void) { return g_outer_capture; }

int b200_graph_launch(void *exec)
{
   B200_CHECK(cudaGraphLaunch((cudaGraphExec_t)exec, g_stream));
   return 0;
}

void b200_graph_destroy(void *exec)
{
   if (exec)
      cudaGraphExecDestroy((cudaGraphExec_t)exec);
}

int b200_device_count(void)
{
   int c = 0;
   if (cudaGetDeviceCount(&c) != cudaSuccess)
   {
      cudaGetLastError();
      return 0;
   }
   return c;
}

int b200_initialized(void) { return g_device >= 0; }

int b200_init(int device)
{
   if (g_device >= 0 && (device < 0 || device == g_device))
      return 0;
   int count = b200_device_count();
   if (count <= 0)
   {
      fprintf(stderr, "nfft4gp_b200: FATAL: no CUDA device visible; this library has no CPU fallback.\n");
      return -1;
   }
   if (device < 0)
   {
      const char *lr = getenv("LOCAL_RANK");
      device = lr ? atoi(lr) % count : 0;
   }
   if (device >= count)
   {
      fprintf(stderr, "nfft4gp_b200: FATAL: device %d requested but only %d visible.\n", device, count);
      return -1;
   }
   B200_CHECK(cudaSetDevice(device));
   cudaDeviceProp prop;
   B200_CHECK(cudaGetDeviceProperties(&prop, device));
   if (prop.major < 10)
   {
      fprintf(stderr, "nfft4gp_b200: FATAL: device %d is sm_%d%d; kernels are built for sm_100a only.\n", device, prop.major, prop.minor);
      return -1;
   }
   if (g_stream)
      cudaStreamDestroy(g_stream);
   B200_CHECK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
   g_device = device;
   return 0;
}

static int ensure_init(void) { return g_device >= 0 ? 0 : b200_init(-1); }

void *b200_malloc(size_t bytes)
{
   void *p = nullptr;
   if (ensure_init())
      return nullptr;
   if (bytes == 0)
      bytes = 8;
   cudaError_t e = cudaMalloc(&p, bytes);
   if (e != cudaSuccess)
   {
      fprintf(stderr, "nfft4gp_b200: cudaMalloc(%zu) failed: %s\n", bytes, cudaGetErrorString(e));
      exit(EXIT_FAILURE); /* reference semantics: allocation failure terminates (SRC/utils/memory.h:85-95) */
   }
   return p;
}

void b200_free(void *p)
{
   if (p)
      cudaFree(p);
}

void *b200_host_alloc(size_t bytes)
{
   void *p = nullptr;
   if (ensure_init())
      return nullptr;
   if (bytes == 0)
      bytes = 8;
   if (cudaMallocHost(&p, bytes) != cudaSuccess)
   {
      fprintf(stderr, "nfft4gp_b200: cudaMallocHost(%zu) failed\n", bytes);
      exit(EXIT_FAILURE);
   }
   return p;
}

void b200_host_free(void *p)
{
   if (p)
      cudaFreeHost(p);
}

int b200_h2d(void *dst, const void *src, size_t bytes)
{
   if (ensure_init())
      return -1;
   B200_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_stream));
   return 0;
}

int b200_d2h(void *dst, const void *src, size_t bytes)
{
   if (ensure_init())
      return -1;
   B200_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_stream));
   B200_CHECK(cudaStreamSynchronize(g_stream));
   return 0;
}

// enqueue only (the destination must be pinned host memory); used inside captured iteration graphs
int b200_d2h_async(void *dst, const void *src, size_t bytes)
{
   if (ensure_init())
      return -1;
   B200_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_stream));
   return 0;
}

int b200_d2d(void *dst, const void *src, size_t bytes)
{
   if (ensure_init())
      return -1;
   B200_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, g_stream));
   return 0;
}

int b200_memset(void *dst, int v, size_t bytes)
{
   if (ensure_init())
      return -1;
   B200_CHECK(cudaMemsetAsync(dst, v, bytes, g_stream));
   return 0;
}

int b200_sync(void)
{
   if (ensure_init())
      return -1;
   B200_CHECK(cudaStreamSynchronize(g_stream));
   return 0;
}

void *b200_event_create(void)
{
   cudaEvent_t ev;
   if (ensure_init())
      return nullptr;
   if (cudaEventCreate(&ev) != cudaSuccess)
      return nullptr;
   return (void *)ev;
}

void b200_event_destroy(void *ev)
{
   if (ev)
      cudaEventDestroy((cudaEvent_t)ev);
}

int b200_event_record(void *ev)
{
   B200_CHECK(cudaEventRecord((cudaEvent_t)ev, g_stream));
   return 0;
}

float b200_event_elapsed_ms(void *a, void *b)
{
   float ms = -1.0f;
   if (cudaEventSynchronize((cudaEvent_t)b) != cudaSuccess)
      return -1.0f;
   if (cudaEventElapsedTime(&ms, (cudaEvent_t)a, (cudaEvent_t)b) != cudaSuccess)
      return -1.0f;
   return ms;
}

long long b200_launch_count(void) { return g_launches.load(); }
void b200_launch_count_reset(void) { g_launches.store(0); }
void b200_launch_count_add(long long k) { g_launches.fetch_add(k, std::memory_order_relaxed); }
void *b200_stream_handle(void) { return (void *)g_stream; }


/* In-run measurement of the roof the 3-D spread / gather kernels are bound by: mma.sync.m8n8k4.f64 (DMMA.8x8x4) throughput of
 * this GPU at its clocks of the moment -- 16 independent accumulator chains per warp, 16 warps per SM, operands in registers,
 * no memory traffic (scripts/fp64_dmma.cu as a library entry, so that bench.py reports the fraction against a number measured
 * in the same process instead of a constant).  Returns TFLOP/s (2 flop per FMA) or a negative value on error. */
double b200_fp64_tensor_peak_tflops(void);
} // extern "C"

__global__ void __launch_bounds__(256) k_fp64_tensor_peak(double *out, int iters, double a, double b)
{
   double m[16][2];
#pragma unroll
   for (int i = 0; i < 16; i++)
   {
      m[i][0] = i + threadIdx.x;
      m[i][1] = -i;
   }
   for (int it = 0; it < iters; it++)
   {
#pragma unroll
      for (int i = 0; i < 16; i++)
         asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(m[i][0]), "+d"(m[i][1]) : "d"(a), "d"(b));
   }
   double s = 0.0;
#pragma unroll
   for (int i = 0; i < 16; i++)
      s += m[i][0] + m[i][1];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" {
double b200_fp64_tensor_peak_tflops(void)
{
   int sm = 148, dev = 0;
   const int iters = 4000, warps_per_sm = 16;
   double *out = nullptr, best = -1.0;
   cudaEvent_t e0, e1;
   if (cudaGetDevice(&dev) != cudaSuccess)
      return -1.0;
   cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
   const int blocks = sm * warps_per_sm / 8;
   if (cudaMalloc(&out, sizeof(double) * (size_t)blocks * 256) != cudaSuccess)
      return -1.0;
   cudaEventCreate(&e0);
   cudaEventCreate(&e1);
   k_fp64_tensor_peak<<<blocks, 256, 0, b200_stream()>>>(out, 200, 1.0000001, 1e-9); /* warm-up */
   for (int rep = 0; rep < 3; rep++)
   {
      float ms = 0.f;
      cudaEventRecord(e0, b200_stream());
      k_fp64_tensor_peak<<<blocks, 256, 0, b200_stream()>>>(out, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1, b200_stream());
      if (cudaEventSynchronize(e1) != cudaSuccess || cudaEventElapsedTime(&ms, e0, e1) != cudaSuccess || ms <= 0.f)
      {
         best = -1.0;
         break;
      }
      const double tf = (double)blocks * 8.0 * 16.0 * iters * 512.0 / ((double)ms * 1e9); /* 8 x 8 x 4 x 2 flop per warp-level mma */
      if (tf > best)
         best = tf;
   }
   cudaEventDestroy(e0);
   cudaEventDestroy(e1);
   cudaFree(out);
   return best;
}
} // extern "C"
