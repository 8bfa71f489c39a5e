/* Part 2 of the C ABI (B200 extensions): device selection, memory helpers, events, counters. */
#include <stdio.h>
#include <stdlib.h>

#include "nfft4gp_b200.h"
#include "b200_dev.h"
#include "b200_host.h"

int Nfft4GPB200Init(int device) { return b200_init(device); }
int Nfft4GPB200DeviceCount(void) { return b200_device_count(); }
const char *Nfft4GPB200Version(void) { return "nfft4gp_b200 0.1 (sm_100a; N=32, n_os=64, m=4 Kaiser-Bessel)"; }
long long Nfft4GPB200LaunchCount(void) { return b200_launch_count(); }
void Nfft4GPB200LaunchCountReset(void) { b200_launch_count_reset(); }

void *Nfft4GPB200HostAlloc(size_t bytes) { return b200_host_alloc(bytes); }
void Nfft4GPB200HostFree(void *p) { b200_host_free(p); }
void *Nfft4GPB200DeviceAlloc(size_t bytes) { return b200_malloc(bytes); }
void Nfft4GPB200DeviceFree(void *p) { b200_free(p); }
int Nfft4GPB200MemcpyH2D(void *dst_dev, const void *src_host, size_t bytes) { return b200_h2d(dst_dev, src_host, bytes) || b200_sync(); }
int Nfft4GPB200MemcpyD2H(void *dst_host, const void *src_dev, size_t bytes) { return b200_d2h(dst_host, src_dev, bytes); }
int Nfft4GPB200DeviceSynchronize(void) { return b200_sync(); }

void *Nfft4GPB200EventCreate(void) { return b200_event_create(); }
void Nfft4GPB200EventDestroy(void *ev) { b200_event_destroy(ev); }
int Nfft4GPB200EventRecord(void *ev) { return b200_event_record(ev); }
float Nfft4GPB200EventElapsedMs(void *start, void *stop) { return b200_event_elapsed_ms(start, stop); }

/* per-kernel-class CUDA-event timing on the library stream (bench.py roofline) */
void Nfft4GPB200ProfileStart(void) { b200_prof_start(); }
int Nfft4GPB200ProfileStop(double *sums_ms, long long *counts, int ncls) { return b200_prof_stop(sums_ms, counts, ncls); }
