// rt.h — tiny runtime shim: CUDA runtime for the product build, libc for the -DRCZ_EMU test build.
#pragma once
#include "simt.h"
#include <stdlib.h>
#include <string.h>

#ifdef RCZ_EMU
typedef void* rt_stream_t;
struct rt_event_emu { int dummy; };
typedef rt_event_emu* rt_event_t;
inline int rt_set_device(int) { return 0; }
inline int rt_l2_fetch_granularity(int) { return 0; }
inline int rt_device_count(int* n) { *n = 1; return 0; }
inline int rt_sm_count(int, int* n) { *n = 4; return 0; }
inline int rt_malloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? 0 : 2; }
inline int rt_free(void* p) { free(p); return 0; }
inline int rt_host_alloc(void** p, size_t n) { return rt_malloc(p, n); }
inline int rt_host_free(void* p) { free(p); return 0; }
inline int rt_h2d(void* d, const void* s, size_t n, rt_stream_t) { if (n) memcpy(d, s, n); return 0; }
inline int rt_d2h(void* d, const void* s, size_t n, rt_stream_t) { if (n) memcpy(d, s, n); return 0; }
inline int rt_d2d(void* d, const void* s, size_t n, rt_stream_t) { if (n) memmove(d, s, n); return 0; }
inline int rt_memset(void* d, int v, size_t n, rt_stream_t) { if (n) memset(d, v, n); return 0; }
inline int rt_stream_create(rt_stream_t* s) { *s = nullptr; return 0; }
inline int rt_stream_destroy(rt_stream_t) { return 0; }
inline int rt_stream_sync(rt_stream_t) { return 0; }
inline int rt_last_error() { return 0; }
inline const char* rt_error_string(int) { return "emu"; }
inline int rt_event_create(rt_event_t* e) { *e = nullptr; return 0; }
inline int rt_event_destroy(rt_event_t) { return 0; }
inline int rt_event_record(rt_event_t, rt_stream_t) { return 0; }
inline int rt_event_elapsed(float* ms, rt_event_t, rt_event_t) { *ms = 0.f; return 0; }
inline int rt_stream_wait_event(rt_stream_t, rt_event_t) { return 0; }
inline int rt_event_sync(rt_event_t) { return 0; }
#else
typedef cudaStream_t rt_stream_t;
typedef cudaEvent_t rt_event_t;
inline int rt_set_device(int d) { return (int)cudaSetDevice(d); }
// performance hint: DRAM -> L2 fill size for a sector miss (random 4-byte gathers want 32 B, not 64/128 B)
inline int rt_l2_fetch_granularity(int bytes) { return (int)cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes); }
inline int rt_device_count(int* n) { return (int)cudaGetDeviceCount(n); }
inline int rt_sm_count(int d, int* n) { return (int)cudaDeviceGetAttribute(n, cudaDevAttrMultiProcessorCount, d); }
inline int rt_malloc(void** p, size_t n) { return (int)cudaMalloc(p, n ? n : 256); }
inline int rt_free(void* p) { return (int)cudaFree(p); }
inline int rt_host_alloc(void** p, size_t n) { return (int)cudaMallocHost(p, n ? n : 256); }
inline int rt_host_free(void* p) { return (int)cudaFreeHost(p); }
inline int rt_h2d(void* d, const void* s, size_t n, rt_stream_t st) { return n ? (int)cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, st) : 0; }
inline int rt_d2h(void* d, const void* s, size_t n, rt_stream_t st) { return n ? (int)cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, st) : 0; }
inline int rt_d2d(void* d, const void* s, size_t n, rt_stream_t st) { return n ? (int)cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, st) : 0; }
inline int rt_memset(void* d, int v, size_t n, rt_stream_t st) { return n ? (int)cudaMemsetAsync(d, v, n, st) : 0; }
inline int rt_stream_create(rt_stream_t* s) { return (int)cudaStreamCreateWithFlags(s, cudaStreamNonBlocking); }
inline int rt_stream_destroy(rt_stream_t s) { return (int)cudaStreamDestroy(s); }
inline int rt_stream_sync(rt_stream_t s) { return (int)cudaStreamSynchronize(s); }
inline int rt_last_error() { return (int)cudaGetLastError(); }
inline const char* rt_error_string(int e) { return cudaGetErrorString((cudaError_t)e); }
inline int rt_event_create(rt_event_t* e) { return (int)cudaEventCreate(e); }
inline int rt_event_destroy(rt_event_t e) { return (int)cudaEventDestroy(e); }
inline int rt_event_record(rt_event_t e, rt_stream_t s) { return (int)cudaEventRecord(e, s); }
inline int rt_event_elapsed(float* ms, rt_event_t a, rt_event_t b) { return (int)cudaEventElapsedTime(ms, a, b); }
inline int rt_stream_wait_event(rt_stream_t s, rt_event_t e) { return (int)cudaStreamWaitEvent(s, e, 0); }
inline int rt_event_sync(rt_event_t e) { return (int)cudaEventSynchronize(e); }
#endif
