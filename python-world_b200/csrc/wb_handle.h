// Library handle: device id, error text and the small constant tables the kernels
// share (FFT twiddles, Nuttall band windows).  Internal header.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "../../include/world_b200.h"
#include "wb_platform.h"

#define WB_TW_N 16384  // twiddle table length = largest in-kernel FFT size

#ifdef WB_HOST_EMU
inline int wb_dev_alloc(void** p, size_t bytes) {
  *p = std::malloc(bytes ? bytes : 1);
  return *p ? 0 : -1;
}
inline void wb_dev_free(void* p) { std::free(p); }
inline int wb_h2d(void* dst, const void* src, size_t bytes, wb_stream_t) {
  std::memcpy(dst, src, bytes);
  return 0;
}
inline int wb_d2h(void* dst, const void* src, size_t bytes, wb_stream_t) {
  std::memcpy(dst, src, bytes);
  return 0;
}
inline int wb_dev_memset(void* dst, int v, size_t bytes, wb_stream_t) {
  std::memset(dst, v, bytes);
  return 0;
}
inline int wb_stream_sync(wb_stream_t) { return 0; }
inline int wb_d2d(void* dst, const void* src, size_t bytes, wb_stream_t) {
  std::memcpy(dst, src, bytes);
  return 0;
}
#else
inline int wb_d2d(void* dst, const void* src, size_t bytes, wb_stream_t s) {
  return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s) == cudaSuccess ? 0 : -1;
}
inline int wb_dev_alloc(void** p, size_t bytes) { return cudaMalloc(p, bytes ? bytes : 1) == cudaSuccess ? 0 : -1; }
inline void wb_dev_free(void* p) { cudaFree(p); }
inline int wb_h2d(void* dst, const void* src, size_t bytes, wb_stream_t s) {
  return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s) == cudaSuccess ? 0 : -1;
}
inline int wb_d2h(void* dst, const void* src, size_t bytes, wb_stream_t s) {
  return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s) == cudaSuccess ? 0 : -1;
}
inline int wb_dev_memset(void* dst, int v, size_t bytes, wb_stream_t s) {
  return cudaMemsetAsync(dst, v, bytes, s) == cudaSuccess ? 0 : -1;
}
inline int wb_stream_sync(wb_stream_t s) { return cudaStreamSynchronize(s) == cudaSuccess ? 0 : -1; }
#endif

struct wb_handle {
  int device;
  wb_cplx* tw;  // exp(-2 pi i m / WB_TW_N)
  std::map<std::string, void*> tables;  // small device-resident constant tables, keyed by name
  std::string err;
};

inline int wb_fail(wb_handle* h, int code, const char* fmt, ...) __attribute__((format(printf, 3, 4)));
#include <cstdarg>
inline int wb_fail(wb_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  return code;
}

// Fetch (or build and upload) a constant table.  `make` fills a host vector.
template <class T, class F>
inline const T* wb_table(wb_handle* h, const std::string& key, F make) {
  auto it = h->tables.find(key);
  if (it != h->tables.end()) return (const T*)it->second;
  std::vector<T> host;
  make(host);
  void* d = nullptr;
  if (wb_dev_alloc(&d, host.size() * sizeof(T))) return nullptr;
  // synchronous upload (setup path, once per handle and key)
#ifdef WB_HOST_EMU
  std::memcpy(d, host.data(), host.size() * sizeof(T));
#else
  // pageable source: cudaMemcpy may return before the DMA has landed, and the callers' streams (torch side streams
  // are non-blocking) are not ordered against the legacy stream -- wait for the device once per table
  if (cudaMemcpy(d, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaDeviceSynchronize() != cudaSuccess) {
    cudaFree(d);
    return nullptr;
  }
#endif
  h->tables[key] = d;
  return (const T*)d;
}

inline void wb_nuttall(int n, std::vector<double>& w) {
  // 4-term Nuttall, endpoints included (d4c.py:245-249, dio.py:208-212, harvest.py:563-567)
  // Term order and fused multiply-adds follow what the reference's matrix product evaluates to in the
  // pinned environment (checked bit for bit by tests/test_tables.py): DIO takes argmax of an even-length
  // window whose two central samples differ only in the last bit (dio.py:131).
  w.resize(n);
  for (int i = 0; i < n; ++i) {
    const double t = (double)i * 2 * WB_PI / (n - 1);
    double acc = -0.487396 * std::cos(1.0 * t);
    acc = std::fma(0.355768, std::cos(0.0 * t), acc);
    acc = std::fma(0.144232, std::cos(2.0 * t), acc);
    acc = std::fma(-0.012604, std::cos(3.0 * t), acc);
    w[i] = acc;
  }
}

inline int wb_pow2_ceil_log2(double v) {  // 2 ** ceil(log2(v))
  return (int)std::pow(2.0, std::ceil(std::log2(v)));
}
inline int wb_ilog2(int n) {
  int l = 0;
  while ((1 << l) < n) ++l;
  return l;
}
inline bool wb_is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

#ifndef WB_HOST_EMU
// Entry points run on the handle's device and leave the caller's current device as they found it.
struct wb_device_guard {
  int prev = -1;
  bool ok = true;
  explicit wb_device_guard(int want) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != want) ok = cudaSetDevice(want) == cudaSuccess;
    else prev = -1;  // nothing to restore
  }
  ~wb_device_guard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};
#define WB_SET_DEVICE(h)                                                            \
  wb_device_guard wb_guard__((h)->device);                                          \
  if (!wb_guard__.ok) return wb_fail(h, WB_E_CUDA, "cudaSetDevice failed")
#else
#define WB_SET_DEVICE(h) ((void)0)
#endif

#define WB_CHECK_LAUNCH(h, rc, what)                                                       \
  do {                                                                                     \
    int rc__ = (rc);                                                                       \
    if (rc__ != 0) return wb_fail(h, WB_E_CUDA, "%s: launch failed (code %d)", what, rc__); \
  } while (0)
