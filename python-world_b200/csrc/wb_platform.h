// world_b200 -- platform layer shared by every kernel.
//
// Kernels are written as "block bodies": functors with
//     WB_DEV void operator()(int block, int tid, int nthr, double* smem) const
// in which every thread-parallel loop has the form  for (i = tid; i < n; i += nthr)
// and phases are separated by WB_SYNC().  nvcc compiles them into __global__
// kernels for sm_100a (the product).  With -DWB_HOST_EMU the same bodies compile
// with g++ and run with nthr == 1 (barriers become no-ops): a TEST-ONLY build used
// by tests/ to check kernel logic against the oracle on a box without a GPU.  The
// product library (libworld_b200.so) never contains or falls back to that path.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#ifdef WB_HOST_EMU
inline void wb_host_sincos(double x, double* s, double* c) {
  *s = std::sin(x);
  *c = std::cos(x);
}
#define sincos wb_host_sincos
#define WB_DEV inline
#define WB_DEV_NI inline
#define WB_DEV_COLD inline
#define WB_HD inline
#define WB_SYNC() ((void)0)
#define WB_LDG(p) (*(p))
typedef void* wb_stream_t;
#else
#include <cuda_runtime.h>
#define WB_DEV __device__ __forceinline__
// large block-cooperative helpers (FFT, windows, running integrals): measured faster inlined at every call site
// (constant folding of n / dir) than as real functions, despite the larger instruction footprint
#define WB_DEV_NI __device__ __forceinline__
// rarely taken fallbacks: real functions, so that the hot path stays small in the instruction cache
#define WB_DEV_COLD __device__ __noinline__
#define WB_HD __host__ __device__ __forceinline__
#define WB_SYNC() __syncthreads()
#define WB_LDG(p) __ldg(p)
typedef cudaStream_t wb_stream_t;
#endif

#define WB_PI 3.14159265358979323846
#define WB_EPS 2.220446049250313e-16

struct alignas(16) wb_cplx {
  double x, y;
};

WB_HD wb_cplx wb_mk(double x, double y) {
  wb_cplx c;
  c.x = x;
  c.y = y;
  return c;
}
WB_HD wb_cplx wb_ldg_cplx(const wb_cplx* p) {
#if defined(WB_HOST_EMU) || !defined(__CUDA_ARCH__)
  return *p;
#else
  const double2 v = __ldg(reinterpret_cast<const double2*>(p));
  return wb_mk(v.x, v.y);
#endif
}
WB_HD wb_cplx wb_cmul(wb_cplx a, wb_cplx b) { return wb_mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
WB_HD wb_cplx wb_cadd(wb_cplx a, wb_cplx b) { return wb_mk(a.x + b.x, a.y + b.y); }
WB_HD wb_cplx wb_csub(wb_cplx a, wb_cplx b) { return wb_mk(a.x - b.x, a.y - b.y); }
WB_HD wb_cplx wb_conj(wb_cplx a) { return wb_mk(a.x, -a.y); }

// sin(pi x), cos(pi x) with exact argument reduction
WB_HD void wb_sincospi(double x, double* s, double* c) {
#ifdef WB_HOST_EMU
  double r = std::fmod(x, 2.0);
  *s = std::sin(WB_PI * r);
  *c = std::cos(WB_PI * r);
#else
  sincospi(x, s, c);
#endif
}

// a / 1000.0, correctly rounded, without the ~25-instruction division: q = a * r with r = RN(1/1000), one
// FMA for the exact remainder and one for the correction (Markstein).  Checked exhaustively against the division for
// a = j * period, j < 5e7, periods 0.5 .. 10 ms (frame times are formed as j * period / 1000 throughout the reference).
WB_HD double wb_div1000(double a) {
  const double r = 1.0 / 1000.0;
  const double q = a * r;
  return fma(fma(-q, 1000.0, a), r, q);
}

WB_HD int wb_imin(int a, int b) { return a < b ? a : b; }
WB_HD int wb_imax(int a, int b) { return a > b ? a : b; }
WB_HD double wb_dmin(double a, double b) { return a < b ? a : b; }
WB_HD double wb_dmax(double a, double b) { return a > b ? a : b; }

// ---------------------------------------------------------------------------------
// Block-wide sum / max of one double per thread.  `scratch` needs 33 doubles.
// ---------------------------------------------------------------------------------
#ifdef WB_HOST_EMU
WB_DEV double wb_block_sum(double v, double*, int, int) { return v; }
WB_DEV double wb_block_max(double v, double*, int, int) { return v; }
WB_DEV void wb_block_sum3(double& a, double& b, double& c, double*, int, int) {}
WB_DEV void wb_atomic_add(double* p, double v) { *p += v; }
WB_DEV int wb_atomic_add_int(int* p, int v) {
  int o = *p;
  *p += v;
  return o;
}
#else
WB_DEV double wb_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
WB_DEV double wb_warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// second stage of the block reductions: the per-warp partials (nw <= 32 of them) are combined by a shuffle tree
// in every warp (one shared-memory load per lane instead of nw dependent loads per thread); the tree is the same
// in every warp, so all threads get bit-identical results
WB_DEV double wb_partials_sum(const double* scratch, int lane, int nw) {
  if (nw <= 8) {  // the steps over lanes 8..31 of the tree below would add zeros: three steps over 8 lanes, same bits
    double t = (lane & 7) < nw ? scratch[lane & 7] : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
  }
  double t = lane < nw ? scratch[lane] : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}
WB_DEV double wb_block_sum(double v, double* scratch, int tid, int nthr) {
  v = wb_warp_sum(v);
  const int w = tid >> 5, nw = (nthr + 31) >> 5;
  __syncthreads();
  if ((tid & 31) == 0) scratch[w] = v;
  __syncthreads();
  return wb_partials_sum(scratch, tid & 31, nw);
}
WB_DEV double wb_block_max(double v, double* scratch, int tid, int nthr) {
  v = wb_warp_max(v);
  const int w = tid >> 5, nw = (nthr + 31) >> 5;
  __syncthreads();
  if ((tid & 31) == 0) scratch[w] = v;
  __syncthreads();
  double t = scratch[(tid & 31) < nw ? (tid & 31) : 0];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t = fmax(t, __shfl_xor_sync(0xffffffffu, t, o));
  return t;
}
WB_DEV void wb_block_sum3(double& a, double& b, double& c, double* scratch, int tid, int nthr) {
  a = wb_warp_sum(a);
  b = wb_warp_sum(b);
  c = wb_warp_sum(c);
  const int w = tid >> 5, nw = (nthr + 31) >> 5;
  __syncthreads();
  if ((tid & 31) == 0) {
    scratch[w] = a;
    scratch[32 + w] = b;
    scratch[64 + w] = c;
  }
  __syncthreads();
  a = wb_partials_sum(scratch, tid & 31, nw);
  b = wb_partials_sum(scratch + 32, tid & 31, nw);
  c = wb_partials_sum(scratch + 64, tid & 31, nw);
}
WB_DEV void wb_atomic_add(double* p, double v) { atomicAdd(p, v); }
WB_DEV int wb_atomic_add_int(int* p, int v) { return atomicAdd(p, v); }
#endif
#define WB_REDUCE_SCRATCH 96  // doubles

// Lane-level helpers for "one warp per item" code.  In the host emulation a warp
// is a single lane.
#ifdef WB_HOST_EMU
#define WB_LANES 1
WB_DEV double wb_lanes_sum(double v) { return v; }
WB_DEV void wb_lanes_sync() {}
// pick the lane value with the smallest key; ties -> larger tag.  Returns through refs.
WB_DEV void wb_lanes_argmin(double& key, int& tag, double& payload) {}
WB_DEV void wb_lanes_argmax_first(double& key, int& tag, double& payload) {}
WB_DEV double wb_lanes_max(double v) { return v; }
WB_DEV int wb_lanes_bcast_int(int v) { return v; }
#else
#define WB_LANES 32
WB_DEV double wb_lanes_sum(double v) { return wb_warp_sum(v); }
WB_DEV void wb_lanes_sync() { __syncwarp(); }
WB_DEV void wb_lanes_argmin(double& key, int& tag, double& payload) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double k2 = __shfl_xor_sync(0xffffffffu, key, o);
    const int t2 = __shfl_xor_sync(0xffffffffu, tag, o);
    const double p2 = __shfl_xor_sync(0xffffffffu, payload, o);
    if (k2 < key || (k2 == key && t2 > tag)) {
      key = k2;
      tag = t2;
      payload = p2;
    }
  }
}
// largest key; ties -> smaller tag
WB_DEV void wb_lanes_argmax_first(double& key, int& tag, double& payload) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double k2 = __shfl_xor_sync(0xffffffffu, key, o);
    const int t2 = __shfl_xor_sync(0xffffffffu, tag, o);
    const double p2 = __shfl_xor_sync(0xffffffffu, payload, o);
    if (k2 > key || (k2 == key && t2 < tag)) {
      key = k2;
      tag = t2;
      payload = p2;
    }
  }
}
WB_DEV double wb_lanes_max(double v) { return wb_warp_max(v); }
WB_DEV int wb_lanes_bcast_int(int v) { return __shfl_sync(0xffffffffu, v, 0); }
#endif

// ---------------------------------------------------------------------------------
// Bulk asynchronous copy global -> shared (TMA engine, `cp.async.bulk`) completing on an mbarrier: one thread
// arms the barrier with the byte count and issues the copy, the data lands without passing through registers
// while the block computes, and every thread waits on the barrier's phase parity before reading.  Sizes and both
// addresses must be multiples of 16 bytes.  GPU only (the host emulation reads global memory directly).
// ---------------------------------------------------------------------------------
#ifndef WB_HOST_EMU
WB_DEV unsigned wb_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
WB_DEV void wb_mbar_init(unsigned long long* bar, int arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(wb_smem_addr(bar)), "r"(arrivals) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // visible to the async proxy
}
WB_DEV void wb_bulk_load(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  const unsigned b = wb_smem_addr(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   wb_smem_addr(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(b)
               : "memory");
}
// orders this thread's earlier generic-proxy accesses to shared memory (made visible to it by a barrier) before
// the async-proxy writes of a bulk copy issued afterwards into the same buffer
WB_DEV void wb_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
WB_DEV void wb_mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned b = wb_smem_addr(bar);
  unsigned done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(b), "r"(parity)
        : "memory");
  } while (!done);
}
#endif

// ---------------------------------------------------------------------------------
// In-place inclusive prefix sum of s[0..n) in shared memory.  `carry` needs nthr+1
// doubles.  Each thread scans one contiguous chunk, chunk totals are scanned by
// thread 0 (nthr <= 1024, this is a few hundred adds), then offsets are applied.
// ---------------------------------------------------------------------------------
WB_DEV void wb_block_scan(double* s, int n, double* carry, int tid, int nthr) {
  const int chunk = (n + nthr - 1) / nthr;
  const int lo = wb_imin(n, tid * chunk), hi = wb_imin(n, lo + chunk);
  double run = 0.0;
  for (int i = lo; i < hi; ++i) {
    run += s[i];
    s[i] = run;
  }
#ifdef WB_HOST_EMU
  (void)carry;
  WB_SYNC();
#else
  // exclusive prefix of the per-thread totals: shuffle scan inside each warp, then across warps
  const int lane = tid & 31, w = tid >> 5, nw = (nthr + 31) >> 5;
  double v = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  if (lane == 31) carry[w] = v;
  __syncthreads();
  if (w == 0) {
    double c = lane < nw ? carry[lane] : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, c, o);
      if (lane >= o) c += t;
    }
    if (lane < nw) carry[lane] = c;  // inclusive totals of warps 0..lane
  }
  __syncthreads();
  const double off = (v - run) + (w > 0 ? carry[w - 1] : 0.0);
  if (off != 0.0)
    for (int i = lo; i < hi; ++i) s[i] += off;
  __syncthreads();
#endif
}

// ---------------------------------------------------------------------------------
// Launch of a block body, and of a barrier-free per-item body
//     WB_DEV void operator()(long long item) const
// ---------------------------------------------------------------------------------
#ifdef WB_HOST_EMU
template <class Body>
inline int wb_launch_flat(const Body& body, long long items, int /*block*/, wb_stream_t) {
  for (long long i = 0; i < items; ++i) body(i);
  return 0;
}
#else
template <class Body>
__global__ void wb_kernel_flat(const Body body, long long items) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < items) body(i);
}
template <class Body>
inline int wb_launch_flat(const Body& body, long long items, int block, wb_stream_t stream) {
  if (items <= 0) return 0;
  const long long grid = (items + block - 1) / block;
  wb_kernel_flat<Body><<<(unsigned)grid, block, 0, stream>>>(body, items);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e - 1000;
}
#endif
#ifdef WB_HOST_EMU
template <class Body>
inline int wb_launch(const Body& body, long long grid, int /*block*/, size_t smem_bytes, wb_stream_t) {
  double* smem = (double*)std::malloc(smem_bytes + 64);
  if (!smem) return -1;
  for (long long b = 0; b < grid; ++b) body((int)b, 0, 1, smem);
  std::free(smem);
  return 0;
}
template <class Body>
inline int wb_launch_spectral(const Body& body, long long grid, int block, size_t smem_bytes, wb_stream_t st) {
  return wb_launch(body, grid, block, smem_bytes, st);
}
template <class Body>
inline int wb_launch_spectral3(const Body& body, long long grid, int block, size_t smem_bytes, wb_stream_t st) {
  return wb_launch(body, grid, block, smem_bytes, st);
}
template <class Body, int MAXT, int MINB>
inline int wb_launch_b(const Body& body, long long grid, int block, size_t smem_bytes, wb_stream_t st) {
  return wb_launch(body, grid, block, smem_bytes, st);
}
#else
// MAXT / MINB: __launch_bounds__ (register cap = 65536 / (MAXT * MINB)); the block size must be <= MAXT
template <class Body, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) wb_kernel(const Body body) {
  extern __shared__ double wb_smem[];
  body((int)blockIdx.x, (int)threadIdx.x, (int)blockDim.x, wb_smem);
}
template <class Body, int MAXT, int MINB>
inline int wb_launch_b(const Body& body, long long grid, int block, size_t smem_bytes, wb_stream_t stream) {
  if (grid <= 0) return 0;
  if (block > MAXT) return -2000;
  if (smem_bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(wb_kernel<Body, MAXT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem_bytes);
    if (e != cudaSuccess) return -(int)e - 1000;
  }
  if (smem_bytes > 8 * 1024)  // these kernels live in shared memory: ask for the largest carve-out
    cudaFuncSetAttribute(wb_kernel<Body, MAXT, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
  wb_kernel<Body, MAXT, MINB><<<(unsigned)grid, block, smem_bytes, stream>>>(body);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e - 1000;
}
template <class Body>
inline int wb_launch(const Body& body, long long grid, int block, size_t smem_bytes, wb_stream_t stream) {
  return wb_launch_b<Body, 256, 1>(body, grid, block, smem_bytes, stream);  // every generic launch uses <= 256 threads
}
// spectral per-frame kernels: cap at 64 registers (4 blocks of 256 or 2 blocks of 512 threads per SM)
template <class Body>
inline int wb_launch_spectral(const Body& body, long long grid, int block, size_t smem_bytes, wb_stream_t stream) {
  if (block <= 256) return wb_launch_b<Body, 256, 4>(body, grid, block, smem_bytes, stream);
  return wb_launch_b<Body, 512, 2>(body, grid, block, smem_bytes, stream);
}
// kernels whose shared memory allows three blocks of 256 threads per SM anyway: 80 registers per thread
template <class Body>
inline int wb_launch_spectral3(const Body& body, long long grid, int block, size_t smem_bytes, wb_stream_t stream) {
  if (block <= 256) return wb_launch_b<Body, 256, 3>(body, grid, block, smem_bytes, stream);
  return wb_launch_b<Body, 512, 2>(body, grid, block, smem_bytes, stream);
}
#endif
