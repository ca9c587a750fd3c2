// Shared helpers for libmfish_sm100 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <vector>

#include "../../include/mfish.h"

namespace mfish {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return MFISH_ERR_CUDA;
  }
  return MFISH_OK;
}

#define MFISH_REQUIRE(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      mfish::set_error(__VA_ARGS__);        \
      return MFISH_ERR_INVALID;             \
    }                                       \
  } while (0)

#define MFISH_CUDA(call)                                              \
  do {                                                                \
    cudaError_t e__ = (call);                                         \
    if (e__ != cudaSuccess) {                                         \
      mfish::set_error("%s: %s", #call, cudaGetErrorString(e__));     \
      return MFISH_ERR_CUDA;                                          \
    }                                                                 \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
__device__ __forceinline__ bool aligned16_dev(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// edt.cu: distance along y to the nearest site of each row (u16, 0xFFFF = none); sites are the zero
// voxels, or the nonzero ones when site_nonzero
int launch_scan_y(const uint8_t* mask, uint16_t* dy, int64_t nrows, int64_t ny, int site_nonzero, cudaStream_t st);

// bigconv.cu: one plain-Gaussian pass with a large radius (FMA-bound regime); taps_dev = radius + 1 half taps
int big_gauss_pass(const float* in, float* out, int64_t nz, int64_t nx, int64_t ny, int axis, const float* taps_dev,
                   int radius, int boundary, const float* mm, cudaStream_t st);

// Number of SMs on the current device (cached).
int num_sms();
// launch accounting (mfish_launch_count)
void count_launch(int k);

// Optional per-kernel timing (mfish_profile_enable / mfish_profile_collect): a scope records a
// CUDA event pair on the launching stream around the launches made inside it.
struct ProfScope {
  const char* name;
  cudaStream_t st;
  void* rec;
  ProfScope(const char* n, cudaStream_t s);
  ~ProfScope();
};

// ---- device helpers -------------------------------------------------------

// Monotone float <-> uint32 key (total order, -0 < +0, NaNs at the ends).
__host__ __device__ __forceinline__ uint32_t f32_to_key(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float key_to_f32(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

// Boundary extension of an index into [0, n): MFISH_BOUNDARY_REFLECT is
// scipy 'reflect' (d c b a | a b c d | d c b a, period 2n), NEAREST clamps.
template <int BMODE>
__device__ __forceinline__ int bmap(int i, int n) {
  if (BMODE == MFISH_BOUNDARY_NEAREST) {
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
  } else {
    if (i >= 0 && i < n) return i;
    int p = 2 * n;
    int m = i % p;
    if (m < 0) m += p;
    return m < n ? m : p - 1 - m;
  }
}

// Single-fold variant, valid for -n <= i <= 2n-1 (no modulo: used in the hot loops).
template <int BMODE>
__device__ __forceinline__ int bmap1(int i, int n) {
  if (BMODE == MFISH_BOUNDARY_NEAREST) {
    return min(max(i, 0), n - 1);
  } else {
    if (i < 0) i = -i - 1;
    if (i >= n) i = 2 * n - 1 - i;
    return i;
  }
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block reduction of (lo, hi) followed by one atomic pair per block.
__device__ __forceinline__ void block_minmax_commit(float lo, float hi, uint32_t* mm) {
  __shared__ float s_lo[32], s_hi[32];
  lo = warp_min(lo);
  hi = warp_max(hi);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int nw = (blockDim.x + 31) >> 5;
  if (lane == 0) {
    s_lo[w] = lo;
    s_hi[w] = hi;
  }
  __syncthreads();
  if (w == 0) {
    lo = lane < nw ? s_lo[lane] : INFINITY;
    hi = lane < nw ? s_hi[lane] : -INFINITY;
    lo = warp_min(lo);
    hi = warp_max(hi);
    if (lane == 0) {
      atomicMin(&mm[0], f32_to_key(lo));
      atomicMax(&mm[1], f32_to_key(hi));
    }
  }
}

// --------------------------------------------------------------------------------------------
// TMA bulk copy + mbarrier helpers (sm_90+ PTX; SASS: UBLKCP / SYNCS)
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMA bulk copy global -> shared, completion signalled on the mbarrier (16-byte aligned, size % 16 == 0)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes,
                                            unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

}  // namespace mfish
