// Ingest: reference-layout (x,y,z) host data -> device layout [z][x][y] float32 / uint8,
// fused with the global min/max that su.normalize_image needs
// (modules/scope_utils3.py:381-386).  HBM-bound: R sizeof(S) + W 4 (or 1) bytes per voxel.
#include <algorithm>

#include "common.cuh"

namespace mfish {
void count_launch(int k);

__global__ void minmax_init_kernel(uint32_t* mm) {
  mm[0] = 0xFFFFFFFFu;
  mm[1] = 0u;
}
__global__ void minmax_finalize_kernel(uint32_t* mm) {
  float lo = key_to_f32(mm[0]);
  float hi = key_to_f32(mm[1]);
  reinterpret_cast<float*>(mm)[0] = lo;
  reinterpret_cast<float*>(mm)[1] = hi;
}

template <typename S>
__device__ __forceinline__ float cvt_value(S v) {
  return static_cast<float>(v);
}
template <typename S, int DST>
__device__ __forceinline__ float cvt_flag(S v) {
  // MFISH_DST_NONZERO: 1 where v != 0;  MFISH_DST_ISZERO: 1 where v == 0 (np.logical_not)
  return ((v != S(0)) == (DST == MFISH_DST_NONZERO)) ? 1.0f : 0.0f;
}

constexpr int TY = 32;  // y-tile (becomes a 128-byte output row segment)
constexpr int ZC = 64;  // z-chunk held in shared memory

// src is (x,y,z) C-order.  Block = (y-tile, x) x z-chunk; 256 threads.
template <typename S, int DST>
__global__ void __launch_bounds__(256)
ingest_xyz_kernel(const S* __restrict__ src, void* __restrict__ dst, int nx, int ny, int nz,
                  int nytiles, uint32_t* mm) {
  __shared__ float tile[TY][ZC + 1];
  const int x = blockIdx.x / nytiles;
  const int y0 = (blockIdx.x - x * nytiles) * TY;
  const int z0 = blockIdx.y * ZC;
  const int zc = min(ZC, nz - z0);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float lo = INFINITY, hi = -INFINITY;
  for (int yy = w; yy < TY; yy += 8) {
    int y = y0 + yy;
    if (y < ny) {
      const S* row = src + ((int64_t)x * ny + y) * nz + z0;
      for (int zz = lane; zz < zc; zz += 32) {
        float v = (DST == MFISH_DST_F32) ? cvt_value<S>(row[zz]) : cvt_flag<S, DST>(row[zz]);
        tile[yy][zz] = v;
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
      }
    }
  }
  __syncthreads();
  const int y = y0 + lane;
  if (y < ny) {
    for (int zz = w; zz < zc; zz += 8) {
      int64_t o = ((int64_t)(z0 + zz) * nx + x) * ny + y;
      float v = tile[lane][zz];
      if (DST == MFISH_DST_F32)
        reinterpret_cast<float*>(dst)[o] = v;
      else
        reinterpret_cast<uint8_t*>(dst)[o] = v != 0.0f ? 1 : 0;
    }
  }
  if (mm != nullptr) block_minmax_commit(lo, hi, mm);
}

// Same element order in and out (already z-first): pure conversion.
template <typename S, int DST>
__global__ void __launch_bounds__(256)
ingest_flat_kernel(const S* __restrict__ src, void* __restrict__ dst, int64_t n, uint32_t* mm) {
  float lo = INFINITY, hi = -INFINITY;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = (DST == MFISH_DST_F32) ? cvt_value<S>(src[i]) : cvt_flag<S, DST>(src[i]);
    if (DST == MFISH_DST_F32)
      reinterpret_cast<float*>(dst)[i] = v;
    else
      reinterpret_cast<uint8_t*>(dst)[i] = v != 0.0f ? 1 : 0;
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  if (mm != nullptr) block_minmax_commit(lo, hi, mm);
}

// Vectorised min/max over a float volume: 4 B/voxel, one 128-bit load per 4 voxels.
__global__ void __launch_bounds__(256)
minmax_kernel(const float* __restrict__ v, int64_t n, uint32_t* mm) {
  float lo = INFINITY, hi = -INFINITY;
  int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // head until 16-byte aligned
  int64_t head = ((16 - (reinterpret_cast<uintptr_t>(v) & 15)) & 15) / 4;
  if (head > n) head = n;
  if (tid < head) {
    lo = fminf(lo, v[tid]);
    hi = fmaxf(hi, v[tid]);
  }
  const float4* v4 = reinterpret_cast<const float4*>(v + head);
  int64_t n4 = (n - head) / 4;
  for (int64_t i = tid; i < n4; i += stride) {
    float4 a = __ldg(&v4[i]);
    lo = fminf(fminf(lo, a.x), fminf(a.y, fminf(a.z, a.w)));
    hi = fmaxf(fmaxf(hi, a.x), fmaxf(a.y, fmaxf(a.z, a.w)));
  }
  int64_t tail0 = head + n4 * 4;
  if (tail0 + tid < n) {
    lo = fminf(lo, v[tail0 + tid]);
    hi = fmaxf(hi, v[tail0 + tid]);
  }
  block_minmax_commit(lo, hi, mm);
}

template <typename S>
static int launch_ingest(const void* src, int layout, int64_t nx, int64_t ny, int64_t nz, void* dst,
                         int dst_kind, uint32_t* mm, cudaStream_t st) {
  const S* s = reinterpret_cast<const S*>(src);
  ProfScope prof(layout == MFISH_LAYOUT_XYZ ? "ingest_xyz" : "ingest_flat", st);
  if (layout == MFISH_LAYOUT_XYZ) {
    int nyt = (int)((ny + TY - 1) / TY);
    int64_t gx = (int64_t)nyt * nx;
    int64_t gy = (nz + ZC - 1) / ZC;
    MFISH_REQUIRE(gx < (1ll << 31) && gy < 65536, "ingest: volume too large for one launch");
    dim3 grid((unsigned)gx, (unsigned)gy);
    if (dst_kind == MFISH_DST_F32)
      ingest_xyz_kernel<S, MFISH_DST_F32><<<grid, 256, 0, st>>>(s, dst, (int)nx, (int)ny, (int)nz, nyt, mm);
    else if (dst_kind == MFISH_DST_NONZERO)
      ingest_xyz_kernel<S, MFISH_DST_NONZERO><<<grid, 256, 0, st>>>(s, dst, (int)nx, (int)ny, (int)nz, nyt, mm);
    else
      ingest_xyz_kernel<S, MFISH_DST_ISZERO><<<grid, 256, 0, st>>>(s, dst, (int)nx, (int)ny, (int)nz, nyt, mm);
  } else {
    int64_t n = nx * ny * nz;
    int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 16);
    if (dst_kind == MFISH_DST_F32)
      ingest_flat_kernel<S, MFISH_DST_F32><<<blocks, 256, 0, st>>>(s, dst, n, mm);
    else if (dst_kind == MFISH_DST_NONZERO)
      ingest_flat_kernel<S, MFISH_DST_NONZERO><<<blocks, 256, 0, st>>>(s, dst, n, mm);
    else
      ingest_flat_kernel<S, MFISH_DST_ISZERO><<<blocks, 256, 0, st>>>(s, dst, n, mm);
  }
  count_launch(1);
  return check_launch("ingest");
}

}  // namespace mfish

using namespace mfish;

extern "C" int mfish_ingest(const void* src_dev, int src_dtype, int src_layout, int64_t nx, int64_t ny,
                            int64_t nz, void* dst_dev, int dst_kind, float* minmax_dev, void* stream) {
  MFISH_REQUIRE(src_dev && dst_dev, "ingest: null pointer");
  MFISH_REQUIRE(nx > 0 && ny > 0 && nz > 0, "ingest: empty volume");
  MFISH_REQUIRE(nx < (1ll << 31) && ny < (1ll << 31) && nz < (1ll << 31), "ingest: dimension too large");
  MFISH_REQUIRE(src_layout == MFISH_LAYOUT_XYZ || src_layout == MFISH_LAYOUT_ZXY, "ingest: bad layout");
  MFISH_REQUIRE(dst_kind == MFISH_DST_F32 || dst_kind == MFISH_DST_NONZERO || dst_kind == MFISH_DST_ISZERO,
                "ingest: bad dst kind");
  cudaStream_t st = as_stream(stream);
  uint32_t* mm = reinterpret_cast<uint32_t*>(minmax_dev);
  if (mm) {
    minmax_init_kernel<<<1, 1, 0, st>>>(mm);
    count_launch(1);
  }
  int rc;
  switch (src_dtype) {
    case MFISH_U8: rc = launch_ingest<uint8_t>(src_dev, src_layout, nx, ny, nz, dst_dev, dst_kind, mm, st); break;
    case MFISH_U16: rc = launch_ingest<uint16_t>(src_dev, src_layout, nx, ny, nz, dst_dev, dst_kind, mm, st); break;
    case MFISH_I16: rc = launch_ingest<int16_t>(src_dev, src_layout, nx, ny, nz, dst_dev, dst_kind, mm, st); break;
    case MFISH_I32: rc = launch_ingest<int32_t>(src_dev, src_layout, nx, ny, nz, dst_dev, dst_kind, mm, st); break;
    case MFISH_I64: rc = launch_ingest<int64_t>(src_dev, src_layout, nx, ny, nz, dst_dev, dst_kind, mm, st); break;
    case MFISH_F32: rc = launch_ingest<float>(src_dev, src_layout, nx, ny, nz, dst_dev, dst_kind, mm, st); break;
    case MFISH_F64: rc = launch_ingest<double>(src_dev, src_layout, nx, ny, nz, dst_dev, dst_kind, mm, st); break;
    default: set_error("ingest: unsupported dtype %d", src_dtype); return MFISH_ERR_UNSUPPORTED;
  }
  if (rc != MFISH_OK) return rc;
  if (mm) {
    minmax_finalize_kernel<<<1, 1, 0, st>>>(mm);
    count_launch(1);
  }
  return check_launch("ingest/finalize");
}

extern "C" int mfish_minmax_f32(const float* vol_dev, int64_t n, float* minmax_dev, void* stream) {
  MFISH_REQUIRE(vol_dev && minmax_dev && n > 0, "minmax: bad arguments");
  cudaStream_t st = as_stream(stream);
  uint32_t* mm = reinterpret_cast<uint32_t*>(minmax_dev);
  ProfScope prof("minmax", st);
  minmax_init_kernel<<<1, 1, 0, st>>>(mm);
  int blocks = (int)std::min<int64_t>((n / 4 + 255) / 256 + 1, (int64_t)num_sms() * 8);
  minmax_kernel<<<blocks, 256, 0, st>>>(vol_dev, n, mm);
  minmax_finalize_kernel<<<1, 1, 0, st>>>(mm);
  count_launch(3);
  return check_launch("minmax");
}

// --------------------------------------------------------------------------------------------
// su.normalize_image as a stand-alone call (modules/scope_utils3.py:381-386): float64 result in
// the caller's element order, arithmetic identical to np.interp with two knots
// (slope*(x - lo) + vmin, rounded after the multiply and after the add; exact fp at the knots).
// --------------------------------------------------------------------------------------------
namespace mfish {

__host__ __device__ __forceinline__ unsigned long long f64_to_key(double f) {
#ifdef __CUDA_ARCH__
  unsigned long long u = (unsigned long long)__double_as_longlong(f);
#else
  unsigned long long u;
  memcpy(&u, &f, 8);
#endif
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_to_f64(unsigned long long k) {
  unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

__global__ void minmax64_init_kernel(unsigned long long* mm) {
  mm[0] = ~0ull;
  mm[1] = 0ull;
}

template <typename S>
__global__ void __launch_bounds__(256) minmax64_kernel(const S* __restrict__ src, int64_t n,
                                                        unsigned long long* mm) {
  double lo = INFINITY, hi = -INFINITY;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double v = (double)src[i];
    lo = fmin(lo, v);
    hi = fmax(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ double s_lo[8], s_hi[8];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
    s_lo[w] = lo;
    s_hi[w] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
      lo = fmin(lo, s_lo[k]);
      hi = fmax(hi, s_hi[k]);
    }
    atomicMin(&mm[0], f64_to_key(lo));
    atomicMax(&mm[1], f64_to_key(hi));
  }
}

template <typename S>
__global__ void __launch_bounds__(256) normalize64_kernel(const S* __restrict__ src, int64_t n,
                                                           const unsigned long long* __restrict__ mm,
                                                           double vmin, double vmax, double* __restrict__ out) {
  const double lo = key_to_f64(mm[0]), hi = key_to_f64(mm[1]);
  const double slope = (vmax - vmin) / (hi - lo);
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double x = (double)src[i];
    double r;
    if (x >= hi)
      r = vmax;  // last knot (also the degenerate lo == hi case)
    else if (x == lo)
      r = vmin;
    else
      r = __dadd_rn(__dmul_rn(slope, x - lo), vmin);
    out[i] = r;
  }
}

template <typename S>
static int launch_normalize(const void* src, int64_t n, double vmin, double vmax, double* out,
                            unsigned long long* mm, cudaStream_t st) {
  const S* s = reinterpret_cast<const S*>(src);
  int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 16);
  minmax64_init_kernel<<<1, 1, 0, st>>>(mm);
  minmax64_kernel<S><<<blocks, 256, 0, st>>>(s, n, mm);
  normalize64_kernel<S><<<blocks, 256, 0, st>>>(s, n, mm, vmin, vmax, out);
  count_launch(3);
  return check_launch("normalize_f64");
}

// device [z][x][y] float32 -> reference order (x,y,z) float64 (dense outputs handed back to the
// reference's callers: distance map, filtered volumes).
__global__ void __launch_bounds__(256)
export_xyz_f64_kernel(const float* __restrict__ src, double* __restrict__ dst, int nx, int ny, int nz,
                      int nytiles) {
  __shared__ float tile[ZC][TY + 1];
  const int x = blockIdx.x / nytiles;
  const int y0 = (blockIdx.x - x * nytiles) * TY;
  const int z0 = blockIdx.y * ZC;
  const int zc = min(ZC, nz - z0);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int y = y0 + lane;
  for (int zz = w; zz < zc; zz += 8) {
    if (y < ny) tile[zz][lane] = src[((int64_t)(z0 + zz) * nx + x) * ny + y];
  }
  __syncthreads();
  for (int yy = w; yy < TY; yy += 8) {
    int yo = y0 + yy;
    if (yo < ny) {
      double* row = dst + ((int64_t)x * ny + yo) * nz + z0;
      for (int zz = lane; zz < zc; zz += 32) row[zz] = (double)tile[zz][yy];
    }
  }
}

__global__ void __launch_bounds__(256)
export_flat_f64_kernel(const float* __restrict__ src, double* __restrict__ dst, int64_t n) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (double)src[i];
}

}  // namespace mfish

extern "C" int mfish_normalize_f64(const void* src_dev, int src_dtype, int64_t n, double vmin, double vmax,
                                   double* out_dev, void* ws16_dev, void* stream) {
  MFISH_REQUIRE(src_dev && out_dev && ws16_dev && n > 0, "normalize: bad arguments");
  cudaStream_t st = as_stream(stream);
  unsigned long long* mm = reinterpret_cast<unsigned long long*>(ws16_dev);
  switch (src_dtype) {
    case MFISH_U8: return launch_normalize<uint8_t>(src_dev, n, vmin, vmax, out_dev, mm, st);
    case MFISH_U16: return launch_normalize<uint16_t>(src_dev, n, vmin, vmax, out_dev, mm, st);
    case MFISH_I16: return launch_normalize<int16_t>(src_dev, n, vmin, vmax, out_dev, mm, st);
    case MFISH_I32: return launch_normalize<int32_t>(src_dev, n, vmin, vmax, out_dev, mm, st);
    case MFISH_I64: return launch_normalize<int64_t>(src_dev, n, vmin, vmax, out_dev, mm, st);
    case MFISH_F32: return launch_normalize<float>(src_dev, n, vmin, vmax, out_dev, mm, st);
    case MFISH_F64: return launch_normalize<double>(src_dev, n, vmin, vmax, out_dev, mm, st);
    default: set_error("normalize: unsupported dtype %d", src_dtype); return MFISH_ERR_UNSUPPORTED;
  }
}

extern "C" int mfish_export_f64(const float* vol_zxy_dev, int64_t nx, int64_t ny, int64_t nz, int dst_layout,
                                double* out_dev, void* stream) {
  MFISH_REQUIRE(vol_zxy_dev && out_dev && nx > 0 && ny > 0 && nz > 0, "export: bad arguments");
  MFISH_REQUIRE(dst_layout == MFISH_LAYOUT_XYZ || dst_layout == MFISH_LAYOUT_ZXY, "export: bad layout");
  if (dst_layout == MFISH_LAYOUT_ZXY) {
    int64_t n = nx * ny * nz;
    int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 16);
    export_flat_f64_kernel<<<blocks, 256, 0, as_stream(stream)>>>(vol_zxy_dev, out_dev, n);
    count_launch(1);
    return check_launch("export_flat_f64");
  }
  int nyt = (int)((ny + TY - 1) / TY);
  int64_t gx = (int64_t)nyt * nx;
  int64_t gy = (nz + ZC - 1) / ZC;
  MFISH_REQUIRE(gx < (1ll << 31) && gy < 65536, "export: volume too large for one launch");
  export_xyz_f64_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, as_stream(stream)>>>(
      vol_zxy_dev, out_dev, (int)nx, (int)ny, (int)nz, nyt);
  count_launch(1);
  return check_launch("export_xyz_f64");
}
