// Exact Euclidean distance transform to the nearest ZERO voxel with per-axis spacing:
// scipy.ndimage.distance_transform_edt(mask, sampling) as called at
// nocodazole/fish_analysis_noco.py:218 and simulation/analyze_simulation.py:101.
//
// Separable, three bandwidth-bound line passes over the device layout [z][x][y]:
//   pass 1 (y, contiguous)  edt_scan_y_kernel: one warp per row.  The row becomes a bit string
//                           (1 = zero voxel = site), 32 voxels per lane-word; nearest site to the
//                           left / right = clz / ffs inside the word + a warp prefix-max / suffix-min
//                           carry across words.  u8 in, u16 |dy| out (0xFFFF = no site in the row).
//   pass 2 (x) and 3 (z)    edt_envelope_kernel: one thread per line, consecutive threads on
//                           consecutive y, so every access of a warp is one contiguous segment.
//                           Lower envelope of the parabolas F(q) + w^2 (p-q)^2 (Felzenszwalb &
//                           Huttenlocher) kept as an IN-PLACE linked stack: link[q] = site below q
//                           on the hull when q was pushed (same layout as the volume, so pushes are
//                           coalesced and no per-thread stack memory exists).  The hidden-site test
//                           is a cross-multiplication in float64 -- exact for the integer
//                           (isotropic / equal in-plane spacing) path, where every operand is an
//                           integer < 2^53.  A backward sweep walks the hull and evaluates.
// Algorithmic bytes per voxel: pass 1 R1+W2, pass 2 R2+W2(link)+W4, pass 3 R4+W2(link)+W4.
#include <algorithm>
#include <type_traits>

#include "common.cuh"

namespace mfish {

constexpr int EDT_NO_SITE = -1;
constexpr uint16_t EDT_INF_U16 = 0xFFFFu;
constexpr int32_t EDT_INF_I32 = 0x7FFFFFFF;

// ------------------------------------------------------------------------------------------
// pass 1: distance along y to the nearest zero voxel of the row
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t zero_bytes_to_nibble(uint32_t four_bytes) {
  // bit k of the result = (byte k == 0)
  uint32_t eq = __vcmpeq4(four_bytes, 0u) & 0x01010101u;
  return ((eq * 0x00204081u) >> 21) & 0xFu;
}

struct ScanArgs {
  const uint8_t* mask;
  uint16_t* dy;
  int64_t nrows;
  int ny;
  int nwords;  // ceil(ny / 32)
};

constexpr int SCAN_WARPS = 8;
constexpr int SCAN_FAR = 1 << 24;

__global__ void __launch_bounds__(SCAN_WARPS * 32) edt_scan_y_kernel(const ScanArgs a) {
  extern __shared__ uint32_t scan_smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t* bits = scan_smem + (size_t)wid * 3 * a.nwords;
  int* carryL = reinterpret_cast<int*>(bits + a.nwords);
  int* carryR = carryL + a.nwords;
  const int ny = a.ny, nwords = a.nwords;

  for (int64_t row = (int64_t)blockIdx.x * SCAN_WARPS + wid; row < a.nrows;
       row += (int64_t)gridDim.x * SCAN_WARPS) {
    const uint8_t* __restrict__ src = a.mask + row * ny;
    uint16_t* __restrict__ dst = a.dy + row * ny;
    const bool vin = ((ny & 15) == 0) && aligned16_dev(a.mask);
    // 1. site bit words
    for (int w = lane; w < nwords; w += 32) {
      const int y0 = w * 32;
      uint32_t b = 0;
      if (vin && y0 + 32 <= ny) {
        const uint4 lo = __ldg(reinterpret_cast<const uint4*>(src + y0));
        const uint4 hi = __ldg(reinterpret_cast<const uint4*>(src + y0 + 16));
        b = zero_bytes_to_nibble(lo.x) | (zero_bytes_to_nibble(lo.y) << 4) |
            (zero_bytes_to_nibble(lo.z) << 8) | (zero_bytes_to_nibble(lo.w) << 12) |
            (zero_bytes_to_nibble(hi.x) << 16) | (zero_bytes_to_nibble(hi.y) << 20) |
            (zero_bytes_to_nibble(hi.z) << 24) | (zero_bytes_to_nibble(hi.w) << 28);
      } else {
        for (int j = 0; j < 32; ++j) {
          const int y = y0 + j;
          if (y < ny && __ldg(src + y) == 0) b |= (1u << j);
        }
      }
      bits[w] = b;
    }
    __syncwarp();
    // 2. exclusive prefix-max of "last site" and exclusive suffix-min of "first site" over words
    int carry = -SCAN_FAR;
    for (int c = 0; c < nwords; c += 32) {
      const int w = c + lane;
      const uint32_t b = w < nwords ? bits[w] : 0u;
      int v = b ? (w * 32 + 31 - __clz(b)) : -SCAN_FAR;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = max(v, t);
      }
      int ex = __shfl_up_sync(0xffffffffu, v, 1);
      ex = lane == 0 ? carry : max(ex, carry);
      if (w < nwords) carryL[w] = ex;
      carry = max(carry, __shfl_sync(0xffffffffu, v, 31));
    }
    carry = SCAN_FAR;
    for (int c = ((nwords - 1) / 32) * 32; c >= 0; c -= 32) {
      const int w = c + lane;
      const uint32_t b = w < nwords ? bits[w] : 0u;
      int v = b ? (w * 32 + __ffs(b) - 1) : SCAN_FAR;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_down_sync(0xffffffffu, v, o);
        if (lane + o < 32) v = min(v, t);
      }
      int ex = __shfl_down_sync(0xffffffffu, v, 1);
      ex = lane == 31 ? carry : min(ex, carry);
      if (w < nwords) carryR[w] = ex;
      carry = min(carry, __shfl_sync(0xffffffffu, v, 0));
    }
    __syncwarp();
    // 3. per voxel distance
    const bool vout = ((ny & 7) == 0) && aligned16_dev(a.dy);
    for (int w = lane; w < nwords; w += 32) {
      const uint32_t b = bits[w];
      const int y0 = w * 32;
      const int cl = carryL[w], cr = carryR[w];
      const bool full = vout && y0 + 32 <= ny;
#pragma unroll 1
      for (int k = 0; k < 4; ++k) {  // 8 voxels -> one 128-bit store (keeps the register count low)
        uint32_t packed[4];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int j = 8 * k + jj;
          const uint32_t upto = b & (0xFFFFFFFFu >> (31 - j));
          const uint32_t from = b >> j;
          const int y = y0 + j;
          const int l = upto ? (y0 + 31 - __clz(upto)) : cl;
          const int r = from ? (y + __ffs(from) - 1) : cr;
          const int d = min(min(y - l, r - y), (int)EDT_INF_U16);
          if (jj & 1)
            packed[jj >> 1] |= (uint32_t)d << 16;
          else
            packed[jj >> 1] = (uint32_t)d;
        }
        if (full) {
          reinterpret_cast<uint4*>(dst + y0)[k] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        } else {
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int y = y0 + 8 * k + jj;
            if (y < ny) dst[y] = (uint16_t)((packed[jj >> 1] >> ((jj & 1) * 16)) & 0xFFFFu);
          }
        }
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// passes 2 and 3: lower envelope along a strided axis
// ------------------------------------------------------------------------------------------
enum { EDT_OUT_I32 = 0, EDT_OUT_F32 = 1, EDT_OUT_DIST = 2 };

// uint32 -> double without the (slow) I2F pipe: 2^52 + v is exactly representable
__device__ __forceinline__ double u2d(uint32_t v) {
  return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;
}
// non-negative double < 2^31 holding an integer -> int32, without F2I
__device__ __forceinline__ int32_t d2i_exact(double v) { return __double2loint(v + 4503599627370496.0); }

// raw input value -> (valid, integer payload).  u16 rows hold |dy| (squared here).
template <typename TIn>
__device__ __forceinline__ bool edt_raw(const TIn* p, uint32_t& v);
template <>
__device__ __forceinline__ bool edt_raw<uint16_t>(const uint16_t* p, uint32_t& v) {
  const uint32_t d = __ldg(p);
  v = d * d;
  return d != EDT_INF_U16;
}
template <>
__device__ __forceinline__ bool edt_raw<int32_t>(const int32_t* p, uint32_t& v) {
  v = (uint32_t)__ldg(p);
  return v != (uint32_t)EDT_INF_I32;
}
template <>
__device__ __forceinline__ bool edt_raw<float>(const float* p, uint32_t& v) {
  const float f = __ldg(p);
  v = __float_as_uint(f);
  return f < INFINITY;
}
template <typename TIn>
__device__ __forceinline__ double edt_value(uint32_t v) {
  return u2d(v);
}
template <>
__device__ __forceinline__ double edt_value<float>(uint32_t v) {
  return (double)__uint_as_float(v);
}

template <typename TIn, typename LinkT>
struct EnvArgs {
  const TIn* in;
  LinkT* link;
  void* out;        // int32 / float, by OUT
  int32_t* out_sq;  // optional exact integer squared distance (EDT_OUT_DIST only)
  int64_t inner;
  int64_t outer_stride;
  int64_t stride;
  int L;
  double sin;        // F = sin * input value            (float64 arithmetic only)
  double w2;         // squared spacing of this axis     (float64 arithmetic only)
  double inv_w2;     // 1 / w2
  double out_scale;  // EDT_OUT_DIST: dist = sqrt(val * out_scale)
  float out_mul;     // (float)sqrt(out_scale)
};

constexpr int ENV_THREADS = 128;
constexpr int ENV_BATCH = 8;

// INT = true : unit weights, every quantity an integer < 2^31 -> hull tests are int32 x int32 ->
//              int64 products (one IMAD.WIDE each), no floating point at all.
// INT = false: float64 hull tests (exact whenever the operands are integers < 2^53).
// IdxT: element offsets inside a line bundle (uint32 when the volume has < 2^32 elements, so an
//       address is one IMAD.WIDE.U32 on top of the base pointer).
template <typename TIn, typename LinkT, int OUT, bool INT, typename IdxT>
__global__ void __launch_bounds__(ENV_THREADS) edt_envelope_kernel(const EnvArgs<TIn, LinkT> a) {
  const int64_t i = (int64_t)blockIdx.x * ENV_THREADS + threadIdx.x;
  if (i >= a.inner) return;
  const int64_t base = (int64_t)blockIdx.y * a.outer_stride + i;
  const TIn* __restrict__ in = a.in + base;
  LinkT* __restrict__ link = a.link + base;
  int32_t* __restrict__ out_i = reinterpret_cast<int32_t*>(a.out) + base;
  float* __restrict__ out_f = reinterpret_cast<float*>(a.out) + base;
  int32_t* __restrict__ out_sq = a.out_sq ? a.out_sq + base : nullptr;
  const bool has_out = a.out != nullptr;
  const IdxT st = (IdxT)a.stride;
  const int L = a.L;
  using GT = typename std::conditional<INT, int32_t, double>::type;
  const double sin = a.sin, inv_w2 = a.inv_w2, w2 = a.w2;
  const double sG = sin * inv_w2;  // G(u) = F(u)/w2 + u^2 = sG * raw + u^2

  auto makeG = [&](uint32_t raw, int u) -> GT {
    if (INT) return (GT)(int32_t)(raw + (uint32_t)(u * u));
    return (GT)fma(sG, edt_value<TIn>(raw), u2d((uint32_t)(u * u)));
  };
  // site `top` is hidden by (sec, q)  <=>  (Gtop - Gsec)(q - top) >= (Gq - Gtop)(top - sec)
  auto hidden = [&](GT Gsec, GT Gtop, GT Gq, int sec, int top, int q) -> bool {
    if (INT) {
      const long long lhs = (long long)((int32_t)Gtop - (int32_t)Gsec) * (long long)(q - top);
      const long long rhs = (long long)((int32_t)Gq - (int32_t)Gtop) * (long long)(top - sec);
      return lhs >= rhs;
    }
    const double lhs = ((double)Gtop - (double)Gsec) * u2d((uint32_t)(q - top));
    const double rhs = ((double)Gq - (double)Gtop) * u2d((uint32_t)(top - sec));
    return lhs >= rhs;
  };

  // The top three hull entries live in registers (t0 = top); deeper ones are re-read through the
  // link chain only when a burst of pops runs the cache dry.  EDT_UNKNOWN = "not cached".
  constexpr int EDT_UNKNOWN = -2;
  int t0 = EDT_NO_SITE, t1 = EDT_NO_SITE, t2 = EDT_NO_SITE;
  GT G0 = 0, G1 = 0, G2 = 0;

  const TIn* __restrict__ pin = in;
  LinkT* __restrict__ plink = link;
  for (int q0 = 0; q0 < L; q0 += ENV_BATCH) {
    uint32_t raw[ENV_BATCH];
    bool ok[ENV_BATCH];
#pragma unroll
    for (int u = 0; u < ENV_BATCH; ++u) {
      ok[u] = false;
      raw[u] = 0;
      if (q0 + u < L) ok[u] = edt_raw<TIn>(pin + (IdxT)u * st, raw[u]);
    }
#pragma unroll
    for (int u = 0; u < ENV_BATCH; ++u) {
      if (ok[u]) {
        const int q = q0 + u;
        const GT Gq = makeG(raw[u], q);
        while (t0 != EDT_NO_SITE) {
          if (t1 == EDT_UNKNOWN) {
            t1 = (int)link[(IdxT)t0 * st];
            if (t1 != EDT_NO_SITE) {
              uint32_t r;
              edt_raw<TIn>(in + (IdxT)t1 * st, r);
              G1 = makeG(r, t1);
            }
          }
          if (t1 == EDT_NO_SITE || !hidden(G1, G0, Gq, t1, t0, q)) break;
          t0 = t1;
          G0 = G1;
          t1 = t2;
          G1 = G2;
          t2 = EDT_UNKNOWN;
        }
        plink[(IdxT)u * st] = (LinkT)t0;
        t2 = t1;
        G2 = G1;
        t1 = t0;
        G1 = G0;
        t0 = q;
        G0 = Gq;
      }
    }
    pin += (IdxT)ENV_BATCH * st;
    plink += (IdxT)ENV_BATCH * st;
  }

  // backward sweep over the hull.  Parabola u at p (divided by w2) is G(u) - 2 p u + p^2, so with
  //   D(p) = (Gprev - Gcur) + 2 p (cur - prev)     prev is at least as good as cur  <=>  D(p) <= 0
  //   val(p) = F(cur) + w2 (p - cur)^2
  // both are advanced incrementally as p steps down (exact for integer-valued operands):
  //   D(p-1) = D(p) - 2 (cur - prev),   val(p-1) = val(p) - e(p),  e(p) = w2 (2 (p-cur) - 1),  e(p-1) = e(p) - 2 w2
  int cur = t0, prev = t1;
  GT Gcur = G0, Gprev = G1;
  if (cur != EDT_NO_SITE && prev == EDT_UNKNOWN) {
    prev = (int)link[(IdxT)cur * st];
    if (prev != EDT_NO_SITE) {
      uint32_t r;
      edt_raw<TIn>(in + (IdxT)prev * st, r);
      Gprev = makeG(r, prev);
    }
  }
  // one entry ahead: index of the entry below prev, plus its payload and its own link, are
  // requested when prev is entered and consumed only at the next hull move
  int nxt = EDT_NO_SITE;
  uint32_t nxt_raw = 0;
  int nxt_link = EDT_NO_SITE;
  auto prefetch_below = [&](int idx) {  // idx = entry whose link names the new nxt
    nxt = EDT_NO_SITE;
    if (idx != EDT_NO_SITE) nxt = (int)link[(IdxT)idx * st];
  };
  auto request_nxt = [&]() {
    if (nxt != EDT_NO_SITE) {
      edt_raw<TIn>(in + (IdxT)nxt * st, nxt_raw);
      nxt_link = (int)link[(IdxT)nxt * st];
    }
  };
  prefetch_below(prev);
  request_nxt();
  IdxT o = (IdxT)(L - 1) * st;
  if (cur == EDT_NO_SITE) {  // no site on this line
    for (int p = L - 1; p >= 0; --p, o -= st) {
      if (OUT == EDT_OUT_I32) {
        out_i[o] = EDT_INF_I32;
      } else {
        if (has_out) out_f[o] = INFINITY;
        if (OUT == EDT_OUT_DIST && out_sq) out_sq[o] = EDT_INF_I32;
      }
    }
    return;
  }
  using WT = typename std::conditional<INT, long long, double>::type;
  const float out_mul = a.out_mul;
  // INT: D and val advance incrementally (exact integers).  float64: both are re-evaluated from
  // cached per-pair constants with one fma each (an incremental sum would cancel catastrophically
  // when val falls from far-field magnitudes to ~0 along a run).
  WT D = 0, Dstep = 0, Gdiff = 0;  // valid while prev != EDT_NO_SITE
  WT val = 0, e = 0;
  double Fcur = 0.0, curd = 0.0;
  auto enter = [&](int p) {  // (re)initialise the running quantities for the pair (prev, cur) at p
    const int dp = p - cur;
    if (INT) {
      val = (WT)((int32_t)Gcur - cur * cur + dp * dp);
      e = (WT)(2 * dp - 1);
      if (prev != EDT_NO_SITE) {
        Dstep = (WT)(2 * (cur - prev));
        D = (WT)((int32_t)Gprev - (int32_t)Gcur) + (long long)p * (long long)Dstep;
      }
    } else {
      curd = u2d((uint32_t)cur);
      Fcur = ((double)Gcur - curd * curd) * w2;
      if (prev != EDT_NO_SITE) {
        Dstep = (WT)u2d((uint32_t)(2 * (cur - prev)));
        Gdiff = (WT)((double)Gprev - (double)Gcur);
      }
    }
  };
  enter(L - 1);
  double pd = u2d((uint32_t)(L - 1));
  for (int p = L - 1; p >= 0; --p, o -= st) {
    if (!INT) D = (WT)fma(pd, (double)Dstep, (double)Gdiff);
    while (prev != EDT_NO_SITE && D <= (WT)0) {  // move down the hull
      cur = prev;
      Gcur = Gprev;
      prev = nxt;
      if (prev != EDT_NO_SITE) Gprev = makeG(nxt_raw, prev);
      nxt = (prev != EDT_NO_SITE) ? nxt_link : EDT_NO_SITE;
      request_nxt();
      enter(p);
      if (!INT) D = (WT)fma(pd, (double)Dstep, (double)Gdiff);
    }
    if (INT) {
      const int32_t v32 = (int32_t)val;
      if (OUT == EDT_OUT_I32) {
        out_i[o] = v32;
      } else if (OUT == EDT_OUT_F32) {
        out_f[o] = (float)v32;
      } else {
        // float32 sqrt of the (exact) integer: |rel err| < 1e-7 after the two roundings
        if (has_out) out_f[o] = sqrtf((float)v32) * out_mul;
        if (out_sq) out_sq[o] = v32;
      }
      val -= e;
      e -= (WT)2;
      D -= Dstep;
    } else {
      const double dpd = pd - curd;
      const double v = fma(w2, dpd * dpd, Fcur);
      if (OUT == EDT_OUT_I32) {
        out_i[o] = d2i_exact(v);
      } else if (OUT == EDT_OUT_F32) {
        out_f[o] = (float)v;
      } else {
        if (has_out) out_f[o] = sqrtf((float)(v * a.out_scale));
        if (out_sq) out_sq[o] = d2i_exact(v);
      }
      pd -= 1.0;
    }
  }
}

template <typename TIn, typename LinkT, int OUT, bool INT>
static int launch_env(const TIn* in, void* link, void* out, int32_t* out_sq, int axis, int64_t nz, int64_t nx,
                      int64_t ny, double sin, double w2, double out_scale, cudaStream_t st) {
  EnvArgs<TIn, LinkT> a;
  a.in = in;
  a.link = reinterpret_cast<LinkT*>(link);
  a.out = out;
  a.out_sq = out_sq;
  int64_t nouter;
  if (axis == MFISH_AXIS_Z) {
    a.inner = nx * ny;
    a.stride = nx * ny;
    a.outer_stride = 0;
    a.L = (int)nz;
    nouter = 1;
  } else {
    a.inner = ny;
    a.stride = ny;
    a.outer_stride = nx * ny;
    a.L = (int)nx;
    nouter = nz;
  }
  a.sin = sin;
  a.w2 = w2;
  a.inv_w2 = 1.0 / w2;
  a.out_scale = out_scale;
  a.out_mul = (float)sqrt(out_scale);
  int64_t bx = (a.inner + ENV_THREADS - 1) / ENV_THREADS;
  MFISH_REQUIRE(bx < (1ll << 31) && nouter < 65536, "edt: grid too large");
  dim3 grid((unsigned)bx, (unsigned)nouter);
  ProfScope prof(axis == MFISH_AXIS_Z ? "edt_envelope_z" : "edt_envelope_x", st);
  if (nz * nx * ny < (1ll << 32))
    edt_envelope_kernel<TIn, LinkT, OUT, INT, uint32_t><<<grid, ENV_THREADS, 0, st>>>(a);
  else
    edt_envelope_kernel<TIn, LinkT, OUT, INT, int64_t><<<grid, ENV_THREADS, 0, st>>>(a);
  count_launch(1);
  return check_launch("edt_envelope");
}

// int_arith: unit weights and max(F) + L^2 < 2^31 (checked by the caller)
template <typename TIn, int OUT>
static int launch_env_link(const TIn* in, void* link, bool link16, bool int_arith, void* out, int32_t* out_sq,
                           int axis, int64_t nz, int64_t nx, int64_t ny, double sin, double w2, double out_scale,
                           cudaStream_t st) {
  if (int_arith) {
    if (link16)
      return launch_env<TIn, int16_t, OUT, true>(in, link, out, out_sq, axis, nz, nx, ny, sin, w2, out_scale, st);
    return launch_env<TIn, int32_t, OUT, true>(in, link, out, out_sq, axis, nz, nx, ny, sin, w2, out_scale, st);
  }
  if (link16)
    return launch_env<TIn, int16_t, OUT, false>(in, link, out, out_sq, axis, nz, nx, ny, sin, w2, out_scale, st);
  return launch_env<TIn, int32_t, OUT, false>(in, link, out, out_sq, axis, nz, nx, ny, sin, w2, out_scale, st);
}

static inline int64_t align256(int64_t v) { return (v + 255) & ~int64_t(255); }

}  // namespace mfish

using namespace mfish;

// workspace: u16 dy | link (4 B) | f2 (4 B)
extern "C" int64_t mfish_edt_workspace_bytes(int64_t nz, int64_t nx, int64_t ny) {
  if (nz <= 0 || nx <= 0 || ny <= 0) return 0;
  const int64_t n = nz * nx * ny;
  return align256(n * 2) + align256(n * 4) + align256(n * 4);
}

static int edt_spacing(const double* spacing_zxy_host, double& wz, double& wx, double& wy) {
  wz = wx = wy = 1.0;
  if (spacing_zxy_host) {
    wz = spacing_zxy_host[0];
    wx = spacing_zxy_host[1];
    wy = spacing_zxy_host[2];
  }
  MFISH_REQUIRE(wz > 0.0 && wx > 0.0 && wy > 0.0, "edt: spacing must be positive");
  return MFISH_OK;
}

// passes 1 + 2: squared in-plane distance.  *f2_kind_host tells how f2 is encoded:
// MFISH_EDT_F2_I32 (exact integer voxel^2, x and y spacing equal) or MFISH_EDT_F2_F32 (physical^2).
extern "C" int mfish_edt_inplane_u8(const uint8_t* mask_dev, int64_t nz, int64_t nx, int64_t ny,
                                    const double* spacing_zxy_host, void* f2_out_dev, int* f2_kind_host,
                                    void* workspace_dev, void* stream) {
  MFISH_REQUIRE(mask_dev && workspace_dev && f2_out_dev && f2_kind_host, "edt: null pointer");
  MFISH_REQUIRE(nz > 0 && nx > 0 && ny > 0, "edt: empty volume");
  MFISH_REQUIRE(ny < 65535, "edt: ny must be < 65535 (u16 row distances)");
  MFISH_REQUIRE(nz < (1 << 24) && nx < (1 << 24), "edt: dimension too large");
  double wz, wx, wy;
  int rc = edt_spacing(spacing_zxy_host, wz, wx, wy);
  if (rc != MFISH_OK) return rc;
  const double plane_sq = (double)(nx - 1) * (nx - 1) + (double)(ny - 1) * (ny - 1);
  const bool plane_int = (wx == wy) && plane_sq < 2147483000.0;
  cudaStream_t st = as_stream(stream);
  const int64_t n = nz * nx * ny;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace_dev);
  uint16_t* dy = reinterpret_cast<uint16_t*>(ws);
  void* link = ws + align256(n * 2);
  {
    ScanArgs a{mask_dev, dy, nz * nx, (int)ny, (int)((ny + 31) / 32)};
    size_t smem = (size_t)SCAN_WARPS * 3 * a.nwords * sizeof(uint32_t);
    MFISH_REQUIRE(smem <= 200 * 1024, "edt: row too long");
    if (smem > 48 * 1024)
      MFISH_CUDA(cudaFuncSetAttribute(edt_scan_y_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = std::min<int64_t>((a.nrows + SCAN_WARPS - 1) / SCAN_WARPS, (int64_t)num_sms() * 32);
    ProfScope prof("edt_scan_y", st);
    edt_scan_y_kernel<<<(unsigned)blocks, SCAN_WARPS * 32, smem, st>>>(a);
    count_launch(1);
    rc = check_launch("edt_scan_y");
    if (rc != MFISH_OK) return rc;
  }
  const bool link16 = nx <= 32767;
  if (plane_int) {
    *f2_kind_host = MFISH_EDT_F2_I32;
    // all-integer hull tests need max(F) + L^2 = (ny-1)^2 + nx^2 below 2^31
    const bool int_arith = (double)(ny - 1) * (ny - 1) + (double)nx * nx < 2147483000.0;
    return launch_env_link<uint16_t, EDT_OUT_I32>(dy, link, link16, int_arith, f2_out_dev, nullptr, MFISH_AXIS_X, nz,
                                                  nx, ny, 1.0, 1.0, 1.0, st);
  }
  *f2_kind_host = MFISH_EDT_F2_F32;
  return launch_env_link<uint16_t, EDT_OUT_F32>(dy, link, link16, false, f2_out_dev, nullptr, MFISH_AXIS_X, nz, nx,
                                                ny, wy * wy, wx * wx, 1.0, st);
}

// pass 3 along z.  link_ws_dev: nz*nx*ny*4 bytes of scratch.
extern "C" int mfish_edt_zpass(const void* f2_dev, int f2_kind, int64_t nz, int64_t nx, int64_t ny,
                               const double* spacing_zxy_host, float* out_dist_dev, int32_t* out_sq_dev,
                               void* link_ws_dev, void* stream) {
  MFISH_REQUIRE(f2_dev && link_ws_dev, "edt: null pointer");
  MFISH_REQUIRE(out_dist_dev || out_sq_dev, "edt: no output requested");
  MFISH_REQUIRE(nz > 0 && nx > 0 && ny > 0 && nz < (1 << 24), "edt: bad dimensions");
  MFISH_REQUIRE(f2_kind == MFISH_EDT_F2_I32 || f2_kind == MFISH_EDT_F2_F32, "edt: bad f2 kind");
  double wz, wx, wy;
  int rc = edt_spacing(spacing_zxy_host, wz, wx, wy);
  if (rc != MFISH_OK) return rc;
  const bool iso = (wz == wx) && (wx == wy);
  if (out_sq_dev) {
    if (!iso || f2_kind != MFISH_EDT_F2_I32) {
      set_error("edt: integer squared distances are defined for isotropic spacing only");
      return MFISH_ERR_UNSUPPORTED;
    }
    const double max_sq =
        (double)(nz - 1) * (nz - 1) + (double)(nx - 1) * (nx - 1) + (double)(ny - 1) * (ny - 1);
    MFISH_REQUIRE(max_sq < 2147483000.0, "edt: volume too large for int32 squared distances");
  }
  cudaStream_t st = as_stream(stream);
  const bool link16 = nz <= 32767;
  if (f2_kind == MFISH_EDT_F2_I32) {
    const int32_t* f2 = reinterpret_cast<const int32_t*>(f2_dev);
    if (iso) {  // keep integers, scale once at the end: dist = sqrt(sq * w^2)
      const double gmax = (double)(nx - 1) * (nx - 1) + (double)(ny - 1) * (ny - 1) + (double)nz * nz;
      return launch_env_link<int32_t, EDT_OUT_DIST>(f2, link_ws_dev, link16, gmax < 2147483000.0, out_dist_dev,
                                                    out_sq_dev, MFISH_AXIS_Z, nz, nx, ny, 1.0, 1.0, wz * wz, st);
    }
    return launch_env_link<int32_t, EDT_OUT_DIST>(f2, link_ws_dev, link16, false, out_dist_dev, nullptr,
                                                  MFISH_AXIS_Z, nz, nx, ny, wx * wx, wz * wz, 1.0, st);
  }
  return launch_env_link<float, EDT_OUT_DIST>(reinterpret_cast<const float*>(f2_dev), link_ws_dev, link16, false,
                                              out_dist_dev, nullptr, MFISH_AXIS_Z, nz, nx, ny, 1.0, wz * wz, 1.0,
                                              st);
}

extern "C" int mfish_edt3d_u8(const uint8_t* mask_dev, int64_t nz, int64_t nx, int64_t ny,
                              const double* spacing_zxy_host, float* out_dist_dev, int32_t* out_sq_dev,
                              void* workspace_dev, void* stream) {
  MFISH_REQUIRE(mask_dev && workspace_dev, "edt: null pointer");
  MFISH_REQUIRE(out_dist_dev || out_sq_dev, "edt: no output requested");
  MFISH_REQUIRE(nz > 0 && nx > 0 && ny > 0, "edt: empty volume");
  const int64_t n = nz * nx * ny;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace_dev);
  void* link = ws + align256(n * 2);
  void* f2 = ws + align256(n * 2) + align256(n * 4);
  int kind = 0;
  int rc = mfish_edt_inplane_u8(mask_dev, nz, nx, ny, spacing_zxy_host, f2, &kind, workspace_dev, stream);
  if (rc != MFISH_OK) return rc;
  return mfish_edt_zpass(f2, kind, nz, nx, ny, spacing_zxy_host, out_dist_dev, out_sq_dev, link, stream);
}
