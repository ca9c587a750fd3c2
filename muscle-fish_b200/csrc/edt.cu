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
#include <cstdlib>
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
  int nwords;       // ceil(ny / 32)
  int site_nonzero;  // 0: sites are the ZERO voxels (EDT); 1: the nonzero ones (dilation of a mask)
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
      if (a.site_nonzero) {  // complement inside the row
        const int rem = ny - y0;
        b = ~b & (rem >= 32 ? 0xFFFFFFFFu : ((1u << rem) - 1u));
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
      if (b == 0u && full) {
        // no site inside the word (the common case: nuclei fill a few percent of the volume): the
        // distance is a ramp from the carried neighbours, no bit scans
        int dl = min(y0 - cl, (int)EDT_INF_U16);      // y - l at j = 0, grows by 1 per voxel
        int dr = min(cr - y0, 2 * (int)EDT_INF_U16);  // r - y at j = 0, shrinks by 1 per voxel
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t packed[4];
#pragma unroll
          for (int jj = 0; jj < 8; jj += 2) {
            const int j = 8 * k + jj;
            const int d0 = min(min(dl + j, dr - j), (int)EDT_INF_U16);
            const int d1 = min(min(dl + j + 1, dr - j - 1), (int)EDT_INF_U16);
            packed[jj >> 1] = (uint32_t)d0 | ((uint32_t)d1 << 16);
          }
          reinterpret_cast<uint4*>(dst + y0)[k] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
        continue;
      }
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

// distance along y to the nearest site of each of the nrows rows (u16, 0xFFFF = no site in the row)
int launch_scan_y(const uint8_t* mask, uint16_t* dy, int64_t nrows, int64_t ny, int site_nonzero, cudaStream_t st) {
  MFISH_REQUIRE(ny > 0 && ny < 65535, "row scan: ny must be < 65535 (u16 row distances)");
  ScanArgs a{mask, dy, nrows, (int)ny, (int)((ny + 31) / 32), site_nonzero};
  size_t smem = (size_t)SCAN_WARPS * 3 * a.nwords * sizeof(uint32_t);
  MFISH_REQUIRE(smem <= 200 * 1024, "row scan: row too long");
  if (smem > 48 * 1024)
    MFISH_CUDA(cudaFuncSetAttribute(edt_scan_y_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks = std::min<int64_t>((a.nrows + SCAN_WARPS - 1) / SCAN_WARPS, (int64_t)num_sms() * 32);
  ProfScope prof("edt_scan_y", st);
  edt_scan_y_kernel<<<(unsigned)blocks, SCAN_WARPS * 32, smem, st>>>(a);
  count_launch(1);
  return check_launch("edt_scan_y");
}

// ------------------------------------------------------------------------------------------
// passes 2 and 3: lower envelope along a strided axis
// ------------------------------------------------------------------------------------------
enum { EDT_OUT_I32 = 0, EDT_OUT_F32 = 1, EDT_OUT_DIST = 2 };

// uint32 -> double without the (slow) I2F pipe: 2^52 + v is exactly representable
__device__ __forceinline__ double u2d(uint32_t v) {
  return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;
}
// MUFU.SQRT: relative error <= 2^-23, one instruction instead of the ~8 of the correctly rounded sqrtf
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// non-negative double < 2^31 holding an integer -> int32, without F2I
__device__ __forceinline__ int32_t d2i_exact(double v) { return __double2loint(v + 4503599627370496.0); }

// raw input value -> integer payload; the per-type sentinel marks "no site" (never a valid payload).
// u16 rows hold |dy| (squared here: at most 65534^2 < 0xFFFFFFFF).
template <typename TIn>
struct EdtRaw;
// fetch() only issues the load (the software-pipelined batches must not touch the value before it is
// consumed, or the thread waits for it on the spot); decode() turns the fetched bits into the payload.
template <>
struct EdtRaw<uint16_t> {
  static constexpr uint32_t kNone = 0xFFFFFFFFu;
  static constexpr uint32_t kFetchNone = EDT_INF_U16;
  static __device__ __forceinline__ uint32_t fetch(const uint16_t* p) { return __ldg(p); }
  static __device__ __forceinline__ uint32_t decode(uint32_t d) { return d == EDT_INF_U16 ? kNone : d * d; }
  static __device__ __forceinline__ uint32_t load(const uint16_t* p) { return decode(fetch(p)); }
};
template <>
struct EdtRaw<int32_t> {
  static constexpr uint32_t kNone = (uint32_t)EDT_INF_I32;
  static constexpr uint32_t kFetchNone = kNone;
  static __device__ __forceinline__ uint32_t fetch(const int32_t* p) { return (uint32_t)__ldg(p); }
  static __device__ __forceinline__ uint32_t decode(uint32_t v) { return v; }
  static __device__ __forceinline__ uint32_t load(const int32_t* p) { return fetch(p); }
};
template <>
struct EdtRaw<float> {
  static constexpr uint32_t kNone = 0x7F800000u;  // +inf
  static constexpr uint32_t kFetchNone = kNone;
  static __device__ __forceinline__ uint32_t fetch(const float* p) { return __float_as_uint(__ldg(p)); }
  static __device__ __forceinline__ uint32_t decode(uint32_t v) { return __uint_as_float(v) < INFINITY ? v : kNone; }
  static __device__ __forceinline__ uint32_t load(const float* p) { return decode(fetch(p)); }
};
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
  void* out;        // int32 / float, by OUT (may be null for EDT_OUT_DIST with SQ == 2)
  int32_t* out_sq;  // exact integer squared distance (EDT_OUT_DIST, SQ >= 1)
  int64_t inner;
  int64_t outer_stride;
  int64_t stride;
  int L;
  double sG;         // G(u) = sG * raw(u) + u^2          (float64 arithmetic only)
  double w2;         // squared spacing of this axis       (float64 arithmetic only)
  double out_scale;  // EDT_OUT_DIST: dist = sqrt(val * out_scale)
  float out_mul;     // (float)sqrt(out_scale)
};

constexpr int ENV_THREADS = 128;
// elements per software-pipelined input batch (template parameter ENV_BATCH of the kernel): 8 for long lines; 4
// for short ones (the z pass of a 100-plane stack: the rolled, bounds-checked tail then covers 4 planes instead of
// 12 -- measured 1.76 -> 1.67 ms on C2, while 16 costs 1.90 ms and the x pass is indifferent)
constexpr int ENV_BATCH_LONG = 8;
constexpr int ENV_BATCH_SHORT = 4;
#ifndef ENV_BANDED_MIN_BLOCKS
#define ENV_BANDED_MIN_BLOCKS 10
#endif

// Arithmetic of the hull tests.  With G(u) = F(u)/w2 + u^2, site b is hidden by (a, q) iff
//     (Gb - Ga) (q - b)  >=  (Gq - Gb) (b - a).
// The kernel keeps A = Gb - Ga and B = b - a of the two top entries as state, so a test costs one
// subtraction, two products and a compare, and a push re-uses the test's own operands as the
// next (A, B).
//  INT : unit weights, every quantity an integer < 2^31 -> int32 differences, int64 products
//        (one IMAD.WIDE each), no floating point at all.
//  F64 : float64 (exact whenever the operands are integers < 2^53; conversions through the
//        2^52 trick, never the I2F/F2I pipe).
template <bool INT>
struct Hull;

template <>
struct Hull<true> {
  using G = int32_t;   // G values and their differences
  using B = int32_t;   // index differences
  using W = long long;
  template <typename TIn>
  static __device__ __forceinline__ G make(uint32_t raw, int u, double, double) {
    return (G)(raw + (uint32_t)(u * u));
  }
  static __device__ __forceinline__ B idiff(int hi, int lo) { return hi - lo; }
  static __device__ __forceinline__ bool hidden(G A, B Bb, G dG, B dq) { return (W)A * (W)dq >= (W)dG * (W)Bb; }
};

template <>
struct Hull<false> {
  using G = double;
  using B = double;
  using W = double;
  template <typename TIn>
  static __device__ __forceinline__ G make(uint32_t raw, int u, double ud, double sG) {
    return fma(sG, edt_value<TIn>(raw), ud * ud);
  }
  static __device__ __forceinline__ B idiff(int hi, int lo) { return u2d((uint32_t)(hi - lo)); }
  static __device__ __forceinline__ bool hidden(G A, B Bb, G dG, B dq) { return A * dq >= dG * Bb; }
};

// OUT: EDT_OUT_I32 (squared, int32) / EDT_OUT_F32 (squared, float) / EDT_OUT_DIST (sqrt, float)
// SQ : EDT_OUT_DIST only -- 0 distances, 1 distances + int32 squares, 2 int32 squares only
// IdxT: element offsets inside a line bundle (uint32 when the volume has < 2^32 elements)
// BANDS: 1 = one thread per line.  2 = two threads per line (all-integer instances only): the line is cut in the
//        middle, the thread of the upper band works in MIRRORED coordinates (t' = L-1-u, negative memory stride), so
//        both threads scan from their outer end towards the cut and both hull stacks end with their top next to it.
//        With G(u) = F(u) + u^2 the hull is the lower convex hull of the points (u, G(u)) (adding the linear term of
//        the mirrored frame does not change it), so the two chains are joined by their lower common tangent: entries
//        are dropped from the two tops -- the only direction the in-place links allow, and the one the tangent
//        needs -- until neither top is hidden by its neighbour below and the other chain's top.  Both threads
//        run that (short) walk redundantly, then each fills from the tangent's crossing point outwards along its
//        own chain.  Twice the threads: the x pass of a 2048 x 2048 x 100 stack no longer runs as a single partial wave.
template <typename TIn, typename LinkT, int OUT, int SQ, bool INT, typename IdxT, int ENV_BATCH, int BANDS>
// (min-blocks 1 on the long-line instances lifts ptxas' own cap of 40 registers -- 48, no spills: x pass 1.45 ->
// 1.38 ms; the short-line instances are faster under the cap, 0 = unspecified)
__global__ void __launch_bounds__(ENV_THREADS, BANDS == 2 ? ENV_BANDED_MIN_BLOCKS : (ENV_BATCH == ENV_BATCH_LONG ? 1 : 0))
    edt_envelope_kernel(const EnvArgs<TIn, LinkT> a) {
  static_assert(BANDS == 1 || (BANDS == 2 && INT), "banded lines: all-integer instances only");
  using H = Hull<INT>;
  using GT = typename H::G;
  using BT = typename H::B;
  using WT = typename H::W;
  using R = EdtRaw<TIn>;
  // pointer walks along a line: unsigned element steps (one thread per line) / signed ones (the mirrored band)
  using StepT = typename std::conditional<BANDS == 1, IdxT, int64_t>::type;
  constexpr int LINES = ENV_THREADS / BANDS;  // lines per block
  const int band = BANDS == 1 ? 0 : (int)threadIdx.x / LINES;  // warp-uniform
  const int64_t i = (int64_t)blockIdx.x * LINES + (BANDS == 1 ? (int)threadIdx.x : (int)threadIdx.x % LINES);
  if (BANDS == 1 && i >= a.inner) return;
  const bool valid = i < a.inner;  // BANDS == 2: every thread reaches the barrier below
  const int Lfull = a.L;
  // this thread's share of the line, in its own frame: L sites, element 0 at `base`, step `stp`
  const int L = BANDS == 1 ? Lfull : (!valid ? 0 : (band == 0 ? Lfull / 2 : Lfull - Lfull / 2));
  const int64_t line0 = (int64_t)blockIdx.y * a.outer_stride + (valid ? i : 0);
  const int64_t base = line0 + (band ? (int64_t)(Lfull - 1) * a.stride : 0);
  const StepT stp = (StepT)(band ? -a.stride : a.stride);
  const TIn* __restrict__ in = a.in + base;
  // hull links and payloads are addressed as parameter base + one IdxT element index (a single
  // IMAD.WIDE per access when the volume has < 2^32 elements), not through 64-bit pointer arithmetic;
  // the mirrored band's negative step wraps modulo 2^32 there, the sums stay true element offsets
  const IdxT bi = (IdxT)base;
  const IdxT st = (IdxT)stp;
  const double sG = a.sG;

  // ---- forward sweep: build the hull as an in-place linked stack (link[q] = entry below q).
  // Registers hold the two top entries (t0 above t1) and A = G0 - G1, B = t0 - t1.
  // Input is software-pipelined: batch k+1 is in flight (registers) while batch k, parked in this
  // thread's own shared-memory column, is consumed by a rolled loop.
  __shared__ uint32_t s_raw[ENV_BATCH][ENV_THREADS];
  int t0 = EDT_NO_SITE, t1 = EDT_NO_SITE;
  GT G0 = 0, A = 0;
  BT Bb = 0;
  {
    uint32_t nb[ENV_BATCH];
#pragma unroll
    for (int u = 0; u < ENV_BATCH; ++u) nb[u] = u < L ? R::fetch(in + (StepT)u * stp) : R::kFetchNone;
    IdxT lidx = bi;
    const TIn* __restrict__ pnext = in + (StepT)ENV_BATCH * stp;
    double qd = 0.0;
    // one hull step for site q with payload rq (skipped for "no site")
    auto step = [&](uint32_t rq, int q) {
      if (rq != R::kNone) {
        const GT Gq = H::template make<TIn>(rq, q, qd, sG);
        GT dG = Gq - G0;
        BT dq = H::idiff(q, t0);
        if (t1 != EDT_NO_SITE && H::hidden(A, Bb, dG, dq)) {
          // pop until the top is visible again (rare: ~0.5 pops per voxel, usually one at a time)
          do {
            t0 = t1;
            G0 = G0 - A;  // = G1
            t1 = (int)a.link[bi + (IdxT)t0 * st];
            dG = Gq - G0;
            dq = H::idiff(q, t0);
            if (t1 == EDT_NO_SITE) break;
            const GT G1 = H::template make<TIn>(R::load(a.in + (bi + (IdxT)t1 * st)), t1, u2d((uint32_t)t1), sG);
            A = G0 - G1;
            Bb = H::idiff(t0, t1);
          } while (H::hidden(A, Bb, dG, dq));
        }
        a.link[lidx] = (LinkT)t0;
        t1 = t0;
        t0 = q;
        G0 = Gq;
        A = dG;
        Bb = dq;
      }
      lidx += st;
      qd += 1.0;
    };
    int q0 = 0;
#pragma unroll 1
    for (; q0 + 2 * ENV_BATCH <= L; q0 += ENV_BATCH) {  // full batch, and a full batch to prefetch
#pragma unroll
      for (int u = 0; u < ENV_BATCH; ++u) s_raw[u][threadIdx.x] = nb[u];
#pragma unroll
      for (int u = 0; u < ENV_BATCH; ++u) nb[u] = R::fetch(pnext + (StepT)u * stp);
      pnext += (StepT)ENV_BATCH * stp;
#pragma unroll 4
      for (int u = 0; u < ENV_BATCH; ++u) step(R::decode(s_raw[u][threadIdx.x]), q0 + u);
    }
#pragma unroll 1
    for (; q0 < L; q0 += ENV_BATCH) {  // the last one or two batches, bounds-checked
#pragma unroll
      for (int u = 0; u < ENV_BATCH; ++u) s_raw[u][threadIdx.x] = nb[u];
#pragma unroll
      for (int u = 0; u < ENV_BATCH; ++u)
        nb[u] = (q0 + ENV_BATCH + u < L) ? R::fetch(pnext + (StepT)u * stp) : R::kFetchNone;
      pnext += (StepT)ENV_BATCH * stp;
      const int qe = min(L, q0 + ENV_BATCH);
#pragma unroll 1
      for (int q = q0; q < qe; ++q) step(R::decode(s_raw[q - q0][threadIdx.x]), q);
    }
  }

  // ---- backward sweep over the hull.  Parabola u at p (divided by w2) is G(u) - 2 p u + p^2:
  //   D(p) = (Gprev - Gcur) + 2 p (cur - prev) = -A + p * 2B   prev at least as good  <=>  D(p) <= 0
  //   val(p) = F(cur) + w2 (p - cur)^2
  // output pointers walk backwards with p (element pstart first; L-1 when the thread owns the whole line)
  int pstart = L - 1;
  if constexpr (BANDS == 2) {
    // ---- join the two chains of the line (see BANDS above).  Everything below is in the line's own frame
    // (u = 0..Lfull-1, G(u) = F(u) + u^2); the upper band's entries are stored mirrored (u = Lfull-1 - t').
    __shared__ int s_top[4][ENV_THREADS];
    s_top[0][threadIdx.x] = t0;
    s_top[1][threadIdx.x] = t1;
    s_top[2][threadIdx.x] = (int)G0;
    s_top[3][threadIdx.x] = (int)(G0 - A);  // G of t1 (unused when t1 is absent)
    __syncthreads();
    const int other = (int)threadIdx.x ^ LINES;
    const int o0 = s_top[0][other];
    if (t0 == EDT_NO_SITE) {
      if (o0 != EDT_NO_SITE || !valid) return;  // the other band's thread fills the whole line
      // no site on the line at all: each thread writes its own band (the no-site loop below)
    } else if (o0 == EDT_NO_SITE) {
      pstart = Lfull - 1;  // the whole line from this band's chain
    } else {
      const int o1 = s_top[1][other], oG0 = s_top[2][other], oG1 = s_top[3][other];
      const int top = Lfull - 1;
      // l / lb: top of the lower band's chain and the entry below it; r / rb: the same for the upper band
      int l = band ? o0 : t0, lb = band ? o1 : t1;
      int rm = band ? t0 : o0, rbm = band ? t1 : o1;  // mirrored indices of the upper band
      int r = top - rm, rb = rbm == EDT_NO_SITE ? EDT_NO_SITE : top - rbm;
      // G in the line's frame: the lower band's is stored as it is; the upper band's G' = F + t'^2
      int32_t Gl = band ? oG0 : (int32_t)G0, Glb = band ? oG1 : (int32_t)(G0 - A);
      const int32_t Grm = band ? (int32_t)G0 : oG0, Grbm = band ? (int32_t)(G0 - A) : oG1;
      int32_t Gr = Grm - rm * rm + r * r, Grb = rb == EDT_NO_SITE ? 0 : Grbm - rbm * rbm + rb * rb;
      const IdxT l0 = (IdxT)line0, sa = (IdxT)a.stride;
      for (;;) {
        bool moved = false;
        // l is hidden by (lb, r)
        while (lb != EDT_NO_SITE && (long long)(Gl - Glb) * (r - l) >= (long long)(Gr - Gl) * (l - lb)) {
          l = lb;
          Gl = Glb;
          lb = (int)a.link[l0 + (IdxT)l * sa];
          if (lb != EDT_NO_SITE) Glb = (int32_t)(R::load(a.in + (l0 + (IdxT)lb * sa)) + (uint32_t)(lb * lb));
          moved = true;
        }
        // r is hidden by (l, rb)
        while (rb != EDT_NO_SITE && (long long)(Gr - Gl) * (rb - r) >= (long long)(Grb - Gr) * (r - l)) {
          r = rb;
          Gr = Grb;
          const int t = (int)a.link[l0 + (IdxT)r * sa];  // mirrored index of the entry below
          rb = t == EDT_NO_SITE ? EDT_NO_SITE : top - t;
          if (rb != EDT_NO_SITE) Grb = (int32_t)(R::load(a.in + (l0 + (IdxT)rb * sa)) + (uint32_t)(rb * rb));
          moved = true;
        }
        if (!moved) break;
      }
      // l is at least as good as r at p  <=>  2 p (r - l) <= Gr - Gl: the lower band's thread fills p <= m,
      // the upper band's thread p > m (ties go to l; the value is the same)
      const int num = Gr - Gl, den = 2 * (r - l);
      int m = num >= 0 ? num / den : -((den - 1 - num) / den);
      m = max(-1, min(m, top));
      if (band == 0) {
        pstart = m;
        t0 = l;
        t1 = lb;
        G0 = (GT)Gl;
        if (lb != EDT_NO_SITE) {
          A = (GT)(Gl - Glb);
          Bb = H::idiff(l, lb);
        }
      } else {
        pstart = top - (m + 1);
        t0 = top - r;
        t1 = rb == EDT_NO_SITE ? EDT_NO_SITE : top - rb;
        G0 = (GT)(Gr - r * r + t0 * t0);
        if (rb != EDT_NO_SITE) {
          A = (GT)((int32_t)G0 - (Grb - rb * rb + t1 * t1));
          Bb = H::idiff(t0, t1);
        }
      }
      if (pstart < 0) return;
    }
  }
  float* __restrict__ out_f = reinterpret_cast<float*>(a.out) + base + (StepT)pstart * stp;
  int32_t* __restrict__ out_i = reinterpret_cast<int32_t*>(a.out) + base + (StepT)pstart * stp;
  int32_t* __restrict__ out_sq = a.out_sq + base + (StepT)pstart * stp;
  if (t0 == EDT_NO_SITE) {  // no site on this line
#pragma unroll 1
    for (int p = pstart; p >= 0; --p, out_i -= stp, out_f -= stp, out_sq -= stp) {
      if (OUT == EDT_OUT_I32) *out_i = EDT_INF_I32;
      if (OUT == EDT_OUT_F32 || (OUT == EDT_OUT_DIST && SQ != 2)) *out_f = INFINITY;
      if (OUT == EDT_OUT_DIST && SQ != 0) *out_sq = EDT_INF_I32;
    }
    return;
  }
  int cur = t0, prev = t1;
  GT Gcur = G0;
  // one entry ahead: index, payload and link of the entry below prev are requested when prev is
  // entered and consumed only at the next hull move
  int nxt = EDT_NO_SITE, nxt_link = EDT_NO_SITE;
  uint32_t nxt_raw = 0;
  if (prev != EDT_NO_SITE) nxt = (int)a.link[bi + (IdxT)prev * st];
  if (nxt != EDT_NO_SITE) {
    nxt_raw = R::load(a.in + (bi + (IdxT)nxt * st));
    nxt_link = (int)a.link[bi + (IdxT)nxt * st];
  }
  const double w2s = a.w2 * a.out_scale;  // F64 + EDT_OUT_DIST: the output scale is folded in
  const double w2 = (OUT == EDT_OUT_DIST) ? w2s : a.w2;
  const float out_mul = a.out_mul;
  // INT: D and val advance incrementally (exact integers).  F64: both are re-evaluated from cached
  // per-pair constants with one fma each (an incremental sum would cancel catastrophically when
  // val falls from far-field magnitudes to ~0 along a run).
  WT D = 1, Dstep = 0, Gdiff = 1;
  int32_t val = 0, e = 0;
  double Fcur = 0.0, curd = 0.0;
  double pd = u2d((uint32_t)pstart);
  // float64 arithmetic decides which site owns p; the DISTANCE itself (EDT_OUT_DIST) is then evaluated
  // in float32 -- three roundings of 6e-8 before the square root and 1.2e-7 in it (MUFU), well inside
  // the 1e-6 contract of the anisotropic transform
  float Fcurf = 0.f, curf = 0.f, pf = (float)pstart;
  const float w2f = (float)w2;
  auto enter = [&](int p, GT Aa, BT Ba) {  // pair (prev, cur) at p;  Aa = Gcur - Gprev, Ba = cur - prev
    if (INT) {
      const int dp = p - cur;
      val = (int32_t)Gcur - cur * cur + dp * dp;
      e = 2 * dp - 1;
      D = 1;
      Dstep = 0;
      if (prev != EDT_NO_SITE) {
        Dstep = (WT)(2 * (int32_t)Ba);
        D = (WT)p * Dstep - (WT)(int32_t)Aa;
      }
    } else {
      curd = u2d((uint32_t)cur);
      Fcur = ((double)Gcur - curd * curd) * w2;
      Fcurf = (float)Fcur;
      curf = (float)cur;
      Dstep = 0;
      Gdiff = 1;  // D stays positive: no entry below
      if (prev != EDT_NO_SITE) {
        Dstep = (WT)((double)Ba + (double)Ba);
        Gdiff = (WT)(-(double)Aa);
      }
    }
  };
  enter(pstart, A, Bb);
#pragma unroll 4
  for (int p = pstart; p >= 0; --p, out_i -= stp, out_f -= stp, out_sq -= stp) {
    if (!INT) D = (WT)fma(pd, (double)Dstep, (double)Gdiff);
    if (D <= (WT)0) {
      do {  // move down the hull (D > 0 whenever prev does not exist)
        const GT Gprev = Gcur - A;
        cur = prev;
        Gcur = Gprev;
        prev = nxt;
        if (prev != EDT_NO_SITE) {
          const GT Gp = H::template make<TIn>(nxt_raw, prev, u2d((uint32_t)prev), sG);
          A = Gcur - Gp;
          Bb = H::idiff(cur, prev);
          nxt = nxt_link;
          if (nxt != EDT_NO_SITE) {
            nxt_raw = R::load(a.in + (bi + (IdxT)nxt * st));
            nxt_link = (int)a.link[bi + (IdxT)nxt * st];
          }
        }
        enter(p, A, Bb);
        if (!INT) D = (WT)fma(pd, (double)Dstep, (double)Gdiff);
      } while (D <= (WT)0);
    }
    if (INT) {
      if (OUT == EDT_OUT_I32) *out_i = val;
      if (OUT == EDT_OUT_F32) *out_f = (float)val;
      if (OUT == EDT_OUT_DIST) {
        // float32 sqrt of the (exact) integer: |rel err| < 1e-7 after the two roundings
        if (SQ != 2) *out_f = sqrtf((float)val) * out_mul;
        if (SQ != 0) *out_sq = val;
      }
      val -= e;
      e -= 2;
      D -= Dstep;
    } else {
      if (OUT == EDT_OUT_DIST) {
        const float dpf = pf - curf;
        *out_f = sqrt_approx(fmaf(w2f, dpf * dpf, Fcurf));  // w2 carries out_scale here
        pf -= 1.0f;
      } else {
        const double dpd = pd - curd;
        const double v = fma(w2, dpd * dpd, Fcur);  // squared distance
        if (OUT == EDT_OUT_I32) *out_i = d2i_exact(v);
        if (OUT == EDT_OUT_F32) *out_f = (float)v;
      }
      pd -= 1.0;
    }
  }
}

template <typename TIn, typename LinkT, int OUT, int SQ, bool INT>
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
  a.sG = sin / w2;
  a.w2 = w2;
  a.out_scale = out_scale;
  a.out_mul = (float)sqrt(out_scale);
  const bool small_idx = nz * nx * ny < (1ll << 32);
  // two threads per line (BANDS of the kernel) where one per line leaves the machine under-filled: the all-integer
  // x pass of a z-slab of a wide mosaic (16384 x 2048 x 25 per rank, 51 200 lines: 5.19 -> 3.79 ms measured; the
  // 204 800 lines of a 2048 x 2048 x 100 stack fill the SMs already and measure the same either way, 1.39 / 1.40 ms).
  // MFISH_EDT_BANDS=1 never / =2 wherever the instance exists (tests, A/B timing).
  bool banded = false;
  if constexpr (INT && std::is_same<TIn, uint16_t>::value) {
    const int forced = [] {
      const char* e = getenv("MFISH_EDT_BANDS");
      return e ? atoi(e) : 0;
    }();
    const int64_t lines = a.inner * nouter;
    banded = forced == 2 ? a.L >= 2 : (forced != 1 && a.L >= 512 && lines <= (int64_t)num_sms() * 1024);
  }
  const int lines_per_block = banded ? ENV_THREADS / 2 : ENV_THREADS;
  int64_t bx = (a.inner + lines_per_block - 1) / lines_per_block;
  MFISH_REQUIRE(bx < (1ll << 31) && nouter < 65536, "edt: grid too large");
  dim3 grid((unsigned)bx, (unsigned)nouter);
  ProfScope prof(axis == MFISH_AXIS_Z ? "edt_envelope_z" : "edt_envelope_x", st);
  if (banded) {
    if constexpr (INT && std::is_same<TIn, uint16_t>::value) {
      if (a.L <= 256) {
        if (small_idx)
          edt_envelope_kernel<TIn, LinkT, OUT, SQ, INT, uint32_t, ENV_BATCH_SHORT, 2><<<grid, ENV_THREADS, 0, st>>>(a);
        else
          edt_envelope_kernel<TIn, LinkT, OUT, SQ, INT, int64_t, ENV_BATCH_SHORT, 2><<<grid, ENV_THREADS, 0, st>>>(a);
      } else {
        if (small_idx)
          edt_envelope_kernel<TIn, LinkT, OUT, SQ, INT, uint32_t, ENV_BATCH_LONG, 2><<<grid, ENV_THREADS, 0, st>>>(a);
        else
          edt_envelope_kernel<TIn, LinkT, OUT, SQ, INT, int64_t, ENV_BATCH_LONG, 2><<<grid, ENV_THREADS, 0, st>>>(a);
      }
    }
  } else if (a.L <= 256) {
    if (small_idx)
      edt_envelope_kernel<TIn, LinkT, OUT, SQ, INT, uint32_t, ENV_BATCH_SHORT, 1><<<grid, ENV_THREADS, 0, st>>>(a);
    else
      edt_envelope_kernel<TIn, LinkT, OUT, SQ, INT, int64_t, ENV_BATCH_SHORT, 1><<<grid, ENV_THREADS, 0, st>>>(a);
  } else {
    if (small_idx)
      edt_envelope_kernel<TIn, LinkT, OUT, SQ, INT, uint32_t, ENV_BATCH_LONG, 1><<<grid, ENV_THREADS, 0, st>>>(a);
    else
      edt_envelope_kernel<TIn, LinkT, OUT, SQ, INT, int64_t, ENV_BATCH_LONG, 1><<<grid, ENV_THREADS, 0, st>>>(a);
  }
  count_launch(1);
  return check_launch("edt_envelope");
}

template <typename TIn, int OUT, int SQ, bool INT>
static int launch_env_l(const TIn* in, void* link, bool link16, void* out, int32_t* out_sq, int axis, int64_t nz,
                        int64_t nx, int64_t ny, double sin, double w2, double out_scale, cudaStream_t st) {
  if (link16)
    return launch_env<TIn, int16_t, OUT, SQ, INT>(in, link, out, out_sq, axis, nz, nx, ny, sin, w2, out_scale, st);
  return launch_env<TIn, int32_t, OUT, SQ, INT>(in, link, out, out_sq, axis, nz, nx, ny, sin, w2, out_scale, st);
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
// pass 1 alone: |dy| to the nearest zero voxel of every row (u16, 0xFFFF = none)
extern "C" int mfish_edt_rowscan_u8(const uint8_t* mask_dev, int64_t nrows, int64_t ny, uint16_t* dy_out_dev,
                                    void* stream) {
  MFISH_REQUIRE(mask_dev && dy_out_dev, "edt: null pointer");
  MFISH_REQUIRE(nrows > 0 && ny > 0, "edt: empty volume");
  return launch_scan_y(mask_dev, dy_out_dev, nrows, ny, 0, as_stream(stream));
}

// pass 2 alone, on row distances: lower envelope along x.  ny_full = length of the rows the distances were
// measured over (a y-slab of a wider mosaic holds |dy| up to ny_full - 1).  link_ws_dev: nz*nx*ny*4 bytes.
extern "C" int mfish_edt_xpass_u16(const uint16_t* dy_dev, int64_t nz, int64_t nx, int64_t ny, int64_t ny_full,
                                   const double* spacing_zxy_host, void* f2_out_dev, int* f2_kind_host,
                                   void* link_ws_dev, void* stream) {
  MFISH_REQUIRE(dy_dev && link_ws_dev && f2_out_dev && f2_kind_host, "edt: null pointer");
  MFISH_REQUIRE(nz > 0 && nx > 0 && ny > 0, "edt: empty volume");
  MFISH_REQUIRE(nz < (1 << 24) && nx < (1 << 24), "edt: dimension too large");
  if (ny_full < ny) ny_full = ny;
  MFISH_REQUIRE(ny_full < 65535, "edt: ny must be < 65535 (u16 row distances)");
  double wz, wx, wy;
  int rc = edt_spacing(spacing_zxy_host, wz, wx, wy);
  if (rc != MFISH_OK) return rc;
  const double plane_sq = (double)(nx - 1) * (nx - 1) + (double)(ny_full - 1) * (ny_full - 1);
  const bool plane_int = (wx == wy) && plane_sq < 2147483000.0;
  cudaStream_t st = as_stream(stream);
  const bool link16 = nx <= 32767;
  if (plane_int) {
    *f2_kind_host = MFISH_EDT_F2_I32;
    // all-integer hull tests need max(F) + L^2 = (ny-1)^2 + nx^2 below 2^31
    const bool int_arith = (double)(ny_full - 1) * (ny_full - 1) + (double)nx * nx < 2147483000.0;
    if (int_arith)
      return launch_env_l<uint16_t, EDT_OUT_I32, 0, true>(dy_dev, link_ws_dev, link16, f2_out_dev, nullptr,
                                                          MFISH_AXIS_X, nz, nx, ny, 1.0, 1.0, 1.0, st);
    return launch_env_l<uint16_t, EDT_OUT_I32, 0, false>(dy_dev, link_ws_dev, link16, f2_out_dev, nullptr,
                                                         MFISH_AXIS_X, nz, nx, ny, 1.0, 1.0, 1.0, st);
  }
  *f2_kind_host = MFISH_EDT_F2_F32;
  return launch_env_l<uint16_t, EDT_OUT_F32, 0, false>(dy_dev, link_ws_dev, link16, f2_out_dev, nullptr, MFISH_AXIS_X,
                                                       nz, nx, ny, wy * wy, wx * wx, 1.0, st);
}

extern "C" int mfish_edt_inplane_u8(const uint8_t* mask_dev, int64_t nz, int64_t nx, int64_t ny,
                                    const double* spacing_zxy_host, void* f2_out_dev, int* f2_kind_host,
                                    void* workspace_dev, void* stream) {
  MFISH_REQUIRE(mask_dev && workspace_dev && f2_out_dev && f2_kind_host, "edt: null pointer");
  MFISH_REQUIRE(nz > 0 && nx > 0 && ny > 0, "edt: empty volume");
  MFISH_REQUIRE(ny < 65535, "edt: ny must be < 65535 (u16 row distances)");
  const int64_t n = nz * nx * ny;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace_dev);
  uint16_t* dy = reinterpret_cast<uint16_t*>(ws);
  void* link = ws + align256(n * 2);
  int rc = launch_scan_y(mask_dev, dy, nz * nx, ny, 0, as_stream(stream));
  if (rc != MFISH_OK) return rc;
  return mfish_edt_xpass_u16(dy, nz, nx, ny, ny, spacing_zxy_host, f2_out_dev, f2_kind_host, link, stream);
}

// pass 3 along z.  link_ws_dev: nz*nx*ny*4 bytes of scratch.
extern "C" int mfish_edt_zpass(const void* f2_dev, int f2_kind, int64_t nz, int64_t nx, int64_t ny,
                               const double* spacing_zxy_host, int64_t f2_bound, float* out_dist_dev,
                               int32_t* out_sq_dev, void* link_ws_dev, void* stream) {
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
  // largest finite in-plane value: the caller's bound (x-slabs of a wider mosaic), else this volume's
  const double plane_max =
      f2_bound > 0 ? (double)f2_bound : (double)(nx - 1) * (nx - 1) + (double)(ny - 1) * (ny - 1);
  if (f2_kind == MFISH_EDT_F2_I32) {
    const int32_t* f2 = reinterpret_cast<const int32_t*>(f2_dev);
    if (iso) {  // keep integers, scale once at the end: dist = sqrt(sq) * w
      const double gmax = plane_max + (double)nz * nz;
      const double sc = wz * wz;
      if (gmax < 2147483000.0) {
        if (out_dist_dev && out_sq_dev)
          return launch_env_l<int32_t, EDT_OUT_DIST, 1, true>(f2, link_ws_dev, link16, out_dist_dev, out_sq_dev,
                                                              MFISH_AXIS_Z, nz, nx, ny, 1.0, 1.0, sc, st);
        if (out_sq_dev)
          return launch_env_l<int32_t, EDT_OUT_DIST, 2, true>(f2, link_ws_dev, link16, nullptr, out_sq_dev,
                                                              MFISH_AXIS_Z, nz, nx, ny, 1.0, 1.0, sc, st);
        return launch_env_l<int32_t, EDT_OUT_DIST, 0, true>(f2, link_ws_dev, link16, out_dist_dev, nullptr,
                                                            MFISH_AXIS_Z, nz, nx, ny, 1.0, 1.0, sc, st);
      }
      MFISH_REQUIRE(!out_sq_dev, "edt: volume too large for int32 squared distances");
      return launch_env_l<int32_t, EDT_OUT_DIST, 0, false>(f2, link_ws_dev, link16, out_dist_dev, nullptr,
                                                           MFISH_AXIS_Z, nz, nx, ny, 1.0, 1.0, sc, st);
    }
    return launch_env_l<int32_t, EDT_OUT_DIST, 0, false>(f2, link_ws_dev, link16, out_dist_dev, nullptr, MFISH_AXIS_Z,
                                                         nz, nx, ny, wx * wx, wz * wz, 1.0, st);
  }
  return launch_env_l<float, EDT_OUT_DIST, 0, false>(reinterpret_cast<const float*>(f2_dev), link_ws_dev, link16,
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
  return mfish_edt_zpass(f2, kind, nz, nx, ny, spacing_zxy_host, 0, out_dist_dev, out_sq_dev, link, stream);
}
