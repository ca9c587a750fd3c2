// Separable 1-D correlation passes (orders 0 and 2 of the sampled Gaussian), the arithmetic
// behind scipy.ndimage.gaussian_filter / gaussian_laplace as skimage.feature.blob_log and
// skimage.filters.gaussian drive them (modules/muscle_fish.py:333,357,661).
//
// Device layout [z][x][y], y contiguous.
//  * conv_y_kernel          : contiguous axis.  A 512-voxel row segment + halo is staged in shared
//                             memory by the TMA unit (cp.async.bulk + mbarrier, double-buffered over
//                             4 rows per block; row-end segments copy their in-range part and patch
//                             the <= R boundary-mapped voxels); each thread produces 4 consecutive
//                             outputs from a register window of aligned and one-shifted float2 pairs,
//                             every tap a packed FFMA2 / FADD2.   R4 + W4 (G) or R4 + W8 (G,D) B/voxel.
//  * conv_strided_tma_kernel: z or x axis, R <= 8, rows a multiple of 4 floats.  A block owns a
//                             256-float row segment and walks the axis; rows arrive through a
//                             shared-memory stage ring filled by TMA bulk copies (full / empty
//                             mbarriers), each thread keeps the 2R+1 window as float2 column pairs.
//                             Every voxel is read once and written once: 8 / 12 / 16 / 12 B/voxel
//                             for ops G / GD / MID / LAST.  LAST can also flag outputs above a
//                             threshold (fused NMS candidates).
//  * conv_strided_kernel    : the same passes with a (2R+1+D)-deep register ring per thread (D planes
//                             of load look-ahead): any radius <= 16, any row length.
//  * *_generic kernels      : gather fallbacks (radius > 16, lines shorter than the radius).
// Taps live in the kernel parameter block (constant bank / uniform registers).
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <vector>
#include <type_traits>

#include "common.cuh"

namespace mfish {
void count_launch(int k);

template <int R>
struct Taps {
  float g[R + 1];
  float d[R + 1];
};

struct NormSrc {
  const float* mm;  // device {min, max} or nullptr
};

__device__ __forceinline__ void load_norm(const float* mm, float& mn, float& sc, bool& degenerate) {
  mn = 0.f;
  sc = 1.f;
  degenerate = false;
  if (mm != nullptr) {
    float lo = __ldg(mm), hi = __ldg(mm + 1);
    mn = lo;
    if (hi > lo)
      sc = (float)(1.0 / ((double)hi - (double)lo));
    else
      degenerate = true;
  }
}

// packed fp32 pairs (SASS: FFMA2 / FMUL2 / FADD2; the scalar operand is broadcast)
__device__ __forceinline__ float2 fma2(float a, float2 b, float2 c) {
  float2 aa = make_float2(a, a);
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&aa);
  unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
  unsigned long long rc = *reinterpret_cast<unsigned long long*>(&c);
  unsigned long long rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 mul2(float a, float2 b) {
  float2 aa = make_float2(a, a);
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&aa);
  unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
  unsigned long long rd;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
  unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
  unsigned long long rd;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}

// --------------------------------------------------------------------------------------------
// contiguous (y) axis
// --------------------------------------------------------------------------------------------
constexpr int YSEG = 512;     // outputs per block
constexpr int YTHREADS = 128;  // 4 outputs per thread

template <int R>
struct YArgs {
  const float* in;
  float* out0;
  float* out1;
  int64_t nrows;  // nz*nx
  int ny;
  int nseg;
  int bmode;
  const float* mm;
  Taps<R> t;
};

constexpr int YROWS = 4;  // consecutive rows handled by one block (double-buffered stages)

template <int R, int OP>
__global__ void __launch_bounds__(YTHREADS) conv_y_kernel(const YArgs<R> a) {
  constexpr int RP = (R + 3) & ~3;  // halo rounded up to a float4
  constexpr int NW = 4 + 2 * RP;    // register window per thread
  __shared__ __align__(128) float buf[2][YSEG + 2 * RP];
  __shared__ __align__(8) unsigned long long bar[2];

  const int64_t rblk = blockIdx.x / a.nseg;
  const int seg0 = (int)(blockIdx.x - rblk * a.nseg) * YSEG;
  const int64_t row_begin = rblk * YROWS;
  const int nrow = (int)min((int64_t)YROWS, a.nrows - row_begin);
  const int ny = a.ny;
  const int t = threadIdx.x;

  float mn, sc;
  bool degen;
  load_norm(a.mm, mn, sc, degen);

  // Rows whose length is a multiple of 4: the in-range part of the segment + halo (raw voxels) is
  // fetched by the TMA unit, one bulk copy per row, the copy of row i+1 in flight while row i is
  // convolved; the at most R out-of-range positions on either side of an edge segment are filled
  // through the boundary map by the first R threads.
  const bool tma = ((ny & 3) == 0) && aligned16_dev(a.in);
  const int lo = max(seg0 - RP, 0), hi = min(seg0 + YSEG + RP, ny);
  const unsigned bytes = (unsigned)(hi - lo) * (unsigned)sizeof(float);
  const int dst_off = lo - (seg0 - RP);
  const bool edge_lo = seg0 - RP < 0, edge_hi = seg0 + YSEG + RP > ny;
  const float* __restrict__ row0 = a.in + row_begin * ny;
  auto fill_halo = [&](float* stage, const float* __restrict__ src) {
    if (t < R) {
      if (edge_lo) {
        const int y = -1 - t;
        const int yy = a.bmode == MFISH_BOUNDARY_NEAREST ? bmap<MFISH_BOUNDARY_NEAREST>(y, ny)
                                                         : bmap<MFISH_BOUNDARY_REFLECT>(y, ny);
        stage[y - (seg0 - RP)] = __ldg(&src[yy]);
      }
      if (edge_hi) {
        const int y = ny + t;
        const int j = y - (seg0 - RP);
        if (j < YSEG + 2 * RP) {
          const int yy = a.bmode == MFISH_BOUNDARY_NEAREST ? bmap<MFISH_BOUNDARY_NEAREST>(y, ny)
                                                           : bmap<MFISH_BOUNDARY_REFLECT>(y, ny);
          stage[j] = __ldg(&src[yy]);
        }
      }
    }
  };
  if (tma) {
    if (t == 0) {
      mbar_init(&bar[0], 1);
      mbar_init(&bar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fill_halo(buf[0], row0);
    __syncthreads();
    if (t == 0) {
      mbar_expect_tx(&bar[0], bytes);
      tma_load_1d(buf[0] + dst_off, row0 + lo, bytes, &bar[0]);
    }
  }
  const int y0 = seg0 + 4 * t;
  const bool vst = ((ny & 3) == 0) && aligned16_dev(a.out0) && (OP != MFISH_CONV_GD || aligned16_dev(a.out1));

  for (int i = 0; i < nrow; ++i) {
    const int64_t row = row_begin + i;
    float* cur = buf[i & 1];
    if (tma) {
      if (i + 1 < nrow) {  // the other stage was released by the barrier that ended row i-1
        const float* __restrict__ nsrc = row0 + (int64_t)(i + 1) * ny;
        if (t == 0) {
          mbar_expect_tx(&bar[(i + 1) & 1], bytes);
          tma_load_1d(buf[(i + 1) & 1] + dst_off, nsrc + lo, bytes, &bar[(i + 1) & 1]);
        }
        fill_halo(buf[(i + 1) & 1], nsrc);
      }
      mbar_wait(&bar[i & 1], (unsigned)((i >> 1) & 1));
    } else {
      const float* __restrict__ src = a.in + row * ny;
      for (int j = t; j < YSEG + 2 * RP; j += YTHREADS) {
        int y = seg0 - RP + j;
        float v = 0.f;
        // positions further than R outside the row are never read by a valid output
        if (y >= -R && y < ny + R) {
          int yy = a.bmode == MFISH_BOUNDARY_NEAREST ? bmap<MFISH_BOUNDARY_NEAREST>(y, ny)
                                                     : bmap<MFISH_BOUNDARY_REFLECT>(y, ny);
          v = __ldg(&src[yy]);
        }
        cur[j] = v;
      }
      __syncthreads();
    }
    if (y0 < ny) {
      // register window as aligned pairs e[j] = (w[2j], w[2j+1]) and the shifted pairs
      // o[j] = (w[2j+1], w[2j+2]); su.normalize_image is applied here, on load
      float2 e[NW / 2], o[NW / 2 - 1];
      const float4* b4 = reinterpret_cast<const float4*>(cur + 4 * t);
      const float2 nmn = make_float2(-mn, -mn);
#pragma unroll
      for (int j = 0; j < NW / 4; ++j) {
        float4 v = b4[j];
        e[2 * j + 0] = mul2(sc, add2(make_float2(v.x, v.y), nmn));
        e[2 * j + 1] = mul2(sc, add2(make_float2(v.z, v.w), nmn));
      }
      if (degen) {
#pragma unroll
        for (int j = 0; j < NW / 2; ++j) e[j] = make_float2(1.0f, 1.0f);
      }
#pragma unroll
      for (int j = 0; j < NW / 2 - 1; ++j) o[j] = make_float2(e[j].y, e[j + 1].x);
      auto pair_at = [&](int idx) -> float2 { return (idx & 1) ? o[idx >> 1] : e[idx >> 1]; };
      float2 o0[2], o1[2];
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int c = 2 * p + RP;
        float2 acc0 = mul2(a.t.g[0], pair_at(c));
        float2 acc1 = (OP == MFISH_CONV_GD) ? mul2(a.t.d[0], pair_at(c)) : make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 1; k <= R; ++k) {
          const float2 sN = add2(pair_at(c - k), pair_at(c + k));
          acc0 = fma2(a.t.g[k], sN, acc0);
          if (OP == MFISH_CONV_GD) acc1 = fma2(a.t.d[k], sN, acc1);
        }
        o0[p] = acc0;
        o1[p] = acc1;
      }
      float* __restrict__ d0 = a.out0 + row * ny + y0;
      if (vst && y0 + 3 < ny) {
        *reinterpret_cast<float4*>(d0) = make_float4(o0[0].x, o0[0].y, o0[1].x, o0[1].y);
        if (OP == MFISH_CONV_GD)
          *reinterpret_cast<float4*>(a.out1 + row * ny + y0) = make_float4(o1[0].x, o1[0].y, o1[1].x, o1[1].y);
      } else {
        const float r0[4] = {o0[0].x, o0[0].y, o0[1].x, o0[1].y};
        const float r1[4] = {o1[0].x, o1[0].y, o1[1].x, o1[1].y};
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          if (y0 + k4 < ny) {
            d0[k4] = r0[k4];
            if (OP == MFISH_CONV_GD) a.out1[row * ny + y0 + k4] = r1[k4];
          }
        }
      }
    }
    __syncthreads();  // every thread is done with this stage
  }
}

// runtime-radius fallback (large sigma: segmentation blurs, FMA-bound by construction)
struct YGenArgs {
  const float* in;
  float* out0;
  float* out1;
  int64_t nrows;
  int ny, nseg, bmode, radius, op;
  const float* mm;
  const float* taps_g;  // device, radius+1
  const float* taps_d;
};

__global__ void __launch_bounds__(YTHREADS) conv_y_generic_kernel(const YGenArgs a) {
  extern __shared__ float gbuf[];
  const int R = a.radius;
  float* tg = gbuf;               // R+1
  float* td = gbuf + (R + 1);     // R+1
  float* buf = gbuf + 2 * (R + 1);  // YSEG + 2R
  const int64_t row = blockIdx.x / a.nseg;
  const int seg0 = (int)(blockIdx.x - row * a.nseg) * YSEG;
  const int ny = a.ny;
  const float* __restrict__ src = a.in + row * ny;
  float mn, sc;
  bool degen;
  load_norm(a.mm, mn, sc, degen);
  for (int j = threadIdx.x; j <= R; j += YTHREADS) {
    tg[j] = a.taps_g[j];
    td[j] = a.op == MFISH_CONV_GD ? a.taps_d[j] : 0.f;
  }
  for (int j = threadIdx.x; j < YSEG + 2 * R; j += YTHREADS) {
    int y = seg0 - R + j;
    float v = 0.f;
    if (y < ny + R) {
      int yy = a.bmode == MFISH_BOUNDARY_NEAREST ? bmap<MFISH_BOUNDARY_NEAREST>(y, ny)
                                                 : bmap<MFISH_BOUNDARY_REFLECT>(y, ny);
      v = degen ? 1.0f : (__ldg(&src[yy]) - mn) * sc;
    }
    buf[j] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < YSEG; i += YTHREADS) {
    int y = seg0 + i;
    if (y >= ny) break;
    const float* c = buf + i + R;
    float acc0 = tg[0] * c[0], acc1 = td[0] * c[0];
    for (int k = 1; k <= R; ++k) {
      float s = c[-k] + c[k];
      acc0 = fmaf(tg[k], s, acc0);
      acc1 = fmaf(td[k], s, acc1);
    }
    a.out0[row * ny + y] = acc0;
    if (a.op == MFISH_CONV_GD) a.out1[row * ny + y] = acc1;
  }
}

// --------------------------------------------------------------------------------------------
// strided (z or x) axis: register-ring walker
// --------------------------------------------------------------------------------------------
// planes of load look-ahead in the register ring.  The two-input z pass of the LoG (MID) runs best with
// 8 planes in flight at 4 blocks/SM (1.63 ms against 1.81 ms with 6 planes on the 2048x2048x100 stack:
// fewer partial, boundary-mapped sweeps on a 100-plane line and more bytes in flight per thread).
template <int OP>
struct StridedTune {
  static constexpr int kLookahead = (OP == MFISH_CONV_MID) ? 8 : 6;
  static constexpr int kMinBlocks = (OP == MFISH_CONV_MID) ? 4 : 5;
};
constexpr int STHREADS = 128;

template <int R>
struct SArgs {
  const float* in0;
  const float* in1;
  float* out0;
  float* out1;
  int64_t inner;         // contiguous positions per line bundle
  int64_t outer_stride;  // elements between bundles
  int stride;            // elements between successive samples along the axis (< 2^31)
  int L;                 // samples along the axis (>= R)
  int out_begin;         // outputs are produced for [out_begin, out_end) only (inputs: the whole line)
  int out_end;
  int chunk_len;         // samples handled per blockIdx.z
  int bmode;
  float scale;
  // optional (OP == LAST only): outputs above cand_thr that are not beaten by their two neighbours along
  // the walked axis are appended (element offset) to cand[]; cand_count may exceed cand_cap (overflow)
  float cand_thr;
  uint32_t* cand;
  uint32_t* cand_count;
  uint32_t cand_cap;
  Taps<R> t;
};

// All four volumes are addressed with one 32-bit element offset (volume < 2^32 elements), so an
// access costs a single IMAD.WIDE.U32 on top of the load/store.
__device__ __forceinline__ int bmap1_rt(int i, int n, int bmode) {
  return bmode == MFISH_BOUNDARY_NEAREST ? bmap1<MFISH_BOUNDARY_NEAREST>(i, n)
                                         : bmap1<MFISH_BOUNDARY_REFLECT>(i, n);
}

template <int R, int OP, bool EMIT = false>
__global__ void __launch_bounds__(STHREADS, StridedTune<OP>::kMinBlocks) conv_strided_kernel(const SArgs<R> a) {
  constexpr int LOOKAHEAD = StridedTune<OP>::kLookahead;
  constexpr int W = 2 * R + 1;
  constexpr int RS = W + LOOKAHEAD;
  constexpr bool TWO_IN = (OP == MFISH_CONV_MID || OP == MFISH_CONV_LAST);

  const int64_t i = (int64_t)blockIdx.x * STHREADS + threadIdx.x;
  if (i >= a.inner) return;
  const unsigned i0 = (unsigned)((int64_t)blockIdx.y * a.outer_stride + i);
  const int kb = a.out_begin + blockIdx.z * a.chunk_len;
  const int ke = min(a.out_end, kb + a.chunk_len);
  if (kb >= ke) return;
  const int L = a.L;
  const int st = a.stride;
  const int bmode = a.bmode;
  const float* __restrict__ p0 = a.in0;
  const float* __restrict__ p1 = a.in1;
  float* __restrict__ q0 = a.out0;
  float* __restrict__ q1 = a.out1;
  const int klast = ke - 1 + R;  // last plane any output of this chunk needs

  // candidate emission (fused threshold + 1-D non-max test of the LoG response, see SArgs)
  constexpr bool emit = EMIT && (OP == MFISH_CONV_LAST);
  float vm1 = -INFINITY, vm2 = -INFINITY;  // the two previous outputs of this thread
  auto emit_cand = [&](unsigned off, float v) {
    const unsigned active = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(active) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(a.cand_count, (uint32_t)__popc(active));
    base = __shfl_sync(active, base, leader);
    const uint32_t pos = base + __popc(active & ((1u << lane) - 1));
    if (pos < a.cand_cap) a.cand[pos] = off;
    (void)v;
  };

  float r0[RS];
  float r1[TWO_IN ? RS : 1];
#pragma unroll
  for (int s = 0; s < RS - 1; ++s) {
    int k = kb - R + s;
    r0[s] = 0.f;
    if (TWO_IN) r1[s] = 0.f;
    if (k <= klast) {
      const unsigned on = i0 + (unsigned)(bmap1_rt(k, L, bmode) * st);
      r0[s] = __ldg(p0 + on);
      if (TWO_IN) r1[s] = __ldg(p1 + on);
    }
  }
  r0[RS - 1] = 0.f;
  if (TWO_IN) r1[RS - 1] = 0.f;

  // one unrolled sweep of RS outputs; FAST = every output and every look-ahead plane is interior
  auto sweep = [&](auto fast_tag, int k0, unsigned off_k) {
    constexpr bool FAST = decltype(fast_tag)::value;
    const float* __restrict__ p0n = p0 + (unsigned)((R + LOOKAHEAD) * st);
    const float* __restrict__ p1n = TWO_IN ? p1 + (unsigned)((R + LOOKAHEAD) * st) : nullptr;
    const unsigned off_first = off_k;
    unsigned cmask = 0;  // bit u: the output BEFORE output u of this sweep is a candidate
#pragma unroll
    for (int u = 0; u < RS; ++u) {
      const int k = k0 + u;
      if (FAST || k < ke) {
        if (FAST) {
          r0[(u + RS - 1) % RS] = __ldg(p0n + off_k);
          if (TWO_IN) r1[(u + RS - 1) % RS] = __ldg(p1n + off_k);
        } else {
          const int kn = k + R + LOOKAHEAD;
          if (kn <= klast) {
            const unsigned on = i0 + (unsigned)(bmap1_rt(kn, L, bmode) * st);
            r0[(u + RS - 1) % RS] = __ldg(p0 + on);
            if (TWO_IN) r1[(u + RS - 1) % RS] = __ldg(p1 + on);
          }
        }
        const float c0 = r0[(u + R) % RS];
        float accG0 = a.t.g[0] * c0;  // G * in0
        float accD0 = a.t.d[0] * c0;  // D * in0
        float accG1 = 0.f;            // G * in1
        if (TWO_IN) accG1 = a.t.g[0] * r1[(u + R) % RS];
#pragma unroll
        for (int j = 1; j <= R; ++j) {
          const float s0 = r0[(u + R - j + RS) % RS] + r0[(u + R + j) % RS];
          if (OP != MFISH_CONV_LAST) accG0 = fmaf(a.t.g[j], s0, accG0);
          if (OP != MFISH_CONV_G) accD0 = fmaf(a.t.d[j], s0, accD0);
          if (TWO_IN) {
            const float s1 = r1[(u + R - j + RS) % RS] + r1[(u + R + j) % RS];
            accG1 = fmaf(a.t.g[j], s1, accG1);
          }
        }
        if (OP == MFISH_CONV_G) {
          q0[off_k] = accG0;
        } else if (OP == MFISH_CONV_GD) {
          q0[off_k] = accG0;
          q1[off_k] = accD0;
        } else if (OP == MFISH_CONV_MID) {
          q0[off_k] = accG0;
          q1[off_k] = accD0 + accG1;
        } else {
          const float v = a.scale * (accD0 + accG1);
          q0[off_k] = v;
          if (emit) {
            cmask |= (unsigned)(vm1 > a.cand_thr && vm1 >= vm2 && vm1 >= v) << u;
            vm2 = vm1;
            vm1 = v;
          }
        }
        off_k += (unsigned)st;
      }
    }
    if (emit) {
      while (cmask) {  // rare: a few candidates per thousand voxels
        const int u = __ffs(cmask) - 1;
        cmask &= cmask - 1;
        emit_cand(off_first + (unsigned)((u - 1) * st), 0.f);
      }
    }
  };

  unsigned off_k = i0 + (unsigned)(kb * st);
  for (int k0 = kb; k0 < ke; k0 += RS) {
    const bool fast = (k0 + RS <= ke) && (k0 + RS - 1 + R + LOOKAHEAD < L);
    if (fast)
      sweep(std::true_type{}, k0, off_k);
    else
      sweep(std::false_type{}, k0, off_k);
    off_k += (unsigned)(RS * st);
  }
  // the last output of the chunk: its successor belongs to another thread block (or does not exist)
  if (emit && vm1 > a.cand_thr && vm1 >= vm2) emit_cand(i0 + (unsigned)((ke - 1) * st), vm1);
}

struct SGenArgs {
  const float* in0;
  const float* in1;
  float* out0;
  float* out1;
  int64_t inner, stride, outer_stride;
  int L, bmode, radius, op;
  int out_begin;
  float scale;
  const float* taps_g;
  const float* taps_d;
};

// runtime-radius fallback: direct gather through L1/L2.
__global__ void __launch_bounds__(STHREADS) conv_strided_generic_kernel(const SGenArgs a) {
  extern __shared__ float tapbuf[];
  const int R = a.radius;
  float* tg = tapbuf;
  float* td = tapbuf + (R + 1);
  for (int j = threadIdx.x; j <= R; j += STHREADS) {
    tg[j] = a.taps_g[j];
    td[j] = a.taps_d ? a.taps_d[j] : 0.f;
  }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * STHREADS + threadIdx.x;
  if (i >= a.inner) return;
  const int k = a.out_begin + blockIdx.z;
  const int64_t base = (int64_t)blockIdx.y * a.outer_stride + i;
  const float* __restrict__ p0 = a.in0 + base;
  const float* __restrict__ p1 = a.in1 ? a.in1 + base : nullptr;
  const bool two_in = (a.op == MFISH_CONV_MID || a.op == MFISH_CONV_LAST);
  float accG0 = 0.f, accD0 = 0.f, accG1 = 0.f;
  for (int j = -R; j <= R; ++j) {
    int kk = a.bmode == MFISH_BOUNDARY_NEAREST ? bmap<MFISH_BOUNDARY_NEAREST>(k + j, a.L)
                                               : bmap<MFISH_BOUNDARY_REFLECT>(k + j, a.L);
    int aj = j < 0 ? -j : j;
    float v0 = __ldg(p0 + (int64_t)kk * a.stride);
    accG0 = fmaf(tg[aj], v0, accG0);
    accD0 = fmaf(td[aj], v0, accD0);
    if (two_in) accG1 = fmaf(tg[aj], __ldg(p1 + (int64_t)kk * a.stride), accG1);
  }
  const int64_t o = base + (int64_t)k * a.stride;
  switch (a.op) {
    case MFISH_CONV_G: a.out0[o] = accG0; break;
    case MFISH_CONV_GD: a.out0[o] = accG0; a.out1[o] = accD0; break;
    case MFISH_CONV_MID: a.out0[o] = accG0; a.out1[o] = accD0 + accG1; break;
    default: a.out0[o] = a.scale * (accD0 + accG1); break;
  }
}

// --------------------------------------------------------------------------------------------
// fused y + z pass of the LoG: in -> (C = Gz Gy in, F = (Dz Gy + Gz Dy) in), R4 + W8 bytes per voxel
// (the separate y and z passes move 12 + 16).
//
// One block owns a 512-voxel y segment of one x row and walks z.  Each input plane's row segment
// (+ halo) is brought into shared memory by the TMA unit (cp.async.bulk + mbarrier, YZ_STAGES
// planes in flight, issued by one elected thread), normalised in place, convolved along y from a
// register window (2 outputs per thread), and pushed into a (2R+1)-deep register ring of packed
// float2 values; the z convolution runs on the ring with packed FFMA2 / FADD2.
// --------------------------------------------------------------------------------------------
constexpr int YZ_THREADS = 256;
constexpr int YZ_SEG = 2 * YZ_THREADS;
constexpr int YZ_STAGES = 8;

// --------------------------------------------------------------------------------------------
// strided (z or x) axis, TMA-staged walker (R <= 8, rows a multiple of 4 floats)
// --------------------------------------------------------------------------------------------
// A block owns a 256-float row segment and walks it along the axis.  One elected thread keeps
// NST-2 rows in flight with bulk copies into a shared-memory stage ring (full/empty mbarriers per
// stage; boundary handling is nothing but the source row of the copy), so the depth of the
// prefetch costs shared memory instead of registers: each thread keeps only the 2R+1 window, as
// float2 column pairs, and all tap arithmetic is packed (FADD2/FFMA2).  Same summation order as
// conv_strided_kernel, hence bit-identical results.
constexpr int TSEG = 256;
constexpr int TTHREADS = 128;
constexpr int TWARPS = TTHREADS / 32;
constexpr int TMA_MAX_R = 8;
constexpr int TSLACK = 3;  // a stage is refilled TSLACK steps after it was consumed
template <int R, int OP>
struct TmaTune {
  static constexpr bool kTwoIn = (OP == MFISH_CONV_MID || OP == MFISH_CONV_LAST);
  static constexpr int kW = 2 * R + 1;
  // a multiple of the window length, so that one unrolled sweep of kStages steps sees every ring
  // slot AND every stage at a compile-time index
  static constexpr int kStages = kW >= 12 ? kW : kW * ((12 + kW - 1) / kW);
  static constexpr int kMinBlocks = (OP == MFISH_CONV_LAST) ? 4 : 5;
  static constexpr size_t kSmem = (size_t)kStages * (kTwoIn ? 2 : 1) * TSEG * sizeof(float) + 2 * kStages * 8;
};

__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __noinline__ void tma_issue_row(const float* in0, const float* in1, unsigned base, int k, int L, int bmode,
                                           int st, unsigned dst_smem, unsigned bar_smem, unsigned bytes) {
  if ((unsigned)k >= (unsigned)L) k = bmap1_rt(k, L, bmode);
  const size_t off = (size_t)base + (size_t)k * (size_t)st;
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_smem),
               "r"(in1 != nullptr ? 2u * bytes : bytes)
               : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(in0 + off), "r"(bytes), "r"(bar_smem)
               : "memory");
  if (in1 != nullptr)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     dst_smem + (unsigned)(TSEG * sizeof(float))),
                 "l"(in1 + off), "r"(bytes), "r"(bar_smem)
                 : "memory");
}

// Rare path of the fused candidate emission (a few voxels per thousand are above the threshold), kept
// out of line so that it costs the walker no registers: 1-D non-max test against the two neighbours
// along the walked axis, re-read from the calling thread's own earlier stores.  A neighbour that is
// not written yet (next sweep / another block) counts as passed: the verify kernel applies the full
// 26-neighbour rule to every candidate anyway.
__device__ __noinline__ void emit_sweep_candidates(const float* out0, unsigned off_first, unsigned st, unsigned cm,
                                                   int o0, int nout, int nst, float thr, uint32_t* cand,
                                                   uint32_t* cand_count, uint32_t cand_cap) {
  auto emit_cand = [&](unsigned off) {
    const unsigned active = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(active) - 1;
    uint32_t b = 0;
    if (lane == leader) b = atomicAdd(cand_count, (uint32_t)__popc(active));
    b = __shfl_sync(active, b, leader);
    const uint32_t pos = b + __popc(active & ((1u << lane) - 1));
    if (pos < cand_cap) cand[pos] = off;
  };
  while (cm) {
    const int u = __ffs(cm) - 1;
    cm &= cm - 1;
    const unsigned off = off_first + (unsigned)u * st;
    const float2 v = *reinterpret_cast<const float2*>(out0 + off);
    float2 pv = make_float2(-INFINITY, -INFINITY), nx = pv;
    if (o0 + u > 0) pv = *reinterpret_cast<const float2*>(out0 + (off - st));
    if (u + 1 < nst && o0 + u + 1 < nout) nx = *reinterpret_cast<const float2*>(out0 + (off + st));
    if (v.x > thr && v.x >= pv.x && v.x >= nx.x) emit_cand(off);
    if (v.y > thr && v.y >= pv.y && v.y >= nx.y) emit_cand(off + 1u);
  }
}

template <int R, int OP, bool EMIT = false>
__global__ void __launch_bounds__(TTHREADS, TmaTune<R, OP>::kMinBlocks) conv_strided_tma_kernel(const SArgs<R> a) {
  constexpr int W = 2 * R + 1;
  constexpr bool TWO_IN = TmaTune<R, OP>::kTwoIn;
  constexpr int NIN = TWO_IN ? 2 : 1;
  constexpr int NST = TmaTune<R, OP>::kStages;
  constexpr int LEAD = NST - TSLACK;  // rows in flight ahead of the consumers
  constexpr unsigned STAGE_BYTES = NIN * TSEG * sizeof(float);
  static_assert(NST % W == 0 && 2 * R < NST && LEAD >= 4, "stage ring geometry");
  extern __shared__ __align__(128) unsigned char tsm[];
  unsigned long long* full = reinterpret_cast<unsigned long long*>(tsm + (size_t)NST * STAGE_BYTES);
  unsigned long long* empty = full + NST;

  const int t = threadIdx.x;
  const int wid = __shfl_sync(0xffffffffu, t >> 5, 0);  // warp-uniform by construction
  const int64_t seg0 = (int64_t)blockIdx.x * TSEG;
  const int nvalid = (int)min((int64_t)TSEG, a.inner - seg0);  // a multiple of 4 (launch precondition)
  const unsigned base = (unsigned)((int64_t)blockIdx.y * a.outer_stride + seg0);
  const int kb = a.out_begin + blockIdx.z * a.chunk_len;
  const int ke = min(a.out_end, kb + a.chunk_len);
  if (kb >= ke) return;
  const int nout = ke - kb;
  const int nrows = nout + 2 * R;  // line positions kb-R .. ke-1+R, boundary-mapped
  const unsigned bytes = (unsigned)nvalid * (unsigned)sizeof(float);
  const int st = a.stride;
  const bool live = 2 * t < nvalid;

  if (t == 0) {
#pragma unroll 1
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], TWARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // one thread: fetch row m of the chunk into stage `sidx` (out of line: one copy of the address
  // arithmetic instead of one per unrolled step keeps the sweep inside the instruction cache)
  auto issue = [&](int m, int sidx) {
    tma_issue_row(a.in0, TWO_IN ? a.in1 : nullptr, base, kb - R + m, a.L, a.bmode, st,
                  smem_u32(tsm) + (unsigned)sidx * STAGE_BYTES, smem_u32(&full[sidx]), bytes);
  };
  if (wid == 0) {
    if (elect_one()) {
#pragma unroll 1
      for (int m = 0; m < LEAD && m < nrows; ++m) issue(m, m);
    }
  }

  // Consume row m (stage m % NST); the warp on duty refills the stage consumed TSLACK steps ago with
  // row m + LEAD.  `mc` is m modulo 2*NST at compile time (stage and barrier parity), `m` the row.
  auto fetch = [&](int m, int mc, unsigned phase, int duty, float2& v0, float2& v1) {
    const int sidx = mc % NST;
    const unsigned par = ((unsigned)(mc / NST) & 1u) ^ phase;
    if (wid == duty) {
      const int mn = m + LEAD;
      if (mn < nrows) {
        const int pc = mc + 2 * NST - TSLACK;  // (m - TSLACK) mod 2*NST
        if (m >= TSLACK) mbar_wait(&empty[pc % NST], ((unsigned)(pc / NST) & 1u) ^ phase);
        if (elect_one()) issue(mn, pc % NST);
      }
    }
    mbar_wait(&full[sidx], par);
    const float2* sp = reinterpret_cast<const float2*>(tsm + (size_t)sidx * STAGE_BYTES) + t;
    v0 = sp[0];
    if (TWO_IN) v1 = sp[TSEG / 2];
    __syncwarp();
    if ((t & 31) == 0) mbar_arrive(&empty[sidx]);
  };

  // candidate emission (fused threshold + 1-D non-max test of the LoG response, see SArgs)
  constexpr bool emit = EMIT && (OP == MFISH_CONV_LAST);

  float2 r0[W];
  float2 r1[TWO_IN ? W : 1];
  float2 dummy;
#pragma unroll
  for (int j = 0; j < 2 * R; ++j) fetch(j, j, 0u, j % TWARPS, r0[j], TWO_IN ? r1[j] : dummy);

  unsigned off_k = base + (unsigned)(kb * st) + 2u * (unsigned)t;
  unsigned phase = 0;  // (o0 / NST) & 1
#pragma unroll 1
  for (int o0 = 0; o0 < nout; o0 += NST, phase ^= 1) {
    const unsigned off_first = off_k;
    unsigned cm = 0;  // bit u: output u of this sweep exceeds the threshold in one of the two columns
#pragma unroll
    for (int u = 0; u < NST; ++u) {
      if (o0 + u < nout) {
        const int mc = u + 2 * R;  // row index modulo 2*NST when phase == 0
        fetch(o0 + mc, mc, phase, u % TWARPS, r0[mc % W], TWO_IN ? r1[mc % W] : dummy);
        const float2 c0 = r0[(u + R) % W];
        float2 accG0 = mul2(a.t.g[0], c0);  // G * in0
        float2 accD0 = mul2(a.t.d[0], c0);  // D * in0
        float2 accG1 = make_float2(0.f, 0.f);  // G * in1
        if (TWO_IN) accG1 = mul2(a.t.g[0], r1[(u + R) % W]);
#pragma unroll
        for (int j = 1; j <= R; ++j) {
          const float2 s0 = add2(r0[(u + R - j + W) % W], r0[(u + R + j) % W]);
          if (OP != MFISH_CONV_LAST) accG0 = fma2(a.t.g[j], s0, accG0);
          if (OP != MFISH_CONV_G) accD0 = fma2(a.t.d[j], s0, accD0);
          if (TWO_IN) {
            const float2 s1 = add2(r1[(u + R - j + W) % W], r1[(u + R + j) % W]);
            accG1 = fma2(a.t.g[j], s1, accG1);
          }
        }
        if (OP == MFISH_CONV_G) {
          if (live) *reinterpret_cast<float2*>(a.out0 + off_k) = accG0;
        } else if (OP == MFISH_CONV_GD) {
          if (live) {
            *reinterpret_cast<float2*>(a.out0 + off_k) = accG0;
            *reinterpret_cast<float2*>(a.out1 + off_k) = accD0;
          }
        } else if (OP == MFISH_CONV_MID) {
          if (live) {
            *reinterpret_cast<float2*>(a.out0 + off_k) = accG0;
            *reinterpret_cast<float2*>(a.out1 + off_k) = add2(accD0, accG1);
          }
        } else {
          const float2 v = mul2(a.scale, add2(accD0, accG1));
          if (live) *reinterpret_cast<float2*>(a.out0 + off_k) = v;
          if (emit) cm |= (unsigned)(v.x > a.cand_thr || v.y > a.cand_thr) << u;
        }
        off_k += (unsigned)st;
      }
    }
    if (emit && live && cm != 0)
      emit_sweep_candidates(a.out0, off_first, (unsigned)st, cm, o0, nout, NST, a.cand_thr, a.cand, a.cand_count,
                            a.cand_cap);
  }
}

template <int R>
struct YZArgs {
  const float* in;
  float* outC;
  float* outF;
  int nz, nx, ny, nseg, bmode;
  const float* mm;
  Taps<R> t;
};


// Stages cycle through full (TMA bytes landed) / empty (all warps have copied their window to
// registers) mbarriers; there is no block-wide barrier in the plane loop.  Thread 0 doubles as the
// TMA producer: at step p it refills the stage of step p - 1, which the other warps have normally
// released long before, so its wait is free.
constexpr int YZ_CONSUMER_WARPS = YZ_THREADS / 32;

// Front half of one step, shared by all ring positions (kept out of line: the plane loop is
// unrolled by the ring depth and must stay inside the instruction cache): wait for the stage, write
// the row-end boundary extension if this warp needs it, copy the window to registers, release the
// stage, normalise, y pass.  Returns (G*in, G*in', D*in, D*in') for the thread's two outputs.
template <int R>
__device__ __noinline__ float4 yz_front(float* row, unsigned long long* full_bar, unsigned long long* empty_bar,
                                        unsigned parity, int wbeg, int dstL, int srcL, int dstR, int srcR, float sc,
                                        float noff, const Taps<R>* __restrict__ taps) {
  constexpr int W = 2 * R + 1;
  mbar_wait(full_bar, parity);
  if (dstL >= 0) row[dstL] = row[srcL];
  if (dstR >= 0) row[dstR] = row[srcR];
  __syncwarp();
  float w[W + 1];
  const float2* wp = reinterpret_cast<const float2*>(row + wbeg);
#pragma unroll
  for (int j = 0; j < (W + 1) / 2; ++j) {
    const float2 v = wp[j];
    w[2 * j] = v.x;
    w[2 * j + 1] = v.y;
  }
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(empty_bar);  // this warp no longer needs the stage
#pragma unroll
  for (int j = 0; j <= W; ++j) w[j] = fmaf(w[j], sc, noff);  // su.normalize_image folded into the load
  const float tg0 = taps->g[0], td0 = taps->d[0];
  float g0 = tg0 * w[R], g1 = tg0 * w[R + 1];
  float d0 = td0 * w[R], d1 = td0 * w[R + 1];
#pragma unroll
  for (int k = 1; k <= R; ++k) {
    const float tg = taps->g[k], td = taps->d[k];
    const float s0 = w[R - k] + w[R + k];
    const float s1 = w[R + 1 - k] + w[R + 1 + k];
    g0 = fmaf(tg, s0, g0);
    g1 = fmaf(tg, s1, g1);
    d0 = fmaf(td, s0, d0);
    d1 = fmaf(td, s1, d1);
  }
  return make_float4(g0, g1, d0, d1);
}

// Producer step (one thread per block, out of line for the same reason): once every warp has
// released the stage of step q, fetch the row segment of step q + YZ_STAGES into it.
__device__ __noinline__ void yz_refill(const float* in_row_lo, int64_t plane, int nz, int bmode, int R, int q, int P,
                                       float* stage_dst0, int stage_floats, unsigned long long* full,
                                       unsigned long long* empty, unsigned bytes, bool wait_empty) {
  const int p = wait_empty ? q + YZ_STAGES : q;
  if (p >= P) return;
  const int stg = p % YZ_STAGES;
  if (wait_empty) mbar_wait(&empty[stg], (unsigned)((q / YZ_STAGES) & 1));
  const int zs = bmode == MFISH_BOUNDARY_NEAREST ? bmap<MFISH_BOUNDARY_NEAREST>(p - R, nz)
                                                 : bmap<MFISH_BOUNDARY_REFLECT>(p - R, nz);
  mbar_expect_tx(&full[stg], bytes);
  tma_load_1d(stage_dst0 + (size_t)stg * stage_floats, in_row_lo + (int64_t)zs * plane, bytes, &full[stg]);
}

template <int R>
__global__ void __launch_bounds__(YZ_THREADS, 2) conv_yz_kernel(const __grid_constant__ YZArgs<R> a) {
  constexpr int RP = (R + 3) & ~3;  // halo rounded up to 16 bytes
  constexpr int W = 2 * R + 1;      // ring depth = taps
  constexpr int ROW = YZ_SEG + 2 * RP;
  __shared__ __align__(128) float s_in[YZ_STAGES][ROW];
  __shared__ __align__(8) unsigned long long s_full[YZ_STAGES];
  __shared__ __align__(8) unsigned long long s_empty[YZ_STAGES];

  const int tid = threadIdx.x;
  const int x = blockIdx.x / a.nseg;
  const int seg0 = (blockIdx.x - x * a.nseg) * YZ_SEG;
  const int ny = a.ny, nz = a.nz, nx = a.nx;
  const int row0 = seg0 - RP;  // y of smem index 0
  const int lo = max(row0, 0), hi = min(seg0 + YZ_SEG + RP, ny);
  const int P = nz + 2 * R;  // planes walked: source plane of step p is bmap(p - R)
  const int bmode = a.bmode;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < YZ_STAGES; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], YZ_CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const unsigned bytes = (unsigned)(hi - lo) * 4u;
  const float* in_row_lo = a.in + (int64_t)x * ny + lo;  // + zs * plane
  const int64_t plane = (int64_t)nx * ny;
  float* stage_dst0 = &s_in[0][lo - row0];
  if (tid == 0) {
    for (int p = 0; p < YZ_STAGES; ++p)
      yz_refill(in_row_lo, plane, nz, bmode, R, p, P, stage_dst0, ROW, s_full, s_empty, bytes, false);
  }

  float mn, sc;
  bool degen;
  load_norm(a.mm, mn, sc, degen);
  float noff = -mn * sc;  // normalised = v * sc + noff
  if (degen) {            // constant image: np.interp gives 1 everywhere
    sc = 0.f;
    noff = 1.f;
  }
  const int y0 = seg0 + 2 * tid;
  const bool active = y0 < ny;  // ny % 4 == 0, so a float2 never straddles the row end
  // window of this thread: smem indices [wbeg, wbeg + W].  In a row-end segment the TMA fetch stops at
  // the row end; the warps whose windows reach past it write the boundary extension (R values,
  // taken from the fetched range) into the stage themselves before loading their windows.
  const int wbeg = 2 * tid + (RP - R);
  const int lane = tid & 31;
  const int wfirst = 2 * (tid - lane) + (RP - R);          // first smem index the warp reads
  const int wlast = 2 * (tid - lane + 31) + (RP - R) + W;  // last one
  int dstL = -1, srcL = 0, dstR = -1, srcR = 0;
  if ((lo != row0) && (row0 + wfirst < lo) && lane < R) {
    const int y = lo - 1 - lane;
    const int yy = bmode == MFISH_BOUNDARY_NEAREST ? bmap1<MFISH_BOUNDARY_NEAREST>(y, ny)
                                                   : bmap1<MFISH_BOUNDARY_REFLECT>(y, ny);
    dstL = y - row0;
    srcL = yy - row0;
  }
  if ((hi != seg0 + YZ_SEG + RP) && (row0 + wlast >= hi) && (row0 + wfirst < hi + R) && lane < R) {
    const int y = hi + lane;
    const int yy = bmode == MFISH_BOUNDARY_NEAREST ? bmap1<MFISH_BOUNDARY_NEAREST>(y, ny)
                                                   : bmap1<MFISH_BOUNDARY_REFLECT>(y, ny);
    dstR = y - row0;
    srcR = yy - row0;
  }
  float2 A2[W], B2[W];
#pragma unroll
  for (int i = 0; i < W; ++i) A2[i] = B2[i] = make_float2(0.f, 0.f);
  float2* __restrict__ outC = reinterpret_cast<float2*>(a.outC + ((int64_t)x * ny + y0));
  float2* __restrict__ outF = reinterpret_cast<float2*>(a.outF + ((int64_t)x * ny + y0));
  const int64_t plane2 = (int64_t)nx * ny / 2;  // float2 elements per z plane

  for (int p0 = 0; p0 < P; p0 += W) {
#pragma unroll
    for (int u = 0; u < W; ++u) {
      const int p = p0 + u;
      if (p < P) {
        const int stg = p % YZ_STAGES;
        const float4 gd = yz_front<R>(s_in[stg], &s_full[stg], &s_empty[stg], (unsigned)((p / YZ_STAGES) & 1), wbeg,
                                      dstL, srcL, dstR, srcR, sc, noff, &a.t);
        A2[u] = make_float2(gd.x, gd.y);
        B2[u] = make_float2(gd.z, gd.w);
        if (tid == 0 && p >= 1)
          yz_refill(in_row_lo, plane, nz, bmode, R, p - 1, P, stage_dst0, ROW, s_full, s_empty, bytes, true);
        // z pass on the ring: the centre of the window that ends at step p is step p - R
        if (p >= 2 * R && active) {
          const int c = (u - R + W) % W;
          float2 C = mul2(a.t.g[0], A2[c]);
          float2 F = fma2(a.t.g[0], B2[c], mul2(a.t.d[0], A2[c]));
#pragma unroll
          for (int k = 1; k <= R; ++k) {
            const float2 sA = add2(A2[(c - k + W) % W], A2[(c + k) % W]);
            const float2 sB = add2(B2[(c - k + W) % W], B2[(c + k) % W]);
            C = fma2(a.t.g[k], sA, C);
            F = fma2(a.t.d[k], sA, F);
            F = fma2(a.t.g[k], sB, F);
          }
          const int64_t o = (int64_t)(p - 2 * R) * plane2;
          outC[o] = C;
          outF[o] = F;
        }
      }
    }
  }
}

// --------------------------------------------------------------------------------------------
// host-side dispatch
// --------------------------------------------------------------------------------------------
template <int R>
static void fill_taps(Taps<R>& t, const double* g, const double* d, int radius) {
  for (int k = 0; k <= R; ++k) {
    t.g[k] = (k <= radius && g) ? (float)g[k] : 0.f;
    t.d[k] = (k <= radius && d) ? (float)d[k] : 0.f;
  }
}

struct PassDesc {
  const float* in0;
  const float* in1;
  float* out0;
  float* out1;
  int64_t nz, nx, ny;
  int axis;
  const double* g;
  const double* d;
  int radius, boundary, op;
  float scale;
  const float* mm;
  cudaStream_t st;
  int out_begin = 0;  // z axis only: produce outputs for [out_begin, out_end); out_end < 0 = whole line
  int out_end = -1;
  float cand_thr = 0.f;  // LAST only: candidate emission (see SArgs)
  uint32_t* cand = nullptr;
  uint32_t* cand_count = nullptr;
  uint32_t cand_cap = 0;
};

static int upload_taps(const PassDesc& p, float** dg, float** dd) {
  int n = p.radius + 1;
  std::vector<float> hg(n), hd(n);
  for (int k = 0; k < n; ++k) {
    hg[k] = p.g ? (float)p.g[k] : 0.f;
    hd[k] = p.d ? (float)p.d[k] : 0.f;
  }
  MFISH_CUDA(cudaMallocAsync((void**)dg, 2 * n * sizeof(float), p.st));
  *dd = *dg + n;
  MFISH_CUDA(cudaMemcpyAsync(*dg, hg.data(), n * sizeof(float), cudaMemcpyHostToDevice, p.st));
  MFISH_CUDA(cudaMemcpyAsync(*dd, hd.data(), n * sizeof(float), cudaMemcpyHostToDevice, p.st));
  // pageable-source async copies return after staging, so the vectors may go out of scope
  return MFISH_OK;
}

template <int R>
static int launch_y(const PassDesc& p) {
  YArgs<R> a;
  a.in = p.in0;
  a.out0 = p.out0;
  a.out1 = p.out1;
  a.nrows = p.nz * p.nx;
  a.ny = (int)p.ny;
  a.nseg = (int)((p.ny + YSEG - 1) / YSEG);
  a.bmode = p.boundary;
  a.mm = p.mm;
  fill_taps<R>(a.t, p.g, p.d, p.radius);
  int64_t blocks = ((a.nrows + YROWS - 1) / YROWS) * a.nseg;
  MFISH_REQUIRE(blocks < (1ll << 31), "sepconv(y): too many row segments");
  ProfScope prof(p.op == MFISH_CONV_G ? "conv_y_g" : "conv_y_gd", p.st);
  if (p.op == MFISH_CONV_G)
    conv_y_kernel<R, MFISH_CONV_G><<<(unsigned)blocks, YTHREADS, 0, p.st>>>(a);
  else
    conv_y_kernel<R, MFISH_CONV_GD><<<(unsigned)blocks, YTHREADS, 0, p.st>>>(a);
  count_launch(1);
  return check_launch("conv_y");
}

static bool strided_tma_enabled() {
  static const bool on = [] {
    const char* e = getenv("MFISH_STRIDED_TMA");
    return !(e && e[0] == '0');
  }();
  return on;
}

template <int R>
static int launch_strided(const PassDesc& p) {
  SArgs<R> a;
  a.in0 = p.in0;
  a.in1 = p.in1;
  a.out0 = p.out0;
  a.out1 = p.out1;
  int64_t nouter;
  if (p.axis == MFISH_AXIS_Z) {
    a.inner = p.nx * p.ny;
    a.stride = (int)(p.nx * p.ny);
    a.outer_stride = 0;
    a.L = (int)p.nz;
    nouter = 1;
  } else {
    a.inner = p.ny;
    a.stride = (int)p.ny;
    a.outer_stride = p.nx * p.ny;
    a.L = (int)p.nx;
    nouter = p.nz;
  }
  a.bmode = p.boundary;
  a.scale = p.scale;
  a.out_begin = p.out_begin;
  a.out_end = p.out_end < 0 ? a.L : p.out_end;
  a.cand_thr = p.cand_thr;
  a.cand = p.op == MFISH_CONV_LAST ? p.cand : nullptr;
  a.cand_count = p.cand_count;
  a.cand_cap = p.cand_cap;
  const int nout = a.out_end - a.out_begin;
  fill_taps<R>(a.t, p.g, p.d, p.radius);
  static const char* const names[2][4] = {{"conv_z_g", "conv_z_gd", "conv_z_mid", "conv_z_last"},
                                          {"conv_x_g", "conv_x_gd", "conv_x_mid", "conv_x_last"}};
  const int min_chunk = std::max(64, 8 * R);
  const int64_t want = (int64_t)num_sms() * 16 * 2;  // blocks
  // TMA-staged walker: rows must be 16-byte addressable pieces
  const bool two_in = p.op == MFISH_CONV_MID || p.op == MFISH_CONV_LAST;
  const bool two_out = p.op == MFISH_CONV_GD || p.op == MFISH_CONV_MID;
  const bool tma_ok = R <= TMA_MAX_R && strided_tma_enabled() && (a.inner % 4 == 0) && (a.stride % 4 == 0) &&
                      (a.outer_stride % 4 == 0) && aligned16(a.in0) && (!two_in || aligned16(a.in1)) &&
                      aligned16(a.out0) && (!two_out || aligned16(a.out1));
  if (tma_ok) {
    if constexpr (R <= TMA_MAX_R) {
      int64_t bx = (a.inner + TSEG - 1) / TSEG;
      int chunks = 1;
      static const int64_t want_tma = [] {
        const char* e = getenv("MFISH_TMA_WANT");
        return e ? (int64_t)atoll(e) : (int64_t)0;
      }();
      const int64_t wt = want_tma > 0 ? want_tma : want / 2;
      while (chunks < 16384 && bx * nouter * chunks < wt && nout / (chunks * 2) >= min_chunk) chunks *= 2;
      a.chunk_len = (nout + chunks - 1) / chunks;
      MFISH_REQUIRE(bx < (1ll << 31) && nouter < 65536, "sepconv(strided): grid too large");
      dim3 grid((unsigned)bx, (unsigned)nouter, (unsigned)chunks);
      ProfScope prof(names[p.axis == MFISH_AXIS_Z ? 0 : 1][p.op], p.st);
      auto go = [&](auto kern, size_t smem) {
        // every op of one radius shares a function-pointer type, so remember the configured kernels by
        // address: the walker wants the large shared-memory carve-out (5 blocks x 35 KB per SM)
        // (function attributes are per device: the key carries the current device)
        static std::vector<std::pair<int, const void*>> configured;
        static std::mutex configured_lock;
        std::lock_guard<std::mutex> guard(configured_lock);
        int dev = 0;
        cudaGetDevice(&dev);
        const std::pair<int, const void*> key(dev, reinterpret_cast<const void*>(kern));
        if (std::find(configured.begin(), configured.end(), key) == configured.end()) {
          cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
          if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          configured.push_back(key);
        }
        kern<<<grid, TTHREADS, smem, p.st>>>(a);
      };
      switch (p.op) {
        case MFISH_CONV_G: go(conv_strided_tma_kernel<R, MFISH_CONV_G>, TmaTune<R, MFISH_CONV_G>::kSmem); break;
        case MFISH_CONV_GD: go(conv_strided_tma_kernel<R, MFISH_CONV_GD>, TmaTune<R, MFISH_CONV_GD>::kSmem); break;
        case MFISH_CONV_MID: go(conv_strided_tma_kernel<R, MFISH_CONV_MID>, TmaTune<R, MFISH_CONV_MID>::kSmem); break;
        default:
          if (a.cand != nullptr)
            go(conv_strided_tma_kernel<R, MFISH_CONV_LAST, true>, TmaTune<R, MFISH_CONV_LAST>::kSmem);
          else
            go(conv_strided_tma_kernel<R, MFISH_CONV_LAST>, TmaTune<R, MFISH_CONV_LAST>::kSmem);
          break;
      }
      count_launch(1);
      return check_launch("conv_strided_tma");
    }
  }
  // split long lines into chunks until the grid holds >= ~4 waves of threads
  int64_t bx = (a.inner + STHREADS - 1) / STHREADS;
  int chunks = 1;
  while (chunks < 16384 && bx * nouter * chunks < want && nout / (chunks * 2) >= min_chunk) chunks *= 2;
  a.chunk_len = (nout + chunks - 1) / chunks;
  MFISH_REQUIRE(bx < (1ll << 31) && nouter < 65536, "sepconv(strided): grid too large");
  dim3 grid((unsigned)bx, (unsigned)nouter, (unsigned)chunks);
  ProfScope prof(names[p.axis == MFISH_AXIS_Z ? 0 : 1][p.op], p.st);
  switch (p.op) {
    case MFISH_CONV_G: conv_strided_kernel<R, MFISH_CONV_G><<<grid, STHREADS, 0, p.st>>>(a); break;
    case MFISH_CONV_GD: conv_strided_kernel<R, MFISH_CONV_GD><<<grid, STHREADS, 0, p.st>>>(a); break;
    case MFISH_CONV_MID: conv_strided_kernel<R, MFISH_CONV_MID><<<grid, STHREADS, 0, p.st>>>(a); break;
    default:
      if (a.cand != nullptr)
        conv_strided_kernel<R, MFISH_CONV_LAST, true><<<grid, STHREADS, 0, p.st>>>(a);
      else
        conv_strided_kernel<R, MFISH_CONV_LAST><<<grid, STHREADS, 0, p.st>>>(a);
      break;
  }
  count_launch(1);
  return check_launch("conv_strided");
}

static int launch_generic(const PassDesc& p);

static bool big_gauss_enabled() {
  static const bool on = [] {
    const char* e = getenv("MFISH_BIG_GAUSS");
    return !(e && e[0] == '0');
  }();
  return on;
}

template <int R>
static int launch_R(const PassDesc& p);

static int launch_generic(const PassDesc& p) {
  float *dg = nullptr, *dd = nullptr;
  int rc = upload_taps(p, &dg, &dd);
  if (rc != MFISH_OK) return rc;
  if (p.op == MFISH_CONV_G && p.radius > 16 && p.out_begin == 0 && p.out_end < 0 && big_gauss_enabled()) {
    // plain Gaussian with many taps (segmentation blurs): the FMA-organised kernels of bigconv.cu
    rc = big_gauss_pass(p.in0, p.out0, p.nz, p.nx, p.ny, p.axis, dg, p.radius, p.boundary, p.mm, p.st);
    cudaFreeAsync(dg, p.st);
    return rc;
  }
  ProfScope prof(p.axis == MFISH_AXIS_Y ? "conv_y_generic" : "conv_strided_generic", p.st);
  if (p.axis == MFISH_AXIS_Y) {
    YGenArgs a;
    a.in = p.in0;
    a.out0 = p.out0;
    a.out1 = p.out1;
    a.nrows = p.nz * p.nx;
    a.ny = (int)p.ny;
    a.nseg = (int)((p.ny + YSEG - 1) / YSEG);
    a.bmode = p.boundary;
    a.radius = p.radius;
    a.op = p.op;
    a.mm = p.mm;
    a.taps_g = dg;
    a.taps_d = dd;
    int64_t blocks = a.nrows * a.nseg;
    size_t smem = (size_t)(2 * (p.radius + 1) + YSEG + 2 * p.radius) * sizeof(float);
    conv_y_generic_kernel<<<(unsigned)blocks, YTHREADS, smem, p.st>>>(a);
  } else {
    SGenArgs a;
    a.in0 = p.in0;
    a.in1 = p.in1;
    a.out0 = p.out0;
    a.out1 = p.out1;
    int64_t nouter;
    if (p.axis == MFISH_AXIS_Z) {
      a.inner = p.nx * p.ny;
      a.stride = p.nx * p.ny;
      a.outer_stride = 0;
      a.L = (int)p.nz;
      nouter = 1;
    } else {
      a.inner = p.ny;
      a.stride = p.ny;
      a.outer_stride = p.nx * p.ny;
      a.L = (int)p.nx;
      nouter = p.nz;
    }
    a.bmode = p.boundary;
    a.radius = p.radius;
    a.op = p.op;
    a.scale = p.scale;
    a.taps_g = dg;
    a.taps_d = dd;
    a.out_begin = p.out_begin;
    const int nout = (p.out_end < 0 ? a.L : p.out_end) - p.out_begin;
    int64_t bx = (a.inner + STHREADS - 1) / STHREADS;
    MFISH_REQUIRE(nouter < 65536 && a.L < 65536, "sepconv(generic): grid too large");
    dim3 grid((unsigned)bx, (unsigned)nouter, (unsigned)nout);
    size_t smem = (size_t)2 * (p.radius + 1) * sizeof(float);
    conv_strided_generic_kernel<<<grid, STHREADS, smem, p.st>>>(a);
  }
  count_launch(1);
  rc = check_launch("conv_generic");
  cudaFreeAsync(dg, p.st);
  return rc;
}

template <int R>
static int launch_R(const PassDesc& p) {
  if (p.axis == MFISH_AXIS_Y) return launch_y<R>(p);
  // the ring walker folds indices once, so the line must be at least R long and the plane
  // stride must fit 32 bits; anything else goes through the gather fallback
  const int64_t L = p.axis == MFISH_AXIS_Z ? p.nz : p.nx;
  if (L < R || p.nz * p.nx * p.ny >= (1ll << 32)) {
    if (p.cand != nullptr) {
      set_error("sepconv: candidate emission needs the register-ring kernel");
      return MFISH_ERR_UNSUPPORTED;
    }
    return launch_generic(p);
  }
  return launch_strided<R>(p);
}

int run_pass(const PassDesc& p) {
  MFISH_REQUIRE(p.in0 && p.out0, "sepconv: null in0/out0");
  MFISH_REQUIRE(p.nz > 0 && p.nx > 0 && p.ny > 0, "sepconv: empty volume");
  MFISH_REQUIRE(p.nz < (1 << 30) && p.nx < (1 << 30) && p.ny < (1 << 30), "sepconv: dimension too large");
  MFISH_REQUIRE(p.radius >= 0 && p.radius <= MFISH_MAX_RADIUS, "sepconv: radius %d out of range", p.radius);
  MFISH_REQUIRE(p.g != nullptr, "sepconv: taps_g required");
  MFISH_REQUIRE(p.op >= MFISH_CONV_G && p.op <= MFISH_CONV_LAST, "sepconv: bad op");
  MFISH_REQUIRE(p.op == MFISH_CONV_G || p.d != nullptr, "sepconv: taps_d required for this op");
  const bool two_in = (p.op == MFISH_CONV_MID || p.op == MFISH_CONV_LAST);
  const bool two_out = (p.op == MFISH_CONV_GD || p.op == MFISH_CONV_MID);
  MFISH_REQUIRE(!two_in || p.in1, "sepconv: in1 required for this op");
  MFISH_REQUIRE(!two_out || p.out1, "sepconv: out1 required for this op");
  MFISH_REQUIRE(p.boundary == MFISH_BOUNDARY_REFLECT || p.boundary == MFISH_BOUNDARY_NEAREST,
                "sepconv: bad boundary mode");
  MFISH_REQUIRE(p.out0 != p.in0 && p.out0 != p.in1 && (!two_out || (p.out1 != p.in0 && p.out1 != p.in1)),
                "sepconv: in-place operation is not supported");
  MFISH_REQUIRE(p.out_end < 0 || (p.axis == MFISH_AXIS_Z && p.out_begin >= 0 && p.out_begin < p.out_end &&
                                  p.out_end <= p.nz),
                "sepconv: bad output range");
  if (p.axis == MFISH_AXIS_Y) {
    if (two_in) {
      set_error("sepconv: ops MID/LAST are implemented for the z and x axes only");
      return MFISH_ERR_UNSUPPORTED;
    }
  } else {
    MFISH_REQUIRE(p.axis == MFISH_AXIS_Z || p.axis == MFISH_AXIS_X, "sepconv: bad axis");
    MFISH_REQUIRE(p.mm == nullptr, "sepconv: normalise-on-load is implemented for the y axis only");
  }
  const int r = p.radius;
  if (r <= 1) return launch_R<1>(p);
  if (r <= 2) return launch_R<2>(p);
  if (r <= 3) return launch_R<3>(p);
  if (r <= 4) return launch_R<4>(p);
  if (r <= 6) return launch_R<6>(p);
  if (r <= 8) return launch_R<8>(p);
  if (r <= 10) return launch_R<10>(p);
  if (r <= 12) return launch_R<12>(p);
  if (r <= 16) return launch_R<16>(p);
  return launch_generic(p);
}

// The fused, TMA-staged y+z pass is correct (same parity tests) but measured 2.9-3.4 ms against 2.93 ms for
// the separate y and z kernels on the 2048x2048x100 stack: 125 registers cap it at 16 warps per SM and the
// ring-unrolled body (50-98 KB of SASS) misses the instruction cache.  It stays opt-in until it wins.
static const bool g_disable_fused_yz = getenv("MFISH_FUSED_YZ") == nullptr;

template <int R>
static int launch_yz(const float* in, float* outC, float* outF, int64_t nz, int64_t nx, int64_t ny, const double* g,
                     const double* d, int radius, const float* mm, cudaStream_t st) {
  YZArgs<R> a;
  a.in = in;
  a.outC = outC;
  a.outF = outF;
  a.nz = (int)nz;
  a.nx = (int)nx;
  a.ny = (int)ny;
  a.nseg = (int)((ny + YZ_SEG - 1) / YZ_SEG);
  a.bmode = MFISH_BOUNDARY_REFLECT;
  a.mm = mm;
  fill_taps<R>(a.t, g, d, radius);
  ProfScope prof("conv_yz_fused", st);
  conv_yz_kernel<R><<<(unsigned)(nx * a.nseg), YZ_THREADS, 0, st>>>(a);
  count_launch(1);
  return check_launch("conv_yz");
}

// Host-side taps, identical formula to muscle-fish_b200/taps.py (scipy _gaussian_kernel1d).
static void host_taps(double sigma, std::vector<double>& g, std::vector<double>& d, int& radius) {
  radius = (int)(4.0 * sigma + 0.5);
  g.assign(radius + 1, 0.0);
  d.assign(radius + 1, 0.0);
  if (sigma <= 0.0) {
    radius = 0;
    g.assign(1, 1.0);
    d.assign(1, 0.0);
    return;
  }
  double s2 = sigma * sigma, sum = 0.0;
  std::vector<double> phi(2 * radius + 1);
  for (int x = -radius; x <= radius; ++x) {
    phi[x + radius] = exp(-0.5 / s2 * (double)(x * x));
  }
  for (double v : phi) sum += v;
  for (int k = 0; k <= radius; ++k) {
    double ph = phi[k + radius] / sum;
    g[k] = ph;
    d[k] = ((double)(k * k) / (s2 * s2) - 1.0 / s2) * ph;
  }
}

}  // namespace mfish

using namespace mfish;

extern "C" int mfish_sepconv_f32(const float* in0_dev, const float* in1_dev, float* out0_dev,
                                 float* out1_dev, int64_t nz, int64_t nx, int64_t ny, int axis,
                                 const double* taps_g_host, const double* taps_d_host, int radius,
                                 int boundary, int op, float out_scale, const float* norm_minmax_dev,
                                 void* stream) {
  PassDesc p{in0_dev, in1_dev, out0_dev, out1_dev, nz, nx, ny, axis, taps_g_host, taps_d_host,
             radius, boundary, op, out_scale, norm_minmax_dev, as_stream(stream)};
  return run_pass(p);
}

extern "C" int mfish_gaussian3d_f32(const float* in_dev, float* out_dev, float* tmp_dev, int64_t nz,
                                    int64_t nx, int64_t ny, const double* sigma_zxy_host, int boundary,
                                    const float* norm_minmax_dev, void* stream) {
  MFISH_REQUIRE(in_dev && out_dev && tmp_dev && sigma_zxy_host, "gaussian3d: null pointer");
  MFISH_REQUIRE(in_dev != out_dev && in_dev != tmp_dev && out_dev != tmp_dev, "gaussian3d: buffers must differ");
  const double sz = sigma_zxy_host[0], sx = sigma_zxy_host[1], sy = sigma_zxy_host[2];
  // pass order y, z, x; a zero sigma skips the pass (y is kept when normalising on load)
  struct P { int axis; double s; };
  P passes[3];
  int np = 0;
  if (sy > 0.0 || norm_minmax_dev) passes[np++] = {MFISH_AXIS_Y, sy};
  if (sz > 0.0) passes[np++] = {MFISH_AXIS_Z, sz};
  if (sx > 0.0) passes[np++] = {MFISH_AXIS_X, sx};
  cudaStream_t st = as_stream(stream);
  if (np == 0) {
    MFISH_CUDA(cudaMemcpyAsync(out_dev, in_dev, sizeof(float) * nz * nx * ny, cudaMemcpyDeviceToDevice, st));
    return MFISH_OK;
  }
  const float* src = in_dev;
  float* dst = (np & 1) ? out_dev : tmp_dev;
  for (int i = 0; i < np; ++i) {
    std::vector<double> g, d;
    int radius;
    host_taps(passes[i].s, g, d, radius);
    PassDesc p{src, nullptr, dst, nullptr, nz, nx, ny, passes[i].axis, g.data(), nullptr, radius,
               boundary, MFISH_CONV_G, 1.0f, (passes[i].axis == MFISH_AXIS_Y) ? norm_minmax_dev : nullptr, st};
    int rc = run_pass(p);
    if (rc != MFISH_OK) return rc;
    src = dst;
    dst = (dst == out_dev) ? tmp_dev : out_dev;
  }
  return MFISH_OK;
}

// LoG of the planes [out_z0, out_z1) of a z-extended slab: the y pass runs on every plane, the z pass
// produces only the requested planes (its inputs reach `radius` planes further), the x pass runs on those.
static int log3d_impl(const float* in_dev, float* out_dev, float* tmp1_dev, float* tmp2_dev, float* tmp3_dev,
                      int64_t nz, int64_t nx, int64_t ny, double sigma, const float* norm_minmax_dev, int64_t out_z0,
                      int64_t out_z1, float cand_thr, uint32_t* cand, uint32_t* cand_count, uint32_t cand_cap,
                      void* stream) {
  MFISH_REQUIRE(in_dev && out_dev && tmp1_dev && tmp2_dev && tmp3_dev, "log3d: null pointer");
  MFISH_REQUIRE(sigma > 0.0, "log3d: sigma must be positive");
  MFISH_REQUIRE(out_z0 >= 0 && out_z0 < out_z1 && out_z1 <= nz, "log3d: bad plane range");
  std::vector<double> g, d;
  int radius;
  host_taps(sigma, g, d, radius);
  cudaStream_t st = as_stream(stream);
  const bool whole = (out_z0 == 0 && out_z1 == nz);
  int rc = MFISH_ERR_UNSUPPORTED;
  // fused y+z pass (TMA staged) when the rows allow 16-byte bulk copies
  const bool fuse_ok = whole && (ny % 4 == 0) && ny >= 32 && aligned16(in_dev) && aligned16(tmp2_dev) &&
                       aligned16(tmp3_dev) && nz < (1 << 20) && nx * ((ny + YZ_SEG - 1) / YZ_SEG) < (1ll << 31) &&
                       !g_disable_fused_yz;
  if (fuse_ok && radius == 8)
    rc = launch_yz<8>(in_dev, tmp2_dev, tmp3_dev, nz, nx, ny, g.data(), d.data(), radius, norm_minmax_dev, st);
  else if (fuse_ok && radius == 4)
    rc = launch_yz<4>(in_dev, tmp2_dev, tmp3_dev, nz, nx, ny, g.data(), d.data(), radius, norm_minmax_dev, st);
  else if (fuse_ok && radius == 6)
    rc = launch_yz<6>(in_dev, tmp2_dev, tmp3_dev, nz, nx, ny, g.data(), d.data(), radius, norm_minmax_dev, st);
  if (rc == MFISH_ERR_UNSUPPORTED) {
    // y: in -> (A = G in, B = D in)
    PassDesc py{in_dev, nullptr, out_dev, tmp1_dev, nz, nx, ny, MFISH_AXIS_Y, g.data(), d.data(), radius,
                MFISH_BOUNDARY_REFLECT, MFISH_CONV_GD, 1.0f, norm_minmax_dev, st};
    rc = run_pass(py);
    if (rc != MFISH_OK) return rc;
    // z: (A, B) -> (C = G A, F = D A + G B)
    PassDesc pz{out_dev, tmp1_dev, tmp2_dev, tmp3_dev, nz, nx, ny, MFISH_AXIS_Z, g.data(), d.data(), radius,
                MFISH_BOUNDARY_REFLECT, MFISH_CONV_MID, 1.0f, nullptr, st};
    if (!whole) {
      pz.out_begin = (int)out_z0;
      pz.out_end = (int)out_z1;
    }
    rc = run_pass(pz);
  }
  if (rc != MFISH_OK) return rc;
  // x: (C, F) -> -sigma^2 (D C + G F), on the requested planes only
  const int64_t off = out_z0 * nx * ny;
  PassDesc px{tmp2_dev + off, tmp3_dev + off, out_dev + off, nullptr, out_z1 - out_z0, nx, ny, MFISH_AXIS_X, g.data(),
              d.data(), radius, MFISH_BOUNDARY_REFLECT, MFISH_CONV_LAST, (float)(-(sigma * sigma)), nullptr, st};
  if (cand != nullptr) {
    if (radius > 16) {
      set_error("log3d: fused candidate emission supports radius <= 16");
      return MFISH_ERR_UNSUPPORTED;
    }
    px.cand_thr = cand_thr;
    px.cand = cand;
    px.cand_count = cand_count;
    px.cand_cap = cand_cap;
  }
  return run_pass(px);
}

extern "C" int mfish_log3d_range_f32(const float* in_dev, float* out_dev, float* tmp1_dev, float* tmp2_dev,
                                     float* tmp3_dev, int64_t nz, int64_t nx, int64_t ny, double sigma,
                                     const float* norm_minmax_dev, int64_t out_z0, int64_t out_z1, void* stream) {
  return log3d_impl(in_dev, out_dev, tmp1_dev, tmp2_dev, tmp3_dev, nz, nx, ny, sigma, norm_minmax_dev, out_z0, out_z1,
                    0.f, nullptr, nullptr, 0, stream);
}

namespace mfish {
int nms_verify_candidates(const float* cube, int64_t nz, int64_t nx, int64_t ny, float thr, const uint32_t* cand,
                          const uint32_t* cand_count, uint32_t cand_cap, int64_t* out_idx, float* out_val,
                          uint32_t* count, int64_t capacity, cudaStream_t st);
}

// LoG + thresholded 26-neighbour non-max suppression + compaction in one call (single scale): the x pass
// emits, next to the LoG volume, the voxels above the threshold that are 1-D maxima along x (warp-ballot
// append); a second, tiny kernel checks the full 3x3x3 neighbourhood of those candidates only.  Saves the
// stand-alone NMS pass over the volume.  cand_ws_dev: (cand_cap + 1) uint32 of scratch.  *count_dev =
// 0xFFFFFFFF reports a candidate overflow: run mfish_nms_compact_f32 on out_dev instead.
// The planes [out_z0, out_z1) of a z-extended slab are treated as a cube of their own: LoG as
// mfish_log3d_range_f32, candidates and peak indices relative to that cube ((x ny + y)(out_z1 - out_z0) + z -
// out_z0), zero padding at its two faces -- a z-slab keeps the peaks of its own planes, which lie inside.
extern "C" int mfish_log3d_peaks_range_f32(const float* in_dev, float* out_dev, float* tmp1_dev, float* tmp2_dev,
                                           float* tmp3_dev, int64_t nz, int64_t nx, int64_t ny, double sigma,
                                           const float* norm_minmax_dev, int64_t out_z0, int64_t out_z1,
                                           float threshold, void* cand_ws_dev, int64_t cand_cap, int64_t* out_idx_dev,
                                           float* out_val_dev, uint32_t* count_dev, int64_t capacity, void* stream) {
  MFISH_REQUIRE(cand_ws_dev && out_idx_dev && out_val_dev && count_dev, "log3d_peaks: null pointer");
  MFISH_REQUIRE(cand_cap > 0 && cand_cap < (1ll << 31) && capacity > 0, "log3d_peaks: bad capacity");
  MFISH_REQUIRE(threshold >= 0.f, "log3d_peaks: threshold must be >= 0");
  MFISH_REQUIRE(out_z0 >= 0 && out_z0 < out_z1 && out_z1 <= nz, "log3d_peaks: bad plane range");
  cudaStream_t st = as_stream(stream);
  uint32_t* cand_count = reinterpret_cast<uint32_t*>(cand_ws_dev);
  uint32_t* cand = cand_count + 1;
  MFISH_CUDA(cudaMemsetAsync(cand_count, 0, sizeof(uint32_t), st));
  int rc = log3d_impl(in_dev, out_dev, tmp1_dev, tmp2_dev, tmp3_dev, nz, nx, ny, sigma, norm_minmax_dev, out_z0,
                      out_z1, threshold, cand, cand_count, (uint32_t)cand_cap, stream);
  if (rc != MFISH_OK) return rc;
  return nms_verify_candidates(out_dev + out_z0 * nx * ny, out_z1 - out_z0, nx, ny, threshold, cand, cand_count,
                               (uint32_t)cand_cap, out_idx_dev, out_val_dev, count_dev, capacity, st);
}

extern "C" int mfish_log3d_peaks_f32(const float* in_dev, float* out_dev, float* tmp1_dev, float* tmp2_dev,
                                     float* tmp3_dev, int64_t nz, int64_t nx, int64_t ny, double sigma,
                                     const float* norm_minmax_dev, float threshold, void* cand_ws_dev,
                                     int64_t cand_cap, int64_t* out_idx_dev, float* out_val_dev, uint32_t* count_dev,
                                     int64_t capacity, void* stream) {
  return mfish_log3d_peaks_range_f32(in_dev, out_dev, tmp1_dev, tmp2_dev, tmp3_dev, nz, nx, ny, sigma,
                                     norm_minmax_dev, 0, nz, threshold, cand_ws_dev, cand_cap, out_idx_dev,
                                     out_val_dev, count_dev, capacity, stream);
}

extern "C" int mfish_log3d_f32(const float* in_dev, float* out_dev, float* tmp1_dev, float* tmp2_dev,
                               float* tmp3_dev, int64_t nz, int64_t nx, int64_t ny, double sigma,
                               const float* norm_minmax_dev, void* stream) {
  return mfish_log3d_range_f32(in_dev, out_dev, tmp1_dev, tmp2_dev, tmp3_dev, nz, nx, ny, sigma, norm_minmax_dev, 0,
                               nz, stream);
}
