// Statistics kernels of find_spots_snrfilter / fix_bleaching (modules/muscle_fish.py:335-400,
// 441-507) and the point look-ups (mask[spot], distance-map gather).
//  * masked_quantile: exact order statistics of the masked voxels by a 3-pass (11/11/10 bit)
//    radix select on the monotone float key: 3 x (R4 + R1) B/voxel, no sort, no host round trip.
//  * blur_at_points : skimage.filters.gaussian(img, (3,3,0.3)) is only ever read at spot voxels
//    (muscle_fish.py:362-365,385-388), so it is evaluated sparsely, one warp per spot, in float64.
#include <algorithm>

#include <stdlib.h>

#include <memory>

#include "common.cuh"

namespace mfish {

// ------------------------------------------------------------------------------------------
// masked radix select
// ------------------------------------------------------------------------------------------
constexpr int QBINS = 2048;
enum { QMODE_QUANTILE = 0, QMODE_BRACKET = 1, QMODE_RANKS = 2 };
struct QState {
  uint32_t prefix[2];
  unsigned long long k[2];
  unsigned long long n;
  double gamma;
  double q;
  int mode;  // how qsel_scan_kernel<0> chooses the two ranks
};
// bracketed select (mfish_masked_quantile_f32): control block shared by its kernels
constexpr int QSEG = 64;  // (layout constant of the control block)
struct QCtl {
  unsigned int seg_count[QSEG * 32];  // [0]: overflow flag of the candidate buffer (rest: reserved)
  unsigned long long n_masked, n_below, n_cand;
  float v_lo, v_hi;    // candidates are the masked values in [v_lo, v_hi]
  int lo_open, hi_open;
  int status;          // 0: exact result delivered, 1: bracket missed, run the three-pass select
  unsigned long long cap;
};
struct QWs {
  uint32_t hist[2][QBINS];
  QState st;
  QCtl ctl;
};
static_assert(sizeof(QWs) <= MFISH_QUANTILE_WS_BYTES, "quantile workspace too small");

template <int PASS>
__device__ __forceinline__ void qsel_digits(uint32_t key, const uint32_t pfx0, const uint32_t pfx1,
                                            int& d0, int& d1) {
  // PASS 0: bits 31..21, PASS 1: bits 20..10, PASS 2: bits 9..0
  if (PASS == 0) {
    d0 = d1 = (int)(key >> 21);
  } else if (PASS == 1) {
    int d = (int)((key >> 10) & 0x7ffu);
    uint32_t hi = key >> 21;
    d0 = hi == pfx0 ? d : -1;
    d1 = hi == pfx1 ? d : -1;
  } else {
    int d = (int)(key & 0x3ffu);
    uint32_t hi = key >> 10;
    d0 = hi == pfx0 ? d : -1;
    d1 = hi == pfx1 ? d : -1;
  }
}

// sample: 0 = every voxel; S > 0 = one 128-byte group (32 voxels) out of every S groups.
// m may be null (every voxel selected); n_dev (optional) caps n by a device-side count.
template <int PASS>
__global__ void __launch_bounds__(256) qsel_hist_kernel(const float* __restrict__ v,
                                                        const uint8_t* __restrict__ m, int64_t n, QWs* ws,
                                                        int sample, const unsigned long long* n_dev) {
  __shared__ uint32_t sh[2][QBINS];
  if (ws->ctl.status != 0 && n_dev != nullptr) return;  // bracket missed: the candidate passes are moot
  if (n_dev != nullptr) n = min((long long)n, (long long)*n_dev);
  for (int i = threadIdx.x; i < 2 * QBINS; i += blockDim.x) (&sh[0][0])[i] = 0;
  __syncthreads();
  const uint32_t pfx0 = PASS == 0 ? 0u : ws->st.prefix[0];
  const uint32_t pfx1 = PASS == 0 ? 0u : ws->st.prefix[1];
  const bool same = (PASS == 0) || (pfx0 == pfx1);
  // run-length aggregation in registers: neighbouring voxels usually fall in the same bin
  int run_d0 = -1, run_d1 = -1;
  uint32_t run_c0 = 0, run_c1 = 0;
  auto push = [&](float val) {
    int d0, d1;
    qsel_digits<PASS>(f32_to_key(val), pfx0, pfx1, d0, d1);
    if (d0 >= 0) {
      if (d0 == run_d0) {
        ++run_c0;
      } else {
        if (run_c0) atomicAdd(&sh[0][run_d0], run_c0);
        run_d0 = d0;
        run_c0 = 1;
      }
    }
    if (!same && d1 >= 0) {
      if (d1 == run_d1) {
        ++run_c1;
      } else {
        if (run_c1) atomicAdd(&sh[1][run_d1], run_c1);
        run_d1 = d1;
        run_c1 = 1;
      }
    }
  };
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool vec = aligned16_dev(v) && ((reinterpret_cast<uintptr_t>(m) & 3u) == 0);
  const int64_t n4 = vec ? n / 4 : 0;
  if (sample > 0 && vec) {
    // groups of 8 float4 (128 B): one pseudo-randomly placed group out of every `sample` consecutive
    // ones (a plain stride would alias with the row length and sample a single column of the volume)
    const int64_t ngroups = (n4 + 7) / 8;
    const int64_t nwin = (ngroups + sample - 1) / sample;
    const int64_t wtid = tid >> 3, wstride = stride >> 3;  // 8 threads share a group
    const int sub = (int)(tid & 7);
    for (int64_t w = wtid; w < nwin; w += wstride) {
      const int64_t g = w * sample + (int64_t)((((uint32_t)w * 0x9E3779B1u) >> 16) % (uint32_t)sample);
      const int64_t i = g * 8 + sub;
      if (i < n4) {
        const uchar4 mk = m ? __ldg(reinterpret_cast<const uchar4*>(m) + i) : make_uchar4(1, 1, 1, 1);
        if (mk.x | mk.y | mk.z | mk.w) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(v) + i);
          if (mk.x) push(a.x);
          if (mk.y) push(a.y);
          if (mk.z) push(a.z);
          if (mk.w) push(a.w);
        }
      }
    }
  } else {
    for (int64_t i = tid; i < n4; i += stride) {
      const uchar4 mk = m ? __ldg(reinterpret_cast<const uchar4*>(m) + i) : make_uchar4(1, 1, 1, 1);
      if (mk.x | mk.y | mk.z | mk.w) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(v) + i);
        if (mk.x) push(a.x);
        if (mk.y) push(a.y);
        if (mk.z) push(a.z);
        if (mk.w) push(a.w);
      }
    }
    for (int64_t i = n4 * 4 + tid; i < n; i += stride) {
      if (!m || m[i]) push(v[i]);
    }
  }
  if (run_c0) atomicAdd(&sh[0][run_d0], run_c0);
  if (run_c1) atomicAdd(&sh[1][run_d1], run_c1);
  __syncthreads();
  for (int i = threadIdx.x; i < QBINS; i += blockDim.x) {
    uint32_t c0 = sh[0][i];
    if (c0) atomicAdd(&ws->hist[0][i], c0);
    if (!same) {
      uint32_t c1 = sh[1][i];
      if (c1) atomicAdd(&ws->hist[1][i], c1);
    }
  }
}

__global__ void qsel_init_kernel(QWs* ws, double q, int mode) {
  for (int i = threadIdx.x; i < 2 * QBINS; i += blockDim.x) (&ws->hist[0][0])[i] = 0;
  if (threadIdx.x == 0) {
    ws->st.prefix[0] = ws->st.prefix[1] = 0;
    ws->st.k[0] = ws->st.k[1] = 0;
    ws->st.n = 0;
    ws->st.gamma = 0.0;
    ws->st.q = q;
    ws->st.mode = mode;
    if (mode != QMODE_RANKS) {
      ws->ctl.n_masked = ws->ctl.n_below = ws->ctl.n_cand = 0;
      ws->ctl.status = 0;
      for (int i = 0; i < QSEG; ++i) ws->ctl.seg_count[i * 32] = 0;
    }
  }
}

// one block of QBINS/8 threads: locate the bins holding ranks k[0], k[1] (block-wide prefix sum
// over the histogram), narrow the prefixes, clear the histogram for the next pass
constexpr int QSCAN_THREADS = QBINS / 8;

template <int PASS>
__global__ void __launch_bounds__(QSCAN_THREADS) qsel_scan_kernel(QWs* ws, float* out, double* info,
                                                                  const unsigned long long* h64 = nullptr) {
  __shared__ unsigned long long s_part[2][QSCAN_THREADS];
  __shared__ unsigned long long s_tot[2];
  __shared__ int s_bin[2];
  __shared__ unsigned long long s_before[2];
  QState& s = ws->st;
  const int t = threadIdx.x;
  const int nb = PASS == 2 ? 1024 : QBINS;
  const bool same = (PASS == 0) || (s.prefix[0] == s.prefix[1]);
  // per-thread sums of 8 consecutive bins, then an inclusive scan of the partials (Hillis-Steele)
  unsigned long long loc[2][8];
  for (int j = 0; j < 2; ++j) {
    const uint32_t* h = ws->hist[(same ? 0 : j)];
    unsigned long long acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int bin = 8 * t + i;
      const unsigned long long c = bin < nb ? (h64 ? h64[(same ? 0 : j) * QBINS + bin] : (unsigned long long)h[bin]) : 0ull;
      loc[j][i] = c;
      acc += c;
    }
    s_part[j][t] = acc;
  }
  __syncthreads();
  for (int o = 1; o < QSCAN_THREADS; o <<= 1) {
    unsigned long long v0 = t >= o ? s_part[0][t - o] : 0ull;
    unsigned long long v1 = t >= o ? s_part[1][t - o] : 0ull;
    __syncthreads();
    s_part[0][t] += v0;
    s_part[1][t] += v1;
    __syncthreads();
  }
  if (t == 0) {
    s_tot[0] = s_part[0][QSCAN_THREADS - 1];
    s_tot[1] = s_part[1][QSCAN_THREADS - 1];
    s_bin[0] = s_bin[1] = nb - 1;  // unreachable default when counts are consistent
    s_before[0] = s_before[1] = 0;
    if (PASS == 0) {
      const unsigned long long n = s_tot[0];
      s.n = n;
      if (s.mode == QMODE_QUANTILE) {
        if (n == 0) {
          out[0] = out[1] = nanf("");
          info[0] = 0.0;
          info[1] = 0.0;
        } else {
          // numpy method 'linear': virtual index = (n - 1) * q
          double vi = (double)(n - 1) * s.q;
          double fl = floor(vi);
          s.gamma = vi - fl;
          long long k0 = (long long)fl;
          k0 = k0 < 0 ? 0 : (k0 > (long long)n - 1 ? (long long)n - 1 : k0);
          long long k1 = k0 + 1 > (long long)n - 1 ? (long long)n - 1 : k0 + 1;
          s.k[0] = (unsigned long long)k0;
          s.k[1] = (unsigned long long)k1;
        }
      } else if (s.mode == QMODE_BRACKET) {
        // ranks of the SAMPLE that bracket the wanted population quantile with a wide margin:
        // 12 standard deviations of the rank of a random sample (cluster sampling is less efficient)
        ws->ctl.lo_open = ws->ctl.hi_open = 1;
        if (n >= 4096) {
          const double r = (double)(n - 1) * s.q;
          const double delta = 12.0 * sqrt((double)n * s.q * (1.0 - s.q)) + 16.0;
          const double lo = floor(r - delta), hi = ceil(r + delta);
          if (lo >= 0.0) {
            s.k[0] = (unsigned long long)lo;
            ws->ctl.lo_open = 0;
          }
          if (hi <= (double)(n - 1)) {
            s.k[1] = (unsigned long long)hi;
            ws->ctl.hi_open = 0;
          } else {
            s.k[1] = n - 1;
          }
        } else if (n > 0) {
          s.k[1] = n - 1;
        }
      }  // QMODE_RANKS: k[0], k[1] were set by qsel_ranks_kernel
    }
  }
  __syncthreads();
  const unsigned long long n_all = s.n;
  if (n_all != 0) {
    for (int j = 0; j < 2; ++j) {
      // this thread owns bins [8t, 8t+8): exclusive prefix = inclusive partial - own sum
      unsigned long long own = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) own += loc[j][i];
      unsigned long long cum = s_part[j][t] - own;
      const unsigned long long k = s.k[j];
      if (k >= cum && k < cum + own) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (k >= cum && k < cum + loc[j][i]) {
            s_bin[j] = 8 * t + i;
            s_before[j] = cum;
          }
          cum += loc[j][i];
        }
      }
    }
  }
  __syncthreads();
  if (t == 0 && n_all != 0) {
    for (int j = 0; j < 2; ++j) {
      const int bsel = s_bin[j];
      s.k[j] -= s_before[j];
      s.prefix[j] = PASS == 0 ? (uint32_t)bsel : ((s.prefix[j] << (PASS == 2 ? 10 : 11)) | (uint32_t)bsel);
    }
    if (PASS == 2) {
      out[0] = key_to_f32(s.prefix[0]);
      out[1] = key_to_f32(s.prefix[1]);
      if (s.mode == QMODE_QUANTILE) {
        info[0] = s.gamma;
        info[1] = (double)s.n;
      }
    }
  }
  __syncthreads();
  for (int i = t; i < 2 * QBINS; i += QSCAN_THREADS) (&ws->hist[0][0])[i] = 0;
}

__global__ void qsel_set_cap_kernel(QWs* ws, unsigned long long cap) { ws->ctl.cap = cap; }

// bracket -> control block (after the sample select)
__global__ void qsel_bracket_kernel(QWs* ws, const float* pair) {
  if (threadIdx.x != 0) return;
  QCtl& c = ws->ctl;
  c.v_lo = c.lo_open ? -INFINITY : pair[0];
  c.v_hi = c.hi_open ? INFINITY : pair[1];
}

// one full pass: count the masked voxels, those below the bracket, and collect those inside it.
// The candidate buffer (pre-filled with a huge value; unused slots sort to the top) is split into one
// private region per warp of the grid, so an append is a register increment: no atomic, and no round
// trip to wait for, on the path of the streaming loads.  A warp that outgrows its region raises the
// overflow flag (seg_count[0]) and the caller falls back to the exact three-pass select.
template <bool MM>
__global__ void __launch_bounds__(256) qsel_range_kernel(const float* __restrict__ v, const uint8_t* __restrict__ m,
                                                         int64_t n, QWs* ws, float* __restrict__ cand, uint32_t* mm) {
  const float lo = ws->ctl.v_lo, hi = ws->ctl.v_hi;
  const unsigned nwarps = gridDim.x * (256 / 32);
  const unsigned seg_cap = (unsigned)(ws->ctl.cap / nwarps);
  float* seg_buf = cand + (size_t)(blockIdx.x * (256 / 32) + (threadIdx.x >> 5)) * seg_cap;
  unsigned wcount = 0;  // candidates this warp has appended (same value in every lane)
  unsigned nm = 0, nb = 0;
  float vmin = INFINITY, vmax = -INFINITY;  // MM: min / max over ALL voxels (su.normalize_image)
  const int lane = threadIdx.x & 31;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool vec = aligned16_dev(v) && ((reinterpret_cast<uintptr_t>(m) & 3u) == 0);
  const int64_t n4 = vec ? n / 4 : 0;
  // uniform control flow: every lane classifies its 4 voxels, then one ballot per voxel slot ranks the
  // warp's candidates behind the warp's running count
  const unsigned lt_mask = (1u << lane) - 1u;
  auto classify = [&](const uchar4 mk, const float4 a, bool have) {
    float val[4] = {a.x, a.y, a.z, a.w};
    unsigned is_c = 0;
    if (have) {
      if (MM) {
        vmin = fminf(fminf(vmin, a.x), fminf(a.y, fminf(a.z, a.w)));
        vmax = fmaxf(fmaxf(vmax, a.x), fmaxf(a.y, fmaxf(a.z, a.w)));
      }
      const unsigned char mb[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {  // branch-free: predicates and selects only
        const unsigned in_mask = mb[k] != 0;
        const unsigned below = in_mask & (unsigned)(val[k] < lo);
        const unsigned inside = in_mask & (unsigned)(val[k] >= lo) & (unsigned)(val[k] <= hi);
        nm += in_mask;
        nb += below;
        is_c |= inside << k;
      }
    }
    if (__any_sync(0xffffffffu, is_c != 0)) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned bk = __ballot_sync(0xffffffffu, (is_c >> k) & 1u);
        if ((is_c >> k) & 1u) {
          const unsigned pos = wcount + __popc(bk & lt_mask);
          if (pos < seg_cap) seg_buf[pos] = val[k];
        }
        wcount += __popc(bk);
      }
    }
  };
  // QGROUPS independent groups of 4 voxels per iteration: every load is in flight before the first is
  // used.  Without MM the voxels of an unmasked group are never fetched.
  constexpr int QGROUPS = 2;
  const int64_t n4_warp = (n4 + 31) / 32 * 32;  // keep whole warps in the loop for the shuffles
  for (int64_t i = tid; i < n4_warp; i += QGROUPS * stride) {
    uchar4 mk[QGROUPS];
    float4 a[QGROUPS];
    bool f[QGROUPS];
#pragma unroll
    for (int g = 0; g < QGROUPS; ++g) {
      const int64_t j = i + g * stride;
      mk[g] = make_uchar4(0, 0, 0, 0);
      if (j < n4) mk[g] = __ldg(reinterpret_cast<const uchar4*>(m) + j);
    }
#pragma unroll
    for (int g = 0; g < QGROUPS; ++g) {
      const int64_t j = i + g * stride;
      f[g] = j < n4 && (MM || (mk[g].x | mk[g].y | mk[g].z | mk[g].w));
      a[g] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f[g]) a[g] = __ldg(reinterpret_cast<const float4*>(v) + j);
    }
#pragma unroll
    for (int g = 0; g < QGROUPS; ++g)
      if (i + g * stride < n4_warp) classify(mk[g], a[g], f[g]);
  }
  // scalar tail (up to 3 voxels, or the whole of an unaligned volume), whole warps again
  const int64_t tail0 = n4 * 4;
  const int64_t ntail = n - tail0;
  const int64_t ntail_warp = (ntail + 31) / 32 * 32;
  for (int64_t r = tid; r < ntail_warp; r += stride) {
    bool c = false;
    float x = 0.f;
    if (r < ntail) {
      x = v[tail0 + r];
      if (MM) {
        vmin = fminf(vmin, x);
        vmax = fmaxf(vmax, x);
      }
      if (m[tail0 + r]) {
        ++nm;
        if (x < lo) ++nb;
        else if (x <= hi) c = true;
      }
    }
    const unsigned bk = __ballot_sync(0xffffffffu, c);
    if (c) {
      const unsigned pos = wcount + __popc(bk & lt_mask);
      if (pos < seg_cap) seg_buf[pos] = x;
    }
    wcount += __popc(bk);
  }
  if (lane == 0 && wcount != 0) {
    atomicAdd(&ws->ctl.n_cand, (unsigned long long)min(wcount, seg_cap));
    if (wcount > seg_cap) atomicOr(&ws->ctl.seg_count[0], 1u);
  }
  // block reduction of the two counters, one atomic pair per block
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nm += __shfl_xor_sync(0xffffffffu, nm, o);
    nb += __shfl_xor_sync(0xffffffffu, nb, o);
  }
  __shared__ unsigned s_nm[8], s_nb[8];
  if ((threadIdx.x & 31) == 0) {
    s_nm[threadIdx.x >> 5] = nm;
    s_nb[threadIdx.x >> 5] = nb;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long tm = 0, tb = 0;
    for (int k = 0; k < 8; ++k) {
      tm += s_nm[k];
      tb += s_nb[k];
    }
    atomicAdd(&ws->ctl.n_masked, tm);
    atomicAdd(&ws->ctl.n_below, tb);
  }
  if (MM) {
    __syncthreads();
    block_minmax_commit(vmin, vmax, mm);
  }
}

__global__ void qsel_mm_init_kernel(uint32_t* mm) {
  mm[0] = 0xFFFFFFFFu;
  mm[1] = 0u;
}
__global__ void qsel_mm_finalize_kernel(uint32_t* mm) {
  const float lo = key_to_f32(mm[0]);
  const float hi = key_to_f32(mm[1]);
  reinterpret_cast<float*>(mm)[0] = lo;
  reinterpret_cast<float*>(mm)[1] = hi;
}

// exact ranks of the population quantile, relative to the candidate buffer; verifies the bracket
// (counts: {n_masked, n_below, n_cand, overflow} summed over the ranks of a split volume, else null)
__global__ void qsel_ranks_kernel(QWs* ws, float* out, double* info, const long long* counts = nullptr) {
  if (threadIdx.x != 0) return;
  QCtl& c = ws->ctl;
  QState& s = ws->st;
  if (counts) {
    c.n_masked = (unsigned long long)counts[0];
    c.n_below = (unsigned long long)counts[1];
    c.n_cand = (unsigned long long)counts[2];
    c.seg_count[0] = counts[3] != 0 ? 1u : 0u;
  }
  const unsigned long long n = c.n_masked;
  info[1] = (double)n;
  const bool overflow = c.seg_count[0] != 0;  // a warp outgrew its region of the candidate buffer
  if (n == 0) {
    out[0] = out[1] = nanf("");
    info[0] = 0.0;
    info[2] = 0.0;
    c.status = 0;
    s.k[0] = s.k[1] = 0;
    return;
  }
  const double vi = (double)(n - 1) * s.q;
  const double fl = floor(vi);
  long long k0 = (long long)fl;
  k0 = k0 < 0 ? 0 : (k0 > (long long)n - 1 ? (long long)n - 1 : k0);
  const long long k1 = k0 + 1 > (long long)n - 1 ? (long long)n - 1 : k0 + 1;
  s.gamma = vi - fl;
  info[0] = s.gamma;
  const bool inside = !overflow && (unsigned long long)k0 >= c.n_below &&
                      (unsigned long long)k1 < c.n_below + c.n_cand;
  c.status = inside ? 0 : 1;
  info[2] = inside ? 0.0 : 1.0;
  s.k[0] = inside ? (unsigned long long)k0 - c.n_below : 0;
  s.k[1] = inside ? (unsigned long long)k1 - c.n_below : 0;
}

// ------------------------------------------------------------------------------------------
// sparse Gaussian blur at points, float64
// ------------------------------------------------------------------------------------------
struct BlurArgs {
  const float* vol;
  int nz, nx, ny;
  const int32_t* pts;  // n x 3 (z, x, y)
  int64_t npts;
  const double* tz;
  const double* tx;
  const double* ty;
  int rz, rx, ry;
  int bmode;
  const float* mm;
  double* out;
};

__global__ void __launch_bounds__(128) blur_points_kernel(const BlurArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= a.npts) return;
  const int z = a.pts[3 * p], x = a.pts[3 * p + 1], y = a.pts[3 * p + 2];
  double mn = 0.0, slope = 1.0;
  bool degen = false;
  if (a.mm) {
    double lo = (double)a.mm[0], hi = (double)a.mm[1];
    mn = lo;
    if (hi > lo) slope = 1.0 / (hi - lo); else degen = true;
  }
  const int wz = 2 * a.rz + 1, wx = 2 * a.rx + 1, wy = 2 * a.ry + 1;
  const int total = wz * wx * wy;
  double acc = 0.0;
  for (int t = lane; t < total; t += 32) {
    int dy = t % wy - a.ry;
    int r = t / wy;
    int dx = r % wx - a.rx;
    int dz = r / wx - a.rz;
    int zz, xx, yy;
    if (a.bmode == MFISH_BOUNDARY_NEAREST) {
      zz = bmap<MFISH_BOUNDARY_NEAREST>(z + dz, a.nz);
      xx = bmap<MFISH_BOUNDARY_NEAREST>(x + dx, a.nx);
      yy = bmap<MFISH_BOUNDARY_NEAREST>(y + dy, a.ny);
    } else {
      zz = bmap<MFISH_BOUNDARY_REFLECT>(z + dz, a.nz);
      xx = bmap<MFISH_BOUNDARY_REFLECT>(x + dx, a.nx);
      yy = bmap<MFISH_BOUNDARY_REFLECT>(y + dy, a.ny);
    }
    double w = a.tz[abs(dz)] * a.tx[abs(dx)] * a.ty[abs(dy)];
    double v = (double)__ldg(&a.vol[((int64_t)zz * a.nx + xx) * a.ny + yy]);
    v = degen ? 1.0 : (v - mn) * slope;
    acc = fma(w, v, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) a.out[p] = acc;
}

// ------------------------------------------------------------------------------------------
// gathers, plane sums, plane scaling
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void gather_kernel(const T* __restrict__ vol, int nz, int nx, int ny, const int32_t* pts,
                              int64_t n, T* out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int z = pts[3 * i], x = pts[3 * i + 1], y = pts[3 * i + 2];
  bool ok = z >= 0 && z < nz && x >= 0 && x < nx && y >= 0 && y < ny;
  out[i] = ok ? vol[((int64_t)z * nx + x) * ny + y] : T(0);
}

__global__ void __launch_bounds__(256) plane_sums_kernel(const float* __restrict__ v,
                                                         const uint8_t* __restrict__ m, int64_t plane,
                                                         double* sums, double* counts) {
  const int z = blockIdx.y;
  const float* pv = v + (int64_t)z * plane;
  const uint8_t* pm = m ? m + (int64_t)z * plane : nullptr;
  double s = 0.0;
  unsigned long long c = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (!pm || pm[i]) {
      s += (double)pv[i];
      ++c;
    }
  }
  double cd = (double)c;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    cd += __shfl_xor_sync(0xffffffffu, cd, o);
  }
  __shared__ double ss[8], sc[8];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
    ss[w] = s;
    sc[w] = cd;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tc = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
      ts += ss[i];
      tc += sc[i];
    }
    atomicAdd(&sums[z], ts);
    atomicAdd(&counts[z], tc);
  }
}

__global__ void __launch_bounds__(256) scale_planes_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                           int64_t plane, const float* __restrict__ f) {
  const int z = blockIdx.y;
  const float fz = f[z];
  const float* pi = in + (int64_t)z * plane;
  float* po = out + (int64_t)z * plane;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane;
       i += (int64_t)gridDim.x * blockDim.x)
    po[i] = pi[i] * fz;
}

}  // namespace mfish

using namespace mfish;

// The select in its three steps, so that a z-slab decomposed volume can all-reduce the histogram
// (first 2*2048 uint32 of the workspace) between `hist` and `scan` of every pass.
extern "C" int mfish_qsel_begin(void* ws_dev, double q, void* stream) {
  MFISH_REQUIRE(ws_dev && q >= 0.0 && q <= 1.0, "qsel_begin: bad arguments");
  qsel_init_kernel<<<1, 256, 0, as_stream(stream)>>>(reinterpret_cast<QWs*>(ws_dev), q, QMODE_QUANTILE);
  count_launch(1);
  return check_launch("qsel_begin");
}

extern "C" int mfish_qsel_hist(const float* vol_dev, const uint8_t* mask_dev, int64_t n, int pass, void* ws_dev,
                               void* stream) {
  MFISH_REQUIRE(vol_dev && mask_dev && ws_dev && n > 0, "qsel_hist: bad arguments");
  MFISH_REQUIRE(pass >= 0 && pass <= 2, "qsel_hist: pass must be 0, 1 or 2");
  cudaStream_t st = as_stream(stream);
  QWs* ws = reinterpret_cast<QWs*>(ws_dev);
  int blocks = (int)std::min<int64_t>((n / 4 + 255) / 256 + 1, (int64_t)num_sms() * 8);
  if (pass == 0)
    qsel_hist_kernel<0><<<blocks, 256, 0, st>>>(vol_dev, mask_dev, n, ws, 0, nullptr);
  else if (pass == 1)
    qsel_hist_kernel<1><<<blocks, 256, 0, st>>>(vol_dev, mask_dev, n, ws, 0, nullptr);
  else
    qsel_hist_kernel<2><<<blocks, 256, 0, st>>>(vol_dev, mask_dev, n, ws, 0, nullptr);
  count_launch(1);
  return check_launch("qsel_hist");
}

extern "C" int mfish_qsel_scan(int pass, void* ws_dev, float* out_dev, double* info_dev, void* stream) {
  MFISH_REQUIRE(ws_dev && out_dev && info_dev, "qsel_scan: null pointer");
  MFISH_REQUIRE(pass >= 0 && pass <= 2, "qsel_scan: pass must be 0, 1 or 2");
  cudaStream_t st = as_stream(stream);
  QWs* ws = reinterpret_cast<QWs*>(ws_dev);
  if (pass == 0)
    qsel_scan_kernel<0><<<1, QSCAN_THREADS, 0, st>>>(ws, out_dev, info_dev);
  else if (pass == 1)
    qsel_scan_kernel<1><<<1, QSCAN_THREADS, 0, st>>>(ws, out_dev, info_dev);
  else
    qsel_scan_kernel<2><<<1, QSCAN_THREADS, 0, st>>>(ws, out_dev, info_dev);
  count_launch(1);
  return check_launch("qsel_scan");
}

extern "C" int mfish_qsel_scan_u64(int pass, void* ws_dev, const unsigned long long* hist64_dev, float* out_dev,
                                  double* info_dev, void* stream) {
  MFISH_REQUIRE(ws_dev && out_dev && info_dev && hist64_dev, "qsel_scan_u64: null pointer");
  MFISH_REQUIRE(pass >= 0 && pass <= 2, "qsel_scan_u64: pass must be 0, 1 or 2");
  cudaStream_t st = as_stream(stream);
  QWs* ws = reinterpret_cast<QWs*>(ws_dev);
  if (pass == 0)
    qsel_scan_kernel<0><<<1, QSCAN_THREADS, 0, st>>>(ws, out_dev, info_dev, hist64_dev);
  else if (pass == 1)
    qsel_scan_kernel<1><<<1, QSCAN_THREADS, 0, st>>>(ws, out_dev, info_dev, hist64_dev);
  else
    qsel_scan_kernel<2><<<1, QSCAN_THREADS, 0, st>>>(ws, out_dev, info_dev, hist64_dev);
  count_launch(1);
  return check_launch("qsel_scan_u64");
}

// ---- the bracketed select in steps, for a volume split across ranks (see include/mfish.h) ----
__global__ void qsel_export_counts_kernel(const QWs* ws, long long* counts) {
  counts[0] = (long long)ws->ctl.n_masked;
  counts[1] = (long long)ws->ctl.n_below;
  counts[2] = (long long)ws->ctl.n_cand;
  counts[3] = ws->ctl.seg_count[0] != 0 ? 1 : 0;
}

extern "C" int mfish_qselb_begin(void* ws_dev, double q, void* stream) {
  MFISH_REQUIRE(ws_dev && q >= 0.0 && q <= 1.0, "qselb_begin: bad arguments");
  qsel_init_kernel<<<1, 256, 0, as_stream(stream)>>>(reinterpret_cast<QWs*>(ws_dev), q, QMODE_BRACKET);
  count_launch(1);
  return check_launch("qselb_begin");
}

extern "C" int mfish_qselb_sample_hist(const float* vol_dev, const uint8_t* mask_dev, int64_t n, int pass,
                                       void* ws_dev, void* stream) {
  MFISH_REQUIRE(vol_dev && mask_dev && ws_dev && n > 0, "qselb_sample_hist: bad arguments");
  MFISH_REQUIRE(pass >= 0 && pass <= 2, "qselb_sample_hist: pass must be 0, 1 or 2");
  cudaStream_t st = as_stream(stream);
  QWs* ws = reinterpret_cast<QWs*>(ws_dev);
  constexpr int SAMPLE = 64;
  const int sblocks = (int)std::min<int64_t>((n / 4 / 64 + 255) / 256 + 1, (int64_t)num_sms() * 8);
  if (pass == 0)
    qsel_hist_kernel<0><<<sblocks, 256, 0, st>>>(vol_dev, mask_dev, n, ws, SAMPLE, nullptr);
  else if (pass == 1)
    qsel_hist_kernel<1><<<sblocks, 256, 0, st>>>(vol_dev, mask_dev, n, ws, SAMPLE, nullptr);
  else
    qsel_hist_kernel<2><<<sblocks, 256, 0, st>>>(vol_dev, mask_dev, n, ws, SAMPLE, nullptr);
  count_launch(1);
  return check_launch("qselb_sample_hist");
}

extern "C" int mfish_qselb_collect(const float* vol_dev, const uint8_t* mask_dev, int64_t n, void* ws_dev,
                                   const float* pair_dev, float* cand_dev, int64_t cap, long long* counts_dev,
                                   float* minmax_dev, void* stream) {
  MFISH_REQUIRE(vol_dev && mask_dev && ws_dev && pair_dev && cand_dev && counts_dev, "qselb_collect: null pointer");
  MFISH_REQUIRE(n > 0 && cap >= QSEG && cap % QSEG == 0, "qselb_collect: bad sizes");
  cudaStream_t st = as_stream(stream);
  QWs* ws = reinterpret_cast<QWs*>(ws_dev);
  ProfScope prof("masked_quantile_collect", st);
  const int blocks = (int)std::min<int64_t>((n / 4 + 255) / 256 + 1, (int64_t)num_sms() * 8);
  qsel_bracket_kernel<<<1, 32, 0, st>>>(ws, pair_dev);
  MFISH_CUDA(cudaMemsetAsync(cand_dev, 0x7f, (size_t)cap * sizeof(float), st));
  qsel_set_cap_kernel<<<1, 1, 0, st>>>(ws, (unsigned long long)cap);
  if (minmax_dev) {  // the slab's own {min, max} over ALL voxels from the same pass
    uint32_t* mm = reinterpret_cast<uint32_t*>(minmax_dev);
    qsel_mm_init_kernel<<<1, 1, 0, st>>>(mm);
    qsel_range_kernel<true><<<blocks, 256, 0, st>>>(vol_dev, mask_dev, n, ws, cand_dev, mm);
    qsel_mm_finalize_kernel<<<1, 1, 0, st>>>(mm);
    count_launch(2);
  } else {
    qsel_range_kernel<false><<<blocks, 256, 0, st>>>(vol_dev, mask_dev, n, ws, cand_dev, nullptr);
  }
  qsel_export_counts_kernel<<<1, 1, 0, st>>>(ws, counts_dev);
  count_launch(4);
  return check_launch("qselb_collect");
}

extern "C" int mfish_qselb_ranks(void* ws_dev, double q, const long long* counts_dev, float* out_dev,
                                 double* info_dev, void* stream) {
  MFISH_REQUIRE(ws_dev && counts_dev && out_dev && info_dev, "qselb_ranks: null pointer");
  cudaStream_t st = as_stream(stream);
  QWs* ws = reinterpret_cast<QWs*>(ws_dev);
  qsel_init_kernel<<<1, 256, 0, st>>>(ws, q, QMODE_RANKS);
  qsel_ranks_kernel<<<1, 32, 0, st>>>(ws, out_dev, info_dev, counts_dev);
  count_launch(2);
  return check_launch("qselb_ranks");
}

extern "C" int mfish_qselb_cand_hist(const float* cand_dev, int64_t cap, int pass, void* ws_dev, void* stream) {
  MFISH_REQUIRE(cand_dev && ws_dev && cap > 0, "qselb_cand_hist: bad arguments");
  MFISH_REQUIRE(pass >= 0 && pass <= 2, "qselb_cand_hist: pass must be 0, 1 or 2");
  cudaStream_t st = as_stream(stream);
  QWs* ws = reinterpret_cast<QWs*>(ws_dev);
  const int cblocks = (int)std::min<int64_t>((cap / 4 + 255) / 256 + 1, (int64_t)num_sms() * 4);
  if (pass == 0)
    qsel_hist_kernel<0><<<cblocks, 256, 0, st>>>(cand_dev, nullptr, cap, ws, 0, &ws->ctl.cap);
  else if (pass == 1)
    qsel_hist_kernel<1><<<cblocks, 256, 0, st>>>(cand_dev, nullptr, cap, ws, 0, &ws->ctl.cap);
  else
    qsel_hist_kernel<2><<<cblocks, 256, 0, st>>>(cand_dev, nullptr, cap, ws, 0, &ws->ctl.cap);
  count_launch(1);
  return check_launch("qselb_cand_hist");
}

extern "C" int mfish_masked_quantile_f32(const float* vol_dev, const uint8_t* mask_dev, int64_t n, double q,
                                         void* ws_dev, float* out_dev, double* info_dev, void* stream) {
  MFISH_REQUIRE(vol_dev && mask_dev && ws_dev && out_dev && info_dev, "masked_quantile: null pointer");
  MFISH_REQUIRE(n > 0 && q >= 0.0 && q <= 1.0, "masked_quantile: bad arguments");
  ProfScope prof("masked_quantile_3pass", as_stream(stream));
  int rc = mfish_qsel_begin(ws_dev, q, stream);
  for (int pass = 0; pass < 3 && rc == MFISH_OK; ++pass) {
    rc = mfish_qsel_hist(vol_dev, mask_dev, n, pass, ws_dev, stream);
    if (rc == MFISH_OK) rc = mfish_qsel_scan(pass, ws_dev, out_dev, info_dev, stream);
  }
  return rc;
}

// Bracketed select: the same two order statistics from ONE full pass over the volume.
//   1. three-pass select on a 1/64 sample (128-byte groups) -> values [v_lo, v_hi] that bracket the
//      wanted ranks with a 12-sigma margin;
//   2. one full pass counts the masked voxels below v_lo and collects those inside the bracket;
//   3. the exact ranks, shifted by the count below, are selected inside the (small) candidate buffer.
// The bracket is verified on the device; info_dev[2] = 1 means it missed (or the buffer overflowed) and
// the caller must run mfish_masked_quantile_f32.  info_dev = {gamma, n_masked, status}.
static int quantile_fast_impl(const float* vol_dev, const uint8_t* mask_dev, int64_t n, double q, void* ws_dev,
                              float* out_dev, double* info_dev, float* minmax_dev, void* stream) {
  MFISH_REQUIRE(vol_dev && mask_dev && ws_dev && out_dev && info_dev, "masked_quantile: null pointer");
  MFISH_REQUIRE(n > 0 && q >= 0.0 && q <= 1.0, "masked_quantile: bad arguments");
  cudaStream_t st = as_stream(stream);
  // three timing scopes: the sample's select (1/64 of the volume, three times), THE full pass (5 B/voxel: the
  // kernel the roofline is about), the select among the collected candidates
  std::unique_ptr<ProfScope> prof(new ProfScope("masked_quantile_sample", st));
  uint32_t* mm = reinterpret_cast<uint32_t*>(minmax_dev);
  QWs* ws = reinterpret_cast<QWs*>(ws_dev);
  const int64_t cap = (std::max<int64_t>(1 << 20, n / 64) + QSEG - 1) / QSEG * QSEG;
  float* cand = nullptr;
  float* pair = nullptr;
  MFISH_CUDA(cudaMallocAsync((void**)&cand, (size_t)cap * sizeof(float), st));
  MFISH_CUDA(cudaMallocAsync((void**)&pair, 2 * sizeof(float) + 3 * sizeof(double), st));
  double* dummy = reinterpret_cast<double*>(pair + 2);
  static const int range_bps = [] {  // blocks per SM of the one full pass (diagnostic switch)
    const char* e = getenv("MFISH_QSEL_BPS");
    const int v = e ? atoi(e) : 0;
    return v > 0 && v <= 64 ? v : 8;
  }();
  const int blocks = (int)std::min<int64_t>((n / 4 + 255) / 256 + 1, (int64_t)num_sms() * range_bps);
  const int sblocks = (int)std::min<int64_t>((n / 4 / 64 + 255) / 256 + 1, (int64_t)num_sms() * 8);
  constexpr int SAMPLE = 64;
  // 1. bracket from the sample
  qsel_init_kernel<<<1, 256, 0, st>>>(ws, q, QMODE_BRACKET);
  qsel_hist_kernel<0><<<sblocks, 256, 0, st>>>(vol_dev, mask_dev, n, ws, SAMPLE, nullptr);
  qsel_scan_kernel<0><<<1, QSCAN_THREADS, 0, st>>>(ws, pair, dummy);
  qsel_hist_kernel<1><<<sblocks, 256, 0, st>>>(vol_dev, mask_dev, n, ws, SAMPLE, nullptr);
  qsel_scan_kernel<1><<<1, QSCAN_THREADS, 0, st>>>(ws, pair, dummy);
  qsel_hist_kernel<2><<<sblocks, 256, 0, st>>>(vol_dev, mask_dev, n, ws, SAMPLE, nullptr);
  qsel_scan_kernel<2><<<1, QSCAN_THREADS, 0, st>>>(ws, pair, dummy);
  qsel_bracket_kernel<<<1, 32, 0, st>>>(ws, pair);
  // 2. the one full pass
  cudaMemsetAsync(cand, 0x7f, (size_t)cap * sizeof(float), st);  // 3.39e38: sorts above any voxel
  qsel_set_cap_kernel<<<1, 1, 0, st>>>(ws, (unsigned long long)cap);  // (a pageable H2D copy would stall the stream)
  prof.reset();
  prof.reset(new ProfScope(minmax_dev ? "masked_quantile_minmax" : "masked_quantile_bracketed", st));
  if (mm) {
    qsel_mm_init_kernel<<<1, 1, 0, st>>>(mm);
    qsel_range_kernel<true><<<blocks, 256, 0, st>>>(vol_dev, mask_dev, n, ws, cand, mm);
    qsel_mm_finalize_kernel<<<1, 1, 0, st>>>(mm);
    count_launch(2);
  } else {
    qsel_range_kernel<false><<<blocks, 256, 0, st>>>(vol_dev, mask_dev, n, ws, cand, nullptr);
  }
  // 3. exact ranks inside the candidates
  prof.reset();
  prof.reset(new ProfScope("masked_quantile_select", st));
  qsel_init_kernel<<<1, 256, 0, st>>>(ws, q, QMODE_RANKS);
  qsel_ranks_kernel<<<1, 32, 0, st>>>(ws, out_dev, info_dev);
  const int cblocks = (int)std::min<int64_t>((cap / 4 + 255) / 256 + 1, (int64_t)num_sms() * 4);
  qsel_hist_kernel<0><<<cblocks, 256, 0, st>>>(cand, nullptr, cap, ws, 0, &ws->ctl.cap);
  qsel_scan_kernel<0><<<1, QSCAN_THREADS, 0, st>>>(ws, out_dev, info_dev);
  qsel_hist_kernel<1><<<cblocks, 256, 0, st>>>(cand, nullptr, cap, ws, 0, &ws->ctl.cap);
  qsel_scan_kernel<1><<<1, QSCAN_THREADS, 0, st>>>(ws, out_dev, info_dev);
  qsel_hist_kernel<2><<<cblocks, 256, 0, st>>>(cand, nullptr, cap, ws, 0, &ws->ctl.cap);
  qsel_scan_kernel<2><<<1, QSCAN_THREADS, 0, st>>>(ws, out_dev, info_dev);
  count_launch(17);
  prof.reset();
  int rc = check_launch("masked_quantile_fast");
  cudaFreeAsync(cand, st);
  cudaFreeAsync(pair, st);
  return rc;
}

extern "C" int mfish_masked_quantile_fast_f32(const float* vol_dev, const uint8_t* mask_dev, int64_t n, double q,
                                              void* ws_dev, float* out_dev, double* info_dev, void* stream) {
  return quantile_fast_impl(vol_dev, mask_dev, n, q, ws_dev, out_dev, info_dev, nullptr, stream);
}

// Same selection; the one full pass over the voxels also reduces the global {min, max} of ALL voxels
// (su.normalize_image's bounds) into minmax_dev, saving the separate pass of mfish_minmax_f32.  The
// pair is valid whatever the bracket status says.
extern "C" int mfish_masked_quantile_minmax_f32(const float* vol_dev, const uint8_t* mask_dev, int64_t n, double q,
                                                void* ws_dev, float* out_dev, double* info_dev, float* minmax_dev,
                                                void* stream) {
  MFISH_REQUIRE(minmax_dev != nullptr, "masked_quantile_minmax: null minmax");
  return quantile_fast_impl(vol_dev, mask_dev, n, q, ws_dev, out_dev, info_dev, minmax_dev, stream);
}

extern "C" int mfish_blur_at_points(const float* vol_dev, int64_t nz, int64_t nx, int64_t ny,
                                    const int32_t* pts_dev, int64_t npts, const double* taps_z_host, int rz,
                                    const double* taps_x_host, int rx, const double* taps_y_host, int ry,
                                    int boundary, const float* norm_minmax_dev, double* out_dev, void* stream) {
  MFISH_REQUIRE(npts >= 0, "blur_at_points: negative npts");
  if (npts == 0) return MFISH_OK;
  MFISH_REQUIRE(vol_dev && pts_dev && out_dev && taps_z_host && taps_x_host && taps_y_host,
                "blur_at_points: null pointer");
  MFISH_REQUIRE(rz >= 0 && rx >= 0 && ry >= 0 && rz <= 4096 && rx <= 4096 && ry <= 4096, "blur_at_points: bad radius");
  cudaStream_t st = as_stream(stream);
  int nt = (rz + 1) + (rx + 1) + (ry + 1);
  std::vector<double> h(nt);
  std::copy(taps_z_host, taps_z_host + rz + 1, h.begin());
  std::copy(taps_x_host, taps_x_host + rx + 1, h.begin() + rz + 1);
  std::copy(taps_y_host, taps_y_host + ry + 1, h.begin() + rz + 1 + rx + 1);
  double* d = nullptr;
  MFISH_CUDA(cudaMallocAsync((void**)&d, nt * sizeof(double), st));
  MFISH_CUDA(cudaMemcpyAsync(d, h.data(), nt * sizeof(double), cudaMemcpyHostToDevice, st));
  BlurArgs a{vol_dev, (int)nz, (int)nx, (int)ny, pts_dev, npts, d, d + rz + 1, d + rz + 1 + rx + 1,
             rz, rx, ry, boundary, norm_minmax_dev, out_dev};
  ProfScope prof("blur_at_points", st);
  blur_points_kernel<<<(unsigned)((npts + 3) / 4), 128, 0, st>>>(a);
  count_launch(1);
  int rc = check_launch("blur_at_points");
  cudaFreeAsync(d, st);
  return rc;
}

extern "C" int mfish_gather_f32(const float* vol_dev, int64_t nz, int64_t nx, int64_t ny, const int32_t* pts_dev,
                                int64_t npts, float* out_dev, void* stream) {
  if (npts == 0) return MFISH_OK;
  MFISH_REQUIRE(vol_dev && pts_dev && out_dev && npts > 0, "gather: bad arguments");
  gather_kernel<float><<<(unsigned)((npts + 255) / 256), 256, 0, as_stream(stream)>>>(
      vol_dev, (int)nz, (int)nx, (int)ny, pts_dev, npts, out_dev);
  count_launch(1);
  return check_launch("gather_f32");
}

extern "C" int mfish_gather_u8(const uint8_t* vol_dev, int64_t nz, int64_t nx, int64_t ny, const int32_t* pts_dev,
                               int64_t npts, uint8_t* out_dev, void* stream) {
  if (npts == 0) return MFISH_OK;
  MFISH_REQUIRE(vol_dev && pts_dev && out_dev && npts > 0, "gather: bad arguments");
  gather_kernel<uint8_t><<<(unsigned)((npts + 255) / 256), 256, 0, as_stream(stream)>>>(
      vol_dev, (int)nz, (int)nx, (int)ny, pts_dev, npts, out_dev);
  count_launch(1);
  return check_launch("gather_u8");
}

extern "C" int mfish_plane_sums_f32(const float* vol_dev, const uint8_t* mask_dev, int64_t nz, int64_t nx,
                                    int64_t ny, double* sums_dev, double* counts_dev, void* stream) {
  MFISH_REQUIRE(vol_dev && sums_dev && counts_dev && nz > 0 && nx > 0 && ny > 0, "plane_sums: bad arguments");
  MFISH_REQUIRE(nz < 65536, "plane_sums: too many planes");
  cudaStream_t st = as_stream(stream);
  MFISH_CUDA(cudaMemsetAsync(sums_dev, 0, nz * sizeof(double), st));
  MFISH_CUDA(cudaMemsetAsync(counts_dev, 0, nz * sizeof(double), st));
  int64_t plane = nx * ny;
  int bx = (int)std::min<int64_t>((plane + 255) / 256, 64);
  plane_sums_kernel<<<dim3(bx, (unsigned)nz), 256, 0, st>>>(vol_dev, mask_dev, plane, sums_dev, counts_dev);
  count_launch(1);
  return check_launch("plane_sums");
}

extern "C" int mfish_scale_planes_f32(const float* in_dev, float* out_dev, int64_t nz, int64_t nx, int64_t ny,
                                      const float* factor_dev, void* stream) {
  MFISH_REQUIRE(in_dev && out_dev && factor_dev && nz > 0 && nx > 0 && ny > 0, "scale_planes: bad arguments");
  MFISH_REQUIRE(nz < 65536, "scale_planes: too many planes");
  int64_t plane = nx * ny;
  int bx = (int)std::min<int64_t>((plane + 255) / 256, 256);
  scale_planes_kernel<<<dim3(bx, (unsigned)nz), 256, 0, as_stream(stream)>>>(in_dev, out_dev, plane, factor_dev);
  count_launch(1);
  return check_launch("scale_planes");
}
