// Threshold statistics of the segmentation steps (modules/muscle_fish.py:165-198, 238-272):
//   skimage.filters.threshold_otsu(img_blur)  -> 256-bin np.histogram of the blurred volume
//   skimage.filters.threshold_li(img_blur)    -> iterated means of the voxels below / above a threshold
//   np.where(img > thresh, 1, 0)              -> binarisation
// One streaming read of the float32 volume each (4 B/voxel).  The bin of a voxel is computed with
// np.histogram's own float64 arithmetic (uniform-bin fast path: truncation of (x - first) / (last - first) * nbins,
// then the two corrections against the linspace edges), so a given volume lands in the same bins as on the host.
#include <algorithm>

#include "common.cuh"

namespace mfish {

constexpr int SEG_THREADS = 256;
constexpr int SEG_MAX_BINS = 1024;

// np.histogram(a, bins=nbins, range=(first, last)) for first < last; edges_dev = np.linspace(first, last, nbins + 1)
__global__ void __launch_bounds__(SEG_THREADS) hist_kernel(const float* __restrict__ v, int64_t n, double first,
                                                           double last, int nbins, const double* __restrict__ edges,
                                                           unsigned long long* __restrict__ hist) {
  extern __shared__ unsigned int sh[];  // [warps][nbins]
  const int nw = SEG_THREADS / 32;
  for (int i = threadIdx.x; i < nw * nbins; i += SEG_THREADS) sh[i] = 0u;
  __syncthreads();
  unsigned int* mine = sh + (threadIdx.x >> 5) * nbins;
  const double denom = last - first;
  const int64_t stride = (int64_t)gridDim.x * SEG_THREADS;
  for (int64_t i = (int64_t)blockIdx.x * SEG_THREADS + threadIdx.x; i < n; i += stride) {
    const double x = (double)__ldg(v + i);
    if (!(x >= first && x <= last)) continue;  // also drops NaN, as np.histogram's keep mask does
    int b = (int)(((x - first) / denom) * (double)nbins);
    if (b == nbins) --b;
    if (x < edges[b]) --b;
    if (b != nbins - 1 && x >= edges[b + 1]) ++b;
    atomicAdd(&mine[b], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < nbins; b += SEG_THREADS) {
    unsigned long long s = 0;
    for (int w = 0; w < nw; ++w) s += sh[w * nbins + b];
    if (s) atomicAdd(&hist[b], s);
  }
}

// {count, sum} of the voxels with v <= thr and with v > thr, in float64 (NaN voxels count nowhere)
__global__ void __launch_bounds__(SEG_THREADS) split_moments_kernel(const float* __restrict__ v, int64_t n, float thr,
                                                                    double* __restrict__ out) {
  double c0 = 0, s0 = 0, c1 = 0, s1 = 0;
  const int64_t stride = (int64_t)gridDim.x * SEG_THREADS;
  for (int64_t i = (int64_t)blockIdx.x * SEG_THREADS + threadIdx.x; i < n; i += stride) {
    const float x = __ldg(v + i);
    if (x <= thr) {
      c0 += 1.0;
      s0 += (double)x;
    } else if (x > thr) {
      c1 += 1.0;
      s1 += (double)x;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c0 += __shfl_xor_sync(0xffffffffu, c0, o);
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    c1 += __shfl_xor_sync(0xffffffffu, c1, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  __shared__ double red[4][SEG_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
    red[0][w] = c0;
    red[1][w] = s0;
    red[2][w] = c1;
    red[3][w] = s1;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0;
    for (int k = 0; k < SEG_THREADS / 32; ++k) t += red[threadIdx.x][k];
    atomicAdd(&out[threadIdx.x], t);
  }
}

// out = (v > thr), optionally of the normalised value (v - min) / (max - min)
__global__ void __launch_bounds__(SEG_THREADS) binarize_kernel(const float* __restrict__ v, int64_t n, float thr,
                                                               const float* __restrict__ mm,
                                                               uint8_t* __restrict__ out) {
  float mn = 0.f, sc = 1.f;
  bool degen = false;
  if (mm) {
    mn = mm[0];
    if (mm[1] > mm[0])
      sc = 1.0f / (mm[1] - mm[0]);
    else
      degen = true;
  }
  const int64_t stride = (int64_t)gridDim.x * SEG_THREADS;
  for (int64_t i = (int64_t)blockIdx.x * SEG_THREADS + threadIdx.x; i < n; i += stride) {
    float x = __ldg(v + i);
    if (mm) x = degen ? 1.0f : (x - mn) * sc;
    out[i] = (uint8_t)(x > thr);
  }
}

}  // namespace mfish

using namespace mfish;

extern "C" int mfish_histogram_f32(const float* vol_dev, int64_t n, double first, double last, int nbins,
                                   const double* edges_dev, unsigned long long* hist_dev, void* stream) {
  MFISH_REQUIRE(vol_dev && edges_dev && hist_dev, "histogram: null pointer");
  MFISH_REQUIRE(n > 0 && nbins > 0 && nbins <= SEG_MAX_BINS && last > first, "histogram: bad arguments");
  cudaStream_t st = as_stream(stream);
  MFISH_CUDA(cudaMemsetAsync(hist_dev, 0, sizeof(unsigned long long) * nbins, st));
  const int blocks = (int)std::min<int64_t>((n + SEG_THREADS - 1) / SEG_THREADS, (int64_t)num_sms() * 8);
  const size_t smem = (size_t)(SEG_THREADS / 32) * nbins * sizeof(unsigned int);
  ProfScope prof("histogram", st);
  hist_kernel<<<blocks, SEG_THREADS, smem, st>>>(vol_dev, n, first, last, nbins, edges_dev, hist_dev);
  count_launch(1);
  return check_launch("histogram");
}

extern "C" int mfish_split_moments_f32(const float* vol_dev, int64_t n, float threshold, double* out4_dev,
                                       void* stream) {
  MFISH_REQUIRE(vol_dev && out4_dev && n > 0, "split_moments: bad arguments");
  cudaStream_t st = as_stream(stream);
  MFISH_CUDA(cudaMemsetAsync(out4_dev, 0, 4 * sizeof(double), st));
  const int blocks = (int)std::min<int64_t>((n + SEG_THREADS - 1) / SEG_THREADS, (int64_t)num_sms() * 8);
  ProfScope prof("split_moments", st);
  split_moments_kernel<<<blocks, SEG_THREADS, 0, st>>>(vol_dev, n, threshold, out4_dev);
  count_launch(1);
  return check_launch("split_moments");
}

extern "C" int mfish_binarize_f32(const float* vol_dev, int64_t n, float threshold, const float* norm_minmax_dev,
                                  uint8_t* out_dev, void* stream) {
  MFISH_REQUIRE(vol_dev && out_dev && n > 0, "binarize: bad arguments");
  cudaStream_t st = as_stream(stream);
  const int blocks = (int)std::min<int64_t>((n + SEG_THREADS - 1) / SEG_THREADS, (int64_t)num_sms() * 16);
  ProfScope prof("binarize", st);
  binarize_kernel<<<blocks, SEG_THREADS, 0, st>>>(vol_dev, n, threshold, norm_minmax_dev, out_dev);
  count_launch(1);
  return check_launch("binarize");
}
