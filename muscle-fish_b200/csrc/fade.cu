// Per-nucleus "fade" volumes: nocodazole/fish_analysis_noco.py:233-238 runs
//     nuc = np.where(region_nuc_labeled == i, 1, 0);  nuc_fade = filters.gaussian(nuc.astype(float), (3, 3, 1))
// once per nucleus over the WHOLE volume (O(nuclei x V) on the host).  The one-hot mask of label i is zero
// outside the label's bounding box, so its Gaussian (truncate 4: radius 12 / 12 / 4) is exactly zero outside the
// box grown by the radius; inside, boundary rule `nearest` only matters where the grown box meets the volume's own
// faces.  All labels are therefore filtered in ONE batch over their grown boxes ("crops", concatenated in one
// buffer): a bounding-box kernel (atomics on labelled voxels only) and three passes that each run one thread per
// crop voxel -- the first reads the int32 label volume and forms the one-hot value on the fly, the other two read
// the previous pass inside the crop and take zero outside it.  Sum of the crops is ~n_nuclei x 10^5 voxels: launch
// latency, not bandwidth, is what this costs.
#include <vector>

#include "common.cuh"

namespace mfish {

constexpr int FADE_MAX_RADIUS = 40;

struct FadeTaps {
  float w[3][FADE_MAX_RADIUS + 1];  // half taps per device axis (z, x, y)
  int r[3];
};

struct FadeArgs {
  const int32_t* labels;
  const int32_t* crops;    // [n][8]: z0, z1, x0, x1, y0, y1 (half-open), label, unused
  const int64_t* offsets;  // [n + 1] element offsets of the crops in the concatenated buffers
  const float* src;        // previous pass (null for the first)
  float* dst;
  int n;
  int nz, nx, ny;
  int axis;                // MFISH_AXIS_*
};

__global__ void label_bbox_kernel(const int32_t* __restrict__ labels, int64_t n, int nx, int ny, int32_t first,
                                  int32_t count, int32_t* __restrict__ boxes) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int32_t l = __ldg(labels + i) - first;
    if (l < 0 || l >= count) continue;
    const int y = (int)(i % ny);
    const int64_t r = i / ny;
    const int x = (int)(r % nx);
    const int z = (int)(r / nx);
    int32_t* b = boxes + (int64_t)l * 6;
    atomicMin(b + 0, z);
    atomicMax(b + 1, z + 1);
    atomicMin(b + 2, x);
    atomicMax(b + 3, x + 1);
    atomicMin(b + 4, y);
    atomicMax(b + 5, y + 1);
  }
}

__global__ void label_bbox_init_kernel(int32_t* boxes, int32_t count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count * 6) boxes[i] = (i % 2 == 0) ? 0x7FFFFFFF : 0;
}

__global__ void __launch_bounds__(256) fade_pass_kernel(const FadeArgs a, const __grid_constant__ FadeTaps t) {
  const int64_t total = a.offsets[a.n];
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  // crop of this element: the last offset <= e
  int lo = 0, hi = a.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(a.offsets + mid) <= e)
      lo = mid;
    else
      hi = mid - 1;
  }
  const int32_t* c = a.crops + (int64_t)lo * 8;
  const int z0 = c[0], z1 = c[1], x0 = c[2], x1 = c[3], y0 = c[4], y1 = c[5], label = c[6];
  const int cx = x1 - x0, cy = y1 - y0;
  const int64_t off = __ldg(a.offsets + lo);
  const int64_t le = e - off;
  const int ly = (int)(le % cy);
  const int lx = (int)((le / cy) % cx);
  const int lz = (int)(le / ((int64_t)cy * cx));
  const int z = z0 + lz, x = x0 + lx, y = y0 + ly;
  const int ax = a.axis;
  const int r = t.r[ax];
  const float* __restrict__ w = t.w[ax];
  float acc = 0.f;
  if (a.src == nullptr) {
    // first pass: one-hot of the label volume, boundary rule `nearest` on the volume itself
    const int pos = ax == MFISH_AXIS_Z ? z : (ax == MFISH_AXIS_X ? x : y);
    const int len = ax == MFISH_AXIS_Z ? a.nz : (ax == MFISH_AXIS_X ? a.nx : a.ny);
    const int64_t st = ax == MFISH_AXIS_Z ? (int64_t)a.nx * a.ny : (ax == MFISH_AXIS_X ? (int64_t)a.ny : 1);
    const int32_t* __restrict__ line = a.labels + ((int64_t)z * a.nx + x) * a.ny + y - (int64_t)pos * st;
    for (int k = -r; k <= r; ++k) {
      const int p = min(max(pos + k, 0), len - 1);
      if (__ldg(line + (int64_t)p * st) == label) acc += w[k < 0 ? -k : k];
    }
  } else {
    const int pos = ax == MFISH_AXIS_Z ? z : (ax == MFISH_AXIS_X ? x : y);
    const int len = ax == MFISH_AXIS_Z ? a.nz : (ax == MFISH_AXIS_X ? a.nx : a.ny);
    const int c0 = ax == MFISH_AXIS_Z ? z0 : (ax == MFISH_AXIS_X ? x0 : y0);
    const int c1 = ax == MFISH_AXIS_Z ? z1 : (ax == MFISH_AXIS_X ? x1 : y1);
    const int64_t st = ax == MFISH_AXIS_Z ? (int64_t)cx * cy : (ax == MFISH_AXIS_X ? (int64_t)cy : 1);
    const float* __restrict__ line = a.src + off + le - (int64_t)(pos - c0) * st;
    for (int k = -r; k <= r; ++k) {
      const int p = min(max(pos + k, 0), len - 1);
      if (p >= c0 && p < c1) acc = fmaf(w[k < 0 ? -k : k], __ldg(line + (int64_t)(p - c0) * st), acc);
    }
  }
  a.dst[e] = acc;
}

}  // namespace mfish

using namespace mfish;

extern "C" int mfish_label_bbox_i32(const int32_t* labels_dev, int64_t nz, int64_t nx, int64_t ny, int32_t first_label,
                                    int32_t n_labels, int32_t* boxes_dev, void* stream) {
  MFISH_REQUIRE(labels_dev && boxes_dev, "label_bbox: null pointer");
  MFISH_REQUIRE(nz > 0 && nx > 0 && ny > 0 && nz < (1 << 30) && nx < (1 << 30) && ny < (1 << 30),
                "label_bbox: bad dimensions");
  MFISH_REQUIRE(n_labels >= 0, "label_bbox: negative label count");
  if (n_labels == 0) return MFISH_OK;
  cudaStream_t st = as_stream(stream);
  ProfScope prof("label_bbox", st);
  label_bbox_init_kernel<<<(n_labels * 6 + 255) / 256, 256, 0, st>>>(boxes_dev, n_labels);
  const int64_t n = nz * nx * ny;
  const int64_t blocks = std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 16);
  label_bbox_kernel<<<(unsigned)blocks, 256, 0, st>>>(labels_dev, n, (int)nx, (int)ny, first_label, n_labels,
                                                      boxes_dev);
  count_launch(2);
  return check_launch("label_bbox");
}

// scipy _gaussian_kernel1d (order 0), as sepconv.cu's host_taps
static int fade_taps(double sigma, float* w, int& radius) {
  radius = (int)(4.0 * sigma + 0.5);
  if (sigma <= 0.0) radius = 0;
  MFISH_REQUIRE(radius <= FADE_MAX_RADIUS, "onehot_gaussian: sigma too large (radius <= 40)");
  if (radius == 0) {
    w[0] = 1.f;
    return MFISH_OK;
  }
  std::vector<double> phi(radius + 1);
  double sum = 0.0;
  for (int k = 0; k <= radius; ++k) {
    phi[k] = exp(-0.5 / (sigma * sigma) * (double)(k * k));
    sum += k == 0 ? phi[k] : 2.0 * phi[k];
  }
  for (int k = 0; k <= radius; ++k) w[k] = (float)(phi[k] / sum);
  return MFISH_OK;
}

extern "C" int mfish_onehot_gaussian_f32(const int32_t* labels_dev, int64_t nz, int64_t nx, int64_t ny,
                                         const int32_t* crops_dev, const int64_t* offsets_dev, int32_t n_crops,
                                         int64_t total_elems, const double* sigma_zxy_host, float* out_dev,
                                         float* tmp_dev, void* stream) {
  MFISH_REQUIRE(labels_dev && crops_dev && offsets_dev && sigma_zxy_host && out_dev && tmp_dev,
                "onehot_gaussian: null pointer");
  MFISH_REQUIRE(nz > 0 && nx > 0 && ny > 0 && nz < (1 << 30) && nx < (1 << 30) && ny < (1 << 30),
                "onehot_gaussian: bad dimensions");
  MFISH_REQUIRE(n_crops >= 0 && total_elems >= 0, "onehot_gaussian: negative size");
  if (n_crops == 0 || total_elems == 0) return MFISH_OK;
  FadeTaps t;
  memset(&t, 0, sizeof(t));
  for (int ax = 0; ax < 3; ++ax) {
    int rc = fade_taps(sigma_zxy_host[ax], t.w[ax], t.r[ax]);
    if (rc != MFISH_OK) return rc;
  }
  cudaStream_t st = as_stream(stream);
  ProfScope prof("onehot_gaussian", st);
  FadeArgs a;
  a.labels = labels_dev;
  a.crops = crops_dev;
  a.offsets = offsets_dev;
  a.n = n_crops;
  a.nz = (int)nz;
  a.nx = (int)nx;
  a.ny = (int)ny;
  const unsigned blocks = (unsigned)((total_elems + 255) / 256);
  // scipy's order: axis 0 first = x, then y, then z of the reference's (x, y, z) arrays
  const int order[3] = {MFISH_AXIS_X, MFISH_AXIS_Y, MFISH_AXIS_Z};
  const float* src = nullptr;
  float* bufs[3] = {out_dev, tmp_dev, out_dev};
  for (int k = 0; k < 3; ++k) {
    a.axis = order[k];
    a.src = src;
    a.dst = bufs[k];
    fade_pass_kernel<<<blocks, 256, 0, st>>>(a, t);
    src = bufs[k];
  }
  count_launch(3);
  return check_launch("onehot_gaussian");
}
