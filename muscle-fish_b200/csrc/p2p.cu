// Block moves between slabs over NVLink peer memory.  The re-partitions of a
// slab-decomposed mosaic (halo planes / rows of the raw input, the all-to-all in front of the EDT's envelope
// passes) are strided blocks: n2 x n1 rows of `row_bytes`, with one pitch per level at the source and at the
// destination.  With the destination mapped from a peer GPU (torch symmetric memory hands out the
// pointer) one kernel packs AND transfers: 16-byte loads from local HBM, 16-byte stores straight into the peer's
// buffer through NVSwitch -- no staging copy, no send/recv pairing, and the block lands where the consumer
// reads it.
#include "common.cuh"

namespace mfish {

// element e of the block = (i2, i1, w): i2 < n2 planes (pitch2), i1 < n1 rows (pitch1), w < row16 words
__global__ void __launch_bounds__(256) copy3d_v16_kernel(uint4* __restrict__ dst, int64_t dp1, int64_t dp2,
                                                         const uint4* __restrict__ src, int64_t sp1, int64_t sp2,
                                                         int row16, int n1, int64_t total16) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  auto locate = [&](int64_t e, int64_t& so, int64_t& doff) {
    const int64_t r = e / row16;
    const int w = (int)(e - r * row16);
    const int64_t i2 = r / n1;
    const int64_t i1 = r - i2 * n1;
    so = i2 * sp2 + i1 * sp1 + w;
    doff = i2 * dp2 + i1 * dp1 + w;
  };
  // four independent 16-byte words in flight per thread
  for (; i + 3 * stride < total16; i += 4 * stride) {
    uint4 v[4];
    int64_t d[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int64_t so;
      locate(i + k * stride, so, d[k]);
      v[k] = __ldg(src + so);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) dst[d[k]] = v[k];
  }
  for (; i < total16; i += stride) {
    int64_t so, doff;
    locate(i, so, doff);
    dst[doff] = __ldg(src + so);
  }
}

}  // namespace mfish

using namespace mfish;

static int copy3d_impl(void* dst_dev, int64_t dst_pitch1, int64_t dst_pitch2, const void* src_dev, int64_t src_pitch1,
                       int64_t src_pitch2, int64_t row_bytes, int64_t n1, int64_t n2, void* stream) {
  MFISH_REQUIRE(dst_dev && src_dev, "copy3d: null pointer");
  MFISH_REQUIRE(row_bytes >= 0 && n1 >= 0 && n2 >= 0 && dst_pitch1 >= row_bytes && src_pitch1 >= row_bytes,
                "copy3d: bad geometry");
  if (row_bytes == 0 || n1 == 0 || n2 == 0) return MFISH_OK;
  MFISH_REQUIRE(((row_bytes | dst_pitch1 | dst_pitch2 | src_pitch1 | src_pitch2) & 15) == 0 && aligned16(dst_dev) &&
                    aligned16(src_dev),
                "copy3d: rows, pitches and base addresses must be multiples of 16 bytes");
  MFISH_REQUIRE(row_bytes / 16 < (1ll << 31) && n1 < (1ll << 31), "copy3d: block too large");
  cudaStream_t st = as_stream(stream);
  const int64_t total16 = n2 * n1 * (row_bytes / 16);
  const int64_t blocks = std::min<int64_t>((total16 + 255) / 256, (int64_t)num_sms() * 16);
  ProfScope prof("copy3d", st);
  copy3d_v16_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<uint4*>(dst_dev), dst_pitch1 / 16,
                                                      dst_pitch2 / 16, reinterpret_cast<const uint4*>(src_dev),
                                                      src_pitch1 / 16, src_pitch2 / 16, (int)(row_bytes / 16), (int)n1,
                                                      total16);
  count_launch(1);
  return check_launch("copy3d");
}

extern "C" int mfish_copy3d(void* dst_dev, int64_t dst_pitch1, int64_t dst_pitch2, const void* src_dev,
                            int64_t src_pitch1, int64_t src_pitch2, int64_t row_bytes, int64_t n1, int64_t n2,
                            void* stream) {
  return copy3d_impl(dst_dev, dst_pitch1, dst_pitch2, src_dev, src_pitch1, src_pitch2, row_bytes, n1, n2, stream);
}

extern "C" int mfish_copy2d(void* dst_dev, int64_t dst_pitch, const void* src_dev, int64_t src_pitch,
                            int64_t row_bytes, int64_t nrows, void* stream) {
  return copy3d_impl(dst_dev, dst_pitch, dst_pitch * nrows, src_dev, src_pitch, src_pitch * nrows, row_bytes, nrows,
                     1, stream);
}
