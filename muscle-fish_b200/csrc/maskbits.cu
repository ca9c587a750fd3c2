// Masks over PCIe as bits.
//
// The reference hands 0/1 masks as whole arrays (np.where(mask, 1, 0) int64, or uint8 / bool:
// general_pipeline/fish_analysis.py:196-199, nocodazole/fish_analysis_noco.py:218); a mask is one BIT of
// information per voxel.  PCIe is the slow link of the host-buffer path (two uint8 masks are a third of the bytes
// of a float32 stack), so the host side packs "voxel != 0" into bits (mfish_host_pack_mask: plain host code, run
// by the caller's worker threads while the image is crossing PCIe) and the device side expands the bits straight
// into the uint8 device layout [z][x][y] (mfish_ingest_mask_bits: the unpack is fused with the (x,y,z) ->
// [z][x][y] transposition mfish_ingest does for whole-byte masks).  Bit i of the packed stream (LSB first inside a
// byte, numpy's packbits(bitorder="little")) is element i of the C-ordered host array.
#include <algorithm>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

#include "common.cuh"

namespace mfish {

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
template <typename S>
static void pack_generic(const S* src, int64_t n, uint8_t* dst, uint8_t flip) {
  const int64_t nb = n >> 3;
  for (int64_t b = 0; b < nb; ++b) {
    const S* s = src + (b << 3);
    unsigned v = 0;
    for (int k = 0; k < 8; ++k) v |= (unsigned)(s[k] != S(0)) << k;
    dst[b] = (uint8_t)(v ^ flip);
  }
  const int rem = (int)(n & 7);
  if (rem) {
    unsigned v = 0;
    for (int k = 0; k < rem; ++k) v |= (unsigned)(src[(nb << 3) + k] != S(0)) << k;
    dst[nb] = (uint8_t)((v ^ flip) & ((1u << rem) - 1u));
  }
}

static void pack_u8(const uint8_t* src, int64_t n, uint8_t* dst, uint8_t flip) {
  int64_t i = 0;
#if defined(__x86_64__)
  // 16 voxels per step: compare with zero, one bit per byte lane (SSE2 is part of the x86-64 baseline)
  const __m128i zero = _mm_setzero_si128();
  const unsigned flip16 = flip ? 0xFFFFu : 0u;
  for (; i + 64 <= n; i += 64) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32));
    const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
    // movemask of (x == 0) is the complement of the wanted bits
    const unsigned ma = ~(unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(a, zero)) & 0xFFFFu;
    const unsigned mb = ~(unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(b, zero)) & 0xFFFFu;
    const unsigned mc = ~(unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(c, zero)) & 0xFFFFu;
    const unsigned md = ~(unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(d, zero)) & 0xFFFFu;
    const uint64_t w = (uint64_t)(ma ^ flip16) | ((uint64_t)(mb ^ flip16) << 16) | ((uint64_t)(mc ^ flip16) << 32) |
                       ((uint64_t)(md ^ flip16) << 48);
    memcpy(dst + (i >> 3), &w, 8);
  }
#endif
  pack_generic<uint8_t>(src + i, n - i, dst + (i >> 3), flip);
}

// ---------------------------------------------------------------------------------------------------------
// device side: bits -> uint8 [z][x][y]
// ---------------------------------------------------------------------------------------------------------
// 32 bits of the stream starting at bit `b` (the buffer is readable up to the word holding bit b + 31: the caller
// pads the allocation)
__device__ __forceinline__ uint32_t bits32_at(const uint32_t* __restrict__ words, int64_t b) {
  const int64_t w = b >> 5;
  const uint32_t lo = __ldg(words + w), hi = __ldg(words + w + 1);
  return __funnelshift_r(lo, hi, (uint32_t)(b & 31));
}

// source order (x,y,z), z fastest: element (x,y,z) is bit (x*ny + y)*nz + z.  A thread owns VEC consecutive y of
// one x and walks z: every store of a warp is one contiguous run of 32*VEC bytes of the plane.
template <int VEC>
__global__ void __launch_bounds__(256)
unpack_xyz_kernel(const uint32_t* __restrict__ words, uint8_t* __restrict__ out, int nx, int ny, int nz,
                  uint32_t flip) {
  const int64_t nq = (int64_t)(ny / VEC);
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq * nx) return;
  const int x = (int)(t / nq);
  const int y = (int)(t - (int64_t)x * nq) * VEC;
  const int64_t plane = (int64_t)nx * ny;
  uint8_t* o = out + (int64_t)x * ny + y;
  const int64_t b0 = ((int64_t)x * ny + y) * nz;
  for (int z0 = 0; z0 < nz; z0 += 32) {
    uint32_t w[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) w[k] = bits32_at(words, b0 + (int64_t)k * nz + z0) ^ flip;
    const int zn = min(32, nz - z0);
#pragma unroll 4
    for (int zz = 0; zz < zn; ++zz) {
      if (VEC == 4) {
        const uint32_t v = ((w[0] >> zz) & 1u) | (((w[1] >> zz) & 1u) << 8) | (((w[2] >> zz) & 1u) << 16) |
                           (((w[3] >> zz) & 1u) << 24);
        *reinterpret_cast<uint32_t*>(o + (int64_t)(z0 + zz) * plane) = v;
      } else {
        o[(int64_t)(z0 + zz) * plane] = (uint8_t)((w[0] >> zz) & 1u);
      }
    }
  }
}

// source already in device order: bit i -> byte i.  A thread expands 32 bits into 32 bytes (two 16-byte stores).
__global__ void __launch_bounds__(256)
unpack_flat_kernel(const uint32_t* __restrict__ words, uint8_t* __restrict__ out, int64_t n, uint32_t flip) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i0 = t * 32;
  if (i0 >= n) return;
  const uint32_t w = __ldg(words + t) ^ flip;
  if (i0 + 32 <= n && aligned16_dev(out)) {
    uint32_t q[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t nib = (w >> (4 * k)) & 0xFu;
      // spread 4 bits over 4 bytes
      q[k] = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
    }
    uint4* o = reinterpret_cast<uint4*>(out + i0);
    o[0] = make_uint4(q[0], q[1], q[2], q[3]);
    o[1] = make_uint4(q[4], q[5], q[6], q[7]);
  } else {
    for (int k = 0; k < 32 && i0 + k < n; ++k) out[i0 + k] = (uint8_t)((w >> k) & 1u);
  }
}

}  // namespace mfish

using namespace mfish;

extern "C" int64_t mfish_mask_bits_bytes(int64_t n) {
  if (n <= 0) return 0;
  // whole 32-bit words, plus one word so that any 32-bit window starting inside the stream can be read
  return ((n + 31) / 32 + 1) * 4;
}

extern "C" int mfish_host_pack_mask(const void* src_host, int src_dtype, int64_t n, uint8_t* bits_host, int negate) {
  MFISH_REQUIRE(src_host && bits_host, "pack mask: null pointer");
  MFISH_REQUIRE(n >= 0, "pack mask: negative length");
  const uint8_t flip = negate ? 0xFFu : 0u;
  switch (src_dtype) {
    case MFISH_U8: pack_u8(reinterpret_cast<const uint8_t*>(src_host), n, bits_host, flip); break;
    case MFISH_U16: pack_generic(reinterpret_cast<const uint16_t*>(src_host), n, bits_host, flip); break;
    case MFISH_I16: pack_generic(reinterpret_cast<const int16_t*>(src_host), n, bits_host, flip); break;
    case MFISH_I32: pack_generic(reinterpret_cast<const int32_t*>(src_host), n, bits_host, flip); break;
    case MFISH_I64: pack_generic(reinterpret_cast<const int64_t*>(src_host), n, bits_host, flip); break;
    case MFISH_F32: pack_generic(reinterpret_cast<const float*>(src_host), n, bits_host, flip); break;
    case MFISH_F64: pack_generic(reinterpret_cast<const double*>(src_host), n, bits_host, flip); break;
    default: set_error("pack mask: unsupported dtype %d", src_dtype); return MFISH_ERR_UNSUPPORTED;
  }
  return MFISH_OK;
}

extern "C" int mfish_ingest_mask_bits(const void* bits_dev, int src_layout, int64_t nx, int64_t ny, int64_t nz,
                                      uint8_t* dst_dev, int dst_kind, void* stream) {
  MFISH_REQUIRE(bits_dev && dst_dev, "ingest bits: null pointer");
  MFISH_REQUIRE(nx > 0 && ny > 0 && nz > 0, "ingest bits: empty volume");
  MFISH_REQUIRE(nx < (1ll << 31) && ny < (1ll << 31) && nz < (1ll << 31), "ingest bits: dimension too large");
  MFISH_REQUIRE(src_layout == MFISH_LAYOUT_XYZ || src_layout == MFISH_LAYOUT_ZXY, "ingest bits: bad layout");
  MFISH_REQUIRE(dst_kind == MFISH_DST_NONZERO || dst_kind == MFISH_DST_ISZERO, "ingest bits: bad dst kind");
  MFISH_REQUIRE((reinterpret_cast<uintptr_t>(bits_dev) & 3u) == 0, "ingest bits: the bit stream must be 4-byte aligned");
  cudaStream_t st = as_stream(stream);
  const uint32_t* words = reinterpret_cast<const uint32_t*>(bits_dev);
  const uint32_t flip = dst_kind == MFISH_DST_ISZERO ? 0xFFFFFFFFu : 0u;
  const int64_t n = nx * ny * nz;
  ProfScope prof("ingest_mask_bits", st);
  if (src_layout == MFISH_LAYOUT_ZXY) {
    const int64_t threads = (n + 31) / 32;
    unpack_flat_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(words, dst_dev, n, flip);
  } else if ((ny & 3) == 0 && (reinterpret_cast<uintptr_t>(dst_dev) & 3u) == 0) {
    const int64_t threads = nx * (ny / 4);
    MFISH_REQUIRE((threads + 255) / 256 < (1ll << 31), "ingest bits: grid too large");
    unpack_xyz_kernel<4><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(words, dst_dev, (int)nx, (int)ny, (int)nz, flip);
  } else {
    const int64_t threads = nx * ny;
    MFISH_REQUIRE((threads + 255) / 256 < (1ll << 31), "ingest bits: grid too large");
    unpack_xyz_kernel<1><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(words, dst_dev, (int)nx, (int)ny, (int)nz, flip);
  }
  count_launch(1);
  return check_launch("ingest_mask_bits");
}
