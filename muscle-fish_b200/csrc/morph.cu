// Compartment masks: n-fold binary erosion / dilation with the cross structuring element, as the reference's
// scripts build the nuclear / perinuclear / sarcoplasmic regions (general_pipeline/fish_analysis.py:169-187,
// nocodazole/fish_analysis_noco.py:193-208, rna_decay/fish_analysis_actd.py:165-183):
//
//     for n in range(n3): mask = skimage.morphology.binary_dilation(mask)            # 3-D cross
//     for n in range(n2): for z: mask[:,:,z] = binary_dilation(mask[:,:,z])          # 2-D cross, per plane
//
// n dilations by the cross are ONE dilation by the L1 ball of radius n, and the two loops compose into
//     out(x,y,z) = OR over |dz| <= n3 of [ in-plane city-block distance to mask(.,.,z+dz)  <=  n2 + n3 - |dz| ]
// so the 50 volume sweeps of the reference become one separable city-block distance transform, capped at
// n2 + n3:  row scan along y (edt.cu) -> min-plus pass with |.| along x (two sweeps) -> windowed pass along z.
// Erosion is the dilation of the complement (skimage erodes with border_value=True: voxels outside the
// volume count as set, i.e. they are never sites of the complement).  Exact: every quantity is an integer.
//
// Bytes per voxel: scan R1+W2, x pass R2+W1+R1+W1, z pass R1(+window through L1)+W1  ~ 10.
#include <algorithm>

#include "common.cuh"

namespace mfish {

constexpr int MORPH_THREADS = 128;

// ---- x pass: d2(x) = min_u min(dy(u), 255) + |x - u|, saturating at 255; four y columns per thread -------------
struct MorphXArgs {
  const uint16_t* dy;  // [nz][nx][ny]
  uint8_t* d2;         // [nz][nx][ny]
  int nx, ny;
};

__device__ __forceinline__ uint32_t pack_cap4(uint16_t a, uint16_t b, uint16_t c, uint16_t d) {
  return (uint32_t)min((int)a, 255) | ((uint32_t)min((int)b, 255) << 8) | ((uint32_t)min((int)c, 255) << 16) |
         ((uint32_t)min((int)d, 255) << 24);
}

__global__ void __launch_bounds__(MORPH_THREADS) morph_x_kernel(const MorphXArgs a) {
  const int y0 = (blockIdx.x * MORPH_THREADS + threadIdx.x) * 4;
  if (y0 >= a.ny) return;
  const int64_t plane = (int64_t)blockIdx.y * a.nx * a.ny;
  const uint16_t* __restrict__ src = a.dy + plane + y0;
  uint8_t* __restrict__ dst = a.d2 + plane + y0;
  const bool vec = (y0 + 4 <= a.ny) && ((a.ny & 3) == 0) && aligned16_dev(a.dy) && aligned16_dev(a.d2);
  auto load4 = [&](int x) -> uint32_t {
    const uint16_t* p = src + (int64_t)x * a.ny;
    if (vec) {
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
      return pack_cap4((uint16_t)(v.x & 0xFFFF), (uint16_t)(v.x >> 16), (uint16_t)(v.y & 0xFFFF), (uint16_t)(v.y >> 16));
    }
    uint16_t e[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) e[j] = (y0 + j < a.ny) ? __ldg(p + j) : (uint16_t)0xFFFF;
    return pack_cap4(e[0], e[1], e[2], e[3]);
  };
  auto store4 = [&](int x, uint32_t v) {
    uint8_t* p = dst + (int64_t)x * a.ny;
    if (vec) {
      *reinterpret_cast<uint32_t*>(p) = v;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (y0 + j < a.ny) p[j] = (uint8_t)(v >> (8 * j));
    }
  };
  auto read4 = [&](int x) -> uint32_t {
    const uint8_t* p = dst + (int64_t)x * a.ny;
    if (vec) return *reinterpret_cast<const uint32_t*>(p);
    uint32_t v = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) v |= (uint32_t)((y0 + j < a.ny) ? p[j] : 255) << (8 * j);
    return v;
  };
  uint32_t f = 0xFFFFFFFFu;
#pragma unroll 4
  for (int x = 0; x < a.nx; ++x) {  // sites at or before x
    f = __vminu4(load4(x), __vaddus4(f, 0x01010101u));
    store4(x, f);
  }
  uint32_t b = 0xFFFFFFFFu;
#pragma unroll 4
  for (int x = a.nx - 1; x >= 0; --x) {  // ... and after x
    b = __vminu4(read4(x), __vaddus4(b, 0x01010101u));
    store4(x, b);
  }
}

// ---- z pass: out = [ min over |dz| <= r3 of d2(z+dz) + |dz|  <=  R ] (optionally complemented) -----------------
struct MorphZArgs {
  const uint8_t* d2;
  uint8_t* out;
  int64_t plane;  // nx * ny
  int nz, r3, R, complement;
};

__global__ void __launch_bounds__(MORPH_THREADS) morph_z_kernel(const MorphZArgs a) {
  const int64_t i0 = ((int64_t)blockIdx.x * MORPH_THREADS + threadIdx.x) * 4;
  if (i0 >= a.plane) return;
  const int z = blockIdx.y;
  const bool vec = (i0 + 4 <= a.plane) && ((a.plane & 3) == 0) && aligned16_dev(a.d2) && aligned16_dev(a.out);
  uint32_t best = 0xFFFFFFFFu;
  const int lo = max(z - a.r3, 0), hi = min(z + a.r3, a.nz - 1);
  for (int u = lo; u <= hi; ++u) {
    const uint8_t* p = a.d2 + (int64_t)u * a.plane + i0;
    uint32_t v;
    if (vec) {
      v = __ldg(reinterpret_cast<const uint32_t*>(p));
    } else {
      v = 0;
      for (int j = 0; j < 4; ++j) v |= (uint32_t)((i0 + j < a.plane) ? __ldg(p + j) : 255) << (8 * j);
    }
    const uint32_t k = (uint32_t)abs(u - z);
    best = __vminu4(best, __vaddus4(v, k * 0x01010101u));
  }
  const uint32_t Rv = (uint32_t)a.R * 0x01010101u;
  uint32_t hit = __vcmpleu4(best, Rv) & 0x01010101u;  // 1 where inside the dilation
  if (a.complement) hit ^= 0x01010101u;
  uint8_t* q = a.out + (int64_t)z * a.plane + i0;
  if (vec) {
    *reinterpret_cast<uint32_t*>(q) = hit;
  } else {
    for (int j = 0; j < 4; ++j)
      if (i0 + j < a.plane) q[j] = (uint8_t)(hit >> (8 * j));
  }
}

// region_peri = (dil > nuc_eroded) & (fiber | nuc);  region_cyt = fiber > dil   (fish_analysis.py:182-186)
__global__ void compartments_combine_kernel(const uint8_t* nuc, const uint8_t* fiber, const uint8_t* ero,
                                            const uint8_t* dil, uint8_t* peri, uint8_t* cyt, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const bool nu = nuc[i] != 0, fi = fiber[i] != 0, er = ero[i] != 0, di = dil[i] != 0;
    peri[i] = (uint8_t)((di && !er) && (fi || nu));
    cyt[i] = (uint8_t)(fi && !di);
  }
}

static inline int64_t align256m(int64_t v) { return (v + 255) & ~int64_t(255); }

// op: 0 dilate, 1 erode.  mask_dev nonzero = set.  out_dev 0/1.
static int morph_run(const uint8_t* mask, int64_t nz, int64_t nx, int64_t ny, int op, int n3, int n2, uint8_t* out,
                     void* ws, cudaStream_t st) {
  const int64_t n = nz * nx * ny;
  uint16_t* dy = reinterpret_cast<uint16_t*>(ws);
  uint8_t* d2 = reinterpret_cast<uint8_t*>(ws) + align256m(n * 2);
  // dilation: sites are the set voxels; erosion: sites are the UNSET voxels, result complemented
  int rc = launch_scan_y(mask, dy, nz * nx, ny, op == 0 ? 1 : 0, st);
  if (rc != MFISH_OK) return rc;
  {
    MorphXArgs a{dy, d2, (int)nx, (int)ny};
    dim3 grid((unsigned)((ny + 4 * MORPH_THREADS - 1) / (4 * MORPH_THREADS)), (unsigned)nz);
    ProfScope prof("morph_x", st);
    morph_x_kernel<<<grid, MORPH_THREADS, 0, st>>>(a);
    count_launch(1);
    rc = check_launch("morph_x");
    if (rc != MFISH_OK) return rc;
  }
  {
    MorphZArgs a{d2, out, nx * ny, (int)nz, n3, n2 + n3, op == 1 ? 1 : 0};
    dim3 grid((unsigned)((nx * ny + 4 * MORPH_THREADS - 1) / (4 * MORPH_THREADS)), (unsigned)nz);
    ProfScope prof("morph_z", st);
    morph_z_kernel<<<grid, MORPH_THREADS, 0, st>>>(a);
    count_launch(1);
    rc = check_launch("morph_z");
  }
  return rc;
}

}  // namespace mfish

using namespace mfish;

extern "C" int64_t mfish_morph_workspace_bytes(int64_t nz, int64_t nx, int64_t ny) {
  if (nz <= 0 || nx <= 0 || ny <= 0) return 0;
  const int64_t n = nz * nx * ny;
  return align256m(n * 2) + align256m(n);
}

extern "C" int mfish_binary_morph_u8(const uint8_t* mask_dev, int64_t nz, int64_t nx, int64_t ny, int op, int n3d,
                                     int n2d, uint8_t* out_dev, void* workspace_dev, void* stream) {
  MFISH_REQUIRE(mask_dev && out_dev && workspace_dev, "morph: null pointer");
  MFISH_REQUIRE(nz > 0 && nx > 0 && ny > 0 && nz < 65536, "morph: bad dimensions");
  MFISH_REQUIRE(op == MFISH_MORPH_DILATE || op == MFISH_MORPH_ERODE, "morph: bad op");
  MFISH_REQUIRE(n3d >= 0 && n2d >= 0 && n3d + n2d <= 254, "morph: iteration counts must satisfy n3d + n2d <= 254");
  return morph_run(mask_dev, nz, nx, ny, op, n3d, n2d, out_dev, workspace_dev, as_stream(stream));
}

extern "C" int mfish_compartments_u8(const uint8_t* nuc_dev, const uint8_t* fiber_dev, int64_t nz, int64_t nx,
                                     int64_t ny, int erode3d, int erode2d, int dilate3d, int dilate2d,
                                     uint8_t* region_nuc_dev, uint8_t* region_peri_dev, uint8_t* region_cyt_dev,
                                     void* workspace_dev, void* stream) {
  MFISH_REQUIRE(nuc_dev && fiber_dev && region_nuc_dev && region_peri_dev && region_cyt_dev && workspace_dev,
                "compartments: null pointer");
  MFISH_REQUIRE(nz > 0 && nx > 0 && ny > 0 && nz < 65536, "compartments: bad dimensions");
  MFISH_REQUIRE(erode3d >= 0 && erode2d >= 0 && erode3d + erode2d <= 254 && dilate3d >= 0 && dilate2d >= 0 &&
                    dilate3d + dilate2d <= 254,
                "compartments: iteration counts out of range");
  cudaStream_t st = as_stream(stream);
  const int64_t n = nz * nx * ny;
  int rc = morph_run(nuc_dev, nz, nx, ny, MFISH_MORPH_ERODE, erode3d, erode2d, region_nuc_dev, workspace_dev, st);
  if (rc != MFISH_OK) return rc;
  // region_cyt_dev doubles as the home of the dilated mask until the combine kernel has read it
  rc = morph_run(nuc_dev, nz, nx, ny, MFISH_MORPH_DILATE, dilate3d, dilate2d, region_cyt_dev, workspace_dev, st);
  if (rc != MFISH_OK) return rc;
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 16);
  ProfScope prof("compartments_combine", st);
  compartments_combine_kernel<<<blocks, 256, 0, st>>>(nuc_dev, fiber_dev, region_nuc_dev, region_cyt_dev,
                                                      region_peri_dev, region_cyt_dev, n);
  count_launch(1);
  return check_launch("compartments_combine");
}
