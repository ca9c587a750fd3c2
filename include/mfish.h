/* libmfish_sm100 -- C ABI of the B200-native muscle-FISH spot-analysis path.
 *
 * The reference (cpkelley94/muscle-FISH) has no FFI layer: its boundary is the
 * Python function surface of modules/muscle_fish.py and modules/scope_utils3.py
 * plus two direct scipy calls.  Each entry point below names the reference
 * call (file:line, relative to the upstream repository root) whose arithmetic
 * it replaces; INTEGRATION.md shows the ctypes binding the Python host side
 * uses.
 *
 * Conventions
 *  - plain C: pointers + sizes only, no torch / C++ types.
 *  - every function returns an int status (MFISH_OK == 0, < 0 on error);
 *    mfish_last_error() gives a thread-local message for the last failure.
 *  - pointers named *_dev are DEVICE pointers owned by the caller; pointers
 *    named *_host are host pointers read before the call returns.
 *  - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it.
 *  - device volumes are laid out [nz][nx][ny], y contiguous ("zxy").  The
 *    reference's logical order is (x, y, z) with z fastest ("xyz",
 *    modules/muscle_fish.py:60-71); mfish_ingest() converts.  Spot indices
 *    returned by the library are REFERENCE raster indices
 *    ((x*ny + y)*nz + z)*ns + s so that ordering matches np.nonzero().
 */
#ifndef MFISH_H_
#define MFISH_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFISH_ABI_VERSION 6

#define MFISH_OK 0
#define MFISH_ERR_INVALID (-1)
#define MFISH_ERR_CUDA (-2)
#define MFISH_ERR_CAPACITY (-3)
#define MFISH_ERR_UNSUPPORTED (-4)

/* element types accepted by mfish_ingest */
#define MFISH_U8 0
#define MFISH_U16 1
#define MFISH_I16 2
#define MFISH_I32 3
#define MFISH_I64 4
#define MFISH_F32 5
#define MFISH_F64 6

#define MFISH_LAYOUT_XYZ 0 /* reference order, z fastest */
#define MFISH_LAYOUT_ZXY 1 /* device order, y fastest  */

#define MFISH_DST_F32 0     /* numeric value as float32              */
#define MFISH_DST_NONZERO 1 /* uint8: 1 where src != 0 else 0 (masks) */
#define MFISH_DST_ISZERO 2  /* uint8: 1 where src == 0 else 0 (np.logical_not(mask),
                               nocodazole/fish_analysis_noco.py:218)   */

#define MFISH_BOUNDARY_REFLECT 0 /* scipy 'reflect'  (gaussian_laplace default) */
#define MFISH_BOUNDARY_NEAREST 1 /* scipy 'nearest'  (skimage.filters.gaussian) */

#define MFISH_AXIS_Z 0
#define MFISH_AXIS_X 1
#define MFISH_AXIS_Y 2

/* one separable pass: G = order-0 taps, D = order-2 taps */
#define MFISH_CONV_G 0    /* out0 = G*in0                               */
#define MFISH_CONV_GD 1   /* out0 = G*in0, out1 = D*in0                 */
#define MFISH_CONV_MID 2  /* out0 = G*in0, out1 = D*in0 + G*in1         */
#define MFISH_CONV_LAST 3 /* out0 = out_scale * (D*in0 + G*in1)         */

#define MFISH_MAX_RADIUS 512

int mfish_abi_version(void);
const char* mfish_last_error(void);
/* number of kernels this library has launched in the calling process */
int64_t mfish_launch_count(void);
/* per-kernel CUDA-event timing on the launching stream (bench / roofline reporting):
 * enable, run, then collect "name calls total_ms" lines (returns bytes needed). */
void mfish_profile_enable(int on);
int64_t mfish_profile_collect(char* buf, int64_t cap);

/* ---- ingest / normalisation ------------------------------------------- *
 * replaces the float64 copies made by su.normalize_image
 * (modules/scope_utils3.py:381-386) and open_image's transposed view
 * (modules/muscle_fish.py:60-71).  Converts `src_dev` (dtype/layout given)
 * into the device layout; when minmax_dev != NULL also reduces min and max of
 * the converted values into minmax_dev[0..1] (float).                      */
int mfish_ingest(const void* src_dev, int src_dtype, int src_layout, int64_t nx, int64_t ny,
                 int64_t nz, void* dst_dev, int dst_kind, float* minmax_dev, void* stream);

/* Masks over PCIe as bits.  The masks the reference passes (`mask=` of find_spots*, modules/muscle_fish.py:281,
 * 631; the nucleus mask of nocodazole/fish_analysis_noco.py:218) carry one bit per voxel; the host-buffer path
 * packs them on the host and expands them on the device instead of moving whole bytes.
 *  mfish_mask_bits_bytes   size of the bit stream of n voxels (whole 32-bit words + one pad word)
 *  mfish_host_pack_mask    HOST code (no CUDA call; thread-safe on disjoint ranges): bit i (LSB first inside a
 *                          byte) of bits_host = (src_host[i] != 0) != negate, for n elements of src_dtype.
 *                          Worker threads split a mask at multiples of 8 elements.
 *  mfish_ingest_mask_bits  the packed stream of a C-ordered array in src_layout -> uint8 mask [nz][nx][ny]
 *                          (dst_kind MFISH_DST_NONZERO / MFISH_DST_ISZERO), i.e. what mfish_ingest gives for
 *                          the unpacked array.  bits_dev: 4-byte aligned, mfish_mask_bits_bytes long.        */
int64_t mfish_mask_bits_bytes(int64_t n);
int mfish_host_pack_mask(const void* src_host, int src_dtype, int64_t n, uint8_t* bits_host, int negate);
int mfish_ingest_mask_bits(const void* bits_dev, int src_layout, int64_t nx, int64_t ny, int64_t nz,
                           uint8_t* dst_dev, int dst_kind, void* stream);

/* global min / max (np.min, np.max inside np.interp, scope_utils3.py:386) */
int mfish_minmax_f32(const float* vol_dev, int64_t n, float* minmax_dev, void* stream);

/* su.normalize_image as a stand-alone call (modules/scope_utils3.py:381-386):
 * out = np.interp(src, (min, max), (vmin, vmax)) in float64, same element
 * order as src (n elements of src_dtype).  ws16_dev: 16 bytes of scratch.     */
int mfish_normalize_f64(const void* src_dev, int src_dtype, int64_t n, double vmin, double vmax,
                        double* out_dev, void* ws16_dev, void* stream);

/* device [z][x][y] float32 -> float64 in dst_layout (MFISH_LAYOUT_XYZ: the
 * reference order its callers index as arr[x, y, z]): the dense outputs handed
 * back to the reference's scripts (the distance map of
 * nocodazole/fish_analysis_noco.py:218-221, fix_bleaching's corrected image,
 * filtered volumes).                                                          */
int mfish_export_f64(const float* vol_zxy_dev, int64_t nx, int64_t ny, int64_t nz, int dst_layout,
                     double* out_dev, void* stream);

/* ---- separable filtering ---------------------------------------------- *
 * replaces scipy.ndimage.correlate1d as driven by gaussian_filter /
 * gaussian_laplace (call sites modules/muscle_fish.py:333,357,661 through
 * skimage.feature.blob_log / skimage.filters.gaussian).  `taps_*_host` hold
 * radius+1 half taps (offset 0..radius; the kernels are symmetric).  When
 * norm_minmax_dev != NULL, in0 is mapped (v-min)/(max-min) on load
 * (ops G / GD only).                                                       */
int mfish_sepconv_f32(const float* in0_dev, const float* in1_dev, float* out0_dev, float* out1_dev,
                      int64_t nz, int64_t nx, int64_t ny, int axis, const double* taps_g_host,
                      const double* taps_d_host, int radius, int boundary, int op, float out_scale,
                      const float* norm_minmax_dev, void* stream);

/* skimage.filters.gaussian(img, sigma) == gaussian_filter(mode, truncate=4)
 * (modules/muscle_fish.py:166,239,357; nocodazole/fish_analysis_noco.py:235).
 * sigma_zxy_host[3] in device axis order; tmp_dev is one scratch volume.   */
int mfish_gaussian3d_f32(const float* in_dev, float* out_dev, float* tmp_dev, int64_t nz, int64_t nx,
                         int64_t ny, const double* sigma_zxy_host, int boundary,
                         const float* norm_minmax_dev, void* stream);

/* -sigma^2 * scipy.ndimage.gaussian_laplace(normalize(img), sigma), the
 * per-scale volume of skimage.feature.blob_log (modules/muscle_fish.py:333,
 * 661; modules/scope_utils3.py:403).  Three fused passes (y: G,D; z: MID;
 * x: LAST).  tmp1..3 are scratch volumes; out_dev is also used as scratch.  */
int mfish_log3d_f32(const float* in_dev, float* out_dev, float* tmp1_dev, float* tmp2_dev,
                    float* tmp3_dev, int64_t nz, int64_t nx, int64_t ny, double sigma,
                    const float* norm_minmax_dev, void* stream);

/* The same for the planes [out_z0, out_z1) of a z-extended slab only (z-slab mosaics: the planes
 * within `radius` of an interior cut see the boundary rule instead of real neighbours and are not
 * wanted).  Planes outside the range of out_dev are left unspecified.        */
int mfish_log3d_range_f32(const float* in_dev, float* out_dev, float* tmp1_dev, float* tmp2_dev,
                          float* tmp3_dev, int64_t nz, int64_t nx, int64_t ny, double sigma,
                          const float* norm_minmax_dev, int64_t out_z0, int64_t out_z1, void* stream);

/* LoG response + thresholded 26-neighbour non-max suppression + compaction in one call (one scale;
 * skimage.feature.blob_log with num_sigma=1 as modules/muscle_fish.py:333,661 call it).  The x pass of
 * the LoG appends, next to the volume it writes, every voxel above the threshold that is a maximum
 * along x (warp ballot + one atomic per warp); the 3x3x3 test then runs on those candidates only.
 * out_dev receives the LoG volume as mfish_log3d_f32 would write it.  cand_ws_dev: (cand_cap + 1) uint32.
 * *count_dev == 0xFFFFFFFF: candidate overflow, call mfish_nms_compact_f32 on out_dev.             */
int mfish_log3d_peaks_f32(const float* in_dev, float* out_dev, float* tmp1_dev, float* tmp2_dev,
                          float* tmp3_dev, int64_t nz, int64_t nx, int64_t ny, double sigma,
                          const float* norm_minmax_dev, float threshold, void* cand_ws_dev, int64_t cand_cap,
                          int64_t* out_idx_dev, float* out_val_dev, uint32_t* count_dev, int64_t capacity,
                          void* stream);
/* The same on the planes [out_z0, out_z1) of a z-extended slab (z-slab mosaics), taken as a cube of their own:
 * LoG as mfish_log3d_range_f32, peak indices relative to that cube, zero padding at its two faces.       */
int mfish_log3d_peaks_range_f32(const float* in_dev, float* out_dev, float* tmp1_dev, float* tmp2_dev,
                                float* tmp3_dev, int64_t nz, int64_t nx, int64_t ny, double sigma,
                                const float* norm_minmax_dev, int64_t out_z0, int64_t out_z1, float threshold,
                                void* cand_ws_dev, int64_t cand_cap, int64_t* out_idx_dev, float* out_val_dev,
                                uint32_t* count_dev, int64_t capacity, void* stream);

/* ---- spot detection --------------------------------------------------- *
 * skimage.feature.peak_local_max(cube, threshold_abs=t, footprint=ones(3^4))
 * as blob_log calls it: v > t and v >= every in-volume neighbour in
 * (scale, z, x, y); ties kept.  cube_dev is [ns][nz][nx][ny].  Writes up to
 * `capacity` (index, value) pairs in arbitrary order and the total number of
 * peaks found into *count_dev (which may exceed capacity: re-run larger).   */
int mfish_nms_compact_f32(const float* cube_dev, int64_t ns, int64_t nz, int64_t nx, int64_t ny,
                          float threshold, int64_t* out_idx_dev, float* out_val_dev,
                          uint32_t* count_dev, int64_t capacity, void* stream);

/* order peaks like peak_local_max and apply blob_log's _prune_blobs for equal sigma.
 * order = MFISH_ORDER_RESPONSE: descending value, ties by ascending raster index (skimage >= 0.18);
 * order = MFISH_ORDER_REVERSE_RASTER: descending raster index (skimage 0.14-0.17, the reference's era:
 * _get_high_intensity_peaks returns np.column_stack(np.nonzero(mask))[::-1]).  With integer
 * coordinates, the peak of rank r is dropped iff some peak of HIGHER rank index lies within squared
 * voxel distance cutoff_sq (the lower index of every overlapping pair is zeroed).
 * n is the number of valid entries.  alive_dev[r] in {0,1}.                */
#define MFISH_ORDER_RESPONSE 0
#define MFISH_ORDER_REVERSE_RASTER 1
int mfish_rank_prune(const int64_t* idx_dev, const float* val_dev, int64_t n, int64_t ns,
                     int64_t nz, int64_t nx, int64_t ny, int64_t cutoff_sq, int order,
                     int64_t* idx_out_dev, float* val_out_dev, uint8_t* alive_out_dev, void* stream);

/* candidate pairs for the unequal-sigma prune (multi-scale blob_log,
 * modules/scope_utils3.py:403): all i<j (rank order) whose squared voxel
 * distance is <= cutoff_sq_host[si*ns+sj].  Pairs are written unordered.   */
int mfish_overlap_pairs(const int64_t* idx_ranked_dev, int64_t n, int64_t ns, int64_t nz, int64_t nx,
                        int64_t ny, const int64_t* cutoff_sq_host, int32_t* pairs_dev,
                        uint32_t* count_dev, int64_t capacity, void* stream);

/* ---- statistics used by find_spots_snrfilter / fix_bleaching ----------- */
#define MFISH_QUANTILE_WS_BYTES (64 * 1024)
/* np.percentile(img[mask], q) order statistics (modules/muscle_fish.py:355-356):
 * out_dev = {value[k], value[k+1]} of the masked RAW values, k = floor of
 * numpy's 'linear' virtual index; info_dev = {gamma, n_masked}.            */
int mfish_masked_quantile_f32(const float* vol_dev, const uint8_t* mask_dev, int64_t n, double q,
                              void* ws_dev, float* out_dev, double* info_dev, void* stream);

/* The same two order statistics from ONE full pass: a 1/64 sample brackets the wanted ranks, one pass
 * counts what lies below the bracket and collects what lies inside, the exact ranks are selected among
 * the collected values.  info_dev = {gamma, n_masked, status}: status 1 means the bracket missed (checked
 * on the device) and the caller must fall back to mfish_masked_quantile_f32; status 0 is exact.          */
int mfish_masked_quantile_fast_f32(const float* vol_dev, const uint8_t* mask_dev, int64_t n, double q,
                                   void* ws_dev, float* out_dev, double* info_dev, void* stream);

/* mfish_masked_quantile_fast_f32 whose one full pass also reduces {min, max} over ALL voxels into
 * minmax_dev (float[2]): su.normalize_image's bounds (modules/scope_utils3.py:381-386) and the fibre
 * baseline percentile (modules/muscle_fish.py:355-356) from a single read of the volume.  minmax_dev is
 * valid whatever the bracket status says.                                                              */
int mfish_masked_quantile_minmax_f32(const float* vol_dev, const uint8_t* mask_dev, int64_t n, double q,
                                     void* ws_dev, float* out_dev, double* info_dev, float* minmax_dev,
                                     void* stream);

/* the same select in its three steps, for a z-slab decomposed volume: after every
 * mfish_qsel_hist the caller all-reduces (sum) the first 2*2048 uint32 of ws_dev
 * across the slabs, then calls mfish_qsel_scan; passes 0, 1, 2 in order.      */
int mfish_qsel_begin(void* ws_dev, double q, void* stream);
int mfish_qsel_hist(const float* vol_dev, const uint8_t* mask_dev, int64_t n, int pass, void* ws_dev,
                    void* stream);
int mfish_qsel_scan(int pass, void* ws_dev, float* out_dev, double* info_dev, void* stream);
/* mfish_qsel_scan on a histogram the caller holds in 64 bits (hist64_dev: 2*2048 uint64, the 32-bit bins of
 * every slab widened BEFORE the all-reduce): mosaics of 2^32 voxels and more (config 4 is 6.7e9).        */
int mfish_qsel_scan_u64(int pass, void* ws_dev, const unsigned long long* hist64_dev, float* out_dev,
                        double* info_dev, void* stream);

/* The bracketed select (mfish_masked_quantile_fast_f32: ONE full pass) in steps, for a volume split across
 * ranks -- the fibre baseline np.percentile(img[mask], 25) of modules/muscle_fish.py:355-356 on a mosaic:
 *   mfish_qselb_begin(ws, q)
 *   pass 0, 1, 2:  mfish_qselb_sample_hist(...)  -> all-reduce the first 2*2048 uint32 of ws (1/64 sample:
 *                  no overflow below 2^38 voxels) -> mfish_qsel_scan(pass, ws, pair_dev, scratch_dev)
 *   mfish_qselb_collect(...)   pair_dev = the bracketing values; one pass over the slab counts the masked
 *                  voxels, those below the bracket, and appends those inside it to cand_dev (cap floats,
 *                  cap % 64 == 0); counts_dev = int64 {n_masked, n_below, n_cand, overflow}
 *                  -> all-reduce counts_dev (sum); minmax_dev (optional float[2]): {min, max} over ALL
 *                  voxels of the slab from the same pass (su.normalize_image's bounds: all-reduce min / max)
 *   mfish_qselb_ranks(ws, q, counts_dev, out, info)   global ranks relative to the candidates;
 *                  info_dev = {gamma, n_masked, status}; status 1: bracket missed / a buffer overflowed,
 *                  run the three-pass select instead
 *   pass 0, 1, 2:  mfish_qselb_cand_hist(cand, cap, pass, ws) -> all-reduce the histogram
 *                  -> mfish_qsel_scan(pass, ws, out_dev, info_dev);  out_dev = the two order statistics  */
int mfish_qselb_begin(void* ws_dev, double q, void* stream);
int mfish_qselb_sample_hist(const float* vol_dev, const uint8_t* mask_dev, int64_t n, int pass, void* ws_dev,
                            void* stream);
int mfish_qselb_collect(const float* vol_dev, const uint8_t* mask_dev, int64_t n, void* ws_dev,
                        const float* pair_dev, float* cand_dev, int64_t cap, long long* counts_dev,
                        float* minmax_dev, void* stream);
int mfish_qselb_ranks(void* ws_dev, double q, const long long* counts_dev, float* out_dev, double* info_dev,
                      void* stream);
int mfish_qselb_cand_hist(const float* cand_dev, int64_t cap, int pass, void* ws_dev, void* stream);

/* img_blur[pos] of skimage.filters.gaussian(normalize(img), sigma)
 * (modules/muscle_fish.py:357-365,385-388), evaluated only at the given
 * points in float64.  pts_dev is npts x 3 int32 (z, x, y).                 */
int mfish_blur_at_points(const float* vol_dev, int64_t nz, int64_t nx, int64_t ny,
                         const int32_t* pts_dev, int64_t npts, const double* taps_z_host, int rz,
                         const double* taps_x_host, int rx, const double* taps_y_host, int ry,
                         int boundary, const float* norm_minmax_dev, double* out_dev, void* stream);

/* value lookups at integer voxel positions: mask[spot] tests
 * (modules/muscle_fish.py:335-343,663-670) and the interpn gather of the
 * distance map (nocodazole/fish_analysis_noco.py:297-298).                 */
int mfish_gather_f32(const float* vol_dev, int64_t nz, int64_t nx, int64_t ny, const int32_t* pts_dev,
                     int64_t npts, float* out_dev, void* stream);
int mfish_gather_u8(const uint8_t* vol_dev, int64_t nz, int64_t nx, int64_t ny, const int32_t* pts_dev,
                    int64_t npts, uint8_t* out_dev, void* stream);

/* per-z-plane masked sum and count (fix_bleaching, modules/muscle_fish.py:
 * 456-461).  sums_dev / counts_dev are double[nz]; mask_dev may be NULL.   */
int mfish_plane_sums_f32(const float* vol_dev, const uint8_t* mask_dev, int64_t nz, int64_t nx,
                         int64_t ny, double* sums_dev, double* counts_dev, void* stream);
/* out = in * factor[z] (modules/muscle_fish.py:506-507)                    */
int mfish_scale_planes_f32(const float* in_dev, float* out_dev, int64_t nz, int64_t nx, int64_t ny,
                           const float* factor_dev, void* stream);

/* ---- Euclidean distance transform ------------------------------------- *
 * scipy.ndimage.distance_transform_edt(mask, sampling)
 * (nocodazole/fish_analysis_noco.py:218; simulation/analyze_simulation.py:101):
 * distance from every voxel to the nearest ZERO voxel of mask_dev.  Exact,
 * separable: one scan pass (y) + two lower-envelope passes (x, z).
 * spacing_zxy_host[3] in device axis order.  out_dist_dev (float32) and/or
 * out_sq_dev (int32 squared voxel distance; isotropic spacing only) may be
 * NULL.  A mask with no zero voxel yields +inf (scipy returns garbage).    */
int64_t mfish_edt_workspace_bytes(int64_t nz, int64_t nx, int64_t ny);
int mfish_edt3d_u8(const uint8_t* mask_dev, int64_t nz, int64_t nx, int64_t ny,
                   const double* spacing_zxy_host, float* out_dist_dev, int32_t* out_sq_dev,
                   void* workspace_dev, void* stream);

/* The same transform in two halves, for a z-slab decomposed mosaic: the y and
 * x passes are local to a z-slab; the z pass runs after the slabs have been
 * re-partitioned into x-slabs (all-to-all).  f2 is the squared in-plane
 * distance, 4 bytes per voxel, encoded as *f2_kind_host says.
 * mfish_edt_inplane_u8 needs mfish_edt_workspace_bytes() of scratch,
 * mfish_edt_zpass needs nz*nx*ny*4 bytes.                                    */
/* ... and in three, for an x-slab decomposition (halo cost of the filters 1 % instead of 70 %): the row
 * scan is local to an x-slab, the slabs are then re-partitioned into y-slabs (2 bytes per voxel on the
 * wire), where both envelope passes are local: mfish_edt_xpass_u16 (ny_full = the mosaic's ny, which
 * bounds the row distances) + mfish_edt_zpass.                                                       */
int mfish_edt_rowscan_u8(const uint8_t* mask_dev, int64_t nrows, int64_t ny, uint16_t* dy_out_dev,
                         void* stream);
int mfish_edt_xpass_u16(const uint16_t* dy_dev, int64_t nz, int64_t nx, int64_t ny, int64_t ny_full,
                        const double* spacing_zxy_host, void* f2_out_dev, int* f2_kind_host,
                        void* link_ws_dev, void* stream);
#define MFISH_EDT_F2_I32 0 /* exact integer voxel^2 (x spacing == y spacing) */
#define MFISH_EDT_F2_F32 1 /* float32 physical units^2                       */
int mfish_edt_inplane_u8(const uint8_t* mask_dev, int64_t nz, int64_t nx, int64_t ny,
                         const double* spacing_zxy_host, void* f2_out_dev, int* f2_kind_host,
                         void* workspace_dev, void* stream);
/* f2_bound: upper bound on the finite values of f2 (integer kind), or 0 to derive it from nx, ny --
 * an x-slab of a wider mosaic holds in-plane distances measured over the GLOBAL plane.          */
int mfish_edt_zpass(const void* f2_dev, int f2_kind, int64_t nz, int64_t nx, int64_t ny,
                    const double* spacing_zxy_host, int64_t f2_bound, float* out_dist_dev,
                    int32_t* out_sq_dev, void* link_ws_dev, void* stream);

/* ---- compartment masks ---------------------------------------------------- *
 * The erosion / dilation loops that build the nuclear, perinuclear and sarcoplasmic
 * regions (general_pipeline/fish_analysis.py:169-187, nocodazole/fish_analysis_noco.py:193-208,
 * rna_decay/fish_analysis_actd.py:165-183):
 *     n3d x skimage.morphology.binary_dilation(mask)            (3-D cross)
 *     n2d x binary_dilation(mask[:, :, z]) for every plane z    (2-D cross)
 * computed as one capped city-block distance transform (n dilations by the cross = one
 * dilation by the L1 ball of radius n).  MFISH_MORPH_ERODE follows skimage's
 * binary_erosion (border_value=True: the volume's faces do not erode).  mask_dev: nonzero =
 * set; out_dev: 0/1.  n3d + n2d <= 254.  Workspace: mfish_morph_workspace_bytes().       */
#define MFISH_MORPH_DILATE 0
#define MFISH_MORPH_ERODE 1
int64_t mfish_morph_workspace_bytes(int64_t nz, int64_t nx, int64_t ny);
int mfish_binary_morph_u8(const uint8_t* mask_dev, int64_t nz, int64_t nx, int64_t ny, int op, int n3d,
                          int n2d, uint8_t* out_dev, void* workspace_dev, void* stream);
/* region_nuc  = erode(nuc; erode3d, erode2d)
 * dilated     = dilate(nuc; dilate3d, dilate2d)
 * region_peri = (dilated > region_nuc) & (fiber | nuc)
 * region_cyt  = fiber > dilated                    (general_pipeline/fish_analysis.py:169-186) */
int mfish_compartments_u8(const uint8_t* nuc_dev, const uint8_t* fiber_dev, int64_t nz, int64_t nx,
                          int64_t ny, int erode3d, int erode2d, int dilate3d, int dilate2d,
                          uint8_t* region_nuc_dev, uint8_t* region_peri_dev, uint8_t* region_cyt_dev,
                          void* workspace_dev, void* stream);

/* ---- segmentation thresholds ---------------------------------------------- *
 * The statistics behind threshold_fiber / threshold_nuclei (modules/muscle_fish.py:165-198,
 * 238-272) on the blurred float32 volume.
 * mfish_histogram_f32: np.histogram(vol, bins=nbins, range=(first, last)) as skimage's
 *   threshold_otsu takes it; edges_dev = np.linspace(first, last, nbins + 1) (float64);
 *   hist_dev: nbins uint64 counts.  nbins <= 1024.
 * mfish_split_moments_f32: out4 = {count, sum} of the voxels <= threshold, {count, sum} of
 *   those > threshold (float64) -- one iteration of skimage's threshold_li.
 * mfish_binarize_f32: out = (vol > threshold) as uint8; with norm_minmax_dev the compare is
 *   on (vol - min) / (max - min) (threshold_nuclei binarises the normalised, unblurred
 *   image, modules/muscle_fish.py:181).                                                   */
int mfish_histogram_f32(const float* vol_dev, int64_t n, double first, double last, int nbins,
                        const double* edges_dev, unsigned long long* hist_dev, void* stream);
int mfish_split_moments_f32(const float* vol_dev, int64_t n, float threshold, double* out4_dev,
                            void* stream);
int mfish_binarize_f32(const float* vol_dev, int64_t n, float threshold, const float* norm_minmax_dev,
                       uint8_t* out_dev, void* stream);

/* ---- per-nucleus faded masks (nuclear-mesh step) --------------------------- *
 * nocodazole/fish_analysis_noco.py:233-235 computes, once per nucleus over the whole volume,
 *     nuc_fade = filters.gaussian(np.where(region_nuc_labeled == i, 1, 0).astype(float), (3, 3, 1)).
 * mfish_label_bbox_i32: tight bounding boxes of the labels first_label .. first_label+n_labels-1 of an
 *   int32 label volume [z][x][y]; boxes_dev[l] = {z0, z1, x0, x1, y0, y1} (half-open; z1 <= z0: absent).
 * mfish_onehot_gaussian_f32: the Gaussian (boundary rule `nearest`, truncate 4) of the one-hot mask of
 *   each listed label inside its crop.  crops_dev[c] = {z0, z1, x0, x1, y0, y1, label, 0}; a crop must
 *   contain the label's bounding box grown by the filter radius (clipped to the volume) -- outside of it
 *   the result is exactly 0.  offsets_dev[c] .. offsets_dev[c+1] (int64, n_crops + 1 entries) is the crop's
 *   range in out_dev / tmp_dev (total_elems floats each), layout [z][x][y] per crop.                     */
int mfish_label_bbox_i32(const int32_t* labels_dev, int64_t nz, int64_t nx, int64_t ny, int32_t first_label,
                         int32_t n_labels, int32_t* boxes_dev, void* stream);
int mfish_onehot_gaussian_f32(const int32_t* labels_dev, int64_t nz, int64_t nx, int64_t ny,
                              const int32_t* crops_dev, const int64_t* offsets_dev, int32_t n_crops,
                              int64_t total_elems, const double* sigma_zxy_host, float* out_dev,
                              float* tmp_dev, void* stream);

/* ---- block moves between slabs (NVLink peer memory) ------------------------ *
 * nrows rows of row_bytes from src (rows src_pitch apart) to dst (rows dst_pitch apart); all four and both
 * base addresses multiples of 16.  dst may be a buffer of a PEER GPU mapped into this process (symmetric
 * memory): the kernel then packs and transfers in one go -- the halo exchange of the filters and the
 * all-to-all in front of the EDT's envelope passes of a slab-decomposed mosaic (BASELINE configs[3]).        */
int mfish_copy2d(void* dst_dev, int64_t dst_pitch, const void* src_dev, int64_t src_pitch, int64_t row_bytes,
                 int64_t nrows, void* stream);
/* the same for n2 x n1 rows with two pitches each (rows pitch1 apart, groups of n1 rows pitch2 apart): a
 * [nz][rows][ny] block whose rows are a sub-range of both volumes.                                           */
int mfish_copy3d(void* dst_dev, int64_t dst_pitch1, int64_t dst_pitch2, const void* src_dev, int64_t src_pitch1,
                 int64_t src_pitch2, int64_t row_bytes, int64_t n1, int64_t n2, void* stream);

/* ---- connected components (segmentation clean-up and labelling) ------------ *
 * skimage.morphology.label / remove_small_objects / remove_small_holes as modules/muscle_fish.py:182-184,
 * 266-272 and nocodazole/fish_analysis_noco.py:227 call them.  mask_dev: C-ordered (na, nb, nc) uint8 volume
 * (nonzero = foreground; na = 1 for 2-D), laid out in the order the CALLER's array is -- a voxel's linear index
 * is then the raster index scipy / skimage scan in.  parent_dev (int32, same size) receives, for every
 * foreground voxel, the smallest linear index of its component (-1 for background): the component's label is the
 * rank of that index among the roots.  connectivity_full: 26- (8-) neighbourhood, else the 6- (4-) one.      */
int mfish_ccl_roots_u8(const uint8_t* mask_dev, int64_t na, int64_t nb, int64_t nc, int connectivity_full,
                       int32_t* parent_dev, void* stream);
/* np.bincount(labels.ravel(), minlength=nvalues) of an int32 label volume (values outside [0, nvalues) ignored):
 * the component sizes remove_small_objects / remove_small_holes and threshold_fiber's "largest object" need.   */
int mfish_label_sizes_i32(const int32_t* labels_dev, int64_t n, int32_t nvalues, unsigned long long* sizes_dev,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MFISH_H_ */
