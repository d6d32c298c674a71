/* pcs.h -- C ABI of libpcs, the B200 (sm_100a) segmentation hot path.
 *
 * The reference (ssilverman16/particle_col_image_segmentation) has no FFI: its
 * boundary is Python functions over numpy arrays that call scipy.ndimage /
 * scikit-image.  Each entry point below replaces one of those
 * library calls (or one reference loop) and cites the reference file:line it
 * stands in for.  INTEGRATION.md shows the ctypes stub that binds them.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless it says "host";
 *  - images are batches of independent 2-D slices (B, H, W), C order;
 *  - a "bit image" is uint32[B][H][WW], WW = (W + 31) / 32, bit j of word k is
 *    pixel x = 32 k + j, bits at x >= W are zero;
 *  - the library never allocates or frees memory: scratch comes from the caller,
 *    sized by the matching *_bytes query;
 *  - all work is enqueued asynchronously on `stream` (a cudaStream_t);
 *  - return value: 0, or a negative PCS_ERR_* with pcs_last_error_string();
 *  - no global mutable state besides the per-thread error string: one host
 *    thread per GPU is safe.
 */
#ifndef PCS_H
#define PCS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCS_VERSION 200
#define PCS_TABLE_COLS 10 /* area, sum_y, sum_x, min_y, min_x, max_y, max_x, first_index, sum_intensity, overlap */

int pcs_version(void);
const char* pcs_last_error_string(void);
int pcs_device_sm_count(int device);
/* launch accounting: kernels launched since load; optional per-kernel CUDA-event timing on the
 * launching stream (enable, run, collect -> distinct kernel names (64-char slots), total ms, launches) */
uint64_t pcs_kernel_launches(void);
int pcs_profile_enable(int on);
int pcs_profile_collect(char* names, double* total_ms, int32_t* launches, int n_max); /* host pointers */

/* ---- K2: thresholds, class masks, LUT relabels --------------------------------
 * cmp: 0 '>', 1 '>=', 2 '<', 3 '<=', 4 '==', 5 '!='.  thr_dev (optional) holds one
 * threshold per slice and overrides thr.  bits and/or mask (uint8 0/1) may be NULL.
 *   boundary_map < threshold            refine_boundaries.py:44-45
 *   ds_arr == label                     tiff_analysis.py:256-257, :812, :818, :984, :987
 *   image > threshold_otsu(image)       north_star (no reference call site) */
int pcs_compare_u8(const uint8_t* img, int thr, const int32_t* thr_dev, int cmp, uint32_t* bits, uint8_t* mask, int B, int H, int W, void* stream);
int pcs_compare_u16(const uint16_t* img, int thr, const int32_t* thr_dev, int cmp, uint32_t* bits, uint8_t* mask, int B, int H, int W, void* stream);
int pcs_compare_i32(const int32_t* img, int thr, const int32_t* thr_dev, int cmp, uint32_t* bits, uint8_t* mask, int B, int H, int W, void* stream);
int pcs_compare_f32(const float* img, float thr, const float* thr_dev, int cmp, uint32_t* bits, uint8_t* mask, int B, int H, int W, void* stream);
int pcs_compare_f64(const double* img, double thr, const double* thr_dev, int cmp, uint32_t* bits, uint8_t* mask, int B, int H, int W, void* stream);
/* np.isin(img, values) through a 256-entry membership table (tiff_analysis.py:816-819) */
int pcs_member_u8(const uint8_t* img, const uint8_t* member256, uint32_t* bits, uint8_t* mask, int B, int H, int W, void* stream);
int pcs_unpack_bits(const uint32_t* bits, uint8_t* mask, int B, int H, int W, void* stream);
/* op: 0 and, 1 or, 2 a & ~b, 3 xor, 4 ~a   (tiff_analysis.py:1000-1007) */
int pcs_bits_logic(const uint32_t* a, const uint32_t* b, uint32_t* out, int op, int B, int H, int W, void* stream);
/* np.sum(mask) per slice (tiff_analysis.py:1015) */
int pcs_bits_count(const uint32_t* bits, uint64_t* counts, int B, int H, int W, void* stream);
/* ordered arr[arr == a] = b chains collapsed to one table (tiff_analysis.py:177-181, :224-231) */
int pcs_lut_u8(uint8_t* img, const uint8_t* lut256, int64_t n, void* stream);
/* img[bits] = value (tiff_analysis.py:240, :286, :1013) */
int pcs_assign_where_u8(uint8_t* img, const uint32_t* bits, int value, int B, int H, int W, void* stream);
/* out[i] = img[slice[i]][idx[i]]; dtype 0 u8, 2 i32, 5 i64 (tiff_analysis.py:845-852, :1041-1044) */
int pcs_gather(const void* img, int dtype, const int64_t* slice, const int64_t* idx, int64_t* out, int64_t n, int64_t slice_elems, void* stream);

/* dst[i] = value for n_words uint32 (multiple of 4, dst 16-byte aligned): grid-stride 128-bit stores; doubles as
 * the store-only bandwidth probe of bench.py */
/* *out (device int64) = largest label of an int32 / int64 label image (0 for an empty one): regionprops on a label image
 * whose label count is not known (tiff_analysis.py:263, :746 pass label images straight from label()) */
int pcs_max_label(const void* labels, int label_bytes, int64_t n, int64_t* out, void* stream);
/* out (W, H) = in (H, W) transposed, elem_bytes 1 or 4: MATLAB numbers components in column-major order (.m:104, :173) */
int pcs_transpose(const void* in, void* out, int elem_bytes, int H, int W, void* stream);
int pcs_fill_u32(void* dst, uint32_t value, size_t n_words, void* stream);
/* zero `bytes` (multiple of 32, 32-byte aligned) with a small persistent grid (ctas_per_sm CTAs of 128 threads per SM):
 * for a side stream, beside kernels that leave DRAM idle */
int pcs_zero_background(void* dst, size_t bytes, int ctas_per_sm, void* stream);

/* ---- K1: histogram + Otsu ------------------------------------------------------
 * skimage.filters.threshold_otsu on uint16 slices (north_star).
 * hist is uint32[B][65536]; thr receives the threshold, minmax (optional) 2 per slice. */
size_t pcs_histogram_bytes(int B);
int pcs_histogram_u16(const uint16_t* img, uint32_t* hist, int B, int H, int W, void* stream);
int pcs_otsu_u16(const uint32_t* hist, int32_t* thr, int32_t* minmax, int B, int64_t npix, void* stream);

/* ---- K3: median ----------------------------------------------------------------
 * scipy.ndimage.median_filter(a, size) with mode='reflect' (tiff_analysis.py:122, :643);
 * size in {3, 5, 7}.  pcs_majority_bits is the same filter on a binary image. */
int pcs_median_u8(const uint8_t* img, uint8_t* out, int size, int B, int H, int W, void* stream);
int pcs_majority_bits(const uint32_t* in, uint32_t* out, int size, int B, int H, int W, void* stream);
/* same, also writing the result as a uint8 0/1 image (fused) */
int pcs_majority_bits_mask(const uint32_t* in, uint32_t* out, uint8_t* mask, int size, int B, int H, int W, void* stream);

/* ---- K4: connected-component labelling -----------------------------------------
 * skimage.measure.label / scipy.ndimage.label (tiff_analysis.py:260, :743, :829;
 * refine_boundaries.py:64): labels 1..N in raster order of the first pixel, per slice.
 * connectivity 4 or 8; labels int32 (label_bytes 4) or int64 (8).
 * counts[B] = N per slice; offsets[B+1] (optional) = exclusive scan of counts;
 * first_out (optional, int64[cap]) = first raster index of every label, row
 * offsets[b] + label - 1 (column 7 of the region table). */
size_t pcs_ccl_workspace_bytes(int B, int H, int W, int with_aux);
int pcs_label_bits(const uint32_t* bits, int B, int H, int W, int connectivity, int invert, void* labels, int label_bytes,
                   int32_t* counts, int32_t* offsets, int64_t* first_out, int64_t cap, void* ws, size_t ws_bytes, void* stream);
/* multi-valued images: six connectivity planes (F, S, U, UL, UR, J), uint32[6][B][H][WW].
 * dtype 0 u8, 1 u16, 2 i32, 3 f32, 4 f64; all_fg labels every pixel (plateaus);
 * higher (optional bit image) flags pixels with a strictly greater neighbour. */
size_t pcs_conn_planes_bytes(int B, int H, int W);
int pcs_conn_planes(const void* img, int dtype, uint32_t* planes, uint32_t* higher, int all_fg, int connectivity, int B, int H, int W, void* stream);
int pcs_label_conn(const uint32_t* planes, int B, int H, int W, int connectivity, void* labels, int label_bytes,
                   int32_t* counts, int32_t* offsets, int64_t* first_out, int64_t cap, void* ws, size_t ws_bytes, void* stream);

/* ---- K6 and friends on the same forest ------------------------------------------
 * scipy.ndimage.binary_fill_holes (tiff_analysis.py:880) */
int pcs_fill_holes_bits(const uint32_t* bits, uint32_t* out, int B, int H, int W, void* ws, size_t ws_bytes, void* stream);
/* the same result when the 8-connected components of `bits` with area >= min_size are described by a
 * region table (bounding boxes): only background inside a box is labelled (holes lie inside their
 * component's box), which is what the pipeline uses */
size_t pcs_fill_holes_table_workspace_bytes(int B, int H, int W);
int pcs_fill_holes_table_bits(const uint32_t* bits, const int64_t* table, int64_t cap, const int32_t* offsets, int64_t min_size,
                              uint32_t* out, uint8_t* out_mask /* optional uint8 copy */, int B, int H, int W, void* ws, size_t ws_bytes,
                              void* stream);
/* remove_small_objects(min_size, 8-connected) followed by binary_fill_holes when the 8-connected components
 * of `bits` are already labelled (`labels` int32, any numbering) and their areas sit in column 0 of `table`
 * (tiff_analysis.py:769-773 then :880).  Only row gaps bounded by two runs of one label can hold hole pixels,
 * so just those are labelled.  table / offsets may be null when min_size <= 1.  Workspace:
 * pcs_fill_holes_table_workspace_bytes. */
int pcs_refine_labeled_bits(const uint32_t* bits, const int32_t* labels, const int64_t* table, int64_t cap, const int32_t* offsets,
                            int64_t min_size, uint32_t* out, uint8_t* out_mask /* optional uint8 copy */, int B, int H, int W, void* ws,
                            size_t ws_bytes, void* stream);
/* skimage.morphology.remove_small_objects (area filter of tiff_analysis.py:769-773); ws needs with_aux = 1 */
int pcs_remove_small_bits(const uint32_t* bits, uint32_t* out, int B, int H, int W, int connectivity, int min_size, void* ws, size_t ws_bytes, void* stream);
/* components of `bits` that contain a seed pixel (merged_image, tiff_analysis.py:843-878) */
int pcs_select_components_bits(const uint32_t* bits, const uint32_t* seeds, uint32_t* out, int B, int H, int W, int connectivity, void* ws, size_t ws_bytes, void* stream);
/* K9 skimage.morphology.local_maxima (refine_boundaries.py:63): planes/higher from pcs_conn_planes(all_fg = 1) */
int pcs_local_maxima_conn(const uint32_t* planes, const uint32_t* higher, uint32_t* out, int32_t* counts, int B, int H, int W, int connectivity, void* ws, size_t ws_bytes, void* stream);

/* ---- K5: binary morphology -------------------------------------------------------
 * out(y,x) = OR over runs r, dx in [lo_r, hi_r] of in'(y - dy_r, x - dx); runs is int32[n_runs][3]
 * = (dy, lo, hi) with hi - lo < 32; in' = in ^ invert_in; outside the image reads `border`;
 * out is complemented when invert_out.  Dilation: flags 0,0,0.  Erosion by S with
 * border_value v: runs of the reflected S, flags 1, !v, 1.
 *   skimage.morphology.binary_dilation(mask, disk(2))  tiff_analysis.py:827-828 */
int pcs_dilate_bits(const uint32_t* in, uint32_t* out, const int32_t* runs, int n_runs, int invert_in, int border, int invert_out, int B, int H, int W, void* stream);

/* ---- K7: exact Euclidean distance transform ---------------------------------------
 * scipy.ndimage.distance_transform_edt (tiff_analysis.py:996; refine_boundaries.py:60) of
 * bits ^ invert.  Any of: dist (float64), sq (int32 squared distance), thr_bits
 * (sq <= thr_sq: binary_dilation(mask, disk(r)) == EDT(~mask)^2 <= r^2, tiff_analysis.py:990). */
size_t pcs_edt_workspace_bytes(int B, int H, int W);
int pcs_edt_bits(const uint32_t* bits, int invert, int B, int H, int W, double* dist, int32_t* sq, uint32_t* thr_bits, int thr_sq, void* ws, size_t ws_bytes, void* stream);

/* ---- K8 / K10 / K11: per-label reductions ------------------------------------------
 * table is int64[PCS_TABLE_COLS][cap]; row offsets[b] + label - 1.
 *   regionprops area / centroid / bbox / coords[0]   tiff_analysis.py:263-275, :746-773
 *   overlap pixels per label                          tiff_analysis.py:268-279
 * intensity_dtype: -1 none, 0 u8, 1 u16; fg_bits (optional) lets empty words be skipped. */
int pcs_table_init(int64_t* table, int64_t cap, void* stream);
/* same, touching only rows [0, offsets[B]) -- for a table whose row count is already on the device */
int pcs_table_init_rows(int64_t* table, int64_t cap, const int32_t* offsets, int B, void* stream);
int pcs_region_table(const void* labels, int label_bytes, const void* intensity, int intensity_dtype, const uint32_t* fg_bits,
                     const uint32_t* ov_bits, const int32_t* offsets, int64_t* table, int64_t cap, int B, int H, int W, void* stream);
/* the float64 table the callers consume, row-major double[cap][13], rows [0, offsets[B]) filled:
 * z (= z0 + slice), label, area, centroid_y, centroid_x, min_row, min_col, max_row + 1, max_col + 1,
 * first_row, first_col, intensity_sum, intensity_mean (regionprops .area/.centroid/.bbox/.coords[0],
 * tiff_analysis.py:754-773; exact integer sums, one IEEE division each) */
int pcs_table_finalize(const int64_t* table, int64_t cap, const int32_t* offsets, int B, int W, double z0, double* out, void* stream);
/* the same into a caller-chosen row buffer of out_cap rows (rows beyond out_cap are dropped, never written) with the
 * true row count stored as a double at *count_out (optional): lets a pipeline finalise straight into the message
 * buffer of the table gather (dist.TableGather staging: header with the count + a speculative number of rows) */
int pcs_table_finalize_ex(const int64_t* table, int64_t cap, const int32_t* offsets, int B, int W, double z0, double* out, int64_t out_cap,
                          double* count_out, void* stream);
/* pixels whose label has keep[b][label] != 0 (merged_image |= labels == v, tiff_analysis.py:878) */
int pcs_select_labels(const void* labels, int label_bytes, const uint8_t* keep, int64_t lut_stride, uint32_t* out, int B, int H, int W, void* stream);
/* pixels whose component area >= min_size, from the table (tiff_analysis.py:769-773) */
int pcs_select_by_area(const int32_t* labels, const uint32_t* fg_bits, const int64_t* table, int64_t cap, const int32_t* offsets,
                       int64_t min_size, uint32_t* out, int B, int H, int W, void* stream);
/* sum(sum(plane .* roimask)) for K float64 planes (.m:122-132, :186-196); out float64[n_rois][K] */
int pcs_roi_sums_f64(const int32_t* labels, const double* planes, int K, int64_t npix, int n_rois, double* out, void* stream);
/* nearest neighbour of every a_i among the b_j (float64 pairs): distance and (optional) index, first minimum
 * on ties; exclude_self skips j == i (nearest OTHER cell of one strain).  No b point -> inf, index -1.
 * Cell-cell distances within and across strains: refine_boundaries.py:8-12 (goal 3), model .m:260-263. */
int pcs_nearest_f64(const double* a, int64_t na, const double* b, int64_t nb, int exclude_self, double* out_dist, int64_t* out_index, void* stream);
/* out[i] = min_j |a_i - b_j| for (x, y) float64 pairs (pdist2 + min, .m:260-263, :301-304) */
int pcs_min_dist_f64(const double* a, int64_t na, const double* b, int64_t nb, double* out, void* stream);

/* ---- marker-controlled watershed ----------------------------------------------------------------
 * skimage.segmentation.watershed(image, markers, mask=mask) with connectivity 1, compactness 0, no line
 * (refine_boundaries.py:73): every pixel joins the 4-neighbour of smallest bottleneck cost, computed by
 * tiled relaxation (pcs_watershed.cu).  Bit-identical to the sequential flood on tie-free images.
 * image float64, markers int32 (> 0 = seed; seeds outside the mask are dropped), mask_bits optional bit
 * image, labels int32 out.  BLOCKING (reads the convergence flags once per batch of 8 sweeps); max_sweeps <= 0 picks a bound
 * from the image size; sweeps_out (host int, optional) receives the number of sweeps. */
size_t pcs_watershed_workspace_bytes(int B, int H, int W);
int pcs_watershed_f64(const double* image, const int32_t* markers, const uint32_t* mask_bits, int32_t* labels, int B, int H, int W,
                      int max_sweeps, int* sweeps_out, void* ws, size_t ws_bytes, void* stream);

/* ---- NanoSIMS ratio images (.m:17-69) --------------------------------------------------------
 * imgaussfilt(A, sigma): taps 2*ceil(2*sigma)+1, replicate border, columns then rows, float64; tmp is
 * scratch of the same size as the images */
int pcs_gauss_f64(const double* in, double* out, double* tmp, double sigma, int B, int H, int W, void* stream);
/* ratio = num ./ ((d0 + d1) + d2) (null denominators are skipped; all null: ratio = num);
 * *maxv = max over the non-NaN ratios (.m:45 `max(N15gauss(:)./(N15gauss(:)+N14gauss(:)))`) */
int pcs_ratio_f64(const double* num, const double* d0, const double* d1, const double* d2, double* ratio, double* maxv, int64_t n, void* stream);
/* one dimension of MATLAB's imresize (.m:125, :189): out[i, x] = sum_p wts[i][p] * in[idx[i][p], x], taps in order, no FMA;
 * idx < 0 marks padding.  Strides (in elements) pick which dimension is resized; with the transposed tap table the same
 * call applies the adjoint (per-ROI sums under resized masks: one resize per ion plane, not one per ROI). */
int pcs_resize_taps_f64(const double* in, double* out, const int32_t* idx, const double* wts, int P, int64_t n_main, int64_t n_other,
                        int64_t in_stride_main, int64_t in_stride_other, int64_t out_stride_main, int64_t out_stride_other, void* stream);
/* uint8(x .* (255 / *maxv)) with MATLAB's conversion: round half away from zero, saturate, NaN -> 0 (.m:31-37) */
int pcs_scale_u8_f64(const double* x, const double* maxv, uint8_t* out, int64_t n, void* stream);

/* ---- the whole segment pipeline for one chunk of slices -----------------------------------
 * threshold (Otsu) -> size x size binary median -> label (8-connected) -> per-label table ->
 * small objects (< min_size) out, holes filled -> exact EDT.  Stands in for the chain
 * split_zstack.py:52 (slice loop) -> tiff_analysis.py:643, :743, :746, :769-773, :880 ->
 * refine_boundaries.py:60; CPU statement: oracle/pipeline.py.  Outputs: mask / refined uint8,
 * labels int32, edt float64 (all (B, H, W)), thr / counts int32[B], offsets int32[B+1],
 * table int64[PCS_TABLE_COLS][cap]. */
size_t pcs_segment_workspace_bytes(int B, int H, int W);
int pcs_segment_chunk(const uint16_t* img, int B, int H, int W, int denoise_size, int min_size, uint8_t* mask, int32_t* labels,
                      uint8_t* refined, double* edt, int32_t* thr, int32_t* counts, int32_t* offsets, int64_t* table,
                      int64_t cap, double* ftable /* optional double[cap][13], see pcs_table_finalize */, int z0, void* ws,
                      size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PCS_H */
