// The whole per-slice segment pipeline as ONE C-ABI call per chunk of slices:
//   Otsu threshold -> 5x5 binary median -> label -> per-label table
//   -> small objects out, holes filled -> exact EDT
// (oracle/pipeline.py states the same composition on the CPU; each stage cites the
// reference line it stands in for in its own file).  Issuing all ~30 launches from C keeps
// the stream saturated: no interpreter or allocator work sits between two kernels.
#include "pcs_ccl.cuh"

#include "pcs.h"

namespace {
struct SegWs {
  uint32_t* hist;
  uint32_t *raw, *bits, *keep, *refined;
  void* ccl;
  size_t ccl_bytes;
  void* edt;
  size_t edt_bytes;
  int* rsum;    // per-run intensity sums, laid out like the labeller's parent planes
  int* wlist;   // per slice: the non-empty words (slice-local word indices)
  int* wcount;  // per slice: how many
};

size_t seg_carve(void* ws, int B, int H, int W, SegWs* out) {
  const size_t bits_bytes = pcs_align256((size_t)B * H * pcs_words(W) * 4);
  char* p = (char*)ws;
  size_t n = 0;
  auto take = [&](size_t bytes) {
    char* q = p ? p + n : nullptr;
    n += pcs_align256(bytes);
    return (void*)q;
  };
  SegWs w;
  w.hist = (uint32_t*)take(pcs_histogram_bytes(B));
  w.raw = (uint32_t*)take(bits_bytes);
  w.bits = (uint32_t*)take(bits_bytes);
  w.keep = (uint32_t*)take(bits_bytes);
  w.refined = (uint32_t*)take(bits_bytes);
  w.ccl_bytes = pcs_fill_holes_table_workspace_bytes(B, H, W);  // covers the labelling pass too
  w.ccl = take(w.ccl_bytes);
  w.edt_bytes = pcs_edt_workspace_bytes(B, H, W);
  w.edt = take(w.edt_bytes);
  w.rsum = (int*)take((size_t)B * H * pcs_words(W) * 16 * 4);
  w.wlist = (int*)take((size_t)B * H * pcs_words(W) * 4);
  w.wcount = (int*)take((size_t)(B + 1) * 4);
  if (out) *out = w;
  return n;
}
}  // namespace

extern "C" {

size_t pcs_segment_workspace_bytes(int B, int H, int W) { return seg_carve(nullptr, B, H, W, nullptr); }

int pcs_segment_chunk(const uint16_t* img, int B, int H, int W, int denoise_size, int min_size, uint8_t* mask, int32_t* labels,
                      uint8_t* refined, double* edt, int32_t* thr, int32_t* counts, int32_t* offsets, int64_t* table,
                      int64_t cap, double* ftable, int z0, void* ws, size_t ws_bytes, void* stream) {
  PCS_REQUIRE(B >= 1 && H >= 1 && W >= 1, "empty batch or image");
  PCS_REQUIRE(img && mask && labels && refined && edt && thr && counts && offsets && table, "null argument");
  if (ws == nullptr || ws_bytes < pcs_segment_workspace_bytes(B, H, W)) {
    pcs_set_error("segment workspace too small (see pcs_segment_workspace_bytes)");
    return PCS_ERR_WORKSPACE;
  }
  SegWs w;
  seg_carve(ws, B, H, W, &w);
  int rc;
#define STEP(call) \
  if ((rc = (call)) != PCS_OK) return rc
  STEP(pcs_histogram_u16(img, w.hist, B, H, W, stream));
  STEP(pcs_otsu_u16(w.hist, thr, nullptr, B, (int64_t)H * W, stream));
  const uint32_t* bits = w.raw;
  // fused path (default parameters, images at least as large as the median window): the image is read once more,
  // by the kernel that thresholds, denoises, writes the mask and runs the labeller's tile pass
  const bool fused = (denoise_size == 5 || denoise_size <= 1) && H >= 5 && W >= 5 && H <= 16384 && W <= 16384 && B <= 65535;
  if (fused) {
    cudaStream_t st = (cudaStream_t)stream;
    PcsCclWs cw;
    STEP(pcs_ccl_ws_carve(w.ccl, w.ccl_bytes, B, H, W, 0, &cw));
    bits = w.bits;
    STEP(pcs_seg_label_stage(img, thr, denoise_size == 5, w.bits, mask, labels, counts, table, cap, cw, w.rsum, w.wlist, w.wcount, B, H, W, st));
    cudaMemcpyAsync(offsets, cw.offsets, (size_t)(B + 1) * 4, cudaMemcpyDeviceToDevice, st);
  } else {
    STEP(pcs_compare_u16(img, 0, thr, 0, w.raw, nullptr, B, H, W, stream));
    if (denoise_size > 1) {
      STEP(pcs_majority_bits_mask(w.raw, w.bits, mask, denoise_size, B, H, W, stream));  // bits + uint8 mask in one pass
      bits = w.bits;
    } else {
      STEP(pcs_unpack_bits(bits, mask, B, H, W, stream));
    }
    STEP(pcs_label_bits(bits, B, H, W, 8, 0, labels, 4, counts, offsets, nullptr, 0, w.ccl, w.ccl_bytes, stream));
    STEP(pcs_table_init_rows(table, cap, offsets, B, stream));
    STEP(pcs_region_table(labels, 4, img, 1, bits, nullptr, offsets, table, cap, B, H, W, stream));
  }
  if (ftable) STEP(pcs_table_finalize(table, cap, offsets, B, W, (double)z0, ftable, stream));
  // small objects out and holes filled in one call: areas come from the table just built, and only
  // row gaps between two runs of one label can hold hole pixels
  if (fused) {
    // labels come from the parent planes, hole candidates are resolved over their list; the run-sum planes and the
    // word list of the labelling stage are free by now and serve as the hole forest and the candidate list
    PcsCclWs cw;
    STEP(pcs_ccl_ws_carve(w.ccl, w.ccl_bytes, B, H, W, 0, &cw));
    STEP(pcs_seg_refine_stage(bits, cw.parent, table, cap, offsets, min_size, w.refined, refined, w.raw, w.rsum, w.wlist, w.wcount, B, H, W, (cudaStream_t)stream));
  } else {
    STEP(pcs_refine_labeled_bits(bits, labels, table, cap, offsets, min_size > 1 ? min_size : 1, w.refined, refined, B, H, W, w.ccl, w.ccl_bytes, stream));
  }
  STEP(pcs_edt_bits(w.refined, 0, B, H, W, edt, nullptr, nullptr, 0, w.edt, w.edt_bytes, stream));
#undef STEP
  return PCS_OK;
}

}  // extern "C"
