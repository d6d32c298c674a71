// Connected-component machinery shared by label / fill-holes / small-object
// removal / local-maxima / component selection.
//
// Replaces skimage.measure.label and scipy.ndimage.label
// (tiff_analysis.py:260, :743, :829; refine_boundaries.py:64).
//
// Design (B200-first, not a port of scipy's line-by-line run labeller nor of
// scikit-image's pixel union-find): the image is never touched pixel by pixel
// during the union phase.  A *connectivity provider* describes each 32-pixel
// word with bit masks -- F (foreground), S (starts of within-word runs of equal
// value), U / UL / UR (pixel equals its upper / upper-left / upper-right
// neighbour) and J (pixel 0 continues the run of the previous word).  Nodes of
// the union-find forest are the within-word runs only.
//
// Node ids and storage.  id = word * SPW + ordinal, where word = y * WW + k is the
// slice-local word index and ordinal counts the runs of that word that start below
// it (SPW = 16 for binary masks -- a 32-bit word holds at most 16 runs -- and 32 for
// multi-valued images).  Ids ascend in raster order of the runs' first pixels, so
// a union that keeps the smaller id as root (atomicMin) makes every root the first
// raster pixel of its component: numbering roots in id order (a two-level scan over
// 32-word chunks) reproduces scipy / scikit-image label order bit-exactly.
// Parents are stored PLANE-MAJOR: slot(id) = ordinal * NW + word (NW = H * WW).
// Plane 0 -- the first run of every word, 96 % of all runs on blob-like masks -- is a
// dense int32 image of 4 B per word (0.125 B / pixel) that stays in L2 and shares
// sectors between neighbouring words; the other planes are touched only where a
// word really has several runs.  (Round 1 addressed a node by the pixel position of
// its run start, one int per PIXEL: every parent access was its own 64-byte DRAM
// transaction, and the pointer-chasing kernels were bound by exactly that.)
#pragma once
#include "pcs_common.cuh"

struct PcsConnWords {
  uint32_t F, S, U, UL, UR;
  int J;
  uint32_t Sa[3];  // run-start words of the row above, columns k-1, k, k+1
};

// -------------------------------------------------------------- binary provider
// Connectivity derived on the fly from a bit-row mask (optionally inverted).
struct PcsBinProv {
  const uint32_t* bits;  // slice base
  int H, W, WW;
  int inv;
  __device__ __forceinline__ PcsBinProv slice(long long b) const {
    PcsBinProv p = *this;
    p.bits = bits + b * (long long)H * WW;
    return p;
  }
  __device__ __forceinline__ uint32_t word(int y, int k) const {
    if (y < 0 || y >= H || k < 0 || k >= WW) return 0u;
    uint32_t w = __ldg(bits + (long long)y * WW + k);
    return inv ? (~w & pcs_valid_mask(k, W)) : w;
  }
  __device__ __forceinline__ void FS(int y, int k, uint32_t& F, uint32_t& S) const {
    F = word(y, k);
    S = F & ~(F << 1);
  }
  __device__ __forceinline__ void conn(int y, int k, PcsConnWords& c) const {
    uint32_t f = word(y, k);
    c.F = f;
    c.S = f & ~(f << 1);
    uint32_t l = word(y, k - 1);
    c.J = (f & 1u) && (l >> 31);
    uint32_t al = word(y - 1, k - 1), ac = word(y - 1, k), ar = word(y - 1, k + 1);
    c.U = f & ac;
    c.UL = f & ((ac << 1) | (al >> 31));
    c.UR = f & ((ac >> 1) | (ar << 31));
    c.Sa[0] = al & ~(al << 1);
    c.Sa[1] = ac & ~(ac << 1);
    c.Sa[2] = ar & ~(ar << 1);
  }
  __device__ __forceinline__ uint32_t Sword(int y, int k) const {
    uint32_t f = word(y, k);
    return f & ~(f << 1);
  }
};

// -------------------------------------------------------------- general provider
// Connectivity planes precomputed from a multi-valued image (six uint32 planes
// of shape (B, H, WW): F, S, U, UL, UR, J).
struct PcsGenProv {
  const uint32_t* planes;  // slice base of plane 0
  long long plane_stride;  // words between planes (= B*H*WW)
  int H, W, WW;
  __device__ __forceinline__ PcsGenProv slice(long long b) const {
    PcsGenProv p = *this;
    p.planes = planes + b * (long long)H * WW;
    return p;
  }
  __device__ __forceinline__ uint32_t ld(int plane, int y, int k) const {
    if (y < 0 || y >= H || k < 0 || k >= WW) return 0u;
    return __ldg(planes + plane * plane_stride + (long long)y * WW + k);
  }
  __device__ __forceinline__ void FS(int y, int k, uint32_t& F, uint32_t& S) const {
    F = ld(0, y, k);
    S = ld(1, y, k);
  }
  __device__ __forceinline__ void conn(int y, int k, PcsConnWords& c) const {
    c.F = ld(0, y, k);
    c.S = ld(1, y, k);
    c.U = ld(2, y, k);
    c.UL = ld(3, y, k);
    c.UR = ld(4, y, k);
    c.J = (int)(ld(5, y, k) & 1u);
    c.Sa[0] = ld(1, y - 1, k - 1);
    c.Sa[1] = ld(1, y - 1, k);
    c.Sa[2] = ld(1, y - 1, k + 1);
  }
  __device__ __forceinline__ uint32_t Sword(int y, int k) const { return ld(1, y, k); }
};

// -------------------------------------------------------------- run iteration
// Pops the lowest run of (F, S): returns its start bit and mask, clears it from S.
__device__ __forceinline__ uint32_t pcs_pop_run(uint32_t F, uint32_t& S, int& s) {
  s = __ffs(S) - 1;
  S &= S - 1;
  uint32_t upper = ~(F >> s);
  int len = upper ? (__ffs(upper) - 1) : 32;
  if (S) {
    int ns = __ffs(S) - 1 - s;
    len = min(len, ns);
  }
  uint32_t m = (len >= 32) ? 0xffffffffu : ((1u << len) - 1u);
  return m << s;
}

// -------------------------------------------------------------- node ids
template <class P> struct PcsNodes { static constexpr int LOG_SPW = 4; };            // binary masks: <= 16 runs per word
template <> struct PcsNodes<PcsGenProv> { static constexpr int LOG_SPW = 5; };       // multi-valued: <= 32
template <int LSPW> __device__ __forceinline__ int pcs_node(int word, int ord) { return (word << LSPW) | ord; }
template <int LSPW> __device__ __forceinline__ int pcs_slot(int id, int NW) { return (id & ((1 << LSPW) - 1)) * NW + (id >> LSPW); }
// ordinal of the run of S (run-start word) that starts at bit s
__device__ __forceinline__ int pcs_run_ord(uint32_t S, int s) { return __popc(S & ((1u << s) - 1u)); }

// -------------------------------------------------------------- union-find
template <int LSPW> __device__ __forceinline__ int pcs_uf_find(int* par, int NW, int n) {
  int r = n, p = pcs_ld_cg(par + pcs_slot<LSPW>(r, NW));
  int first = p;
  while (p != r) {
    r = p;
    p = pcs_ld_cg(par + pcs_slot<LSPW>(r, NW));
  }
  if (first != r) atomicMin(par + pcs_slot<LSPW>(n, NW), r);  // one-step compression (monotone, race-safe)
  return r;
}

template <int LSPW> __device__ __forceinline__ void pcs_uf_union(int* par, int NW, int a, int b) {
  while (true) {
    a = pcs_uf_find<LSPW>(par, NW, a);
    b = pcs_uf_find<LSPW>(par, NW, b);
    if (a == b) return;
    if (a < b) {
      int t = a;
      a = b;
      b = t;
    }
    int old = atomicMin(par + pcs_slot<LSPW>(a, NW), b);  // link the larger root under the smaller
    if (old == a) return;
    a = old;  // a had been linked meanwhile: keep uniting its (old) parent with b
  }
}

// Workspace layout of one connected-component job.
struct PcsCclWs {
  int* parent;         // B * SPW * NW ints in use (B * H * Wp allocated: enough for SPW = 32)
  uint32_t* rootbits;  // B * H * WW  (bit j: the word's run of ordinal j is a root)
  int* chunk;          // B * H * CPR  (roots per 32-word chunk, then exclusive base)
  int* offsets;        // B + 1        (exclusive scan of per-slice counts)
  int* aux;            // like parent   (optional: per-root accumulators, e.g. area)
};

size_t pcs_ccl_ws_bytes(int B, int H, int W, int with_aux);
int pcs_ccl_ws_carve(void* ws, size_t ws_bytes, int B, int H, int W, int with_aux, PcsCclWs* out);

// ---- pieces of the labeller the fused segment pipeline drives itself (pcs_segment.cu, pcs_pipeline.cu)
// exclusive scan of ws.chunk per slice (in place), per-slice totals -> counts, exclusive scan of those -> ws.offsets
int pcs_ccl_scan_offsets(const PcsCclWs& ws, int32_t* counts, int B, int H, int W, cudaStream_t st);
// labels of a binary mask whose parent planes hold the (negated) label or the root of every run -> int32 image
int pcs_ccl_relabel_bin(const uint32_t* bits, const PcsCclWs& ws, int32_t* labels, int B, int H, int W, cudaStream_t st);
int pcs_seg_label_stage(const uint16_t* img, const int32_t* thr, int median, uint32_t* bits, uint8_t* mask, int32_t* labels,
                        int32_t* counts, int64_t* table, int64_t cap, const PcsCclWs& ws, int* rsum, int* wlist, int* wcount, int B, int H,
                        int W, cudaStream_t st);
int pcs_seg_refine_stage(const uint32_t* bits, const int* labpar, const int64_t* table, int64_t cap, const int32_t* offsets,
                         int64_t min_size, uint32_t* out, uint8_t* out_mask, uint32_t* cand, int* hpar, int* clist, int* ccount, int B,
                         int H, int W, cudaStream_t st);
