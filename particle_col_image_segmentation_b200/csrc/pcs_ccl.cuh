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
// the union-find forest are the run starts only, so parent traffic is sparse
// and lives in L2.  Union keeps the smaller raster index as root (atomicMin),
// which makes every root the first raster pixel of its component: numbering
// roots in raster order (a two-level scan over 32-word chunks) reproduces
// scipy / scikit-image label order bit-exactly.
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

// -------------------------------------------------------------- union-find
__device__ __forceinline__ int pcs_uf_find(int* par, int n) {
  int r = n, p = pcs_ld_cg(par + r);
  int first = p;
  while (p != r) {
    r = p;
    p = pcs_ld_cg(par + r);
  }
  if (first != r) atomicMin(par + n, r);  // one-step compression (monotone, race-safe)
  return r;
}

__device__ __forceinline__ void pcs_uf_union(int* par, int a, int b) {
  while (true) {
    a = pcs_uf_find(par, a);
    b = pcs_uf_find(par, b);
    if (a == b) return;
    if (a < b) {
      int t = a;
      a = b;
      b = t;
    }
    int old = atomicMin(par + a, b);  // link the larger root under the smaller
    if (old == a) return;
    a = old;  // a had been linked meanwhile: keep uniting its (old) parent with b
  }
}

// Workspace layout of one connected-component job.
struct PcsCclWs {
  int* parent;         // B * H * Wp
  uint32_t* rootbits;  // B * H * WW
  int* chunk;          // B * H * CPR  (roots per 32-word chunk, then exclusive base)
  int* offsets;        // B + 1        (exclusive scan of per-slice counts)
  int* aux;            // B * H * Wp   (optional: per-root accumulators, e.g. area)
};

size_t pcs_ccl_ws_bytes(int B, int H, int W, int with_aux);
int pcs_ccl_ws_carve(void* ws, size_t ws_bytes, int B, int H, int W, int with_aux, PcsCclWs* out);

