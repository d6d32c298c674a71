// Shared device/host helpers for libpcs (sm_100a).
//
// Data model used by every kernel in this library:
//   * images are batches of independent 2-D slices, shape (B, H, W), C order
//     (split_zstack.py:52 iterates slices; tiff_analysis.py:727-737 only ever
//     processes 2-D images), one launch covers the whole batch;
//   * binary masks travel as BIT ROWS: uint32 words, WW = ceil(W/32) words per
//     row, bit j of word k is pixel x = 32*k + j; bits at x >= W are always 0;
//   * connected-component nodes are the starts of within-word runs, addressed
//     in the padded index space  node = y * (32*WW) + x  (raster order kept).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PCS_OK 0
#define PCS_ERR_INVALID (-1)
#define PCS_ERR_CUDA (-2)
#define PCS_ERR_WORKSPACE (-3)
#define PCS_ERR_UNSUPPORTED (-4)

extern "C" void pcs_set_error(const char* msg);
int pcs_check_launch(const char* what);
// launch accounting / optional per-kernel CUDA-event timing (pcs_profile_* in pcs.h)
void pcs_prof_begin(const char* name, cudaStream_t st);
void pcs_prof_end(cudaStream_t st);
#define PCS_LAUNCH(name, st, ...)  \
  do {                             \
    pcs_prof_begin(name, st);      \
    __VA_ARGS__;                   \
    pcs_prof_end(st);              \
  } while (0)

#define PCS_REQUIRE(cond, msg)        \
  do {                                \
    if (!(cond)) {                    \
      pcs_set_error(msg);             \
      return PCS_ERR_INVALID;         \
    }                                 \
  } while (0)

static inline int pcs_words(int W) { return (W + 31) >> 5; }
// true the first time it is called for (slot, current device): kernels that need a function attribute
// (opt-in shared memory) set it once per device, so a process that drives several GPUs stays correct
static inline bool pcs_first_use(bool (&done)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return true;
  if (done[dev]) return false;
  done[dev] = true;
  return true;
}

static inline size_t pcs_align256(size_t n) { return (n + 255) & ~(size_t)255; }
static inline unsigned pcs_blocks(long long n, int per_block) {
  long long b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > 0x7fffffffLL) b = 0x7fffffffLL;
  return (unsigned)b;
}

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ uint32_t pcs_valid_mask(int k, int W) {
  // valid-pixel mask of word k in a row of W pixels
  int rem = W - (k << 5);
  return rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
}

__device__ __forceinline__ int pcs_ld_cg(const int* p) { return __ldcg(p); }
// t = (i2 * n1 + i1) * n0 + i0 with 32-bit divisions whenever t fits (64-bit division by a
// run-time value costs ~100 instructions, which dominated the word-parallel kernels)
__device__ __forceinline__ void pcs_split3(long long t, int n0, int n1, int& i0, int& i1, long long& i2) {
  if (t <= 0xffffffffLL) {
    const unsigned u = (unsigned)t, q = u / (unsigned)n0, q2 = q / (unsigned)n1;
    i0 = (int)(u - q * (unsigned)n0);
    i1 = (int)(q - q2 * (unsigned)n1);
    i2 = q2;
  } else {
    const long long q = t / n0, q2 = q / n1;
    i0 = (int)(t - q * n0);
    i1 = (int)(q - q2 * n1);
    i2 = q2;
  }
}


// start bit of the within-word run of `w` that contains bit j (bit j must be set)
__device__ __forceinline__ int pcs_run_start(uint32_t w, int j) {
  uint32_t zeros_below = ~w & ((1u << j) - 1u);
  return zeros_below ? 32 - __clz(zeros_below) : 0;
}
// highest set bit of s at or below j (s must have one)
__device__ __forceinline__ int pcs_start_at_or_below(uint32_t s, int j) {
  uint32_t m = s & (0xffffffffu >> (31 - j));
  return 31 - __clz(m);
}

__device__ __forceinline__ int pcs_warp_excl_scan(int v, int lane, int* total) {
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += t;
  }
  *total = __shfl_sync(0xffffffffu, x, 31);
  return x - v;
}

// 32 mask bits -> 32 bytes (0/1) with two 16-byte stores when the row allows it
__device__ __forceinline__ void pcs_store_mask_bytes(uint8_t* __restrict__ row, int k, int W, uint32_t w) {
  const int x0 = k << 5;
  if (x0 + 32 <= W && ((((uintptr_t)(row + x0)) & 15) == 0)) {
    uint32_t o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t n = (w >> (4 * i)) & 0xfu;
      // spread 4 bits to 4 bytes: n * (1 + 2^7 + 2^14 + 2^21) puts bit j at position 8j (no carries: the
      // four shifted copies of a 4-bit value do not overlap), the mask keeps just those
      o[i] = (n * 0x00204081u) & 0x01010101u;
    }
    uint4* dst = reinterpret_cast<uint4*>(row + x0);
    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
  } else {
    int n = min(32, W - x0);
    for (int i = 0; i < n; ++i) row[x0 + i] = (w >> i) & 1u;
  }
}

// ---------------------------------------------------------------- binary 5x5 median helpers (scipy 'reflect' borders)
__device__ __forceinline__ int pcs_reflect(int i, int n) {
  // scipy.ndimage mode='reflect': (d c b a | a b c d | d c b a)
  if (i < 0) i = -i - 1;
  if (i >= n) i = 2 * n - i - 1;
  return i;
}

__device__ __forceinline__ uint32_t pcs_getbit(const uint32_t* row, int x) { return (row[x >> 5] >> (x & 31)) & 1u; }

// 64-bit window of a bit row: window bit i <-> x = 32k - 16 + i, reflected at the row ends
__device__ __forceinline__ unsigned long long pcs_window_reflect(const uint32_t* row, int k, int W, int WW, int r) {
  unsigned long long win = (unsigned long long)row[k] << 16;
  if (k > 0) win |= row[k - 1] >> 16;
  if (k + 1 < WW) win |= (unsigned long long)(row[k + 1] & 0xffffu) << 48;
  if (k == 0)
    for (int q = 1; q <= r; ++q) win |= (unsigned long long)pcs_getbit(row, q - 1) << (16 - q);
  const int xhi = (k << 5) + 47;
  if (xhi >= W)
    for (int q = 1; q <= r; ++q) {
      int x = W - 1 + q;
      int i = x - (k << 5) + 16;
      if (i >= 0 && i < 64) win |= (unsigned long long)pcs_getbit(row, W - q) << i;
    }
  return win;
}

// 5x5 binary median, bit-sliced: the 25 neighbours of 32 pixels are counted with carry-save
// adders on whole words (about 8 instructions per pixel instead of one popc per pixel and row).
__device__ __forceinline__ void pcs_add5(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t& s0, uint32_t& s1,
                                         uint32_t& s2) {
  const uint32_t t = a ^ b ^ c, m1 = (a & b) | (c & (a ^ b));
  s0 = t ^ d ^ e;
  const uint32_t m2 = (t & d) | (e & (t ^ d));
  s1 = m1 ^ m2;
  s2 = m1 & m2;
}

// the five rows' bit-sliced horizontal counts (r0 = ones, r1 = twos, r2 = fours) -> bits where at least 13 of the
// 25 neighbours are set (the median of a 5x5 binary window)
__device__ __forceinline__ uint32_t pcs_majority5_word(const uint32_t (&r0)[5], const uint32_t (&r1)[5], const uint32_t (&r2)[5]) {
  uint32_t a0, a1, a2, b1, b2, b3, c2, c3, c4;
  pcs_add5(r0[0], r0[1], r0[2], r0[3], r0[4], a0, a1, a2);  // ones   (weight 1)
  pcs_add5(r1[0], r1[1], r1[2], r1[3], r1[4], b1, b2, b3);  // twos   (weight 2)
  pcs_add5(r2[0], r2[1], r2[2], r2[3], r2[4], c2, c3, c4);  // fours  (weight 4)
  // total = A + 2B + 4C, bit by bit
  const uint32_t s0 = a0;
  const uint32_t s1 = a1 ^ b1, k2 = a1 & b1;
  const uint32_t t2 = a2 ^ b2 ^ c2, m2 = (a2 & b2) | (c2 & (a2 ^ b2));
  const uint32_t s2 = t2 ^ k2, n2 = t2 & k2;
  const uint32_t t3 = b3 ^ c3 ^ m2, m3 = (b3 & c3) | (m2 & (b3 ^ c3));
  const uint32_t s3 = t3 ^ n2, n3 = t3 & n2;
  const uint32_t s4 = c4 ^ m3 ^ n3;
  return s4 | (s3 & s2 & (s1 | s0));  // total >= 13
}
