// K5 binary morphology on bit rows (flat footprint as horizontal runs) and
// K3 median filtering (uint8 images, and the binary majority special case).
//
// Replaces (file:line in /root/reference):
//   skimage.morphology.binary_dilation(mask, disk(2))   tiff_analysis.py:827-828
//   skimage.morphology.binary_dilation(mask, disk(20))  tiff_analysis.py:990 (via pcs_edt_bits)
//   scipy.ndimage.median_filter(ds_arr, size=5)         tiff_analysis.py:122, :643
// plus erosion / opening / closing (north_star; no reference call site) by duality:
//   erode(X, S, border) = ~dilate(~X, reflect(S), !border).
#include "pcs_common.cuh"

#include "pcs.h"

#define MORPH_THREADS 256

struct PcsRun {
  int dy, lo, hi;  // offsets covered: (dy, dx) for dx in [lo, hi], hi - lo + 1 <= 32
};

// out(y, x) = OR over runs, dx in [lo, hi] of in'(y - dy, x - dx), in' = in ^ invert_in,
// pixels outside the image read as `border`.
__global__ void __launch_bounds__(MORPH_THREADS)
    k_dilate_bits(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, const PcsRun* __restrict__ runs, int n_runs,
                  int invert_in, int border, int invert_out, int B, int H, int W, int WW) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * H * WW;
  if (t >= total) return;
  int k, y;
  long long b;
  pcs_split3(t, WW, H, k, y, b);
  const uint32_t* src = in + b * (long long)H * WW;
  const uint32_t bw = border ? 0xffffffffu : 0u;
  uint32_t acc = 0;
  for (int r = 0; r < n_runs; ++r) {
    const PcsRun run = runs[r];
    const int ys = y - run.dy;
    const int len = run.hi - run.lo + 1;
    const bool row_out = ys < 0 || ys >= H;
    // source word j smears onto x in [32j + lo, 32j + lo + 30 + len]; keep those meeting word k
    const int jmax = ((k << 5) + 31 - run.lo) >> 5;
    const int jmin = ((k << 5) - run.lo - len + 1) >> 5;  // arithmetic shift = floor
    for (int j = jmin; j <= jmax; ++j) {
      uint32_t w;
      if (row_out || j < 0 || j >= WW) {
        w = bw;
      } else {
        w = __ldg(src + (long long)ys * WW + j);
        const uint32_t vm = pcs_valid_mask(j, W);
        if (invert_in) w = ~w;
        w = (w & vm) | (bw & ~vm);
      }
      if (!w) continue;
      unsigned long long S = w;
      int cover = 1;
      while (cover * 2 <= len) {
        S |= S << cover;
        cover *= 2;
      }
      if (len > cover) S |= S << (len - cover);
      const int sft = 32 * (k - j) - run.lo;  // out bit t <- S bit (t + sft)
      uint32_t c;
      if (sft >= 0)
        c = sft < 64 ? (uint32_t)(S >> sft) : 0u;
      else
        c = -sft < 32 ? ((uint32_t)S) << (-sft) : 0u;
      acc |= c;
    }
  }
  const uint32_t vm = pcs_valid_mask(k, W);
  out[t] = (invert_out ? ~acc : acc) & vm;
}

// ---------------------------------------------------------------- median
// (pcs_reflect / pcs_window_reflect / pcs_add5 / pcs_majority5_word live in pcs_common.cuh: the fused pipeline kernel uses them too)
// binary median (majority) of a size x size window, mode reflect; thread per word
template <int size>
__global__ void __launch_bounds__(MORPH_THREADS)
    k_majority_bits(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int B, int H, int W, int WW) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * H * WW;
  if (t >= total) return;
  int k, y;
  long long b;
  pcs_split3(t, WW, H, k, y, b);
  const uint32_t* src = in + b * (long long)H * WW;
  const int r = size >> 1;
  const int need = (size * size) / 2 + 1;  // ones needed for the median to be 1
  const uint32_t wm = (1u << size) - 1u;
  unsigned long long win[size];
#pragma unroll
  for (int dy = -r; dy <= r; ++dy) win[dy + r] = pcs_window_reflect(src + (long long)pcs_reflect(y + dy, H) * WW, k, W, WW, r);
  uint32_t o = 0;
  for (int tb = 0; tb < 32; ++tb) {
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < size; ++q) cnt += __popc((uint32_t)(win[q] >> (tb + 16 - r)) & wm);
    o |= (uint32_t)(cnt >= need) << tb;
  }
  out[t] = o & pcs_valid_mask(k, W);
}

#define MAJ_ROWS 8  // output rows per thread: each input row's horizontal sums are computed once and reused 5 times
__global__ void __launch_bounds__(MORPH_THREADS)
    k_majority5_bits(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint8_t* __restrict__ mask, int B, int H, int W,
                     int WW, int strips) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * strips * WW;
  if (t >= total) return;
  int k, strip;
  long long b;
  pcs_split3(t, WW, strips, k, strip, b);
  const uint32_t* src = in + b * (long long)H * WW;
  const int y0 = strip * MAJ_ROWS;
  const uint32_t vm = pcs_valid_mask(k, W);
  uint32_t r0[5], r1[5], r2[5];  // sliding window: bit-sliced count of the 5 horizontal neighbours per input row
#pragma unroll
  for (int i = 0; i < MAJ_ROWS + 4; ++i) {
    const int yi = y0 - 2 + i;  // input row entering the window
    if (yi - 2 < H) {           // still needed by some output row of the image
      const unsigned long long win = pcs_window_reflect(src + (long long)pcs_reflect(yi, H) * WW, k, W, WW, 2);
      pcs_add5((uint32_t)(win >> 14), (uint32_t)(win >> 15), (uint32_t)(win >> 16), (uint32_t)(win >> 17), (uint32_t)(win >> 18),
               r0[i % 5], r1[i % 5], r2[i % 5]);
    }
    if (i >= 4) {
      const int y = y0 + i - 4;  // output row whose 5 input rows are now in the window
      if (y < H) {
        const uint32_t ge13 = pcs_majority5_word(r0, r1, r2) & vm;
        out[(b * H + y) * (long long)WW + k] = ge13;
        if (mask) pcs_store_mask_bytes(mask + (b * H + y) * (long long)W, k, W, ge13);  // fused uint8 output
      }
    }
  }
}

// generic uint8 median, size in {3, 5, 7}, mode reflect; CTA tile 32 x 8 with halo in shared memory
#define MED_TX 32
#define MED_TY 8
template <int size>
__global__ void __launch_bounds__(MED_TX* MED_TY)
    k_median_u8(const uint8_t* __restrict__ img, uint8_t* __restrict__ out, int H, int W) {
  __shared__ uint8_t tile[MED_TY + 6][MED_TX + 6 + 2];
  const int r = size >> 1;
  const int rank = (size * size) / 2;
  const long long b = blockIdx.z;
  const uint8_t* src = img + b * (long long)H * W;
  const int x0 = blockIdx.x * MED_TX, y0 = blockIdx.y * MED_TY;
  const int tw = MED_TX + 2 * r, th = MED_TY + 2 * r;
  for (int i = threadIdx.y * MED_TX + threadIdx.x; i < tw * th; i += MED_TX * MED_TY) {
    int ty = i / tw, tx = i % tw;
    // tile cells far outside the image (partial edge tiles) are never read back: clamp them
    int yy = min(max(pcs_reflect(y0 + ty - r, H), 0), H - 1), xx = min(max(pcs_reflect(x0 + tx - r, W), 0), W - 1);
    tile[ty][tx] = src[(long long)yy * W + xx];
  }
  __syncthreads();
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  if (x >= W || y >= H) return;
  // largest v with #(window < v) <= rank is the rank-th smallest value
  int v = 0;
  for (int bit = 7; bit >= 0; --bit) {
    const int cand = v | (1 << bit);
    int cnt = 0;
#pragma unroll
    for (int dy = 0; dy < size; ++dy)
#pragma unroll
      for (int dx = 0; dx < size; ++dx) cnt += tile[threadIdx.y + dy][threadIdx.x + dx] < cand;
    if (cnt <= rank) v = cand;
  }
  out[b * (long long)H * W + (long long)y * W + x] = (uint8_t)v;
}

// ---------------------------------------------------------------- images smaller than the window
// scipy's 'reflect' keeps folding an index until it lands inside the image ((d c b a | a b c d | d c b a)
// repeated), which only matters when a side is shorter than the window.  These per-pixel kernels do the
// general fold; they are only launched for such tiny images, where speed is irrelevant.
__device__ __forceinline__ int pcs_reflect_any(int i, int n) {
  const int p = 2 * n;
  i %= p;
  if (i < 0) i += p;
  return i < n ? i : p - 1 - i;
}

__global__ void __launch_bounds__(256)
    k_majority_small(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int size, int B, int H, int W, int WW) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // thread per output word
  if (t >= (long long)B * H * WW) return;
  int k, y;
  long long b;
  pcs_split3(t, WW, H, k, y, b);
  const uint32_t* src = in + b * (long long)H * WW;
  const int r = size >> 1, need = (size * size) / 2 + 1;
  uint32_t o = 0;
  for (int j = 0; j < 32 && (k << 5) + j < W; ++j) {
    int cnt = 0;
    for (int dy = -r; dy <= r; ++dy)
      for (int dx = -r; dx <= r; ++dx) {
        const int yy = pcs_reflect_any(y + dy, H), xx = pcs_reflect_any((k << 5) + j + dx, W);
        cnt += (src[(long long)yy * WW + (xx >> 5)] >> (xx & 31)) & 1u;
      }
    o |= (uint32_t)(cnt >= need) << j;
  }
  out[t] = o;
}

__global__ void __launch_bounds__(256)
    k_median_small(const uint8_t* __restrict__ img, uint8_t* __restrict__ out, int size, int B, int H, int W) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // thread per pixel
  if (t >= (long long)B * H * W) return;
  int x, y;
  long long b;
  pcs_split3(t, W, H, x, y, b);
  const uint8_t* src = img + b * (long long)H * W;
  const int r = size >> 1, rank = (size * size) / 2;
  int v = 0;  // largest v with #(window < v) <= rank is the rank-th smallest value
  for (int bit = 7; bit >= 0; --bit) {
    const int cand = v | (1 << bit);
    int cnt = 0;
    for (int dy = -r; dy <= r; ++dy)
      for (int dx = -r; dx <= r; ++dx) cnt += src[(long long)pcs_reflect_any(y + dy, H) * W + pcs_reflect_any(x + dx, W)] < cand;
    if (cnt <= rank) v = cand;
  }
  out[t] = (uint8_t)v;
}

extern "C" {

int pcs_dilate_bits(const uint32_t* in, uint32_t* out, const int32_t* runs, int n_runs, int invert_in, int border,
                    int invert_out, int B, int H, int W, void* stream) {
  PCS_REQUIRE(B >= 1 && H >= 1 && W >= 1, "empty batch or image");
  PCS_REQUIRE(n_runs >= 1 && runs != nullptr, "empty footprint");
  PCS_REQUIRE(in != out, "dilation cannot run in place");
  int WW = pcs_words(W);
  PCS_LAUNCH("k_dilate_bits", (cudaStream_t)stream, k_dilate_bits<<<pcs_blocks((long long)B * H * WW, MORPH_THREADS), MORPH_THREADS, 0, (cudaStream_t)stream>>>(
      in, out, (const PcsRun*)runs, n_runs, invert_in, border, invert_out, B, H, W, WW));
  return pcs_check_launch("dilate");
}

int pcs_majority_bits(const uint32_t* in, uint32_t* out, int size, int B, int H, int W, void* stream) {
  return pcs_majority_bits_mask(in, out, nullptr, size, B, H, W, stream);
}

int pcs_majority_bits_mask(const uint32_t* in, uint32_t* out, uint8_t* mask, int size, int B, int H, int W, void* stream) {
  PCS_REQUIRE(B >= 1 && H >= 1 && W >= 1, "empty batch or image");
  PCS_REQUIRE(size == 3 || size == 5 || size == 7, "median size must be 3, 5 or 7");
  PCS_REQUIRE(in != out, "median cannot run in place");
  int WW = pcs_words(W);
  unsigned g = pcs_blocks((long long)B * H * WW, MORPH_THREADS);
  cudaStream_t st = (cudaStream_t)stream;
  if (H < size || W < size) {  // a side shorter than the window: general reflect fold
    PCS_LAUNCH("k_majority_small", st, k_majority_small<<<pcs_blocks((long long)B * H * WW, 256), 256, 0, st>>>(in, out, size, B, H, W, WW));
    if (mask) return pcs_unpack_bits(out, mask, B, H, W, stream);
    return pcs_check_launch("majority");
  }
  if (size == 3)
    PCS_LAUNCH("k_majority_bits", st, k_majority_bits<3><<<g, MORPH_THREADS, 0, st>>>(in, out, B, H, W, WW));
  else if (size == 5)
  {
    const int strips = (H + MAJ_ROWS - 1) / MAJ_ROWS;
    PCS_LAUNCH("k_majority5_bits", st,
               k_majority5_bits<<<pcs_blocks((long long)B * strips * WW, MORPH_THREADS), MORPH_THREADS, 0, st>>>(in, out, mask, B, H, W, WW, strips));
  }
  else
    PCS_LAUNCH("k_majority_bits", st, k_majority_bits<7><<<g, MORPH_THREADS, 0, st>>>(in, out, B, H, W, WW));
  if (mask && size != 5) return pcs_unpack_bits(out, mask, B, H, W, stream);
  return pcs_check_launch("majority");
}

int pcs_median_u8(const uint8_t* img, uint8_t* out, int size, int B, int H, int W, void* stream) {
  PCS_REQUIRE(B >= 1 && H >= 1 && W >= 1, "empty batch or image");
  PCS_REQUIRE(size == 3 || size == 5 || size == 7, "median size must be 3, 5 or 7");
  PCS_REQUIRE(img != out, "median cannot run in place");
  PCS_REQUIRE(B <= 65535, "batch above 65535");
  if (H < size || W < size) {  // a side shorter than the window: general reflect fold
    PCS_LAUNCH("k_median_small", (cudaStream_t)stream, k_median_small<<<pcs_blocks((long long)B * H * W, 256), 256, 0, (cudaStream_t)stream>>>(img, out, size, B, H, W));
    return pcs_check_launch("median");
  }
  dim3 grid((W + MED_TX - 1) / MED_TX, (H + MED_TY - 1) / MED_TY, B);
  dim3 block(MED_TX, MED_TY);
  cudaStream_t st = (cudaStream_t)stream;
  if (size == 3)
    PCS_LAUNCH("k_median_u8", st, k_median_u8<3><<<grid, block, 0, st>>>(img, out, H, W));
  else if (size == 5)
    PCS_LAUNCH("k_median_u8", st, k_median_u8<5><<<grid, block, 0, st>>>(img, out, H, W));
  else
    PCS_LAUNCH("k_median_u8", st, k_median_u8<7><<<grid, block, 0, st>>>(img, out, H, W));
  return pcs_check_launch("median");
}

}  // extern "C"
