// K2 element-wise passes: thresholds, class-equality masks, LUT relabels, masked
// overwrite, bit packing, and the connectivity planes for multi-valued labelling.
//
// Replaces (file:line in /root/reference):
//   boundary_map < threshold                   refine_boundaries.py:44-45
//   ds_arr == label                            tiff_analysis.py:256-257, :812, :818, :984, :987
//   arr[arr == a] = b chains                   tiff_analysis.py:177-181, :224-231
//   base[other == 1] = val                     tiff_analysis.py:240
//   updated[overlap] = overlap_label, np.sum   tiff_analysis.py:1000-1015
//
// All of these are HBM-bound byte passes: one thread produces one 32-pixel mask
// word from 128-bit vector loads, so global traffic is the algorithmic minimum.
#include "pcs_common.cuh"

#include "pcs.h"

#define PW_THREADS 256

// ---------------------------------------------------------------- generic word builder
// Calls pred(value) for the 32 pixels of word (row, k) and packs the answers.
template <typename T, class Pred>
__device__ __forceinline__ uint32_t pcs_pack_word(const T* __restrict__ row, int k, int W, Pred pred) {
  const int x0 = k << 5;
  uint32_t w = 0;
  constexpr int VEC = 16 / sizeof(T);  // elements per 128-bit load
  if (x0 + 32 <= W && ((((uintptr_t)(row + x0)) & 15) == 0)) {
#pragma unroll
    for (int v = 0; v < 32 / VEC; ++v) {
      uint4 q = __ldg(reinterpret_cast<const uint4*>(row + x0) + v);
      const T* e = reinterpret_cast<const T*>(&q);
#pragma unroll
      for (int i = 0; i < VEC; ++i) w |= (uint32_t)(pred(e[i]) ? 1u : 0u) << (v * VEC + i);
    }
  } else {
    int n = min(32, W - x0);
    for (int i = 0; i < n; ++i) w |= (uint32_t)(pred(row[x0 + i]) ? 1u : 0u) << i;
  }
  return w;
}

#define PCS_WORD_INDEX                                            \
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; \
  long long total = (long long)B * H * WW;                        \
  if (t >= total) return;                                         \
  int k, y_;                                                      \
  long long b_;                                                   \
  pcs_split3(t, WW, H, k, y_, b_);                                \
  long long rowi = b_ * H + y_; /* b*H + y */

// cmp: 0 '>', 1 '>=', 2 '<', 3 '<=', 4 '==', 5 '!='
template <typename T, typename TT>
__device__ __forceinline__ bool pcs_cmp(T v, TT thr, int cmp) {
  switch (cmp) {
    case 0: return v > thr;
    case 1: return v >= thr;
    case 2: return v < thr;
    case 3: return v <= thr;
    case 4: return v == thr;
    default: return v != thr;
  }
}

// thr_dev: optional per-slice thresholds (device), else the scalar `thr`
template <typename T, typename TT>
__global__ void __launch_bounds__(PW_THREADS)
    k_compare(const T* __restrict__ img, TT thr, const TT* __restrict__ thr_dev, int cmp, uint32_t* __restrict__ bits,
              uint8_t* __restrict__ mask, int B, int H, int W, int WW) {
  PCS_WORD_INDEX
  TT th = thr_dev ? thr_dev[rowi / H] : thr;
  const T* row = img + rowi * (long long)W;
  uint32_t w = pcs_pack_word<T>(row, k, W, [=](T v) { return pcs_cmp<T, TT>(v, th, cmp); });
  if (bits) bits[t] = w;
  if (mask) pcs_store_mask_bytes(mask + rowi * (long long)W, k, W, w);
}

// values listed in a 256-entry membership table (covers == v and np.isin)
__global__ void __launch_bounds__(PW_THREADS)
    k_member_u8(const uint8_t* __restrict__ img, const uint8_t* __restrict__ member, uint32_t* __restrict__ bits,
                uint8_t* __restrict__ mask, int B, int H, int W, int WW) {
  __shared__ uint8_t m[256];
  if (threadIdx.x < 256) m[threadIdx.x] = member[threadIdx.x];
  __syncthreads();
  PCS_WORD_INDEX
  const uint8_t* row = img + rowi * (long long)W;
  uint32_t w = pcs_pack_word<uint8_t>(row, k, W, [&](uint8_t v) { return m[v] != 0; });
  if (bits) bits[t] = w;
  if (mask) pcs_store_mask_bytes(mask + rowi * (long long)W, k, W, w);
}

__global__ void __launch_bounds__(PW_THREADS)
    k_unpack(const uint32_t* __restrict__ bits, uint8_t* __restrict__ mask, int B, int H, int W, int WW) {
  PCS_WORD_INDEX
  pcs_store_mask_bytes(mask + rowi * (long long)W, k, W, bits[t]);
}

// op: 0 and, 1 or, 2 andnot (a & ~b), 3 xor, 4 not a
__global__ void __launch_bounds__(PW_THREADS)
    k_bits_logic(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint32_t* __restrict__ out, int op, int B,
                 int H, int W, int WW) {
  PCS_WORD_INDEX
  (void)rowi;
  uint32_t x = a[t], y = b ? b[t] : 0u, r;
  switch (op) {
    case 0: r = x & y; break;
    case 1: r = x | y; break;
    case 2: r = x & ~y; break;
    case 3: r = x ^ y; break;
    default: r = ~x & pcs_valid_mask(k, W); break;
  }
  out[t] = r;
}

// per-slice population count (np.sum of a mask, tiff_analysis.py:1015)
__global__ void __launch_bounds__(PW_THREADS)
    k_bits_count(const uint32_t* __restrict__ bits, unsigned long long* __restrict__ counts, long long words_per_slice) {
  long long b = blockIdx.y;
  const uint32_t* p = bits + b * words_per_slice;
  unsigned long long c = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < words_per_slice; i += (long long)gridDim.x * blockDim.x)
    c += __popc(p[i]);
  c = __reduce_add_sync(0xffffffffu, (unsigned)c);  // < 2^32 per warp is safe: <= 32 * words handled
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(counts + b, c);
}

// in-place 256-entry LUT (ordered arr[arr==a]=b chains collapse to one table)
__global__ void __launch_bounds__(PW_THREADS) k_lut_u8(uint8_t* __restrict__ img, const uint8_t* __restrict__ lut, long long n) {
  __shared__ uint8_t l[256];
  if (threadIdx.x < 256) l[threadIdx.x] = lut[threadIdx.x];
  __syncthreads();
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (i >= n) return;
  if (i + 16 <= n && ((((uintptr_t)(img + i)) & 15) == 0)) {
    uint4 q = *reinterpret_cast<uint4*>(img + i);
    uint8_t* e = reinterpret_cast<uint8_t*>(&q);
#pragma unroll
    for (int j = 0; j < 16; ++j) e[j] = l[e[j]];
    *reinterpret_cast<uint4*>(img + i) = q;
  } else {
    for (long long j = i; j < min(n, i + 16); ++j) img[j] = l[img[j]];
  }
}

// img[bits] = value (tiff_analysis.py:240, :1013, :286)
__global__ void __launch_bounds__(PW_THREADS)
    k_assign_where(uint8_t* __restrict__ img, const uint32_t* __restrict__ bits, uint8_t value, int B, int H, int W, int WW) {
  PCS_WORD_INDEX
  uint32_t w = bits[t];
  if (!w) return;
  uint8_t* row = img + rowi * (long long)W;
  const int x0 = k << 5;
  if (x0 + 32 <= W && ((((uintptr_t)(row + x0)) & 15) == 0)) {
    uint4* p = reinterpret_cast<uint4*>(row + x0);
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      uint32_t h = (w >> (16 * v)) & 0xffffu;
      if (!h) continue;
      uint4 q = p[v];
      uint8_t* e = reinterpret_cast<uint8_t*>(&q);
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if ((h >> j) & 1u) e[j] = value;
      p[v] = q;
    }
  } else {
    int n = min(32, W - x0);
    for (int i = 0; i < n; ++i)
      if ((w >> i) & 1u) row[x0 + i] = value;
  }
}

// ---------------------------------------------------------------- connectivity planes
// Six planes (F, S, U, UL, UR, J) describing equal-value adjacency of a
// multi-valued image; `all_fg` labels every pixel (plateaus), otherwise value 0
// is background (skimage.measure.label, tiff_analysis.py:743).
// `higher` (optional) gets a bit where some 8-/4-neighbour is strictly greater
// (skimage.morphology.local_maxima, refine_boundaries.py:63).
template <typename T>
__global__ void __launch_bounds__(PW_THREADS)
    k_conn_planes(const T* __restrict__ img, uint32_t* __restrict__ planes, uint32_t* __restrict__ higher, int all_fg,
                  int conn8, int B, int H, int W, int WW) {
  PCS_WORD_INDEX
  int y = (int)(rowi % H);
  const T* row = img + rowi * (long long)W;
  const T* up = y > 0 ? row - W : nullptr;
  const T* dn = y < H - 1 ? row + W : nullptr;
  const int x0 = k << 5;
  const int n = min(32, W - x0);
  uint32_t F = 0, S = 0, U = 0, UL = 0, UR = 0, Hh = 0;
  T prev = x0 > 0 ? row[x0 - 1] : T(0);
  bool prev_ok = x0 > 0;
  for (int i = 0; i < n; ++i) {
    int x = x0 + i;
    T v = row[x];
    bool fg = all_fg || v != T(0);
    bool has_next = x + 1 < W;
    if (fg) {
      F |= 1u << i;
      bool eq_left = prev_ok && prev == v;
      if (i == 0 || !eq_left) S |= 1u << i;
      if (up) {
        if (up[x] == v) U |= 1u << i;
        if (x > 0 && up[x - 1] == v) UL |= 1u << i;
        if (has_next && up[x + 1] == v) UR |= 1u << i;
      }
    }
    if (higher) {
      bool hi = false;
      if (prev_ok && prev > v) hi = true;
      if (has_next && row[x + 1] > v) hi = true;
      if (up) {
        if (up[x] > v) hi = true;
        if (conn8 && ((x > 0 && up[x - 1] > v) || (has_next && up[x + 1] > v))) hi = true;
      }
      if (dn) {
        if (dn[x] > v) hi = true;
        if (conn8 && ((x > 0 && dn[x - 1] > v) || (has_next && dn[x + 1] > v))) hi = true;
      }
      if (hi) Hh |= 1u << i;
    }
    prev = v;
    prev_ok = true;
  }
  int J = 0;
  if (x0 > 0 && (F & 1u)) J = row[x0 - 1] == row[x0];
  long long ps = (long long)B * H * WW;
  planes[t] = F;
  planes[ps + t] = S;
  planes[2 * ps + t] = U;
  planes[3 * ps + t] = UL;
  planes[4 * ps + t] = UR;
  planes[5 * ps + t] = (uint32_t)J;
  if (higher) higher[t] = Hh;
}

// gather table values: out[i] = img[slice[i]][idx[i]]  (class at the first pixel,
// tiff_analysis.py:1041-1044; dilated label under a centroid, :845-852)
template <typename T>
__global__ void k_gather(const T* __restrict__ img, const long long* __restrict__ slice, const long long* __restrict__ idx,
                         long long* __restrict__ out, long long n, long long slice_elems) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long s = slice ? slice[i] : 0;
  const long long j = idx[i];
  // an index outside the slice reads nothing (0 = background for every caller): a centroid of a region measured on
  // another, larger image must not become an out-of-bounds load
  out[i] = (j >= 0 && j < slice_elems && s >= 0) ? (long long)img[s * slice_elems + j] : 0;
}

// ---------------------------------------------------------------- C ABI
static int dims_ok(int B, int H, int W) {
  PCS_REQUIRE(B >= 1 && H >= 1 && W >= 1, "empty batch or image");
  return PCS_OK;
}

template <typename T, typename TT>
static int launch_compare(const void* img, TT thr, const void* thr_dev, int cmp, uint32_t* bits, uint8_t* mask, int B, int H,
                          int W, void* stream) {
  int rc = dims_ok(B, H, W);
  if (rc) return rc;
  PCS_REQUIRE(cmp >= 0 && cmp <= 5, "bad comparison code");
  PCS_REQUIRE(bits || mask, "no output requested");
  int WW = pcs_words(W);
  PCS_LAUNCH("k_compare", (cudaStream_t)stream, k_compare<T, TT><<<pcs_blocks((long long)B * H * WW, PW_THREADS), PW_THREADS, 0, (cudaStream_t)stream>>>(
      (const T*)img, thr, (const TT*)thr_dev, cmp, bits, mask, B, H, W, WW));
  return pcs_check_launch("compare");
}

// grid-stride fill with 128-bit stores: the store-only bandwidth probe bench.py reports next to the copy
// figure (a plain kernel writes HBM at ~6.8 TB/s on B200; torch's uint8 fill_ reaches only ~3.9)
__global__ void __launch_bounds__(256) k_fill128(uint4* __restrict__ p, uint32_t v, size_t n16) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n16; i += stride) p[i] = make_uint4(v, v, v, v);
}

// largest label of an int32 / int64 label image (regionprops on a label image whose label count is not known)
template <typename T>
__global__ void __launch_bounds__(256) k_max_label(const T* __restrict__ lab, long long n, long long* __restrict__ out) {
  long long m = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) m = max(m, (long long)lab[i]);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

// out (W, H) = in (H, W) transposed, 1- or 4-byte elements (MATLAB's column-major component order = raster order of
// the transposed mask, nanosims.matlab_label); 32 x 32 tiles through shared memory
template <typename T>
__global__ void __launch_bounds__(256) k_transpose(const T* __restrict__ in, T* __restrict__ out, int H, int W) {
  __shared__ T tile[32][33];
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8)
    if (y0 + r < H && x0 + tx < W) tile[r][tx] = in[(long long)(y0 + r) * W + x0 + tx];
  __syncthreads();
  for (int r = ty; r < 32; r += 8)
    if (x0 + r < W && y0 + tx < H) out[(long long)(x0 + r) * H + y0 + tx] = tile[tx][r];
}

extern "C" {

int pcs_compare_u16(const uint16_t* img, int thr, const int32_t* thr_dev, int cmp, uint32_t* bits, uint8_t* mask, int B,
                    int H, int W, void* stream) {
  return launch_compare<uint16_t, int>(img, thr, thr_dev, cmp, bits, mask, B, H, W, stream);
}
int pcs_compare_u8(const uint8_t* img, int thr, const int32_t* thr_dev, int cmp, uint32_t* bits, uint8_t* mask, int B,
                   int H, int W, void* stream) {
  return launch_compare<uint8_t, int>(img, thr, thr_dev, cmp, bits, mask, B, H, W, stream);
}
int pcs_compare_i32(const int32_t* img, int thr, const int32_t* thr_dev, int cmp, uint32_t* bits, uint8_t* mask, int B,
                    int H, int W, void* stream) {
  return launch_compare<int32_t, int>(img, thr, thr_dev, cmp, bits, mask, B, H, W, stream);
}
int pcs_compare_f32(const float* img, float thr, const float* thr_dev, int cmp, uint32_t* bits, uint8_t* mask, int B, int H,
                    int W, void* stream) {
  return launch_compare<float, float>(img, thr, thr_dev, cmp, bits, mask, B, H, W, stream);
}
int pcs_compare_f64(const double* img, double thr, const double* thr_dev, int cmp, uint32_t* bits, uint8_t* mask, int B,
                    int H, int W, void* stream) {
  return launch_compare<double, double>(img, thr, thr_dev, cmp, bits, mask, B, H, W, stream);
}

int pcs_member_u8(const uint8_t* img, const uint8_t* member256, uint32_t* bits, uint8_t* mask, int B, int H, int W,
                  void* stream) {
  int rc = dims_ok(B, H, W);
  if (rc) return rc;
  int WW = pcs_words(W);
  PCS_LAUNCH("k_member_u8", (cudaStream_t)stream, k_member_u8<<<pcs_blocks((long long)B * H * WW, PW_THREADS), PW_THREADS, 0, (cudaStream_t)stream>>>(img, member256, bits,
                                                                                                     mask, B, H, W, WW));
  return pcs_check_launch("member");
}

int pcs_unpack_bits(const uint32_t* bits, uint8_t* mask, int B, int H, int W, void* stream) {
  int rc = dims_ok(B, H, W);
  if (rc) return rc;
  int WW = pcs_words(W);
  PCS_LAUNCH("k_unpack", (cudaStream_t)stream, k_unpack<<<pcs_blocks((long long)B * H * WW, PW_THREADS), PW_THREADS, 0, (cudaStream_t)stream>>>(bits, mask, B, H, W, WW));
  return pcs_check_launch("unpack");
}

int pcs_bits_logic(const uint32_t* a, const uint32_t* b, uint32_t* out, int op, int B, int H, int W, void* stream) {
  int rc = dims_ok(B, H, W);
  if (rc) return rc;
  PCS_REQUIRE(op >= 0 && op <= 4, "bad logic op");
  PCS_REQUIRE(op == 4 || b != nullptr, "binary op needs two operands");
  int WW = pcs_words(W);
  PCS_LAUNCH("k_bits_logic", (cudaStream_t)stream, k_bits_logic<<<pcs_blocks((long long)B * H * WW, PW_THREADS), PW_THREADS, 0, (cudaStream_t)stream>>>(a, b, out, op, B, H,
                                                                                                      W, WW));
  return pcs_check_launch("bits logic");
}

int pcs_bits_count(const uint32_t* bits, uint64_t* counts, int B, int H, int W, void* stream) {
  int rc = dims_ok(B, H, W);
  if (rc) return rc;
  long long wps = (long long)H * pcs_words(W);
  cudaMemsetAsync(counts, 0, (size_t)B * 8, (cudaStream_t)stream);
  dim3 grid((unsigned)min((long long)64, (wps + PW_THREADS - 1) / PW_THREADS), B);
  PCS_LAUNCH("k_bits_count", (cudaStream_t)stream, k_bits_count<<<grid, PW_THREADS, 0, (cudaStream_t)stream>>>(bits, (unsigned long long*)counts, wps));
  return pcs_check_launch("bits count");
}

int pcs_lut_u8(uint8_t* img, const uint8_t* lut256, int64_t n, void* stream) {
  PCS_REQUIRE(n >= 1, "empty array");
  PCS_LAUNCH("k_lut_u8", (cudaStream_t)stream, k_lut_u8<<<pcs_blocks((n + 15) / 16, PW_THREADS), PW_THREADS, 0, (cudaStream_t)stream>>>(img, lut256, n));
  return pcs_check_launch("lut");
}

int pcs_assign_where_u8(uint8_t* img, const uint32_t* bits, int value, int B, int H, int W, void* stream) {
  int rc = dims_ok(B, H, W);
  if (rc) return rc;
  int WW = pcs_words(W);
  PCS_LAUNCH("k_assign_where", (cudaStream_t)stream, k_assign_where<<<pcs_blocks((long long)B * H * WW, PW_THREADS), PW_THREADS, 0, (cudaStream_t)stream>>>(img, bits,
                                                                                                        (uint8_t)value, B, H, W, WW));
  return pcs_check_launch("assign where");
}

size_t pcs_conn_planes_bytes(int B, int H, int W) { return (size_t)6 * B * H * pcs_words(W) * 4; }

// dtype: 0 u8, 1 u16, 2 i32, 3 f32, 4 f64
int pcs_conn_planes(const void* img, int dtype, uint32_t* planes, uint32_t* higher, int all_fg, int connectivity, int B, int H,
                    int W, void* stream) {
  int rc = dims_ok(B, H, W);
  if (rc) return rc;
  PCS_REQUIRE(connectivity == 4 || connectivity == 8, "connectivity must be 4 or 8");
  int WW = pcs_words(W);
  unsigned g = pcs_blocks((long long)B * H * WW, PW_THREADS);
  cudaStream_t st = (cudaStream_t)stream;
  int c8 = connectivity == 8;
  switch (dtype) {
    case 0: PCS_LAUNCH("k_conn_planes", st, k_conn_planes<uint8_t><<<g, PW_THREADS, 0, st>>>((const uint8_t*)img, planes, higher, all_fg, c8, B, H, W, WW)); break;
    case 1: PCS_LAUNCH("k_conn_planes", st, k_conn_planes<uint16_t><<<g, PW_THREADS, 0, st>>>((const uint16_t*)img, planes, higher, all_fg, c8, B, H, W, WW)); break;
    case 2: PCS_LAUNCH("k_conn_planes", st, k_conn_planes<int32_t><<<g, PW_THREADS, 0, st>>>((const int32_t*)img, planes, higher, all_fg, c8, B, H, W, WW)); break;
    case 3: PCS_LAUNCH("k_conn_planes", st, k_conn_planes<float><<<g, PW_THREADS, 0, st>>>((const float*)img, planes, higher, all_fg, c8, B, H, W, WW)); break;
    case 4: PCS_LAUNCH("k_conn_planes", st, k_conn_planes<double><<<g, PW_THREADS, 0, st>>>((const double*)img, planes, higher, all_fg, c8, B, H, W, WW)); break;
    default: pcs_set_error("unsupported dtype for connectivity planes"); return PCS_ERR_UNSUPPORTED;
  }
  return pcs_check_launch("connectivity planes");
}

// dtype: 0 u8, 2 i32, 5 i64
int pcs_gather(const void* img, int dtype, const int64_t* slice, const int64_t* idx, int64_t* out, int64_t n,
               int64_t slice_elems, void* stream) {
  if (n <= 0) return PCS_OK;
  unsigned g = pcs_blocks(n, 256);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case 0: PCS_LAUNCH("k_gather", st, k_gather<uint8_t><<<g, 256, 0, st>>>((const uint8_t*)img, (const long long*)slice, (const long long*)idx, (long long*)out, n, slice_elems)); break;
    case 2: PCS_LAUNCH("k_gather", st, k_gather<int32_t><<<g, 256, 0, st>>>((const int32_t*)img, (const long long*)slice, (const long long*)idx, (long long*)out, n, slice_elems)); break;
    case 5: PCS_LAUNCH("k_gather", st, k_gather<long long><<<g, 256, 0, st>>>((const long long*)img, (const long long*)slice, (const long long*)idx, (long long*)out, n, slice_elems)); break;
    default: pcs_set_error("unsupported dtype for gather"); return PCS_ERR_UNSUPPORTED;
  }
  return pcs_check_launch("gather");
}

// Background fill: a SMALL persistent grid (ctas_per_sm CTAs of 128 threads per SM, 256-bit stores) meant to run on a side
// stream beside kernels that leave DRAM idle.  Every CTA is resident from the start, so the block scheduler keeps
// dispatching the other streams' kernels into the rest of each SM (a grid with pending CTAs keeps them out).
__global__ void __launch_bounds__(128) k_fill_background(uint4* __restrict__ p, size_t n16) {
  size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 2;
  const size_t stride = (size_t)gridDim.x * blockDim.x * 2;
  for (; i + 1 < n16; i += stride)
    asm volatile("st.global.v8.u32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(p + i), "r"(0u) : "memory");
  if (i < n16) p[i] = make_uint4(0, 0, 0, 0);
}

int pcs_zero_background(void* dst, size_t bytes, int ctas_per_sm, void* stream) {
  PCS_REQUIRE(dst != nullptr && (bytes & 31) == 0 && ((((uintptr_t)dst) & 31) == 0), "background fill needs a 32-byte aligned buffer of 32k bytes");
  if (bytes == 0) return PCS_OK;
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  PCS_LAUNCH("k_fill_background", (cudaStream_t)stream, k_fill_background<<<sms * ctas_per_sm, 128, 0, (cudaStream_t)stream>>>((uint4*)dst, bytes / 16));
  return pcs_check_launch("background fill");
}

int pcs_fill_u32(void* dst, uint32_t value, size_t n_words, void* stream) {
  PCS_REQUIRE(dst != nullptr && (n_words & 3) == 0 && ((((uintptr_t)dst) & 15) == 0), "fill needs a 16-byte aligned buffer of 4k words");
  if (n_words == 0) return PCS_OK;
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  PCS_LAUNCH("k_fill128", (cudaStream_t)stream, k_fill128<<<sms * 16, 256, 0, (cudaStream_t)stream>>>((uint4*)dst, value, n_words / 4));
  return pcs_check_launch("fill");
}

int pcs_max_label(const void* labels, int label_bytes, int64_t n, int64_t* out, void* stream) {
  PCS_REQUIRE(labels && out && n >= 0, "bad arguments");
  PCS_REQUIRE(label_bytes == 4 || label_bytes == 8, "label dtype must be int32 or int64");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(out, 0, 8, st);
  if (n == 0) return PCS_OK;
  const unsigned g = pcs_blocks(n, 256 * 8);
  if (label_bytes == 4)
    PCS_LAUNCH("k_max_label", st, k_max_label<int32_t><<<g, 256, 0, st>>>((const int32_t*)labels, n, (long long*)out));
  else
    PCS_LAUNCH("k_max_label", st, k_max_label<long long><<<g, 256, 0, st>>>((const long long*)labels, n, (long long*)out));
  return pcs_check_launch("max label");
}

int pcs_transpose(const void* in, void* out, int elem_bytes, int H, int W, void* stream) {
  PCS_REQUIRE(in && out && in != out && H >= 1 && W >= 1, "bad arguments");
  PCS_REQUIRE(elem_bytes == 1 || elem_bytes == 4, "transpose handles 1- and 4-byte elements");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 g((W + 31) / 32, (H + 31) / 32);
  if (elem_bytes == 1)
    PCS_LAUNCH("k_transpose", st, k_transpose<uint8_t><<<g, 256, 0, st>>>((const uint8_t*)in, (uint8_t*)out, H, W));
  else
    PCS_LAUNCH("k_transpose", st, k_transpose<uint32_t><<<g, 256, 0, st>>>((const uint32_t*)in, (uint32_t*)out, H, W));
  return pcs_check_launch("transpose");
}

}  // extern "C"
