// K8 per-label reductions, K10 label-keyed selection, K11 per-ROI plane sums and
// pairwise minimum distances.
//
// Replaces (file:line in /root/reference):
//   skimage.measure.regionprops + .area/.centroid/.bbox/.coords[0]   tiff_analysis.py:263-275, :746-773
//   per-cell overlap loop (labeled == label, logical_and, np.sum)     tiff_analysis.py:268-279
//   merged_image |= (dilated_labels == v)                             tiff_analysis.py:878
//   sum(sum(plane .* roimask)) per ROI and plane                      .m:122-132, :186-196
//   pdist2 + min (nearest neighbour, distance to boundary pixels)     .m:260-263, :301-304
//
// Reductions are run based: one thread walks a column of 32-pixel words down a
// strip of rows, so a component that crosses the strip is accumulated in
// registers and reaches the table with a handful of 64-bit atomics per strip
// instead of one set per pixel or per run.  Sums are integers (exact); means and
// centroids are one fp64 division on the host, which reproduces numpy's
// float64 results bit for bit.
#include "pcs_common.cuh"

#include "pcs.h"

#define PROPS_THREADS 128
#define PROPS_ROWS 8
#define T_AREA 0
#define T_SUMY 1
#define T_SUMX 2
#define T_MINY 3
#define T_MINX 4
#define T_MAXY 5
#define T_MAXX 6
#define T_FIRST 7
#define T_SUMI 8
#define T_OVERLAP 9

__global__ void k_table_init(long long* __restrict__ table, long long cap, const int* __restrict__ offsets, int B) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap || (offsets && i >= (long long)offsets[B])) return;  // only the rows in use when the count is known
#pragma unroll
  for (int c = 0; c < PCS_TABLE_COLS; ++c) {
    long long v = 0;
    if (c == T_MINY || c == T_MINX || c == T_FIRST) v = 0x7fffffffffffffffLL;
    if (c == T_MAXY || c == T_MAXX) v = -1;
    table[c * cap + i] = v;
  }
}

struct RegionAcc {
  int label;
  int area, minx, maxx, miny, maxy;
  long long sx, sy, si, first;
  int ov;
};

__device__ __noinline__ void acc_flush(const RegionAcc& a, long long* __restrict__ table, long long cap, long long base,
                                       bool has_int, bool has_ov) {
  if (a.label <= 0) return;
  long long row = base + a.label - 1;
  if (row >= cap) return;
  typedef unsigned long long ull;
  atomicAdd((ull*)(table + T_AREA * cap + row), (ull)a.area);
  atomicAdd((ull*)(table + T_SUMY * cap + row), (ull)a.sy);
  atomicAdd((ull*)(table + T_SUMX * cap + row), (ull)a.sx);
  atomicMin(table + T_MINY * cap + row, (long long)a.miny);
  atomicMin(table + T_MINX * cap + row, (long long)a.minx);
  atomicMax(table + T_MAXY * cap + row, (long long)a.maxy);
  atomicMax(table + T_MAXX * cap + row, (long long)a.maxx);
  atomicMin(table + T_FIRST * cap + row, a.first);
  if (has_int) atomicAdd((ull*)(table + T_SUMI * cap + row), (ull)a.si);
  if (has_ov && a.ov) atomicAdd((ull*)(table + T_OVERLAP * cap + row), (ull)a.ov);
}

// thread = (slice, strip of PROPS_ROWS rows, word column).  Runs come from the
// label image itself: a run is a maximal stretch of equal non-zero labels inside
// the word, so the same kernel serves binary and multi-valued labelling.  Words
// without foreground (fg_bits) are skipped without touching the label image.
template <typename LabT, typename IntT>
__global__ void __launch_bounds__(PROPS_THREADS)
    k_region_table(const LabT* __restrict__ labels, const IntT* __restrict__ intensity, const uint32_t* __restrict__ fg_bits,
                   const uint32_t* __restrict__ ov_bits, const int* __restrict__ offsets, long long* __restrict__ table,
                   long long cap, int B, int H, int W, int WW, int strips) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * strips * WW;
  if (t >= total) return;
  int k, strip;
  long long b;
  pcs_split3(t, WW, strips, k, strip, b);
  const long long base = offsets ? (long long)offsets[b] : 0;
  const int x0 = k << 5;
  const int n = min(32, W - x0);
  const bool has_int = intensity != nullptr, has_ov = ov_bits != nullptr;
  RegionAcc acc;
  acc.label = 0;
  const int y0 = strip * PROPS_ROWS;
  // all foreground words of the strip first: one round trip instead of PROPS_ROWS dependent ones
  uint32_t fw[PROPS_ROWS];
#pragma unroll
  for (int r = 0; r < PROPS_ROWS; ++r)
    fw[r] = (y0 + r < H) ? (fg_bits ? __ldg(fg_bits + (b * H + y0 + r) * (long long)WW + k) : 0xffffffffu) : 0u;
#pragma unroll 1
  for (int r = 0; r < PROPS_ROWS; ++r) {
    if (fw[r] == 0u) continue;
    const int y = y0 + r;
    const long long wi = (b * H + y) * (long long)WW + k;
    const LabT* lrow = labels + (b * H + y) * (long long)W + x0;
    int L[32];
    if (sizeof(LabT) == 4 && n == 32 && ((((uintptr_t)lrow) & 15) == 0)) {
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        int4 q = __ldg(reinterpret_cast<const int4*>(lrow) + v);
        L[4 * v] = q.x;
        L[4 * v + 1] = q.y;
        L[4 * v + 2] = q.z;
        L[4 * v + 3] = q.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) L[i] = i < n ? (int)lrow[i] : 0;
    }
    int I[32];
    if (has_int) {
      const IntT* irow = intensity + (b * H + y) * (long long)W + x0;
      if (n == 32 && ((((uintptr_t)irow) & 15) == 0)) {
        constexpr int VEC = 16 / sizeof(IntT);
#pragma unroll
        for (int v = 0; v < 32 / VEC; ++v) {
          uint4 q = __ldg(reinterpret_cast<const uint4*>(irow) + v);
          const IntT* e = reinterpret_cast<const IntT*>(&q);
#pragma unroll
          for (int i = 0; i < VEC; ++i) I[v * VEC + i] = L[v * VEC + i] != 0 ? (int)e[i] : 0;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) I[i] = (i < n && L[i] != 0) ? (int)irow[i] : 0;
      }
    }
    const uint32_t ovw = has_ov ? ov_bits[wi] : 0u;
    int runlab = 0, runstart = 0;
    long long runsi = 0;
#pragma unroll
    for (int i = 0; i <= 32; ++i) {
      const int lab = i < 32 ? L[i] : 0;
      if (lab != runlab) {
        if (runlab != 0) {
          const int len = i - runstart;
          const int xs = x0 + runstart;
          if (runlab != acc.label) {
            acc_flush(acc, table, cap, base, has_int, has_ov);
            acc.label = runlab;
            acc.area = 0;
            acc.sx = acc.sy = acc.si = 0;
            acc.minx = xs;
            acc.maxx = xs + len - 1;
            acc.miny = acc.maxy = y;
            acc.first = (long long)y * W + xs;  // rows then columns ascend inside a thread
            acc.ov = 0;
          }
          acc.area += len;
          acc.sx += (long long)len * xs + (long long)(len * (len - 1) / 2);
          acc.sy += (long long)len * y;
          acc.si += runsi;
          acc.minx = min(acc.minx, xs);
          acc.maxx = max(acc.maxx, xs + len - 1);
          acc.maxy = y;
          if (has_ov) {
            const uint32_t rm = (len >= 32 ? 0xffffffffu : ((1u << len) - 1u)) << runstart;
            acc.ov += __popc(ovw & rm);
          }
        }
        runlab = lab;
        runstart = i;
        runsi = 0;
      }
      if (has_int && i < 32) runsi += I[i];
    }
  }
  acc_flush(acc, table, cap, base, has_int, has_ov);
}

// Binary-labelling fast path: the runs are read off the foreground bit image, so the label
// image is touched once per RUN (4 bytes) instead of once per pixel, background words cost one
// 4-byte load, and the register footprint stays small enough for full occupancy.
template <typename IntT>
__global__ void __launch_bounds__(PROPS_THREADS)
    k_region_table_bits(const int32_t* __restrict__ labels, const IntT* __restrict__ intensity, const uint32_t* __restrict__ fg_bits,
                        const uint32_t* __restrict__ ov_bits, const int* __restrict__ offsets, long long* __restrict__ table,
                        long long cap, int B, int H, int W, int WW, int strips) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * strips * WW;
  if (t >= total) return;
  int k, strip;
  long long b;
  pcs_split3(t, WW, strips, k, strip, b);
  const long long base = offsets ? (long long)offsets[b] : 0;
  const int x0 = k << 5;
  const bool has_int = intensity != nullptr, has_ov = ov_bits != nullptr;
  // rows start 16-byte aligned and the word is complete: vector loads of the intensities
  const bool vec = has_int && x0 + 32 <= W && ((W * sizeof(IntT)) & 15) == 0 && (((uintptr_t)intensity) & 15) == 0;
  const int y0 = strip * PROPS_ROWS;
  RegionAcc acc;
  acc.label = 0;
  const uint32_t* fcol = fg_bits + (b * H + y0) * (long long)WW + k;
  uint32_t fnext = __ldg(fcol);  // y0 < H always; the next row's word is requested one iteration ahead
#pragma unroll 1
  for (int r = 0; r < PROPS_ROWS; ++r) {
    const uint32_t f = fnext;
    fnext = (r + 1 < PROPS_ROWS && y0 + r + 1 < H) ? __ldg(fcol + (long long)(r + 1) * WW) : 0u;
    if (f == 0u) continue;
    const int y = y0 + r;
    const long long rowo = (b * H + y) * (long long)W + x0;
    const uint32_t ovw = has_ov ? ov_bits[(b * H + y) * (long long)WW + k] : 0u;
    // the 32 intensities of the word in one go (16-byte loads) when the row allows it
    constexpr int NW = 8 * (int)sizeof(IntT);  // 32-bit registers holding the word's pixels
    uint32_t iw[NW];
    if (has_int && vec) {
#pragma unroll
      for (int v = 0; v < NW / 4; ++v) {
        const uint4 q4 = __ldg(reinterpret_cast<const uint4*>(intensity + rowo) + v);
        iw[4 * v] = q4.x;
        iw[4 * v + 1] = q4.y;
        iw[4 * v + 2] = q4.z;
        iw[4 * v + 3] = q4.w;
      }
    }
    uint32_t S = f & ~(f << 1);
    while (S) {
      const int s = __ffs(S) - 1;
      S &= S - 1;
      const uint32_t upper = ~(f >> s);
      const int len = upper ? (__ffs(upper) - 1) : 32;
      const int lab = labels[rowo + s];
      const int xs = x0 + s;
      long long si = 0;
      if (has_int) {
        if (vec) {
          const uint32_t rm = (len >= 32 ? 0xffffffffu : ((1u << len) - 1u)) << s;
          uint32_t sum = 0;  // 32 pixels of at most 16 bits: no overflow
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const uint32_t px = sizeof(IntT) == 2 ? ((i & 1) ? (iw[i >> 1] >> 16) : (iw[i >> 1] & 0xffffu))
                                                  : ((iw[i >> 2] >> (8 * (i & 3))) & 0xffu);
            sum += (rm >> i) & 1u ? px : 0u;
          }
          si = sum;
        } else {
          for (int i = 0; i < len; ++i) si += (long long)intensity[rowo + s + i];
        }
      }
      if (lab != acc.label) {
        acc_flush(acc, table, cap, base, has_int, has_ov);
        acc.label = lab;
        acc.area = 0;
        acc.sx = acc.sy = acc.si = 0;
        acc.minx = xs;
        acc.maxx = xs + len - 1;
        acc.miny = acc.maxy = y;
        acc.first = (long long)y * W + xs;
        acc.ov = 0;
      }
      acc.area += len;
      acc.sx += (long long)len * xs + (long long)(len * (len - 1) / 2);
      acc.sy += (long long)len * y;
      acc.si += si;
      acc.minx = min(acc.minx, xs);
      acc.maxx = max(acc.maxx, xs + len - 1);
      acc.maxy = y;
      if (has_ov) {
        const uint32_t rm = (len >= 32 ? 0xffffffffu : ((1u << len) - 1u)) << s;
        acc.ov += __popc(ovw & rm);
      }
    }
  }
  acc_flush(acc, table, cap, base, has_int, has_ov);
}

// thread per table row: the float64 table the callers consume (same columns and arithmetic as
// oracle/pipeline.py::region_table -- exact integer sums, one IEEE division for centroid and mean):
//   z, label, area, centroid_y, centroid_x, min_row, min_col, max_row + 1, max_col + 1,
//   first_row, first_col, intensity_sum, intensity_mean
__global__ void __launch_bounds__(256)
    k_table_finalize(const long long* __restrict__ table, long long cap, const int* __restrict__ offsets, int B, int W, double z0,
                     double* __restrict__ out, long long out_cap, double* __restrict__ count_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && count_out) *count_out = (double)offsets[B];  // the true row count travels with the rows (it may exceed out_cap)
  const long long n = min(min((long long)offsets[B], cap), out_cap);
  if (i >= n) return;
  int lo = 0, hi = B;  // slice of this row: offsets[lo] <= i < offsets[lo + 1]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((long long)offsets[mid] <= i) lo = mid; else hi = mid;
  }
  const double area = (double)table[T_AREA * cap + i];
  const long long first = table[T_FIRST * cap + i];
  const double si = (double)table[T_SUMI * cap + i];
  double* o = out + i * 13;
  o[0] = z0 + (double)lo;
  o[1] = (double)(i - offsets[lo] + 1);
  o[2] = area;
  o[3] = __ddiv_rn((double)table[T_SUMY * cap + i], area);
  o[4] = __ddiv_rn((double)table[T_SUMX * cap + i], area);
  o[5] = (double)table[T_MINY * cap + i];
  o[6] = (double)table[T_MINX * cap + i];
  o[7] = (double)(table[T_MAXY * cap + i] + 1);
  o[8] = (double)(table[T_MAXX * cap + i] + 1);
  o[9] = (double)(first / W);
  o[10] = (double)(first % W);
  o[11] = si;
  o[12] = __ddiv_rn(si, area);
}

// out bits = pixels whose label has keep[label] != 0 (per-slice LUT rows of `lut_stride` entries)
template <typename LabT>
__global__ void __launch_bounds__(256)
    k_select_labels(const LabT* __restrict__ labels, const uint8_t* __restrict__ keep, long long lut_stride,
                    uint32_t* __restrict__ out, int B, int H, int W, int WW) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * H * WW;
  if (t >= total) return;
  int k, y_;
  long long b;
  pcs_split3(t, WW, H, k, y_, b);
  const long long rowi = b * H + y_;
  const LabT* lrow = labels + rowi * (long long)W + (k << 5);
  const uint8_t* kp = keep + b * lut_stride;
  const int n = min(32, W - (k << 5));
  uint32_t o = 0;
  int last = 0, lastk = 0;
  for (int i = 0; i < n; ++i) {
    int lab = (int)lrow[i];
    if (lab != last) {
      last = lab;
      lastk = lab > 0 && lab < lut_stride ? kp[lab] : 0;
    }
    o |= (uint32_t)(lastk != 0) << i;
  }
  out[t] = o;
}

// out bits = pixels whose component has area >= min_size, read straight from the region
// table (small-object filter, tiff_analysis.py:769-773 expressed on the mask)
__global__ void __launch_bounds__(256)
    k_select_by_area(const int32_t* __restrict__ labels, const uint32_t* __restrict__ fg_bits, const long long* __restrict__ table,
                     long long cap, const int* __restrict__ offsets, long long min_size, uint32_t* __restrict__ out, int B, int H,
                     int W, int WW) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * H * WW;
  if (t >= total) return;
  uint32_t f = fg_bits[t];
  if (!f) {
    out[t] = 0u;
    return;
  }
  int k, y_;
  long long b;
  pcs_split3(t, WW, H, k, y_, b);
  const long long rowi = b * H + y_;
  const int32_t* lrow = labels + rowi * (long long)W + (k << 5);
  const long long base = offsets[b];
  uint32_t o = 0, S = f & ~(f << 1);
  while (S) {
    int s = __ffs(S) - 1;
    S &= S - 1;
    uint32_t upper = ~(f >> s);
    int len = upper ? (__ffs(upper) - 1) : 32;
    uint32_t rm = (len >= 32 ? 0xffffffffu : ((1u << len) - 1u)) << s;
    long long row = base + lrow[s] - 1;
    if (row >= 0 && row < cap && table[T_AREA * cap + row] >= min_size) o |= rm;
  }
  out[t] = o;
}

// per-ROI sums of K float64 planes (labels int32, 0 = outside any ROI)
__global__ void __launch_bounds__(256)
    k_roi_sums(const int32_t* __restrict__ labels, const double* __restrict__ planes, int K, long long npix, int n_rois,
               double* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  int lab = labels[i];
  if (lab <= 0 || lab > n_rois) return;
  for (int k = 0; k < K; ++k) atomicAdd(out + (long long)(lab - 1) * K + k, planes[(long long)k * npix + i]);
}

// out[i] = min_j sqrt((ax-bx)^2 + (ay-by)^2), numpy evaluation order, no FMA contraction
__global__ void __launch_bounds__(128)
    k_min_dist(const double* __restrict__ a, long long na, const double* __restrict__ b, long long nb, double* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= na) return;
  const double ax = a[2 * i], ay = a[2 * i + 1];
  double best = __longlong_as_double(0x7ff0000000000000LL);
  for (long long j = 0; j < nb; ++j) {
    double dx = __dsub_rn(ax, b[2 * j]), dy = __dsub_rn(ay, b[2 * j + 1]);
    double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    best = d2 < best ? d2 : best;
  }
  out[i] = sqrt(best);
}

// nearest neighbour of every a_i among the b_j (optionally skipping j == i: nearest OTHER cell of the same
// strain, refine_boundaries.py:8-12 goal 3; the MATLAB model is pdist2 + min, .m:260-263).  b is staged
// through shared memory 128 points at a time; squared distances in numpy order without FMA, first minimum
// wins (np.argmin), one IEEE sqrt at the end.
__global__ void __launch_bounds__(128)
    k_nearest(const double* __restrict__ a, long long na, const double* __restrict__ b, long long nb, int exclude_self,
              double* __restrict__ out_d, long long* __restrict__ out_j) {
  __shared__ double sb[128][2];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const double ax = i < na ? a[2 * i] : 0.0, ay = i < na ? a[2 * i + 1] : 0.0;
  double best = __longlong_as_double(0x7ff0000000000000LL);
  long long at = -1;
  for (long long j0 = 0; j0 < nb; j0 += 128) {
    const long long j = j0 + threadIdx.x;
    __syncthreads();
    if (j < nb) {
      sb[threadIdx.x][0] = b[2 * j];
      sb[threadIdx.x][1] = b[2 * j + 1];
    }
    __syncthreads();
    const int n = (int)min(128LL, nb - j0);
    for (int t = 0; t < n; ++t) {
      const double dx = __dsub_rn(ax, sb[t][0]), dy = __dsub_rn(ay, sb[t][1]);
      const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
      if (d2 < best && !(exclude_self && j0 + t == i)) {
        best = d2;
        at = j0 + t;
      }
    }
  }
  if (i < na) {
    out_d[i] = sqrt(best);
    if (out_j) out_j[i] = at;
  }
}

extern "C" {

int pcs_table_init(int64_t* table, int64_t cap, void* stream) {
  PCS_REQUIRE(cap >= 1 && table != nullptr, "empty table");
  PCS_LAUNCH("k_table_init", (cudaStream_t)stream, k_table_init<<<pcs_blocks(cap, 256), 256, 0, (cudaStream_t)stream>>>((long long*)table, cap, nullptr, 0));
  return pcs_check_launch("table init");
}

int pcs_table_init_rows(int64_t* table, int64_t cap, const int32_t* offsets, int B, void* stream) {
  PCS_REQUIRE(cap >= 1 && table != nullptr && offsets != nullptr && B >= 1, "empty table");
  PCS_LAUNCH("k_table_init", (cudaStream_t)stream, k_table_init<<<pcs_blocks(cap, 256), 256, 0, (cudaStream_t)stream>>>((long long*)table, cap, offsets, B));
  return pcs_check_launch("table init");
}

// label_bytes 4|8; intensity_dtype: -1 none, 0 u8, 1 u16
int pcs_region_table(const void* labels, int label_bytes, const void* intensity, int intensity_dtype, const uint32_t* fg_bits,
                     const uint32_t* ov_bits,
                     const int32_t* offsets, int64_t* table, int64_t cap, int B, int H, int W, void* stream) {
  PCS_REQUIRE(B >= 1 && H >= 1 && W >= 1, "empty batch or image");
  PCS_REQUIRE(label_bytes == 4 || label_bytes == 8, "label dtype must be int32 or int64");
  PCS_REQUIRE(table != nullptr && cap >= 1, "empty table");
  const int WW = pcs_words(W);
  const int strips = (H + PROPS_ROWS - 1) / PROPS_ROWS;
  unsigned g = pcs_blocks((long long)B * strips * WW, PROPS_THREADS);
  cudaStream_t st = (cudaStream_t)stream;
  long long* tb = (long long*)table;
  if (intensity == nullptr) intensity_dtype = -1;
#define LAUNCH(LT, IT) \
  PCS_LAUNCH("k_region_table", st, (k_region_table<LT, IT><<<g, PROPS_THREADS, 0, st>>>((const LT*)labels, (const IT*)intensity, fg_bits, ov_bits, offsets, tb, cap, B, H, W, WW, strips)))
  if (label_bytes == 4 && fg_bits != nullptr) {
    // binary labelling: runs come from the bit image
    if (intensity_dtype == 1)
      PCS_LAUNCH("k_region_table_bits", st, (k_region_table_bits<uint16_t><<<g, PROPS_THREADS, 0, st>>>((const int32_t*)labels, (const uint16_t*)intensity, fg_bits, ov_bits, offsets, tb, cap, B, H, W, WW, strips)));
    else if (intensity_dtype == 0 || intensity_dtype == -1)
      PCS_LAUNCH("k_region_table_bits", st, (k_region_table_bits<uint8_t><<<g, PROPS_THREADS, 0, st>>>((const int32_t*)labels, (const uint8_t*)intensity, fg_bits, ov_bits, offsets, tb, cap, B, H, W, WW, strips)));
    else {
      pcs_set_error("unsupported intensity dtype");
      return PCS_ERR_UNSUPPORTED;
    }
  } else if (label_bytes == 4) {
    if (intensity_dtype == 1)
      LAUNCH(int32_t, uint16_t);
    else if (intensity_dtype == 0 || intensity_dtype == -1)
      LAUNCH(int32_t, uint8_t);
    else {
      pcs_set_error("unsupported intensity dtype");
      return PCS_ERR_UNSUPPORTED;
    }
  } else {
    if (intensity_dtype == 1)
      LAUNCH(long long, uint16_t);
    else if (intensity_dtype == 0 || intensity_dtype == -1)
      LAUNCH(long long, uint8_t);
    else {
      pcs_set_error("unsupported intensity dtype");
      return PCS_ERR_UNSUPPORTED;
    }
  }
#undef LAUNCH
  return pcs_check_launch("region table");
}

int pcs_table_finalize(const int64_t* table, int64_t cap, const int32_t* offsets, int B, int W, double z0, double* out, void* stream) {
  return pcs_table_finalize_ex(table, cap, offsets, B, W, z0, out, cap, nullptr, stream);
}

int pcs_table_finalize_ex(const int64_t* table, int64_t cap, const int32_t* offsets, int B, int W, double z0, double* out, int64_t out_cap,
                          double* count_out, void* stream) {
  PCS_REQUIRE(table && offsets && out && cap >= 1 && out_cap >= 1 && B >= 1 && W >= 1, "bad table arguments");
  const long long rows = out_cap < cap ? out_cap : cap;
  PCS_LAUNCH("k_table_finalize", (cudaStream_t)stream, k_table_finalize<<<pcs_blocks(rows, 256), 256, 0, (cudaStream_t)stream>>>(
      (const long long*)table, cap, offsets, B, W, z0, out, out_cap, count_out));
  return pcs_check_launch("table finalize");
}

int pcs_select_labels(const void* labels, int label_bytes, const uint8_t* keep, int64_t lut_stride, uint32_t* out, int B, int H,
                      int W, void* stream) {
  PCS_REQUIRE(B >= 1 && H >= 1 && W >= 1, "empty batch or image");
  PCS_REQUIRE(label_bytes == 4 || label_bytes == 8, "label dtype must be int32 or int64");
  const int WW = pcs_words(W);
  unsigned g = pcs_blocks((long long)B * H * WW, 256);
  if (label_bytes == 4)
    PCS_LAUNCH("k_select_labels", (cudaStream_t)stream, k_select_labels<int32_t><<<g, 256, 0, (cudaStream_t)stream>>>((const int32_t*)labels, keep, lut_stride, out, B, H, W, WW));
  else
    PCS_LAUNCH("k_select_labels", (cudaStream_t)stream, k_select_labels<long long><<<g, 256, 0, (cudaStream_t)stream>>>((const long long*)labels, keep, lut_stride, out, B, H, W, WW));
  return pcs_check_launch("select labels");
}

int pcs_select_by_area(const int32_t* labels, const uint32_t* fg_bits, const int64_t* table, int64_t cap, const int32_t* offsets,
                       int64_t min_size, uint32_t* out, int B, int H, int W, void* stream) {
  PCS_REQUIRE(B >= 1 && H >= 1 && W >= 1, "empty batch or image");
  PCS_REQUIRE(labels && fg_bits && table && offsets && out, "null argument");
  const int WW = pcs_words(W);
  PCS_LAUNCH("k_select_by_area", (cudaStream_t)stream, k_select_by_area<<<pcs_blocks((long long)B * H * WW, 256), 256, 0, (cudaStream_t)stream>>>(
      labels, fg_bits, (const long long*)table, cap, offsets, min_size, out, B, H, W, WW));
  return pcs_check_launch("select by area");
}

int pcs_roi_sums_f64(const int32_t* labels, const double* planes, int K, int64_t npix, int n_rois, double* out, void* stream) {
  PCS_REQUIRE(K >= 1 && npix >= 1 && n_rois >= 0, "empty input");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_rois == 0) return PCS_OK;
  cudaMemsetAsync(out, 0, (size_t)n_rois * K * 8, st);
  PCS_LAUNCH("k_roi_sums", st, k_roi_sums<<<pcs_blocks(npix, 256), 256, 0, st>>>(labels, planes, K, npix, n_rois, out));
  return pcs_check_launch("roi sums");
}

int pcs_nearest_f64(const double* a, int64_t na, const double* b, int64_t nb, int exclude_self, double* out_dist, int64_t* out_index,
                    void* stream) {
  if (na <= 0) return PCS_OK;
  PCS_REQUIRE(a && out_dist && (b || nb == 0) && nb >= 0, "null argument");
  PCS_LAUNCH("k_nearest", (cudaStream_t)stream, k_nearest<<<pcs_blocks(na, 128), 128, 0, (cudaStream_t)stream>>>(
      a, na, b, nb, exclude_self, out_dist, (long long*)out_index));
  return pcs_check_launch("nearest");
}

int pcs_min_dist_f64(const double* a, int64_t na, const double* b, int64_t nb, double* out, void* stream) {
  if (na <= 0) return PCS_OK;
  PCS_LAUNCH("k_min_dist", (cudaStream_t)stream, k_min_dist<<<pcs_blocks(na, 128), 128, 0, (cudaStream_t)stream>>>(a, na, b, nb, out));
  return pcs_check_launch("min dist");
}

}  // extern "C"
