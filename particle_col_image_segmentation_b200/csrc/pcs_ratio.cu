// NanoSIMS ratio images: the display pipeline at the top of
// HCN_nanosims_rois_activity_distance_5iso_YG.m.
//
// Replaces (file:line in /root/reference):
//   imgaussfilt(plane, sigma)                                        .m:43-44, :51-52, :56-58, :62
//   ratio = num ./ (d0 + d1 + d2), max(ratio(:))                      .m:45, :53-54, :59-60, :63-69
//   uint8(x .* (255 / max))   (round half away from zero, saturate)   .m:31-37 and every ratio image
//
// imgaussfilt: taps 2*ceil(2*sigma)+1, weights exp(-x^2 / (2 sigma^2)) normalised to sum 1, 'replicate'
// border, applied down the columns and then along the rows in double precision.  Every product and sum
// is a separate IEEE operation in tap order (no FMA), so oracle/nanosims.py reproduces it bit for bit;
// MATLAB's own summation order is not documented (parity with MATLAB itself: unpinned).
#include "pcs_common.cuh"

#include "pcs.h"

#define GAUSS_MAX_R 32

struct GaussTaps {
  double w[2 * GAUSS_MAX_R + 1];
  int r;
};

// thread per pixel; dir 0: down the columns (taps over y), dir 1: along the rows (taps over x)
__global__ void __launch_bounds__(256)
    k_gauss_pass(const double* __restrict__ in, double* __restrict__ out, GaussTaps taps, int dir, int B, int H, int W) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * H * W;
  if (i >= total) return;
  int x, y;
  long long b;
  pcs_split3(i, W, H, x, y, b);
  const double* src = in + b * (long long)H * W;
  double acc = 0.0;
  for (int t = -taps.r; t <= taps.r; ++t) {
    const int yy = dir == 0 ? min(max(y + t, 0), H - 1) : y;
    const int xx = dir == 1 ? min(max(x + t, 0), W - 1) : x;
    acc = __dadd_rn(acc, __dmul_rn(taps.w[t + taps.r], src[(long long)yy * W + xx]));
  }
  out[i] = acc;
}

__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
  unsigned long long* p = (unsigned long long*)addr;
  unsigned long long old = *p;
  while (true) {
    const double cur = __longlong_as_double((long long)old);
    if (!(v > cur)) return;  // also leaves when v is NaN
    const unsigned long long seen = atomicCAS(p, old, (unsigned long long)__double_as_longlong(v));
    if (seen == old) return;
    old = seen;
  }
}

__global__ void k_fill_f64(double* p, double v) { *p = v; }

// ratio = num ./ ((d0 + d1) + d2), denominators in the order given; maxv = max over the non-NaN ratios
__global__ void __launch_bounds__(256)
    k_ratio(const double* __restrict__ num, const double* __restrict__ d0, const double* __restrict__ d1,
            const double* __restrict__ d2, double* __restrict__ ratio, double* __restrict__ maxv, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double r = __longlong_as_double(0xfff0000000000000LL);  // -inf
  if (i < n) {
    r = num[i];
    if (d0) {
      double den = d0[i];
      if (d1) den = __dadd_rn(den, d1[i]);
      if (d2) den = __dadd_rn(den, d2[i]);
      r = __ddiv_rn(r, den);
    }
    ratio[i] = r;
  }
  // warp maximum first (NaN never wins), one atomic per warp
  double m = (r == r) ? r : __longlong_as_double(0xfff0000000000000LL);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double other = __shfl_xor_sync(0xffffffffu, m, o);
    m = other > m ? other : m;
  }
  if ((threadIdx.x & 31) == 0) atomic_max_double(maxv, m);
}

// uint8(x .* (255 / max)): MATLAB rounds half away from zero and saturates; NaN -> 0
__global__ void __launch_bounds__(256)
    k_scale_u8(const double* __restrict__ x, const double* __restrict__ maxv, uint8_t* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double s = __ddiv_rn(255.0, *maxv);
  const double v = __dmul_rn(x[i], s);
  uint8_t o = 0;
  if (v == v) {
    const double r = round(v);
    o = r <= 0.0 ? 0 : (r >= 255.0 ? 255 : (uint8_t)r);
  }
  out[i] = o;
}

// ---------------------------------------------------------------- imresize (.m:125, :189)
// MATLAB resizes one dimension at a time: out[i] = sum_p w[i][p] * in[idx[i][p]] with, per output index, P taps
// whose weights and (mirrored) source indices come from imresize's `contributions` (host side, nanosims.py).  The
// same kernel applies the ADJOINT when it is handed the transposed tap table, which is how the per-ROI sums under
// a resized ROI mask are computed with one resize per ion plane instead of one per ROI (nanosims.py).  Products and
// sums are separate IEEE operations in tap order (no FMA): oracle/nanosims.py reproduces the forward resize bit for
// bit.  Thread per output element; `main` is the resized dimension, `other` the one carried along.
__global__ void __launch_bounds__(256)
    k_resize_taps(const double* __restrict__ in, double* __restrict__ out, const int* __restrict__ idx, const double* __restrict__ wts,
                  int P, long long n_main, long long n_other, long long is_main, long long is_other, long long os_main, long long os_other) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_main * n_other) return;
  const long long i = t / n_other, x = t - i * n_other;
  double acc = 0.0;
  for (int p = 0; p < P; ++p) {
    const int j = idx[i * P + p];
    if (j < 0) continue;  // padding of a ragged (transposed) table
    acc = __dadd_rn(acc, __dmul_rn(wts[i * P + p], in[(long long)j * is_main + x * is_other]));
  }
  out[i * os_main + x * os_other] = acc;
}

extern "C" {

int pcs_gauss_f64(const double* in, double* out, double* tmp, double sigma, int B, int H, int W, void* stream) {
  PCS_REQUIRE(B >= 1 && H >= 1 && W >= 1, "empty batch or image");
  PCS_REQUIRE(in && out && tmp, "null argument");
  PCS_REQUIRE(sigma > 0.0, "sigma must be positive");
  const int r = (int)ceil(2.0 * sigma);
  PCS_REQUIRE(r <= GAUSS_MAX_R, "sigma too large for the spatial filter");
  GaussTaps taps;
  taps.r = r;
  double sum = 0.0;
  for (int t = -r; t <= r; ++t) {
    taps.w[t + r] = exp(-(double)(t * t) / (2.0 * sigma * sigma));
    sum += taps.w[t + r];
  }
  for (int t = 0; t <= 2 * r; ++t) taps.w[t] /= sum;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned g = pcs_blocks((long long)B * H * W, 256);
  PCS_LAUNCH("k_gauss_pass", st, k_gauss_pass<<<g, 256, 0, st>>>(in, tmp, taps, 0, B, H, W));
  PCS_LAUNCH("k_gauss_pass", st, k_gauss_pass<<<g, 256, 0, st>>>(tmp, out, taps, 1, B, H, W));
  return pcs_check_launch("gauss");
}

int pcs_ratio_f64(const double* num, const double* d0, const double* d1, const double* d2, double* ratio, double* maxv, int64_t n,
                  void* stream) {
  PCS_REQUIRE(n >= 1 && num && ratio && maxv, "null argument");
  PCS_REQUIRE(d0 || (!d1 && !d2), "denominators must be given in order");
  cudaStream_t st = (cudaStream_t)stream;
  k_fill_f64<<<1, 1, 0, st>>>(maxv, -INFINITY);
  PCS_LAUNCH("k_ratio", st, k_ratio<<<pcs_blocks(n, 256), 256, 0, st>>>(num, d0, d1, d2, ratio, maxv, n));
  return pcs_check_launch("ratio");
}

int pcs_scale_u8_f64(const double* x, const double* maxv, uint8_t* out, int64_t n, void* stream) {
  PCS_REQUIRE(n >= 1 && x && maxv && out, "null argument");
  PCS_LAUNCH("k_scale_u8", (cudaStream_t)stream, k_scale_u8<<<pcs_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(x, maxv, out, n));
  return pcs_check_launch("scale u8");
}

int pcs_resize_taps_f64(const double* in, double* out, const int32_t* idx, const double* wts, int P, int64_t n_main, int64_t n_other,
                        int64_t in_stride_main, int64_t in_stride_other, int64_t out_stride_main, int64_t out_stride_other, void* stream) {
  PCS_REQUIRE(in && out && idx && wts && P >= 1 && n_main >= 1 && n_other >= 1, "bad resize arguments");
  PCS_REQUIRE(in != out, "resize cannot run in place");
  PCS_LAUNCH("k_resize_taps", (cudaStream_t)stream, k_resize_taps<<<pcs_blocks(n_main * n_other, 256), 256, 0, (cudaStream_t)stream>>>(
      in, out, idx, wts, P, n_main, n_other, in_stride_main, in_stride_other, out_stride_main, out_stride_other));
  return pcs_check_launch("resize");
}

}  // extern "C"
