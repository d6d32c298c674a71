// K7 exact Euclidean distance transform (separable, integer squared distances,
// one correctly-rounded fp64 sqrt at the end) and the squared-distance threshold
// that implements disk dilation / erosion.
//
// Replaces (file:line in /root/reference):
//   scipy.ndimage.distance_transform_edt(~particle_mask)   tiff_analysis.py:996
//   scipy.ndimage.distance_transform_edt(binary_mask)      refine_boundaries.py:60
//   binary_dilation(mask, disk(r))  ==  EDT(~mask)^2 <= r^2 tiff_analysis.py:828, :990
//   dist_transform < DISTANCE_THRESHOLD                    tiff_analysis.py:1000
//
// scipy computes an int32 feature transform and then sqrt(float64(dy^2+dx^2)); an
// exact integer squared distance followed by IEEE sqrt is therefore bit-identical.
//
// Pass 1 (columns): g(y,x) = distance to the nearest background pixel in column x.
// Pass 2 (rows): D2(y,x) = min_c (x-c)^2 + g(y,c)^2.  The arg-min is monotone in x
// (the cost matrix is totally monotone), so the row is solved by divide and
// conquer: the middle column first, then each half with the candidate range cut
// at the parent's arg-min -- O(W log W) integer evaluations per row, no divisions,
// no stacks, level-synchronous inside one CTA with shuffle reductions.  The
// column distance also bounds the search window (|x-c| < g(y,x)), which makes
// rows through small particles nearly free.
#include "pcs_common.cuh"

#include "pcs.h"

#define EDT_INF 0xFFFFu
#define EDT_ROW_THREADS 256
#define EDT_KEY_MAX 0xFFFFFFFFFFFFFFFFull

// thread per column
__global__ void __launch_bounds__(128)
    k_edt_cols(const uint32_t* __restrict__ bits, int invert, uint16_t* __restrict__ g, int H, int W, int WW) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  long long b = blockIdx.y;
  const uint32_t* col = bits + b * (long long)H * WW + (x >> 5);
  uint16_t* gc = g + b * (long long)H * W + x;
  const int sh = x & 31;
  const uint32_t flip = invert ? 1u : 0u;
  uint32_t d = EDT_INF;
#pragma unroll 8
  for (int y = 0; y < H; ++y) {
    uint32_t fg = ((__ldg(col + (long long)y * WW) >> sh) & 1u) ^ flip;
    d = fg ? (d == EDT_INF ? EDT_INF : d + 1u) : 0u;
    gc[(long long)y * W] = (uint16_t)d;
  }
  d = EDT_INF;
#pragma unroll 8
  for (int y = H - 1; y >= 0; --y) {
    uint32_t fg = ((__ldg(col + (long long)y * WW) >> sh) & 1u) ^ flip;
    d = fg ? (d == EDT_INF ? EDT_INF : d + 1u) : 0u;
    uint32_t up = gc[(long long)y * W];
    gc[(long long)y * W] = (uint16_t)min(up, d);
  }
}

// CTA per row
__global__ void __launch_bounds__(EDT_ROW_THREADS)
    k_edt_rows(const uint16_t* __restrict__ g, double* __restrict__ dist, int32_t* __restrict__ sq,
               uint32_t* __restrict__ thr_bits, int thr_sq, int H, int W, int WW, int W2, int L) {
  extern __shared__ uint16_t smem[];
  uint16_t* gs = smem;      // column distances of this row
  uint16_t* am = smem + W;  // arg-min column per x
  const int tid = threadIdx.x;
  const int y = blockIdx.x;
  const long long b = blockIdx.y;
  const uint16_t* grow = g + (b * H + y) * (long long)W;
  for (int x = tid; x < W; x += EDT_ROW_THREADS) gs[x] = grow[x];
  __syncthreads();

  for (int l = 0; l < L; ++l) {
    const int step = W2 >> (l + 1);
    const int nodes = 1 << l;
    int gsz = EDT_ROW_THREADS >> l;
    gsz = gsz > 32 ? 32 : (gsz < 1 ? 1 : gsz);
    const int groups = EDT_ROW_THREADS / gsz;
    const int sub = tid & (gsz - 1);
    for (int j = tid / gsz; j < nodes; j += groups) {
      const int xp = step * (2 * j + 1);  // 1-based position
      const bool valid = xp <= W;
      const int x = xp - 1;
      unsigned long long key = EDT_KEY_MAX;
      if (valid) {
        const uint32_t gx = gs[x];
        if (gx == 0u) {
          key = (unsigned long long)x;
        } else {
          int clo = (xp - step >= 1) ? (int)am[xp - step - 1] : 0;
          int chi = (xp + step <= W) ? (int)am[xp + step - 1] : W - 1;
          if (clo == (int)EDT_INF) clo = 0;  // neighbours without a finite site give no bound
          if (chi == (int)EDT_INF) chi = W - 1;
          if (gx != EDT_INF) {
            clo = max(clo, x - (int)gx + 1);
            chi = min(chi, x + (int)gx - 1);
          }
          for (int c = clo + sub; c <= chi; c += gsz) {
            const uint32_t gv = gs[c];
            if (gv != EDT_INF) {
              const int d = x - c;
              const uint32_t val = (uint32_t)(d * d) + gv * gv;
              const unsigned long long k2 = ((unsigned long long)val << 16) | (unsigned)c;
              key = k2 < key ? k2 : key;
            }
          }
        }
      }
      for (int o = gsz >> 1; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other < key ? other : key;
      }
      if (valid && sub == 0) am[x] = key == EDT_KEY_MAX ? (uint16_t)EDT_INF : (uint16_t)(key & 0xffffu);
    }
    __syncthreads();
  }

  const long long orow = (b * H + y) * (long long)W;
  const int Wp = WW << 5;
  for (int x0 = 0; x0 < Wp; x0 += EDT_ROW_THREADS) {
    const int x = x0 + tid;
    long long d2 = 0;
    bool nosite = false;
    if (x < W) {
      const uint32_t a = am[x];
      if (a == EDT_INF) {
        // no background pixel anywhere: scipy measures to the virtual point (-1, 0);
        // for the threshold output (dilation of an empty mask) the distance is infinite
        nosite = true;
        d2 = (long long)(y + 1) * (y + 1) + (long long)x * x;
      } else {
        const int d = x - (int)a;
        const uint32_t gv = gs[a];
        d2 = (long long)d * d + (long long)gv * gv;
      }
      if (dist) dist[orow + x] = sqrt((double)d2);
      if (sq) sq[orow + x] = (int32_t)d2;
    }
    if (thr_bits) {
      unsigned ball = __ballot_sync(0xffffffffu, x < W && !nosite && d2 <= (long long)thr_sq);
      if ((tid & 31) == 0 && (x >> 5) < WW) thr_bits[(b * H + y) * (long long)WW + (x >> 5)] = ball;
    }
  }
}

extern "C" {

size_t pcs_edt_workspace_bytes(int B, int H, int W) { return pcs_align256((size_t)B * H * W * 2); }

int pcs_edt_bits(const uint32_t* bits, int invert, int B, int H, int W, double* dist, int32_t* sq, uint32_t* thr_bits,
                 int thr_sq, void* ws, size_t ws_bytes, void* stream) {
  PCS_REQUIRE(B >= 1 && H >= 1 && W >= 1, "empty batch or image");
  PCS_REQUIRE(H <= 16384 && W <= 16384, "image side above 16384 is not supported");
  PCS_REQUIRE(dist || sq || thr_bits, "no output requested");
  if (ws == nullptr || ws_bytes < pcs_edt_workspace_bytes(B, H, W)) {
    pcs_set_error("EDT workspace too small (see pcs_edt_workspace_bytes)");
    return PCS_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int WW = pcs_words(W);
  uint16_t* g = (uint16_t*)ws;
  dim3 gc((W + 127) / 128, B);
  PCS_LAUNCH("k_edt_cols", st, k_edt_cols<<<gc, 128, 0, st>>>(bits, invert, g, H, W, WW));
  int W2 = 1, L = 0;
  while (W2 <= W) {
    W2 <<= 1;
    ++L;
  }
  size_t smem = (size_t)W * 4;
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    cudaFuncSetAttribute(k_edt_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    smem_set = smem;
  }
  dim3 gr(H, B);
  PCS_LAUNCH("k_edt_rows", st, k_edt_rows<<<gr, EDT_ROW_THREADS, smem, st>>>(g, dist, sq, thr_bits, thr_sq, H, W, WW, W2, L));
  return pcs_check_launch("edt");
}

}  // extern "C"
