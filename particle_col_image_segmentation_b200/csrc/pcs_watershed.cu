// Marker-controlled watershed.
//
// Replaces skimage.segmentation.watershed(boundary_map, markers, mask=binary_mask)
//   refine_boundaries.py:73   (connectivity 1, compactness 0, no watershed line)
//
// scikit-image floods sequentially: pop the queued pixel of lowest (value, age), give its unlabelled
// in-mask 4-neighbours its label, queue them.  That order is not a parallel algorithm, but its RESULT
// has a closed form when no two competing pixels share a value.  Let b(p) be the bottleneck cost of
// p: the smallest L such that p is reached from some marker pixel through in-mask pixels whose values
// (the marker's and p's included) are all <= L.  Pixels are popped in order of b (everything with
// b < L is popped before the pixel of value L that opens the next basin), and pixels with equal b
// hang off the same bottleneck pixel, hence carry the same label.  A pixel is labelled when its first
// neighbour pops, so
//     label(p) = label(q*),   q* = the 4-neighbour with the smallest b,   b(p) = max(value(p), b(q*)).
// That is a fixed point of a local rule and is computed here by relaxation: every free pixel keeps
// (b, d, label) = (max(v, b_q*), d_q* + 1, label_q*) with q* = argmin over neighbours of (b, d, label);
// d (hops to the marker) only breaks cycles inside zones of equal b.  CTAs iterate their 32x16 tile in
// shared memory against a fixed halo (Jacobi, two buffers), global sweeps ping-pong two state arrays
// until a whole sweep changes nothing.  At the fixed point b is the exact bottleneck cost and every
// parent chain ends in a marker, so the labels equal the sequential flood's bit for bit on tie-free
// images (tests/test_gpu_watershed.py); where equal values compete the result is still a valid flood,
// with ties going to the smaller (hops, label) instead of scikit-image's insertion age.
#include "pcs_common.cuh"

#include "pcs.h"

#define WS_TX 32
#define WS_TY 16
#define WS_INNER (2 * (WS_TX + WS_TY))
#define WS_DINF 0x3fffffff
#define WS_BATCH 8  // sweeps between two reads of the convergence flags

struct WsState {
  double b;
  int d, lab;
};

__device__ __forceinline__ bool ws_less(const WsState& a, const WsState& c) {
  return a.b < c.b || (a.b == c.b && (a.d < c.d || (a.d == c.d && a.lab < c.lab)));
}

// kind: 0 outside the mask (or the image), 1 marker, 2 free
__global__ void __launch_bounds__(WS_TX* WS_TY)
    k_ws_init(const double* __restrict__ img, const int32_t* __restrict__ markers, const uint32_t* __restrict__ mask_bits,
              double* __restrict__ b, int* __restrict__ d, int* __restrict__ lab, uint8_t* __restrict__ kind, int B, int H, int W,
              int WW) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * H * W) return;
  int x, y;
  long long s;
  pcs_split3(i, W, H, x, y, s);
  const bool in = mask_bits ? ((mask_bits[(s * H + y) * (long long)WW + (x >> 5)] >> (x & 31)) & 1u) : true;
  const int m = in ? markers[i] : 0;  // markers outside the mask do not flood
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  kind[i] = !in ? 0 : (m > 0 ? 1 : 2);
  b[i] = m > 0 ? img[i] : inf;
  d[i] = m > 0 ? 0 : WS_DINF;
  lab[i] = m > 0 ? m : 0;
}

__global__ void __launch_bounds__(WS_TX* WS_TY)
    k_ws_sweep(const double* __restrict__ img, const uint8_t* __restrict__ kind, const double* __restrict__ b_in,
               const int* __restrict__ d_in, const int* __restrict__ l_in, double* __restrict__ b_out, int* __restrict__ d_out,
               int* __restrict__ l_out, int* __restrict__ changed, const uint8_t* __restrict__ tf_in, uint8_t* __restrict__ tf_out, int H,
               int W) {
  __shared__ double sb[2][WS_TY + 2][WS_TX + 2];
  __shared__ int sd[2][WS_TY + 2][WS_TX + 2], sl[2][WS_TY + 2][WS_TX + 2];
  const int tx = threadIdx.x % WS_TX, ty = threadIdx.x / WS_TX;
  const int x0 = blockIdx.x * WS_TX, y0 = blockIdx.y * WS_TY;
  const long long base = (long long)blockIdx.z * H * W;
  // Active tiles only.  A tile's result depends on its own pixels and the one-pixel halo its four neighbours lend it;
  // if none of the five changed in the previous sweep, this sweep would write what the output buffer (the input of
  // the previous sweep) already holds.  Late sweeps touch only the tiles along the flood paths that are still
  // growing -- on the 4096^2 touching-particle map 123 sweeps used to stream the whole state 123 times.
  const long long tile = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  if (tf_in) {
    int act = tf_in[tile];
    if (blockIdx.x > 0) act |= tf_in[tile - 1];
    if (blockIdx.x + 1 < gridDim.x) act |= tf_in[tile + 1];
    if (blockIdx.y > 0) act |= tf_in[tile - gridDim.x];
    if (blockIdx.y + 1 < gridDim.y) act |= tf_in[tile + gridDim.x];
    if (!act) {
      if (threadIdx.x == 0) tf_out[tile] = 0;
      return;
    }
  }
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  // tile + one-pixel halo; outside the image (and outside the mask: b = inf there) nothing floods
  for (int i = threadIdx.x; i < (WS_TY + 2) * (WS_TX + 2); i += WS_TX * WS_TY) {
    const int r = i / (WS_TX + 2), c = i % (WS_TX + 2);
    const int y = y0 + r - 1, x = x0 + c - 1;
    WsState s{inf, WS_DINF, 0};
    if (y >= 0 && y < H && x >= 0 && x < W) {
      const long long g = base + (long long)y * W + x;
      s.b = b_in[g];
      s.d = d_in[g];
      s.lab = l_in[g];
    }
    sb[0][r][c] = sb[1][r][c] = s.b;
    sd[0][r][c] = sd[1][r][c] = s.d;
    sl[0][r][c] = sl[1][r][c] = s.lab;
  }
  const int y = y0 + ty, x = x0 + tx;
  const bool inside = y < H && x < W;
  const long long g = base + (long long)y * W + x;
  const bool free_px = inside && kind[g] == 2;
  const double v = free_px ? img[g] : 0.0;
  __syncthreads();
  const WsState first{sb[0][ty + 1][tx + 1], sd[0][ty + 1][tx + 1], sl[0][ty + 1][tx + 1]};
  int cur = 0;
  for (int it = 0; it < WS_INNER; ++it) {
    int moved = 0;
    if (free_px) {
      const int r = ty + 1, c = tx + 1;
      WsState best{sb[cur][r - 1][c], sd[cur][r - 1][c], sl[cur][r - 1][c]};
      const WsState q1{sb[cur][r][c - 1], sd[cur][r][c - 1], sl[cur][r][c - 1]};
      const WsState q2{sb[cur][r][c + 1], sd[cur][r][c + 1], sl[cur][r][c + 1]};
      const WsState q3{sb[cur][r + 1][c], sd[cur][r + 1][c], sl[cur][r + 1][c]};
      if (ws_less(q1, best)) best = q1;
      if (ws_less(q2, best)) best = q2;
      if (ws_less(q3, best)) best = q3;
      WsState n{inf, WS_DINF, 0};
      if (best.b < inf) {
        n.b = best.b > v ? best.b : v;
        n.d = best.d + 1;
        n.lab = best.lab;
      }
      moved = n.b != sb[cur][r][c] || n.d != sd[cur][r][c] || n.lab != sl[cur][r][c];
      sb[cur ^ 1][r][c] = n.b;
      sd[cur ^ 1][r][c] = n.d;
      sl[cur ^ 1][r][c] = n.lab;
    }
    cur ^= 1;
    if (!__syncthreads_or(moved)) break;
  }
  int diff = 0;
  if (inside) {
    const WsState last{sb[cur][ty + 1][tx + 1], sd[cur][ty + 1][tx + 1], sl[cur][ty + 1][tx + 1]};
    b_out[g] = last.b;
    d_out[g] = last.d;
    l_out[g] = last.lab;
    diff = last.b != first.b || last.d != first.d || last.lab != first.lab;
  }
  diff = __syncthreads_or(diff);
  if (threadIdx.x == 0) {
    tf_out[tile] = (uint8_t)(diff != 0);
    if (diff) *changed = 1;
  }
}

extern "C" {

size_t pcs_watershed_workspace_bytes(int B, int H, int W) {
  const size_t n = (size_t)B * H * W;
  const size_t tiles = (size_t)B * ((H + WS_TY - 1) / WS_TY) * ((W + WS_TX - 1) / WS_TX);
  return 2 * pcs_align256(n * 8) + 4 * pcs_align256(n * 4) + pcs_align256(n) + 256 + 2 * pcs_align256(tiles);  // 256 B: WS_BATCH convergence flags; then two planes of per-tile "changed" flags
}

// Blocking: the sweep loop reads the convergence flags back once per batch of WS_BATCH sweeps.  Returns PCS_OK and the number
// of sweeps (optional); labels receives the last state's labels (int32, 0 = not flooded).
int pcs_watershed_f64(const double* image, const int32_t* markers, const uint32_t* mask_bits, int32_t* labels, int B, int H, int W,
                      int max_sweeps, int* sweeps_out, void* wsp, size_t ws_bytes, void* stream) {
  PCS_REQUIRE(B >= 1 && H >= 1 && W >= 1, "empty batch or image");
  PCS_REQUIRE(image && markers && labels, "null argument");
  PCS_REQUIRE(B <= 65535 && (H + WS_TY - 1) / WS_TY <= 65535, "image too large for the watershed grid");
  if (wsp == nullptr || ws_bytes < pcs_watershed_workspace_bytes(B, H, W)) {
    pcs_set_error("watershed workspace too small (see pcs_watershed_workspace_bytes)");
    return PCS_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)B * H * W;
  char* p = (char*)wsp;
  double* b[2];
  int *d[2], *l[2];
  for (int i = 0; i < 2; ++i) {
    b[i] = (double*)p;
    p += pcs_align256(n * 8);
  }
  for (int i = 0; i < 2; ++i) {
    d[i] = (int*)p;
    p += pcs_align256(n * 4);
    l[i] = (int*)p;
    p += pcs_align256(n * 4);
  }
  uint8_t* kind = (uint8_t*)p;
  p += pcs_align256(n);
  int* changed = (int*)p;
  p += 256;
  const size_t tiles = (size_t)B * ((H + WS_TY - 1) / WS_TY) * ((W + WS_TX - 1) / WS_TX);
  uint8_t* tf[2] = {(uint8_t*)p, (uint8_t*)p + pcs_align256(tiles)};
  const int WW = pcs_words(W);
  PCS_LAUNCH("k_ws_init", st, k_ws_init<<<pcs_blocks((long long)n, WS_TX * WS_TY), WS_TX * WS_TY, 0, st>>>(
      image, markers, mask_bits, b[0], d[0], l[0], kind, B, H, W, WW));
  dim3 grid((W + WS_TX - 1) / WS_TX, (H + WS_TY - 1) / WS_TY, B);
  if (max_sweeps <= 0) {  // a flood path crosses at most about one tile per sweep; winding paths are bounded by the pixel count
    const long long bound = (long long)H * W / 16 + 4LL * (H + W) + 64;
    max_sweeps = bound > 0x7fffffffLL ? 0x7fffffff : (int)bound;
  }
  // Sweeps are issued in batches of WS_BATCH with one flag per sweep and ONE read-back per batch: the host no longer
  // sits between every two sweeps.  A sweep that changes nothing leaves the state at the fixed point, so the sweeps a
  // batch runs past convergence are no-ops and the reported count is the index of the first quiet sweep.
  int cur = 0, sweeps = 0, launched = 0;
  bool done = false;
  while (!done) {
    if (sweeps >= max_sweeps) {
      pcs_set_error("watershed did not converge within max_sweeps");
      return PCS_ERR_INVALID;
    }
    int flags[WS_BATCH];
    cudaMemsetAsync(changed, 0, sizeof(int) * WS_BATCH, st);
    for (int i = 0; i < WS_BATCH; ++i) {
      // the very first sweep runs every tile (and fills BOTH state buffers' worth of flags: its output flags feed sweep 2)
      PCS_LAUNCH("k_ws_sweep", st, k_ws_sweep<<<grid, WS_TX * WS_TY, 0, st>>>(image, kind, b[cur], d[cur], l[cur], b[cur ^ 1], d[cur ^ 1],
                                                                                l[cur ^ 1], changed + i, launched ? tf[cur] : nullptr, tf[cur ^ 1], H, W));
      cur ^= 1;
      ++launched;
    }
    cudaMemcpyAsync(flags, changed, sizeof(int) * WS_BATCH, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) return pcs_check_launch("watershed sweep");
    for (int i = 0; i < WS_BATCH; ++i) {
      ++sweeps;
      if (!flags[i]) {
        done = true;
        break;
      }
    }
  }
  cudaMemcpyAsync(labels, l[cur], n * 4, cudaMemcpyDeviceToDevice, st);
  if (sweeps_out) *sweeps_out = sweeps;
  return pcs_check_launch("watershed");
}

}  // extern "C"
