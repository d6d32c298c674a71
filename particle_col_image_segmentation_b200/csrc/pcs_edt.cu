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
// Column pass without a column image.  The mask is transposed 32x32 bits at a time
// into VERTICAL words (k_edt_transpose: bit j of vw[band][x] is row 32*band + j of
// column x) and a 64-step scan per column records, for every band, the distance to
// the nearest background row above / below the band (k_edt_carry).  The vertical
// distance g(y, x) of any pixel is then two bit scans (clz / ffs) of one word plus a
// carry -- 0.25 B/pixel of traffic instead of a 2 B/pixel distance image written,
// re-read and rewritten.
//
// Row pass: D2(y, x) = min_c (x-c)^2 + g(y, c)^2.
//   * k_edt_near (CTA per 32-row band x 256-column tile): a pixel with g <= EDT_DMAX only
//     needs candidates with |x-c| < g, so an outward search c = x -+ d that stops as soon
//     as d^2 >= best is exact, stays inside the tile + 40-pixel halo held in shared memory
//     and costs O(distance) per foreground pixel; pixels with a larger g flag their row;
//   * k_edt_far (warp per flagged row): the arg-min is monotone in x (the cost matrix is totally
//     monotone), so the row is solved by divide and conquer -- the middle column first,
//     then each half with the candidate range cut at the parent's arg-min: O(W log W)
//     integer evaluations, no divisions, no stacks, shuffle reductions inside the warp.
#include "pcs_common.cuh"

#include "pcs.h"

#define EDT_INF 0xFFFFu
#define EDT_DMAX 40u
#define EDT_KEY_MAX 0xFFFFFFFFFFFFFFFFull
#define EDT_WARPS 8

// warp per (slice, band, group of 4 words): 32 rows x 128 columns of bits -> 128 vertical words.  Four
// independent loads per lane are in flight before the shuffles start (one load per warp left the kernel
// waiting on memory latency).
__global__ void __launch_bounds__(256)
    k_edt_transpose(const uint32_t* __restrict__ bits, int invert, uint32_t* __restrict__ vw, int B, int H, int W, int WW,
                    int NB) {
  long long g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int WG = (WW + 3) >> 2;
  long long total = (long long)B * NB * WG;
  if (g >= total) return;
  int kg, q;
  long long b;
  pcs_split3(g, WG, NB, kg, q, b);
  const int y = (q << 5) + lane;
  const int k0 = kg << 2;
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    w[i] = 0xffffffffu;  // rows past the image never act as background
    if (y < H && k0 + i < WW) {
      w[i] = __ldg(bits + (b * H + y) * (long long)WW + k0 + i);
      if (invert) w[i] = ~w[i];
    }
  }
  // 32x32 bit transposes across the warp: five block-swap steps (lane = row in, lane = column out)
#pragma unroll
  for (int j = 16; j >= 1; j >>= 1) {
    const uint32_t m = j == 16 ? 0x0000ffffu : j == 8 ? 0x00ff00ffu : j == 4 ? 0x0f0f0f0fu : j == 2 ? 0x33333333u : 0x55555555u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t other = __shfl_xor_sync(0xffffffffu, w[i], j);
      w[i] = (lane & j) ? ((w[i] & ~m) | ((other & ~m) >> j)) : ((w[i] & m) | ((other & m) << j));
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (k0 + i < WW) vw[(b * NB + q) * (long long)(WW << 5) + ((k0 + i) << 5) + lane] = w[i];
}

// thread per (column, direction): distance from the band edge to the nearest background row beyond it.
// The scan is sequential over the bands of a column, but its loads are independent of the carry, so they
// are issued sixteen at a time; the upward and the downward scan run in different threads.
__global__ void __launch_bounds__(128)
    k_edt_carry(const uint32_t* __restrict__ vw, uint16_t* __restrict__ up, uint16_t* __restrict__ dn, int W, int Wp, int NB) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  const long long b = blockIdx.y;
  const bool down = blockIdx.z != 0;
  const uint32_t* v = vw + b * (long long)NB * Wp + x;
  uint16_t* o = (down ? dn : up) + b * (long long)NB * Wp + x;
  uint32_t carry = EDT_INF;
  for (int q0 = 0; q0 < NB; q0 += 16) {
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int q = down ? NB - 1 - (q0 + i) : q0 + i;
      z[i] = (q0 + i < NB) ? ~__ldg(v + (long long)q * Wp) : 0u;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (q0 + i < NB) {
        const int q = down ? NB - 1 - (q0 + i) : q0 + i;
        o[(long long)q * Wp] = (uint16_t)carry;  // up: from row 32q - 1 upwards; down: from row 32(q+1) downwards
        const uint32_t edge = down ? (uint32_t)(__ffs(z[i]) - 1) : (uint32_t)__clz(z[i]);  // rows between the band edge and its nearest zero
        carry = z[i] ? edge : (carry == EDT_INF ? EDT_INF : carry + 32u);
      }
    }
  }
}

// divide-and-conquer row solve by one warp; gs = column distances, am = arg-min per x
__device__ __forceinline__ void edt_row_dc(const uint16_t* gs, uint16_t* am, int W, int W2, int L, int lane) {
  for (int l = 0; l < L; ++l) {
    const int step = W2 >> (l + 1);
    const int nodes = 1 << l;
    const int gsz = l < 5 ? (32 >> l) : 1;
    const int groups = 32 / gsz;
    const int sub = lane & (gsz - 1);
    for (int j = lane / gsz; j < nodes; j += groups) {
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
    __syncwarp();
  }
}

// vertical distance of row r of a band from the column word z (zero bits = background rows)
__device__ __forceinline__ uint32_t edt_vdist(uint32_t z, int r, uint32_t cu, uint32_t cd) {
  if ((z >> r) & 1u) return 0u;
  const uint32_t zu = z & (0xffffffffu >> (31 - r));  // background rows at or above r
  const uint32_t gu = zu ? (uint32_t)(r - (31 - __clz(zu))) : (cu == EDT_INF ? EDT_INF : cu + (uint32_t)r + 1u);
  const uint32_t zd = z >> r;  // background rows at or below r
  const uint32_t gd = zd ? (uint32_t)(__ffs(zd) - 1) : (cd == EDT_INF ? EDT_INF : cd + 32u - (uint32_t)r);
  return min(gu, gd);
}

// sqrt of the squared distances the near phase can produce (<= EDT_DMAX^2): one L1-resident load
// instead of a 45-instruction fp64 sqrt per foreground pixel.  The table is a compile-time constant
// (exact hexadecimal literals of the correctly rounded square roots, identical to what the far phase's
// sqrt() returns), so no initialisation kernel exists that a first call on another stream, inside a
// graph capture or from another thread could race with.
#define EDT_LUT_N (EDT_DMAX * EDT_DMAX + 1)
__device__ const double g_edt_sqrt_lut[EDT_LUT_N] = {
#include "pcs_edt_sqrt_lut.inc"
};

// Phase 1: CTA per (column tile, band, slice).  Every pixel whose vertical distance is at most
// EDT_DMAX is final after an outward search inside the tile + halo; the others flag their row.
// Only foreground pixels are visited individually (a balanced work list built off the vertical bit
// words).  The float64 tile is cleared first -- full tiles by the SM's copy engine (cp.async.bulk from a
// zeroed piece of shared memory), ragged ones by 256-bit thread stores -- and the foreground pixels are
// then overwritten one by one; both land in L2 before the lines leave for DRAM: 8 B/pixel of output traffic.
#define EDT_TW 256
#define EDT_HALO 40
#ifndef EDT_TMA_FILL
#define EDT_TMA_FILL 1  // zero fill of the float64 tile by the SM's copy engine (cp.async.bulk shared -> global, SASS UBLKCP)
#endif
__device__ __forceinline__ void edt_bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}
#define EDT_TWH (EDT_TW + 2 * EDT_HALO)
#ifndef EDT_ER
#define EDT_ER 16        // rows per CTA: half a band keeps 8 CTAs (2048 threads) resident per SM
#endif
#define EDT_GCLAMP 255u  // vertical distances above EDT_DMAX never win a near search: 8 bits are enough
__global__ void __launch_bounds__(EDT_TW)
    k_edt_near(const uint32_t* __restrict__ vw, const uint16_t* __restrict__ up, const uint16_t* __restrict__ dn,
               double* __restrict__ dist, int32_t* __restrict__ sq, uint32_t* __restrict__ thr_bits, int thr_sq,
               uint8_t* __restrict__ row_far, int H, int W, int WW, int NB) {
  __shared__ __align__(16) uint8_t g[EDT_ER][EDT_TWH];
  __shared__ __align__(16) uint16_t d2s[EDT_ER][EDT_TW];
  __shared__ uint32_t tb[EDT_ER][EDT_TW / 32];
  __shared__ unsigned short items[EDT_ER * EDT_TWH];  // (row << 9 | column) of the foreground pixels of tile + halo
  __shared__ uint32_t zcol[EDT_TWH];                   // background rows of every column (vertical word, inverted)
  __shared__ unsigned short cucol[EDT_TWH], cdcol[EDT_TWH];  // carries beyond the band, per column
  __shared__ int nitems;
  const int tid = threadIdx.x;
  if (tid == 0) nitems = 0;
  const int Wp = WW << 5;
  const int x0 = blockIdx.x * EDT_TW;
  const int q = blockIdx.y / (32 / EDT_ER);                 // band of 32 rows (one vertical word per column)
  const int r0 = (blockIdx.y % (32 / EDT_ER)) * EDT_ER;     // first row of the band this CTA owns
  const long long b = blockIdx.z;
  const long long band = (b * NB + q) * (long long)Wp;
  const int rows = min(EDT_ER, H - (q << 5) - r0);
  if (rows <= 0) return;
  const int cols = min(EDT_TW, W - x0);
  // column words and carries of both column slots of this thread are requested first, so one
  // global round trip is in flight while the tiles are cleared (background: g = 0, distance 0)
  const int xa = x0 - EDT_HALO + tid, xb = xa + EDT_TW;
  const bool ina = xa >= 0 && xa < W, inb = tid < 2 * EDT_HALO && xb < W;  // xb >= 0 always
  uint32_t fa = 0, fb = 0, cua = 0, cda = 0, cub = 0, cdb = 0;
  if (ina) {
    fa = __ldg(vw + band + xa);  // set bits = foreground rows of this column
    cua = __ldg(up + band + xa);
    cda = __ldg(dn + band + xa);
  }
  if (inb) {
    fb = __ldg(vw + band + xb);
    cub = __ldg(up + band + xb);
    cdb = __ldg(dn + band + xb);
  }
  // The float64 output: the whole tile is cleared here with 256-bit stores (background pixels are at
  // distance 0, and they are the majority); pass 2 then overwrites the foreground pixels one by one.  Both
  // sets of stores come from this CTA with barriers in between, so they reach memory in order and merge in
  // L2: DRAM still sees every line once.  Converting a shared tile of squared distances on the way out
  // instead cost 30 % of the kernel's instructions (four table lookups per store, predicated for all).
  const long long obase = (b * H + (q << 5) + r0) * (long long)W + x0;
  // full tiles: the zeros leave through the copy engine (one bulk copy of a 2 KB row segment per row, issued by one
  // thread from a zeroed piece of shared memory).  The per-thread 256-bit stores this replaces filled the load /
  // store queues the shared-memory traffic of the passes below goes through: stores and compute added up
  // (155 + 47 us per 32 slices) instead of overlapping.
  const bool bulk = EDT_TMA_FILL && dist && cols == EDT_TW && (W & 1) == 0 && ((((uintptr_t)dist) & 15) == 0);
  if (bulk) {
    uint4* dz = reinterpret_cast<uint4*>(&d2s[0][0]);  // 8 KB, zero from here on unless sq is asked for (pass 2)
    for (int i = tid; i < EDT_ER * EDT_TW / 8; i += EDT_TW) dz[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy zeros before the copy engine reads them
  } else if (dist) {
    if (cols == EDT_TW && (W & 3) == 0 && ((((uintptr_t)dist) & 31) == 0)) {
      // thread = (row tid / 64 of every group of four rows, four columns): one pointer, a constant stride
      double* dst = dist + obase + (long long)(tid >> 6) * W + ((tid & 63) << 2);
      const long long step = 4ll * W;
#pragma unroll
      for (int r = tid >> 6; r < EDT_ER; r += EDT_TW / 64, dst += step)
        if (r < rows) asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(dst), "d"(0.0) : "memory");
    } else {
      for (int i = tid; i < rows * EDT_TW; i += EDT_TW) {
        const int r = i / EDT_TW, c = i % EDT_TW;
        if (c < cols) dist[obase + (long long)r * W + c] = 0.0;
      }
    }
  }
  // this CTA's rows of the two column words; a tile whose own columns hold no foreground (the halo only lends its g
  // to pixels of the tile) is done once its zeros are on their way -- most tiles of a blob image: no shared tiles
  // to clear, no lists, no further barriers
  const uint32_t rmask = (rows < 32 ? (1u << rows) - 1u : 0xffffffffu) & ((EDT_ER == 32) ? 0xffffffffu : ((1u << EDT_ER) - 1u));
  const uint32_t fra = (fa >> r0) & rmask, frb = (fb >> r0) & rmask;  // zero where the column is outside the image
  const uint32_t own = tid >= EDT_HALO ? fra : frb;  // slot 0 holds columns -HALO .. TW-HALO-1 of the tile, slot 1 the rest
  const int any = __syncthreads_or(own != 0u);
  if (bulk && tid == 0) {
    for (int r = 0; r < rows; ++r) edt_bulk_store(dist + obase + (long long)r * W, &d2s[0][0], EDT_TW * 8);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  if (!any && !sq && !thr_bits) {
    // shared memory must outlive the copy engine's reads of it
    if (bulk && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    return;
  }
  {
    // g = 0 for the columns of the image, EDT_GCLAMP (no site) for tile or halo columns outside it
    uint4* gz = reinterpret_cast<uint4*>(&g[0][0]);
    for (int i = tid; i < EDT_ER * EDT_TWH / 16; i += EDT_TW) {
      const int xl = x0 - EDT_HALO + (i % (EDT_TWH / 16)) * 16;  // image column of the first of these 16 bytes
      uint4 z = make_uint4(0, 0, 0, 0);
      if (xl < 0 || xl + 15 >= W) {
        uint32_t wv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          wv[k] = 0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int x = xl + 4 * k + j;
            if (x < 0 || x >= W) wv[k] |= EDT_GCLAMP << (8 * j);
          }
        }
        z = make_uint4(wv[0], wv[1], wv[2], wv[3]);
      }
      gz[i] = z;
    }
    if (sq && !bulk) {
      uint4* dz = reinterpret_cast<uint4*>(&d2s[0][0]);
      for (int i = tid; i < EDT_ER * EDT_TW / 8; i += EDT_TW) dz[i] = make_uint4(0, 0, 0, 0);
    }
  }
  // pass 1a, thread per column: list the foreground pixels of tile + halo (warp-level exclusive scan of the
  // per-column counts) and park the column's word and carries in shared memory
#pragma unroll
  for (int slot = 0; slot < 2; ++slot) {
    if (slot == 1 && tid >= 2 * EDT_HALO) break;
    const int col = tid + slot * EDT_TW;
    const uint32_t fw = slot ? fb : fa;
    const uint32_t f = slot ? frb : fra;
    if (slot ? inb : ina) {
      zcol[col] = ~fw;
      cucol[col] = (unsigned short)(slot ? cub : cua);
      cdcol[col] = (unsigned short)(slot ? cdb : cda);
    }
    const int cnt = __popc(f);
    const unsigned act = __activemask();
    if (__ballot_sync(act, cnt != 0) == 0u) continue;  // none of this warp's columns holds foreground (warp-uniform)
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int tv = __shfl_up_sync(act, incl, o);
      if ((tid & 31) >= o) incl += tv;
    }
    const int last = 31 - __clz(act);
    int wtot = __shfl_sync(act, incl, last);
    int wbase = 0;
    if ((tid & 31) == last) wbase = atomicAdd(&nitems, wtot);
    wbase = __shfl_sync(act, wbase, last);
    if (cnt) {
      int pos = wbase + incl - cnt;
      uint32_t ff = f;
      while (ff) {
        const int r = __ffs(ff) - 1;
        ff &= ff - 1;
        items[pos++] = (unsigned short)((r << 9) | col);
      }
    }
  }
  __syncthreads();
  // pass 1b, thread per listed pixel: its vertical distance.  Done per pixel, not per column, because the
  // foreground comes in blobs: a thread per column left most warps idle behind the few that cross a blob.
  const int n = nitems;
  for (int it = tid; it < n; it += EDT_TW) {
    const int item = items[it];
    const int r = item >> 9, c = item & 511;
    g[r][c] = (uint8_t)min(edt_vdist(zcol[c], r0 + r, cucol[c], cdcol[c]), EDT_GCLAMP);
  }
  // pass 2 overwrites foreground pixels of the tile (and d2s when sq is asked for): the zeros must have landed first
  if (bulk && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncthreads();
  if (thr_bits) {
    // background pixels are at distance 0: start every row word from them (warp w owns word w)
    const int x = x0 + tid;
    for (int r = 0; r < EDT_ER; ++r) {
      unsigned bg = __ballot_sync(0xffffffffu, x < W && g[r][tid + EDT_HALO] == 0u);
      if ((tid & 31) == 0) tb[r][tid >> 5] = thr_sq >= 0 ? bg : 0u;
    }
    __syncthreads();
  }
  // pass 2: thread per foreground pixel of the tile proper (halo pixels only lend their g)
  for (int it = tid; it < n; it += EDT_TW) {
    const int item = items[it];
    const int r = item >> 9, c = item & 511;
    if (c < EDT_HALO || c >= EDT_HALO + EDT_TW) continue;
    const uint32_t gx = g[r][c];
    if (gx > EDT_DMAX) {
      row_far[b * H + (q << 5) + r0 + r] = 1;  // solved by k_edt_far, which rewrites the whole row
      continue;
    }
    uint32_t best = gx * gx;
    for (uint32_t dd = 1; dd * dd < best; ++dd) {  // dd < gx <= EDT_HALO: stays inside the tile + halo
      const uint32_t d2d = dd * dd;
      const uint32_t g1 = g[r][c - (int)dd], g2 = g[r][c + (int)dd];
      best = min(best, min(d2d + g1 * g1, d2d + g2 * g2));  // clamped columns (255^2) never win
    }
    if (dist) dist[obase + (long long)r * W + (c - EDT_HALO)] = g_edt_sqrt_lut[best];
    if (sq) d2s[r][c - EDT_HALO] = (uint16_t)best;
    if (thr_bits && (int)best <= thr_sq) atomicOr(&tb[r][(c - EDT_HALO) >> 5], 1u << ((c - EDT_HALO) & 31));
  }
  __syncthreads();
  // the squared-distance output (when asked for) leaves through the shared tile in one coalesced sweep
  if (sq) {
    if (cols == EDT_TW && (W & 3) == 0 && ((((uintptr_t)sq) & 15) == 0)) {
      for (int i = tid; i < rows * (EDT_TW / 4); i += EDT_TW) {
        const int r = i / (EDT_TW / 4), c = (i % (EDT_TW / 4)) * 4;
        const uint2 pr = *reinterpret_cast<const uint2*>(&d2s[r][c]);
        *reinterpret_cast<int4*>(sq + obase + (long long)r * W + c) =
            make_int4((int)(pr.x & 0xffffu), (int)(pr.x >> 16), (int)(pr.y & 0xffffu), (int)(pr.y >> 16));
      }
    } else {
      for (int i = tid; i < rows * EDT_TW; i += EDT_TW) {
        const int r = i / EDT_TW, c = i % EDT_TW;
        if (c < cols) sq[obase + (long long)r * W + c] = (int32_t)d2s[r][c];
      }
    }
  }
  if (thr_bits) {
    for (int i = tid; i < rows * (EDT_TW / 32); i += EDT_TW) {
      const int r = i / (EDT_TW / 32), w = i % (EDT_TW / 32);
      const int kw = (x0 >> 5) + w;
      if (kw < WW) thr_bits[(b * H + (q << 5) + r0 + r) * (long long)WW + kw] = tb[r][w];
    }
  }
}

// Phase 2: warp per flagged row -- full divide-and-conquer solve, rewrites the whole row
__global__ void __launch_bounds__(EDT_WARPS * 32)
    k_edt_far(const uint32_t* __restrict__ vw, const uint16_t* __restrict__ up, const uint16_t* __restrict__ dn,
              double* __restrict__ dist, int32_t* __restrict__ sq, uint32_t* __restrict__ thr_bits, int thr_sq,
              const uint8_t* __restrict__ row_far, int H, int W, int WW, int NB, int W2, int L) {
  extern __shared__ uint16_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int Wp = WW << 5;
  const int y = blockIdx.x * nwarps + warp;
  const long long b = blockIdx.y;
  if (y >= H || !row_far[b * H + y]) return;
  uint16_t* gs = smem + (size_t)warp * 2 * Wp;
  uint16_t* am = gs + Wp;
  const int q = y >> 5, r = y & 31;
  const long long band = (b * NB + q) * (long long)Wp;
  for (int x = lane; x < Wp; x += 32) {
    uint32_t gv = 0;
    if (x < W) gv = edt_vdist(~__ldg(vw + band + x), r, __ldg(up + band + x), __ldg(dn + band + x));
    gs[x] = (uint16_t)gv;
  }
  __syncwarp();
  edt_row_dc(gs, am, W, W2, L, lane);
  const long long orow = (b * H + y) * (long long)W;
  for (int x = lane; x < Wp; x += 32) {
    long long d2 = 0;
    bool nosite = false;
    if (x < W) {
      const uint32_t gx = gs[x];
      if (gx != 0u) {
        const uint32_t a = am[x];
        if (a == EDT_INF) {
          // no background pixel anywhere: scipy measures to the virtual point (-1, 0);
          // for the threshold output (dilation of an empty mask) the distance is infinite
          nosite = true;
          d2 = (long long)(y + 1) * (y + 1) + (long long)x * x;
        } else {
          const int dd = x - (int)a;
          const uint32_t gv = gs[a];
          d2 = (long long)dd * dd + (long long)gv * gv;
        }
      }
      if (dist) dist[orow + x] = sqrt((double)d2);
      if (sq) sq[orow + x] = (int32_t)d2;
    }
    if (thr_bits) {
      unsigned ball = __ballot_sync(0xffffffffu, x < W && !nosite && d2 <= (long long)thr_sq);
      if (lane == 0) thr_bits[(b * H + y) * (long long)WW + (x >> 5)] = ball;
    }
  }
}

extern "C" {

size_t pcs_edt_workspace_bytes(int B, int H, int W) {
  size_t NB = (H + 31) / 32, Wp = (size_t)pcs_words(W) * 32;
  return pcs_align256(B * NB * Wp * 4) + 2 * pcs_align256(B * NB * Wp * 2) + pcs_align256((size_t)B * H);
}

int pcs_edt_bits(const uint32_t* bits, int invert, int B, int H, int W, double* dist, int32_t* sq, uint32_t* thr_bits,
                 int thr_sq, void* ws, size_t ws_bytes, void* stream) {
  PCS_REQUIRE(B >= 1 && H >= 1 && W >= 1, "empty batch or image");
  PCS_REQUIRE(H <= 16384 && W <= 16384, "image side above 16384 is not supported");
  PCS_REQUIRE(B <= 65535, "batch above 65535");
  PCS_REQUIRE(dist || sq || thr_bits, "no output requested");
  if (ws == nullptr || ws_bytes < pcs_edt_workspace_bytes(B, H, W)) {
    pcs_set_error("EDT workspace too small (see pcs_edt_workspace_bytes)");
    return PCS_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int WW = pcs_words(W), Wp = WW << 5, NB = (H + 31) / 32;
  PCS_REQUIRE((long long)NB * (32 / EDT_ER) <= 65535, "too many row bands");
  int nwarps = EDT_WARPS;  // far-field solve: one row buffer pair (4 * Wp bytes) per warp in shared memory
  while (nwarps > 1 && (size_t)nwarps * 4 * Wp > 200 * 1024) nwarps >>= 1;
  size_t smem = (size_t)nwarps * 4 * Wp;
  char* p = (char*)ws;
  uint32_t* vw = (uint32_t*)p;
  p += pcs_align256((size_t)B * NB * Wp * 4);
  uint16_t* up = (uint16_t*)p;
  p += pcs_align256((size_t)B * NB * Wp * 2);
  uint16_t* dn = (uint16_t*)p;
  p += pcs_align256((size_t)B * NB * Wp * 2);
  uint8_t* row_far = (uint8_t*)p;
  cudaMemsetAsync(row_far, 0, (size_t)B * H, st);
  PCS_LAUNCH("k_edt_transpose", st,
             k_edt_transpose<<<pcs_blocks((long long)B * NB * ((WW + 3) / 4) * 32, 256), 256, 0, st>>>(bits, invert, vw, B, H, W, WW, NB));
  dim3 gc((W + 127) / 128, B, 2);
  PCS_LAUNCH("k_edt_carry", st, k_edt_carry<<<gc, 128, 0, st>>>(vw, up, dn, W, Wp, NB));
  dim3 gn((W + EDT_TW - 1) / EDT_TW, NB * (32 / EDT_ER), B);
  PCS_LAUNCH("k_edt_near", st,
             k_edt_near<<<gn, EDT_TW, 0, st>>>(vw, up, dn, dist, sq, thr_bits, thr_sq, row_far, H, W, WW, NB));
  int W2 = 1, L = 0;
  while (W2 <= W) {
    W2 <<= 1;
    ++L;
  }
  static size_t smem_set[64] = {};  // per device: largest opt-in shared-memory size set so far
  if (smem > 48 * 1024) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || smem > smem_set[dev]) {
      cudaFuncSetAttribute(k_edt_far, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (dev >= 0 && dev < 64) smem_set[dev] = smem;
    }
  }
  dim3 gf((H + nwarps - 1) / nwarps, B);
  PCS_LAUNCH("k_edt_far", st,
             k_edt_far<<<gf, nwarps * 32, smem, st>>>(vw, up, dn, dist, sq, thr_bits, thr_sq, row_far, H, W, WW, NB, W2, L));
  return pcs_check_launch("edt");
}

}  // extern "C"
