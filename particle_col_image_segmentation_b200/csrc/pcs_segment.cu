// Fused kernels of the z-stack segment pipeline (pcs_segment_chunk).  Each one merges stages that the
// drop-in primitives run as separate passes, so that a pixel, a mask word or a run label crosses HBM once:
//
//   k_seg_threshold_tile   image > Otsu threshold -> 5x5 binary median -> uint8 mask + bit rows
//                          -> tile-local union-find of the runs (the first pass of the labeller)
//                          -> per-run intensity sums.  The uint16 image is read here for the last time.
//                          (ilastik's role + scipy.ndimage.median_filter, tiff_analysis.py:122, :643;
//                           skimage.measure.label pass 1, tiff_analysis.py:743)
//   k_seg_rank_init        roots get their raster-order label; the root's table row is initialised with the
//                          fields a root knows by itself (first pixel, top row)
//   k_seg_relabel_table    run labels -> int32 label image (16-byte coalesced stores) and, from the same
//                          registers, the per-label reductions (regionprops area / centroid sums / bbox /
//                          integrated intensity, tiff_analysis.py:746-773); the label of every run is left in
//                          the parent plane for the refine stage
//
// Round 1 ran k_compare, k_majority5_bits, k_ccl_tile, k_table_init, k_ccl_rank, k_ccl_relabel and
// k_region_table_bits for this: the image was read three times, the label image written and re-read.
#include "pcs_ccl.cuh"

#include "pcs.h"

#define SEG_TR 32  // tile rows   } the labeller's tile (pcs_ccl.cu: PcsTile<PcsBinProv>)
#define SEG_TW 8   // tile words  }
#define SEG_THREADS 128
#ifndef SEG_MINBLOCKS
#define SEG_MINBLOCKS 8
#endif
#define SEG_LSPW 4
#define SEG_SPW 16
#define SEG_RR (SEG_TR + 4)  // staged rows: two halo rows above and below
#define SEG_RW (SEG_TW + 2)  // staged words per row: one halo word left and right (only two bits of each are used)

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

__device__ __forceinline__ int seg_lfind(volatile int* sp, int n) {
  int r = n, p = sp[r];
  const int first = p;
  while (p != r) {
    r = p;
    p = sp[r];
  }
  if (first != r) atomicMin((int*)sp + n, r);
  return r;
}

__device__ __forceinline__ void seg_lunion(int* sp, int a, int b) {
  while (true) {
    a = seg_lfind(sp, a);
    b = seg_lfind(sp, b);
    if (a == b) return;
    if (a < b) {
      int t = a;
      a = b;
      b = t;
    }
    int old = atomicMin(sp + a, b);
    if (old == a) return;
    a = old;
  }
}

// ---------------------------------------------------------------- TMA (bulk async copy) + mbarrier helpers
// cp.async.bulk (1-D): the copy engine of the SM moves a contiguous run of bytes from global to shared memory and
// signals an mbarrier with the byte count; no thread holds the data in registers on the way (SASS: UBLKCP).
// Measured on B200 (2048^2 x 64, profiles/README.md): 335 us per step with the bulk copies against 297 us with plain
// 128-bit loads -- every pixel is used exactly once, straight out of the register it was loaded into, so the detour
// through shared memory (one issuing thread, an mbarrier round trip, an LDS per vector) only adds latency.  Kept
// behind SEG_TMA for the record; the product builds with SEG_TMA = 0.
#ifndef SEG_TMA
#define SEG_TMA 0
#endif
__device__ __forceinline__ uint32_t seg_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void seg_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(seg_smem(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void seg_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(seg_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void seg_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(seg_smem(dst)), "l"(src), "r"(bytes),
               "r"(seg_smem(bar))
               : "memory");
}
__device__ __forceinline__ void seg_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(seg_smem(bar)), "r"(parity) : "memory");
  } while (!done);
}

// 8 pixels of a uint4 against the threshold -> 8 bits
__device__ __forceinline__ uint32_t seg_cmp8(const uint4& q, uint32_t t) {
  uint32_t m = 0;
  m |= (uint32_t)((q.x & 0xffffu) > t) << 0;
  m |= (uint32_t)((q.x >> 16) > t) << 1;
  m |= (uint32_t)((q.y & 0xffffu) > t) << 2;
  m |= (uint32_t)((q.y >> 16) > t) << 3;
  m |= (uint32_t)((q.z & 0xffffu) > t) << 4;
  m |= (uint32_t)((q.z >> 16) > t) << 5;
  m |= (uint32_t)((q.w & 0xffffu) > t) << 6;
  m |= (uint32_t)((q.w >> 16) > t) << 7;
  return m;
}

// CTA = one labelling tile (32 rows x 256 pixels) of one slice.
//  1. stage: rows y0-2 .. y0+33, each as one fully coalesced 512-byte warp load (all of a thread's loads are
//     issued before the first compare), compared against the slice's threshold and packed to bit rows in shared
//     memory; the two halo pixels left and right of the tile come from 4-byte loads;
//  2. median: bit-sliced carry-save count of the 25 neighbours (scipy 'reflect' at the image border), thread =
//     (word column, pair of rows); the final word goes to the bit plane, the uint8 mask and shared memory;
//  3. label pass 1 on the words in shared memory: slots per run, unions inside the tile with shared atomics,
//     every run leaves with the id of its tile-local root (same result as k_ccl_tile<PcsBinProv, 8>);
//  4. per-run sums of the pixel values (the pixels come back from L1/L2: this CTA has just read them).
// The memory-bound stage of one CTA overlaps the latency-bound union-find of its neighbours on the SM.
template <bool MEDIAN>
__global__ void __launch_bounds__(SEG_THREADS, SEG_MINBLOCKS)
    k_seg_threshold_tile(const uint16_t* __restrict__ img, const int32_t* __restrict__ thr, uint32_t* __restrict__ bits,
                         uint8_t* __restrict__ mask, int* __restrict__ parent, int* __restrict__ rsum, int* __restrict__ wlist,
                         int* __restrict__ wcount, int H, int W, int WW) {
  constexpr int NWORDS = SEG_TR * SEG_TW;
  __shared__ int s_lbase;
  __shared__ uint32_t raw[SEG_RR][SEG_RW];
#if SEG_TMA
  // The staged pixel rows (TMA destination) and the union-find slots never live at the same time: the pixels are dead
  // once the raw bits are packed (first barrier), the slots are born after it.  One buffer serves both.
  constexpr int STAGE_BYTES = SEG_RR * 32 * SEG_TW * 2, SLOT_BYTES = NWORDS * SEG_SPW * 4;
  __shared__ __align__(128) unsigned char stage_or_slots[STAGE_BYTES > SLOT_BYTES ? STAGE_BYTES : SLOT_BYTES];
  __shared__ __align__(8) uint64_t mbar;
  int* sp = reinterpret_cast<int*>(stage_or_slots);
#else
  __shared__ int sp[NWORDS * SEG_SPW];
#endif
  __shared__ uint32_t fsm[NWORDS], ssm[NWORDS];
  __shared__ unsigned short items[NWORDS];
  __shared__ int nitems;
  const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
  const int k0 = blockIdx.x * SEG_TW, y0 = blockIdx.y * SEG_TR, x0 = k0 << 5;
  const long long b = blockIdx.z;
  const uint16_t* src = img + b * (long long)H * W;
  const uint32_t t = (uint32_t)max(thr[b], 0);  // a negative threshold cannot occur (Otsu of uint16 data)
  if (tid == 0) nitems = 0;
  const int halo = MEDIAN ? 2 : 0;
  const bool inner = k0 > 0 && ((k0 + SEG_TW) << 5) + 16 <= W;  // no word of this tile has a window beyond the row ends
  // ---------------- 1. stage the raw threshold bits
  const bool fast = x0 + 32 * SEG_TW <= W && (W & 7) == 0 && ((((uintptr_t)src) & 15) == 0);
  if (fast) {
    constexpr int RPW = SEG_RR / 4;  // rows per warp
#if SEG_TMA
    // One thread arms the barrier with the byte count and issues one 512-byte bulk copy per staged row; the SM's copy
    // engine lands the rows in shared memory (UBLKCP), every thread waits on the mbarrier and reads its 8 pixels back.
    uint16_t (*pix)[32 * SEG_TW] = reinterpret_cast<uint16_t (*)[32 * SEG_TW]>(stage_or_slots);
    const int r_lo = max(2 - halo, 2 - y0), r_hi = min(SEG_TR + 2 + halo, H - y0 + 2);  // staged rows [r_lo, r_hi) lie in the image
    if (tid == 0) seg_mbar_init(&mbar, 1);
    __syncthreads();
    if (tid == 0) {
      seg_mbar_expect_tx(&mbar, (uint32_t)(r_hi - r_lo) * 32 * SEG_TW * 2);
      for (int r = r_lo; r < r_hi; ++r) seg_bulk_g2s(&pix[r][0], src + (long long)(y0 - 2 + r) * W + x0, 32 * SEG_TW * 2, &mbar);
    }
    seg_mbar_wait(&mbar, 0);
    uint4 q[RPW];
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
      const int r = wq + 4 * i;
      q[i] = make_uint4(0, 0, 0, 0);
      if (r >= r_lo && r < r_hi) q[i] = *reinterpret_cast<const uint4*>(&pix[r][lane * 8]);
    }
#else
    uint4 q[RPW];
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
      const int r = wq + 4 * i, yi = y0 - 2 + r;
      q[i] = make_uint4(0, 0, 0, 0);
      if (r >= 2 - halo && r < SEG_TR + 2 + halo && yi >= 0 && yi < H) q[i] = __ldg(reinterpret_cast<const uint4*>(src + (long long)yi * W + x0) + lane);
    }
#endif
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
      const int r = wq + 4 * i;
      uint32_t v = seg_cmp8(q[i], t) << (8 * (lane & 3));  // lanes 4w .. 4w+3 hold the four bytes of word w
      v |= __shfl_xor_sync(0xffffffffu, v, 1);
      v |= __shfl_xor_sync(0xffffffffu, v, 2);
      if ((lane & 3) == 0) raw[r][1 + (lane >> 2)] = v;
    }
  } else {
    for (int i = tid; i < SEG_RR * SEG_TW; i += SEG_THREADS) {
      const int r = i / SEG_TW, c = i % SEG_TW, yi = y0 - 2 + r, k = k0 + c;
      uint32_t v = 0;
      if (r >= 2 - halo && r < SEG_TR + 2 + halo && yi >= 0 && yi < H && k < WW) {
        const uint16_t* row = src + (long long)yi * W;
        const int n = min(32, W - (k << 5));
        for (int e = 0; e < n; ++e) v |= (uint32_t)(row[(k << 5) + e] > t) << e;
      }
      raw[r][1 + c] = v;
    }
  }
  if (MEDIAN && tid < 2 * SEG_RR) {  // halo pixels: x0-2, x0-1 (bits 30, 31 of word k0-1) and x0+256, x0+257 (bits 0, 1 of word k0+8)
    const int r = tid >> 1, side = tid & 1, yi = y0 - 2 + r;
    uint32_t v = 0;
    if (yi >= 0 && yi < H) {
      const uint16_t* row = src + (long long)yi * W;
      if (side == 0) {
        if (x0 >= 2) v = ((uint32_t)(row[x0 - 2] > t) << 30) | ((uint32_t)(row[x0 - 1] > t) << 31);
      } else {
        const int xr = x0 + 32 * SEG_TW;
        if (xr < W) v |= (uint32_t)(row[xr] > t);
        if (xr + 1 < W) v |= (uint32_t)(row[xr + 1] > t) << 1;
      }
    }
    raw[r][side ? SEG_RW - 1 : 0] = v;
  }
  __syncthreads();
  // ---------------- 2. median + outputs; thread = (word column c, rows 2 * strip and 2 * strip + 1 of the tile)
  const int c = tid & (SEG_TW - 1), strip = tid >> 3;
  const int k = k0 + c;
  uint32_t fin[2] = {0u, 0u};
  if (k < WW) {
    const uint32_t vm = pcs_valid_mask(k, W);
    if (MEDIAN) {
      // the six input rows' windows first: nine pixels in ten are background, and a thread whose windows hold no set
      // bit at all (bits 14 .. 49 are the ones the 5-wide sums look at) has two all-zero output words and skips the
      // carry-save adders -- most of this phase's instructions
      unsigned long long win[6];
      unsigned long long seen = 0;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int yi = y0 + 2 * strip - 2 + i;  // input row
        win[i] = 0;
        if (yi - 2 < H) {  // still needed by an output row of the image
          const int ry = pcs_reflect(yi, H);
          const uint32_t* rowp = &raw[ry - (y0 - 2)][1] - k0;  // rowp[k'] = word k' of that row
          // words away from the left / right image border need no reflection: three shared loads and two funnel shifts
          win[i] = inner ? ((unsigned long long)rowp[k] << 16) | (rowp[k - 1] >> 16) | ((unsigned long long)(rowp[k + 1] & 0xffffu) << 48)
                         : pcs_window_reflect(rowp, k, W, WW, 2);
          seen |= win[i];
        }
      }
      if (seen & 0x0003ffffffffc000ull) {
        uint32_t r0[6], r1[6], r2[6];  // bit-sliced horizontal counts of the six input rows
#pragma unroll
        for (int i = 0; i < 6; ++i)
          pcs_add5((uint32_t)(win[i] >> 14), (uint32_t)(win[i] >> 15), (uint32_t)(win[i] >> 16), (uint32_t)(win[i] >> 17), (uint32_t)(win[i] >> 18),
                   r0[i], r1[i], r2[i]);
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const uint32_t a0[5] = {r0[o], r0[o + 1], r0[o + 2], r0[o + 3], r0[o + 4]};
          const uint32_t a1[5] = {r1[o], r1[o + 1], r1[o + 2], r1[o + 3], r1[o + 4]};
          const uint32_t a2[5] = {r2[o], r2[o + 1], r2[o + 2], r2[o + 3], r2[o + 4]};
          if (y0 + 2 * strip + o < H) fin[o] = pcs_majority5_word(a0, a1, a2) & vm;
        }
      }
    } else {
#pragma unroll
      for (int o = 0; o < 2; ++o)
        if (y0 + 2 * strip + o < H) fin[o] = raw[2 + 2 * strip + o][1 + c] & vm;
    }
  }
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    const int wr = 2 * strip + o, w = wr * SEG_TW + c, y = y0 + wr;
    const uint32_t F = fin[o];
    uint32_t S = F & ~(F << 1);
    fsm[w] = F;
    ssm[w] = S;
    if (k < WW && y < H) {
      bits[(b * H + y) * (long long)WW + k] = F;
      pcs_store_mask_bytes(mask + (b * H + y) * (long long)W, k, W, F);
    }
    const int sbase = w << SEG_LSPW;
    for (int j = 0; S; ++j) {
      S &= S - 1;
      sp[sbase + j] = sbase + j;
    }
    // list the non-empty words: one shared atomic per warp
    const unsigned has = __ballot_sync(0xffffffffu, F != 0u);
    if (has) {
      const int leader = __ffs(has) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(&nitems, __popc(has));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (F) items[base + __popc(has & ((1u << lane) - 1u))] = (unsigned short)w;
    }
  }
  __syncthreads();
  const int n = nitems;
  if (n == 0) return;  // empty tile (uniform)
  // the slice's list of non-empty words: the sparse passes that follow (tile-edge unions, flatten, ranking) walk this
  // list instead of all words -- four words in five are empty
  if (tid == 0) s_lbase = atomicAdd(wcount + b, n);
  // ---------------- 3. unions between runs of this tile (8-connectivity), thread per non-empty word
  for (int it = tid; it < n; it += SEG_THREADS) {
    const int w = items[it], wr = w / SEG_TW, wc = w % SEG_TW;
    const uint32_t F = fsm[w];
    const int sbase = w << SEG_LSPW;
    if ((F & 1u) && wc > 0 && (fsm[w - 1] >> 31)) seg_lunion(sp, sbase, sbase - SEG_SPW + __popc(ssm[w - 1]) - 1);
    if (wr == 0) continue;
    const uint32_t al = wc > 0 ? fsm[w - SEG_TW - 1] : 0u, ac = fsm[w - SEG_TW], ar = wc < SEG_TW - 1 ? fsm[w - SEG_TW + 1] : 0u;
    const uint32_t U = F & ac, UL = F & ((ac << 1) | (al >> 31)), UR = F & ((ac >> 1) | (ar << 31));
    if (!(U | UL | UR)) continue;
    const uint32_t Sa[3] = {wc > 0 ? ssm[w - SEG_TW - 1] : 0u, ssm[w - SEG_TW], wc < SEG_TW - 1 ? ssm[w - SEG_TW + 1] : 0u};
    uint32_t S = ssm[w];
    for (int j = 0; S; ++j) {
      int s;
      const uint32_t R = pcs_pop_run(F, S, s);
      unsigned long long T = (((unsigned long long)(U & R)) << 1) | (unsigned long long)(UL & R) | (((unsigned long long)(UR & R)) << 2);
      while (T) {
        const int i = __ffsll((long long)T) - 1;
        T &= T + (1ull << i);
        const int rel = (i - 1) >> 5, ca = wc + rel;
        if (ca >= 0 && ca < SEG_TW) {  // the run above lives in this tile
          const uint32_t sa_w = Sa[rel + 1];
          const int sa = pcs_start_at_or_below(sa_w, (i - 1) & 31);
          seg_lunion(sp, sbase + j, ((w - SEG_TW + rel) << SEG_LSPW) + pcs_run_ord(sa_w, sa));
        }
      }
    }
  }
  __syncthreads();
  // ---------------- 4. publish: every run points at the global id of its tile-local root; per-run intensity sums
  const int NW = H * WW;
  int* par = parent + b * ((long long)NW << SEG_LSPW);
  int* rs = rsum + b * ((long long)NW << SEG_LSPW);
  int* wl = wlist + b * (long long)NW + s_lbase;
  for (int it = tid; it < n; it += SEG_THREADS) {
    const int w = items[it], wr = w / SEG_TW, wc = w % SEG_TW;
    const int sbase = w << SEG_LSPW;
    const int y = y0 + wr, kk = k0 + wc, gw = y * WW + kk;
    wl[it] = gw;
    const uint32_t F = fsm[w];
    uint32_t S = ssm[w];
    const uint16_t* px = src + (long long)y * W + (kk << 5);
    const bool vec = (kk << 5) + 32 <= W && ((((uintptr_t)px) & 15) == 0);
    uint32_t iw[16];
    if (vec) {
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const uint4 q4 = __ldg(reinterpret_cast<const uint4*>(px) + v);
        iw[4 * v] = q4.x;
        iw[4 * v + 1] = q4.y;
        iw[4 * v + 2] = q4.z;
        iw[4 * v + 3] = q4.w;
      }
    }
    for (int j = 0; S; ++j) {
      int s;
      const uint32_t R = pcs_pop_run(F, S, s);
      const int root = seg_lfind(sp, sbase + j);
      const int rw = root >> SEG_LSPW, rj = root & (SEG_SPW - 1);
      par[j * NW + gw] = pcs_node<SEG_LSPW>((y0 + rw / SEG_TW) * WW + k0 + rw % SEG_TW, rj);
      uint32_t sum = 0;  // <= 32 pixels of 16 bits: no overflow
      if (vec) {
        // IDP.2A: two 16-bit pixels times two mask bytes (0 / 1) per instruction; the four mask bits of a nibble are
        // spread to four bytes by one multiply (no carries: the shifted copies of a 4-bit value do not overlap)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t bm = (((R >> (4 * i)) & 0xfu) * 0x00204081u) & 0x01010101u;
          sum = __dp2a_lo(iw[2 * i], bm, sum);
          sum = __dp2a_hi(iw[2 * i + 1], bm, sum);
        }
      } else {
        for (uint32_t m = R; m; m &= m - 1) sum += px[__ffs(m) - 1];
      }
      rs[j * NW + gw] = (int)sum;
    }
  }
}

// ---------------------------------------------------------------- sparse passes over the list of non-empty words
// grid = (GX, B); thread per listed word (grid-stride), the slice's count is read from device memory.
#define SEG_LIST_THREADS 256

// the unions k_seg_threshold_tile could not do: adjacencies across a tile edge (top row of a tile: everything above;
// first / last word column: the run to the left and the diagonal neighbours beyond the column)
__global__ void __launch_bounds__(SEG_LIST_THREADS)
    k_seg_merge_list(const uint32_t* __restrict__ bits, int* __restrict__ parent, const int* __restrict__ wlist,
                     const int* __restrict__ wcount, int H, int W, int WW) {
  const long long b = blockIdx.y;
  const int NW = H * WW;
  const int n = min(wcount[b], NW);
  const uint32_t* bb = bits + b * (long long)NW;
  int* par = parent + b * ((long long)NW << SEG_LSPW);
  const int* wl = wlist + b * (long long)NW;
  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < n; it += gridDim.x * blockDim.x) {
    const int gw = wl[it], y = gw / WW, k = gw - y * WW;
    const bool top = (y % SEG_TR) == 0, left = (k % SEG_TW) == 0, right = (k % SEG_TW) == SEG_TW - 1;
    if (!(top || left || right)) continue;
    const uint32_t F = __ldg(bb + gw);
    if (left && k > 0 && (F & 1u)) {
      const uint32_t Fl = __ldg(bb + gw - 1);
      if (Fl >> 31) pcs_uf_union<SEG_LSPW>(par, NW, pcs_node<SEG_LSPW>(gw, 0), pcs_node<SEG_LSPW>(gw - 1, __popc(Fl & ~(Fl << 1)) - 1));
    }
    if (y == 0) continue;
    const uint32_t al = k > 0 ? __ldg(bb + gw - WW - 1) : 0u, ac = __ldg(bb + gw - WW), ar = k + 1 < WW ? __ldg(bb + gw - WW + 1) : 0u;
    const uint32_t U = F & ac, UL = F & ((ac << 1) | (al >> 31)), UR = F & ((ac >> 1) | (ar << 31));
    if (!(U | UL | UR)) continue;
    const uint32_t Sa[3] = {al & ~(al << 1), ac & ~(ac << 1), ar & ~(ar << 1)};
    uint32_t S = F & ~(F << 1);
    for (int j = 0; S; ++j) {
      int s;
      const uint32_t R = pcs_pop_run(F, S, s);
      unsigned long long T = (((unsigned long long)(U & R)) << 1) | (unsigned long long)(UL & R) | (((unsigned long long)(UR & R)) << 2);
      while (T) {
        const int i = __ffsll((long long)T) - 1;
        T &= T + (1ull << i);
        const int rel = (i - 1) >> 5;
        if (!(top || (rel < 0 && left) || (rel > 0 && right))) continue;  // done inside the tile
        const uint32_t sa_w = Sa[rel + 1];
        const int sa = pcs_start_at_or_below(sa_w, (i - 1) & 31);
        pcs_uf_union<SEG_LSPW>(par, NW, pcs_node<SEG_LSPW>(gw, j), pcs_node<SEG_LSPW>(gw - WW + rel, pcs_run_ord(sa_w, sa)));
      }
    }
  }
}

// every run points at its root; root flags per word (bit j: the run of ordinal j is a root); roots per 32-word chunk
// (rootbits and chunk are zeroed beforehand: only listed words are written)
__global__ void __launch_bounds__(SEG_LIST_THREADS)
    k_seg_flatten_list(const uint32_t* __restrict__ bits, int* __restrict__ parent, const int* __restrict__ wlist,
                       const int* __restrict__ wcount, uint32_t* __restrict__ rootbits, int* __restrict__ chunk, int H, int W, int WW,
                       int CPR) {
  const long long b = blockIdx.y;
  const int NW = H * WW;
  const int n = min(wcount[b], NW);
  const uint32_t* bb = bits + b * (long long)NW;
  int* par = parent + b * ((long long)NW << SEG_LSPW);
  const int* wl = wlist + b * (long long)NW;
  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < n; it += gridDim.x * blockDim.x) {
    const int gw = wl[it];
    const int q0 = pcs_ld_cg(par + gw);  // plane 0, requested beside the word: a listed word has at least one run
    const uint32_t F = __ldg(bb + gw);
    const int nr = __popc(F & ~(F << 1));
    uint32_t roots = 0;
    for (int j = 0; j < nr; ++j) {
      const int nd = pcs_node<SEG_LSPW>(gw, j);
      int r = nd, q = j == 0 ? q0 : pcs_ld_cg(par + j * NW + gw);
      const int first = q;
      while (q != r) {
        r = q;
        q = pcs_ld_cg(par + pcs_slot<SEG_LSPW>(r, NW));
      }
      if (r == nd)
        roots |= 1u << j;
      else if (first != r)
        par[j * NW + gw] = r;
    }
    rootbits[b * (long long)NW + gw] = roots;
    if (roots) {
      const int y = gw / WW, k = gw - y * WW;
      atomicAdd(chunk + b * (long long)H * CPR + y * CPR + (k >> 5), __popc(roots));
    }
  }
}

// block per slice: exclusive scan of the slice's chunk counts (raster order) and the slice total; the block that
// finishes last also scans the totals into the table row offsets (one launch instead of two: both are latency-only)
__global__ void __launch_bounds__(1024)
    k_seg_scan_offsets(int* __restrict__ chunk, int32_t* __restrict__ counts, int n, int* __restrict__ offsets, int* __restrict__ done, int B) {
  __shared__ int wsum[32];
  __shared__ int carry_s;
  __shared__ int is_last;
  int* c = chunk + (long long)blockIdx.x * n;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + tid;
    const int v = i < n ? c[i] : 0;
    const int carry = carry_s;  // stable here: last written before the previous iteration's barriers
    int tot;
    const int ex = pcs_warp_excl_scan(v, lane, &tot);
    if (lane == 0) wsum[wid] = tot;
    __syncthreads();
    if (wid == 0) {
      const int w = lane < (int)(blockDim.x >> 5) ? wsum[lane] : 0;
      int wt;
      const int wex = pcs_warp_excl_scan(w, lane, &wt);
      wsum[lane] = wex;
      if (lane == 0) carry_s = carry + wt;
    }
    __syncthreads();
    if (i < n) c[i] = carry + wsum[wid] + ex;
    __syncthreads();
  }
  if (tid == 0) {
    counts[blockIdx.x] = carry_s;
    __threadfence();
    is_last = atomicAdd(done, 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < B; base += blockDim.x) {
    const int i = base + tid;
    const int v = i < B ? __ldcg(counts + i) : 0;
    const int carry = carry_s;
    int tot;
    const int ex = pcs_warp_excl_scan(v, lane, &tot);
    if (lane == 0) wsum[wid] = tot;
    __syncthreads();
    if (wid == 0) {
      const int w = wsum[lane];
      int wt;
      const int wex = pcs_warp_excl_scan(w, lane, &wt);
      wsum[lane] = wex;
      if (lane == 0) carry_s = carry + wt;
    }
    __syncthreads();
    if (i < B) offsets[i] = carry + wsum[wid] + ex;
    __syncthreads();
  }
  if (tid == 0) offsets[B] = carry_s;
}

// roots get their raster-order rank (stored negated in the parent plane) and their table row is initialised.  The
// rank of a root = roots of the slice before its 32-word chunk (scanned chunk counts) + roots in the chunk's earlier
// words + its ordinal among the roots of its own word.  A root is the first raster pixel of its component, so the
// first-pixel column and the top row of the bounding box are final here; the sums start at zero, the other bounds
// at their neutral values.
__global__ void __launch_bounds__(SEG_LIST_THREADS)
    k_seg_rank_list(const uint32_t* __restrict__ bits, int* __restrict__ parent, const int* __restrict__ wlist,
                    const int* __restrict__ wcount, const uint32_t* __restrict__ rootbits, const int* __restrict__ chunk,
                    const int* __restrict__ offsets, long long* __restrict__ table, long long cap, int H, int W, int WW, int CPR) {
  const long long b = blockIdx.y;
  const int NW = H * WW;
  const int n = min(wcount[b], NW);
  const uint32_t* bb = bits + b * (long long)NW;
  const uint32_t* rb = rootbits + b * (long long)NW;
  int* par = parent + b * ((long long)NW << SEG_LSPW);
  const int* wl = wlist + b * (long long)NW;
  const long long trow = offsets[b];
  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < n; it += gridDim.x * blockDim.x) {
    const int gw = wl[it];
    const uint32_t roots = rb[gw];
    if (!roots) continue;
    const int y = gw / WW, k = gw - y * WW;
    int rank = chunk[b * (long long)H * CPR + y * CPR + (k >> 5)];
    for (int w = gw - (k & 31); w < gw; ++w) rank += __popc(__ldg(rb + w));
    const uint32_t F = __ldg(bb + gw);
    uint32_t S = F & ~(F << 1);
    for (int j = 0; S; ++j) {
      const int s = __ffs(S) - 1;
      S &= S - 1;
      if (!((roots >> j) & 1u)) continue;
      ++rank;
      par[j * NW + gw] = -rank;
      const long long row = trow + rank - 1;
      if (row < cap) {
        table[T_AREA * cap + row] = 0;
        table[T_SUMY * cap + row] = 0;
        table[T_SUMX * cap + row] = 0;
        table[T_MINY * cap + row] = y;
        table[T_MINX * cap + row] = 0x7fffffffffffffffLL;
        table[T_MAXY * cap + row] = -1;
        table[T_MAXX * cap + row] = -1;
        table[T_FIRST * cap + row] = (long long)y * W + (k << 5) + s;
        table[T_SUMI * cap + row] = 0;
        table[T_OVERLAP * cap + row] = 0;
      }
    }
  }
}

struct SegAcc {
  int label;
  int area, minx, maxx, maxy;
  long long sx, sy, si;
};

__device__ __noinline__ void seg_flush(const SegAcc& a, long long* __restrict__ table, long long cap, long long base) {
  if (a.label <= 0) return;
  const long long row = base + a.label - 1;
  if (row >= cap) return;
  typedef unsigned long long ull;
  atomicAdd((ull*)(table + T_AREA * cap + row), (ull)a.area);
  atomicAdd((ull*)(table + T_SUMY * cap + row), (ull)a.sy);
  atomicAdd((ull*)(table + T_SUMX * cap + row), (ull)a.sx);
  atomicAdd((ull*)(table + T_SUMI * cap + row), (ull)a.si);
  atomicMin(table + T_MINX * cap + row, (long long)a.minx);
  atomicMax(table + T_MAXY * cap + row, (long long)a.maxy);
  atomicMax(table + T_MAXX * cap + row, (long long)a.maxx);
}

// warp per (32-word chunk, strip of SEG_RROWS rows): labels out + per-label reductions.
//  * a lane owns one word column and walks down the strip.  The label of a run comes from the parent plane (node ->
//    root -> rank: two dependent loads after the word itself, all L2 hits on the dense planes).  The three levels are
//    software-pipelined over the rows -- while row r is expanded, the word of row r + 3, the parent / run sum of row
//    r + 2 and the root rank of row r + 1 are in flight -- so the chain costs one round trip per strip, not three per
//    row (stall samples of the unpipelined version: 40 % on that chain).  The label of every run is written back
//    (negated) so that the refine stage reads labels from that plane instead of the label image;
//  * a row's 32 words are expanded to pixels with 16-byte stores, a lane writing 4 consecutive pixels: single-run
//    words broadcast their label by shuffle, words with several runs (4 % of the non-empty ones) go through a small
//    shared table.  (Letting the owner lane of such a word write its 32 pixels itself looked cheaper and was not: one
//    warp row in five holds such a word, and the whole warp then walks the per-pixel loop -- 157 vs 103 us per 16
//    slices under ncu);
//  * the same lane accumulates area, coordinate sums, bbox and intensity (from the per-run sums k_seg_threshold_tile
//    left) of the component it is walking through and reaches the table with 7 atomics only when the label under
//    it changes: a blob crossing the strip costs one flush per column, not one per run.
#ifndef SEG_RROWS
#define SEG_RROWS 8
#endif
#ifndef SEG_RL_WARPS
#define SEG_RL_WARPS 4
#endif

__device__ __forceinline__ void seg_acc_run(SegAcc& acc, int l, int len, int xs, int y, int si, long long* __restrict__ table,
                                            long long cap, long long tbase) {
  if (l != acc.label) {
    seg_flush(acc, table, cap, tbase);
    acc.label = l;
    acc.area = 0;
    acc.sx = acc.sy = acc.si = 0;
    acc.minx = xs;
    acc.maxx = xs + len - 1;
  }
  acc.area += len;
  acc.sx += (long long)len * xs + (long long)(len * (len - 1) / 2);
  acc.sy += (long long)len * y;
  acc.si += si;
  acc.minx = min(acc.minx, xs);
  acc.maxx = max(acc.maxx, xs + len - 1);
  acc.maxy = y;
}

__device__ __forceinline__ bool seg_single_run(uint32_t F) {
  const uint32_t S = F & ~(F << 1);
  return S != 0u && (S & (S - 1)) == 0u;
}

template <typename OutT>
__global__ void __launch_bounds__(SEG_RL_WARPS * 32)
    k_seg_relabel_table(const uint32_t* __restrict__ bits, int* __restrict__ parent, const int* __restrict__ rsum,
                        const int* __restrict__ offsets, long long* __restrict__ table, long long cap, OutT* __restrict__ out, int H, int W,
                        int WW, int CPR, int strips) {
  __shared__ int lab[SEG_RL_WARPS][32][33];
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // (strip, chunk) of the slice; the slice is blockIdx.y
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  if (g >= strips * CPR) return;
  const long long b = blockIdx.y;
  const int ch = g % CPR, strip = g / CPR;
  const int k = ch * 32 + lane;
  const int NW = H * WW;
  int* par = parent + b * ((long long)NW << SEG_LSPW);
  const int* rs = rsum + b * ((long long)NW << SEG_LSPW);
  const long long tbase = offsets[b];
  const int nk = min(32, WW - ch * 32);
  const bool vec_out = sizeof(OutT) == 4 && (W & 3) == 0;
  const int y0 = strip * SEG_RROWS;
  const int yend = min(H, y0 + SEG_RROWS);
  const int x0 = k << 5;
  const uint32_t* fcol = bits + b * (long long)NW + k;
  auto ldF = [&](int y) -> uint32_t { return (k < WW && y < yend) ? __ldg(fcol + (long long)y * WW) : 0u; };
  // pipeline registers: F of rows y, y+1, y+2; parent / run sum of rows y, y+1; label of row y
  uint32_t Fa = ldF(y0), Fb = ldF(y0 + 1), Fc = ldF(y0 + 2);
  int pa = 0, sa = 0, pb = 0, sb = 0, oa = 0;
  if (seg_single_run(Fa)) {
    pa = par[y0 * WW + k];
    sa = rs[y0 * WW + k];
  }
  if (seg_single_run(Fb)) {
    pb = par[(y0 + 1) * WW + k];
    sb = rs[(y0 + 1) * WW + k];
  }
  if (seg_single_run(Fa)) oa = pa < 0 ? -pa : -par[pcs_slot<SEG_LSPW>(pa, NW)];
  SegAcc acc;
  acc.label = 0;
#pragma unroll 1
  for (int y = y0; y < yend; ++y) {
    // issue the next levels
    const uint32_t Fd = ldF(y + 3);
    int pc = 0, sc = 0, ob = 0;
    if (seg_single_run(Fc)) {
      pc = par[(y + 2) * WW + k];
      sc = rs[(y + 2) * WW + k];
    }
    if (seg_single_run(Fb)) ob = pb < 0 ? -pb : -par[pcs_slot<SEG_LSPW>(pb, NW)];
    // row y
    const uint32_t F = Fa;
    const uint32_t S = F & ~(F << 1);
    const int one = oa;  // label of the word's only run (0: empty word or several runs)
    const int gw = y * WW + k;
    OutT* orow = out + (b * H + y) * (long long)W;
    if (S) {
      if (one) {
        if (pa >= 0) par[gw] = -one;
        seg_acc_run(acc, one, __popc(F), x0 + __ffs(S) - 1, y, sa, table, cap, tbase);
      } else {  // several runs: their labels go through shared memory
        uint32_t rem = S;
        for (int j = 0; rem; ++j) {
          int s;
          const uint32_t R = pcs_pop_run(F, rem, s);
          const int p0 = par[j * NW + gw];
          const int l = p0 < 0 ? -p0 : -par[pcs_slot<SEG_LSPW>(p0, NW)];
          if (p0 >= 0) par[j * NW + gw] = -l;
          lab[wl][lane][s] = l;
          seg_acc_run(acc, l, __popc(R), x0 + s, y, rs[j * NW + gw], table, cap, tbase);
        }
      }
    }
    __syncwarp();
    const bool any = __ballot_sync(0xffffffffu, F != 0u) != 0u;
    if (vec_out) {
      const int sub = lane >> 3, nib = (lane & 7) << 2;
      for (int k4 = 0; k4 < nk; k4 += 4) {
        const int kk = k4 + sub;
        int4 v = make_int4(0, 0, 0, 0);
        if (any) {
          const uint32_t f = __shfl_sync(0xffffffffu, F, kk & 31);
          const int l1 = __shfl_sync(0xffffffffu, one, kk & 31);
          const uint32_t nb = kk < nk ? (f >> nib) & 0xfu : 0u;
          if (nb) {
            if (l1) {
              v.x = (nb & 1u) ? l1 : 0;
              v.y = (nb & 2u) ? l1 : 0;
              v.z = (nb & 4u) ? l1 : 0;
              v.w = (nb & 8u) ? l1 : 0;
            } else {
              const uint32_t s = f & ~(f << 1);
              if (nb & 1u) v.x = lab[wl][kk][pcs_start_at_or_below(s, nib)];
              if (nb & 2u) v.y = lab[wl][kk][pcs_start_at_or_below(s, nib + 1)];
              if (nb & 4u) v.z = lab[wl][kk][pcs_start_at_or_below(s, nib + 2)];
              if (nb & 8u) v.w = lab[wl][kk][pcs_start_at_or_below(s, nib + 3)];
            }
          }
        }
        const int x = ((ch * 32 + kk) << 5) + nib;
        if (kk < nk && x < W) *reinterpret_cast<int4*>(orow + x) = v;  // W % 4 == 0: x + 3 < W too
      }
    } else {
      for (int kk = 0; kk < nk; ++kk) {
        const uint32_t f = any ? __shfl_sync(0xffffffffu, F, kk) : 0u;
        const int l1 = any ? __shfl_sync(0xffffffffu, one, kk) : 0;
        const int x = ((ch * 32 + kk) << 5) + lane;
        int v = 0;
        if ((f >> lane) & 1u) v = l1 ? l1 : lab[wl][kk][pcs_start_at_or_below(f & ~(f << 1), lane)];
        if (x < W) orow[x] = (OutT)v;
      }
    }
    __syncwarp();  // the next row reuses lab[wl]
    // rotate the pipeline
    Fa = Fb;
    Fb = Fc;
    Fc = Fd;
    pa = pb;
    sa = sb;
    pb = pc;
    sb = sc;
    oa = ob;
  }
  seg_flush(acc, table, cap, tbase);
}

// every run gets its label (root rank) written back, negated, in place of its parent -- the expansion to pixels and
// the refine stage then read labels with ONE load per run -- and adds its area, coordinate sums, horizontal extent,
// bottom row and intensity sum to its component's table row.  All of a thread's loads are independent of every other
// thread's, so the node -> root -> rank chain is hidden by thread-level parallelism alone.
// Warp aggregation before the global atomics: consecutive list entries are neighbouring words of a tile, so several
// lanes of a warp usually carry the same label.  __match_any_sync groups them, REDUX (__reduce_*_sync over the group's
// mask; per-warp partial sums fit 32 bits) folds the seven quantities and the group's first lane issues the atomics:
// about a third of the atomics of the one-per-run version.
__device__ __forceinline__ void seg_table_add(long long* __restrict__ table, long long cap, long long row, int area, int sy, int sx,
                                              int si, int minx, int maxx, int maxy) {
  typedef unsigned long long ull;
  atomicAdd((ull*)(table + T_AREA * cap + row), (ull)area);
  atomicAdd((ull*)(table + T_SUMY * cap + row), (ull)sy);
  atomicAdd((ull*)(table + T_SUMX * cap + row), (ull)sx);
  atomicAdd((ull*)(table + T_SUMI * cap + row), (ull)si);
  atomicMin(table + T_MINX * cap + row, (long long)minx);
  atomicMax(table + T_MAXY * cap + row, (long long)maxy);
  atomicMax(table + T_MAXX * cap + row, (long long)maxx);
}

__global__ void __launch_bounds__(SEG_LIST_THREADS)
    k_seg_label_list(const uint32_t* __restrict__ bits, int* __restrict__ parent, const int* __restrict__ rsum,
                     const int* __restrict__ wlist, const int* __restrict__ wcount, const int* __restrict__ offsets,
                     long long* __restrict__ table, long long cap, int H, int W, int WW) {
  const long long b = blockIdx.y;
  const int NW = H * WW;
  const int n = min(wcount[b], NW);
  const uint32_t* bb = bits + b * (long long)NW;
  int* par = parent + b * ((long long)NW << SEG_LSPW);
  const int* rs = rsum + b * ((long long)NW << SEG_LSPW);
  const int* wl = wlist + b * (long long)NW;
  const long long trow = offsets[b];
  const int lane = threadIdx.x & 31;
  // warp-uniform trip count: lanes past the end of the list take part in the votes with nothing to add
  for (int it0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); it0 < n; it0 += gridDim.x * blockDim.x) {
    const int it = it0 + lane;
    const bool valid = it < n;
    int l0 = 0, area = 0, sy = 0, sx = 0, si = 0, minx = 0x7fffffff, maxx = -1, y = 0;
    if (valid) {
      const int gw = wl[it];
      const int p00 = par[gw], rs00 = rs[gw];  // plane 0 of both, requested beside the word
      const uint32_t F = __ldg(bb + gw);
      y = gw / WW;
      const int x0 = (gw - y * WW) << 5;
      uint32_t S = F & ~(F << 1);
      for (int j = 0; S; ++j) {
        int s;
        const uint32_t R = pcs_pop_run(F, S, s);
        const int p0 = j == 0 ? p00 : par[j * NW + gw];
        const int rsi = j == 0 ? rs00 : rs[j * NW + gw];
        int l;
        if (p0 < 0) {
          l = -p0;  // a root: already ranked
        } else {
          l = -par[pcs_slot<SEG_LSPW>(p0, NW)];
          par[j * NW + gw] = -l;
        }
        const int len = __popc(R), xs = x0 + s;
        const int rsy = len * y, rsx = len * xs + len * (len - 1) / 2;
        if (j == 0) {  // the word's first run goes through the warp aggregation
          l0 = l;
          area = len, sy = rsy, sx = rsx, si = rsi, minx = xs, maxx = xs + len - 1;
        } else if (trow + l - 1 < cap) {
          seg_table_add(table, cap, trow + l - 1, len, rsy, rsx, rsi, xs, xs + len - 1, y);
        }
      }
    }
    const unsigned grp = __match_any_sync(0xffffffffu, l0);
    area = __reduce_add_sync(grp, area);
    sy = __reduce_add_sync(grp, sy);
    sx = __reduce_add_sync(grp, sx);
    si = __reduce_add_sync(grp, si);
    minx = __reduce_min_sync(grp, minx);
    maxx = __reduce_max_sync(grp, maxx);
    y = __reduce_max_sync(grp, y);
    if (l0 > 0 && lane == __ffs(grp) - 1 && trow + l0 - 1 < cap) seg_table_add(table, cap, trow + l0 - 1, area, sy, sx, si, minx, maxx, y);
  }
}

// ============================================================== host side
// The labelling stage of the pipeline: img, thr -> bits, uint8 mask, int32 labels, counts, offsets, region table.
int pcs_seg_label_stage(const uint16_t* img, const int32_t* thr, int median, uint32_t* bits, uint8_t* mask, int32_t* labels,
                        int32_t* counts, int64_t* table, int64_t cap, const PcsCclWs& ws, int* rsum, int* wlist, int* wcount, int B, int H,
                        int W, cudaStream_t st) {
  const int WW = pcs_words(W), CPR = (WW + 31) / 32;
  const long long NW = (long long)H * WW;
  PCS_REQUIRE(B <= 65535 && (H + SEG_TR - 1) / SEG_TR <= 65535, "grid too large for the tile kernel");
  cudaMemsetAsync(wcount, 0, (size_t)(B + 1) * 4, st);
  cudaMemsetAsync(ws.rootbits, 0, (size_t)B * NW * 4, st);
  cudaMemsetAsync(ws.chunk, 0, (size_t)B * H * CPR * 4, st);
  dim3 gt((WW + SEG_TW - 1) / SEG_TW, (H + SEG_TR - 1) / SEG_TR, B);
  if (median)
    PCS_LAUNCH("k_seg_threshold_tile", st, (k_seg_threshold_tile<true><<<gt, SEG_THREADS, 0, st>>>(img, thr, bits, mask, ws.parent, rsum, wlist, wcount, H, W, WW)));
  else
    PCS_LAUNCH("k_seg_threshold_tile", st, (k_seg_threshold_tile<false><<<gt, SEG_THREADS, 0, st>>>(img, thr, bits, mask, ws.parent, rsum, wlist, wcount, H, W, WW)));
  // list passes: about one thread per listed word on blob-like masks (a fifth of the words), grid-stride beyond that
  long long gx = (NW / 5 + SEG_LIST_THREADS - 1) / SEG_LIST_THREADS;
  if (gx < 1) gx = 1;
  if (gx > 1024) gx = 1024;
  dim3 gl((unsigned)gx, B);
  PCS_LAUNCH("k_seg_merge_list", st, (k_seg_merge_list<<<gl, SEG_LIST_THREADS, 0, st>>>(bits, ws.parent, wlist, wcount, H, W, WW)));
  PCS_LAUNCH("k_seg_flatten_list", st, (k_seg_flatten_list<<<gl, SEG_LIST_THREADS, 0, st>>>(bits, ws.parent, wlist, wcount, ws.rootbits, ws.chunk, H, W, WW, CPR)));
  int rc = PCS_OK;
  PCS_LAUNCH("k_seg_scan_offsets", st, (k_seg_scan_offsets<<<B, 1024, 0, st>>>(ws.chunk, counts, H * CPR, ws.offsets, wcount + B, B)));  // wcount[B]: the "blocks done" counter, zeroed with the list counts
  PCS_LAUNCH("k_seg_rank_list", st, (k_seg_rank_list<<<gl, SEG_LIST_THREADS, 0, st>>>(bits, ws.parent, wlist, wcount, ws.rootbits, ws.chunk, ws.offsets, (long long*)table, cap, H, W, WW, CPR)));
#ifdef SEG_FUSED_RELABEL_TABLE
  const int strips = (H + SEG_RROWS - 1) / SEG_RROWS;
  dim3 gr(pcs_blocks((long long)strips * CPR * 32, SEG_RL_WARPS * 32), B);
  PCS_LAUNCH("k_seg_relabel_table", st, (k_seg_relabel_table<int32_t><<<gr, SEG_RL_WARPS * 32, 0, st>>>(bits, ws.parent, rsum, ws.offsets, (long long*)table, cap, labels, H, W, WW, CPR, strips)));
#else
  PCS_LAUNCH("k_seg_label_list", st, (k_seg_label_list<<<gl, SEG_LIST_THREADS, 0, st>>>(bits, ws.parent, rsum, wlist, wcount, ws.offsets, (long long*)table, cap, H, W, WW)));
  rc = pcs_ccl_relabel_bin(bits, ws, labels, B, H, W, st);
  if (rc) return rc;
#endif
  return pcs_check_launch("segment: labelling stage");
}
