// K4 connected-component labelling and everything built on the same forest:
// K6 fill holes, small-object removal, component selection, K9 plateau maxima.
//
// Replaces (file:line in /root/reference):
//   skimage.measure.label            tiff_analysis.py:260, :743, :829; refine_boundaries.py:64
//   scipy.ndimage.binary_fill_holes  tiff_analysis.py:880
//   area filter on regions           tiff_analysis.py:769-773 (as skimage remove_small_objects)
//   merged_image |= (labels == v)    tiff_analysis.py:878
//   skimage.morphology.local_maxima  refine_boundaries.py:63
#include <type_traits>

#include "pcs_ccl.cuh"

#include "pcs.h"

#define PCS_CCL_THREADS 256
#define PCS_MARK (-1)

// ============================================================== workspace
size_t pcs_ccl_ws_bytes(int B, int H, int W, int with_aux) {
  size_t WW = pcs_words(W), Wp = WW * 32, CPR = (WW + 31) / 32;
  size_t n = 0;
  n += pcs_align256((size_t)B * H * Wp * 4);
  n += pcs_align256((size_t)B * H * WW * 4);
  n += pcs_align256((size_t)B * H * CPR * 4);
  n += pcs_align256((size_t)(B + 1) * 4);
  if (with_aux) n += pcs_align256((size_t)B * H * Wp * 4);
  return n;
}

int pcs_ccl_ws_carve(void* ws, size_t ws_bytes, int B, int H, int W, int with_aux, PcsCclWs* out) {
  if (ws == nullptr || ws_bytes < pcs_ccl_ws_bytes(B, H, W, with_aux)) {
    pcs_set_error("connected-component workspace too small (see pcs_ccl_workspace_bytes)");
    return PCS_ERR_WORKSPACE;
  }
  size_t WW = pcs_words(W), Wp = WW * 32, CPR = (WW + 31) / 32;
  char* p = (char*)ws;
  out->parent = (int*)p;
  p += pcs_align256((size_t)B * H * Wp * 4);
  out->rootbits = (uint32_t*)p;
  p += pcs_align256((size_t)B * H * WW * 4);
  out->chunk = (int*)p;
  p += pcs_align256((size_t)B * H * CPR * 4);
  out->offsets = (int*)p;
  p += pcs_align256((size_t)(B + 1) * 4);
  out->aux = with_aux ? (int*)p : nullptr;
  return PCS_OK;
}

// ============================================================== kernels
// ---------------------------------------------------------------- tile-local union-find
// CTA = CCL_TR rows x 32 words (1024 pixels).  Every run of the tile gets a slot in a shared
// parent array (16 slots per word, ordered like the raster), unions between runs of the same
// tile are resolved with shared-memory atomics, and each node leaves the kernel pointing at
// the global index of its tile-local root.  Only adjacencies that cross a tile edge remain for
// the global pass (k_ccl_merge_edges), so a component spanning the image costs a chain over
// tiles, not over pixels or rows.
// Tile shape: TR rows x TW words, handled by 128 threads whatever the shape -- after the staging
// phase only the non-empty words carry work, so a small CTA keeps many tiles in flight per SM.
// Binary masks have at most 16 runs per 32-pixel word, multi-valued images up to 32: both shapes
// (32 x 8 words x 16 slots, 16 x 8 words x 32 slots) hold 16 KB of shared parents.
#define CCL_TILE_THREADS 128
template <class P> struct PcsTile { static constexpr int TR = 32, TW = 8, LOG_SPW = 4; };
template <> struct PcsTile<PcsGenProv> { static constexpr int TR = 16, TW = 8, LOG_SPW = 5; };

__device__ __forceinline__ int pcs_lfind(volatile int* sp, int n) {
  int r = n, p = sp[r];
  const int first = p;
  while (p != r) {
    r = p;
    p = sp[r];
  }
  if (first != r) atomicMin((int*)sp + n, r);  // path compression (monotone, race-safe): later finds are one hop
  return r;
}

__device__ __forceinline__ void pcs_lunion(int* sp, int a, int b) {
  while (true) {
    a = pcs_lfind(sp, a);
    b = pcs_lfind(sp, b);
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

template <class P, int CONN>
__global__ void __launch_bounds__(CCL_TILE_THREADS) k_ccl_tile(P prov, int* __restrict__ parent) {
  constexpr int TR = PcsTile<P>::TR, TW = PcsTile<P>::TW, LSPW = PcsTile<P>::LOG_SPW, SPW = 1 << LSPW, NWORDS = TR * TW;
  constexpr bool kBin = std::is_same<P, PcsBinProv>::value;
  __shared__ int sp[NWORDS * SPW];
  __shared__ uint32_t fsm[NWORDS], ssm[NWORDS];
  __shared__ uint32_t gsm[kBin ? 1 : 4][NWORDS];  // multi-valued: U, UL, UR, J planes of the tile
  __shared__ unsigned short items[NWORDS];
  __shared__ int nitems;
  const int H = prov.H, WW = prov.WW;
  const int k0 = blockIdx.x * TW, y0 = blockIdx.y * TR;
  const long long b = blockIdx.z;
  P p = prov.slice(b);
  if (threadIdx.x == 0) nitems = 0;
  // phase A, thread per word: stage the tile, give every run its own slot, list the non-empty words.
  // The words of all of a thread's iterations are requested first (one memory round trip, not one each).
  constexpr int NPT = (NWORDS + CCL_TILE_THREADS - 1) / CCL_TILE_THREADS;
  uint32_t Fpre[NPT], Spre[NPT];
  uint32_t any = 0u;
#pragma unroll
  for (int i = 0; i < NPT; ++i) {
    const int w = threadIdx.x + i * CCL_TILE_THREADS;
    const int k = k0 + w % TW, y = y0 + w / TW;
    Fpre[i] = Spre[i] = 0u;
    if (w < NWORDS && y < H && k < WW) p.FS(y, k, Fpre[i], Spre[i]);
    any |= Fpre[i];
  }
  if (!__syncthreads_or(any != 0u)) return;  // empty tile: the common case on sparse planes (hole candidates)
#pragma unroll
  for (int i = 0; i < NPT; ++i) {
    const int w = threadIdx.x + i * CCL_TILE_THREADS;
    if (w >= NWORDS) break;
    const int r = w / TW, c = w % TW;
    const int k = k0 + c, y = y0 + r;
    uint32_t F = Fpre[i], S = Spre[i];
    fsm[w] = F;
    ssm[w] = S;
    if (!kBin) {
      PcsConnWords cw;
      cw.U = cw.UL = cw.UR = 0;
      cw.J = 0;
      if (F) p.conn(y, k, cw);
      gsm[0][w] = cw.U;
      gsm[kBin ? 0 : 1][w] = cw.UL;
      gsm[kBin ? 0 : 2][w] = cw.UR;
      gsm[kBin ? 0 : 3][w] = (uint32_t)cw.J;
    }
    const int sbase = w << LSPW;
    int j = 0;
    while (S) {
      S &= S - 1;
      sp[sbase + j] = sbase + j;
      ++j;
    }
    // one shared atomic per warp, not per word
    const unsigned act = __activemask();
    const unsigned has = __ballot_sync(act, F != 0u);
    if (has) {
      const int lane = threadIdx.x & 31;
      const int leader = __ffs(has) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(&nitems, __popc(has));
      base = __shfl_sync(act, base, leader);
      if (F) items[base + __popc(has & ((1u << lane) - 1u))] = (unsigned short)w;
    }
  }
  __syncthreads();
  const int n = nitems;
  if (n == 0) return;  // empty tile (uniform): nothing to unite, nothing to publish
  // phase B, thread per NON-EMPTY word (all lanes busy): unions between runs of this tile
  for (int it = threadIdx.x; it < n; it += CCL_TILE_THREADS) {
    const int w = items[it], wr = w / TW, wc = w % TW;
    const uint32_t F = fsm[w];
    uint32_t U = 0, UL = 0, UR = 0, Sa[3] = {0u, 0u, 0u};
    int J;
    if (kBin) {
      const uint32_t fl = wc > 0 ? fsm[w - 1] : 0u;
      J = (F & 1u) && (fl >> 31);
      if (wr > 0) {
        const uint32_t al = wc > 0 ? fsm[w - TW - 1] : 0u, ac = fsm[w - TW], ar = wc < TW - 1 ? fsm[w - TW + 1] : 0u;
        U = F & ac;
        UL = F & ((ac << 1) | (al >> 31));
        UR = F & ((ac >> 1) | (ar << 31));
      }
    } else {
      J = (int)gsm[kBin ? 0 : 3][w] && wc > 0;
      if (wr > 0) {
        U = gsm[0][w];
        UL = gsm[kBin ? 0 : 1][w];
        UR = gsm[kBin ? 0 : 2][w];
      }
    }
    if (wr > 0) {
      Sa[0] = wc > 0 ? ssm[w - TW - 1] : 0u;
      Sa[1] = ssm[w - TW];
      Sa[2] = wc < TW - 1 ? ssm[w - TW + 1] : 0u;
    }
    const int sbase = w << LSPW;
    if (J) pcs_lunion(sp, sbase, sbase - SPW + __popc(ssm[w - 1]) - 1);
    if (U | UL | UR) {
      uint32_t S = ssm[w];
      int j = 0;
      while (S) {
        int s;
        uint32_t R = pcs_pop_run(F, S, s);
        unsigned long long T = ((unsigned long long)(U & R)) << 1;
        if (CONN == 8) T |= (unsigned long long)(UL & R) | (((unsigned long long)(UR & R)) << 2);
        while (T) {
          int i = __ffsll((long long)T) - 1;
          T &= T + (1ull << i);
          int rel = (i - 1) >> 5;
          int ca = wc + rel;
          if (ca >= 0 && ca < TW) {  // the run above lives in this tile
            int ja = (i - 1) & 31;
            uint32_t sa_w = Sa[rel + 1];
            int sa = pcs_start_at_or_below(sa_w, ja);
            int ord = __popc(sa_w & ((1u << sa) - 1u));
            pcs_lunion(sp, sbase + j, ((w - TW + rel) << LSPW) + ord);
          }
        }
        ++j;
      }
    }
  }
  __syncthreads();
  // phase C, thread per non-empty word: every node points at the global id of its tile-local root
  // (tile slot = word-in-tile * SPW + ordinal, global id = word * SPW + ordinal: only the word changes)
  const int NW = H * WW;
  int* par = parent + b * ((long long)NW << LSPW);
  for (int it = threadIdx.x; it < n; it += CCL_TILE_THREADS) {
    const int w = items[it], wr = w / TW, wc = w % TW;
    const int sbase = w << LSPW;
    const int gw = (y0 + wr) * WW + k0 + wc;
    const int nruns = __popc(ssm[w]);
    for (int j = 0; j < nruns; ++j) {
      const int root = pcs_lfind(sp, sbase + j);
      const int rw = root >> LSPW, rj = root & (SPW - 1);  // word of the root inside the tile, run ordinal
      par[j * NW + gw] = pcs_node<LSPW>((y0 + rw / TW) * WW + k0 + rw % TW, rj);
    }
  }
}

// thread per TILE-EDGE word: the unions k_ccl_tile could not do -- adjacencies across a tile edge.
// Edge words per slice: the top row of every tile (rows y % TR == 0, all words) followed by the
// first / last word column of every tile in the remaining rows.
template <class P, int CONN>
__global__ void __launch_bounds__(PCS_CCL_THREADS) k_ccl_merge_edges(P prov, int* __restrict__ parent, int B, int per_slice) {
  const int H = prov.H, WW = prov.WW;
  constexpr int CCL_TR = PcsTile<P>::TR, TW = PcsTile<P>::TW;
  int e = blockIdx.x * blockDim.x + threadIdx.x;  // edge word of the slice; the slice is blockIdx.y
  if (e >= per_slice) return;
  const long long b = blockIdx.y;
  const int ntop = (H + CCL_TR - 1) / CCL_TR;
  int y, k;
  if (e < ntop * WW) {
    y = (e / WW) * CCL_TR;
    k = e % WW;
  } else {
    e -= ntop * WW;
    const int ncol = 2 * ((WW + TW - 1) / TW);  // first and last word of every tile column
    const int rr = e / ncol, cc = e % ncol;      // rr counts the non-top rows
    y = rr + rr / (CCL_TR - 1) + 1;              // skip rows that are multiples of TR
    k = (cc >> 1) * TW + ((cc & 1) ? TW - 1 : 0);
    if (y >= H || k >= WW) return;
  }
  const bool top = (y % CCL_TR) == 0, left = (k % TW) == 0, right = (k % TW) == TW - 1;
  P p = prov.slice(b);
  uint32_t F0, S0;
  p.FS(y, k, F0, S0);
  if (!F0) return;
  PcsConnWords c;
  p.conn(y, k, c);
  constexpr int LSPW = PcsNodes<P>::LOG_SPW;
  const int NW = H * WW;
  int* par = parent + b * ((long long)NW << LSPW);
  const int gw = y * WW + k;
  if (c.J && left) {
    uint32_t Sl = p.Sword(y, k - 1);
    pcs_uf_union<LSPW>(par, NW, pcs_node<LSPW>(gw, 0), pcs_node<LSPW>(gw - 1, __popc(Sl) - 1));  // last run of the word to the left
  }
  if (y == 0 || !(c.U | c.UL | c.UR)) return;
  uint32_t S = c.S;
  for (int j = 0; S; ++j) {  // j: ordinal of the run
    int s;
    uint32_t R = pcs_pop_run(c.F, S, s);
    unsigned long long T = ((unsigned long long)(c.U & R)) << 1;
    if (CONN == 8) T |= (unsigned long long)(c.UL & R) | (((unsigned long long)(c.UR & R)) << 2);
    while (T) {
      int i = __ffsll((long long)T) - 1;
      T &= T + (1ull << i);
      int rel = (i - 1) >> 5;
      if (!(top || (rel < 0 && left) || (rel > 0 && right))) continue;  // done inside the tile
      int ja = (i - 1) & 31;
      const uint32_t Sa = c.Sa[rel + 1];
      int sa = pcs_start_at_or_below(Sa, ja);
      pcs_uf_union<LSPW>(par, NW, pcs_node<LSPW>(gw, j), pcs_node<LSPW>(gw - WW + rel, pcs_run_ord(Sa, sa)));
    }
  }
}

// warp per PAIR of 32-word chunks: point every node at its root, flag roots, count them.  A lane walks
// the parent chains of its two words side by side, so two dependent load chains are in flight per lane
// (the kernel is bound by that pointer chase).
template <class P>
__global__ void __launch_bounds__(PCS_CCL_THREADS)
    k_ccl_flatten(P prov, int* __restrict__ parent, uint32_t* __restrict__ rootbits, int* __restrict__ chunk,
                  int* __restrict__ aux, int B, int CPR) {
  const int H = prov.H, WW = prov.WW;
  const int pair = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // pair of chunks of the slice; the slice is blockIdx.y
  int lane = threadIdx.x & 31;
  const int nchunks = H * CPR;
  if (2 * pair >= nchunks) return;
  const long long b = blockIdx.y;
  P p = prov.slice(b);
  constexpr int LSPW = PcsNodes<P>::LOG_SPW;
  const int NW = H * WW;
  int* par = parent + b * ((long long)NW << LSPW);
  uint32_t roots[2] = {0u, 0u};
  int gw[2] = {0, 0}, nr[2] = {0, 0}, kk[2], yy[2];
  bool valid[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int g32 = 2 * pair + u;
    valid[u] = g32 < nchunks;
    const int ch = g32 % CPR;
    yy[u] = g32 / CPR;
    kk[u] = ch * 32 + lane;
    if (valid[u] && kk[u] < WW) {
      uint32_t F, S;
      p.FS(yy[u], kk[u], F, S);
      nr[u] = __popc(S);
      gw[u] = yy[u] * WW + kk[u];
    }
  }
  for (int j = 0; j < nr[0] || j < nr[1]; ++j) {  // j: run ordinal, both words side by side
    int n[2], r[2], q[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      n[u] = j < nr[u] ? pcs_node<LSPW>(gw[u], j) : -1;
      r[u] = n[u];
      q[u] = n[u] >= 0 ? pcs_ld_cg(par + j * NW + gw[u]) : -1;
    }
    while (q[0] != r[0] || q[1] != r[1]) {
#pragma unroll
      for (int u = 0; u < 2; ++u)
        if (q[u] != r[u]) {
          r[u] = q[u];
          q[u] = pcs_ld_cg(par + pcs_slot<LSPW>(r[u], NW));
        }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (n[u] < 0) continue;
      if (r[u] != n[u])
        par[j * NW + gw[u]] = r[u];
      else {
        roots[u] |= 1u << j;
        if (aux) aux[b * ((long long)NW << LSPW) + j * NW + gw[u]] = 0;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    if (!valid[u]) continue;  // warp-uniform
    if (kk[u] < WW) rootbits[(b * H + yy[u]) * (long long)WW + kk[u]] = roots[u];
    const int cnt = __reduce_add_sync(0xffffffffu, __popc(roots[u]));
    if (lane == 0) chunk[b * (long long)nchunks + 2 * pair + u] = cnt;
  }
}

// block per slice: exclusive scan of the chunk counts (raster order), slice total
__global__ void __launch_bounds__(1024) k_ccl_scan(int* __restrict__ chunk, int32_t* __restrict__ counts, int n) {
  __shared__ int wsum[32];
  __shared__ int carry_s;
  int* c = chunk + (long long)blockIdx.x * n;
  int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {
    int i = base + tid;
    int v = i < n ? c[i] : 0;
    int carry = carry_s;  // stable here: last written before the previous iteration's barriers
    int tot;
    int ex = pcs_warp_excl_scan(v, lane, &tot);
    if (lane == 0) wsum[wid] = tot;
    __syncthreads();
    if (wid == 0) {
      int w = lane < (int)(blockDim.x >> 5) ? wsum[lane] : 0;
      int wt;
      int wex = pcs_warp_excl_scan(w, lane, &wt);
      wsum[lane] = wex;
      if (lane == 0) carry_s = carry + wt;
    }
    __syncthreads();
    if (i < n) c[i] = carry + wsum[wid] + ex;
    __syncthreads();
  }
  if (tid == 0) counts[blockIdx.x] = carry_s;
}

// one block: exclusive scan of the per-slice counts -> table row offsets
__global__ void __launch_bounds__(1024) k_ccl_offsets(const int32_t* __restrict__ counts, int* __restrict__ offsets, int B) {
  __shared__ int wsum[32];
  __shared__ int carry_s;
  int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < B; base += blockDim.x) {
    int i = base + tid;
    int v = i < B ? counts[i] : 0;
    int carry = carry_s;
    int tot;
    int ex = pcs_warp_excl_scan(v, lane, &tot);
    if (lane == 0) wsum[wid] = tot;
    __syncthreads();
    if (wid == 0) {
      int w = wsum[lane];
      int wt;
      int wex = pcs_warp_excl_scan(w, lane, &wt);
      wsum[lane] = wex;
      if (lane == 0) carry_s = carry + wt;
    }
    __syncthreads();
    if (i < B) offsets[i] = carry + wsum[wid] + ex;
    __syncthreads();
  }
  if (tid == 0) offsets[B] = carry_s;
}

// warp per chunk: roots get their raster-order rank (stored negated in parent),
// and the table's first-pixel column if requested
template <class P>
__global__ void __launch_bounds__(PCS_CCL_THREADS)
    k_ccl_rank(P prov, int* __restrict__ parent, const uint32_t* __restrict__ rootbits, const int* __restrict__ chunk,
               const int* __restrict__ offsets, long long* __restrict__ first_out, long long cap, int B, int CPR) {
  const int H = prov.H, WW = prov.WW, W = prov.W;
  const int g32 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // 32-word chunk of the slice; the slice is blockIdx.y
  int lane = threadIdx.x & 31;
  if (g32 >= H * CPR) return;
  const long long b = blockIdx.y;
  const long long g = b * H * CPR + g32;
  const int ch = g32 % CPR, y = g32 / CPR;
  int k = ch * 32 + lane;
  uint32_t roots = k < WW ? rootbits[(b * H + y) * (long long)WW + k] : 0u;  // bit j: run of ordinal j is a root
  const int cbase = chunk[g];  // requested beside the root bits, not after the scan
  int tot;
  int ex = pcs_warp_excl_scan(__popc(roots), lane, &tot);
  if (!roots) return;
  int rank = cbase + ex;
  constexpr int LSPW = PcsNodes<P>::LOG_SPW;
  const int NW = H * WW;
  int* par = parent + b * ((long long)NW << LSPW);
  const int gw = y * WW + k;
  long long trow = first_out ? (long long)offsets[b] : 0;
  uint32_t S = 0;
  if (first_out) {
    uint32_t F;
    prov.slice(b).FS(y, k, F, S);
  }
  while (roots) {
    int j = __ffs(roots) - 1;
    roots &= roots - 1;
    ++rank;
    par[j * NW + gw] = -rank;
    if (first_out) {
      long long row = trow + rank - 1;
      uint32_t sb = S;  // start bit of the run of ordinal j: drop the j lowest start bits
      for (int q = 0; q < j; ++q) sb &= sb - 1;
      if (row < cap) first_out[row] = (long long)y * W + (k << 5) + (__ffs(sb) - 1);
    }
  }
}

// warp per chunk: expand run labels to pixels with coalesced 128-byte stores
template <class P, typename OutT>
__global__ void __launch_bounds__(PCS_CCL_THREADS)
    k_ccl_relabel(P prov, const int* __restrict__ parent, OutT* __restrict__ out, int B, int CPR) {
  __shared__ int lab[PCS_CCL_THREADS / 32][32][33];
  const int H = prov.H, WW = prov.WW, W = prov.W;
  const int g32 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // 32-word chunk of the slice; the slice is blockIdx.y
  int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  if (g32 >= H * CPR) return;
  const long long b = blockIdx.y;
  const long long g = b * H * CPR + g32;
  const int ch = g32 % CPR, y = g32 / CPR;
  int k = ch * 32 + lane;
  uint32_t F = 0, S = 0;
  int one = 0;  // label of the word's only run (words with several runs go through shared memory)
  constexpr int LSPW = PcsNodes<P>::LOG_SPW;
  const int NW = H * WW;
  const int* par = parent + b * ((long long)NW << LSPW);
  if (k < WW) {
    P p = prov.slice(b);
    const int gw = y * WW + k;
    // plane 0 (the first run of the word) is requested beside the word itself: one coalesced line per 32 words, and the
    // word -> parent chain loses a round trip (what the plane holds for an empty word is never used)
    const int p0pre = par[gw];
    p.FS(y, k, F, S);
    uint32_t rem = S;
    if (rem && !(rem & (rem - 1))) {
      const int p0 = p0pre;  // plane 0: the word's only run
      one = p0 < 0 ? -p0 : -par[pcs_slot<LSPW>(p0, NW)];
      rem = 0;
    }
    for (int j = 0; rem; j += 4) {  // four independent lookups in flight per iteration; j: ordinal of the first
      int s0 = __ffs(rem) - 1;
      rem &= rem - 1;
      int s1 = rem ? __ffs(rem) - 1 : -1;
      rem &= rem - 1 + (rem == 0);
      int s2 = rem ? __ffs(rem) - 1 : -1;
      rem &= rem - 1 + (rem == 0);
      int s3 = rem ? __ffs(rem) - 1 : -1;
      rem &= rem - 1 + (rem == 0);
      int p0 = par[j * NW + gw];
      int p1 = s1 >= 0 ? par[(j + 1) * NW + gw] : -1;
      int p2 = s2 >= 0 ? par[(j + 2) * NW + gw] : -1;
      int p3 = s3 >= 0 ? par[(j + 3) * NW + gw] : -1;
      int l0 = p0 < 0 ? -p0 : -par[pcs_slot<LSPW>(p0, NW)];
      int l1 = p1 < 0 ? -p1 : -par[pcs_slot<LSPW>(p1, NW)];
      int l2 = p2 < 0 ? -p2 : -par[pcs_slot<LSPW>(p2, NW)];
      int l3 = p3 < 0 ? -p3 : -par[pcs_slot<LSPW>(p3, NW)];
      lab[wl][lane][s0] = l0;
      if (s1 >= 0) lab[wl][lane][s1] = l1;
      if (s2 >= 0) lab[wl][lane][s2] = l2;
      if (s3 >= 0) lab[wl][lane][s3] = l3;
    }
  }
  __syncwarp();
  OutT* orow = out + (b * H + y) * (long long)W;
  int nk = min(32, WW - ch * 32);
  if (__ballot_sync(0xffffffffu, F != 0u) == 0u && sizeof(OutT) == 4 && (W & 3) == 0) {
    // no foreground in these 32 words: plain zero fill
    for (int k4 = 0; k4 < nk; k4 += 4) {
      const int kk = k4 + (lane >> 3);
      const int x = ((ch * 32 + kk) << 5) + ((lane & 7) << 2);
      if (kk < nk && x < W) *reinterpret_cast<int4*>(orow + x) = make_int4(0, 0, 0, 0);
    }
  } else if (sizeof(OutT) == 4 && (W & 3) == 0) {
    // 16-byte stores: a lane owns 4 consecutive pixels, 4 words (128 pixels) per iteration
    const int sub = lane >> 3, nib = (lane & 7) << 2;
    for (int k4 = 0; k4 < nk; k4 += 4) {
      const int kk = k4 + sub;
      const uint32_t f = __shfl_sync(0xffffffffu, F, kk & 31);
      const uint32_t s = __shfl_sync(0xffffffffu, S, kk & 31);
      const int l1 = __shfl_sync(0xffffffffu, one, kk & 31);
      int4 v = make_int4(0, 0, 0, 0);
      const uint32_t n = kk < nk ? (f >> nib) & 0xfu : 0u;
      if (n) {
        if (l1) {  // single-run word: every foreground pixel carries the same label
          v.x = (n & 1u) ? l1 : 0;
          v.y = (n & 2u) ? l1 : 0;
          v.z = (n & 4u) ? l1 : 0;
          v.w = (n & 8u) ? l1 : 0;
        } else {
          if (n & 1u) v.x = lab[wl][kk][pcs_start_at_or_below(s, nib)];
          if (n & 2u) v.y = lab[wl][kk][pcs_start_at_or_below(s, nib + 1)];
          if (n & 4u) v.z = lab[wl][kk][pcs_start_at_or_below(s, nib + 2)];
          if (n & 8u) v.w = lab[wl][kk][pcs_start_at_or_below(s, nib + 3)];
        }
      }
      const int x = ((ch * 32 + kk) << 5) + nib;
      if (kk < nk && x < W) *reinterpret_cast<int4*>(orow + x) = v;  // W % 4 == 0: x + 3 < W too
    }
  } else {
    for (int kk = 0; kk < nk; ++kk) {
      uint32_t f = __shfl_sync(0xffffffffu, F, kk);
      uint32_t s = __shfl_sync(0xffffffffu, S, kk);
      const int l1 = __shfl_sync(0xffffffffu, one, kk);
      int x = ((ch * 32 + kk) << 5) + lane;
      int v = 0;
      if ((f >> lane) & 1u) v = l1 ? l1 : lab[wl][kk][pcs_start_at_or_below(s, lane)];
      if (x < W) orow[x] = (OutT)v;
    }
  }
}

// thread per word: mark the roots of runs selected by `mode`
//   mode 0: runs touching the image border       (fill holes)
//   mode 1: runs intersecting the bit plane `m`  (seeds / higher-neighbour flags)
template <class P>
__global__ void __launch_bounds__(PCS_CCL_THREADS)
    k_ccl_mark(P prov, int* __restrict__ parent, const uint32_t* __restrict__ m, int mode, int B) {
  // thread per group of 4 words of a row: the four word loads are independent, and on the sparse planes
  // this kernel usually sees (hole candidates, seeds) most groups leave right after them
  const int H = prov.H, WW = prov.WW, W = prov.W;
  const int WG = (WW + 3) >> 2;
  const int g32 = blockIdx.x * blockDim.x + threadIdx.x;
  if (g32 >= H * WG) return;
  const long long b = blockIdx.y;
  const int y = g32 / WG, k0 = (g32 % WG) << 2;
  P p = prov.slice(b);
  uint32_t F4[4], S4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    F4[i] = S4[i] = 0u;
    if (k0 + i < WW) p.FS(y, k0 + i, F4[i], S4[i]);
  }
  if (!(S4[0] | S4[1] | S4[2] | S4[3])) return;
  constexpr int LSPW = PcsNodes<P>::LOG_SPW;
  const int NW = H * WW;
  int* par = parent + b * ((long long)NW << LSPW);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t F = F4[i], S = S4[i];
    if (!S) continue;
    const int k = k0 + i;
    uint32_t M;
    if (mode == 0) {
      M = (y == 0 || y == H - 1) ? 0xffffffffu : 0u;
      if (k == 0) M |= 1u;
      if (k == WW - 1) M |= 1u << ((W - 1) & 31);
    } else {
      M = m[(b * H + y) * (long long)WW + k];
    }
    M &= F;
    if (!M) continue;
    const int gw = y * WW + k;
    for (int j = 0; S; ++j) {  // j: ordinal of the run
      int s;
      uint32_t R = pcs_pop_run(F, S, s);
      if (!(R & M)) continue;
      int q = par[j * NW + gw];
      if (q < 0) continue;                      // already a marked root
      par[pcs_slot<LSPW>(q, NW)] = PCS_MARK;    // q is the root (its own id for a root)
    }
  }
}

// thread per word: accumulate run lengths at the roots
template <class P>
__global__ void __launch_bounds__(PCS_CCL_THREADS) k_ccl_area(P prov, const int* __restrict__ parent, int* __restrict__ aux, int B) {
  const int H = prov.H, WW = prov.WW;
  const int t32 = blockIdx.x * blockDim.x + threadIdx.x;  // word of the slice; the slice is blockIdx.y
  if (t32 >= H * WW) return;
  const long long b = blockIdx.y;
  const long long t = b * H * WW + t32;
  const int k = t32 % WW, y = t32 / WW;
  P p = prov.slice(b);
  uint32_t F, S;
  p.FS(y, k, F, S);
  if (!S) return;
  constexpr int LSPW = PcsNodes<P>::LOG_SPW;
  const int NW = H * WW;
  const int* par = parent + b * ((long long)NW << LSPW);
  int* ax = aux + b * ((long long)NW << LSPW);
  const int gw = y * WW + k;
  for (int j = 0; S; ++j) {
    int s;
    uint32_t R = pcs_pop_run(F, S, s);
    atomicAdd(ax + pcs_slot<LSPW>(par[j * NW + gw], NW), __popc(R));  // after the flatten every node holds its root's id
  }
}

// thread per word: mark roots whose accumulated area is below min_size
template <int LSPW>
__global__ void __launch_bounds__(PCS_CCL_THREADS)
    k_ccl_mark_small(int* __restrict__ parent, const uint32_t* __restrict__ rootbits, const int* __restrict__ aux,
                     int min_size, int B, int H, int WW) {
  const int t32 = blockIdx.x * blockDim.x + threadIdx.x;  // word of the slice
  const int NW = H * WW;
  if (t32 >= NW) return;
  const long long b = blockIdx.y;
  uint32_t roots = rootbits[b * NW + t32];  // bit j: run of ordinal j is a root
  if (!roots) return;
  const long long sb = b * ((long long)NW << LSPW);
  while (roots) {
    int j = __ffs(roots) - 1;
    roots &= roots - 1;
    if (aux[sb + j * NW + t32] < min_size) parent[sb + j * NW + t32] = PCS_MARK;
  }
}

// thread per word: out = runs whose root is marked (want=1) / unmarked (want=0),
// optionally OR-ed with another bit plane; all-or-nothing veto by counts==1
template <class P>
__global__ void __launch_bounds__(PCS_CCL_THREADS)
    k_ccl_select(P prov, const int* __restrict__ parent, int want_marked, const uint32_t* __restrict__ or_bits,
                 const int32_t* __restrict__ veto_counts, uint32_t* __restrict__ out, uint8_t* __restrict__ mask, int B) {
  const int H = prov.H, WW = prov.WW;
  const int t32 = blockIdx.x * blockDim.x + threadIdx.x;  // word of the slice; the slice is blockIdx.y
  if (t32 >= H * WW) return;
  const long long b = blockIdx.y;
  const long long t = b * H * WW + t32;
  const int k = t32 % WW, y = t32 / WW;
  P p = prov.slice(b);
  const uint32_t ob = or_bits ? __ldg(or_bits + t) : 0u;  // requested beside the plane's own word, not after the parent walk
  uint32_t F, S;
  p.FS(y, k, F, S);
  uint32_t o = 0;
  if (S && !(veto_counts && veto_counts[b] == 1)) {
    constexpr int LSPW = PcsNodes<P>::LOG_SPW;
    const int NW = H * WW;
    const int* par = parent + b * ((long long)NW << LSPW);
    const int gw = y * WW + k;
    for (int j = 0; S; ++j) {
      int s;
      uint32_t R = pcs_pop_run(F, S, s);
      int q = par[j * NW + gw];
      int marked = q < 0 ? 1 : (q == pcs_node<LSPW>(gw, j) ? 0 : (par[pcs_slot<LSPW>(q, NW)] < 0));
      if (marked == want_marked) o |= R;
    }
  }
  o |= ob;
  out[t] = o;
  if (mask) pcs_store_mask_bytes(mask + (b * H + y) * (long long)prov.W, k, prov.W, o);  // fused uint8 output
}

// ---------------------------------------------------------------- hole filling inside bounding boxes
// A hole of an (8-connected) component lies inside that component's bounding box, so background
// outside every box is open by construction and never needs labelling.  k_bbox_raster paints the
// boxes of the kept components (area >= min_size) from the region table; k_hole_candidates turns
// them into the candidate mask C (background inside a box) and the seed mask M (candidates on the
// image border or 4-adjacent to open background).  Holes = 4-connected components of C without a seed.
__global__ void __launch_bounds__(256)
    k_bbox_raster(const long long* __restrict__ table, long long cap, const int* __restrict__ offsets, long long min_size,
                  uint32_t* __restrict__ bb, int B, int H, int WW) {
  long long g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // warp per table row
  const int lane = threadIdx.x & 31;
  const long long nrows = min((long long)offsets[B], cap);
  if (g >= nrows) return;
  if (table[0 * cap + g] < min_size) return;  // column 0: area
  int lo = 0, hi = B;                          // slice of this row: offsets[b] <= g < offsets[b+1]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if ((long long)offsets[mid] <= g) lo = mid; else hi = mid;
  }
  const int y0 = (int)table[3 * cap + g], x0 = (int)table[4 * cap + g];
  const int y1 = (int)table[5 * cap + g], x1 = (int)table[6 * cap + g];  // inclusive
  const int k0 = x0 >> 5, k1 = x1 >> 5;
  uint32_t* base = bb + (long long)lo * H * WW;
  const int nw = k1 - k0 + 1;
  const long long cells = (long long)(y1 - y0 + 1) * nw;
  for (long long i = lane; i < cells; i += 32) {
    const int y = y0 + (int)(i / nw), k = k0 + (int)(i % nw);
    uint32_t m = 0xffffffffu;
    if (k == k0) m &= 0xffffffffu << (x0 & 31);
    if (k == k1) m &= 0xffffffffu >> (31 - (x1 & 31));
    atomicOr(base + (long long)y * WW + k, m);
  }
}

__global__ void __launch_bounds__(256)
    k_hole_candidates(const uint32_t* __restrict__ fbits, const uint32_t* __restrict__ bb, uint32_t* __restrict__ cand,
                      uint32_t* __restrict__ seed, int B, int H, int W, int WW) {
  const int t32 = blockIdx.x * blockDim.x + threadIdx.x;
  if (t32 >= H * WW) return;
  const long long t = (long long)blockIdx.y * H * WW + t32;
  const int k = t32 % WW, y = t32 / WW;
  const uint32_t vm = pcs_valid_mask(k, W);
  const uint32_t box = bb[t];
  const uint32_t C = ~fbits[t] & box & vm;
  cand[t] = C;
  uint32_t M = 0;
  if (C) {
    // open background = background outside every box; outside the image counts as open too
    auto open_at = [&](long long idx, int kk) { return ~fbits[idx] & ~bb[idx] & pcs_valid_mask(kk, W); };
    const uint32_t n_c = ~fbits[t] & ~box & vm;
    uint32_t adj = (n_c << 1) | (n_c >> 1);
    adj |= (k > 0) ? (open_at(t - 1, k - 1) >> 31) : 1u;
    if (k + 1 < WW) adj |= open_at(t + 1, k + 1) << 31;
    adj |= (y > 0) ? open_at(t - WW, k) : 0xffffffffu;
    adj |= (y + 1 < H) ? open_at(t + WW, k) : 0xffffffffu;
    adj |= 1u << ((W - 1) & 31) & ((k == WW - 1) ? 0xffffffffu : 0u);  // right image border
    M = C & adj;
  }
  seed[t] = M;
}


// ---------------------------------------------------------------- refine: small objects out + hole candidates by row spans
// Hole filling when the foreground is already labelled (the pipeline's case).  The outer boundary
// of a hole is an 8-connected curve of ONE component, so every hole pixel has pixels of that
// component both to its left and to its right in its own row (possibly beyond islands that sit in
// the hole).  Background outside every same-label row span can therefore never be part of a hole.
// The candidate set C = background between two runs of one label in a row is tiny for blob-like
// masks (convex blobs have none).  Holes = 4-connected components of C that neither touch the image
// border nor touch background outside C.  Any superset of C gives the same answer, which the
// bounded search below relies on.
//
// k_refine_rows: warp per row.  Drops the runs of components below min_size (area from the region
// table: remove_small_objects), lists the kept runs of the row in shared memory, pairs every run
// with the previous run of the same label (looking back at most REFINE_K runs; if the window is
// exhausted the whole row prefix becomes candidate, a safe superset) and paints the spans: partial
// words with shared atomics, whole words through a difference array + prefix sum.
#define REFINE_MAX_WARPS 8
#define REFINE_K 48
#define REFINE_RMAX 256  // kept runs listed per row; a row with more becomes candidate as a whole (safe superset)
#define REFINE_SMEM_LIMIT (200 * 1024)
static inline size_t refine_words_per_warp(int WW) { return (size_t)WW * 3 + 1 + 2 * REFINE_RMAX; }

__device__ __forceinline__ void refine_paint(uint32_t* cb, int* diff, int a, int e) {
  const int wa = a >> 5, we = e >> 5;
  const uint32_t ma = 0xffffffffu << (a & 31), me = 0xffffffffu >> (31 - (e & 31));
  if (wa == we) {
    atomicOr(cb + wa, ma & me);
  } else {
    atomicOr(cb + wa, ma);
    atomicOr(cb + we, me);
    if (we - wa > 1) {
      atomicAdd(diff + wa + 1, 1);
      atomicAdd(diff + we, -1);
    }
  }
}

// PLANES = false: run labels are read from the int32 label image (the drop-in primitive).
// PLANES = true (the pipeline): run labels are read, negated, from the labeller's parent planes, where
// k_seg_relabel_table left them -- a coalesced load of the dense plane 0 for every word with one run instead of a
// scattered 4-byte read of a 16 MiB label image; candidate words are appended to the slice's list and their runs get
// fresh singleton parents in `hpar`, so the hole stage that follows only walks that list.
struct PcsRefineSparse {
  const int* labpar;  // parent planes holding -label for every run
  int* hpar;          // parent planes of the hole forest
  int* clist;         // per slice: candidate words
  int* ccount;        // per slice: how many
  uint8_t* mask;      // uint8 copy of the kept words (the refined mask before the holes are added), optional
};

template <bool PLANES>
__global__ void __launch_bounds__(REFINE_MAX_WARPS * 32)
    k_refine_rows(const uint32_t* __restrict__ fg, const int32_t* __restrict__ labels, const long long* __restrict__ table,
                  long long cap, const int* __restrict__ offsets, long long min_size, uint32_t* __restrict__ kept_out,
                  uint32_t* __restrict__ cand_out, long long rows, int H, int W, int WW, PcsRefineSparse sp) {
  extern __shared__ uint32_t sm[];
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + wl;
  if (row >= rows) return;
  uint32_t* kw = sm + (size_t)wl * ((size_t)WW * 3 + 1 + 2 * REFINE_RMAX);  // kept words of the row
  uint32_t* cb = kw + WW;                                                     // painted partial words
  int* diff = (int*)(cb + WW);                                                // WW + 1: whole-word coverage, differenced
  int* rl = diff + WW + 1;                                                    // labels of the kept runs, row order
  uint32_t* rse = (uint32_t*)(rl + REFINE_RMAX);                              // start | end << 16 of the kept runs
  const long long b = (long long)((unsigned)row / (unsigned)H);  // rows = B * H <= 65535 * 16384 < 2^31: 32-bit division
  const int y = (int)(row - b * H);
  const int NW = H * WW;
  const long long tbase = offsets ? (long long)offsets[b] : 0;
  const int32_t* lrow = PLANES ? nullptr : labels + row * (long long)W;
  const int* lpl = PLANES ? sp.labpar + b * ((long long)NW << 4) + (long long)y * WW : nullptr;  // plane j of word k: lpl[j * NW + k]
  const uint32_t* frow = fg + row * (long long)WW;
  int n = 0;
  bool overflow = false;
  // Four 32-word groups at a time: the words, then the label of every word's first run, then its area are requested
  // for all four before any is used, so a row waits for three memory round trips per 128 words (the word -> label ->
  // area chain, once per group, was a third of this kernel's stall samples).
  for (int kb = 0; kb < WW; kb += 128) {
    uint32_t f4[4];
    int L4[4];
    long long A4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = kb + 32 * u + lane;
      f4[u] = k < WW ? __ldg(frow + k) : 0u;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = kb + 32 * u + lane;
      L4[u] = 0;
      if (PLANES) {
        // plane 0 holds the label of the word's first run: requested for every word of the row beside the word itself
        // (one coalesced line per 32 words either way; what it holds for an empty word is never used), so the
        // word -> label -> area chain is two round trips, not three
        if (k < WW) L4[u] = -lpl[k];
      } else if (f4[u]) {
        L4[u] = __ldg(lrow + (k << 5) + __ffs(f4[u]) - 1);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      A4[u] = 0;
      if (min_size > 1 && f4[u]) {
        const long long r = tbase + L4[u] - 1;
        A4[u] = (r < 0 || r >= cap) ? 0 : __ldg(table + r);  // column 0 of the table: area
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k0 = kb + 32 * u;
      if (k0 >= WW) break;  // warp-uniform
      const int k = k0 + lane;
      const uint32_t f = f4[u];
      uint32_t kept = f;
      int cnt = __popc(f & ~(f << 1));
      if (min_size > 1) {
        kept = 0u;
        cnt = 0;
        uint32_t S = f & ~(f << 1);
        for (int j = 0; S; ++j) {
          const int s = __ffs(S) - 1;
          S &= S - 1;
          const uint32_t upper = ~(f >> s);
          const int len = upper ? (__ffs(upper) - 1) : 32;
          long long area = A4[u];
          if (j > 0) {
            const long long r = tbase + (PLANES ? -lpl[j * NW + k] : __ldg(lrow + (k << 5) + s)) - 1;
            area = (r < 0 || r >= cap) ? 0 : __ldg(table + r);
          }
          if (area < min_size) continue;
          kept |= (len >= 32 ? 0xffffffffu : ((1u << len) - 1u)) << s;
          ++cnt;
        }
      }
      int tot;
      int at = n + pcs_warp_excl_scan(cnt, lane, &tot);
      n += tot;
      overflow |= n > REFINE_RMAX;
      if (!overflow) {
        uint32_t S = kept & ~(kept << 1);
        const uint32_t Sf = f & ~(f << 1);  // the ordinal of a kept run counts the dropped runs of the word too
        while (S) {
          const int s = __ffs(S) - 1;
          S &= S - 1;
          const uint32_t upper = ~(kept >> s);
          const int len = upper ? (__ffs(upper) - 1) : 32;
          const uint32_t xs = (uint32_t)((k << 5) + s);
          const int ord = pcs_run_ord(Sf, s);
          rl[at] = ord == 0 ? L4[u] : (PLANES ? -lpl[ord * NW + k] : __ldg(lrow + xs));
          rse[at] = xs | ((xs + len - 1) << 16);
          ++at;
        }
      }
      if (k < WW) {
        kw[k] = kept;
        kept_out[row * (long long)WW + k] = kept;
        if (PLANES && sp.mask) pcs_store_mask_bytes(sp.mask + row * (long long)W, k, W, kept);  // holes are patched in by k_hole_apply_list
      }
    }
  }
  __syncwarp();
  // a candidate word leaves for the plane and, in the pipeline, for the slice's list with singleton parents
  auto emit = [&](int k, uint32_t c) {
    if (k < WW) cand_out[row * (long long)WW + k] = c;
    if (PLANES) {
      const unsigned has = __ballot_sync(0xffffffffu, k < WW && c != 0u);
      if (has) {
        const int leader = __ffs(has) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(sp.ccount + b, __popc(has));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (k < WW && c != 0u) {
          const int gw = y * WW + k;
          sp.clist[b * (long long)NW + base + __popc(has & ((1u << lane) - 1u))] = gw;
          int* hp = sp.hpar + b * ((long long)NW << 4);
          const int nr = __popc(c & ~(c << 1));
          for (int j = 0; j < nr; ++j) hp[j * NW + gw] = pcs_node<4>(gw, j);
        }
      }
    }
  };
  if (overflow) {
    for (int k0 = 0; k0 < WW; k0 += 32) {
      const int k = k0 + lane;
      emit(k, k < WW ? ~kw[k] & pcs_valid_mask(k, W) : 0u);
    }
    return;
  }
  // Most rows of a blob-like mask have no gap between two runs of one label at all: the pairing is searched once
  // without painting, and such rows leave with an all-zero candidate row -- no scratch to clear, no prefix scan.
  int fb = -1;  // end of the row prefix painted when a search window ran out
  bool pair = false;
  // Up to 32 kept runs (the usual row has about ten): one __match_any_sync over the labels gives every run the lanes
  // that hold the same label, the nearest one below is its partner -- no backward walk (which cost each lane as many
  // steps as there were runs before it, and the warp as many as the row had runs)
  const bool small = n <= 32;
  int jm = -1;  // partner of run `lane` when small
  if (small) {
    const int L = lane < n ? rl[lane] : -(lane + 1);  // idle lanes: distinct negative sentinels (labels are positive)
    const unsigned same = __match_any_sync(0xffffffffu, L) & ((1u << lane) - 1u);
    jm = same ? 31 - __clz(same) : -1;
    if (lane < n && jm >= 0) pair = (int)(rse[jm] >> 16) + 1 <= (int)(rse[lane] & 0xffffu) - 1;
  } else {
    for (int i = lane; i < n; i += 32) {
      const int L = rl[i];
      const int lo = max(i - REFINE_K, 0);
      int j = i - 1;
      while (j >= lo && rl[j] != L) --j;
      const int e = (int)(rse[i] & 0xffffu) - 1;
      if (j >= lo) {
        pair |= (int)(rse[j] >> 16) + 1 <= e;
      } else if (lo > 0) {
        fb = max(fb, e);
      }
    }
  }
  if (!__any_sync(0xffffffffu, pair || fb >= 0)) {
    for (int k = lane; k < WW; k += 32) cand_out[row * (long long)WW + k] = 0u;
    return;
  }
  for (int k = lane; k < WW; k += 32) {
    cb[k] = 0u;
    diff[k] = 0;
  }
  if (lane == 0) diff[WW] = 0;
  __syncwarp();
  if (small) {
    if (lane < n && jm >= 0) {
      const int a = (int)(rse[jm] >> 16) + 1, e = (int)(rse[lane] & 0xffffu) - 1;
      if (a <= e) refine_paint(cb, diff, a, e);
    }
  } else {
    for (int i = lane; i < n; i += 32) {
      const int L = rl[i];
      const int lo = max(i - REFINE_K, 0);
      int j = i - 1;
      while (j >= lo && rl[j] != L) --j;
      const int e = (int)(rse[i] & 0xffffu) - 1;
      if (j >= lo) {
        const int a = (int)(rse[j] >> 16) + 1;
        if (a <= e) refine_paint(cb, diff, a, e);
      }
    }
  }
  fb = __reduce_max_sync(0xffffffffu, fb);
  if (lane == 0 && fb >= 0) refine_paint(cb, diff, 0, fb);
  __syncwarp();
  int carry = 0;
  for (int k0 = 0; k0 < WW; k0 += 32) {
    const int k = k0 + lane;
    const int d = k < WW ? diff[k] : 0;
    int tot;
    const int cov = carry + pcs_warp_excl_scan(d, lane, &tot) + d;
    carry += tot;
    emit(k, k < WW ? (cb[k] | (cov > 0 ? 0xffffffffu : 0u)) & ~kw[k] & pcs_valid_mask(k, W) : 0u);
  }
}

// ---------------------------------------------------------------- hole stage over the candidate list (pipeline)
// grid = (GX, B), thread per listed candidate word (grid-stride).  Candidates are few (blob-like masks have next to
// none), so the three passes below cost a launch each instead of the five dense passes of the generic path.
#define HOLE_LIST_THREADS 128

// 4-connected unions between candidate runs: with the word to the left and with the row above
__device__ __forceinline__ void hole_union_pass(const uint32_t* __restrict__ cc, int* __restrict__ par, const int* __restrict__ cl, int n,
                                                int NW, int WW, int first, int stride) {
  for (int it = first; it < n; it += stride) {
    const int gw = cl[it], y = gw / WW, k = gw - y * WW;
    const uint32_t C = __ldg(cc + gw);
    if (k > 0 && (C & 1u)) {
      const uint32_t Cl = __ldg(cc + gw - 1);
      if (Cl >> 31) pcs_uf_union<4>(par, NW, pcs_node<4>(gw, 0), pcs_node<4>(gw - 1, __popc(Cl & ~(Cl << 1)) - 1));
    }
    if (y == 0) continue;
    const uint32_t ac = __ldg(cc + gw - WW);
    if (!(C & ac)) continue;
    const uint32_t Sa = ac & ~(ac << 1);
    uint32_t S = C & ~(C << 1);
    for (int j = 0; S; ++j) {
      int s;
      const uint32_t R = pcs_pop_run(C, S, s);
      unsigned long long T = (unsigned long long)(R & ac);  // every stretch of set bits lies in one run of the row above
      while (T) {
        const int i = __ffsll((long long)T) - 1;
        T &= T + (1ull << i);
        pcs_uf_union<4>(par, NW, pcs_node<4>(gw, j), pcs_node<4>(gw - WW, pcs_run_ord(Sa, pcs_start_at_or_below(Sa, i))));
      }
    }
  }
}

__device__ __forceinline__ int pcs_hole_root(const int* par, int NW, int n) {  // root id of n; roots may carry PCS_MARK
  int r = n;
  while (true) {
    const int p = pcs_ld_cg(par + pcs_slot<4>(r, NW));
    if (p < 0 || p == r) return r;
    r = p;
  }
}

// candidate runs that touch the image border or background outside the candidate set are open: mark their roots
__device__ __forceinline__ void hole_mark_pass(const uint32_t* __restrict__ kk, const uint32_t* __restrict__ cc, int* __restrict__ par,
                                               const int* __restrict__ cl, int n, int NW, int H, int W, int WW, int first, int stride) {
  for (int it = first; it < n; it += stride) {
    const int gw = cl[it], y = gw / WW, k = gw - y * WW;
    const uint32_t C = __ldg(cc + gw);
    auto open_at = [&](int w, int kw) { return ~kk[w] & ~__ldg(cc + w) & pcs_valid_mask(kw, W); };  // kk is the refined plane this kernel patches later: plain loads
    const uint32_t n_c = open_at(gw, k);
    uint32_t adj = (n_c << 1) | (n_c >> 1);
    adj |= (k > 0) ? (open_at(gw - 1, k - 1) >> 31) : 1u;
    if (k + 1 < WW) adj |= open_at(gw + 1, k + 1) << 31;
    adj |= (y > 0) ? open_at(gw - WW, k) : 0xffffffffu;
    adj |= (y + 1 < H) ? open_at(gw + WW, k) : 0xffffffffu;
    adj |= 1u << ((W - 1) & 31) & ((k == WW - 1) ? 0xffffffffu : 0u);  // right image border
    const uint32_t M = C & adj;
    if (!M) continue;
    uint32_t S = C & ~(C << 1);
    for (int j = 0; S; ++j) {
      int s;
      const uint32_t R = pcs_pop_run(C, S, s);
      if (!(R & M)) continue;
      const int r = pcs_hole_root(par, NW, pcs_node<4>(gw, j));
      par[pcs_slot<4>(r, NW)] = PCS_MARK;
    }
  }
}

// thread per listed candidate word: the candidate runs whose root is unmarked are closed -- holes -- and are added to
// the refined word and its uint8 copy (which k_refine_rows wrote without them)
__device__ __forceinline__ void hole_apply_pass(const uint32_t* __restrict__ cand, const int* __restrict__ par, const int* __restrict__ cl,
                                                int n, long long b, uint32_t* __restrict__ out, uint8_t* __restrict__ mask, int NW, int H, int W,
                                                int WW, int first, int stride) {
  for (int it = first; it < n; it += stride) {
    const int gw = cl[it];
    const long long t = b * NW + gw;
    const uint32_t C = cand[t];
    uint32_t add = 0;
    uint32_t S = C & ~(C << 1);
    for (int j = 0; S; ++j) {
      int s;
      const uint32_t R = pcs_pop_run(C, S, s);
      const int r = pcs_hole_root(par, NW, pcs_node<4>(gw, j));
      if (pcs_ld_cg(par + pcs_slot<4>(r, NW)) >= 0) add |= R;
    }
    if (!add) continue;
    const uint32_t o = out[t] | add;
    out[t] = o;
    if (mask) {
      const int y = gw / WW, k = gw - y * WW;
      pcs_store_mask_bytes(mask + (b * H + y) * (long long)W, k, W, o);
    }
  }
}

// CTA per slice: the three passes of the hole stage -- unions between candidate runs, marking of the components that
// touch open background, patching of the closed ones into the refined words -- with block barriers in between.
// Slices are independent problems, so one CTA per slice needs no grid-wide synchronisation, and with the usual handful
// of candidates per slice the three passes cost one launch instead of three (each was launch latency only).  A slice
// with very many candidates just takes more trips through the block-stride loops.
__global__ void __launch_bounds__(1024)
    k_hole_resolve_list(const uint32_t* __restrict__ cand, int* __restrict__ hpar, const int* __restrict__ clist,
                        const int* __restrict__ ccount, uint32_t* __restrict__ out, uint8_t* __restrict__ mask, int H, int W, int WW) {
  const long long b = blockIdx.x;
  const int NW = H * WW;
  const int n = min(ccount[b], NW);
  if (n == 0) return;
  const uint32_t* cc = cand + b * (long long)NW;
  int* par = hpar + b * ((long long)NW << 4);
  const int* cl = clist + b * (long long)NW;
  hole_union_pass(cc, par, cl, n, NW, WW, threadIdx.x, blockDim.x);
  __syncthreads();
  hole_mark_pass(out + b * (long long)NW, cc, par, cl, n, NW, H, W, WW, threadIdx.x, blockDim.x);
  __syncthreads();
  hole_apply_pass(cand, par, cl, n, b, out, mask, NW, H, W, WW, threadIdx.x, blockDim.x);
}

// thread per group of 4 words: seeds = candidate pixels on the top / bottom image row or 4-adjacent to
// background that is not a candidate (such background is never part of a hole).  Only words that hold
// candidates are written: k_ccl_mark reads the seed plane nowhere else.
__global__ void __launch_bounds__(256)
    k_hole_seeds(const uint32_t* __restrict__ kept, const uint32_t* __restrict__ cand, uint32_t* __restrict__ seed, int B, int H,
                 int W, int WW) {
  const int WG = (WW + 3) >> 2;
  const int g32 = blockIdx.x * blockDim.x + threadIdx.x;
  if (g32 >= H * WG) return;
  const int y = g32 / WG, k0 = (g32 % WG) << 2;
  const long long t0 = ((long long)blockIdx.y * H + y) * WW + k0;
  uint32_t C4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) C4[i] = k0 + i < WW ? __ldg(cand + t0 + i) : 0u;
  if (!(C4[0] | C4[1] | C4[2] | C4[3])) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t C = C4[i];
    if (!C) continue;
    const int k = k0 + i;
    const long long t = t0 + i;
    auto open_at = [&](long long idx, int kk) { return ~kept[idx] & ~cand[idx] & pcs_valid_mask(kk, W); };
    const uint32_t n_c = open_at(t, k);
    uint32_t adj = (n_c << 1) | (n_c >> 1);
    adj |= (k > 0) ? (open_at(t - 1, k - 1) >> 31) : 1u;
    if (k + 1 < WW) adj |= open_at(t + 1, k + 1) << 31;
    adj |= (y > 0) ? open_at(t - WW, k) : 0xffffffffu;
    adj |= (y + 1 < H) ? open_at(t + WW, k) : 0xffffffffu;
    adj |= 1u << ((W - 1) & 31) & ((k == WW - 1) ? 0xffffffffu : 0u);  // right image border
    seed[t] = C & adj;
  }
}

// ============================================================== host drivers
template <class P>
static int ccl_forest(const P& prov, int B, int conn, const PcsCclWs& ws, int32_t* counts, int zero_aux, cudaStream_t st) {
  const int H = prov.H, WW = prov.WW;
  const int CPR = (WW + 31) / 32;
  dim3 gw(pcs_blocks((long long)H * WW, PCS_CCL_THREADS), B);
  dim3 gq(pcs_blocks((long long)H * ((WW + 3) / 4), PCS_CCL_THREADS), B);  // groups of 4 words
  dim3 gc(pcs_blocks((long long)H * CPR * 32, PCS_CCL_THREADS), B);
  constexpr int CCL_TR = PcsTile<P>::TR, TW = PcsTile<P>::TW;
  PCS_REQUIRE(B <= 65535 && (H + CCL_TR - 1) / CCL_TR <= 65535, "grid too large for the tile kernel");
  dim3 gt((WW + TW - 1) / TW, (H + CCL_TR - 1) / CCL_TR, B);
  const int ntop = (H + CCL_TR - 1) / CCL_TR;
  const int per_slice = ntop * WW + (H - ntop) * 2 * ((WW + TW - 1) / TW);  // tile-edge words of one slice
  dim3 ge(pcs_blocks(per_slice, PCS_CCL_THREADS), B);
  if (conn == 8) {
    PCS_LAUNCH("k_ccl_tile", st, (k_ccl_tile<P, 8><<<gt, CCL_TILE_THREADS, 0, st>>>(prov, ws.parent)));
    PCS_LAUNCH("k_ccl_merge_edges", st, (k_ccl_merge_edges<P, 8><<<ge, PCS_CCL_THREADS, 0, st>>>(prov, ws.parent, B, per_slice)));
  } else {
    PCS_LAUNCH("k_ccl_tile", st, (k_ccl_tile<P, 4><<<gt, CCL_TILE_THREADS, 0, st>>>(prov, ws.parent)));
    PCS_LAUNCH("k_ccl_merge_edges", st, (k_ccl_merge_edges<P, 4><<<ge, PCS_CCL_THREADS, 0, st>>>(prov, ws.parent, B, per_slice)));
  }
  PCS_LAUNCH("k_ccl_flatten", st, k_ccl_flatten<P><<<dim3(pcs_blocks(((long long)H * CPR + 1) / 2 * 32, PCS_CCL_THREADS), B), PCS_CCL_THREADS, 0, st>>>(prov, ws.parent, ws.rootbits, ws.chunk, zero_aux ? ws.aux : nullptr, B, CPR));
  if (counts) {
    PCS_LAUNCH("k_ccl_scan", st, k_ccl_scan<<<B, 1024, 0, st>>>(ws.chunk, counts, H * CPR));
    PCS_LAUNCH("k_ccl_offsets", st, k_ccl_offsets<<<1, 1024, 0, st>>>(counts, ws.offsets, B));
  }
  return pcs_check_launch("ccl forest");
}

template <class P>
static int ccl_label(const P& prov, int B, int conn, void* labels, int label_bytes, int32_t* counts, int32_t* offsets_out,
                     long long* first_out, long long cap, void* wsp, size_t ws_bytes, cudaStream_t st) {
  PCS_REQUIRE(conn == 4 || conn == 8, "connectivity must be 4 or 8");
  PCS_REQUIRE(label_bytes == 4 || label_bytes == 8, "label dtype must be int32 or int64");
  PCS_REQUIRE(counts != nullptr && labels != nullptr, "null output");
  PcsCclWs ws;
  int rc = pcs_ccl_ws_carve(wsp, ws_bytes, B, prov.H, prov.W, 0, &ws);
  if (rc) return rc;
  rc = ccl_forest(prov, B, conn, ws, counts, 0, st);
  if (rc) return rc;
  const int H = prov.H, WW = prov.WW, CPR = (WW + 31) / 32;
  dim3 gc(pcs_blocks((long long)H * CPR * 32, PCS_CCL_THREADS), B);
  PCS_LAUNCH("k_ccl_rank", st, k_ccl_rank<P><<<gc, PCS_CCL_THREADS, 0, st>>>(prov, ws.parent, ws.rootbits, ws.chunk, ws.offsets, first_out, cap, B, CPR));
  if (label_bytes == 4)
    PCS_LAUNCH("k_ccl_relabel", st, k_ccl_relabel<P, int32_t><<<gc, PCS_CCL_THREADS, 0, st>>>(prov, ws.parent, (int32_t*)labels, B, CPR));
  else
    PCS_LAUNCH("k_ccl_relabel", st, k_ccl_relabel<P, long long><<<gc, PCS_CCL_THREADS, 0, st>>>(prov, ws.parent, (long long*)labels, B, CPR));
  if (offsets_out) cudaMemcpyAsync(offsets_out, ws.offsets, (size_t)(B + 1) * 4, cudaMemcpyDeviceToDevice, st);
  return pcs_check_launch("ccl label");
}

int pcs_ccl_relabel_bin(const uint32_t* bits, const PcsCclWs& ws, int32_t* labels, int B, int H, int W, cudaStream_t st) {
  PcsBinProv prov{bits, H, W, pcs_words(W), 0};
  const int CPR = (prov.WW + 31) / 32;
  dim3 gc(pcs_blocks((long long)H * CPR * 32, PCS_CCL_THREADS), B);
  PCS_LAUNCH("k_ccl_relabel", st, (k_ccl_relabel<PcsBinProv, int32_t><<<gc, PCS_CCL_THREADS, 0, st>>>(prov, ws.parent, labels, B, CPR)));
  return pcs_check_launch("ccl relabel");
}

int pcs_ccl_scan_offsets(const PcsCclWs& ws, int32_t* counts, int B, int H, int W, cudaStream_t st) {
  const int CPR = (pcs_words(W) + 31) / 32;
  PCS_LAUNCH("k_ccl_scan", st, k_ccl_scan<<<B, 1024, 0, st>>>(ws.chunk, counts, H * CPR));
  PCS_LAUNCH("k_ccl_offsets", st, k_ccl_offsets<<<1, 1024, 0, st>>>(counts, ws.offsets, B));
  return pcs_check_launch("ccl scan");
}

// The refine stage of the pipeline (remove_small_objects + binary_fill_holes, tiff_analysis.py:769-773, :880) on the
// labelled mask: run labels from the parent planes, hole candidates resolved over their list.  cand: bit
// plane of scratch; hpar: parent planes of scratch (B * 16 * NW ints); clist / ccount: B * NW / B ints.
int pcs_seg_refine_stage(const uint32_t* bits, const int* labpar, const int64_t* table, int64_t cap, const int32_t* offsets,
                         int64_t min_size, uint32_t* out, uint8_t* out_mask, uint32_t* cand, int* hpar, int* clist, int* ccount, int B,
                         int H, int W, cudaStream_t st) {
  const int WW = pcs_words(W);
  const long long NW = (long long)H * WW, rows = (long long)B * H;
  const size_t warp_bytes = refine_words_per_warp(WW) * 4;
  int warps = (int)(REFINE_SMEM_LIMIT / warp_bytes);
  PCS_REQUIRE(warps >= 1, "row too wide for the refine kernel");
  if (warps > REFINE_MAX_WARPS) warps = REFINE_MAX_WARPS;
  static bool attr_set[64] = {};
  if (pcs_first_use(attr_set)) cudaFuncSetAttribute(k_refine_rows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, REFINE_SMEM_LIMIT);
  cudaMemsetAsync(ccount, 0, (size_t)B * 4, st);
  // the kept words go straight to the refined bit plane and its uint8 copy; the hole stage reads that plane as "kept" and
  // patches the few words that gain hole pixels
  PCS_LAUNCH("k_refine_rows", st, (k_refine_rows<true><<<pcs_blocks(rows, warps), warps * 32, warps * warp_bytes, st>>>(
      bits, nullptr, (const long long*)table, cap, offsets, min_size > 1 ? min_size : 1, out, cand, rows, H, W, WW,
      PcsRefineSparse{labpar, hpar, clist, ccount, out_mask})));
  PCS_LAUNCH("k_hole_resolve_list", st, (k_hole_resolve_list<<<B, 1024, 0, st>>>(cand, hpar, clist, ccount, out, out_mask, H, W, WW)));
  return pcs_check_launch("segment: refine stage");
}

static int check_dims(int B, int H, int W) {
  PCS_REQUIRE(B >= 1 && H >= 1 && W >= 1, "empty batch or image");
  PCS_REQUIRE(H <= 16384 && W <= 16384, "image side above 16384 is not supported");
  PCS_REQUIRE(B <= 65535, "batch above 65535 slices");
  return PCS_OK;
}

extern "C" {

size_t pcs_ccl_workspace_bytes(int B, int H, int W, int with_aux) { return pcs_ccl_ws_bytes(B, H, W, with_aux); }

int pcs_label_bits(const uint32_t* bits, int B, int H, int W, int connectivity, int invert, void* labels, int label_bytes,
                   int32_t* counts, int32_t* offsets, int64_t* first_out, int64_t cap, void* ws, size_t ws_bytes,
                   void* stream) {
  int rc = check_dims(B, H, W);
  if (rc) return rc;
  PcsBinProv prov{bits, H, W, pcs_words(W), invert};
  return ccl_label(prov, B, connectivity, labels, label_bytes, counts, offsets, (long long*)first_out, cap, ws, ws_bytes,
                   (cudaStream_t)stream);
}

int pcs_label_conn(const uint32_t* planes, int B, int H, int W, int connectivity, void* labels, int label_bytes,
                   int32_t* counts, int32_t* offsets, int64_t* first_out, int64_t cap, void* ws, size_t ws_bytes,
                   void* stream) {
  int rc = check_dims(B, H, W);
  if (rc) return rc;
  PcsGenProv prov{planes, (long long)B * H * pcs_words(W), H, W, pcs_words(W)};
  return ccl_label(prov, B, connectivity, labels, label_bytes, counts, offsets, (long long*)first_out, cap, ws, ws_bytes,
                   (cudaStream_t)stream);
}

int pcs_fill_holes_bits(const uint32_t* bits, uint32_t* out, int B, int H, int W, void* wsp, size_t ws_bytes, void* stream) {
  int rc = check_dims(B, H, W);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  PcsCclWs ws;
  rc = pcs_ccl_ws_carve(wsp, ws_bytes, B, H, W, 0, &ws);
  if (rc) return rc;
  PcsBinProv prov{bits, H, W, pcs_words(W), 1};  // background, 4-connected (tiff_analysis.py:880)
  rc = ccl_forest(prov, B, 4, ws, nullptr, 0, st);
  if (rc) return rc;
  dim3 gw(pcs_blocks((long long)H * prov.WW, PCS_CCL_THREADS), B);
  dim3 gq(pcs_blocks((long long)H * ((prov.WW + 3) / 4), PCS_CCL_THREADS), B);  // groups of 4 words
  PCS_LAUNCH("k_ccl_mark", st, k_ccl_mark<PcsBinProv><<<gq, PCS_CCL_THREADS, 0, st>>>(prov, ws.parent, nullptr, 0, B));
  PCS_LAUNCH("k_ccl_select", st, k_ccl_select<PcsBinProv><<<gw, PCS_CCL_THREADS, 0, st>>>(prov, ws.parent, 0, bits, nullptr, out, nullptr, B));
  return pcs_check_launch("fill holes");
}

int pcs_fill_holes_table_bits(const uint32_t* bits, const int64_t* table, int64_t cap, const int32_t* offsets, int64_t min_size,
                              uint32_t* out, uint8_t* out_mask, int B, int H, int W, void* wsp, size_t ws_bytes, void* stream) {
  int rc = check_dims(B, H, W);
  if (rc) return rc;
  PCS_REQUIRE(bits && table && offsets && out, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int WW = pcs_words(W);
  const size_t plane = pcs_align256((size_t)B * H * WW * 4);
  const size_t need = pcs_ccl_ws_bytes(B, H, W, 0) + 3 * plane;
  if (wsp == nullptr || ws_bytes < need) {
    pcs_set_error("fill-holes workspace too small (see pcs_fill_holes_table_workspace_bytes)");
    return PCS_ERR_WORKSPACE;
  }
  PcsCclWs ws;
  rc = pcs_ccl_ws_carve(wsp, ws_bytes, B, H, W, 0, &ws);
  if (rc) return rc;
  char* extra = (char*)wsp + pcs_ccl_ws_bytes(B, H, W, 0);
  uint32_t* bb = (uint32_t*)extra;
  uint32_t* cand = (uint32_t*)(extra + plane);
  uint32_t* seed = (uint32_t*)(extra + 2 * plane);
  cudaMemsetAsync(bb, 0, (size_t)B * H * WW * 4, st);
  PCS_LAUNCH("k_bbox_raster", st, k_bbox_raster<<<pcs_blocks(cap * 32, 256), 256, 0, st>>>((const long long*)table, cap, offsets, min_size, bb, B, H, WW));
  dim3 gw(pcs_blocks((long long)H * WW, PCS_CCL_THREADS), B);
  dim3 gq(pcs_blocks((long long)H * ((WW + 3) / 4), PCS_CCL_THREADS), B);  // groups of 4 words
  PCS_LAUNCH("k_hole_candidates", st, k_hole_candidates<<<gw, 256, 0, st>>>(bits, bb, cand, seed, B, H, W, WW));
  PcsBinProv prov{cand, H, W, WW, 0};
  rc = ccl_forest(prov, B, 4, ws, nullptr, 0, st);
  if (rc) return rc;
  PCS_LAUNCH("k_ccl_mark", st, (k_ccl_mark<PcsBinProv><<<gq, PCS_CCL_THREADS, 0, st>>>(prov, ws.parent, seed, 1, B)));
  PCS_LAUNCH("k_ccl_select", st, (k_ccl_select<PcsBinProv><<<gw, PCS_CCL_THREADS, 0, st>>>(prov, ws.parent, 0, bits, nullptr, out, out_mask, B)));
  return pcs_check_launch("fill holes (table)");
}

int pcs_refine_labeled_bits(const uint32_t* bits, const int32_t* labels, const int64_t* table, int64_t cap, const int32_t* offsets,
                            int64_t min_size, uint32_t* out, uint8_t* out_mask, int B, int H, int W, void* wsp, size_t ws_bytes,
                            void* stream) {
  int rc = check_dims(B, H, W);
  if (rc) return rc;
  PCS_REQUIRE(bits && labels && out, "null argument");
  PCS_REQUIRE(min_size <= 1 || (table && offsets && cap >= 1), "min_size > 1 needs the region table");
  cudaStream_t st = (cudaStream_t)stream;
  const int WW = pcs_words(W);
  const size_t plane = pcs_align256((size_t)B * H * WW * 4);
  if (wsp == nullptr || ws_bytes < pcs_ccl_ws_bytes(B, H, W, 0) + 3 * plane) {
    pcs_set_error("refine workspace too small (see pcs_fill_holes_table_workspace_bytes)");
    return PCS_ERR_WORKSPACE;
  }
  PcsCclWs ws;
  rc = pcs_ccl_ws_carve(wsp, ws_bytes, B, H, W, 0, &ws);
  if (rc) return rc;
  char* extra = (char*)wsp + pcs_ccl_ws_bytes(B, H, W, 0);
  uint32_t* kept = (uint32_t*)extra;
  uint32_t* cand = (uint32_t*)(extra + plane);
  uint32_t* seed = (uint32_t*)(extra + 2 * plane);
  const long long rows = (long long)B * H;
  const size_t warp_bytes = refine_words_per_warp(WW) * 4;
  int warps = (int)(REFINE_SMEM_LIMIT / warp_bytes);
  PCS_REQUIRE(warps >= 1, "row too wide for the refine kernel");
  if (warps > REFINE_MAX_WARPS) warps = REFINE_MAX_WARPS;
  static bool attr_set[64] = {};
  if (pcs_first_use(attr_set)) cudaFuncSetAttribute(k_refine_rows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, REFINE_SMEM_LIMIT);
  PCS_LAUNCH("k_refine_rows", st, (k_refine_rows<false><<<pcs_blocks(rows, warps), warps * 32, warps * warp_bytes, st>>>(
      bits, labels, (const long long*)table, cap, offsets, min_size, kept, cand, rows, H, W, WW, PcsRefineSparse{})));
  dim3 gw(pcs_blocks((long long)H * WW, PCS_CCL_THREADS), B);
  dim3 gq(pcs_blocks((long long)H * ((WW + 3) / 4), PCS_CCL_THREADS), B);  // groups of 4 words
  PCS_LAUNCH("k_hole_seeds", st, (k_hole_seeds<<<dim3(pcs_blocks((long long)H * ((WW + 3) / 4), 256), B), 256, 0, st>>>(kept, cand, seed, B, H, W, WW)));
  PcsBinProv prov{cand, H, W, WW, 0};
  rc = ccl_forest(prov, B, 4, ws, nullptr, 0, st);
  if (rc) return rc;
  PCS_LAUNCH("k_ccl_mark", st, (k_ccl_mark<PcsBinProv><<<gq, PCS_CCL_THREADS, 0, st>>>(prov, ws.parent, seed, 1, B)));
  PCS_LAUNCH("k_ccl_select", st, (k_ccl_select<PcsBinProv><<<gw, PCS_CCL_THREADS, 0, st>>>(prov, ws.parent, 0, kept, nullptr, out, out_mask, B)));
  return pcs_check_launch("refine (labelled)");
}

size_t pcs_fill_holes_table_workspace_bytes(int B, int H, int W) {
  return pcs_ccl_ws_bytes(B, H, W, 0) + 3 * pcs_align256((size_t)B * H * pcs_words(W) * 4);
}

int pcs_remove_small_bits(const uint32_t* bits, uint32_t* out, int B, int H, int W, int connectivity, int min_size,
                          void* wsp, size_t ws_bytes, void* stream) {
  int rc = check_dims(B, H, W);
  if (rc) return rc;
  PCS_REQUIRE(connectivity == 4 || connectivity == 8, "connectivity must be 4 or 8");
  cudaStream_t st = (cudaStream_t)stream;
  PcsCclWs ws;
  rc = pcs_ccl_ws_carve(wsp, ws_bytes, B, H, W, 1, &ws);
  if (rc) return rc;
  PcsBinProv prov{bits, H, W, pcs_words(W), 0};
  rc = ccl_forest(prov, B, connectivity, ws, nullptr, 1, st);
  if (rc) return rc;
  dim3 gw(pcs_blocks((long long)H * prov.WW, PCS_CCL_THREADS), B);
  dim3 gq(pcs_blocks((long long)H * ((prov.WW + 3) / 4), PCS_CCL_THREADS), B);  // groups of 4 words
  PCS_LAUNCH("k_ccl_area", st, k_ccl_area<PcsBinProv><<<gw, PCS_CCL_THREADS, 0, st>>>(prov, ws.parent, ws.aux, B));
  PCS_LAUNCH("k_ccl_mark_small", st, k_ccl_mark_small<PcsNodes<PcsBinProv>::LOG_SPW><<<gw, PCS_CCL_THREADS, 0, st>>>(ws.parent, ws.rootbits, ws.aux, min_size, B, H, prov.WW));
  PCS_LAUNCH("k_ccl_select", st, k_ccl_select<PcsBinProv><<<gw, PCS_CCL_THREADS, 0, st>>>(prov, ws.parent, 0, nullptr, nullptr, out, nullptr, B));
  return pcs_check_launch("remove small objects");
}

int pcs_select_components_bits(const uint32_t* bits, const uint32_t* seeds, uint32_t* out, int B, int H, int W,
                               int connectivity, void* wsp, size_t ws_bytes, void* stream) {
  int rc = check_dims(B, H, W);
  if (rc) return rc;
  PCS_REQUIRE(connectivity == 4 || connectivity == 8, "connectivity must be 4 or 8");
  cudaStream_t st = (cudaStream_t)stream;
  PcsCclWs ws;
  rc = pcs_ccl_ws_carve(wsp, ws_bytes, B, H, W, 0, &ws);
  if (rc) return rc;
  PcsBinProv prov{bits, H, W, pcs_words(W), 0};
  rc = ccl_forest(prov, B, connectivity, ws, nullptr, 0, st);
  if (rc) return rc;
  dim3 gw(pcs_blocks((long long)H * prov.WW, PCS_CCL_THREADS), B);
  dim3 gq(pcs_blocks((long long)H * ((prov.WW + 3) / 4), PCS_CCL_THREADS), B);  // groups of 4 words
  PCS_LAUNCH("k_ccl_mark", st, k_ccl_mark<PcsBinProv><<<gq, PCS_CCL_THREADS, 0, st>>>(prov, ws.parent, seeds, 1, B));
  PCS_LAUNCH("k_ccl_select", st, k_ccl_select<PcsBinProv><<<gw, PCS_CCL_THREADS, 0, st>>>(prov, ws.parent, 1, nullptr, nullptr, out, nullptr, B));
  return pcs_check_launch("select components");
}

int pcs_local_maxima_conn(const uint32_t* planes, const uint32_t* higher, uint32_t* out, int32_t* counts, int B, int H,
                          int W, int connectivity, void* wsp, size_t ws_bytes, void* stream) {
  int rc = check_dims(B, H, W);
  if (rc) return rc;
  PCS_REQUIRE(connectivity == 4 || connectivity == 8, "connectivity must be 4 or 8");
  PCS_REQUIRE(counts != nullptr, "null counts");
  cudaStream_t st = (cudaStream_t)stream;
  PcsCclWs ws;
  rc = pcs_ccl_ws_carve(wsp, ws_bytes, B, H, W, 0, &ws);
  if (rc) return rc;
  PcsGenProv prov{planes, (long long)B * H * pcs_words(W), H, W, pcs_words(W)};
  rc = ccl_forest(prov, B, connectivity, ws, counts, 0, st);
  if (rc) return rc;
  dim3 gw(pcs_blocks((long long)H * prov.WW, PCS_CCL_THREADS), B);
  dim3 gq(pcs_blocks((long long)H * ((prov.WW + 3) / 4), PCS_CCL_THREADS), B);  // groups of 4 words
  PCS_LAUNCH("k_ccl_mark", st, k_ccl_mark<PcsGenProv><<<gq, PCS_CCL_THREADS, 0, st>>>(prov, ws.parent, higher, 1, B));
  // a plateau that is the whole image (counts == 1) is not a maximum
  PCS_LAUNCH("k_ccl_select", st, k_ccl_select<PcsGenProv><<<gw, PCS_CCL_THREADS, 0, st>>>(prov, ws.parent, 0, nullptr, counts, out, nullptr, B));
  return pcs_check_launch("local maxima");
}

}  // extern "C"
