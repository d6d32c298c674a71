// K1 per-slice 65536-bin histogram of uint16 pixels + Otsu threshold.
//
// Replaces skimage.filters.threshold_otsu (no reference call site --
// refine_boundaries.py:22 imports `filters` but never calls it; the north_star
// names Otsu).  The arithmetic follows scikit-image 0.25.2
// so the threshold is bit-identical: integer images get one bin per integer
// between min and max, counts are float32, class means are float64, the
// between-class variance is float32(w1*w2) * (m1-m2)^2 and the first arg-max wins.
//
// Histogram design: 65536 x 32-bit counters (256 KB) do not fit in shared memory,
// so each CTA keeps 65536 PACKED 16-bit counters (128 KB) and processes fewer
// than 65536 pixels between flushes, which makes overflow impossible.  A flush
// touches only non-zero bins, so global atomics drop from one per pixel to one
// per distinct value per CTA.
#include <cooperative_groups.h>

#include "pcs_common.cuh"

#include "pcs.h"

#ifndef HIST_THREADS
#define HIST_THREADS 1024
#endif
#ifndef HIST_VEC_PER_THREAD
#define HIST_VEC_PER_THREAD 7  // uint4 loads (8 px each)
#endif
#ifndef HIST_MINBLOCKS
#define HIST_MINBLOCKS 1
#endif
#define HIST_PIX_PER_BLOCK (HIST_THREADS * HIST_VEC_PER_THREAD * 8)  // 57344 < 65536
#define HIST_SMEM_BYTES (32768 * 4)

__device__ __forceinline__ void hist_count8(uint32_t* sh, const uint4& q) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t a = w[j] & 0xffffu, c = w[j] >> 16;
    atomicAdd(&sh[a >> 1], 1u << ((a & 1u) << 4));
    atomicAdd(&sh[c >> 1], 1u << ((c & 1u) << 4));
  }
}

// PERSISTENT kernel: at most one CTA per SM, CTA c walks a contiguous range of the flattened (slice, tile) space
// (tile = HIST_PIX_PER_BLOCK consecutive pixels of one slice).
//  * The pixels of the next tile are requested while the current one is being counted: as soon as a vector has
//    been counted its registers receive the same vector of the next tile, so the HBM round trip hides behind the
//    remaining atomics (the first version loaded and counted vector by vector and mostly waited for loads).
//  * The packed counters are flushed only when one of them could overflow during the next tile, i.e. when
//    some counter has reached 65536 - HIST_PIX_PER_BLOCK = 8192 (a one-instruction test per word while
//    scanning shared memory), when the walk crosses into another slice, and at the end.  On micrographs the
//    fullest bin takes < 1 % of the pixels, so a CTA flushes about once per slice it touches; an image of one
//    value still flushes every tile and stays exact.
//  * Why persistent: the grid never exceeds the SM count, so every CTA is dispatched at once and the block
//    scheduler moves on to kernels of OTHER streams, which fit beside this one (1024 threads and 99 KB of shared
//    memory stay free per SM).  With more CTAs than SMs the pending ones kept other streams' kernels out and the
//    atomics-bound histogram never overlapped the bandwidth-bound kernels of the neighbouring chunk
//    (scratch/probe_r2a.py: 0.274 ms beside a 1 GiB fill vs 0.288 ms one after the other).
__global__ void __launch_bounds__(HIST_THREADS, HIST_MINBLOCKS)
    k_hist_u16(const uint16_t* __restrict__ img, uint32_t* __restrict__ hist, long long npix, int ntiles, long long total_tiles) {
  extern __shared__ uint32_t sh[];  // 32768 words, two 16-bit counters each
  const int tid = threadIdx.x;
  uint4* sh4 = reinterpret_cast<uint4*>(sh);
#pragma unroll
  for (int i = 0; i < 32768 / 4 / HIST_THREADS; ++i) sh4[tid + i * HIST_THREADS] = make_uint4(0, 0, 0, 0);
  // tiles [g0, g1) of the flattened space belong to this CTA
  const long long g0 = total_tiles * blockIdx.x / gridDim.x, g1 = total_tiles * (blockIdx.x + 1) / gridDim.x;
  // (slice, tile) of a flattened tile index; the walk below advances it incrementally (a 64-bit division per tile and
  // thread was an eighth of this kernel's instructions)
  long long cur_b = g0 / ntiles;
  int cur_t = (int)(g0 - cur_b * ntiles);
  auto tile_at = [&](long long b, int t, long long& start) {
    start = (long long)t * HIST_PIX_PER_BLOCK;
    return img + b * npix;
  };
  auto full_tile = [&](const uint16_t* src, long long start) {
    return ((((uintptr_t)src) & 15) == 0) && start + HIST_PIX_PER_BLOCK <= npix;
  };
  uint4 cur[HIST_VEC_PER_THREAD];
  if (g0 < g1) {
    long long start;
    const uint16_t* src = tile_at(cur_b, cur_t, start);
    if (full_tile(src, start)) {
      const uint4* s4 = reinterpret_cast<const uint4*>(src + start) + tid;
#pragma unroll
      for (int v = 0; v < HIST_VEC_PER_THREAD; ++v) cur[v] = __ldg(s4 + v * HIST_THREADS);
    }
  }
  __syncthreads();
  for (long long g = g0; g < g1; ++g) {
    long long start, nstart = 0;
    const uint16_t* src = tile_at(cur_b, cur_t, start);
    const uint16_t* nsrc = src;
    const bool more = g + 1 < g1;
    const long long this_b = cur_b;
    if (++cur_t == ntiles) {  // the next tile opens the next slice
      cur_t = 0;
      ++cur_b;
    }
    if (more) nsrc = tile_at(cur_b, cur_t, nstart);
    const bool have_next = more && full_tile(nsrc, nstart);
    if (full_tile(src, start)) {
      const uint4* n4 = reinterpret_cast<const uint4*>(nsrc + nstart) + tid;
#pragma unroll
      for (int v = 0; v < HIST_VEC_PER_THREAD; ++v) {
        hist_count8(sh, cur[v]);
        if (have_next) cur[v] = __ldg(n4 + v * HIST_THREADS);
      }
    } else {  // ragged tail or unaligned slice
      if (have_next) {  // the next tile (first of the next slice) still wants its vectors in flight
        const uint4* n4 = reinterpret_cast<const uint4*>(nsrc + nstart) + tid;
#pragma unroll
        for (int v = 0; v < HIST_VEC_PER_THREAD; ++v) cur[v] = __ldg(n4 + v * HIST_THREADS);
      }
      const long long end = min(npix, start + HIST_PIX_PER_BLOCK);
      for (long long i = start + tid; i < end; i += HIST_THREADS) {
        const uint32_t a = src[i];
        atomicAdd(&sh[a >> 1], 1u << ((a & 1u) << 4));
      }
    }
    __syncthreads();
    const bool last = !more || nsrc != src;  // end of this CTA's range or of the slice
    int hot = 0;                             // some counter at 8192 or above: the next tile could overflow it
    if (!last) {
#pragma unroll
      for (int i = 0; i < 32768 / 4 / HIST_THREADS; ++i) {
        const uint4 q = sh4[tid + i * HIST_THREADS];
        hot |= ((q.x | q.y | q.z | q.w) & 0xE000E000u) != 0u;
      }
    }
    if (!__syncthreads_or(hot | (int)last)) continue;
    // flush the non-zero counters into the slice's histogram and clear them
    uint32_t* gh = hist + this_b * 65536;
#pragma unroll
    for (int i = 0; i < 32768 / 4 / HIST_THREADS; ++i) {
      const int idx = tid + i * HIST_THREADS;
      const uint4 q = sh4[idx];
      if (q.x | q.y | q.z | q.w) {
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (w[j]) {
            const int bin = (idx * 4 + j) * 2;
            const uint32_t lo = w[j] & 0xffffu, hi = w[j] >> 16;
            if (lo) atomicAdd(gh + bin, lo);
            if (hi) atomicAdd(gh + bin + 1, hi);
          }
        }
        sh4[idx] = make_uint4(0, 0, 0, 0);
      }
    }
    __syncthreads();
  }
}

struct OtsuBest {
  double var;
  int idx;
};

__device__ __forceinline__ OtsuBest otsu_better(OtsuBest a, OtsuBest b) {
  // larger variance wins; ties go to the smaller bin (np.argmax returns the first)
  if (b.idx >= 0 && (a.idx < 0 || b.var > a.var || (b.var == a.var && b.idx < a.idx))) return b;
  return a;
}

// One CLUSTER of four CTAs per slice: CTA r owns bins [16384 r, 16384 (r + 1)), thread t of it the 16 bins
// from 16384 r + 16 t.  The CTAs exchange their totals and their best candidates through distributed shared
// memory (two cluster barriers), so the fp64 evaluation of the 65536 candidate thresholds is spread over four
// SMs instead of one -- with one CTA per slice a 64-slice batch used 64 of the 148 SMs and the kernel was
// bound by the divisions of a single SM.
#define OTSU_CTAS 4
#define OTSU_BINS_PER_THREAD (65536 / OTSU_CTAS / 1024)

struct OtsuCtaTotals {
  long long cnt, sum;
  int lo, hi;
};

__global__ void __cluster_dims__(OTSU_CTAS, 1, 1) __launch_bounds__(1024)
    k_otsu_u16(const uint32_t* __restrict__ hist, int32_t* __restrict__ thr, int32_t* __restrict__ minmax) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  __shared__ long long s_cnt[32], s_sum[32];
  __shared__ int s_lo[32], s_hi[32];
  __shared__ double s_var[32];
  __shared__ int s_idx[32];
  __shared__ OtsuCtaTotals s_tot;  // this CTA's totals, read by the other CTAs of the cluster
  __shared__ OtsuBest s_best;      // this CTA's best candidate, read by CTA 0
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int slice = blockIdx.x / OTSU_CTAS;
  const int bin0 = (int)rank * (65536 / OTSU_CTAS) + tid * OTSU_BINS_PER_THREAD;
  const uint32_t* h = hist + (long long)slice * 65536;
  const uint4* h4 = reinterpret_cast<const uint4*>(h + bin0);
  uint32_t c[OTSU_BINS_PER_THREAD];
#pragma unroll
  for (int i = 0; i < OTSU_BINS_PER_THREAD / 4; ++i) {
    const uint4 q = __ldg(h4 + i);
    c[4 * i] = q.x;
    c[4 * i + 1] = q.y;
    c[4 * i + 2] = q.z;
    c[4 * i + 3] = q.w;
  }
  long long cnt = 0, sum = 0;
  int lo = 65536, hi = -1;
#pragma unroll
  for (int j = 0; j < OTSU_BINS_PER_THREAD; ++j) {
    const int v = bin0 + j;
    cnt += c[j];
    sum += (long long)c[j] * v;
    if (c[j]) {
      lo = min(lo, v);
      hi = max(hi, v);
    }
  }
  // inclusive scans of (cnt, sum) across the warp, min / max of the occupied bins
  long long icnt = cnt, isum = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    long long a = __shfl_up_sync(0xffffffffu, icnt, o), s2 = __shfl_up_sync(0xffffffffu, isum, o);
    if (lane >= o) {
      icnt += a;
      isum += s2;
    }
  }
  int wlo = __reduce_min_sync(0xffffffffu, lo), whi = __reduce_max_sync(0xffffffffu, hi);
  if (lane == 31) {
    s_cnt[wid] = icnt;
    s_sum[wid] = isum;
  }
  if (lane == 0) {
    s_lo[wid] = wlo;
    s_hi[wid] = whi;
  }
  __syncthreads();
  if (wid == 0) {
    long long a = s_cnt[lane], s2 = s_sum[lane];
    int l2 = s_lo[lane], h2 = s_hi[lane];
    long long ia = a, is = s2;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long x = __shfl_up_sync(0xffffffffu, ia, o), y = __shfl_up_sync(0xffffffffu, is, o);
      if (lane >= o) {
        ia += x;
        is += y;
      }
    }
    int mlo = __reduce_min_sync(0xffffffffu, l2), mhi = __reduce_max_sync(0xffffffffu, h2);
    s_cnt[lane] = ia - a;  // exclusive warp prefixes
    s_sum[lane] = is - s2;
    if (lane == 31) {
      s_tot.cnt = ia;
      s_tot.sum = is;
      s_tot.lo = mlo;
      s_tot.hi = mhi;
    }
  }
  cluster.sync();  // every CTA's totals are in its shared memory
  long long tot_cnt = 0, tot_sum = 0, before_cnt = 0, before_sum = 0;
  int vmin = 65536, vmax = -1;
#pragma unroll
  for (unsigned r = 0; r < OTSU_CTAS; ++r) {
    const OtsuCtaTotals* t = cluster.map_shared_rank(&s_tot, r);
    const long long tc = t->cnt, ts = t->sum;
    tot_cnt += tc;
    tot_sum += ts;
    if (r < rank) {
      before_cnt += tc;
      before_sum += ts;
    }
    vmin = min(vmin, t->lo);
    vmax = max(vmax, t->hi);
  }
  long long run_cnt = before_cnt + s_cnt[wid] + (icnt - cnt);  // pixels strictly below this thread's first bin
  long long run_sum = before_sum + s_sum[wid] + (isum - sum);
  OtsuBest best{0.0, -1};
#pragma unroll
  for (int j = 0; j < OTSU_BINS_PER_THREAD; ++j) {
    const int v = bin0 + j;
    run_cnt += c[j];
    run_sum += (long long)c[j] * v;
    // an empty bin leaves both classes as they were: its variance equals the previous bin's and can
    // never be the FIRST maximum, so the (slow, fp64) evaluation is skipped for it
    if (c[j] != 0u && v >= vmin && v < vmax) {
      float w1 = (float)run_cnt, w2 = (float)(tot_cnt - run_cnt);  // exact: npix <= 2^24
      double m1 = (double)run_sum / (double)w1;
      double m2 = (double)(tot_sum - run_sum) / (double)w2;
      double d = m1 - m2;
      double var = __dmul_rn((double)__fmul_rn(w1, w2), __dmul_rn(d, d));
      best = otsu_better(best, OtsuBest{var, v});
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    OtsuBest other{__shfl_xor_sync(0xffffffffu, best.var, o), __shfl_xor_sync(0xffffffffu, best.idx, o)};
    best = otsu_better(best, other);
  }
  if (lane == 0) {
    s_var[wid] = best.var;
    s_idx[wid] = best.idx;
  }
  __syncthreads();
  if (wid == 0) {
    OtsuBest bb{s_var[lane], s_idx[lane]};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      OtsuBest other{__shfl_xor_sync(0xffffffffu, bb.var, o), __shfl_xor_sync(0xffffffffu, bb.idx, o)};
      bb = otsu_better(bb, other);
    }
    if (lane == 0) s_best = bb;
  }
  cluster.sync();  // every CTA's best candidate is in its shared memory
  if (rank == 0 && tid == 0) {
    OtsuBest bb = s_best;
    for (unsigned r = 1; r < OTSU_CTAS; ++r) bb = otsu_better(bb, *cluster.map_shared_rank(&s_best, r));
    thr[slice] = (vmin == vmax || bb.idx < 0) ? vmin : bb.idx;  // single-valued image -> that value
    if (minmax) {
      minmax[2 * slice] = vmin;
      minmax[2 * slice + 1] = vmax;
    }
  }
  cluster.sync();  // no CTA leaves while its shared memory may still be read
}

extern "C" {

size_t pcs_histogram_bytes(int B) { return (size_t)B * 65536 * 4; }

int pcs_histogram_u16(const uint16_t* img, uint32_t* hist, int B, int H, int W, void* stream) {
  PCS_REQUIRE(B >= 1 && H >= 1 && W >= 1, "empty batch or image");
  static bool attr_set[64] = {};
  if (pcs_first_use(attr_set)) cudaFuncSetAttribute(k_hist_u16, cudaFuncAttributeMaxDynamicSharedMemorySize, HIST_SMEM_BYTES);
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(hist, 0, pcs_histogram_bytes(B), st);
  long long npix = (long long)H * W;
  // one persistent CTA per SM (or fewer when there are fewer tiles), each walking a contiguous tile range
  const long long ntiles = (npix + HIST_PIX_PER_BLOCK - 1) / HIST_PIX_PER_BLOCK;
  const long long total = ntiles * B;
  int sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const unsigned grid = (unsigned)(total < sms ? total : sms);
  PCS_LAUNCH("k_hist_u16", st, k_hist_u16<<<grid, HIST_THREADS, HIST_SMEM_BYTES, st>>>(img, hist, npix, (int)ntiles, total));
  return pcs_check_launch("histogram");
}

int pcs_otsu_u16(const uint32_t* hist, int32_t* thr, int32_t* minmax, int B, int64_t npix, void* stream) {
  PCS_REQUIRE(B >= 1, "empty batch");
  PCS_REQUIRE(npix <= (1LL << 24), "Otsu parity needs at most 2^24 pixels per slice (float32 cumulative counts)");
  PCS_LAUNCH("k_otsu_u16", (cudaStream_t)stream, k_otsu_u16<<<B * OTSU_CTAS, 1024, 0, (cudaStream_t)stream>>>(hist, thr, minmax));
  return pcs_check_launch("otsu");
}

}  // extern "C"
