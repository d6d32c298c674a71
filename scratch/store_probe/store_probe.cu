// Store-bandwidth probe: how fast can a kernel write 2 GiB of HBM on this GPU, by store flavour?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_st16(uint4* p, size_t n16) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n16; i += stride) p[i] = make_uint4(1, 2, 3, 4);
}
__global__ void k_st16_cs(uint4* p, size_t n16) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n16; i += stride) __stcs(p + i, make_uint4(1, 2, 3, 4));
}
__global__ void k_st16_wt(uint4* p, size_t n16) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n16; i += stride) __stwt(p + i, make_uint4(1, 2, 3, 4));
}
__global__ void k_st32(uint4* p, size_t n32) {  // 256-bit stores
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n32; i += stride) {
    unsigned long long a = (unsigned long long)(p + 2 * i);
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(a), "l"(1ull), "l"(2ull), "l"(3ull), "l"(4ull) : "memory");
  }
}
// TMA bulk store: every CTA fills a 32 KB shared buffer once, then streams it out in 32 KB bulk copies
__global__ void k_bulk(char* p, size_t bytes) {
  extern __shared__ __align__(128) char sm[];
  const int CH = 32768;
  for (int i = threadIdx.x * 16; i < CH; i += blockDim.x * 16) *reinterpret_cast<uint4*>(sm + i) = make_uint4(1, 2, 3, 4);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(sm);
    for (size_t off = (size_t)blockIdx.x * CH; off + CH <= bytes; off += (size_t)gridDim.x * CH) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + off), "r"(s), "r"(CH) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

template <class F>
static void timeit(const char* name, size_t bytes, F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9f;
  for (int it = 0; it < 6; ++it) {
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (it && ms < best) best = ms;
  }
  cudaError_t err = cudaGetLastError();
  printf("%-28s %8.3f ms  %7.1f GB/s  %s\n", name, best, bytes / best / 1e6, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
  const size_t bytes = 2ull << 30;
  char* p;
  cudaMalloc(&p, bytes);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int mult : {4, 8, 16}) {
    const int grid = sms * mult;
    printf("grid = %d x SMs\n", mult);
    timeit("st.128", bytes, [&] { k_st16<<<grid, 256>>>((uint4*)p, bytes / 16); });
    timeit("st.128 .cs", bytes, [&] { k_st16_cs<<<grid, 256>>>((uint4*)p, bytes / 16); });
    timeit("st.128 .wt", bytes, [&] { k_st16_wt<<<grid, 256>>>((uint4*)p, bytes / 16); });
    timeit("st.256", bytes, [&] { k_st32<<<grid, 256>>>((uint4*)p, bytes / 32); });
  }
  cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  for (int mult : {1, 2, 4}) {
    char name[64];
    snprintf(name, sizeof name, "TMA bulk store 32K x%d", mult);
    timeit(name, bytes, [&] { k_bulk<<<sms * mult, 128, 32768>>>(p, bytes); });
  }
  timeit("cudaMemset", bytes, [&] { cudaMemsetAsync(p, 0, bytes); });
  return 0;
}
