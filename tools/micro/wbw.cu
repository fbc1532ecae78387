// Write-bandwidth ceilings for the padded-batch output pattern (1.71 GB int64 + 0.21 GB bytes per 1M molecules):
//   A  st.global.cs 16-byte stores, grid-stride              (what warp_pad_row issues)
//   B  cp.async.bulk shared->global of `chunk` bytes per warp (what the bulk-store flush issues), 1 or 2 slots
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/wbw.cu -o gpurun_out/wbw && gpurun_out/wbw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_plain(longlong2* __restrict__ x, size_t n16) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) __stcs(x + i, make_longlong2(2, 2));
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int SLOTS>
__global__ void k_bulk(char* __restrict__ x, size_t nchunks, int chunk, int fill) {
  extern __shared__ __align__(16) char sm[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  char* slot0 = sm + (size_t)w * SLOTS * chunk;
  const size_t gw = (size_t)blockIdx.x * nw + w, tw = (size_t)gridDim.x * nw;
  int sl = 0;
  for (size_t c = gw; c < nchunks; c += tw) {
    char* s = slot0 + sl * chunk;
    if (lane == 0) {
      if (SLOTS == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
    if (fill)
      for (int i = lane * 16; i < chunk; i += 512) *reinterpret_cast<longlong2*>(s + i) = make_longlong2((long long)c, 2);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(x + c * (size_t)chunk), "r"(smem_u32(s)), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (SLOTS == 2) sl ^= 1;
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
float timeit(F f, int it = 20) {
  for (int i = 0; i < 3; ++i) f();
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int i = 0; i < it; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / it;
}

int main() {
  const size_t bytes = (size_t)213668224 * 8;   // X of the headline workload
  char* x;
  cudaMalloc(&x, bytes + 4096);
  for (int cps : {4, 8, 16}) {
    float ms = timeit([&] { k_plain<<<148 * cps, 256>>>((longlong2*)x, bytes / 16); });
    printf("plain st.cs.128 grid=148x%d x256: %.4f ms  %.0f GB/s\n", cps, ms, bytes / ms / 1e6);
  }
  for (int chunk : {1712, 3424, 6848, 13696}) {
    const int c16 = chunk & ~15;
    for (int warps : {4}) {
      for (int ctas : {3, 5, 6, 8}) {
        for (int fill : {0, 1}) {
          size_t nch = bytes / c16;
          cudaFuncSetAttribute(k_bulk<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
          cudaFuncSetAttribute(k_bulk<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
          if ((size_t)warps * c16 * ctas < 220000) {
            float ms = timeit([&] { k_bulk<1><<<148 * ctas, warps * 32, warps * c16>>>(x, nch, c16, fill); });
            printf("bulk 1 slot chunk=%5d ctas/sm=%d fill=%d: %.4f ms  %.0f GB/s\n", c16, ctas, fill, ms, nch * (double)c16 / ms / 1e6);
          }
          if ((size_t)warps * c16 * 2 * ctas < 220000) {
            float ms = timeit([&] { k_bulk<2><<<148 * ctas, warps * 32, warps * c16 * 2>>>(x, nch, c16, fill); });
            printf("bulk 2 slot chunk=%5d ctas/sm=%d fill=%d: %.4f ms  %.0f GB/s\n", c16, ctas, fill, ms, nch * (double)c16 / ms / 1e6);
          }
        }
      }
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
