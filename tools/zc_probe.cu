// Development probe: how fast can SMs read SPARSE pieces (one 32-byte sector per 120-byte cell) of pinned,
// mapped host memory over PCIe, compared with the copy engine's dense cudaMemcpyAsync?  Decides whether a
// host-resident variant of the loss kernel (which needs only 12 of every 240 bytes for cells without object)
// can beat the dense H2D pipeline.  Build: nvcc -arch=sm_100a -O3 tools/zc_probe.cu -o tools/zc_probe.bin
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__global__ void sparse_read(const float* __restrict__ P, const float* __restrict__ T, long cells, int words, float* out) {
  float acc = 0.f;
  long stride = (long)gridDim.x * blockDim.x;
  for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < cells; q += stride) {
    const float2 c = *reinterpret_cast<const float2*>(P + q * 30);
    acc += c.x + c.y;
    if (words > 1) acc += T[q * 30];
  }
  if (acc == 12345.678f) out[0] = acc;
}
// each warp reads whole 128-byte lines (dense) from host: upper bound for SM-issued reads
__global__ void dense_read(const float4* __restrict__ P, long n4, float* out) {
  float acc = 0.f;
  long stride = (long)gridDim.x * blockDim.x;
  for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) { float4 v = P[q]; acc += v.x + v.w; }
  if (acc == 12345.678f) out[0] = acc;
}
// grad write straight into mapped host memory (posted PCIe writes), dense
__global__ void dense_write(float4* __restrict__ G, long n4) {
  long stride = (long)gridDim.x * blockDim.x;
  for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) G[q] = make_float4(0.f, 1.f, 0.f, 0.f);
}

int main() {
  const long cells = 65536L * 196 / 4;  // quarter of config 3: 3.2 M cells, 385 MB per tensor
  const size_t bytes = cells * 120;
  float *hp, *ht, *hg, *dp, *out;
  CK(cudaHostAlloc(&hp, bytes, cudaHostAllocMapped));
  CK(cudaHostAlloc(&ht, bytes, cudaHostAllocMapped));
  CK(cudaHostAlloc(&hg, bytes, cudaHostAllocMapped));
  for (size_t i = 0; i < bytes / 4; i += 1024) hp[i] = ht[i] = 1.f;
  CK(cudaMalloc(&dp, bytes));
  CK(cudaMalloc(&out, 4));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float ms;
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(a)); CK(cudaMemcpyAsync(dp, hp, bytes, cudaMemcpyHostToDevice)); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    CK(cudaEventElapsedTime(&ms, a, b));
    printf("memcpy H2D dense       : %.2f ms  %.1f GB/s  (%.0f Mcells/s if 240 B/cell)\n", ms, bytes / ms / 1e6, bytes / ms / 1e6 * 1e3 / 240);
    CK(cudaEventRecord(a)); CK(cudaMemcpyAsync(hg, dp, bytes, cudaMemcpyDeviceToHost)); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    CK(cudaEventElapsedTime(&ms, a, b));
    printf("memcpy D2H dense       : %.2f ms  %.1f GB/s\n", ms, bytes / ms / 1e6);
  }
  {  // copy-engine gather: 8 bytes out of every 120-byte row (cudaMemcpy2DAsync, host pitch 120 -> device pitch 8)
    for (int width = 8; width <= 32; width *= 2) {
      CK(cudaMemcpy2DAsync(dp, width, hp, 120, width, cells, cudaMemcpyHostToDevice));
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(a)); CK(cudaMemcpy2DAsync(dp, width, hp, 120, width, cells, cudaMemcpyHostToDevice)); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
      CK(cudaEventElapsedTime(&ms, a, b));
      printf("copy-engine 2D gather, %2d B of each 120-B row: %.2f ms  %.0f Mrows/s\n", width, ms, cells / ms / 1e3);
    }
  }
  for (int blocks_per_sm = 2; blocks_per_sm <= 8; blocks_per_sm *= 2) {
    for (int words = 1; words <= 2; ++words) {
      sparse_read<<<148 * blocks_per_sm, 256>>>(hp, ht, cells, words, out);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(a)); sparse_read<<<148 * blocks_per_sm, 256>>>(hp, ht, cells, words, out); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
      CK(cudaEventElapsedTime(&ms, a, b));
      printf("SM sparse read  x%d streams, %d blk/SM: %.2f ms  %.0f Mcells/s  (%.1f GB/s of 32-B sectors)\n", words, blocks_per_sm, ms,
             cells / ms / 1e3, cells * 32.0 * words / ms / 1e6);
    }
  }
  dense_read<<<148 * 8, 256>>>((const float4*)hp, bytes / 16, out); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a)); dense_read<<<148 * 8, 256>>>((const float4*)hp, bytes / 16, out); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  CK(cudaEventElapsedTime(&ms, a, b));
  printf("SM dense read          : %.2f ms  %.1f GB/s\n", ms, bytes / ms / 1e6);
  dense_write<<<148 * 8, 256>>>((float4*)hg, bytes / 16); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a)); dense_write<<<148 * 8, 256>>>((float4*)hg, bytes / 16); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  CK(cudaEventElapsedTime(&ms, a, b));
  printf("SM dense write to host : %.2f ms  %.1f GB/s\n", ms, bytes / ms / 1e6);
  return 0;
}
