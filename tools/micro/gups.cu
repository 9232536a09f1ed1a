// Random-sector read rate of HBM (the roofline of the probe): each thread reads `per` independent random
// 16/32-byte pieces from a table of `mb` MiB.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gups gups.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
template <int PER, int VEC>
__global__ void probe(const uint4 *tab, uint64_t mask, uint64_t n, uint32_t *out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i * PER >= n) return;
  uint4 v[PER][VEC];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    uint64_t h = (mix(i * PER + j) & mask) & ~uint64_t(VEC - 1);
#pragma unroll
    for (int w = 0; w < VEC; ++w)
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[j][w].x), "=r"(v[j][w].y), "=r"(v[j][w].z), "=r"(v[j][w].w) : "l"(tab + h + w));
  }
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < PER; ++j)
#pragma unroll
    for (int w = 0; w < VEC; ++w) acc += v[j][w].x ^ v[j][w].w;
  if (acc == 0x12345678u) out[0] = acc;
}
template <int PER>
__global__ void probe256(const uint4 *tab, uint64_t mask, uint64_t n, uint32_t *out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i * PER >= n) return;
  uint32_t v[PER][8];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    uint64_t h = (mix(i * PER + j) & mask) & ~uint64_t(1);
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[j][0]), "=r"(v[j][1]), "=r"(v[j][2]), "=r"(v[j][3]), "=r"(v[j][4]), "=r"(v[j][5]), "=r"(v[j][6]), "=r"(v[j][7]) : "l"(tab + h));
  }
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < PER; ++j) acc += v[j][0] ^ v[j][7];
  if (acc == 0x12345678u) out[0] = acc;
}
int main(int argc, char **argv) {
  const uint64_t n = 200000000ull;
  if (argc > 1) {
    size_t g = 0;
    cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1]));
    cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    printf("cudaLimitMaxL2FetchGranularity <- %s: %s, now %zu\n", argv[1], cudaGetErrorString(e), g);
  }
  uint32_t *out; cudaMalloc(&out, 4);
  for (int mb : {64, 8192}) {
    uint64_t slots = (uint64_t)mb * 1024 * 1024 / 16;
    uint4 *tab; if (cudaMalloc(&tab, slots * 16) != cudaSuccess) { printf("alloc %d MiB failed\n", mb); continue; }
    cudaMemset(tab, 1, slots * 16);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int vec = 1; vec <= 3; ++vec) {
      float best = 1e9;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        if (vec == 1) probe<8, 1><<<(unsigned)((n / 8 + 255) / 256), 256>>>(tab, slots - 1, n, out);
        else if (vec == 2) probe<8, 2><<<(unsigned)((n / 8 + 255) / 256), 256>>>(tab, slots - 1, n, out);
        else probe256<8><<<(unsigned)((n / 8 + 255) / 256), 256>>>(tab, slots - 1, n, out);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
      }
      printf("table %6d MiB, %s: %.2f G reads/s (%.0f GB/s of sectors)\n", mb, vec == 1 ? "one 16-byte load" : vec == 2 ? "two 16-byte loads of one sector" : "one 32-byte load", n / best / 1e6, n / best / 1e6 * 32);
    }
    cudaFree(tab);
  }
  return 0;
}
