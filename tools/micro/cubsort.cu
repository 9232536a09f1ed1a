// Library baseline for the index build's sort: cub::DeviceRadixSort::SortPairs on (uint64 key, uint32 pos)
// records (CUB of CUDA 12.9: one-sweep, 8 bits per pass).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o cubsort cubsort.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cub/cub.cuh>
__global__ void fill(uint64_t *k, uint32_t *v, uint64_t n, int bits) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t x = i * 0x9E3779B97F4A7C15ull + 12345; x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  k[i] = bits < 64 ? (x & ((1ull << bits) - 1)) : x; v[i] = (uint32_t)i;
}
int main(int argc, char **argv) {
  const uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 39999969ull;
  uint64_t *k0, *k1; uint32_t *v0, *v1;
  cudaMalloc(&k0, n * 8); cudaMalloc(&k1, n * 8); cudaMalloc(&v0, n * 4); cudaMalloc(&v1, n * 4);
  for (int bits : {64, 42, 32, 24}) {
    size_t tmp = 0; void *d_tmp = nullptr;
    cub::DeviceRadixSort::SortPairs(d_tmp, tmp, k0, k1, v0, v1, (int)n, 0, bits);
    cudaMalloc(&d_tmp, tmp);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
      fill<<<(unsigned)((n + 255) / 256), 256>>>(k0, v0, n, bits);
      cudaEventRecord(a);
      cub::DeviceRadixSort::SortPairs(d_tmp, tmp, k0, k1, v0, v1, (int)n, 0, bits);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    const int passes = (bits + 7) / 8;
    printf("cub SortPairs<u64,u32> n=%llu bits=%d: %.3f ms = %.1f us per 8-bit pass (incl. histogram) = %.0f GB/s per pass (24 B/record)\n",
           (unsigned long long)n, bits, best, best * 1e3 / passes, 24.0 * n / (best / passes) / 1e6);
    cudaFree(d_tmp);
  }
  return 0;
}
