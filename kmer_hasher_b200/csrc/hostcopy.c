/* hostcopy.c -- the host side of the pageable staging path (api.cu, CopyPool): copy between a pinned staging slot and the
 * caller's ordinary memory.  A large copy with regular stores reads every destination line before overwriting it (read for
 * ownership) and pushes it through the caches: three DRAM transfers per byte instead of two.  Non-temporal stores skip
 * both; measured on the B200 boxes' hosts, 8 threads: 84 GB/s against glibc memcpy's 44 GB/s (tools/micro/hostcopy.c) --
 * the difference between a staging path bound by the host (44 < PCIe's 57 GB/s) and one bound by the link.
 * Plain C so that it is compiled by gcc with the AVX2 target attribute, away from nvcc's front end. */
#include <stddef.h>
#include <string.h>

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
__attribute__((target("avx2"))) static void copy_stream_avx2(char *dst, const char *src, size_t n) {
  size_t head = (32 - ((size_t)dst & 31)) & 31;
  if (head > n) head = n;
  memcpy(dst, src, head);
  dst += head; src += head; n -= head;
  size_t i = 0;
  for (; i + 128 <= n; i += 128) {
    const __m256i a = _mm256_loadu_si256((const __m256i *)(src + i)), b = _mm256_loadu_si256((const __m256i *)(src + i + 32));
    const __m256i c = _mm256_loadu_si256((const __m256i *)(src + i + 64)), d = _mm256_loadu_si256((const __m256i *)(src + i + 96));
    _mm256_stream_si256((__m256i *)(dst + i), a);
    _mm256_stream_si256((__m256i *)(dst + i + 32), b);
    _mm256_stream_si256((__m256i *)(dst + i + 64), c);
    _mm256_stream_si256((__m256i *)(dst + i + 96), d);
  }
  _mm_sfence();                                   /* the stores are globally visible before the caller is told */
  memcpy(dst + i, src + i, n - i);
}
static int have_avx2(void) {
  static int v = -1;
  if (v < 0) { __builtin_cpu_init(); v = __builtin_cpu_supports("avx2") ? 1 : 0; }
  return v;
}
#endif

/* copy n bytes; large pieces with non-temporal stores where the CPU has AVX2, else memcpy */
void kmg_host_copy(void *dst, const void *src, size_t n) {
#if defined(__x86_64__) && defined(__GNUC__)
  if (n >= ((size_t)1 << 20) && have_avx2()) { copy_stream_avx2((char *)dst, (const char *)src, n); return; }
#endif
  memcpy(dst, src, n);
}
