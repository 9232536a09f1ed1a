/* Host memcpy rate with T threads: glibc memcpy against AVX2 non-temporal stores (what the pageable staging path is
 * bound by).  usage: hostcopy [threads] [MB per piece]   gcc -O2 -mavx2 -pthread -o hostcopy hostcopy.c */
#define _GNU_SOURCE
#include <immintrin.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static void copy_nt(char *dst, const char *src, size_t n) {
  size_t head = (32 - ((size_t)dst & 31)) & 31;
  if (head > n) head = n;
  memcpy(dst, src, head); dst += head; src += head; n -= head;
  size_t i = 0;
  for (; i + 128 <= n; i += 128) {
    __m256i a = _mm256_loadu_si256((const __m256i *)(src + i)), b = _mm256_loadu_si256((const __m256i *)(src + i + 32));
    __m256i c = _mm256_loadu_si256((const __m256i *)(src + i + 64)), d = _mm256_loadu_si256((const __m256i *)(src + i + 96));
    _mm256_stream_si256((__m256i *)(dst + i), a); _mm256_stream_si256((__m256i *)(dst + i + 32), b);
    _mm256_stream_si256((__m256i *)(dst + i + 64), c); _mm256_stream_si256((__m256i *)(dst + i + 96), d);
  }
  _mm_sfence();
  memcpy(dst + i, src + i, n - i);
}
struct job { char *dst; const char *src; size_t n; int nt; };
static void *run(void *p) { struct job *j = p; if (j->nt) copy_nt(j->dst, j->src, j->n); else memcpy(j->dst, j->src, j->n); return 0; }
int main(int argc, char **argv) {
  int T = argc > 1 ? atoi(argv[1]) : 8;
  size_t piece = (size_t)(argc > 2 ? atoi(argv[2]) : 8) << 20, total = piece * T, rounds = ((size_t)2 << 30) / total;
  char *src = aligned_alloc(4096, total), *dst = aligned_alloc(4096, total * rounds);
  memset(src, 1, total); memset(dst, 0, total * rounds);
  for (int nt = 0; nt < 2; ++nt) {
    double t0 = now();
    for (size_t r = 0; r < rounds; ++r) {
      pthread_t th[64]; struct job jb[64];
      for (int i = 0; i < T; ++i) { jb[i] = (struct job){dst + r * total + i * piece, src + i * piece, piece, nt}; pthread_create(&th[i], 0, run, &jb[i]); }
      for (int i = 0; i < T; ++i) pthread_join(th[i], 0);
    }
    double dt = now() - t0;
    printf("%d threads x %zu MB pieces, %s: %.1f GB/s\n", T, piece >> 20, nt ? "AVX2 non-temporal stores" : "glibc memcpy", total * rounds / dt / 1e9);
  }
  return 0;
}
