/*
 * oracle/ref_fasta.c -- TEST INFRASTRUCTURE, not product code.
 *
 * The reference reads FASTA/FASTQ (plain or gz) with klib's kseq.h over zlib (src/kseq.h; KSEQ_INIT at
 * src/kmer_reader.h:8, the read loop at src/kmer_reader.c:41-77).  This driver runs that UNMODIFIED parser -- kseq.h is
 * included from where it lies under /root/reference/src -- and hands back every record's name and sequence, so that the
 * device parser of libkmergpu (csrc/reads.cuh) can be compared with it.  Built into oracle/_ref/libkmer_ref.so.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include "kseq.h"   /* from -I/root/reference/src */
KSEQ_INIT(gzFile, gzread)

/* Reads the whole file.  *names_out / *seqs_out: the names / sequences back to back, each followed by '\n';
 * *lens_out: sequence length per record.  Returns the number of records, or -1 if the file cannot be opened. */
int64_t ref_reads_parse(const char *path, char **names_out, char **seqs_out, int64_t **lens_out) {
  gzFile gz = gzopen(path, "r");
  if (!gz) return -1;
  kseq_t *ks = kseq_init(gz);
  size_t ncap = 1 << 16, scap = 1 << 20, lcap = 1 << 10, nl = 0, sl = 0;
  char *names = malloc(ncap), *seqs = malloc(scap);
  int64_t *lens = malloc(lcap * sizeof(int64_t));
  int64_t n = 0;
  while (kseq_read(ks) >= 0) {
    while (nl + ks->name.l + 2 > ncap) names = realloc(names, ncap *= 2);
    while (sl + ks->seq.l + 2 > scap) seqs = realloc(seqs, scap *= 2);
    if ((size_t)n + 1 > lcap) lens = realloc(lens, (lcap *= 2) * sizeof(int64_t));
    memcpy(names + nl, ks->name.s, ks->name.l); nl += ks->name.l; names[nl++] = '\n';
    memcpy(seqs + sl, ks->seq.s, ks->seq.l); sl += ks->seq.l; seqs[sl++] = '\n';
    lens[n++] = (int64_t)ks->seq.l;
  }
  names[nl] = 0; seqs[sl] = 0;
  kseq_destroy(ks);
  gzclose(gz);
  *names_out = names; *seqs_out = seqs; *lens_out = lens;
  return n;
}
void ref_reads_free(void *p) { free(p); }
