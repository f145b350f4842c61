/*
 * sts_generators.h - the chained generators of the reference's tools/generators.c (3, 6, 7, 8, 9) on the host; see
 * sts_generators.c.  Generators 1, 2, 4 and 5 are made on the GPU: stsgpu_generate() in include/stsgpu.h.
 */
#ifndef STS_GENERATORS_H
#define STS_GENERATORS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

struct sts_gen;

/* 1 for the generators this file makes (3, 6, 7, 8, 9) */
int sts_gen_is_chained(int generator);
/* the start of the tool's sequence for streams of n bits (n a multiple of 8); NULL for another generator or a bad n */
struct sts_gen *sts_gen_open(int generator, long n);
/* the NEXT nstreams streams, n / 8 file bytes each (as `generators <g> <count>` writes them), into raw; 0 or -1 */
int sts_gen_next(struct sts_gen *g, uint8_t *raw, long nstreams);
/* generator 8 only, before the first sts_gen_next: the 128 bytes the tool's first block finds in its accumulator (default zero) */
void sts_gen_set_ms_initial(struct sts_gen *g, const uint8_t y[128]);
void sts_gen_close(struct sts_gen *g);

#ifdef __cplusplus
}
#endif
#endif
