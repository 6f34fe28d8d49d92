/*
 * oracle/kbbq_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement ("port") of the reference's read-parallel recalibration
 * hot path, SURVEY.md §8(a) rows a1-a19.  It exists to check the CUDA path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load it.  Nothing under kbbq_b200/ includes, links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function
 * here against the unmodified reference compiled into oracle/_ref
 * (oracle/Makefile target `ref`) and against the known-answer vectors in
 * tests/golden/ that were generated from that build.
 *
 * Read sets use the same flat layout everywhere:
 *   seq[total] ASCII, off[n+1] base offsets, qual[total] phred (no +33),
 *   rg[n] dense read-group ids, second[n] 0/1.
 */
#ifndef KBBQ_ORACLE_H
#define KBBQ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KO_NUM_PATTERNS 65536
#define KO_MAXQ 93

typedef struct ko_bloom {
    uint64_t table_bits;   /* padded to a multiple of 512 */
    uint64_t n_blocks;
    uint32_t n_salts;
    uint32_t salts[128];
    uint64_t random_seed;  /* blocked_bloom_filter::random_seed_ */
    uint32_t optimal_k;
    uint64_t optimal_table_size;
    uint64_t inserted;
    uint64_t *words;       /* n_blocks * 8 */
    uint64_t *patterns;    /* 65536 * 8 */
} ko_bloom;

/* primitives */
int ko_nt4(unsigned char ch);
uint32_t ko_hash_ap8(uint64_t key, uint32_t salt);
void ko_kmer_of(int k, const char *s, size_t n, uint64_t out[5]);
void ko_rng_seed(uint32_t seed, uint64_t state[4]);
uint64_t ko_rng_next(uint64_t state[4]);
void ko_rng_outputs(uint32_t seed, uint64_t skip, uint64_t n, uint64_t state_after_seed[4], uint64_t *out);
void ko_bernoulli(uint32_t seed, double p, uint64_t n, uint8_t *out);

/* Bloom */
ko_bloom *ko_bloom_new(uint64_t projected, double fpr, uint64_t seed);
void ko_bloom_free(ko_bloom *b);
void ko_bloom_insert(ko_bloom *b, uint64_t key);
int ko_bloom_contains(const ko_bloom *b, uint64_t key);
double ko_bloom_fprate(const ko_bloom *b);

/* host maths between passes */
long double ko_phit(const ko_bloom *sampled, long double alpha);
void ko_thresholds(int k, long double p, int *out /* k+1 */);
void ko_thresholds_from_filter(const ko_bloom *sampled, long double alpha_num, long double alpha_den, int k, int *out);

/* passes */
uint64_t ko_subsample(uint64_t n, const char *seq, const uint64_t *off, int k, double alpha, uint32_t seed,
                      ko_bloom *sampled);
uint64_t ko_subsample_skip(uint64_t n, const char *seq, const uint64_t *off, int k, double alpha, uint32_t seed,
                           uint64_t skip, ko_bloom *sampled);
void ko_infer_errors(uint64_t n, const char *seq, const uint64_t *off, const uint8_t *qual, int k,
                     const ko_bloom *sampled, const int *thresholds, uint8_t *errors);
void ko_find_trusted(uint64_t n, const char *seq, const uint64_t *off, const uint8_t *qual, int k,
                     const ko_bloom *sampled, const int *thresholds, ko_bloom *trusted);
/* errors[total] out; seq_after[total] (may be NULL) receives read.seq as consume_read sees it */
void ko_get_errors(uint64_t n, const char *seq, const uint64_t *off, const uint8_t *qual, int k,
                   const ko_bloom *trusted, uint8_t *errors, char *seq_after);
/* dense tables: rg[R][2], q[R][94][2], cy[R][94][2][maxlen][2], di[R][94][16][2] (accumulated into) */
void ko_consume(uint64_t n, const char *seq, const uint64_t *off, const uint8_t *qual, const int32_t *rg,
                const uint8_t *second, const uint8_t *errors, int R, int maxlen, uint64_t *rgc, uint64_t *qc,
                uint64_t *cyc, uint64_t *dic);
/* dense dq tables: meanq[R], rgdq[R], qdq[R][94], cydq[R][94][2][maxlen], didq[R][94][16] */
void ko_get_dqs(int R, int maxlen, const uint64_t *rgc, const uint64_t *qc, const uint64_t *cyc, const uint64_t *dic,
                int32_t *meanq, int32_t *rgdq, int32_t *qdq, int32_t *cydq, int32_t *didq);
void ko_recalibrate(uint64_t n, const char *seq, const uint64_t *off, const uint8_t *qual, const int32_t *rg,
                    const uint8_t *second, int R, int maxlen, const int32_t *meanq, const int32_t *rgdq,
                    const int32_t *qdq, const int32_t *cydq, const int32_t *didq, uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif
