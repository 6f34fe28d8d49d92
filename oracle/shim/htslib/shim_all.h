/*
 * oracle/shim/htslib/shim_all.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A small stand-in for the parts of htslib (>= 1.10) that the unmodified
 * reference sources under /root/reference include and call.  htslib is not
 * installed in this image and is not vendored by the reference
 * (src/kbbq/CMakeLists.txt:13-16 links the system `hts`), so the oracle build
 * (oracle/Makefile -> oracle/_ref/) compiles the reference against this header.
 *
 * What the hot path needs from htslib is pure data: the three nucleotide
 * lookup tables (call sites: include/kbbq/bloom.hh:351,374;
 * include/kbbq/readutils.hh:35-36; src/kbbq/readutils.cc:353,487,494,582-583;
 * include/kbbq/covariateutils.hh:88,93-94).  Their values are htslib's
 * documented ones.  Everything else here is record I/O
 * glue: FASTQ (plain or gzip) through zlib, and a compact BAM reader/writer.
 *
 * Written from the public htslib API documentation; no htslib source is
 * copied.  Nothing under kbbq_b200/ includes this file.
 */
#ifndef KBBQ_ORACLE_HTS_SHIM_H
#define KBBQ_ORACLE_HTS_SHIM_H

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <errno.h>
#include <zlib.h>

#define HTS_VERSION 101000

#ifdef __cplusplus
extern "C" {
#endif

/* ---- nucleotide tables (definitions in shim.c) ------------------------- */
extern const unsigned char seq_nt16_table[256];
extern const char seq_nt16_str[];
extern const int seq_nt16_int[];

/* ---- kstring ------------------------------------------------------------ */
typedef struct kstring_t {
    size_t l, m;
    char *s;
} kstring_t;

static inline const char *ks_c_str(kstring_t *s) { return s->l && s->s ? s->s : ""; }

/* ---- thread pool (no-op: the reference's --threads only sizes htslib's
 *      (de)compression pool, kbbq.cc:159-168) ----------------------------- */
typedef struct hts_tpool hts_tpool;
typedef struct htsThreadPool {
    hts_tpool *pool;
    int qsize;
} htsThreadPool;
hts_tpool *hts_tpool_init(int n);
void hts_tpool_destroy(hts_tpool *p);

/* ---- BGZF: gzip/plain stream through zlib -------------------------------- */
typedef struct BGZF {
    gzFile gz;
    int is_write;
} BGZF;
BGZF *bgzf_open(const char *path, const char *mode);
ssize_t bgzf_read(BGZF *fp, void *data, size_t length);
ssize_t bgzf_write(BGZF *fp, const void *data, size_t length);
int bgzf_close(BGZF *fp);
int bgzf_thread_pool(BGZF *fp, hts_tpool *pool, int qsize);

/* ---- hFILE + format sniffing --------------------------------------------- */
typedef struct hFILE {
    FILE *fp;
    char *path;
} hFILE;
hFILE *hopen(const char *filename, const char *mode, ...);
int hclose(hFILE *fp);
void hclose_abruptly(hFILE *fp);

enum htsExactFormat {
    unknown_format, binary_format, text_format, sam, bam, bai, cram, crai, vcf, bcf, csi, gzi, tbi, bed,
    htsget, json, empty_format, fasta_format, fastq_format, fai_format, fqi_format
};
typedef struct htsFormat {
    int category;
    enum htsExactFormat format;
    struct { short major, minor; } version;
    int compression;
    short compression_level;
    void *specific;
} htsFormat;
int hts_detect_format(hFILE *fp, htsFormat *fmt);

/* ---- SAM/BAM -------------------------------------------------------------- */
typedef struct sam_hdr_t {
    int32_t n_targets;
    uint32_t *target_len;
    char **target_name;
    size_t l_text;
    char *text;
} sam_hdr_t;
typedef sam_hdr_t bam_hdr_t;

typedef struct bam1_core_t {
    int64_t pos;
    int32_t tid;
    uint16_t bin;
    uint8_t qual;
    uint8_t l_extranul;
    uint16_t flag;
    uint16_t l_qname;
    uint32_t n_cigar;
    int32_t l_qseq;
    int32_t mtid;
    int64_t mpos;
    int64_t isize;
} bam1_core_t;

typedef struct bam1_t {
    bam1_core_t core;
    uint64_t id;
    uint8_t *data;
    int l_data;
    uint32_t m_data;
} bam1_t;

#define BAM_FREVERSE 16
#define BAM_FREAD1 64
#define BAM_FREAD2 128

#define bam_is_rev(b) (((b)->core.flag & BAM_FREVERSE) != 0)
#define bam_get_qname(b) ((char *)(b)->data)
#define bam_get_cigar(b) ((uint32_t *)((b)->data + (b)->core.l_qname))
#define bam_get_seq(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname)
#define bam_get_qual(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1))
#define bam_get_aux(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1) + (b)->core.l_qseq)
#define bam_get_l_aux(b) ((b)->l_data - ((b)->core.n_cigar << 2) - (b)->core.l_qname - (b)->core.l_qseq - (((b)->core.l_qseq + 1) >> 1))
#define bam_seqi(s, i) ((s)[(i) >> 1] >> ((~(i) & 1) << 2) & 0xf)

bam1_t *bam_init1(void);
void bam_destroy1(bam1_t *b);
uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]);
char *bam_aux2Z(const uint8_t *s);
int bam_aux_update_str(bam1_t *b, const char tag[2], int len, const char *data);

typedef struct htsFile {
    gzFile gz;
    int is_write;
    int is_bam;
} htsFile;
typedef htsFile samFile;

samFile *sam_open(const char *fn, const char *mode);
int sam_close(samFile *fp);
htsFile *hts_hopen(hFILE *fp, const char *fn, const char *mode);
int hts_close(htsFile *fp);
int hts_set_thread_pool(htsFile *fp, htsThreadPool *p);
sam_hdr_t *sam_hdr_read(samFile *fp);
int sam_hdr_write(samFile *fp, const sam_hdr_t *h);
void sam_hdr_destroy(sam_hdr_t *h);
int sam_read1(samFile *fp, sam_hdr_t *h, bam1_t *b);
int sam_write1(samFile *fp, const sam_hdr_t *h, const bam1_t *b);
static inline int sam_hdr_nref(const sam_hdr_t *h) { return h ? h->n_targets : -1; }
static inline int64_t sam_hdr_tid2len(const sam_hdr_t *h, int tid) {
    return (h && tid >= 0 && tid < h->n_targets) ? (int64_t)h->target_len[tid] : 0;
}

#ifdef __cplusplus
}
#endif

/* ---- kseq: a FASTA/FASTQ record reader with klib's documented contract ----
 * kseq_read returns the sequence length (>=0), -1 at EOF, -2 when the quality
 * string is shorter than the sequence.  name = header up to the first white
 * space, comment = rest of the header line.  Multi-line records are accepted.
 */
#define KSEQ_DECLARE(type_t)

#define KSEQ_INIT(type_t, __read)                                                                                \
    typedef struct kseq_t {                                                                                      \
        kstring_t name, comment, seq, qual;                                                                      \
        int last_char;                                                                                           \
        type_t f;                                                                                                \
        unsigned char *buf;                                                                                      \
        int begin, end, is_eof;                                                                                  \
    } kseq_t;                                                                                                    \
    static inline kseq_t *kseq_init(type_t fd) {                                                                 \
        kseq_t *s = (kseq_t *)calloc(1, sizeof(kseq_t));                                                         \
        s->f = fd;                                                                                               \
        s->buf = (unsigned char *)malloc(65536);                                                                 \
        return s;                                                                                                \
    }                                                                                                            \
    static inline void kseq_destroy(kseq_t *ks) {                                                                \
        if (!ks) return;                                                                                         \
        free(ks->name.s); free(ks->comment.s); free(ks->seq.s); free(ks->qual.s); free(ks->buf);                 \
        free(ks);                                                                                                \
    }                                                                                                            \
    static inline int kseq_shim_getc(kseq_t *ks) {                                                               \
        if (ks->begin >= ks->end) {                                                                              \
            if (ks->is_eof) return -1;                                                                           \
            ks->begin = 0;                                                                                       \
            ks->end = (int)__read(ks->f, ks->buf, 65536);                                                        \
            if (ks->end <= 0) { ks->is_eof = 1; ks->end = 0; return -1; }                                        \
        }                                                                                                        \
        return (int)ks->buf[ks->begin++];                                                                        \
    }                                                                                                            \
    static inline void kseq_shim_putc(kstring_t *s, int c) {                                                     \
        if (s->l + 2 > s->m) { s->m = s->m ? s->m * 2 : 256; s->s = (char *)realloc(s->s, s->m); }               \
        s->s[s->l++] = (char)c; s->s[s->l] = 0;                                                                  \
    }                                                                                                            \
    static inline void kseq_shim_clear(kstring_t *s) {                                                           \
        s->l = 0; if (s->s) s->s[0] = 0;                                                                         \
    }                                                                                                            \
    /* read to end of line into s (without the newline); returns the terminating char or -1 */                  \
    static inline int kseq_shim_line(kseq_t *ks, kstring_t *s) {                                                 \
        int c;                                                                                                   \
        while ((c = kseq_shim_getc(ks)) >= 0 && c != '\n') if (c != '\r') kseq_shim_putc(s, c);                   \
        return c;                                                                                                \
    }                                                                                                            \
    static inline int kseq_read(kseq_t *ks) {                                                                    \
        int c;                                                                                                   \
        if (ks->last_char == 0) {                                                                                \
            while ((c = kseq_shim_getc(ks)) >= 0 && c != '>' && c != '@') {}                                      \
            if (c < 0) return -1;                                                                                \
            ks->last_char = c;                                                                                   \
        }                                                                                                        \
        kseq_shim_clear(&ks->name); kseq_shim_clear(&ks->comment);                                               \
        kseq_shim_clear(&ks->seq); kseq_shim_clear(&ks->qual);                                                   \
        /* header: name up to white space, comment = remainder */                                               \
        while ((c = kseq_shim_getc(ks)) >= 0 && c != '\n' && c != ' ' && c != '\t' && c != '\r')                  \
            kseq_shim_putc(&ks->name, c);                                                                        \
        if (c < 0) return -1;                                                                                    \
        if (c == ' ' || c == '\t') c = kseq_shim_line(ks, &ks->comment);                                          \
        else if (c == '\r') c = kseq_shim_line(ks, &ks->comment);                                                 \
        if (!ks->seq.s) { ks->seq.m = 256; ks->seq.s = (char *)malloc(256); ks->seq.s[0] = 0; }                   \
        /* sequence lines until '+', '>' or '@' at line start */                                                \
        while ((c = kseq_shim_getc(ks)) >= 0 && c != '>' && c != '+' && c != '@') {                               \
            if (c == '\n' || c == '\r') continue;                                                                \
            kseq_shim_putc(&ks->seq, c);                                                                         \
            c = kseq_shim_line(ks, &ks->seq);                                                                    \
            if (c < 0) break;                                                                                    \
        }                                                                                                        \
        if (c == '>' || c == '@') ks->last_char = c; else ks->last_char = 0;                                     \
        if (c != '+') return (int)ks->seq.l;                                                                     \
        if (!ks->qual.s) { ks->qual.m = ks->seq.m; ks->qual.s = (char *)malloc(ks->qual.m); ks->qual.s[0] = 0; }  \
        { kstring_t skip = {0, 0, 0}; c = kseq_shim_line(ks, &skip); free(skip.s); }                              \
        if (c < 0) return -2;                                                                                    \
        while (ks->qual.l < ks->seq.l) {                                                                         \
            size_t before = ks->qual.l;                                                                          \
            c = kseq_shim_line(ks, &ks->qual);                                                                   \
            if (c < 0 && ks->qual.l == before) break;                                                            \
            if (c < 0) break;                                                                                    \
        }                                                                                                        \
        ks->last_char = 0;                                                                                       \
        if (ks->seq.l != ks->qual.l) return -2;                                                                  \
        return (int)ks->seq.l;                                                                                   \
    }

#endif
