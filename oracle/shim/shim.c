/*
 * oracle/shim/shim.c -- TEST INFRASTRUCTURE ONLY (see htslib/shim_all.h).
 * Implements the few htslib entry points the unmodified reference links
 * against: nucleotide tables, gzip-backed BGZF, format sniffing, and a compact
 * BAM reader/writer (BGZF is a sequence of gzip members, which zlib's gz* API
 * reads transparently; output is written as one gzip stream).
 */
#include "htslib/shim_all.h"
#include <unistd.h>

/* htslib's documented IUPAC tables. */
const char seq_nt16_str[] = "=ACMGRSVTWYHKDBN";
const int seq_nt16_int[] = {4, 0, 1, 4, 2, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4};
#define X 15
const unsigned char seq_nt16_table[256] = {
    X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X,
    X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, 1, 2, 4, 8, X, X, X, X, X, X, X, X, X, 0, X, X,
    /*@*/ X, 1, 14, 2, 13, X, X, 4, 11, X, X, 12, X, 3, X, X, X, X, 5, 6, 8, X, 7, 9, X, 10, X, X, X, X, X, X,
    /*`*/ X, 1, 14, 2, 13, X, X, 4, 11, X, X, 12, X, 3, X, X, X, X, 5, 6, 8, X, 7, 9, X, 10, X, X, X, X, X, X,
    X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X,
    X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X,
    X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X,
    X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X, X};
#undef X

/* ---- thread pool: a non-null token, nothing else ------------------------- */
struct hts_tpool { int n; };
hts_tpool *hts_tpool_init(int n) {
    hts_tpool *p = (hts_tpool *)malloc(sizeof(*p));
    if (p) p->n = n;
    return p;
}
void hts_tpool_destroy(hts_tpool *p) { free(p); }

/* ---- BGZF ------------------------------------------------------------------ */
static gzFile shim_gzopen(const char *path, const char *mode) {
    int wr = strchr(mode, 'w') != NULL;
    char zmode[8];
    if (wr) {
        /* "wu" = uncompressed in htslib; everything else default level */
        snprintf(zmode, sizeof zmode, "%s", strchr(mode, 'u') ? "wbT" : "wb");
    } else {
        snprintf(zmode, sizeof zmode, "rb");
    }
    if (strcmp(path, "-") == 0) return gzdopen(dup(wr ? 1 : 0), zmode);
    return gzopen(path, zmode);
}

BGZF *bgzf_open(const char *path, const char *mode) {
    gzFile gz = shim_gzopen(path, mode);
    if (!gz) return NULL;
    gzbuffer(gz, 1 << 20);
    BGZF *fp = (BGZF *)calloc(1, sizeof(BGZF));
    fp->gz = gz;
    fp->is_write = strchr(mode, 'w') != NULL;
    return fp;
}
ssize_t bgzf_read(BGZF *fp, void *data, size_t length) { return gzread(fp->gz, data, (unsigned)length); }
ssize_t bgzf_write(BGZF *fp, const void *data, size_t length) {
    int n = gzwrite(fp->gz, data, (unsigned)length);
    return n == 0 && length ? -1 : n;
}
int bgzf_close(BGZF *fp) {
    int r = gzclose(fp->gz);
    free(fp);
    return r == Z_OK ? 0 : -1;
}
int bgzf_thread_pool(BGZF *fp, hts_tpool *pool, int qsize) {
    (void)fp; (void)pool; (void)qsize;
    return 0;
}

/* ---- hFILE + format sniffing ------------------------------------------------ */
hFILE *hopen(const char *filename, const char *mode, ...) {
    (void)mode;
    hFILE *h = (hFILE *)calloc(1, sizeof(hFILE));
    h->path = strdup(filename);
    h->fp = strcmp(filename, "-") == 0 ? stdin : fopen(filename, "rb");
    if (!h->fp) { free(h->path); free(h); return NULL; }
    return h;
}
int hclose(hFILE *fp) {
    if (!fp) return 0;
    int r = (fp->fp && fp->fp != stdin) ? fclose(fp->fp) : 0;
    free(fp->path);
    free(fp);
    return r;
}
void hclose_abruptly(hFILE *fp) { if (fp) hclose(fp); }

int hts_detect_format(hFILE *fp, htsFormat *fmt) {
    if (!fp || !fp->fp) return -1;
    memset(fmt, 0, sizeof *fmt);
    fmt->format = unknown_format;
    if (strcmp(fp->path, "-") == 0) return -1; /* the reference re-opens its input for every pass */
    gzFile gz = gzopen(fp->path, "rb");
    if (!gz) return -1;
    unsigned char head[4] = {0, 0, 0, 0};
    int n = gzread(gz, head, 4);
    gzclose(gz);
    if (n >= 4 && memcmp(head, "BAM\1", 4) == 0) fmt->format = bam;
    else if (n >= 1 && head[0] == '@') fmt->format = fastq_format;
    else if (n >= 1 && head[0] == '>') fmt->format = fasta_format;
    return 0;
}

/* ---- BAM --------------------------------------------------------------------- */
bam1_t *bam_init1(void) { return (bam1_t *)calloc(1, sizeof(bam1_t)); }
void bam_destroy1(bam1_t *b) {
    if (!b) return;
    free(b->data);
    free(b);
}

static int aux_skip(const uint8_t *s, const uint8_t *end, const uint8_t **next) {
    /* s points at the type byte */
    if (s >= end) return -1;
    uint8_t t = *s++;
    size_t sz;
    switch (t) {
    case 'A': case 'c': case 'C': sz = 1; break;
    case 's': case 'S': sz = 2; break;
    case 'i': case 'I': case 'f': sz = 4; break;
    case 'd': sz = 8; break;
    case 'Z': case 'H':
        while (s < end && *s) ++s;
        if (s >= end) return -1;
        *next = s + 1;
        return 0;
    case 'B': {
        if (end - s < 5) return -1;
        uint8_t st = *s++;
        uint32_t cnt;
        memcpy(&cnt, s, 4);
        s += 4;
        size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
        sz = es * (size_t)cnt;
        break;
    }
    default: return -1;
    }
    if ((size_t)(end - s) < sz) return -1;
    *next = s + sz;
    return 0;
}

uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]) {
    const uint8_t *s = bam_get_aux(b), *end = b->data + b->l_data;
    while (end - s >= 3) {
        const uint8_t *next;
        if (s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1]) return (uint8_t *)(s + 2);
        if (aux_skip(s + 2, end, &next) < 0) { errno = EINVAL; return NULL; }
        s = next;
    }
    errno = ENOENT;
    return NULL;
}

char *bam_aux2Z(const uint8_t *s) {
    if (*s == 'Z' || *s == 'H') return (char *)(s + 1);
    errno = EINVAL;
    return NULL;
}

/* htslib's kroundup32: capacity of bam1_t::data grows to the next power of two */
static uint32_t shim_roundup32(uint32_t x) {
    --x; x |= x >> 1; x |= x >> 2; x |= x >> 4; x |= x >> 8; x |= x >> 16;
    return ++x;
}

/* Like htslib (sam.c bam_aux_update_str): the tag is rewritten IN PLACE -- the data buffer is only reallocated when
 * its capacity (m_data, a power of two kept across records) is too small.  This matters for parity: the reference's
 * BamFile::recalibrate (htsiter.cc:11-35) takes bam_get_qual() BEFORE calling this function and writes the new
 * qualities through that pointer afterwards, which is only valid while the buffer does not move. */
int bam_aux_update_str(bam1_t *b, const char tag[2], int len, const char *data) {
    /* len includes the terminating NUL when data is NUL-terminated (htslib contract) */
    size_t dlen = (len > 0 && data[len - 1] == 0) ? (size_t)len - 1 : (len < 0 ? strlen(data) : (size_t)len);
    uint8_t *old = bam_aux_get(b, tag);
    size_t head, tail_off;
    if (old) {
        if (*old != 'Z') { errno = EINVAL; return -1; }
        const uint8_t *next;
        if (aux_skip(old, b->data + b->l_data, &next) < 0) { errno = EINVAL; return -1; }
        head = (size_t)(old - 2 - b->data);
        tail_off = (size_t)(next - b->data);
    } else {
        head = (size_t)b->l_data;
        tail_off = (size_t)b->l_data;
    }
    size_t tail = (size_t)b->l_data - tail_off;
    size_t newlen = head + 3 + dlen + 1 + tail;
    if (newlen > b->m_data) {
        uint32_t cap = shim_roundup32((uint32_t)newlen);
        uint8_t *nd = (uint8_t *)realloc(b->data, cap);
        if (!nd) { errno = ENOMEM; return -1; }
        b->data = nd;
        b->m_data = cap;
    }
    memmove(b->data + head + 3 + dlen + 1, b->data + tail_off, tail);
    b->data[head] = (uint8_t)tag[0];
    b->data[head + 1] = (uint8_t)tag[1];
    b->data[head + 2] = 'Z';
    memcpy(b->data + head + 3, data, dlen);
    b->data[head + 3 + dlen] = 0;
    b->l_data = (int)newlen;
    return 0;
}

static samFile *shim_sam_from_gz(gzFile gz, int wr) {
    if (!gz) return NULL;
    gzbuffer(gz, 1 << 20);
    samFile *f = (samFile *)calloc(1, sizeof(samFile));
    f->gz = gz;
    f->is_write = wr;
    f->is_bam = 1;
    return f;
}
samFile *sam_open(const char *fn, const char *mode) {
    return shim_sam_from_gz(shim_gzopen(fn, mode), strchr(mode, 'w') != NULL);
}
int sam_close(samFile *fp) {
    if (!fp) return 0;
    int r = gzclose(fp->gz);
    free(fp);
    return r == Z_OK ? 0 : -1;
}
htsFile *hts_hopen(hFILE *hf, const char *fn, const char *mode) {
    htsFile *f = sam_open(fn, mode);
    if (f) hclose(hf); /* htslib takes ownership of the hFILE */
    return f;
}
int hts_close(htsFile *fp) { return sam_close(fp); }
int hts_set_thread_pool(htsFile *fp, htsThreadPool *p) {
    (void)fp; (void)p;
    return 0;
}

static int rd(gzFile gz, void *p, size_t n) { return gzread(gz, p, (unsigned)n) == (int)n ? 0 : -1; }
static int wr(gzFile gz, const void *p, size_t n) { return n == 0 || gzwrite(gz, p, (unsigned)n) == (int)n ? 0 : -1; }

sam_hdr_t *sam_hdr_read(samFile *fp) {
    char magic[4];
    int32_t l_text, n_ref;
    if (rd(fp->gz, magic, 4) || memcmp(magic, "BAM\1", 4)) return NULL;
    if (rd(fp->gz, &l_text, 4)) return NULL;
    sam_hdr_t *h = (sam_hdr_t *)calloc(1, sizeof(sam_hdr_t));
    h->l_text = (size_t)l_text;
    h->text = (char *)calloc((size_t)l_text + 1, 1);
    if (l_text && rd(fp->gz, h->text, (size_t)l_text)) { sam_hdr_destroy(h); return NULL; }
    if (rd(fp->gz, &n_ref, 4)) { sam_hdr_destroy(h); return NULL; }
    h->n_targets = n_ref;
    h->target_len = (uint32_t *)calloc((size_t)n_ref + 1, sizeof(uint32_t));
    h->target_name = (char **)calloc((size_t)n_ref + 1, sizeof(char *));
    for (int i = 0; i < n_ref; ++i) {
        int32_t l_name;
        if (rd(fp->gz, &l_name, 4)) { sam_hdr_destroy(h); return NULL; }
        h->target_name[i] = (char *)calloc((size_t)l_name + 1, 1);
        if (rd(fp->gz, h->target_name[i], (size_t)l_name) || rd(fp->gz, &h->target_len[i], 4)) {
            sam_hdr_destroy(h);
            return NULL;
        }
    }
    return h;
}

int sam_hdr_write(samFile *fp, const sam_hdr_t *h) {
    int32_t l_text = (int32_t)h->l_text, n_ref = h->n_targets;
    if (wr(fp->gz, "BAM\1", 4) || wr(fp->gz, &l_text, 4) || wr(fp->gz, h->text, h->l_text) || wr(fp->gz, &n_ref, 4))
        return -1;
    for (int i = 0; i < n_ref; ++i) {
        int32_t l_name = (int32_t)strlen(h->target_name[i]) + 1;
        if (wr(fp->gz, &l_name, 4) || wr(fp->gz, h->target_name[i], (size_t)l_name) || wr(fp->gz, &h->target_len[i], 4))
            return -1;
    }
    return 0;
}

void sam_hdr_destroy(sam_hdr_t *h) {
    if (!h) return;
    if (h->target_name)
        for (int i = 0; i < h->n_targets; ++i) free(h->target_name[i]);
    free(h->target_name);
    free(h->target_len);
    free(h->text);
    free(h);
}

int sam_read1(samFile *fp, sam_hdr_t *h, bam1_t *b) {
    (void)h;
    int32_t block_size;
    int n = gzread(fp->gz, &block_size, 4);
    if (n == 0) return -1; /* EOF */
    if (n != 4 || block_size < 32) return -2;
    uint8_t core[32];
    if (rd(fp->gz, core, 32)) return -2;
    int32_t refID, pos, l_seq, next_refID, next_pos, tlen;
    uint16_t bin, n_cigar, flag;
    memcpy(&refID, core, 4);
    memcpy(&pos, core + 4, 4);
    uint8_t l_read_name = core[8], mapq = core[9];
    memcpy(&bin, core + 10, 2);
    memcpy(&n_cigar, core + 12, 2);
    memcpy(&flag, core + 14, 2);
    memcpy(&l_seq, core + 16, 4);
    memcpy(&next_refID, core + 20, 4);
    memcpy(&next_pos, core + 24, 4);
    memcpy(&tlen, core + 28, 4);
    b->core.tid = refID; b->core.pos = pos; b->core.l_qname = l_read_name; b->core.qual = mapq;
    b->core.bin = bin; b->core.n_cigar = n_cigar; b->core.flag = flag; b->core.l_qseq = l_seq;
    b->core.mtid = next_refID; b->core.mpos = next_pos; b->core.isize = tlen; b->core.l_extranul = 0;
    int l_data = block_size - 32;
    if ((uint32_t)l_data > b->m_data) { /* htslib: realloc_bam_data rounds the capacity up to a power of two */
        uint32_t cap = shim_roundup32((uint32_t)l_data);
        b->data = (uint8_t *)realloc(b->data, (size_t)cap);
        b->m_data = cap;
    }
    b->l_data = l_data;
    if (l_data && rd(fp->gz, b->data, (size_t)l_data)) return -2;
    return 0;
}

int sam_write1(samFile *fp, const sam_hdr_t *h, const bam1_t *b) {
    (void)h;
    int32_t block_size = 32 + b->l_data;
    uint8_t core[32];
    int32_t refID = b->core.tid, pos = (int32_t)b->core.pos, l_seq = b->core.l_qseq, next_refID = b->core.mtid,
            next_pos = (int32_t)b->core.mpos, tlen = (int32_t)b->core.isize;
    uint16_t bin = b->core.bin, n_cigar = (uint16_t)b->core.n_cigar, flag = b->core.flag;
    memcpy(core, &refID, 4);
    memcpy(core + 4, &pos, 4);
    core[8] = (uint8_t)b->core.l_qname;
    core[9] = b->core.qual;
    memcpy(core + 10, &bin, 2);
    memcpy(core + 12, &n_cigar, 2);
    memcpy(core + 14, &flag, 2);
    memcpy(core + 16, &l_seq, 4);
    memcpy(core + 20, &next_refID, 4);
    memcpy(core + 24, &next_pos, 4);
    memcpy(core + 28, &tlen, 4);
    if (wr(fp->gz, &block_size, 4) || wr(fp->gz, core, 32) || wr(fp->gz, b->data, (size_t)b->l_data)) return -1;
    return block_size + 4;
}
