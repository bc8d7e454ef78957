/*
 * oracle/ref_codecs.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * Thin wrappers over the REAL third-party codecs the reference calls, as installed in this image:
 *   liblz4.so.1 (1.9.4)  <- lz4framed.compress(b)   /root/reference/snacc/pairwise_ncd.py:80
 *   libz.so.1   (1.3)    <- gzip.compress(b) / zlib.compress(b)   pairwise_ncd.py:74,78
 * They are used (a) to pin the C restatements in lz4_oracle.c / deflate_oracle.c, and (b) as the
 * timed CPU baseline ("kind": "reference") in bench.py: a pool of host threads each running the
 * reference's compressor calls on pre-loaded sequences (the codec-only ceiling of BASELINE.md 3.2).
 *
 * liblz4 ships without headers here, so the few prototypes needed are declared by hand and the
 * library is dlopen()ed.  LZ4F_preferences_t is 56 bytes in 1.9.x.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <dlfcn.h>
#include <pthread.h>
#include <zlib.h>

typedef struct {
    int blockSizeID, blockMode, contentChecksumFlag, frameType;
    unsigned long long contentSize;
    unsigned dictID;
    int blockChecksumFlag;
} lz4f_frameinfo;
typedef struct {
    lz4f_frameinfo frameInfo;
    int compressionLevel;
    unsigned autoFlush, favorDecSpeed, reserved[3];
} lz4f_prefs;

typedef size_t (*fn_bound)(size_t, const lz4f_prefs *);
typedef size_t (*fn_frame)(void *, size_t, const void *, size_t, const lz4f_prefs *);
typedef unsigned (*fn_iserr)(size_t);
typedef int (*fn_vernum)(void);

static fn_bound p_bound; static fn_frame p_frame; static fn_iserr p_iserr; static fn_vernum p_ver;
static pthread_once_t once = PTHREAD_ONCE_INIT;
static int lz4_ok = 0;

static void load_lz4(void)
{
    void *h = dlopen("liblz4.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    p_bound = (fn_bound)dlsym(h, "LZ4F_compressFrameBound");
    p_frame = (fn_frame)dlsym(h, "LZ4F_compressFrame");
    p_iserr = (fn_iserr)dlsym(h, "LZ4F_isError");
    p_ver = (fn_vernum)dlsym(h, "LZ4_versionNumber");
    lz4_ok = p_bound && p_frame && p_iserr && p_ver;
}

int ref_lz4_version(void) { pthread_once(&once, load_lz4); return lz4_ok ? p_ver() : -1; }
const char *ref_zlib_version(void) { return zlibVersion(); }

/* len(lz4framed.compress(src)); scratch may be NULL (allocated per call) */
int64_t ref_lz4f_size_flag(const uint8_t *src, uint64_t n, int content_size_flag)
{
    pthread_once(&once, load_lz4);
    if (!lz4_ok) return -1;
    lz4f_prefs prefs; memset(&prefs, 0, sizeof(prefs));
    if (content_size_flag) prefs.frameInfo.contentSize = n;
    size_t cap = p_bound((size_t)n, &prefs);
    void *dst = malloc(cap);
    if (!dst) return -2;
    size_t r = p_frame(dst, cap, src, (size_t)n, &prefs);
    free(dst);
    if (p_iserr(r)) return -3;
    return (int64_t)r;
}
int64_t ref_lz4f_size(const uint8_t *src, uint64_t n) { return ref_lz4f_size_flag(src, n, 1); }

/* raw deflate length exactly as CPython drives zlib: deflateInit2(level, Z_DEFLATED, wbits, 8, 0),
 * deflate(Z_FINISH) into a growing output buffer (the chunking does not influence the stream) */
int64_t ref_deflate_size(const uint8_t *src, uint64_t n, int level)
{
    z_stream zs; memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return -1;
    enum { CH = 1 << 16 };
    static __thread unsigned char out[CH];
    zs.next_in = (Bytef *)src;
    uint64_t left = n, total = 0;
    int ret;
    do {
        if (zs.avail_in == 0 && left) {
            uInt k = left > 0x40000000u ? 0x40000000u : (uInt)left;
            zs.avail_in = k; left -= k;
        }
        zs.next_out = out; zs.avail_out = CH;
        ret = deflate(&zs, left ? Z_NO_FLUSH : Z_FINISH);
        total += CH - zs.avail_out;
    } while (ret == Z_OK || ret == Z_BUF_ERROR);
    deflateEnd(&zs);
    return ret == Z_STREAM_END ? (int64_t)total : -2;
}

/* codec ids shared with include/snacc_b200.h: 0 = LZ4F, 1 = GZIP9, 2 = ZLIB6.
 * Returns len(compressed) as the reference's compressor call would produce it
 * (gzip: raw + 18, zlib: raw + 6). */
int64_t ref_compressed_len(const uint8_t *src, uint64_t n, int codec)
{
    switch (codec) {
    case 0: return ref_lz4f_size(src, n);
    case 1: { int64_t r = ref_deflate_size(src, n, 9); return r < 0 ? r : r + 18; }
    case 2: { int64_t r = ref_deflate_size(src, n, 6); return r < 0 ? r : r + 6; }
    }
    return -10;
}

/* ---- threaded batch of (x, y) jobs over pre-loaded sequences: the CPU baseline ---- */
typedef struct {
    const uint8_t *corpus; const uint64_t *offsets;   /* n_seqs + 1 offsets into corpus */
    const int32_t *job_x, *job_y;                     /* y < 0 -> single */
    int64_t *out; int n_jobs; int codec;
    volatile int next; pthread_mutex_t mu;
} batch_t;

static void *batch_worker(void *arg)
{
    batch_t *b = (batch_t *)arg;
    uint8_t *buf = NULL; uint64_t cap = 0;
    for (;;) {
        pthread_mutex_lock(&b->mu);
        int j = b->next++;
        pthread_mutex_unlock(&b->mu);
        if (j >= b->n_jobs) break;
        int x = b->job_x[j], y = b->job_y[j];
        uint64_t lx = b->offsets[x + 1] - b->offsets[x];
        uint64_t ly = y >= 0 ? b->offsets[y + 1] - b->offsets[y] : 0;
        if (lx + ly + 16 > cap) { cap = lx + ly + 16; buf = (uint8_t *)realloc(buf, cap); }
        memcpy(buf, b->corpus + b->offsets[x], lx);
        if (y >= 0) memcpy(buf + lx, b->corpus + b->offsets[y], ly);
        b->out[j] = ref_compressed_len(buf, lx + ly, b->codec);
    }
    free(buf);
    return NULL;
}

int ref_batch_sizes(const uint8_t *corpus, const uint64_t *offsets, const int32_t *job_x,
                    const int32_t *job_y, int n_jobs, int codec, int n_threads, int64_t *out)
{
    batch_t b = { corpus, offsets, job_x, job_y, out, n_jobs, codec, 0, PTHREAD_MUTEX_INITIALIZER };
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int i = 0; i < n_threads; i++) pthread_create(&th[i], NULL, batch_worker, &b);
    for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
    free(th);
    return 0;
}
