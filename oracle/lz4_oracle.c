/*
 * oracle/lz4_oracle.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * CPU restatement of the byte count produced by the call the reference makes at
 *   /root/reference/snacc/pairwise_ncd.py:80   compressed_seq = lz4framed.compress(sequence)
 * i.e. py-lz4framed's one-shot compress -> LZ4F_compressFrame with the binding's
 * defaults {blockSizeID = default (64 KiB), blockMode = linked, no checksums,
 * contentSize = len(sequence), compressionLevel = 0}.
 *
 * The arithmetic lives in a third-party dependency that is NOT under /root/reference:
 * py-lz4framed (requirements.txt:6, unpinned) which bundles the LZ4 C library.  The oracle is
 * anchored on the LZ4 library installed in this image, liblz4.so.1 = 1.9.4: this file restates
 * the published LZ4 frame format and the `fast` (level 0, acceleration 1) block compressor of
 * that release, and tests/ pins it against liblz4.so.1.9.4 itself (oracle/ref_codecs.c) on
 * sweeps over every regime boundary and against golden vectors in tests/golden/.
 *
 * Everything here only COUNTS bytes; no compressed output is materialised.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>

#define LZ4_BLOCK      65536u      /* default frame block size (blockSizeID 0 -> 64 KiB) */
#define MINMATCH       4
#define MFLIMIT        12
#define LASTLITERALS   5
#define LZ4_MINLENGTH  (MFLIMIT + 1)
#define MAX_DISTANCE   65535u
#define SKIP_TRIGGER   6
#define ML_MASK        15u
#define RUN_MASK       15u

static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

/* single-block frames: 16-bit position table, 13-bit index, 4-byte multiplicative hash */
static inline uint32_t hash4_u16(uint32_t seq) { return (seq * 2654435761u) >> (32 - 13); }
/* linked multi-block frames on a 64-bit little-endian host: 32-bit table, 12-bit index, 5-byte hash */
static inline uint32_t hash5_u32(uint64_t seq) {
    return (uint32_t)(((seq << 24) * 889523592379ull) >> (64 - 12));
}

/* number of equal bytes between a and b, never reading a at or beyond `limit` */
static inline uint32_t count_eq(const uint8_t *a, const uint8_t *b, const uint8_t *limit) {
    const uint8_t *s = a;
    while (a < limit && *a == *b) { a++; b++; }
    return (uint32_t)(a - s);
}

typedef struct {
    uint32_t table[8192];     /* big enough for both regimes (8192 x u16-equivalent or 4096 x u32) */
    uint32_t current_offset;  /* stream position of the next block's first byte */
} lz4_state;

/*
 * One block of the `fast` compressor.
 *   base   : start of the whole stream (positions in the table are offsets from base)
 *   start  : stream offset of this block's first byte; n: block length
 *   u16    : 1 = single-block regime (hash4/13 bits, no distance test), 0 = linked regime
 *   budget : output capacity the frame layer grants (n - 1); exceeding it aborts the block
 * returns compressed payload bytes, or 0 when the block aborts (frame layer then stores it raw).
 * The table keeps whatever was inserted up to the abort point, as in the library.
 */
static uint32_t lz4_fast_block(lz4_state *st, const uint8_t *base, uint32_t start, uint32_t n,
                               int u16, uint32_t budget)
{
    const uint8_t *src = base + start;
    const uint8_t *ip = src;
    const uint8_t *anchor = src;
    const uint8_t *iend = src + n;
    const uint8_t *mflimit_plus1 = iend - MFLIMIT + 1;
    const uint8_t *matchlimit = iend - LASTLITERALS;
    const uint8_t *low_limit = u16 ? src : base;   /* prefix mode: everything before is reachable */
    uint64_t op = 0;                               /* bytes emitted so far */
    uint32_t *tab = st->table;
    uint32_t forward_h;

    st->current_offset = start + n;

#define HASH_AT(p) (u16 ? hash4_u16(rd32(p)) : hash5_u32(rd64(p)))

    if (n < LZ4_MINLENGTH) goto last_literals;

    tab[HASH_AT(ip)] = (uint32_t)(ip - base);
    ip++;
    forward_h = HASH_AT(ip);

    for (;;) {
        const uint8_t *match;
        uint64_t token_pos;
        /* search loop with skip acceleration */
        {
            const uint8_t *forward_ip = ip;
            uint32_t step = 1;
            uint32_t search_nb = 1u << SKIP_TRIGGER;   /* acceleration 1 */
            for (;;) {
                uint32_t h = forward_h;
                uint32_t current = (uint32_t)(forward_ip - base);
                uint32_t match_index = tab[h];
                ip = forward_ip;
                forward_ip += step;
                step = (search_nb++ >> SKIP_TRIGGER);
                if (forward_ip > mflimit_plus1) goto last_literals;
                match = base + match_index;
                forward_h = HASH_AT(forward_ip);
                tab[h] = current;
                if (!u16 && match_index + MAX_DISTANCE < current) continue;   /* too far */
                if (rd32(match) == rd32(ip)) break;
            }
        }
        /* catch up: extend the match backwards over pending literals */
        while (ip > anchor && match > low_limit && ip[-1] == match[-1]) { ip--; match--; }

        /* literal run */
        {
            uint32_t lit = (uint32_t)(ip - anchor);
            token_pos = op; op++;
            if (op + lit + (2 + 1 + LASTLITERALS) + (lit / 255) > budget) return 0;
            if (lit >= RUN_MASK) {
                uint32_t len = lit - RUN_MASK;
                op += len / 255 + 1;
            }
            op += lit;
        }
        (void)token_pos;
next_match:
        op += 2;   /* offset */
        {
            uint32_t mcode = count_eq(ip + MINMATCH, match + MINMATCH, matchlimit);
            ip += (size_t)mcode + MINMATCH;
            if (op + (1 + LASTLITERALS) + (mcode + 240) / 255 > budget) return 0;
            if (mcode >= ML_MASK) {
                mcode -= ML_MASK;
                op += mcode / 255 + 1;
            }
        }
        anchor = ip;
        if (ip >= mflimit_plus1) break;

        tab[HASH_AT(ip - 2)] = (uint32_t)(ip - 2 - base);

        /* immediate re-test at the new position */
        {
            uint32_t h = HASH_AT(ip);
            uint32_t current = (uint32_t)(ip - base);
            uint32_t match_index = tab[h];
            match = base + match_index;
            tab[h] = current;
            if ((u16 || match_index + MAX_DISTANCE >= current) && rd32(match) == rd32(ip)) {
                op++;   /* token with zero literals */
                goto next_match;
            }
        }
        forward_h = HASH_AT(++ip);
    }

last_literals:
    {
        uint64_t last_run = (uint64_t)(iend - anchor);
        if (op + last_run + 1 + ((last_run + 255 - RUN_MASK) / 255) > budget) return 0;
        if (last_run >= RUN_MASK) op += 1 + (last_run - RUN_MASK) / 255 + 1;
        else op += 1;
        op += last_run;
    }
#undef HASH_AT
    return (uint32_t)op;
}

/*
 * Frame size, with optional per-block payload sizes (block_sizes may be NULL; raw-stored blocks
 * are reported with bit 31 set, like the on-wire block header).
 * content_size_flag: 1 = header carries the 8-byte content size (py-lz4framed default assumed by
 * SURVEY.md section 8c), 0 = it does not.
 *
 * The caller must provide 8 readable bytes of slack after src[n-1] (hash5 reads 8 bytes; the
 * library has the same property because it only hashes positions <= n - 12).
 */
uint64_t oracle_lz4f_size_ex(const uint8_t *src, uint64_t n, int content_size_flag,
                             uint32_t *block_sizes, uint32_t max_blocks)
{
    uint64_t total = 4 /*magic*/ + 1 /*FLG*/ + 1 /*BD*/ + 1 /*HC*/;
    if (content_size_flag && n != 0) total += 8;   /* contentSize == 0 means "unknown": no field */
    lz4_state *st = (lz4_state *)calloc(1, sizeof(lz4_state));
    uint32_t nb = 0;
    if (n <= LZ4_BLOCK) {
        /* one-shot rule: a frame that fits one block is switched to independent blocks and is
         * compressed with the 16-bit table (len < 65547) */
        if (n > 0) {
            uint32_t c = lz4_fast_block(st, src, 0, (uint32_t)n, 1, (uint32_t)n - 1);
            uint32_t stored = (c == 0 || c >= n);
            uint32_t sz = stored ? (uint32_t)n : c;
            total += 4 + sz;
            if (block_sizes && nb < max_blocks) block_sizes[nb] = sz | (stored ? 0x80000000u : 0);
            nb++;
        }
    } else {
        uint64_t pos = 0;
        while (pos < n) {
            uint32_t len = (n - pos >= LZ4_BLOCK) ? LZ4_BLOCK : (uint32_t)(n - pos);
            uint32_t c = lz4_fast_block(st, src, (uint32_t)pos, len, 0, len - 1);
            uint32_t stored = (c == 0 || c >= len);
            uint32_t sz = stored ? len : c;
            total += 4 + sz;
            if (block_sizes && nb < max_blocks) block_sizes[nb] = sz | (stored ? 0x80000000u : 0);
            nb++;
            pos += len;
        }
    }
    total += 4;   /* end mark */
    free(st);
    return total;
}

uint64_t oracle_lz4f_size(const uint8_t *src, uint64_t n)
{
    return oracle_lz4f_size_ex(src, n, 1, NULL, 0);
}
