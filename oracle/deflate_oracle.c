/*
 * oracle/deflate_oracle.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * CPU restatement of the raw-deflate byte count behind
 *   /root/reference/snacc/pairwise_ncd.py:74   gzip.compress(sequence)  -> level 9, raw + 18
 *   /root/reference/snacc/pairwise_ncd.py:78   zlib.compress(sequence)  -> level 6, raw + 6
 * The arithmetic is in a third-party dependency that is not under /root/reference: CPython's
 * zlib module over the system zlib, here zlib 1.3 (memLevel 8, windowBits 15, default strategy).
 * This file restates zlib's published algorithm -- the lazy matcher (`deflate_slow`,
 * `longest_match`, the window refill schedule) and the block cost accounting (`_tr_tally`,
 * `_tr_flush_block`, `build_tree`, `gen_bitlen`, `scan_tree`, `build_bl_tree`) -- using absolute
 * stream positions instead of a sliding buffer.  tests/ pin it against the system libz
 * (oracle/ref_codecs.c) and against golden vectors in tests/golden/.
 *
 * Only bit COUNTS are produced; no bitstream is materialised.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>

#define WSIZE        32768u
#define WMASK        32767u
#define MIN_MATCH    3
#define MAX_MATCH    258
#define MIN_LOOKAHEAD (MAX_MATCH + MIN_MATCH + 1)      /* 262 */
#define MAX_DIST     (WSIZE - MIN_LOOKAHEAD)           /* 32506 */
#define TOO_FAR      4096
#define HASH_SIZE    32768u
#define LIT_BUFSIZE  16384u                            /* memLevel 8 */
#define SYMS_PER_BLOCK (LIT_BUFSIZE - 1)

#define L_CODES   286
#define D_CODES   30
#define BL_CODES  19
#define LITERALS  256
#define END_BLOCK 256
#define HEAP_SIZE (2 * L_CODES + 1)
#define MAX_BITS  15
#define MAX_BL_BITS 7
#define REP_3_6     16
#define REPZ_3_10   17
#define REPZ_11_138 18

static const int extra_lbits[29] = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
static const int extra_dbits[30] = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};
static const int extra_blbits[19] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,2,3,7};
static const uint8_t bl_order[19] = {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15};

typedef struct { uint16_t freq; uint16_t dad; uint16_t len; } node_t;

typedef struct {
    node_t *tree;
    int max_code;
    const uint8_t *static_len;   /* NULL for the bit-length tree */
    const int *extra_bits;
    int extra_base;
    int elems;
    int max_length;
} tree_desc;

typedef struct {
    node_t ltree[HEAP_SIZE];
    node_t dtree[2 * D_CODES + 1];
    node_t bltree[2 * BL_CODES + 1];
    int heap[HEAP_SIZE];
    int heap_len, heap_max;
    uint8_t depth[HEAP_SIZE];
    uint16_t bl_count[MAX_BITS + 1];
    uint64_t opt_len, static_len;
    uint32_t sym_count;
    uint8_t length_code[256];
    uint8_t static_llen[L_CODES + 2];
    uint8_t static_dlen[D_CODES];
    uint64_t bits;               /* total bits emitted so far */
} trees_t;

static int dist_code(uint32_t d /* distance - 1 */)
{
    if (d < 4) return (int)d;
    int k = 31 - __builtin_clz(d);
    return 2 * k + (int)((d >> (k - 1)) & 1);
}

static void trees_init(trees_t *t)
{
    memset(t, 0, sizeof(*t));
    int length = 0, code;
    for (code = 0; code < 28; code++)
        for (int n = 0; n < (1 << extra_lbits[code]); n++) t->length_code[length++] = (uint8_t)code;
    t->length_code[length - 1] = (uint8_t)code;      /* length 258 gets its own code 28 */
    for (int n = 0; n <= 143; n++) t->static_llen[n] = 8;
    for (int n = 144; n <= 255; n++) t->static_llen[n] = 9;
    for (int n = 256; n <= 279; n++) t->static_llen[n] = 7;
    for (int n = 280; n <= 287; n++) t->static_llen[n] = 8;
    for (int n = 0; n < D_CODES; n++) t->static_dlen[n] = 5;
}

static void init_block(trees_t *t)
{
    for (int n = 0; n < L_CODES; n++) t->ltree[n].freq = 0;
    for (int n = 0; n < D_CODES; n++) t->dtree[n].freq = 0;
    for (int n = 0; n < BL_CODES; n++) t->bltree[n].freq = 0;
    t->ltree[END_BLOCK].freq = 1;
    t->opt_len = t->static_len = 0;
    t->sym_count = 0;
}

#define SMALLER(tree, n, m) \
    (tree[n].freq < tree[m].freq || (tree[n].freq == tree[m].freq && t->depth[n] <= t->depth[m]))

static void pqdownheap(trees_t *t, node_t *tree, int k)
{
    int v = t->heap[k];
    int j = k << 1;
    while (j <= t->heap_len) {
        if (j < t->heap_len && SMALLER(tree, t->heap[j + 1], t->heap[j])) j++;
        if (SMALLER(tree, v, t->heap[j])) break;
        t->heap[k] = t->heap[j]; k = j;
        j <<= 1;
    }
    t->heap[k] = v;
}

static void gen_bitlen(trees_t *t, tree_desc *desc)
{
    node_t *tree = desc->tree;
    int max_code = desc->max_code;
    int max_length = desc->max_length;
    int h, n, m, bits, xbits, overflow = 0;

    for (bits = 0; bits <= MAX_BITS; bits++) t->bl_count[bits] = 0;
    tree[t->heap[t->heap_max]].len = 0;

    for (h = t->heap_max + 1; h < HEAP_SIZE; h++) {
        n = t->heap[h];
        bits = tree[tree[n].dad].len + 1;
        if (bits > max_length) bits = max_length, overflow++;
        tree[n].len = (uint16_t)bits;
        if (n > max_code) continue;
        t->bl_count[bits]++;
        xbits = 0;
        if (n >= desc->extra_base) xbits = desc->extra_bits[n - desc->extra_base];
        uint64_t f = tree[n].freq;
        t->opt_len += f * (unsigned)(bits + xbits);
        if (desc->static_len) t->static_len += f * (unsigned)(desc->static_len[n] + xbits);
    }
    if (overflow == 0) return;

    do {
        bits = max_length - 1;
        while (t->bl_count[bits] == 0) bits--;
        t->bl_count[bits]--;
        t->bl_count[bits + 1] += 2;
        t->bl_count[max_length]--;
        overflow -= 2;
    } while (overflow > 0);

    for (bits = max_length; bits != 0; bits--) {
        n = t->bl_count[bits];
        while (n != 0) {
            m = t->heap[--h];
            if (m > max_code) continue;
            if ((unsigned)tree[m].len != (unsigned)bits) {
                t->opt_len += ((uint64_t)bits - tree[m].len) * tree[m].freq;
                tree[m].len = (uint16_t)bits;
            }
            n--;
        }
    }
}

/* NOTE: zlib keeps `dad` and `len` in one union; gen_bitlen only reads tree[dad].len after the
 * father's len has been written (fathers precede sons in heap order), so separate fields are
 * equivalent. */
static void build_tree(trees_t *t, tree_desc *desc)
{
    node_t *tree = desc->tree;
    int elems = desc->elems;
    int n, m, max_code = -1, node;

    t->heap_len = 0; t->heap_max = HEAP_SIZE;
    for (n = 0; n < elems; n++) {
        if (tree[n].freq != 0) {
            t->heap[++(t->heap_len)] = max_code = n;
            t->depth[n] = 0;
        } else {
            tree[n].len = 0;
        }
    }
    while (t->heap_len < 2) {
        node = t->heap[++(t->heap_len)] = (max_code < 2 ? ++max_code : 0);
        tree[node].freq = 1;
        t->depth[node] = 0;
        t->opt_len--;
        if (desc->static_len) t->static_len -= desc->static_len[node];
    }
    desc->max_code = max_code;

    for (n = t->heap_len / 2; n >= 1; n--) pqdownheap(t, tree, n);

    node = elems;
    do {
        n = t->heap[1];
        t->heap[1] = t->heap[t->heap_len--];
        pqdownheap(t, tree, 1);
        m = t->heap[1];

        t->heap[--(t->heap_max)] = n;
        t->heap[--(t->heap_max)] = m;

        tree[node].freq = (uint16_t)(tree[n].freq + tree[m].freq);
        t->depth[node] = (uint8_t)((t->depth[n] >= t->depth[m] ? t->depth[n] : t->depth[m]) + 1);
        tree[n].dad = tree[m].dad = (uint16_t)node;
        t->heap[1] = node++;
        pqdownheap(t, tree, 1);
    } while (t->heap_len >= 2);

    t->heap[--(t->heap_max)] = t->heap[1];
    gen_bitlen(t, desc);
}

static void scan_tree(trees_t *t, node_t *tree, int max_code)
{
    int n, prevlen = -1, curlen, nextlen = tree[0].len, count = 0, max_count = 7, min_count = 4;
    if (nextlen == 0) max_count = 138, min_count = 3;
    tree[max_code + 1].len = 0xffff;
    for (n = 0; n <= max_code; n++) {
        curlen = nextlen; nextlen = tree[n + 1].len;
        if (++count < max_count && curlen == nextlen) continue;
        else if (count < min_count) t->bltree[curlen].freq += count;
        else if (curlen != 0) {
            if (curlen != prevlen) t->bltree[curlen].freq++;
            t->bltree[REP_3_6].freq++;
        } else if (count <= 10) t->bltree[REPZ_3_10].freq++;
        else t->bltree[REPZ_11_138].freq++;
        count = 0; prevlen = curlen;
        if (nextlen == 0) max_count = 138, min_count = 3;
        else if (curlen == nextlen) max_count = 6, min_count = 3;
        else max_count = 7, min_count = 4;
    }
}

/* close the current block: stored_len input bytes, can_store = block start still inside the window */
static void flush_block(trees_t *t, uint64_t stored_len, int can_store, int last)
{
    tree_desc ld = { t->ltree, 0, t->static_llen, extra_lbits, LITERALS + 1, L_CODES, MAX_BITS };
    tree_desc dd = { t->dtree, 0, t->static_dlen, extra_dbits, 0, D_CODES, MAX_BITS };
    tree_desc bd = { t->bltree, 0, NULL, extra_blbits, 0, BL_CODES, MAX_BL_BITS };
    int max_blindex;

    build_tree(t, &ld);
    build_tree(t, &dd);
    scan_tree(t, t->ltree, ld.max_code);
    scan_tree(t, t->dtree, dd.max_code);
    build_tree(t, &bd);
    for (max_blindex = BL_CODES - 1; max_blindex >= 3; max_blindex--)
        if (t->bltree[bl_order[max_blindex]].len != 0) break;
    t->opt_len += 3 * ((uint64_t)max_blindex + 1) + 5 + 5 + 4;

    uint64_t opt_lenb = (t->opt_len + 3 + 7) >> 3;
    uint64_t static_lenb = (t->static_len + 3 + 7) >> 3;
    if (static_lenb <= opt_lenb) opt_lenb = static_lenb;

    if (stored_len + 4 <= opt_lenb && can_store) {
        t->bits += 3;
        t->bits = (t->bits + 7) & ~7ull;
        t->bits += 32 + 8 * stored_len;
    } else if (static_lenb == opt_lenb) {
        t->bits += 3 + t->static_len;
    } else {
        t->bits += 3 + t->opt_len;
    }
    init_block(t);
    if (last) t->bits = (t->bits + 7) & ~7ull;
}

static inline int tally_lit(trees_t *t, uint8_t c)
{
    t->ltree[c].freq++;
    return ++t->sym_count == SYMS_PER_BLOCK;
}
static inline int tally_dist(trees_t *t, uint32_t dist, uint32_t len_minus3)
{
    t->ltree[t->length_code[len_minus3] + LITERALS + 1].freq++;
    t->dtree[dist_code(dist - 1)].freq++;
    return ++t->sym_count == SYMS_PER_BLOCK;
}

typedef struct { int good_length, max_lazy, nice_length, max_chain; } dconfig;
static const dconfig cfg6 = { 8, 16, 128, 128 };
static const dconfig cfg9 = { 32, 258, 258, 4096 };

typedef struct {
    const uint8_t *w; uint64_t n;
    int64_t *head;      /* absolute position of the newest string per hash, -1 = none */
    int64_t *prev;      /* ring of WSIZE absolute positions */
    uint64_t base;      /* absolute position of window index 0 */
    uint64_t read;      /* bytes pulled into the window so far */
    uint64_t strstart, match_start, prev_match;
    uint32_t match_length, prev_length;
    int match_available;
    uint64_t block_start;
} dstate;

static inline uint32_t hash3(const uint8_t *p)
{
    return (((uint32_t)p[0] << 10) ^ ((uint32_t)p[1] << 5) ^ p[2]) & (HASH_SIZE - 1);
}

/* zlib stores window-relative positions, where 0 doubles as NIL: the string at window index 0
 * can never be a candidate. */
#define IS_NIL(s, p) ((p) < 0 || (uint64_t)(p) <= (s)->base)

static void fill_window(dstate *s)
{
    for (;;) {
        uint64_t more = 2 * (uint64_t)WSIZE - (s->read - s->base);
        if (s->strstart - s->base >= WSIZE + MAX_DIST) {
            s->base += WSIZE;
            more += WSIZE;
        }
        if (s->read == s->n) break;
        uint64_t k = s->n - s->read; if (k > more) k = more;
        s->read += k;
        if (!(s->read - s->strstart < MIN_LOOKAHEAD && s->read != s->n)) break;
    }
}

static uint32_t longest_match(dstate *s, const dconfig *c, int64_t cur_match)
{
    unsigned chain_length = (unsigned)c->max_chain;
    const uint8_t *scan = s->w + s->strstart;
    int best_len = (int)s->prev_length;
    int nice_match = c->nice_length;
    uint64_t lookahead = s->read - s->strstart;
    /* candidates must be strictly above `limit` */
    int64_t limit = (s->strstart - s->base > MAX_DIST) ? (int64_t)(s->strstart - MAX_DIST) : (int64_t)s->base;
    uint64_t maxcmp = s->n - s->strstart; if (maxcmp > MAX_MATCH) maxcmp = MAX_MATCH;

    if (s->prev_length >= (uint32_t)c->good_length) chain_length >>= 2;
    if ((uint64_t)nice_match > lookahead) nice_match = (int)lookahead;

    do {
        const uint8_t *match = s->w + cur_match;
        /* bytes at or beyond the end of input compare as a mismatch here; zlib compares window
         * garbage there but then clips to lookahead and stops at nice_match <= lookahead, so the
         * outcome is identical. */
        int len = 0;
        while ((uint64_t)len < maxcmp && match[len] == scan[len]) len++;
        if (len > best_len) {
            s->match_start = (uint64_t)cur_match;
            best_len = len;
            if (len >= nice_match) break;
        }
        cur_match = s->prev[cur_match & WMASK];
    } while (cur_match > limit && --chain_length != 0);

    if ((uint64_t)best_len <= lookahead) return (uint32_t)best_len;
    return (uint32_t)lookahead;
}

uint64_t oracle_deflate_bits(const uint8_t *src, uint64_t n, int level,
                             uint64_t *n_blocks_out, uint64_t *n_syms_out)
{
    const dconfig *c = (level == 9) ? &cfg9 : &cfg6;
    trees_t *t = (trees_t *)malloc(sizeof(trees_t));
    dstate S, *s = &S;
    uint64_t nblocks = 0, nsyms = 0;
    memset(s, 0, sizeof(*s));
    trees_init(t); init_block(t);
    s->w = src; s->n = n;
    s->head = (int64_t *)malloc(sizeof(int64_t) * HASH_SIZE);
    s->prev = (int64_t *)malloc(sizeof(int64_t) * WSIZE);
    for (uint32_t i = 0; i < HASH_SIZE; i++) s->head[i] = -1;
    for (uint32_t i = 0; i < WSIZE; i++) s->prev[i] = -1;
    s->match_length = s->prev_length = MIN_MATCH - 1;

#define INSERT_STRING(pos, hh) do { uint32_t h_ = hash3(s->w + (pos)); \
        (hh) = s->head[h_]; s->prev[(pos) & WMASK] = (hh); s->head[h_] = (int64_t)(pos); } while (0)
#define FLUSH(last) do { flush_block(t, s->strstart - s->block_start, s->block_start >= s->base, (last)); \
        s->block_start = s->strstart; nblocks++; } while (0)

    for (;;) {
        int bflush;
        int64_t hash_head = -1;
        if (s->read - s->strstart < MIN_LOOKAHEAD) {
            fill_window(s);
            if (s->read == s->strstart) break;
        }
        uint64_t lookahead = s->read - s->strstart;
        if (lookahead >= MIN_MATCH) INSERT_STRING(s->strstart, hash_head);

        s->prev_length = s->match_length; s->prev_match = s->match_start;
        s->match_length = MIN_MATCH - 1;

        if (!IS_NIL(s, hash_head) && s->prev_length < (uint32_t)c->max_lazy &&
            s->strstart - (uint64_t)hash_head <= MAX_DIST) {
            s->match_length = longest_match(s, c, hash_head);
            if (s->match_length <= 5 &&
                (s->match_length == MIN_MATCH && s->strstart - s->match_start > TOO_FAR))
                s->match_length = MIN_MATCH - 1;
        }
        if (s->prev_length >= MIN_MATCH && s->match_length <= s->prev_length) {
            uint64_t max_insert = s->strstart + lookahead - MIN_MATCH;
            bflush = tally_dist(t, (uint32_t)(s->strstart - 1 - s->prev_match), s->prev_length - MIN_MATCH);
            nsyms++;
            s->prev_length -= 2;
            do {
                if (++s->strstart <= max_insert) { int64_t hh; INSERT_STRING(s->strstart, hh); (void)hh; }
            } while (--s->prev_length != 0);
            s->match_available = 0;
            s->match_length = MIN_MATCH - 1;
            s->strstart++;
            if (bflush) FLUSH(0);
        } else if (s->match_available) {
            bflush = tally_lit(t, s->w[s->strstart - 1]);
            nsyms++;
            if (bflush) FLUSH(0);
            s->strstart++;
        } else {
            s->match_available = 1;
            s->strstart++;
        }
    }
    if (s->match_available) { tally_lit(t, s->w[s->strstart - 1]); nsyms++; s->match_available = 0; }
    FLUSH(1);
#undef INSERT_STRING
#undef FLUSH
    uint64_t bits = t->bits;
    free(s->head); free(s->prev); free(t);
    if (n_blocks_out) *n_blocks_out = nblocks;
    if (n_syms_out) *n_syms_out = nsyms;
    return bits;
}

/* raw deflate stream length in bytes */
uint64_t oracle_deflate_size(const uint8_t *src, uint64_t n, int level)
{
    return oracle_deflate_bits(src, n, level, NULL, NULL) >> 3;
}
