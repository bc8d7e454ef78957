// lz4_packed.cuh -- LZ4-frame compressed-size kernels on 2-bit packed sequences (K1/K2 of SURVEY.md 2.3),
// sm_100a.  This is the fast path of len(lz4framed.compress(b)) (reference call site
// snacc/pairwise_ncd.py:80) for corpora whose sequences use at most four distinct byte values
// (A/C/G/T genomes); everything else takes the byte-wise kernels of lz4.cuh.
//
// Why packing changes the design (DESIGN.md "LZ4 tile kernel"):
//   * with <= 4 symbols the 5-byte (linked regime) / 4-byte (single-block regime) hash of LZ4 only ever
//     sees 1024 / 256 distinct k-mers, so the 4096 x u32 / 8192 x u16 library table collapses to a table
//     with one slot per distinct bucket the k-mers reach (a 1024-entry code -> slot map keeps k-mers that
//     collide in the library's table colliding here too).  4 KiB (512 B) of state per stream
//     instead of 16 KiB -> the table lives in SHARED MEMORY, bank-conflict free (entry e of lane l is word
//     e*LANES + l);
//   * 32 bases fit one 64-bit word: hash input, the 4-byte match test and the match extension are one
//     XOR + count-trailing-zeros;
//   * all streams of a CTA are pair jobs (x_i, y) that share y: y is staged once per CTA in a 128 Ki-base
//     shared-memory ring (64 Ki back-reference window + read-ahead) and every stream resumes from the
//     prefix checkpoint of its own x, taken right before the parse first looks at a byte of y.  Reads
//     that fall outside the ring (early back-references into x, very long matches) take an exact
//     global-memory path, so the ring is only ever a cache.
//
// The parse is the same flat probe loop as lz4.cuh (validated against liblz4 1.9.4), expressed as a
// resumable state machine (PkState) so that a stream can stop at a ring refill and so that the prefix
// pass can stop, without side effects, at the first iteration whose outcome depends on bytes >= xend.
#pragma once
#include "common.cuh"
#include "lz4.cuh"
#include <type_traits>

namespace snacc {

constexpr uint32_t PK_RING_WORDS = 4096;                 // 32 KiB: 131072 bases
constexpr uint32_t PK_RING_BASES = PK_RING_WORDS * 32;
constexpr uint32_t PK_RING_BYTES = PK_RING_WORDS * 8 + 16;   // + mirror of the first two words (pk_turbo_lean reads up to 3 x u32 from one masked address)
constexpr uint32_t PK_CHUNK_BASES = 32768;               // refill granularity (keeps >= 96 Ki bases behind)
constexpr uint32_t PK_GUARD = 128;                       // streams stop this far before the ring's end
constexpr uint32_t PK_ABORT = 0xffffffffu;

enum : uint32_t { PK_SEARCH = 0, PK_RETEST = 1, PK_BLOCK_START = 2, PK_DONE = 3 };

// Parser state between two iterations of the probe loop.  Everything a stream needs to resume.
struct PkState {
    uint32_t ip, fip, step, nb;      // probe position (re-test mode) / forward position + skip counters (search mode)
    uint32_t anchor, op;             // start of pending literals; payload bytes of the open block (PK_ABORT = aborted)
    uint32_t bs, be, mfl1, mlim;     // open block [bs, be), mflimit+1, matchlimit
    uint32_t budget, max_lhs;        // blen-1; largest left-hand side any output-budget test has seen in this block
    uint32_t phase, pad_;
    uint64_t total;                  // sum of (4 + stored payload) over closed blocks
};

// What one stream sees of its input: y through the shared ring, everything else through global memory.
struct PkView {
    const uint64_t *ring;            // PK_RING_WORDS words holding y offsets [rlo, rlo + rspan + 64)
    uint32_t rlo, rspan;             // fast path iff (q - rlo) <= rspan (unsigned), q = p - lx
    const uint64_t *xw, *yw;         // packed x / y in global memory (zero padded)
    uint32_t lx;
    // sequences with bytes outside the alphabet (EXC variants only; see pk_step_exact):
    const uint32_t *dring;           // dirty ring: bit g & 31 of word (g >> 5) & (PK_DRING_WORDS - 1) is set when one of the 16
                                     // bases of y granule g (y offsets [16 g, 16 g + 16)) is flagged; word PK_DRING_WORDS mirrors word 0
    const uint32_t *ymask;           // one bit per base of y (global memory): the base is outside the alphabet
};
constexpr uint32_t PK_DRING_WORDS = PK_RING_WORDS * 32 / 16 / 32;      // 256 words: one bit per 16 bases of the ring
constexpr uint32_t PK_NONE = 0xffffffffu;

SNACC_HD uint32_t pk_ctz64(uint64_t d)
{
#ifdef __CUDA_ARCH__
    return (uint32_t)__ffsll((long long)d) - 1;
#else
    return (uint32_t)__builtin_ctzll(d);
#endif
}

SNACC_HD uint64_t pk_read(const uint64_t *w, uint32_t k)
{
    const uint32_t i = k >> 5, s = (k & 31) * 2;
    const uint64_t a = SNACC_LDG(w + i), b = SNACC_LDG(w + i + 1);
    return (a >> s) | ((b << 1) << (63 - s));
}

// exact path for anything outside the ring
static __host__ __device__ __noinline__ uint64_t pk_get32_slow(const PkView &v, uint32_t p)
{
    if (p >= v.lx) return pk_read(v.yw, p - v.lx);
    const uint64_t a = pk_read(v.xw, p);
    const uint32_t k = v.lx - p;                 // bases of x available from p
    if (k >= 32) return a;
    const uint64_t b = SNACC_LDG(v.yw);          // straddles the x|y boundary
    return (a & ((1ull << (2 * k)) - 1)) | (b << (2 * k));
}

// 32 bases of the stream starting at position p (base p in bits 0-1)
SNACC_HD uint64_t pk_get32(const PkView &v, uint32_t p)
{
    const uint32_t q = p - v.lx;
    if ((uint32_t)(q - v.rlo) <= v.rspan) {
        const uint32_t i = q >> 5, s = (q & 31) * 2;
        const uint64_t a = v.ring[i & (PK_RING_WORDS - 1)], b = v.ring[(i + 1) & (PK_RING_WORDS - 1)];
        return (a >> s) | ((b << 1) << (63 - s));
    }
    return pk_get32_slow(v, p);
}

// Position table.  `lut` maps a k-mer code to its slot: codes whose real bytes fall into the same bucket
// of the library's hash table share a slot (pack.cuh: pk_slot_lut), so the table behaves exactly like
// the library's while holding at most 1024 (256) live entries.  Three storage kinds:
//   KIND 0  linked regime, 32-bit positions                                  (4 B per slot)
//   KIND 1  single-block regime, 16-bit positions (streams <= 64 KiB)         (2 B per slot)
//   KIND 2  linked regime, 17 bits per slot: the low 16 bits of the position plus one bit in a separate bit plane
//           that says "written during the current 64 Ki epoch"              (2 B + 1 bit per slot)
// KIND 2 is what lets twice as many streams stay resident per SM.  An epoch is one LZ4 block: blocks start at
// multiples of 65536 of the stream position, so all positions probed between two block starts share their upper
// bits.  A slot with its bit set was written at (epoch base | low16) <= the probe: always within reach.  A slot
// with its bit clear was written during the PREVIOUS epoch at (epoch base - 65536 + low16), which is within 65535
// of a probe at p iff low16 > (p & 0xffff).  At every block start (new_epoch) the slots whose bit is still clear
// -- not written for a whole epoch, so at least 65536 behind every later probe -- are retired to low16 = 0, which the
// rule above never accepts, and all bits are cleared.  Inserts only ever SET a bit (one shared-memory atomic OR, no
// read-modify-write), and nothing has to be swept in the inner loop.
template <int KIND, int STRIDE> struct PkTab {
    static constexpr bool U16 = KIND == 1;
    typedef typename std::conditional<KIND == 0, uint32_t, uint16_t>::type T;
    static constexpr uint32_t K = U16 ? 4 : 5;
    static constexpr uint32_t MASK = U16 ? 0xffu : 0x3ffu;
    static constexpr uint32_t ENTRIES = MASK + 1;
    static constexpr uint32_t ESZ = STRIDE * sizeof(T);   // bytes between two slots of one lane
    T *t;
    uint32_t *ep;                                         // KIND 2: bit plane, word w of this lane at ep[w * STRIDE]
    uint32_t nslot;                                       // KIND 2: slots actually stored
    uint32_t epoch_base;                                  // KIND 2: start of the epoch the bits refer to (a block start)
    const uint16_t *lut;                                  // code -> slot index
    SNACC_HD uint32_t slot(uint32_t c) const { return lut[c]; }
    // candidate of code c for a probe at position ip; false: nothing within reach
    SNACC_HD bool lookup(uint32_t c, uint32_t ip, uint32_t &m) const { return lookup_idx(slot(c), ip, m); }
    SNACC_HD void put(uint32_t c, uint32_t pos) { put_idx(slot(c), pos); }
    SNACC_HD bool lookup_idx(uint32_t idx, uint32_t ip, uint32_t &m) const
    {
        if (KIND == 0) { m = t[idx * STRIDE]; return m + LZ4_MAX_DISTANCE >= ip; }
        if (KIND == 1) { m = t[idx * STRIDE]; return true; }
        const uint32_t v = t[idx * STRIDE];
        const uint32_t cur = (ep[(idx >> 5) * STRIDE] >> (idx & 31)) & 1;
        m = ip - ((ip - v) & 0xffffu);
        return cur ? m != ip : v > (ip & 0xffffu);
    }
    SNACC_HD void put_idx(uint32_t idx, uint32_t pos)
    {
        t[idx * STRIDE] = (T)pos;
        if (KIND == 2) ep[(idx >> 5) * STRIDE] |= 1u << (idx & 31);
    }
    // KIND 2, at a block start: retire what was not written during the epoch that ends, clear the bits
    SNACC_HD void new_epoch(uint32_t base)
    {
        if constexpr (KIND == 2) {
            for (uint32_t w = 0; w * 32 < nslot; ++w) {
                uint32_t z = ~ep[w * STRIDE];
                if (nslot - w * 32 < 32) z &= (1u << (nslot - w * 32)) - 1;
                for (; z; z &= z - 1) t[(w * 32 + (uint32_t)SNACC_FFS32(z) - 1) * STRIDE] = 0;
                ep[w * STRIDE] = 0;
            }
            epoch_base = base;
        }
    }
    // KIND 2: store an absolute position coming from a 32-bit checkpoint table, for a stream whose open block
    // (= current epoch) starts at bs.  (The epoch word of a slot is shared by 32 slots: callers import a whole word
    // from one thread, or clear the words first.)
    SNACC_HD void import_slot(uint32_t idx, uint32_t pos, uint32_t bs)
    {
        if (KIND != 2) { t[idx * STRIDE] = (T)pos; return; }
        const bool cur = pos >= bs, prev = !cur && pos + 65536u >= bs;
        t[idx * STRIDE] = (T)((cur || prev) ? pos : 0u);
        uint32_t &w = ep[(idx >> 5) * STRIDE];
        w = (w & ~(1u << (idx & 31))) | ((cur ? 1u : 0u) << (idx & 31));
    }
};

SNACC_HD void pk_end_block(PkState &st)
{
    const uint32_t blen = st.be - st.bs;
    uint32_t op = st.op;
    if (op != PK_ABORT) {
        const uint32_t last_run = st.be - st.anchor;
        const uint32_t lhs = op + last_run + 1 + (last_run + 240) / 255;
        st.max_lhs = tmax(st.max_lhs, lhs);
        if (lhs > st.budget) op = PK_ABORT;
        else op += 1 + (last_run >= 15 ? (last_run - 15) / 255 + 1 : 0) + last_run;
    }
    const uint32_t payload = (op == PK_ABORT || op >= blen) ? blen : op;
    st.total += 4 + payload;
    st.bs = st.be;
    st.phase = PK_BLOCK_START;
}

SNACC_HD void pk_fresh(PkState &st)
{
    st = PkState();
    st.phase = PK_BLOCK_START;
}

// position the next iteration will look at first (for the ring-refill stop test)
SNACC_HD uint32_t pk_next_pos(const PkState &st)
{
    return st.phase == PK_SEARCH ? st.fip : st.phase == PK_RETEST ? st.ip : st.bs;
}

// One iteration of the probe loop (or one block start).  `n`: stream length (0xffffffff while the
// prefix pass pretends the stream goes on).  DETECT: return true -- leaving state and table exactly as
// they were before the call -- when the iteration would depend on a base at or beyond `xend`.
template <int KIND, int STRIDE, bool DETECT>
SNACC_HD bool pk_step(PkState &st, PkTab<KIND, STRIDE> &tab, const PkView &v, uint32_t n, uint32_t xend)
{
    constexpr uint32_t K = PkTab<KIND, STRIDE>::K, MASK = PkTab<KIND, STRIDE>::MASK;
    if (st.phase == PK_BLOCK_START) {
        if (st.bs >= n) { st.phase = PK_DONE; return false; }
        // KIND 2: a block start is an epoch boundary of the table (the pair kernel has usually done this already,
        // cooperatively -- pk_run -- and then epoch_base == bs)
        if (KIND == 2 && tab.epoch_base != st.bs) tab.new_epoch(st.bs);
        const uint32_t be = (n - st.bs > LZ4_BLOCK) ? st.bs + LZ4_BLOCK : n;
        const uint32_t blen = be - st.bs;
        if (DETECT && st.bs + K > xend) return true;
        st.be = be; st.budget = blen - 1; st.anchor = st.bs; st.op = 0; st.max_lhs = 0;
        if (blen >= LZ4_MINLENGTH) {
            st.mfl1 = be - LZ4_MFLIMIT + 1; st.mlim = be - LZ4_LASTLITERALS;
            tab.put((uint32_t)pk_get32(v, st.bs) & MASK, st.bs);
            st.ip = st.bs; st.fip = st.bs + 1; st.step = 1; st.nb = 64;
            st.phase = PK_SEARCH;
        } else {
            pk_end_block(st);
        }
        return false;
    }
    const bool searching = st.phase == PK_SEARCH;
    PkState pre;
    if (DETECT) pre = st;
    uint32_t ip;
    if (searching) {
        ip = st.fip; st.fip += st.step; st.step = (st.nb++ >> 6);
        if (DETECT && st.fip > xend) { st = pre; return true; }
        if (st.fip > st.mfl1) { pk_end_block(st); return false; }
    } else {
        ip = st.ip;
    }
    if (DETECT && ip + K > xend) { st = pre; return true; }
    uint64_t w = pk_get32(v, ip);
    const uint32_t c = (uint32_t)w & MASK;
    uint32_t m;
    bool hit = tab.lookup(c, ip, m);
    const uint32_t old_m = m;
    tab.put(c, ip);
    uint32_t common = 0;
    if (hit) {
        const uint64_t d = w ^ pk_get32(v, m);
        common = d ? (pk_ctz64(d) >> 1) : 32;
        hit = common >= 4;
    }
    if (hit) {
        uint32_t op = st.op;
        if (searching) {
            bool moved = false;
            while (ip > st.anchor && m > 0 && (((uint32_t)pk_get32(v, ip - 1) ^ (uint32_t)pk_get32(v, m - 1)) & 3) == 0) {
                --ip; --m; moved = true;
            }
            const uint32_t lit = ip - st.anchor;
            op += 1;
            const uint32_t lhs = op + lit + 8 + lit / 255;
            st.max_lhs = tmax(st.max_lhs, lhs);
            if (lhs > st.budget) { st.op = PK_ABORT; pk_end_block(st); return false; }
            if (lit >= 15) op += (lit - 15) / 255 + 1;
            op += lit;
            if (moved) {
                w = pk_get32(v, ip);
                const uint64_t d = w ^ pk_get32(v, m);
                common = d ? (pk_ctz64(d) >> 1) : 32;
            }
        } else {
            op += 1;                               // token with zero literals
        }
        op += 2;                                   // offset
        uint32_t lim = st.mlim;
        if (DETECT) lim = tmin(lim, xend);
        if (common >= 32) {                        // long match: keep comparing 32 bases at a time
            while (ip + common < lim) {
                const uint64_t d = pk_get32(v, ip + common) ^ pk_get32(v, m + common);
                if (d) { common += pk_ctz64(d) >> 1; break; }
                common += 32;
            }
        }
        const uint32_t mlen = tmin(common, lim - ip);
        const uint32_t ip2 = ip + mlen;
        if (DETECT && (ip2 >= xend || (ip2 < st.mfl1 && ip2 + K - 2 > xend))) {
            tab.put(c, old_m); st = pre; return true;
        }
        const uint32_t mcode = mlen - 4;
        const uint32_t lhs2 = op + 6 + (mcode + 240) / 255;
        st.max_lhs = tmax(st.max_lhs, lhs2);
        if (lhs2 > st.budget) { st.op = PK_ABORT; pk_end_block(st); return false; }
        if (mcode >= 15) op += (mcode - 15) / 255 + 1;
        st.op = op; st.anchor = ip2; st.ip = ip2;
        if (ip2 >= st.mfl1) { pk_end_block(st); return false; }
        const uint32_t off = mlen - 2;
        const uint32_t c2 = (off + K <= 32) ? (uint32_t)(w >> (2 * off)) & MASK : (uint32_t)pk_get32(v, ip2 - 2) & MASK;
        tab.put(c2, ip2 - 2);
        st.phase = PK_RETEST;                      // immediate re-test at ip2
    } else if (!searching) {
        st.phase = PK_SEARCH; st.fip = ip + 1; st.step = 1; st.nb = 64;
    }
    return false;
}

// ---- byte-exact step: sequences with bytes outside the 4-symbol alphabet ------------------------------------------
// A packed sequence may hold a few bytes outside the corpus alphabet (N, IUPAC codes, soft-masked lower case): in the
// 2-bit text they carry a filler code and are flagged in a one-bit-per-base mask.  The fast loop never runs with such a
// base inside its window (the streams stop ahead of it) and treats one inside a candidate's window as a mismatch
// (pk_turbo_lean, EXC); everything else -- the probes around such a base -- takes this step, which works on the TRUE
// bytes of the stream (the ASCII corpus in global memory): LZ4's hash of the five real bytes selects the bucket; a
// bucket that some alphabet k-mer reaches is that k-mer's slot of the shared-memory table (so the two kinds of k-mers
// collide exactly as in the library's table), any other bucket lives in a per-stream overflow table in global memory.
// Same state machine, same DETECT contract as pk_step.
struct PkExact {
    Stream s;                 // the true bytes: x then y
    const uint16_t *b2s;      // LZ4 hash bucket (12 bits) -> slot index, 0xffff: no alphabet k-mer reaches this bucket
    uint32_t *ovf;            // this stream's table for those buckets: 4096 absolute positions (0 = position 0, as in the library)
};
constexpr uint32_t PK_OVF_ENTRIES = 4096;

SNACC_HD uint32_t pkx_bucket(const PkExact &xv, uint32_t p)
{
    return (uint32_t)(((ld64(xv.s, p) << 24) * 889523592379ull) >> (64 - 12));      // LZ4_hash5 on little-endian 64-bit
}
template <int KIND, int STRIDE>
SNACC_HD bool pkx_lookup(const PkTab<KIND, STRIDE> &tab, const PkExact &xv, uint32_t bucket, uint32_t ip, uint32_t &m)
{
    const uint32_t idx = SNACC_LDG(xv.b2s + bucket);
    if (idx != 0xffffu) return tab.lookup_idx(idx, ip, m);
    m = xv.ovf[bucket];
    return m + LZ4_MAX_DISTANCE >= ip;
}
template <int KIND, int STRIDE>
SNACC_HD void pkx_put(PkTab<KIND, STRIDE> &tab, const PkExact &xv, uint32_t bucket, uint32_t pos)
{
    const uint32_t idx = SNACC_LDG(xv.b2s + bucket);
    if (idx != 0xffffu) tab.put_idx(idx, pos);
    else xv.ovf[bucket] = pos;
}
// equal bytes from stream positions a and b forwards, at most max
SNACC_HD uint32_t pkx_common(const PkExact &xv, uint32_t a, uint32_t b, uint32_t max)
{
    uint32_t l = 0;
    while (l < max) {
        const uint64_t d = ld64(xv.s, a + l) ^ ld64(xv.s, b + l);
        if (d) { l += (uint32_t)(SNACC_FFS64(d) - 1) >> 3; break; }
        l += 8;
    }
    return tmin(l, max);
}

template <int KIND, int STRIDE, bool DETECT>
SNACC_HD bool pk_step_exact(PkState &st, PkTab<KIND, STRIDE> &tab, const PkExact &xv, uint32_t n, uint32_t xend)
{
    static_assert(KIND != 1, "the byte-exact step covers the linked regime (5-byte hash) only");
    constexpr uint32_t K = 5;
    if (st.phase == PK_BLOCK_START) {
        if (st.bs >= n) { st.phase = PK_DONE; return false; }
        if (KIND == 2 && tab.epoch_base != st.bs) tab.new_epoch(st.bs);
        const uint32_t be = (n - st.bs > LZ4_BLOCK) ? st.bs + LZ4_BLOCK : n;
        const uint32_t blen = be - st.bs;
        if (DETECT && st.bs + K > xend) return true;
        st.be = be; st.budget = blen - 1; st.anchor = st.bs; st.op = 0; st.max_lhs = 0;
        if (blen >= LZ4_MINLENGTH) {
            st.mfl1 = be - LZ4_MFLIMIT + 1; st.mlim = be - LZ4_LASTLITERALS;
            pkx_put(tab, xv, pkx_bucket(xv, st.bs), st.bs);
            st.ip = st.bs; st.fip = st.bs + 1; st.step = 1; st.nb = 64;
            st.phase = PK_SEARCH;
        } else {
            pk_end_block(st);
        }
        return false;
    }
    const bool searching = st.phase == PK_SEARCH;
    PkState pre;
    if (DETECT) pre = st;
    uint32_t ip;
    if (searching) {
        ip = st.fip; st.fip += st.step; st.step = (st.nb++ >> 6);
        if (DETECT && st.fip > xend) { st = pre; return true; }
        if (st.fip > st.mfl1) { pk_end_block(st); return false; }
    } else {
        ip = st.ip;
    }
    if (DETECT && ip + K > xend) { st = pre; return true; }
    const uint32_t b = pkx_bucket(xv, ip);
    uint32_t m;
    bool hit = pkx_lookup(tab, xv, b, ip, m);
    const uint32_t old_m = m;
    pkx_put(tab, xv, b, ip);
    if (hit) hit = (uint32_t)ld64(xv.s, ip) == (uint32_t)ld64(xv.s, m);        // LZ4_read32(match) == LZ4_read32(ip)
    if (hit) {
        uint32_t op = st.op;
        if (searching) {
            while (ip > st.anchor && m > 0 && ld8(xv.s, ip - 1) == ld8(xv.s, m - 1)) { --ip; --m; }
            const uint32_t lit = ip - st.anchor;
            op += 1;
            const uint32_t lhs = op + lit + 8 + lit / 255;
            st.max_lhs = tmax(st.max_lhs, lhs);
            if (lhs > st.budget) { st.op = PK_ABORT; pk_end_block(st); return false; }
            if (lit >= 15) op += (lit - 15) / 255 + 1;
            op += lit;
        } else {
            op += 1;                               // token with zero literals
        }
        op += 2;                                   // offset
        uint32_t lim = st.mlim;
        if (DETECT) lim = tmin(lim, xend);
        const uint32_t mlen = pkx_common(xv, ip, m, lim - ip);
        const uint32_t ip2 = ip + mlen;
        if (DETECT && (ip2 >= xend || (ip2 < st.mfl1 && ip2 + K - 2 > xend))) {
            pkx_put(tab, xv, b, old_m); st = pre; return true;
        }
        const uint32_t mcode = mlen - 4;
        const uint32_t lhs2 = op + 6 + (mcode + 240) / 255;
        st.max_lhs = tmax(st.max_lhs, lhs2);
        if (lhs2 > st.budget) { st.op = PK_ABORT; pk_end_block(st); return false; }
        if (mcode >= 15) op += (mcode - 15) / 255 + 1;
        st.op = op; st.anchor = ip2; st.ip = ip2;
        if (ip2 >= st.mfl1) { pk_end_block(st); return false; }
        pkx_put(tab, xv, pkx_bucket(xv, ip2 - 2), ip2 - 2);
        st.phase = PK_RETEST;                      // immediate re-test at ip2
    } else if (!searching) {
        st.phase = PK_SEARCH; st.fip = ip + 1; st.step = 1; st.nb = 64;
    }
    return false;
}

#if defined(PK_COUNT_STEPS)
static uint64_t pk_general_steps = 0, pk_lean_steps = 0, pk_turbo_steps = 0, pk_batch_steps = 0;
#endif
// ---- inner-loop helpers ------------------------------------------------------------------------
// 16 bases starting at base index k of a packed array viewed as 32-bit words
SNACC_HD uint32_t pk_fsr(uint32_t lo, uint32_t hi, uint32_t sh)
{
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, sh);
#else
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> (sh & 31));
#endif
}
SNACC_HD uint32_t pk_ctz32(uint32_t d)
{
#ifdef __CUDA_ARCH__
    return (uint32_t)__ffs((int)d) - 1;
#else
    return (uint32_t)__builtin_ctz(d);
#endif
}
SNACC_HD uint32_t pk_ring16(const uint32_t *ring32, uint32_t q)
{
    const uint32_t i = q >> 4;
    return pk_fsr(ring32[i & (2 * PK_RING_WORDS - 1)], ring32[(i + 1) & (2 * PK_RING_WORDS - 1)], (q & 15) * 2);
}
SNACC_HD uint32_t pk_glob16(const uint64_t *w, uint32_t k)
{
    const uint32_t *w32 = reinterpret_cast<const uint32_t *>(w);
    const uint32_t i = k >> 4;
    return pk_fsr(SNACC_LDG(w32 + i), SNACC_LDG(w32 + i + 1), (k & 15) * 2);
}

SNACC_HD uint64_t pk_ring_read(const uint64_t *ring, uint32_t q)
{
    const uint32_t i = q >> 5, s = (q & 31) * 2;
    const uint64_t a = ring[i & (PK_RING_WORDS - 1)], b = ring[(i + 1) & (PK_RING_WORDS - 1)];
    return (a >> s) | ((b << 1) << (63 - s));
}
SNACC_HD uint32_t pk_clz32(uint32_t d)
{
#ifdef __CUDA_ARCH__
    return (uint32_t)__clz((int)d);
#else
    return (uint32_t)__builtin_clz(d);
#endif
}

SNACC_HD bool pk_all(uint32_t mask, bool pred)
{
#ifdef __CUDA_ARCH__
    return __all_sync(mask, pred);
#else
    (void)mask; return pred;
#endif
}
SNACC_HD bool pk_any(uint32_t mask, bool pred)
{
#ifdef __CUDA_ARCH__
    return __any_sync(mask, pred);
#else
    (void)mask; return pred;
#endif
}

template <int KIND, int STRIDE>
static __host__ __device__ __noinline__ void pk_step_general(PkState &st, PkTab<KIND, STRIDE> &tab, const PkView &v, uint32_t n)
{
    pk_step<KIND, STRIDE, false>(st, tab, v, n, 0);
}
template <int KIND, int STRIDE>
static __host__ __device__ __noinline__ void pk_step_exact_general(PkState &st, PkTab<KIND, STRIDE> &tab, const PkExact &xv, uint32_t n)
{
    if constexpr (KIND != 1) pk_step_exact<KIND, STRIDE, false>(st, tab, xv, n, 0);
}

// ---- flagged bases and the fast loop ---------------------------------------------------------------------------------
// words of the per-base mask of a sequence (one u32 per 32 bases, like the packed words, plus zero padding)
SNACC_HD uint32_t pk_mask_words(uint32_t len) { return pk_words(len) + 32; }

// one word of the dirty ring from 16 words of the per-base mask (512 bases = 32 granules)
SNACC_HD uint32_t pk_dirty_word(const uint32_t *mask16)
{
    uint32_t d = 0;
    for (int i = 0; i < 16; ++i) {
        const uint32_t w = SNACC_LDG(mask16 + i);
        d |= ((w & 0xffffu) ? 1u : 0u) << (2 * i) | ((w >> 16) ? 1u : 0u) << (2 * i + 1);
    }
    return d;
}

// first flagged granule at or after y offset q inside the ring's coverage [rlo, hi): its start offset (PK_NONE: none)
// and, in *end, the end of the run of flagged granules it starts (at most hi)
SNACC_HD uint32_t pk_next_dirty(const PkView &v, uint32_t q, uint32_t hi, uint32_t *end)
{
    uint32_t g = tmax(q, v.rlo) >> 4;
    const uint32_t gh = hi >> 4;
    while (g < gh) {
        const uint32_t w = v.dring[(g >> 5) & (PK_DRING_WORDS - 1)] >> (g & 31);
        if (w) { g += (uint32_t)SNACC_FFS32(w) - 1; break; }
        g = (g | 31) + 1;
    }
    if (g >= gh) return PK_NONE;
    uint32_t e = g + 1;
    while (e < gh && ((v.dring[(e >> 5) & (PK_DRING_WORDS - 1)] >> (e & 31)) & 1)) ++e;
    *end = e << 4;
    return g << 4;
}

// ---- inner loop ("turbo") ------------------------------------------------------------------------
// One thread is one serial dependency chain: probe position -> table slot -> candidate -> compare ->
// next probe position.  The lanes of a warp run different streams, so the loop is kept WARP-UNIFORM: every
// lane executes every iteration (lanes that have reached `stop` idle), and the burst ends for all of them
// as soon as one lane meets something the loop does not cover (end of block, candidate outside the ring,
// match of 12+ bases, catch-up of 4+ bases, tight output budget, skip step > 1).  pk_run then gives every
// lane one general pk_step -- in lock-step again -- and starts the next burst.  Without this the lanes
// drift apart and the warp executes them one by one (measured: 4.3 of 16 lanes active per instruction).
//
// Exactness: the table sees precisely the library's sequence of reads and writes; an iteration that cannot
// be completed writes nothing and leaves the state untouched.
// Shared-memory accessors of the loop.  On the device they are explicit ld.shared / st.shared on 32-bit
// shared-window addresses (the compiler otherwise re-derives the generic->shared base inside the loop); on the
// host -- the emulation used by the tests -- they are plain pointer accesses.
#ifdef __CUDA_ARCH__
typedef uint32_t pk_sptr;
__device__ __forceinline__ pk_sptr pk_sptr_of(const void *p) { return (pk_sptr)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t pk_lds32(pk_sptr a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t pk_lds16(pk_sptr a) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return v; }
__device__ __forceinline__ void pk_sts32(pk_sptr a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void pk_sts16(pk_sptr a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "h"((uint16_t)v) : "memory"); }
__device__ __forceinline__ uint32_t pk_sel(bool c, uint32_t a, uint32_t b)
{
    uint32_t r;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\tselp.u32 %0, %1, %2, q;\n\t}" : "=r"(r) : "r"(a), "r"(b), "r"((uint32_t)c));
    return r;
}
__device__ __forceinline__ pk_sptr pk_opaque(pk_sptr a) { pk_sptr r; asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(a)); return r; }
#else
typedef uintptr_t pk_sptr;
inline pk_sptr pk_opaque(pk_sptr a) { return a; }
inline pk_sptr pk_sptr_of(const void *p) { return (pk_sptr)p; }
inline uint32_t pk_lds32(pk_sptr a) { return *reinterpret_cast<const uint32_t *>(a); }
inline uint32_t pk_lds16(pk_sptr a) { return *reinterpret_cast<const uint16_t *>(a); }
inline void pk_sts32(pk_sptr a, uint32_t v) { *reinterpret_cast<uint32_t *>(a) = v; }
inline void pk_sts16(pk_sptr a, uint32_t v) { *reinterpret_cast<uint16_t *>(a) = (uint16_t)v; }
inline uint32_t pk_sel(bool c, uint32_t a, uint32_t b) { return c ? a : b; }
#endif

SNACC_HD uint32_t pk_reduce_or(uint32_t mask, uint32_t v)
{
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 800
    return __reduce_or_sync(mask, v);
#elif defined(__CUDA_ARCH__)
    return (__any_sync(mask, v & 1) ? 1u : 0u) | (__any_sync(mask, v & 2) ? 2u : 0u);
#else
    (void)mask; return v;
#endif
}

// ---- the loop ----------------------------------------------------------------------------------------------------
// With one warp per scheduler (26 or 32 streams per warp in the pair tiles, a single stream per warp in the singles
// pass) a warp iteration costs what its instructions and their dependency stalls cost, so the loop does nothing ahead
// of time: per probe it reads the
// candidate's 16 bases, compares, and only then resolves the two slots it needs (the insert at pn-2 and the next probe
// pn).  What keeps it short:
//   * KIND 2 inserts are a 16-bit store plus one shared-memory atomic OR on the epoch bit plane (PkTab): nothing to read
//     back, nothing to undo -- both inserts are issued after the compare, predicated on the iteration being committed;
//   * the ring carries a 16-byte mirror of its first words behind its end, so a 2- or 3-word read needs one masked address;
//   * the conditions that can only change slowly (output budget, skip counter, pending literals) are evaluated at the
//     warp vote, every 4th iteration, with the slack 4 iterations can use up; per iteration only the block limit and
//     the candidate's residence in the ring are tested;
//   * stores are predicated PTX (no branches inside the body).
// (An earlier version looked the table up speculatively for every likely next probe to shorten the dependency chain;
// measured on B200 the lean loop is 1.8x faster even with a single lane per warp -- profiles/README.md.)
#ifndef PK_UNROLL
#define PK_UNROLL 8              // iterations between two warp votes (the body is unrolled that many times)
#endif
#ifndef PK_EPOCH_ATOMIC
#define PK_EPOCH_ATOMIC 0        // 1: epoch bits set with red.shared.or (measured: the shared-memory atomics stall the in-order LSU)
#endif
#ifdef __CUDA_ARCH__
__device__ __forceinline__ void pk_sor32_if(bool c, pk_sptr a, uint32_t v)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q red.shared.or.b32 [%0], %1;\n\t}" :: "r"(a), "r"(v), "r"((uint32_t)c) : "memory");
}
__device__ __forceinline__ void pk_sts16_if(bool c, pk_sptr a, uint32_t v)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.shared.u16 [%0], %1;\n\t}" :: "r"(a), "h"((uint16_t)v), "r"((uint32_t)c) : "memory");
}
__device__ __forceinline__ void pk_sts32_if(bool c, pk_sptr a, uint32_t v)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.shared.u32 [%0], %1;\n\t}" :: "r"(a), "r"(v), "r"((uint32_t)c) : "memory");
}
#else
inline void pk_sor32_if(bool c, pk_sptr a, uint32_t v) { if (c) *reinterpret_cast<uint32_t *>(a) |= v; }
inline void pk_sts16_if(bool c, pk_sptr a, uint32_t v) { if (c) *reinterpret_cast<uint16_t *>(a) = (uint16_t)v; }
inline void pk_sts32_if(bool c, pk_sptr a, uint32_t v) { if (c) *reinterpret_cast<uint32_t *>(a) = v; }
#endif

// Returns whether THIS lane is the reason the burst ended (it needs a general step); the other lanes just re-enter.
template <int KIND, int STRIDE, bool EXC>
SNACC_HD bool pk_turbo_lean(PkState &st, PkTab<KIND, STRIDE> &tab, const PkView &v, uint32_t stop, uint32_t mask, bool work)
{
    typedef PkTab<KIND, STRIDE> Tab;
    constexpr uint32_t MASK = Tab::MASK;
    constexpr uint32_t ESZ = Tab::ESZ;                      // bytes between two slots of one lane
    constexpr uint32_t EWB = STRIDE * 4;                    // KIND 2: bytes between two epoch words of one lane
    constexpr uint32_t RMASK = (2 * PK_RING_WORDS - 1) * 4; // byte-offset mask of the ring seen as u32 words
    constexpr uint32_t VMASK = KIND == 0 ? 0xffffffffu : 0xffffu;   // what a slot keeps of a position
    const uint32_t lx = v.lx, rlo = v.rlo, rspan = v.rspan;
    uint32_t p = st.phase == PK_SEARCH ? st.fip : st.ip;
    // (the skip counter is not carried: in search mode with step 1 it is 63 + the number of pending literals -- 64 at the
    // first probe after a match or a block start, one more per miss -- so "nb < 128" is a bound on the pending literals)
    uint32_t anchor = st.anchor, op = st.op;
    const uint32_t lim = tmin(stop, st.mfl1 > 16 ? st.mfl1 - 16 : 0u);
    // output budget: PK_UNROLL iterations add at most that many x (token + offset + length byte) + the literals pending
    // at the vote + 11 bases each
    constexpr uint32_t UNROLL = EXC ? PK_UNROLL / 2 : PK_UNROLL;   // (the EXC body is longer: keep the loop inside the instruction cache)
    constexpr uint32_t OP_SLACK = 80 + 15 * UNROLL;
    const uint32_t op_lim = st.budget > OP_SLACK ? st.budget - OP_SLACK : 0u;
    const bool fin = !work || p >= stop;
    const bool ok0 = fin || (st.phase <= PK_RETEST && !(st.phase == PK_SEARCH && st.step != 1) &&
                             (uint32_t)(p - 4 - lx - rlo) <= rspan);
    if (!pk_all(mask, ok0) || pk_all(mask, fin)) return !ok0;
    const pk_sptr ring_a = pk_opaque(pk_sptr_of(v.ring)), lut_a = pk_opaque(pk_sptr_of(tab.lut)),
                  tab_a = pk_opaque(pk_sptr_of(tab.t)), ep_a = KIND == 2 ? pk_opaque(pk_sptr_of(tab.ep)) : 0,
                  dr_a = EXC ? pk_opaque(pk_sptr_of(v.dring)) : 0;
#define PK_TLD(addr) (KIND == 0 ? pk_lds32(addr) : pk_lds16(addr))
#define PK_TST_IF(c, addr, val) do { if (KIND == 0) pk_sts32_if(c, addr, val); else pk_sts16_if(c, addr, val); } while (0)
    // state of the pending probe p: slot / epoch-word address, slot index (bit position), candidate, epoch word as last read
    pk_sptr sa = tab_a, ea = ep_a;
    uint32_t eb = 0, m = 0, ew = 0;
    bool near = false;                                      // the slot of p holds a candidate within reach
    if (!fin) {
        const uint32_t c0 = (uint32_t)(pk_ring_read(v.ring, p - lx)) & MASK;
        eb = tab.lut[c0];
        sa = tab_a + eb * ESZ; ea = ep_a + (eb >> 5) * EWB;
        near = tab.lookup(c0, p, m);
        if (KIND == 2) ew = pk_lds32(ea);
    }
    bool blocked = false, stuck = false;
    for (;;) {
        // warp vote every PK_UNROLL iterations: leave when a live lane cannot go on or nobody runs any more.  What can
        // only change slowly is tested here, with the slack that many iterations can use up.
        const uint32_t pend0 = p - anchor;
        const bool live = !fin && p < stop;
        bool run = live && !blocked && op + pend0 <= op_lim && pend0 <= 57 - UNROLL;
        const bool can = run && p < lim;
        stuck = live && !can;
        if (pk_any(mask, stuck) || !pk_any(mask, can)) break;
#pragma unroll
        for (int u = 0; u < (int)UNROLL; ++u) {
            // ---- straight-line, branch-free body: every lane executes everything, stores are predicated, and a lane
            // that does not commit looks its own probe up again (d = 0), which leaves its state as it was
            const bool go = run & (p < lim);
            const uint32_t pend = p - anchor;               // pending literals; search mode iff != 0
            const uint32_t qm4 = m - 4 - lx, qp4 = p - 4 - lx;
            const uint32_t jm = (qm4 >> 2) & RMASK, jp = (qp4 >> 2) & RMASK;
            // candidate: 16 bases from m-4; window: bases [p-4, p+12) and the 16 after them (the mirror behind the
            // ring's end makes +4 / +8 safe without a second mask)
            const uint32_t ca = pk_lds32(ring_a + jm), cb = pk_lds32(ring_a + jm + 4);
            const uint32_t wa = pk_lds32(ring_a + jp), wb = pk_lds32(ring_a + jp + 4), wc = pk_lds32(ring_a + jp + 8);
            const uint32_t xm = pk_fsr(ca, cb, qm4 * 2);
            const uint32_t Ws = pk_fsr(wa, wb, qp4 * 2), Wt = pk_fsr(wb, wc, qp4 * 2);
            uint32_t x = Ws ^ xm;
            const bool inring = (uint32_t)(qm4 - rlo) <= rspan + 32;
            if (EXC) {
                // a flagged base inside the candidate's 16 bases [m-4, m+12) (two granules at most) never equals the
                // clean base it is compared with: force a mismatch there.  Only a candidate that matches 4+ bases in the
                // 2-bit text can change (a shorter match stays a miss), and only if its window is flagged: rare -- the
                // lane that meets one fetches the per-base mask from global memory while the others wait.
                const uint32_t g0 = qm4 >> 4;
                const pk_sptr da = dr_a + ((g0 >> 5) & (PK_DRING_WORDS - 1)) * 4;
                const uint32_t dirty = pk_fsr(pk_lds32(da), pk_lds32(da + 4), g0) & 3;
                if (dirty && near && inring && ((x >> 8) & 0xffu) == 0) {
                    const uint32_t *mw = v.ymask + (qm4 >> 5);
                    uint32_t em = pk_fsr(SNACC_LDG(mw), SNACC_LDG(mw + 1), qm4) & 0xffffu;
                    em = (em | (em << 8)) & 0x00ff00ffu; em = (em | (em << 4)) & 0x0f0f0f0fu;
                    em = (em | (em << 2)) & 0x33333333u; em = (em | (em << 1)) & 0x55555555u;
                    x |= em;
                }
            }
            // equal bases forwards from p (a sentinel bit caps the count at 12) and backwards from p-1 (at 4)
            uint32_t common = pk_ctz32((x >> 8) | 0x01000000u) >> 1;
            common = near ? common : 0;
            uint32_t k = pk_clz32((x << 24) | 0x00800000u) >> 1;
            const uint32_t kmax = tmin(pend, m);
            const bool hit = common >= 4;
            // candidate outside the ring / long match / long catch-up: not for this loop
            // (bitwise on purpose: no short-circuit branches in the body)
            const bool bail = go & ((near & !inring) | (common > 11) | (hit & (k == 4) & (kmax > 4)));
            blocked = blocked | bail;
            run = run & !bail;
            const bool commit = go & !bail, ch = commit & hit;
            k = tmin(k, kmax);
            const uint32_t lit = pend - k;
            const uint32_t add = 3 + lit + (lit >= 15 ? 1u : 0u);   // token + offset + literals (lit <= 57: one length byte at most)
            const uint32_t d = commit ? (hit ? common : 1u) : 0u;   // the next probe is at p + d
            const uint32_t pn = p + d;
            // first insert: slot of p <- p.  KIND 2: the lookup of p read the epoch word last and nothing has been
            // written to the lane's plane since, so the bit is set from the register copy -- only when it is not set
            // yet (one write per slot and epoch in the steady state)
            PK_TST_IF(commit, sa, p);
            if (KIND == 2) {
                const uint32_t b1 = 1u << (eb & 31);
                pk_sts32_if(commit & ((ew & b1) == 0), ea, ew | b1);
            }
            // slots of pn-2 (second insert) and pn (next probe), from the register window
            const uint32_t in = pk_lds16(lut_a + 2 * (pk_fsr(Ws, Wt, 2 * d + 8) & MASK));
            const uint32_t i2 = pk_lds16(lut_a + 2 * (pk_fsr(Ws, Wt, 2 * d + 4) & MASK));
            const pk_sptr san = tab_a + in * ESZ, ean = ep_a + (in >> 5) * EWB;
            const pk_sptr s2 = tab_a + i2 * ESZ, e2 = ep_a + (i2 >> 5) * EWB;
            // all loads first (the candidate of the next probe must not wait for the second insert's read-modify-write) ...
            uint32_t vn = PK_TLD(san), ewn = 0, w2 = 0;
            if (KIND == 2) { ewn = pk_lds32(ean); w2 = pk_lds32(e2); }
            // ... then the second insert (after a match only): slot of pn-2 <- pn-2 ...
            PK_TST_IF(ch, s2, pn - 2);
            const uint32_t b2 = 1u << (i2 & 31);
            if (KIND == 2) pk_sts32_if(ch & ((w2 & b2) == 0), e2, w2 | b2);
            // ... and what it changes in the values just loaded
            vn = (ch & (i2 == in)) ? ((pn - 2) & VMASK) : vn;
            if (KIND == 2) ewn = (ch & ((i2 >> 5) == (in >> 5))) ? (ewn | b2) : ewn;
            // candidate of the next probe
            uint32_t mn; bool nn;
            if (KIND == 2) {
                mn = pn - ((pn - vn) & 0xffffu);
                nn = ((ewn >> (in & 31)) & 1) ? mn != pn : vn > (pn & 0xffffu);
            } else {
                mn = vn;
                nn = KIND == 1 || (pn - vn <= LZ4_MAX_DISTANCE);
            }
#if defined(PK_COUNT_STEPS) && !defined(__CUDA_ARCH__)
            if (commit) ++pk_turbo_steps;
#endif
            op += ch ? add : 0u;
            anchor = ch ? pn : anchor;
            m = mn; near = nn; sa = san; ea = ean; eb = in; ew = ewn; p = pn;
        }
    }
#undef PK_TLD
#undef PK_TST_IF
    if (work && st.phase <= PK_RETEST) {
        if (p != anchor) { st.phase = PK_SEARCH; st.fip = p; st.step = 1; st.nb = 63 + (p - anchor); }
        else             { st.phase = PK_RETEST; st.ip = p; }
        st.anchor = anchor; st.op = op;
    }
    return stuck;
}

// ---- singles pass: 64 probes at a time ---------------------------------------------------------------------------
// A sequence on its own is ONE dependency chain, and the singles pass sits in front of every pair tile (0.16 s of a
// c4 step with the scalar loop above, the same on 1 or 8 GPUs).  In the singles kernel the whole warp works on it:
// every lane does, for the probes at P + lane and P + 32 + lane, everything the loop above does that does not depend
// on the probes before -- the window, the slot, the candidate as the table has it at the start of the batch, the
// 12-base compare forwards and the 4-base compare backwards -- and packs the outcome into one word.  Then all lanes
// walk the chain together: the probe at p takes the word of position p - P (one shuffle), and the next probe is d
// further; that shuffle and add are all that is left on the dependency chain, the bookkeeping of op / anchor and the
// two inserts hang off it.  A lane's candidate is stale once a probe of the same batch has written its slot (slot of
// p, slot of pn - 2): the walk reads the slot again (its index is in the word) and compares it with what the lane saw
// (a second shuffle); it stops in front of a stale position -- and at anything else the scalar loop does not cover
// either -- and pk_run lets the scalar loop do a few probes before the next batch.
// Same state, same table reads and writes as the scalar loop.
constexpr uint32_t PKB_D = 0xfu, PKB_HIT = 0x10u, PKB_K_SHIFT = 5, PKB_BAIL = 0x100u, PKB_IDX_SHIFT = 9, PKB_IDX2_SHIFT = 19,
                   PKB_K4 = 1u << 29;
constexpr uint32_t PKB_WINDOW = 64, PKB_HOP_MAX = 52;       // a hop from position l touches positions up to l + 11

// the probe at q = P + l on its own
template <int KIND>
SNACC_HD uint32_t pk_batch_probe(const PkTab<KIND, 1> &tab, const PkView &v, uint32_t q, uint32_t l, uint32_t &idx, uint32_t &seen)
{
    constexpr uint32_t MASK = PkTab<KIND, 1>::MASK, RMASK = (2 * PK_RING_WORDS - 1) * 4;
    const pk_sptr ring_a = pk_sptr_of(v.ring);
    const uint32_t qp4 = q - 4 - v.lx, jp = (qp4 >> 2) & RMASK;
    const uint32_t wa = pk_lds32(ring_a + jp), wb = pk_lds32(ring_a + jp + 4), wc = pk_lds32(ring_a + jp + 8);
    const uint32_t Ws = pk_fsr(wa, wb, qp4 * 2), Wt = pk_fsr(wb, wc, qp4 * 2);
    idx = tab.lut[pk_fsr(Ws, Wt, 8) & MASK];
    uint32_t m;
    const bool near = tab.lookup_idx(idx, q, m);
    seen = tab.t[idx];                                                 // (KIND 0 / 1: the slot is the candidate position)
    const uint32_t qm4 = m - 4 - v.lx, jm = (qm4 >> 2) & RMASK;
    const uint32_t ca = pk_lds32(ring_a + jm), cb = pk_lds32(ring_a + jm + 4);
    const uint32_t x = Ws ^ pk_fsr(ca, cb, qm4 * 2);
    const bool inring = (uint32_t)(qm4 - v.rlo) <= v.rspan + 32;
    uint32_t common = pk_ctz32((x >> 8) | 0x01000000u) >> 1;           // forwards from q, at most 12
    common = near ? common : 0;
    const uint32_t k = pk_clz32((x << 24) | 0x00800000u) >> 1;         // backwards from q - 1, at most 4
    const bool hit = common >= 4;
    // not for the walk: candidate outside the ring, match of 12+ bases, a catch-up that may end at the stream start,
    // and the positions a hop must not start from (its second insert would lie beyond the window)
    const bool bail = (near && !inring) || common > 11 || (hit && m < 64) || l > PKB_HOP_MAX;
    return (hit ? common & PKB_D : 1u) | (hit ? PKB_HIT : 0u) | (k << PKB_K_SHIFT) | (bail ? PKB_BAIL : 0u) | (idx << PKB_IDX_SHIFT) |
           (hit && k == 4 ? PKB_K4 : 0u);
}

template <int KIND>
SNACC_HD void pk_batch_run(PkState &st, PkTab<KIND, 1> &tab, const PkView &v, uint32_t stop)
{
    typedef typename PkTab<KIND, 1>::T T;
    if (st.phase > PK_RETEST || (st.phase == PK_SEARCH && st.step != 1)) return;
    uint32_t p = st.phase == PK_SEARCH ? st.fip : st.ip;
    uint32_t anchor = st.anchor, op = st.op;
    const uint32_t p_in = p;
    const uint32_t lim = tmin(stop, st.mfl1 > 16 ? st.mfl1 - 16 : 0u);
    constexpr uint32_t OP_SLACK = 80 + 15 * PKB_WINDOW;
    const uint32_t op_lim = st.budget > OP_SLACK ? st.budget - OP_SLACK : 0u;
#ifdef __CUDA_ARCH__
    const uint32_t lane = threadIdx.x & 31;
    const pk_sptr tab_a = pk_sptr_of(tab.t);
#endif
    for (;;) {
        const uint32_t P = p;
        if (P + PKB_WINDOW > lim || op + (P - anchor) > op_lim || (uint32_t)(P - 4 - v.lx - v.rlo) > v.rspan) break;
        // ---- every lane: its two probes; the slot of q + d - 2 (second insert of a hit at q) comes from the lane that owns it
#ifdef __CUDA_ARCH__
        uint32_t idxA, idxB, seenA, seenB;
        uint32_t packA = pk_batch_probe<KIND>(tab, v, P + lane, lane, idxA, seenA);
        uint32_t packB = pk_batch_probe<KIND>(tab, v, P + 32 + lane, 32 + lane, idxB, seenB);
        {
            const uint32_t la = lane + (packA & PKB_D) - 2, lb = lane + (packB & PKB_D) - 2;       // (a miss: d = 1, unused)
            const uint32_t a_lo = __shfl_sync(0xffffffffu, idxA, la & 31), a_hi = __shfl_sync(0xffffffffu, idxB, la & 31);
            const uint32_t b_hi = __shfl_sync(0xffffffffu, idxB, lb & 31);
            packA |= ((la & 32) ? a_hi : a_lo) << PKB_IDX2_SHIFT;
            packB |= b_hi << PKB_IDX2_SHIFT;                          // (lb >= 32: only for positions no hop starts from)
        }
#define PKB_PACK(l) __shfl_sync(0xffffffffu, ((l) & 32) ? packB : packA, (l) & 31)
#define PKB_SEEN(l) __shfl_sync(0xffffffffu, ((l) & 32) ? seenB : seenA, (l) & 31)
#else
        uint32_t pack[PKB_WINDOW], idx[PKB_WINDOW], seen[PKB_WINDOW];
        for (uint32_t l = 0; l < PKB_WINDOW; ++l) pack[l] = pk_batch_probe<KIND>(tab, v, P + l, l, idx[l], seen[l]);
        for (uint32_t l = 0; l < PKB_WINDOW; ++l) pack[l] |= idx[(l + (pack[l] & PKB_D) - 2) & (PKB_WINDOW - 1)] << PKB_IDX2_SHIFT;
#define PKB_PACK(l) pack[(l) & (PKB_WINDOW - 1)]
#define PKB_SEEN(l) seen[(l) & (PKB_WINDOW - 1)]
#endif
        // ---- the walk, uniform over the warp.  l runs ahead on its own (d comes straight out of the shuffled word);
        // everything that changes state is predicated on `ok`, which drops for good at the first position the walk
        // cannot take, and p / anchor / op stay where they were then.  (While ok holds l <= 63: hops start at l <= 52.)
        uint32_t l = 0;
        bool ok = true;
        do {
            ok = ok & (p - anchor <= 52);                   // at most 4 more literals before the next test: step stays 1
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t pk = PKB_PACK(l);
                const uint32_t sn = PKB_SEEN(l);
                const uint32_t d = pk & PKB_D, kk = (pk >> PKB_K_SHIFT) & 7, pend = p - anchor;
                const bool hit = (pk & PKB_HIT) != 0;
                const uint32_t i1 = (pk >> PKB_IDX_SHIFT) & 1023, i2 = (pk >> PKB_IDX2_SHIFT) & 1023;
                const uint32_t pn = p + d;
                // stale: the slot no longer holds what the lane saw (an insert of this batch; positions only grow)
#ifdef __CUDA_ARCH__
                const uint32_t now = KIND == 0 ? pk_lds32(tab_a + i1 * 4) : pk_lds16(tab_a + i1 * 2);
#else
                const uint32_t now = tab.t[i1];
#endif
                ok = ok & ((pk & PKB_BAIL) == 0) & !(((pk & PKB_K4) != 0) & (pend > 4)) & (now == sn);
                const bool gh = ok & hit;
#ifdef __CUDA_ARCH__
                // the inserts: every lane stores the same value to the same address
                if (KIND == 0) { pk_sts32_if(ok, tab_a + i1 * 4, p); pk_sts32_if(gh, tab_a + i2 * 4, pn - 2); }
                else           { pk_sts16_if(ok, tab_a + i1 * 2, p); pk_sts16_if(gh, tab_a + i2 * 2, pn - 2); }
#else
                if (ok) {
                    tab.t[i1] = (T)p;
                    if (hit) tab.t[i2] = (T)(pn - 2);
#if defined(PK_COUNT_STEPS)
                    ++pk_batch_steps;
#endif
                }
#endif
                const uint32_t lit = pend - tmin(kk, pend);
                op += gh ? 3 + lit + (lit >= 15 ? 1u : 0u) : 0u;
                anchor = gh ? pn : anchor;
                p = ok ? pn : p;
                l += d;
            }
        } while (ok);
#undef PKB_PACK
#undef PKB_SEEN
        if (p == P) break;
    }
    if (p != p_in) {
        if (p != anchor) { st.phase = PK_SEARCH; st.fip = p; st.step = 1; st.nb = 63 + (p - anchor); }
        else             { st.phase = PK_RETEST; st.ip = p; }
        st.anchor = anchor; st.op = op;
    }
}

#ifdef __CUDA_ARCH__
// KIND 2 epoch boundary (PkTab::new_epoch) done by the whole warp for every lane that stands at a block start: 32
// slots per lane and step instead of one lane walking its ~900 slots while the others wait
template <int STRIDE>
__device__ __forceinline__ void pk_epoch_coop(const PkState &st, PkTab<2, STRIDE> &tab, uint32_t n, uint32_t mask, bool work)
{
    const bool need = work && st.phase == PK_BLOCK_START && st.bs < n && tab.epoch_base != st.bs;
    uint32_t todo = __ballot_sync(mask, need);
    if (!todo) return;
    const uint32_t lane = threadIdx.x & 31;
    const int mine = (int)tmin(lane, (uint32_t)STRIDE - 1);   // lanes beyond the tile's streams point at the last stream's table
    const uint32_t rank = __popc(mask & ((1u << lane) - 1)), nact = __popc(mask);
    const uint32_t nw = (tab.nslot + 31) >> 5;
    while (todo) {
        const int L = __ffs(todo) - 1;
        todo &= todo - 1;
        uint16_t *tL = tab.t + (L - mine);
        uint32_t *eL = tab.ep + (L - mine);
        for (uint32_t w = rank; w < nw; w += nact) {
            uint32_t z = ~eL[w * STRIDE];
            if (tab.nslot - w * 32 < 32) z &= (1u << (tab.nslot - w * 32)) - 1;
            for (; z; z &= z - 1) tL[(w * 32 + __ffs(z) - 1) * STRIDE] = 0;
            eL[w * STRIDE] = 0;
        }
    }
    __syncwarp(mask);
    if (need) tab.epoch_base = st.bs;
}
#endif

// Run the streams of the warp (lanes in `mask`, one stream each) until each one's next position reaches
// its `stop` or it is done: turbo bursts, separated by one general pk_step for the lanes that ended the burst.
// BATCH (singles pass, KIND 0 / 1, one stream for the whole warp -- on the device every lane of the warp calls this
// with the same state): pk_batch_run in front, the scalar loop for the probes a batch stops at.
template <int KIND, int STRIDE, bool EXC = false, bool BATCH = false>
SNACC_HD void pk_run(PkState &st, PkTab<KIND, STRIDE> &tab, const PkView &v, const PkExact *xv, uint32_t n, uint32_t stop, uint32_t mask)
{
    for (;;) {
        bool work = st.phase != PK_DONE && pk_next_pos(st) < stop;
        if (!pk_any(mask, work)) return;
        if constexpr (BATCH && STRIDE == 1 && KIND != 2 && !EXC) {
            pk_batch_run<KIND>(st, tab, v, stop);
            work = st.phase != PK_DONE && pk_next_pos(st) < stop;
            if (!work) return;
            const bool stuck = pk_turbo_lean<KIND, STRIDE, EXC>(st, tab, v, tmin(stop, pk_next_pos(st) + 24), mask, work);
            if (stuck && st.phase != PK_DONE && pk_next_pos(st) < stop) {
#if defined(PK_COUNT_STEPS) && !defined(__CUDA_ARCH__)
                ++pk_general_steps;
#endif
                pk_step_general<KIND, STRIDE>(st, tab, v, n);
            }
            continue;
        }
        // only the lanes that ended the burst take a general step: a round then costs one lane's path, not the divergent
        // paths of all 26 (measured: the general steps were 13 % of the pair kernel's time, mostly block ends, which the
        // lanes reach at unrelated times)
        const bool stuck = pk_turbo_lean<KIND, STRIDE, EXC>(st, tab, v, stop, mask, work);
        work = stuck && st.phase != PK_DONE && pk_next_pos(st) < stop;
#ifdef __CUDA_ARCH__
        if constexpr (KIND == 2 && STRIDE != 1) pk_epoch_coop<STRIDE>(st, tab, n, mask, work);
#endif
        if (work) {
#if defined(PK_COUNT_STEPS) && !defined(__CUDA_ARCH__)
            ++pk_general_steps;
#endif
            if constexpr (EXC) pk_step_exact_general<KIND, STRIDE>(st, tab, *xv, n);
            else pk_step_general<KIND, STRIDE>(st, tab, v, n);
        }
    }
}

// The same for sequences that may hold flagged bases (EXC): every lane runs the fast loop up to 20 bases ahead of the
// next flagged granule of y (its window [p-4, p+16) then holds clean bases only), waits there for the other lanes --
// they share y, so they arrive within a few iterations of each other -- and all of them cross the flagged stretch with
// byte-exact steps in lock-step.  `hi`: end of the ring's coverage (y offset).
template <int KIND, int STRIDE>
SNACC_HD void pk_run_exc(PkState &st, PkTab<KIND, STRIDE> &tab, const PkView &v, const PkExact &xv, uint32_t n, uint32_t stop,
                         uint32_t hi, uint32_t mask)
{
    for (;;) {
        uint32_t np = pk_next_pos(st);
        const bool active = st.phase != PK_DONE && np < stop;
        uint32_t lane_stop = stop, until = 0;
        if (active) {
            uint32_t e_end = 0;
            const uint32_t q = np > v.lx + 4 ? np - v.lx - 4 : 0;
            const uint32_t e = pk_next_dirty(v, q, hi, &e_end);
            if (e != PK_NONE) {
                lane_stop = tmin(stop, v.lx + e > 20 ? v.lx + e - 20 : 0u);
                until = v.lx + e_end + 4;
            }
        }
        pk_run<KIND, STRIDE, true>(st, tab, v, &xv, n, lane_stop, mask);
        const bool cross = active && lane_stop < stop;       // stopped by a flagged granule, not by the ring
        if (!pk_any(mask, cross && st.phase != PK_DONE && pk_next_pos(st) < stop)) return;
        for (;;) {
            const bool c = cross && st.phase != PK_DONE && pk_next_pos(st) < tmin(until, stop);
            if (!pk_any(mask, c)) break;
#ifdef __CUDA_ARCH__
            if constexpr (KIND == 2 && STRIDE != 1) pk_epoch_coop<STRIDE>(st, tab, n, mask, c);
#endif
            if (c) {
#if defined(PK_COUNT_STEPS) && !defined(__CUDA_ARCH__)
                ++pk_general_steps;
#endif
                pk_step_exact_general<KIND, STRIDE>(st, tab, xv, n);
            }
        }
    }
}

// Re-derive the open block's limits for the real stream length when a pair stream resumes from the
// prefix checkpoint of x (taken with be = bs + 64 KiB).  Returns false when a budget test that passed in
// the prefix pass could fail under the real, smaller budget (the job then takes the byte-wise kernel).
SNACC_HD bool pk_resume(PkState &st, uint32_t n)
{
    if (st.phase == PK_BLOCK_START || st.phase == PK_DONE) return true;
    const uint32_t be = (n - st.bs > LZ4_BLOCK) ? st.bs + LZ4_BLOCK : n;
    st.be = be; st.budget = be - st.bs - 1;
    st.mfl1 = be - LZ4_MFLIMIT + 1; st.mlim = be - LZ4_LASTLITERALS;
    return st.max_lhs <= st.budget;
}

// bookkeeping of the ring window, identical on every thread of the CTA; the caller fills the word
// range [w0, w1) that start()/advance() return
struct PkRing {
    uint32_t yw_total;     // words of y that may be loaded (incl. zero padding)
    uint32_t hi_w;         // words [.., hi_w) loaded so far
    uint32_t cover;        // words behind hi_w that count as resident: PK_RING_WORDS, or PK_RING_COVER_EXC when the words
                           // the next refill replaces host the dirty ring (EXC kernels)
    SNACC_HD void start(uint32_t ly, uint32_t &w0, uint32_t &w1, uint32_t cover_ = PK_RING_WORDS)
    {
        cover = cover_;
        yw_total = pk_words(ly);
        // EXC: the first fill leaves the ring's tail free for the dirty ring
        hi_w = tmin(yw_total, first_fill());
        w0 = 0; w1 = hi_w;
    }
    SNACC_HD bool complete() const { return hi_w >= yw_total; }
    // a stream over y is `runs()` calls of pk_run, one per ring state: the first fill, then one per advance()
    SNACC_HD uint32_t runs() const
    {
        const uint32_t ff = first_fill();
        return yw_total <= ff ? 1u : 1u + (yw_total - ff + PK_CHUNK_BASES / 32 - 1) / (PK_CHUNK_BASES / 32);
    }
    // the ring as it is before run r >= 1 (after r advances), for a CTA that takes a stream over from another one
    // (tile segments): [w0, w1) = every word resident at that point
    SNACC_HD void restart(uint32_t ly, uint32_t r, uint32_t &w0, uint32_t &w1, uint32_t cover_ = PK_RING_WORDS)
    {
        cover = cover_;
        yw_total = pk_words(ly);
        hi_w = tmin(yw_total, first_fill() + r * (PK_CHUNK_BASES / 32));
        w1 = hi_w;
        w0 = hi_w > PK_RING_WORDS ? hi_w - PK_RING_WORDS : 0u;
    }
    SNACC_HD void advance(uint32_t &w0, uint32_t &w1)
    {
        w0 = hi_w;
        hi_w = tmin(yw_total, hi_w + PK_CHUNK_BASES / 32);
        w1 = hi_w;
    }
    SNACC_HD uint32_t first_fill() const { return cover == PK_RING_WORDS ? PK_RING_WORDS : PK_RING_WORDS - 160; }
    // until the ring wraps every word loaded so far is resident
    SNACC_HD uint32_t lo_w() const { return hi_w <= first_fill() ? 0 : hi_w - cover; }
    SNACC_HD void view(PkView &v) const
    {
        v.rlo = lo_w() * 32;
        v.rspan = (hi_w - lo_w()) * 32 - 64;     // hi_w - lo_w >= 6 words always (padding)
    }
    // streams may run while their next position is below this y offset
    SNACC_HD uint32_t stop_q() const { return complete() ? 0xffffffffu : hi_w * 32 - PK_GUARD; }
    // EXC: where the dirty ring (PK_DRING_WORDS + 1 u32 = 129 ring words) lives inside the ring: in the words that are
    // no longer resident -- the ones the next refill replaces, or the never-filled tail of a short y
    SNACC_HD uint32_t dring_word() const
    {
        const uint32_t b = hi_w & (PK_RING_WORDS - 1);
        return hi_w < PK_RING_WORDS ? hi_w : (b + 129 <= PK_RING_WORDS ? b : 0);
    }
};
// EXC: 992 ring words (31 Ki bases) are given up: the window (64 Ki) + one refill chunk (32 Ki) + the guard still fit
constexpr uint32_t PK_RING_COVER_EXC = PK_RING_WORDS - 992;

// entry k of the dirty ring for the ring's current coverage (the one dirty word j with j mod PK_DRING_WORDS == k that
// overlaps the resident packed words; 0 if none), from y's per-base mask
SNACC_HD uint32_t pk_dring_entry(const PkRing &rg, const uint32_t *ymask, uint32_t k)
{
    const uint32_t j0 = rg.lo_w() >> 4, j1 = (rg.hi_w + 15) >> 4;
    uint32_t j = (j0 & ~(PK_DRING_WORDS - 1)) + k;
    if (j < j0) j += PK_DRING_WORDS;
    return j < j1 ? pk_dirty_word(ymask + 16 * j) : 0u;
}

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// device side
// ------------------------------------------------------------------------------------------------
struct PkCorpus {
    const uint64_t *words;       // all packed sequences
    const uint64_t *woff;        // word offset of sequence i
    const uint32_t *len;         // bases
};

struct PkTile {                  // up to T pair jobs sharing y
    int32_t y, count;
    int64_t first;               // index of the tile's first job in tile_x / tile_out
};

// what the EXC kernels need on top of PkCorpus (corpora with bytes outside the alphabet): the per-base masks, the true
// bytes, the bucket -> slot map and the overflow tables of pk_step_exact
struct PkExcCorpus {
    const uint32_t *mask;        // per-base masks of all sequences
    const uint64_t *moff;        // word offset of sequence i in mask
    const uint8_t *bytes;        // the padded ASCII corpus
    const uint64_t *boff;        // byte offset of sequence i
    const uint16_t *b2s;         // LZ4 bucket -> slot index (0xffff: overflow table)
    uint32_t *ovf_work;          // one overflow table (PK_OVF_ENTRIES words) per resident stream
    uint32_t *ck_ovf;            // overflow table of every sequence's prefix checkpoint
};

// checkpoint storage: slot = seq * 2 + (linked ? 1 : 0); table stored as u32[1024] in both regimes
constexpr uint32_t PK_CKPT_TAB = 1024;

// Tile segments (linked regime): a tile's pass over y is cut at ring-refill boundaries into n_seg segments that are
// separate work items, so that the last round of a launch is 1/n_seg of a tile long instead of a whole tile (a launch
// of 320 tiles on 148 SMs takes 2.25 tile times with 8 segments instead of 3).  At a cut every stream's table, epoch
// plane and state go to HBM (2 KB per stream) and flag[tile] is raised; the CTA that draws the next segment waits for
// the flag -- work items are drawn segment-major from one counter and every CTA of the grid is resident, so the
// producer of a segment is always running or finished.
struct PkSegStore {
    uint16_t *tab;               // [tile * T + stream][nslot]
    uint32_t *ep;                // [tile * T + stream][nw]
    PkState *st;                 // [tile * T + stream]
    uint32_t *eb;                // [tile * T + stream]  epoch_base | (1 if the job left the packed path)
    int32_t *flag;               // [tile]  segments finished
};

// cooperative ring fill: y words [w0, w1) -> ring; w0, w1 even
__device__ __forceinline__ void pk_ring_fill(uint64_t *ring, const uint64_t *yw, uint32_t w0, uint32_t w1)
{
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(yw);
    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(ring);
    for (uint32_t i = (w0 >> 1) + threadIdx.x; i < (w1 >> 1); i += blockDim.x) {
        const ulonglong2 q = __ldg(src + i);
        dst[i & (PK_RING_WORDS / 2 - 1)] = q;
        if ((i & (PK_RING_WORDS / 2 - 1)) == 0) dst[PK_RING_WORDS / 2] = q;   // mirror of the first 16 bytes behind the end
    }
}

// (re)build the dirty ring for the ring's current coverage, inside the ring words that are not resident (PkRing::dring_word)
__device__ __forceinline__ void pk_dring_build(uint64_t *ring, const PkRing &rg, const uint32_t *ymask)
{
    uint32_t *dr = reinterpret_cast<uint32_t *>(ring + rg.dring_word());
    for (uint32_t k = threadIdx.x; k < PK_DRING_WORDS; k += blockDim.x) {
        const uint32_t d = pk_dring_entry(rg, ymask, k);
        dr[k] = d;
        if (k == 0) dr[PK_DRING_WORDS] = d;
    }
}

// ---- pair tiles: one stream per thread, LANES active lanes per warp, tables in shared memory --------
// KIND: table storage (PkTab).  nslot: slots per stream (KIND 2: what pk_slot_lut needs, rounded up to 2).
// tile_out == nullptr: rectangle mode -- the tile's jobs are rows first .. first+count-1 of tile_x against
// column y, and results go to out[(first + k) * out_stride + (y - col0)] (no per-job arrays on the host).
// shared memory: ring | code->slot map | position tables (warp-major, lane-interleaved) | epoch planes
// (EXC: the dirty ring lives inside the ring, in the words that are not resident -- PkRing::dring_word)
// EXC: the corpus holds sequences with bytes outside the alphabet (pk_run_exc, pk_step_exact); xc is unused otherwise.
template <int KIND, int LANES, bool EXC>
__global__ void __launch_bounds__(384, 1)
lz4_pk_pair_kernel(PkCorpus pc, PkExcCorpus xc, const PkTile *__restrict__ tiles, int32_t n_tiles, const int32_t *__restrict__ tile_x,
                   const int64_t *__restrict__ tile_out, const uint32_t *__restrict__ ck_tab,
                   const PkState *__restrict__ ck_state, const uint16_t *__restrict__ lut_g, uint32_t nslot,
                   int64_t out_stride, int32_t col0, unsigned long long *__restrict__ counter, int64_t *__restrict__ out,
                   int32_t n_seg, PkSegStore ss)
{
    typedef PkTab<KIND, LANES> Tab;
    constexpr bool U16 = Tab::U16;
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t *ring = reinterpret_cast<uint64_t *>(smem);
    uint16_t *lut = reinterpret_cast<uint16_t *>(smem + PK_RING_BYTES);
    typename Tab::T *tabs = reinterpret_cast<typename Tab::T *>(smem + PK_RING_BYTES + Tab::ENTRIES * 2);
    const uint32_t n_warps = blockDim.x >> 5;
    if (KIND != 2) nslot = Tab::ENTRIES;
    const uint32_t nw = (nslot + 31) >> 5;                 // epoch words per stream
    uint32_t *eps = reinterpret_cast<uint32_t *>(tabs + (size_t)n_warps * nslot * LANES);
    __shared__ int32_t s_tile;

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t slot = warp * LANES + lane;
    for (uint32_t i = threadIdx.x; i < Tab::ENTRIES; i += blockDim.x) lut[i] = lut_g[i];
    // all 32 lanes of a warp run the stream loop (full-mask votes, no partial-warp synchronisation); the lanes beyond
    // the tile's LANES streams never have work and point at the last stream's table (they only ever load from it)
    const uint32_t tl = lane < LANES ? lane : LANES - 1;
    Tab tab;
    tab.t = tabs + (size_t)warp * (nslot * LANES) + tl;
    tab.ep = eps + (size_t)warp * (nw * LANES) + tl;
    tab.epoch_base = 0;
    tab.nslot = nslot;
    tab.lut = lut;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = (int32_t)atomicAdd(counter, 1ull);
        __syncthreads();
        const int32_t item = s_tile;                       // work items: segment-major
        if (item >= n_tiles * n_seg) break;
        const int32_t tile = item % n_tiles, seg = item / n_tiles;
        const PkTile td = tiles[tile];
        const uint32_t ly = pc.len[td.y];
        const uint64_t *yw = pc.words + pc.woff[td.y];
        PkRing rg;
        uint32_t w0, w1;
        rg.start(ly, w0, w1, EXC ? PK_RING_COVER_EXC : PK_RING_WORDS);
        // this item's runs [r, r1) of the tile's pass over y (host: n_seg <= runs of every tile)
        const uint32_t runs = rg.runs();
        uint32_t r = (uint32_t)((uint64_t)runs * seg / n_seg);
        const uint32_t r1 = (uint32_t)((uint64_t)runs * (seg + 1) / n_seg);
        const size_t seg_base = (size_t)tile * (n_warps * LANES);         // the tile's records in ss
        // EXC: a stream's overflow table stays where it is between segments (one per stream of the launch, not per
        // resident stream)
        const size_t ovf_base = n_seg > 1 ? seg_base : (size_t)blockIdx.x * (n_warps * LANES);
        if (seg > 0) {
            if (threadIdx.x == 0) {
                int32_t f;
                do {
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(ss.flag + tile) : "memory");
                    if (f < seg) __nanosleep(256);
                } while (f < seg);
            }
            __syncthreads();
            rg.restart(ly, r, w0, w1, EXC ? PK_RING_COVER_EXC : PK_RING_WORDS);
        }
        pk_ring_fill(ring, yw, w0, w1);
        const uint32_t *ymask = EXC ? xc.mask + xc.moff[td.y] : nullptr;
        if (EXC) pk_dring_build(ring, rg, ymask);

        const bool has = lane < LANES && (int32_t)slot < td.count;
        PkState st = PkState();
        PkView v;
        v.ring = ring; v.yw = yw; v.xw = yw; v.lx = 0; v.dring = nullptr; v.ymask = ymask;
        PkExact xv;
        xv.b2s = xc.b2s; xv.ovf = nullptr; xv.s.x = xv.s.y = nullptr; xv.s.lx = xv.s.n = 0;
        uint32_t n = 0;
        bool bail = false;
        st.phase = PK_DONE; st.total = 0;
        // each warp copies the checkpoint tables of its streams, one stream at a time (coalesced reads);
        // KIND 2 converts the 32-bit checkpoint positions relative to the stream's open block
        for (int k = 0; k < LANES; ++k) {
            const int32_t sl = (int32_t)(warp * LANES + k);
            if (sl >= td.count) break;
            Tab tk = tab;
            tk.t = tabs + (size_t)warp * (nslot * LANES) + k;
            tk.ep = eps + (size_t)warp * (nw * LANES) + k;
            if (KIND == 2 && seg > 0) {                        // the records the previous segment left
                const uint16_t *s16 = ss.tab + (seg_base + sl) * nslot;
                const uint32_t *s32 = ss.ep + (seg_base + sl) * nw;
                for (uint32_t e = lane; e < nslot; e += 32) tk.t[e * LANES] = s16[e];
                for (uint32_t wd = lane; wd < nw; wd += 32) tk.ep[wd * LANES] = s32[wd];
                continue;
            }
            const int32_t x = tile_x[td.first + sl];
            const uint32_t *src = ck_tab + (size_t)(2 * x + (U16 ? 0 : 1)) * PK_CKPT_TAB;
            const uint32_t bs = ck_state[2 * x + (U16 ? 0 : 1)].bs;
            if (KIND == 2) {
                // one lane per epoch word so that the read-modify-writes of import_slot never collide
                for (uint32_t wd = lane; wd < nw; wd += 32)
                    for (uint32_t e = wd * 32; e < tmin(wd * 32 + 32, nslot); ++e) tk.import_slot(e, src[e], bs);
            } else {
                for (uint32_t e = lane; e < nslot; e += 32) tk.import_slot(e, src[e], bs);
            }
            if (EXC) {                                         // the stream's overflow table starts as its x's
                const uint4 *so = reinterpret_cast<const uint4 *>(xc.ck_ovf + (size_t)x * PK_OVF_ENTRIES);
                uint4 *dd = reinterpret_cast<uint4 *>(xc.ovf_work + (ovf_base + sl) * PK_OVF_ENTRIES);
                for (uint32_t e = lane; e < PK_OVF_ENTRIES / 4; e += 32) dd[e] = so[e];
            }
        }
        if (has) {
            const int32_t x = tile_x[td.first + slot];
            st = ck_state[2 * x + (U16 ? 0 : 1)];
            v.lx = pc.len[x];
            v.xw = pc.words + pc.woff[x];
            n = v.lx + ly;
            if (EXC) {
                xv.s.x = xc.bytes + xc.boff[x]; xv.s.y = xc.bytes + xc.boff[td.y]; xv.s.lx = v.lx; xv.s.n = n;
                xv.ovf = xc.ovf_work + (ovf_base + slot) * PK_OVF_ENTRIES;
            }
            if (KIND == 2 && seg > 0) {
                st = ss.st[seg_base + slot];
                const uint32_t eb = ss.eb[seg_base + slot];
                bail = eb & 1u;
                tab.epoch_base = eb & ~1u;
            } else {
                bail = !pk_resume(st, n);
                if (bail) st.phase = PK_DONE;
                tab.epoch_base = st.bs;            // the imported bits refer to the checkpoint's open block
            }
        }
        for (;;) {
            __syncthreads();                       // ring (and on the first pass the tables) visible
            rg.view(v);
            if (EXC) v.dring = reinterpret_cast<const uint32_t *>(ring + rg.dring_word());
            const uint32_t stop_q = rg.stop_q();
            const uint32_t stop = stop_q == 0xffffffffu ? 0xffffffffu : v.lx + stop_q;
            if constexpr (EXC) pk_run_exc<KIND, LANES>(st, tab, v, xv, n, stop, rg.hi_w * 32, 0xffffffffu);
            else pk_run<KIND, LANES>(st, tab, v, nullptr, n, stop, 0xffffffffu);
            if (++r >= r1) break;                  // r1 == runs: the ring is complete
            __syncthreads();                       // everyone is done reading the slots about to be replaced
            rg.advance(w0, w1);
            pk_ring_fill(ring, yw, w0, w1);
            if (EXC) pk_dring_build(ring, rg, ymask);
        }
        if (KIND == 2 && seg + 1 < n_seg) {
            // hand the tile's streams to whoever draws the next segment
            __syncwarp();
            for (int k = 0; k < LANES; ++k) {
                const int32_t sl = (int32_t)(warp * LANES + k);
                if (sl >= td.count) break;
                const typename Tab::T *tk = tabs + (size_t)warp * (nslot * LANES) + k;
                const uint32_t *ek = eps + (size_t)warp * (nw * LANES) + k;
                uint16_t *d16 = ss.tab + (seg_base + sl) * nslot;
                uint32_t *d32 = ss.ep + (seg_base + sl) * nw;
                for (uint32_t e = lane; e < nslot; e += 32) d16[e] = (uint16_t)tk[e * LANES];
                for (uint32_t wd = lane; wd < nw; wd += 32) d32[wd] = ek[wd * LANES];
            }
            if (has) {
                ss.st[seg_base + slot] = st;
                ss.eb[seg_base + slot] = tab.epoch_base | (bail ? 1u : 0u);
            }
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(ss.flag + tile), "r"(seg + 1) : "memory");
            continue;
        }
        if (has) {
            // explicit output index per job, or (rectangle mode) row-major position of (x, y) in the rectangle
            const int64_t oi = tile_out ? tile_out[td.first + slot] : (td.first + slot) * out_stride + (td.y - col0);
            out[oi] = bail ? -1 : (int64_t)(st.total + lz4_frame_overhead(n));
        }
    }
}

// ---- singles + prefix checkpoints: one sequence per CTA, thread 0 parses, everyone refills the ring ----
// For task t = sequence s: out[out_idx[t]] = frame size of s alone (its own regime; skipped when
// out_idx[t] < 0); checkpoints of s as the x of a pair stream, for the regimes asked for in want[t]
// (bit 0: single-block regime, bit 1: linked regime).
struct PkSingleSmem {
    uint32_t tab[1024];
    uint32_t snap[1024];
};

template <int KIND, bool DETECT, bool EXC>
__device__ void pk_single_run(PkState &st, PkTab<KIND, 1> &tab, PkView &v, PkRing &rg, const uint64_t *yw, uint64_t *ring,
                              uint32_t n, uint32_t xend, uint32_t snap_bs, PkState *snap_st, uint32_t *snap_tab,
                              PkExact *xv, uint32_t *snap_ovf)
{
    // CTA-uniform control flow: warp 0 parses -- all its lanes carry the same state and run the same code, so that the
    // batches of pk_batch_run have the whole warp -- and all threads take part in ring refills
    __shared__ uint32_t s_more;
    for (;;) {
        __syncthreads();
        rg.view(v);
        if (EXC) v.dring = reinterpret_cast<const uint32_t *>(ring + rg.dring_word());
        const uint32_t stop = rg.stop_q();
        if (threadIdx.x < 32) {
            bool touched = false;
            while (st.phase != PK_DONE && pk_next_pos(st) < stop) {
                if (DETECT) {
                    bool t;
                    if constexpr (EXC && KIND != 1) t = pk_step_exact<KIND, 1, true>(st, tab, *xv, n, xend);
                    else t = pk_step<KIND, 1, true>(st, tab, v, n, xend);
                    if (t) { touched = true; break; }
                    continue;
                }
                uint32_t limit = stop;
                if (snap_st) {
                    // stop at the start of the last block: its state is where the prefix pass resumes
                    if (st.phase == PK_BLOCK_START && st.bs == snap_bs) {
                        *snap_st = st;
                        for (uint32_t e = threadIdx.x; e < PkTab<KIND, 1>::ENTRIES; e += 32) snap_tab[e] = tab.t[e];
                        if (EXC && snap_ovf) for (uint32_t e = threadIdx.x; e < PK_OVF_ENTRIES; e += 32) snap_ovf[e] = xv->ovf[e];
                        __syncwarp();
                        snap_st = nullptr;
                    } else {
                        limit = tmin(stop, snap_bs);
                    }
                }
                if constexpr (EXC && KIND != 1) pk_run_exc<KIND, 1>(st, tab, v, *xv, n, limit, rg.hi_w * 32, 0xffffffffu);
                else pk_run<KIND, 1, false, true>(st, tab, v, nullptr, n, limit, 0xffffffffu);
            }
            s_more = (!touched && st.phase != PK_DONE && !rg.complete()) ? 1u : 0u;
        }
        __syncthreads();
        if (!s_more) break;
        uint32_t w0, w1;
        rg.advance(w0, w1);
        pk_ring_fill(ring, yw, w0, w1);
        if (EXC) pk_dring_build(ring, rg, v.ymask);
    }
}

// EXC: the corpus holds sequences with bytes outside the alphabet.  Such a sequence is handled here when it is a
// linked-regime stream (> 64 KiB) or the x of one; the single-block regime (16-bit table, 4-byte hash) has no byte-exact
// step -- the host sends those jobs to the byte-wise kernels.
template <bool EXC>
__global__ void __launch_bounds__(64)
lz4_pk_single_kernel(PkCorpus pc, PkExcCorpus xc, const int32_t *__restrict__ seqs, const int32_t *__restrict__ want, int32_t n_seqs,
                     uint32_t *__restrict__ ck_tab, PkState *__restrict__ ck_state, const uint16_t *__restrict__ lut5_g,
                     const uint16_t *__restrict__ lut4_g, const int64_t *__restrict__ out_idx,
                     int64_t *__restrict__ out)
{
    __shared__ __align__(16) uint64_t ring[PK_RING_WORDS + 2];
    __shared__ uint32_t s_tab[1024];
    __shared__ uint32_t s_snap[1024];
    __shared__ uint16_t s_lut5[1024];
    __shared__ uint16_t s_lut4[256];
    for (uint32_t i = threadIdx.x; i < 1024; i += blockDim.x) s_lut5[i] = lut5_g[i];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_lut4[i] = lut4_g[i];
    for (int32_t t = blockIdx.x; t < n_seqs; t += gridDim.x) {
        const int32_t s = seqs[t];
        const int32_t wn = want[t];
        const uint32_t len = pc.len[s];
        const uint64_t *yw = pc.words + pc.woff[s];
        PkView v;
        v.ring = ring; v.yw = yw; v.xw = yw; v.lx = 0; v.dring = nullptr; v.ymask = EXC ? xc.mask + xc.moff[s] : nullptr;
        PkExact xv;
        uint32_t *ovf_w = nullptr, *ovf_ck = nullptr;
        if (EXC) {
            ovf_w = xc.ovf_work + (size_t)blockIdx.x * PK_OVF_ENTRIES;
            ovf_ck = xc.ck_ovf + (size_t)s * PK_OVF_ENTRIES;
            xv.b2s = xc.b2s; xv.ovf = ovf_w;
            xv.s.x = xc.bytes + xc.boff[s]; xv.s.y = xv.s.x + len; xv.s.lx = len; xv.s.n = len;
        }
        PkRing rg;
        PkState st, snap;
        PkTab<0, 1> tl; tl.t = s_tab; tl.lut = s_lut5; tl.ep = nullptr; tl.nslot = 1024; tl.epoch_base = 0;
        PkTab<1, 1> ts; ts.t = reinterpret_cast<uint16_t *>(s_tab); ts.lut = s_lut4; ts.ep = nullptr; ts.nslot = 256; ts.epoch_base = 0;
        const bool linked_single = len > LZ4_BLOCK;
        const uint32_t last_bs = (len / LZ4_BLOCK) * LZ4_BLOCK;

        // (1) the sequence on its own
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < 1024; i += blockDim.x) s_tab[i] = 0;
        if (EXC) for (uint32_t i = threadIdx.x; i < PK_OVF_ENTRIES; i += blockDim.x) ovf_w[i] = 0;
        uint32_t w0, w1;
        const uint32_t cover = EXC ? PK_RING_COVER_EXC : PK_RING_WORDS;
        rg.start(len, w0, w1, cover);
        pk_ring_fill(ring, yw, w0, w1);
        if (EXC) pk_dring_build(ring, rg, v.ymask);
        pk_fresh(st); pk_fresh(snap);
        if (linked_single) pk_single_run<0, false, EXC>(st, tl, v, rg, yw, ring, len, 0, last_bs, &snap, s_snap, &xv, ovf_ck);
        else               pk_single_run<1, false, EXC>(st, ts, v, rg, yw, ring, len, 0, 0, nullptr, nullptr, &xv, nullptr);
        if (threadIdx.x == 0 && out_idx[t] >= 0) out[out_idx[t]] = (int64_t)(st.total + lz4_frame_overhead(len));

        // (2) linked-regime checkpoint: from the snapshot at the last block start (or from scratch)
        if (wn & 2) {
            __syncthreads();
            if (linked_single) {
                // a sequence that is a whole number of blocks never reaches BLOCK_START(last_bs) before DONE
                // inside the run above only if last_bs == len; pk_step turns that state into DONE, so the
                // snapshot has been taken in both cases
                for (uint32_t i = threadIdx.x; i < 1024; i += blockDim.x) s_tab[i] = s_snap[i];
                st = snap;
            } else {
                for (uint32_t i = threadIdx.x; i < 1024; i += blockDim.x) s_tab[i] = 0;
                if (EXC) for (uint32_t i = threadIdx.x; i < PK_OVF_ENTRIES; i += blockDim.x) ovf_ck[i] = 0;
                pk_fresh(st);
                rg.start(len, w0, w1, cover);
                pk_ring_fill(ring, yw, w0, w1);
                if (EXC) pk_dring_build(ring, rg, v.ymask);
            }
            if (EXC) xv.ovf = ovf_ck;                       // the checkpoint's overflow table is updated in place
            pk_single_run<0, true, EXC>(st, tl, v, rg, yw, ring, 0xffffffffu, len, 0, nullptr, nullptr, &xv, nullptr);
            __syncthreads();
            uint32_t *dst = ck_tab + (size_t)(2 * s + 1) * PK_CKPT_TAB;
            for (uint32_t i = threadIdx.x; i < 1024; i += blockDim.x) dst[i] = s_tab[i];
            if (threadIdx.x == 0) ck_state[2 * s + 1] = st;
        }
        // (3) single-block-regime checkpoint (only meaningful when some pair stream with this x fits one block)
        if ((wn & 1) && len < LZ4_BLOCK) {
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < 1024; i += blockDim.x) s_tab[i] = 0;
            pk_fresh(st);
            rg.start(len, w0, w1);
            pk_ring_fill(ring, yw, w0, w1);
            pk_single_run<1, true, false>(st, ts, v, rg, yw, ring, 0xffffffffu, len, 0, nullptr, nullptr, nullptr, nullptr);
            __syncthreads();
            uint32_t *dst = ck_tab + (size_t)(2 * s) * PK_CKPT_TAB;
            const uint16_t *t16 = reinterpret_cast<const uint16_t *>(s_tab);
            for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) dst[i] = t16[i];
            if (threadIdx.x == 0) ck_state[2 * s] = st;
        }
    }
}
#endif  // __CUDACC__

}  // namespace snacc
