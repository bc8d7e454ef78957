// lz4.cuh -- LZ4-frame compressed-size kernels (K1/K2 of SURVEY.md 2.3), sm_100a.
//
// Reproduces len(lz4framed.compress(b)) (reference call site snacc/pairwise_ncd.py:80), i.e.
// LZ4F_compressFrame of LZ4 1.9.4 with 64 KiB blocks, linked, no checksums, level 0, content size in
// the header.  Only byte COUNTS are produced.
//
// One compressor stream per thread.  The stream's hash table (16 KiB) lives in a per-thread slab in
// global memory (L2 resident for the default number of streams in flight); stream bytes are read
// straight from the padded corpus through the two-segment accessor (x then y), nothing is copied.
// The parse is written as a single flat probe loop so the lanes of a warp -- which work on different
// streams -- re-converge on every probe.
//
// Prefix checkpoint: in the linked regime every 64 KiB block that lies wholly inside x is identical
// in frame(x) and frame(x+y).  lz4_prefix_kernel parses those blocks once per x and stores the hash
// table + byte total; pair jobs start from that state.
#pragma once
#include "common.cuh"

namespace snacc {

constexpr uint32_t LZ4_BLOCK = 65536;
constexpr uint32_t LZ4_TABLE_BYTES = 16384;      // 4096 x u32 (linked) or 8192 x u16 (single block)
constexpr uint32_t LZ4_MFLIMIT = 12;
constexpr uint32_t LZ4_LASTLITERALS = 5;
constexpr uint32_t LZ4_MINLENGTH = 13;
constexpr uint32_t LZ4_MAX_DISTANCE = 65535;

struct Lz4Job {
    int32_t x;
    int32_t y;     // < 0: single
};

// frame bytes that are not block payload: magic 4 + FLG 1 + BD 1 + HC 1 + content size 8 (absent for
// empty input) + end mark 4
SNACC_HD uint64_t lz4_frame_overhead(uint32_t n) { return n ? 19 : 11; }

template <bool U16> struct Lz4Table;
template <> struct Lz4Table<true> {      // single-block regime: 16-bit positions, hash4 -> 13 bits
    uint16_t *t;
    SNACC_HD static uint32_t hash(uint64_t seq) {
        return ((uint32_t)seq * 2654435761u) >> (32 - 13);
    }
    SNACC_HD uint32_t get(uint32_t h) const { return t[h]; }
    SNACC_HD void put(uint32_t h, uint32_t pos) { t[h] = (uint16_t)pos; }
};
template <> struct Lz4Table<false> {     // linked regime: 32-bit positions, hash5 -> 12 bits
    uint32_t *t;
    SNACC_HD static uint32_t hash(uint64_t seq) {
        return (uint32_t)(((seq << 24) * 889523592379ull) >> (64 - 12));
    }
    SNACC_HD uint32_t get(uint32_t h) const { return t[h]; }
    SNACC_HD void put(uint32_t h, uint32_t pos) { t[h] = pos; }
};

// number of equal bytes of stream[a..] and stream[b..] (b < a), a never reaching `lim`
SNACC_HD uint32_t lz4_count(const Stream &s, uint32_t a, uint32_t b, uint32_t lim)
{
    uint32_t a0 = a;
    while (a < lim) {
        uint64_t d = ld64(s, a) ^ ld64(s, b);
        if (d) {
            a += (uint32_t)(SNACC_FFS64(d) - 1) >> 3;
            break;
        }
        a += 8; b += 8;
    }
    if (a > lim) a = lim;
    return a - a0;
}

// Parse blocks [first_block, last_block) of the stream with the fast compressor and return the sum of
// (4 + stored payload) over them.  Table state is carried in `tab` exactly as the library carries it.
template <bool U16>
__host__ __device__ uint64_t lz4_parse_blocks(const Stream &s, Lz4Table<U16> tab, uint32_t first_block,
                                     uint32_t last_block)
{
    uint64_t total = 0;
    for (uint32_t blk = first_block; blk < last_block; ++blk) {
        const uint32_t bs = blk * LZ4_BLOCK;
        const uint32_t be = tmin(s.n, bs + LZ4_BLOCK);
        const uint32_t blen = be - bs;
        const uint32_t budget = blen - 1;          // frame layer grants srcSize - 1 output bytes
        uint32_t op = 0;                           // payload bytes so far; 0xffffffff = aborted
        uint32_t anchor = bs;
        if (blen >= LZ4_MINLENGTH) {
            const uint32_t mfl1 = be - LZ4_MFLIMIT + 1;
            const uint32_t mlim = be - LZ4_LASTLITERALS;
            tab.put(tab.hash(ld64(s, bs)), bs);
            uint32_t ip = bs;                      // position probed in this iteration
            uint32_t fip = bs + 1, step = 1, nb = 64;
            bool searching = true;
            for (;;) {
                if (searching) {
                    ip = fip; fip += step; step = (nb++ >> 6);
                    if (fip > mfl1) break;
                }
                const uint64_t cur = ld64(s, ip);
                const uint32_t h = tab.hash(cur);
                uint32_t m = tab.get(h);
                tab.put(h, ip);
                bool hit = U16 || (m + LZ4_MAX_DISTANCE >= ip);
                uint64_t diff = 0;
                if (hit) {
                    diff = cur ^ ld64(s, m);
                    hit = ((uint32_t)diff == 0);
                }
                if (hit) {
                    bool moved = false;
                    if (searching) {
                        // catch up over pending literals, then account the literal run
                        while (ip > anchor && m > 0 && ld8(s, ip - 1) == ld8(s, m - 1)) { --ip; --m; moved = true; }
                        const uint32_t lit = ip - anchor;
                        op += 1;
                        if (op + lit + 8 + lit / 255 > budget) { op = 0xffffffffu; break; }
                        if (lit >= 15) op += (lit - 15) / 255 + 1;
                        op += lit;
                    } else {
                        op += 1;                   // token with zero literals
                    }
                    op += 2;                       // offset
                    // match length beyond the 4 verified bytes, never reaching mlim
                    uint32_t mcode;
                    const uint32_t hi = (uint32_t)(diff >> 32);
                    if (moved) {
                        mcode = lz4_count(s, ip + 4, m + 4, mlim);
                    } else if (hi != 0) {
                        mcode = tmin((uint32_t)(SNACC_FFS32(hi) - 1) >> 3, mlim - ip - 4);
                    } else if (ip + 8 >= mlim) {
                        mcode = mlim - ip - 4;
                    } else {
                        mcode = 4 + lz4_count(s, ip + 8, m + 8, mlim);
                    }
                    ip += 4 + mcode;
                    if (op + 6 + (mcode + 240) / 255 > budget) { op = 0xffffffffu; break; }
                    if (mcode >= 15) op += (mcode - 15) / 255 + 1;
                    anchor = ip;
                    if (ip >= mfl1) break;
                    tab.put(tab.hash(ld64(s, ip - 2)), ip - 2);
                    searching = false;             // immediate re-test at ip
                } else if (!searching) {
                    searching = true; fip = ip + 1; step = 1; nb = 64;
                }
            }
        }
        if (op != 0xffffffffu) {
            const uint32_t last_run = be - anchor;
            if (op + last_run + 1 + (last_run + 240) / 255 > budget) op = 0xffffffffu;
            else op += 1 + (last_run >= 15 ? (last_run - 15) / 255 + 1 : 0) + last_run;
        }
        const uint32_t payload = (op == 0xffffffffu || op >= blen) ? blen : op;
        total += 4 + payload;
    }
    return total;
}

SNACC_HD Stream make_stream(const uint8_t *corpus, const uint64_t *off, const uint32_t *len,
                                              int32_t x, int32_t y)
{
    Stream s;
    s.x = corpus + off[x];
    s.lx = len[x];
    if (y >= 0) { s.y = corpus + off[y]; s.n = s.lx + len[y]; }
    else        { s.y = s.x + s.lx;      s.n = s.lx; }
    return s;
}

// Prefix state of x in the linked regime: zero the table, parse the full blocks inside x.
// Returns the sum of (4 + payload) over those blocks; `tab` (16 KiB) is left in the carried state.
__host__ __device__ inline uint64_t lz4_prefix_state(Stream s, uint8_t *tab)
{
    const uint32_t full = s.lx / LZ4_BLOCK;
    s.n = full * LZ4_BLOCK;          // only whole blocks; each is parsed exactly as inside a longer stream
    uint4 *z = reinterpret_cast<uint4 *>(tab);
    for (uint32_t i = 0; i < LZ4_TABLE_BYTES / 16; ++i) z[i] = make_uint4(0, 0, 0, 0);
    return lz4_parse_blocks<false>(s, Lz4Table<false>{reinterpret_cast<uint32_t *>(tab)}, 0, full);
}

// len(lz4framed.compress(stream)).  `ckpt` / `ckpt_total`: prefix state of x (may be null when x owns no
// full block); `mytab`: 16 KiB of scratch owned by the caller.
__host__ __device__ inline uint64_t lz4_frame_size(const Stream &s, uint8_t *mytab, const uint8_t *ckpt,
                                                   uint64_t ckpt_total)
{
    uint64_t total = lz4_frame_overhead(s.n);
    uint4 *z = reinterpret_cast<uint4 *>(mytab);
    if (s.n == 0) {
        // empty frame: header + end mark only
    } else if (s.n <= LZ4_BLOCK) {
        // one-shot rule: the frame becomes a single independent block, 16-bit table
        for (uint32_t i = 0; i < LZ4_TABLE_BYTES / 16; ++i) z[i] = make_uint4(0, 0, 0, 0);
        total += lz4_parse_blocks<true>(s, Lz4Table<true>{reinterpret_cast<uint16_t *>(mytab)}, 0, 1);
    } else {
        const uint32_t full_x = s.lx / LZ4_BLOCK;     // blocks wholly inside x: from the checkpoint
        const uint32_t nblk = (s.n + LZ4_BLOCK - 1) / LZ4_BLOCK;
        if (full_x > 0) {
            const uint4 *c = reinterpret_cast<const uint4 *>(ckpt);
            for (uint32_t i = 0; i < LZ4_TABLE_BYTES / 16; ++i) z[i] = c[i];
            total += ckpt_total;
        } else {
            for (uint32_t i = 0; i < LZ4_TABLE_BYTES / 16; ++i) z[i] = make_uint4(0, 0, 0, 0);
        }
        total += lz4_parse_blocks<false>(s, Lz4Table<false>{reinterpret_cast<uint32_t *>(mytab)}, full_x, nblk);
    }
    return total;
}

// ---- K2a: per-sequence prefix state (linked regime): parse the full blocks inside x once ----
// ckpt_tab: n_ck slabs of 16 KiB; ckpt_total[slot]: sum of (4 + payload) over those blocks.
__global__ void lz4_prefix_kernel(const uint8_t *__restrict__ corpus, const uint64_t *__restrict__ off,
                                  const uint32_t *__restrict__ len, const int32_t *__restrict__ todo,
                                  int32_t n_todo, const int32_t *__restrict__ slot_of,
                                  uint8_t *__restrict__ ckpt_tab, uint64_t *__restrict__ ckpt_total)
{
    const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_todo) return;
    const int32_t x = todo[t];
    const int32_t slot = slot_of[x];
    const Stream s = make_stream(corpus, off, len, x, -1);
    ckpt_total[slot] = lz4_prefix_state(s, ckpt_tab + (size_t)slot * LZ4_TABLE_BYTES);
}

// ---- K1/K2b: frame size of every job; persistent threads pull jobs from a counter ----
// `sel` (optional): the launch covers jobs sel[0..n_jobs) of the job arrays instead of 0..n_jobs.
__global__ void lz4_stream_kernel(const uint8_t *__restrict__ corpus, const uint64_t *__restrict__ off,
                                  const uint32_t *__restrict__ len, const int32_t *__restrict__ job_x,
                                  const int32_t *__restrict__ job_y, const int64_t *__restrict__ sel, int64_t n_jobs,
                                  const int32_t *__restrict__ slot_of, const uint8_t *__restrict__ ckpt_tab,
                                  const uint64_t *__restrict__ ckpt_total, uint8_t *__restrict__ work_tab,
                                  unsigned long long *__restrict__ counter, int64_t *__restrict__ out)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint8_t *mytab = work_tab + (size_t)tid * LZ4_TABLE_BYTES;
    for (;;) {
        const long long k = (long long)atomicAdd(counter, 1ull);
        if (k >= n_jobs) break;
        const long long j = sel ? sel[k] : k;
        const int32_t x = job_x[j];
        const int32_t y = job_y ? job_y[j] : -1;
        const Stream s = make_stream(corpus, off, len, x, y);
        const int32_t slot = slot_of[x];
        const bool use = slot >= 0 && s.n > LZ4_BLOCK;
        out[j] = (int64_t)lz4_frame_size(s, mytab, use ? ckpt_tab + (size_t)slot * LZ4_TABLE_BYTES : nullptr,
                                         use ? ckpt_total[slot] : 0);
    }
}

}  // namespace snacc
