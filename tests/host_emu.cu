// tests/host_emu.cu -- TEST INFRASTRUCTURE.  Compiles the product's __host__ __device__ parse functions
// (snacc_b200/csrc/*.cuh) for the CPU so the exact kernel logic can be checked against the oracle in the
// GPU-less container.  Never loaded by the product package.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../snacc_b200/csrc/common.cuh"
#include "../snacc_b200/csrc/lz4.cuh"

using namespace snacc;

static uint8_t *padded_copy(const uint8_t *p, uint64_t n)
{
    uint8_t *q = (uint8_t *)aligned_alloc(16, (n + SEQ_PAD + 64 + 15) & ~15ull);
    memset(q, 0, (n + SEQ_PAD + 64 + 15) & ~15ull);
    if (n) memcpy(q, p, n);
    return q;
}

extern "C" int64_t emu_lz4_size(const uint8_t *x, uint32_t lx, const uint8_t *y, int64_t ly)
{
    uint8_t *px = padded_copy(x, lx);
    uint8_t *py = ly >= 0 ? padded_copy(y, (uint64_t)ly) : nullptr;
    Stream s;
    s.x = px; s.lx = lx;
    if (ly >= 0) { s.y = py; s.n = lx + (uint32_t)ly; } else { s.y = px + lx; s.n = lx; }
    std::vector<uint8_t> ck(LZ4_TABLE_BYTES), tab(LZ4_TABLE_BYTES);
    uint64_t ck_total = 0;
    const bool use = lx >= LZ4_BLOCK && s.n > LZ4_BLOCK;
    if (lx >= LZ4_BLOCK) {
        Stream sx = s; sx.y = px + lx; sx.n = lx;
        ck_total = lz4_prefix_state(sx, ck.data());
    }
    uint64_t r = lz4_frame_size(s, tab.data(), use ? ck.data() : nullptr, use ? ck_total : 0);
    free(px); free(py);
    return (int64_t)r;
}
