// tests/host_emu.cu -- TEST INFRASTRUCTURE.  Compiles the product's __host__ __device__ parse functions
// (snacc_b200/csrc/*.cuh) for the CPU so the exact kernel logic can be checked against the oracle in the
// GPU-less container.  Never loaded by the product package.
#define PK_COUNT_STEPS 1
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

// ---------------------------------------------------------------------------------------------
// packed (2-bit) LZ4 path: the same pk_step / PkRing / checkpoint logic the tile kernels run, driven
// sequentially.  ring refills, the global-memory fallback, prefix checkpoints and resume are all exercised.
// ---------------------------------------------------------------------------------------------
#include "../snacc_b200/csrc/pack.cuh"
#include "../snacc_b200/csrc/lz4_packed.cuh"

static std::vector<uint64_t> pack_host(const PkAlphabet &a, const uint8_t *p, uint32_t n, bool &ok)
{
    std::vector<uint64_t> w(pk_words(n) + 2, 0);
    for (uint32_t i = 0; i < n; ++i) {
        const uint8_t c = a.code_of[p[i]];
        if (c > 3) { ok = false; return w; }
        w[i >> 5] |= (uint64_t)c << (2 * (i & 31));
    }
    return w;
}

static void ring_fill_host(std::vector<uint64_t> &ring, const std::vector<uint64_t> &yw, uint32_t w0, uint32_t w1)
{
    for (uint32_t i = w0; i < w1; ++i) {
        ring[i & (PK_RING_WORDS - 1)] = yw[i];
        if ((i & (PK_RING_WORDS - 1)) < 2) ring[PK_RING_WORDS + (i & 1)] = yw[i];      // the mirror pk_ring_fill keeps
    }
}

struct EmuCkpt { PkState st; std::vector<uint32_t> tab; };
static int emu_batch = 0;            // 1: the KIND 0 / 1 streams run pk_batch_run (the singles kernel's 32-probe batches)
extern "C" void emu_set_batch(int on) { emu_batch = on; }

// run a stream until DONE (or until DETECT touches); mirrors pk_single_run / the pair kernel loop
template <int KIND, bool DETECT>
static bool emu_run(PkState &st, PkTab<KIND, 1> &tab, std::vector<uint32_t> *tab32, PkView &v, PkRing &rg,
                    std::vector<uint64_t> &ring, const std::vector<uint64_t> &yw, uint32_t n, uint32_t xend,
                    uint32_t snap_bs, EmuCkpt *snap)
{
    for (;;) {
        rg.view(v);
        const uint32_t sq = rg.stop_q();
        const uint32_t stop = sq == 0xffffffffu ? sq : v.lx + sq;
        while (st.phase != PK_DONE && pk_next_pos(st) < stop) {
            if (DETECT) {
                if (pk_step<KIND, 1, true>(st, tab, v, n, xend)) return true;
                continue;
            }
            uint32_t limit = stop;
            if (snap) {
                if (st.phase == PK_BLOCK_START && st.bs == snap_bs) { snap->st = st; snap->tab = *tab32; snap = nullptr; }
                else limit = tmin(stop, snap_bs);
            }
            if (emu_batch) pk_run<KIND, 1, false, true>(st, tab, v, nullptr, n, limit, 1u);
            else pk_run<KIND, 1>(st, tab, v, nullptr, n, limit, 1u);
        }
        if (st.phase == PK_DONE || rg.complete()) return false;
        uint32_t w0, w1;
        rg.advance(w0, w1);
        ring_fill_host(ring, yw, w0, w1);
    }
}

// the shared-memory code->slot map of the kernels: slot indices
template <int KIND> static std::vector<uint16_t> emu_lut(const uint16_t *raw)
{
    return std::vector<uint16_t>(raw, raw + PkTab<KIND, 1>::ENTRIES);
}

static int emu_segments = 1;
extern "C" void emu_set_segments(int k) { emu_segments = k < 1 ? 1 : k; }

// returns the size, -1 when the packed pair path bails out, -2 when the input is not packable
extern "C" int64_t emu_lz4_packed(const uint8_t *x, uint32_t lx, const uint8_t *y, int64_t ly_)
{
    unsigned long long hist[256] = {0};
    for (uint32_t i = 0; i < lx; ++i) hist[x[i]]++;
    for (int64_t i = 0; i < ly_; ++i) hist[y[i]]++;
    const PkAlphabet a = pk_choose_alphabet(hist);
    bool ok = true;
    std::vector<uint64_t> xw = pack_host(a, x, lx, ok);
    std::vector<uint64_t> yw = ly_ >= 0 ? pack_host(a, y, (uint32_t)ly_, ok) : std::vector<uint64_t>();
    if (!ok) return -2;
    uint16_t raw5[1024], raw4[256];
    const uint32_t nslot5 = pk_slot_lut(a, false, raw5);
    pk_slot_lut(a, true, raw4);
    const std::vector<uint16_t> lut32 = emu_lut<0>(raw5), lut16 = emu_lut<1>(raw4), lut17 = emu_lut<2>(raw5);
    std::vector<uint64_t> ring(PK_RING_WORDS + 2, 0);
    std::vector<uint32_t> tab(1024, 0);
    PkTab<0, 1> t32; t32.t = tab.data(); t32.ep = nullptr; t32.nslot = 1024; t32.epoch_base = 0; t32.lut = lut32.data();
    PkTab<1, 1> t16; t16.t = reinterpret_cast<uint16_t *>(tab.data()); t16.ep = nullptr; t16.nslot = 256; t16.epoch_base = 0; t16.lut = lut16.data();
    PkView v; PkRing rg; PkState st; uint32_t w0, w1;

    // ---- the sequence x on its own (kernel lz4_pk_single_kernel, step 1) ----
    const bool linked_single = lx > LZ4_BLOCK;
    const uint32_t last_bs = (lx / LZ4_BLOCK) * LZ4_BLOCK;
    EmuCkpt snap; pk_fresh(snap.st); snap.tab.assign(1024, 0);
    v.ring = ring.data(); v.yw = xw.data(); v.xw = xw.data(); v.lx = 0;
    rg.start(lx, w0, w1); ring_fill_host(ring, xw, w0, w1);
    pk_fresh(st);
    if (linked_single) emu_run<0, false>(st, t32, &tab, v, rg, ring, xw, lx, 0, last_bs, &snap);
    else               emu_run<1, false>(st, t16, &tab, v, rg, ring, xw, lx, 0, 0, nullptr);
    const int64_t single = (int64_t)(st.total + lz4_frame_overhead(lx));
    if (ly_ < 0) return single;

    // ---- prefix checkpoint of x in the regime of the pair stream (steps 2 / 3) ----
    const uint32_t ly = (uint32_t)ly_;
    const uint32_t n = lx + ly;
    const bool u16 = n <= LZ4_BLOCK;
    if (ly < 16) return -1;          // host-side eligibility rule of the packed pair path
    EmuCkpt ck;
    if (!u16) {
        if (linked_single) { tab = snap.tab; st = snap.st; }
        else { tab.assign(1024, 0); pk_fresh(st); rg.start(lx, w0, w1); ring_fill_host(ring, xw, w0, w1); }
        t32.t = tab.data();
        if (!emu_run<0, true>(st, t32, &tab, v, rg, ring, xw, 0xffffffffu, lx, 0, nullptr)) return -3;
        ck.st = st; ck.tab = tab;
    } else {
        tab.assign(1024, 0); pk_fresh(st); rg.start(lx, w0, w1); ring_fill_host(ring, xw, w0, w1);
        t16.t = reinterpret_cast<uint16_t *>(tab.data());
        if (!emu_run<1, true>(st, t16, &tab, v, rg, ring, xw, 0xffffffffu, lx, 0, nullptr)) return -3;
        ck.st = st; ck.tab.assign(1024, 0);
        const uint16_t *p16 = reinterpret_cast<const uint16_t *>(tab.data());
        for (int i = 0; i < 256; ++i) ck.tab[i] = p16[i];
    }

    // ---- the pair stream resumes from the checkpoint (kernel lz4_pk_pair_kernel) ----
    st = ck.st;
    if (!pk_resume(st, n)) return -1;
    v.ring = ring.data(); v.yw = yw.data(); v.xw = xw.data(); v.lx = lx;
    rg.start(ly, w0, w1); ring_fill_host(ring, yw, w0, w1);
    if (u16) {
        std::vector<uint32_t> t(1024, 0);
        uint16_t *p16 = reinterpret_cast<uint16_t *>(t.data());
        for (int i = 0; i < 256; ++i) p16[i] = (uint16_t)ck.tab[i];
        t16.t = p16;
        emu_run<1, false>(st, t16, &t, v, rg, ring, yw, n, 0, 0, nullptr);
    } else {
        // linked regime: 16-bit slots + epoch bit plane (PkTab KIND 2), imported from the 32-bit checkpoint
        const uint32_t nslot = (nslot5 + 1) & ~1u;
        std::vector<uint16_t> lo(nslot, 0);
        std::vector<uint32_t> ep((nslot + 31) / 32, 0);
        PkTab<2, 1> t17; t17.t = lo.data(); t17.ep = ep.data(); t17.nslot = nslot; t17.epoch_base = ck.st.bs; t17.lut = lut17.data();
        for (uint32_t e = 0; e < nslot; ++e) t17.import_slot(e, ck.tab[e], ck.st.bs);
        if (emu_segments > 1 && rg.runs() >= (uint32_t)emu_segments) {
            // tile segments (PkSegStore): at a cut, everything but the stream's records is thrown away and the ring is
            // rebuilt by PkRing::restart, as when another CTA draws the next segment
            const uint32_t runs = rg.runs();
            for (int seg = 0; seg < emu_segments; ++seg) {
                uint32_t r = (uint32_t)((uint64_t)runs * seg / emu_segments);
                const uint32_t r1 = (uint32_t)((uint64_t)runs * (seg + 1) / emu_segments);
                if (seg > 0) {
                    std::fill(ring.begin(), ring.end(), 0x5a5a5a5a5a5a5a5aull);
                    rg.restart(ly, r, w0, w1);
                    ring_fill_host(ring, yw, w0, w1);
                }
                for (;;) {
                    rg.view(v);
                    const uint32_t sq = rg.stop_q();
                    pk_run<2, 1>(st, t17, v, nullptr, n, sq == 0xffffffffu ? sq : v.lx + sq, 1u);
                    if (++r >= r1) break;
                    rg.advance(w0, w1);
                    ring_fill_host(ring, yw, w0, w1);
                }
                if (seg + 1 < emu_segments) {
                    const std::vector<uint16_t> s_lo = lo; const std::vector<uint32_t> s_ep = ep;
                    const PkState s_st = st; const uint32_t s_eb = t17.epoch_base;
                    std::fill(lo.begin(), lo.end(), (uint16_t)0xdead); std::fill(ep.begin(), ep.end(), 0xdeadbeefu);
                    st = PkState(); t17.epoch_base = 12345;
                    lo = s_lo; ep = s_ep; st = s_st; t17.epoch_base = s_eb;
                    t17.t = lo.data(); t17.ep = ep.data();
                }
            }
            if (!rg.complete()) return -4;
        } else {
            emu_run<2, false>(st, t17, nullptr, v, rg, ring, yw, n, 0, 0, nullptr);
        }
    }
    return (int64_t)(st.total + lz4_frame_overhead(n));
}

// ---------------------------------------------------------------------------------------------
// byte-exact LZ4 step (pk_step_exact) on its own: every step of the stream through the true-byte path -- any byte
// values, the alphabet's buckets in the slot table, all other buckets in the overflow table -- singles and pairs
// (prefix checkpoint by the DETECT variant, resume on the 17-bit table).  Linked regime only: returns -5 otherwise.
// ---------------------------------------------------------------------------------------------
static std::vector<uint16_t> emu_b2s(const PkAlphabet &a, const uint16_t *raw5)
{
    std::vector<uint16_t> b2s(PK_OVF_ENTRIES, 0xffff);
    for (uint32_t c = 0; c < 1024; ++c) b2s[pk_bucket(a, c, false)] = raw5[c];
    return b2s;
}

extern "C" int64_t emu_lz4_exact(const uint8_t *x, uint32_t lx, const uint8_t *y, int64_t ly_)
{
    unsigned long long hist[256] = {0};
    for (uint32_t i = 0; i < lx; ++i) hist[x[i]]++;
    for (int64_t i = 0; i < ly_; ++i) hist[y[i]]++;
    const PkAlphabet a = pk_choose_alphabet(hist);
    uint16_t raw5[1024];
    const uint32_t nslot5 = pk_slot_lut(a, false, raw5);
    const std::vector<uint16_t> b2s = emu_b2s(a, raw5);
    uint8_t *px = padded_copy(x, lx);
    uint8_t *py = ly_ >= 0 ? padded_copy(y, (uint64_t)ly_) : nullptr;
    std::vector<uint32_t> tab(1024, 0), ovf(PK_OVF_ENTRIES, 0);
    PkTab<0, 1> t32; t32.t = tab.data(); t32.ep = nullptr; t32.nslot = 1024; t32.epoch_base = 0; t32.lut = raw5;
    PkExact xv; xv.b2s = b2s.data(); xv.ovf = ovf.data();
    PkState st;
    int64_t result = -5;
    if (ly_ < 0) {
        if (lx > LZ4_BLOCK) {
            xv.s.x = px; xv.s.y = px + lx; xv.s.lx = lx; xv.s.n = lx;
            pk_fresh(st);
            while (st.phase != PK_DONE) pk_step_exact<0, 1, false>(st, t32, xv, lx, 0);
            result = (int64_t)(st.total + lz4_frame_overhead(lx));
        }
    } else if ((uint64_t)lx + (uint64_t)ly_ > LZ4_BLOCK) {
        const uint32_t n = lx + (uint32_t)ly_;
        // prefix checkpoint of x: the state right before the first iteration that looks at a byte of y
        xv.s.x = px; xv.s.y = px + lx; xv.s.lx = lx; xv.s.n = lx;
        pk_fresh(st);
        bool touched = false;
        while (st.phase != PK_DONE && !(touched = pk_step_exact<0, 1, true>(st, t32, xv, 0xffffffffu, lx))) {}
        if (touched && pk_resume(st, n)) {
            const uint32_t nslot = (nslot5 + 1) & ~1u;
            std::vector<uint16_t> lo(nslot, 0);
            std::vector<uint32_t> ep((nslot + 31) / 32, 0);
            PkTab<2, 1> t17; t17.t = lo.data(); t17.ep = ep.data(); t17.nslot = nslot; t17.epoch_base = st.bs; t17.lut = raw5;
            for (uint32_t e = 0; e < nslot; ++e) t17.import_slot(e, tab[e], st.bs);
            xv.s.y = py; xv.s.n = n;
            while (st.phase != PK_DONE) pk_step_exact<2, 1, false>(st, t17, xv, n, 0);
            result = (int64_t)(st.total + lz4_frame_overhead(n));
        } else {
            result = -1;
        }
    }
    free(px); free(py);
    return result;
}

// ---------------------------------------------------------------------------------------------
// packed path for sequences with a few bytes outside the alphabet (the EXC kernels): filler code + per-base mask,
// dirty ring, fast loop up to the flagged granules (pk_run_exc), byte-exact steps across them and for every general
// step, overflow table, prefix checkpoint by the byte-exact DETECT step.  Linked regime only (-5 otherwise).
// ---------------------------------------------------------------------------------------------
struct EmuSeqX { std::vector<uint64_t> words; std::vector<uint32_t> mask; uint8_t *bytes; uint32_t len; };

static EmuSeqX emu_pack_x(const PkAlphabet &a, const uint8_t *p, uint32_t n)
{
    EmuSeqX q;
    q.len = n;
    q.words.assign(pk_words(n) + 2, 0);
    q.mask.assign(pk_mask_words(n) + 64, 0);
    for (uint32_t i = 0; i < n; ++i) {
        const uint8_t c = a.code_of[p[i]];
        if (c > 3) q.mask[i >> 5] |= 1u << (i & 31);
        else q.words[i >> 5] |= (uint64_t)c << (2 * (i & 31));
    }
    q.bytes = padded_copy(p, n);
    return q;
}

// what the EXC kernels do at every ring (re)fill: packed words [w0, w1), then the dirty ring rebuilt inside the ring
static void ring_fill_x(std::vector<uint64_t> &ring, const PkRing &rg, const EmuSeqX &y, uint32_t w0, uint32_t w1)
{
    ring_fill_host(ring, y.words, w0, w1);
    uint32_t *dr = reinterpret_cast<uint32_t *>(ring.data() + rg.dring_word());
    for (uint32_t k = 0; k < PK_DRING_WORDS; ++k) dr[k] = pk_dring_entry(rg, y.mask.data(), k);
    dr[PK_DRING_WORDS] = dr[0];
}

// one stream through pk_run_exc with ring refills (the loop of lz4_pk_pair_kernel<.., EXC> / the singles kernel)
template <int KIND>
static void emu_run_x(PkState &st, PkTab<KIND, 1> &tab, PkView &v, const PkExact &xv, const EmuSeqX &y, uint32_t n,
                      std::vector<uint64_t> &ring, uint32_t limit)
{
    PkRing rg; uint32_t w0, w1;
    rg.start(y.len, w0, w1, PK_RING_COVER_EXC); ring_fill_x(ring, rg, y, w0, w1);
    // tile segments (emu_set_segments; pair streams only): at a cut the ring is thrown away and rebuilt by
    // PkRing::restart, as when another CTA draws the next segment (table, state and overflow table stay where they are)
    const uint32_t runs = rg.runs(), k = (KIND == 2 && limit == 0xffffffffu && emu_segments > 1 && runs >= (uint32_t)emu_segments) ? emu_segments : 1;
    uint32_t r = 0;
    for (uint32_t seg = 0; seg < k; ++seg) {
        const uint32_t r1 = (uint32_t)((uint64_t)runs * (seg + 1) / k);
        if (seg > 0) {
            std::fill(ring.begin(), ring.end(), 0x5a5a5a5a5a5a5a5aull);
            rg.restart(y.len, r, w0, w1, PK_RING_COVER_EXC);
            ring_fill_x(ring, rg, y, w0, w1);
        }
        for (;;) {
            rg.view(v);
            v.dring = reinterpret_cast<const uint32_t *>(ring.data() + rg.dring_word());
            const uint32_t sq = rg.stop_q();
            const uint32_t stop = tmin(limit, sq == 0xffffffffu ? sq : v.lx + sq);
            pk_run_exc<KIND, 1>(st, tab, v, xv, n, stop, rg.hi_w * 32, 1u);
            if (k == 1 && (st.phase == PK_DONE || rg.complete() || pk_next_pos(st) >= limit)) return;
            if (++r >= r1) break;
            rg.advance(w0, w1);
            ring_fill_x(ring, rg, y, w0, w1);
        }
    }
}

extern "C" int64_t emu_lz4_packed_exc(const uint8_t *x, uint32_t lx, const uint8_t *y, int64_t ly_)
{
    unsigned long long hist[256] = {0};
    for (uint32_t i = 0; i < lx; ++i) hist[x[i]]++;
    for (int64_t i = 0; i < ly_; ++i) hist[y[i]]++;
    const PkAlphabet a = pk_choose_alphabet(hist);
    uint16_t raw5[1024];
    const uint32_t nslot5 = pk_slot_lut(a, false, raw5);
    const std::vector<uint16_t> b2s = emu_b2s(a, raw5);
    EmuSeqX sx = emu_pack_x(a, x, lx), sy;
    if (ly_ >= 0) sy = emu_pack_x(a, y, (uint32_t)ly_); else sy.bytes = nullptr;
    std::vector<uint64_t> ring(PK_RING_WORDS + 2, 0);
    std::vector<uint32_t> tab(1024, 0), ovf(PK_OVF_ENTRIES, 0);
    PkTab<0, 1> t32; t32.t = tab.data(); t32.ep = nullptr; t32.nslot = 1024; t32.epoch_base = 0; t32.lut = raw5;
    PkExact xv; xv.b2s = b2s.data(); xv.ovf = ovf.data();
    PkView v; v.ring = ring.data(); v.dring = nullptr;
    PkState st;
    int64_t result = -5;
    const uint32_t last_bs = (lx / LZ4_BLOCK) * LZ4_BLOCK;
    if (lx > LZ4_BLOCK || (ly_ >= 16 && (uint64_t)lx + (uint64_t)ly_ > LZ4_BLOCK)) {
        // ---- x on its own (singles kernel): size when it is a linked-regime stream; the state at its last block start ----
        v.yw = sx.words.data(); v.xw = sx.words.data(); v.lx = 0; v.ymask = sx.mask.data();
        xv.s.x = sx.bytes; xv.s.y = sx.bytes + lx; xv.s.lx = lx; xv.s.n = lx;
        pk_fresh(st);
        if (lx > LZ4_BLOCK) {
            // up to the last block start with the fast loop, snapshot, then on to the end for the size of x alone
            emu_run_x<0>(st, t32, v, xv, sx, lx, ring, last_bs);
            while (!(st.phase == PK_BLOCK_START && st.bs == last_bs) && st.phase != PK_DONE) pk_step_exact<0, 1, false>(st, t32, xv, lx, 0);
            const PkState snap = st; const std::vector<uint32_t> stab = tab, sovf = ovf;
            emu_run_x<0>(st, t32, v, xv, sx, lx, ring, 0xffffffffu);
            result = (int64_t)(st.total + lz4_frame_overhead(lx));
            st = snap; tab = stab; ovf = sovf;
        }
        if (ly_ >= 0) {
            // ---- prefix checkpoint: byte-exact DETECT steps over the last partial block of x ----
            const uint32_t n = lx + (uint32_t)ly_;
            bool touched = false;
            while (st.phase != PK_DONE && !(touched = pk_step_exact<0, 1, true>(st, t32, xv, 0xffffffffu, lx))) {}
            result = -1;
            if (touched && pk_resume(st, n)) {
                // ---- the pair stream (pair kernel, EXC) ----
                const uint32_t nslot = (nslot5 + 1) & ~1u;
                std::vector<uint16_t> lo(nslot, 0);
                std::vector<uint32_t> ep((nslot + 31) / 32, 0);
                PkTab<2, 1> t17; t17.t = lo.data(); t17.ep = ep.data(); t17.nslot = nslot; t17.epoch_base = st.bs; t17.lut = raw5;
                for (uint32_t e = 0; e < nslot; ++e) t17.import_slot(e, tab[e], st.bs);
                v.yw = sy.words.data(); v.xw = sx.words.data(); v.lx = lx; v.ymask = sy.mask.data();
                xv.s.y = sy.bytes; xv.s.n = n;
                emu_run_x<2>(st, t17, v, xv, sy, n, ring, 0xffffffffu);
                result = (int64_t)(st.total + lz4_frame_overhead(n));
            }
        }
    }
    free(sx.bytes); free(sy.bytes);
    return result;
}

// ---------------------------------------------------------------------------------------------
// deflate path: the same index / F table / junction / checkpoint / parse logic the kernels run
// ---------------------------------------------------------------------------------------------
#define DFL_CHECK_COMPACT 1
#include "../snacc_b200/csrc/deflate.cuh"

struct EmuIndex { std::vector<uint32_t> order, bstart; };
static EmuIndex emu_index(const uint8_t *p, uint32_t len)
{
    EmuIndex ix; ix.bstart.assign(DFL_HASH + 1, 0);
    const uint32_t nidx = len >= 3 ? len - 2 : 0;
    ix.order.assign(nidx + 4, 0);
    std::vector<uint32_t> cnt(DFL_HASH, 0);
    for (uint32_t i = 0; i < nidx; ++i) cnt[dfl_hash3(p[i], p[i + 1], p[i + 2])]++;
    uint32_t acc = 0;
    for (uint32_t h = 0; h < DFL_HASH; ++h) { ix.bstart[h] = acc; acc += cnt[h]; cnt[h] = ix.bstart[h]; }
    ix.bstart[DFL_HASH] = acc;
    for (uint32_t i = 0; i < nidx; ++i) ix.order[cnt[dfl_hash3(p[i], p[i + 1], p[i + 2])]++] = i;
    return ix;
}

// F (and, like the kernels at level 6, the quartered-chain table FQ) of a sequence alone, plus what
// dfl_match_kernel / dfl_head_kernel keep about its head
struct EmuSeq { std::vector<uint32_t> F, FQ; std::vector<uint16_t> head_order, head_visit; };
static int32_t emu_match_mismatches = 0;     // dfl_match_word (6-byte index shortcut) vs dfl_f_word, over everything computed so far
static EmuSeq emu_F_single(const DflStream &d, const DflConfig &cfg, bool want_q)
{
    EmuSeq e;
    // transient 6-byte index of the whole sequence (dfl_index_kernel<1>), level 9 only -- as deflate_run does
    std::vector<uint32_t> order6, bstart6(DFL_HASH + 1, 0);
    const bool i6on = !want_q;
    if (i6on) {
        const uint32_t n6 = d.s.n >= 6 ? d.s.n - 5 : 0;
        order6.assign(n6 + 4, 0);
        std::vector<uint32_t> cnt(DFL_HASH + 1, 0);
        for (uint32_t i = 0; i < n6; ++i) cnt[dfl_hash6w(ldu64(d.s.x + i)) + 1]++;
        for (uint32_t h = 0; h < DFL_HASH; ++h) cnt[h + 1] += cnt[h];
        for (uint32_t h = 0; h <= DFL_HASH; ++h) bstart6[h] = cnt[h];
        for (uint32_t i = 0; i < n6; ++i) order6[cnt[dfl_hash6w(ldu64(d.s.x + i))]++] = i;
    }
    const DflIndex6 i6{order6.data(), bstart6.data()};
    e.F.assign(d.s.n + 16, 0);
    if (want_q) e.FQ.assign(d.s.n + 16, 0);
    e.head_visit.assign(DFL_JY, 0);
    const uint32_t nidx = d.s.n >= 3 ? d.s.n - 2 : 0;
    for (uint32_t k = 0; k < nidx; ++k) {
        const uint32_t p = d.ix.order[k];
        uint32_t q, visit;
        e.F[p] = dfl_match_word(d, p, cfg, k, i6on ? &i6 : nullptr, &q, &visit);
        if (i6on && p >= DFL_JY) {
            uint32_t q0, v0;
            if (dfl_f_word(d, p, cfg, k, &q0, &v0) != e.F[p] || q0 != q) ++emu_match_mismatches;
        }
        if (want_q) e.FQ[p] = q;
        if (p < DFL_JY) { e.head_visit[p] = (uint16_t)visit; e.head_order.push_back((uint16_t)p); }   // index order = (hash, position)
    }
    return e;
}

// canonical symbol stream of a sequence alone, built the way dfl_parse_kernel (kind 3) + the dfl_cum kernels do
struct EmuCanon { std::vector<uint32_t> end, cum; std::vector<uint16_t> code; uint32_t n_sym = 0; };

// returns the raw deflate size of x (ly < 0) or of x followed by y.  use_canon: take the remaining blocks of y
// from the canonical stream of y (the product path); info[0] = 1 when the shortcut was taken, info[1] = 1 when a
// block that might be stored sent the job back to the serial parse
extern "C" int64_t emu_deflate_size_ex(const uint8_t *x, uint32_t lx, const uint8_t *y, int64_t ly_, int level, int use_canon,
                                       int32_t *info)
{
    const DflConfig cfg = dfl_config(level);
    if (info) info[0] = info[1] = info[2] = info[3] = 0;
    uint8_t *px = padded_copy(x, lx);
    uint8_t *py = ly_ >= 0 ? padded_copy(y, (uint64_t)ly_) : nullptr;
    EmuIndex ix = emu_index(px, lx), iy;
    if (ly_ >= 0) iy = emu_index(py, (uint32_t)ly_);
    DflTrees *tr = new DflTrees();
    std::vector<uint16_t> lf(DFL_L_CODES, 0), df(DFL_D_CODES, 0);
    // x alone
    DflStream dx; dx.s.x = px; dx.s.lx = lx; dx.s.y = px + lx; dx.s.n = lx; dx.pair = false;
    dx.ix.order = ix.order.data(); dx.ix.bstart = ix.bstart.data(); dx.iy = dx.ix;
    const bool want_q = level != 9;
    EmuSeq ex = emu_F_single(dx, cfg, want_q);
    std::vector<uint32_t> &Fx = ex.F;
    DflFView fv; fv.fx = fv.fy = fv.fj = Fx.data(); fv.jx0 = fv.jend = fv.lx = lx;
    fv.qx = fv.qy = fv.qj = want_q ? ex.FQ.data() : nullptr;
    DflParseState st;
    int64_t result;
    if (ly_ < 0) {
        // kind 3: checkpoint stop in the middle, symbols recorded -- the size must not care
        std::vector<uint32_t> rend(lx / 4 + 1024); std::vector<uint16_t> rcode(lx / 4 + 1024);
        DflRec rec{rend.data(), rcode.data(), (uint32_t)rend.size(), 0};
        dfl_parse_fresh(st); lf[256] = 1;
        if (dfl_parse(dx, fv, cfg, st, lf.data(), 1, df.data(), 1, *tr, dfl_jx0(lx), &rec) == 0)
            dfl_parse(dx, fv, cfg, st, lf.data(), 1, df.data(), 1, *tr, 0xffffffffu, &rec);
        result = (int64_t)(st.bits >> 3);
    } else {
        const uint32_t ly = (uint32_t)ly_;
        // checkpoint of x: parse x alone up to its junction
        dfl_parse_fresh(st); lf[256] = 1;
        dfl_parse(dx, fv, cfg, st, lf.data(), 1, df.data(), 1, *tr, dfl_jx0(lx));
        // y alone
        DflStream dy; dy.s.x = py; dy.s.lx = ly; dy.s.y = py + ly; dy.s.n = ly; dy.pair = false;
        dy.ix.order = iy.order.data(); dy.ix.bstart = iy.bstart.data(); dy.iy = dy.ix;
        EmuSeq ey = emu_F_single(dy, cfg, want_q);
        std::vector<uint32_t> &Fy = ey.F;
        EmuCanon ec;
        if (use_canon) {
            const uint32_t cap = (ly / 4 + 1024 + DFL_CUM_G - 1) / DFL_CUM_G * DFL_CUM_G;
            ec.end.assign(cap + 16, 0); ec.code.assign(cap + 16, 0);
            DflRec rec{ec.end.data(), ec.code.data(), cap, 0};
            DflFView fy; fy.fx = fy.fy = fy.fj = Fy.data(); fy.jx0 = fy.jend = fy.lx = ly;
            fy.qx = fy.qy = fy.qj = want_q ? ey.FQ.data() : nullptr;
            DflParseState sy; dfl_parse_fresh(sy);
            std::vector<uint16_t> l2(DFL_L_CODES, 0), d2(DFL_D_CODES, 0); l2[256] = 1;
            DflTrees *t2 = new DflTrees();
            if (dfl_parse(dy, fy, cfg, sy, l2.data(), 1, d2.data(), 1, *t2, dfl_jx0(ly), &rec) == 0)
                dfl_parse(dy, fy, cfg, sy, l2.data(), 1, d2.data(), 1, *t2, 0xffffffffu, &rec);
            delete t2;
            ec.n_sym = rec.n == DFL_NONE ? 0 : rec.n;
            const uint32_t rows = ec.n_sym / DFL_CUM_G;
            ec.cum.assign((size_t)(cap / DFL_CUM_G + 1) * DFL_CUM_W, 0);
            for (uint32_t r = 1; r <= rows; ++r) {                       // dfl_cum_chunk_kernel + dfl_cum_scan_kernel
                uint32_t *row = ec.cum.data() + (size_t)r * DFL_CUM_W;
                const uint32_t *prev = row - DFL_CUM_W;
                for (uint32_t c = 0; c < DFL_CUM_W; ++c) row[c] = prev[c];
                for (uint32_t k = (r - 1) * DFL_CUM_G; k < r * DFL_CUM_G; ++k) {
                    const uint32_t cd = ec.code[k];
                    row[cd & 511]++;
                    if ((cd >> 9) != DFL_LIT) row[DFL_L_CODES + (cd >> 9)]++;
                }
            }
        }
        // pair stream: junction F, then resume
        DflStream d; d.s.x = px; d.s.lx = lx; d.s.y = py; d.s.n = lx + ly; d.pair = true; d.ix = dx.ix; d.iy = dy.ix;
        const uint32_t jx0 = dfl_jx0(lx), jlen = dfl_jlen(lx, ly);
        std::vector<uint32_t> FJ(jlen + 16, 0), FJQ(jlen + 16, 0);
        // dfl_junction_kernel: general walk for the last positions of x, continued walk for the head of y
        const uint32_t jxl = lx - jx0, jyl = jlen - jxl, n_head = (uint32_t)ey.head_order.size();
        int32_t mismatches = 0;
        for (uint32_t u = 0; u < jxl; ++u) FJ[u] = dfl_f_word(d, jx0 + u, cfg, 0xffffffffu, &FJQ[u]);
        // what dfl_head_kernel / dfl_tail6_kernel keep about the tail of x
        std::vector<uint16_t> tail_cnt(DFL_HASH, 0), t6_order(DFL_T6, 0), t6_start(DFL_H6 + 1, 0);
        {
            const uint32_t from = dfl_tail3_from(lx);
            for (uint32_t h = 0; h < DFL_HASH; ++h)
                for (uint32_t k = ix.bstart[h]; k < ix.bstart[h + 1]; ++k) if (ix.order[k] >= from) tail_cnt[h]++;
            const uint32_t t0 = dfl_tail6_t0(lx), n6 = dfl_tail6_count(lx);
            std::vector<uint32_t> cnt(DFL_H6 + 1, 0);
            for (uint32_t i = 0; i < n6; ++i) cnt[dfl_hash6(ldu64(px + t0 + i)) + 1]++;
            for (uint32_t h = 0; h < DFL_H6; ++h) cnt[h + 1] += cnt[h];
            for (uint32_t h = 0; h <= DFL_H6; ++h) t6_start[h] = (uint16_t)cnt[h];
            for (uint32_t i = 0; i < n6; ++i) t6_order[cnt[dfl_hash6(ldu64(px + t0 + i))]++] = (uint16_t)i;
        }
        const DflTail6 t6{t6_order.data(), t6_start.data(), dfl_tail6_t0(lx)};
        int32_t shortcuts = 0;
        for (uint32_t t = 0; t < jyl; ++t) {
            if (t >= n_head) { FJ[jxl + t] = 0; FJQ[jxl + t] = 0; continue; }
            const uint32_t yq = ey.head_order[t];
            uint32_t q = 0;
            const uint32_t w = dfl_junction_word(d, lx + yq, cfg, Fy[yq], ey.head_visit[yq], want_q, want_q ? ey.FQ[yq] : 0u, &t6,
                                                 tail_cnt.data(), &q);
            {   // how often the 6-byte shortcut applies (reported, not checked)
                const uint32_t v = ey.head_visit[yq], cnt = v & DFL_V_COUNT;
                if (!(v & (DFL_V_NICE | DFL_V_HEADFAR)) && (((Fy[yq] & ~DFL_QDIFF) >> 16) >= 5 || (w >> 16) >= 6) && d.s.n - (lx + yq) >= 6 &&
                    cnt + 3 + tail_cnt[dfl_hash_at(d.s, lx + yq)] < ((uint32_t)cfg.max_chain >> 2)) ++shortcuts;
            }
            FJ[jxl + yq] = w; FJQ[jxl + yq] = q;
            // self-check against the general walk over the pair stream
            uint32_t q0 = 0;
            const uint32_t w0 = dfl_f_word(d, lx + yq, cfg, 0xffffffffu, &q0);
            if ((w & ~DFL_QDIFF) != (w0 & ~DFL_QDIFF) || ((w0 & DFL_QDIFF) && !(w & DFL_QDIFF)) || (want_q && q != q0)) ++mismatches;
        }
        if (info) { info[2] = mismatches + emu_match_mismatches + (int32_t)dfl_compact_mismatch; info[3] = shortcuts; }
        fv.fx = Fx.data(); fv.fy = Fy.data(); fv.fj = FJ.data(); fv.jx0 = jx0; fv.jend = jx0 + jlen; fv.lx = lx;
        if (want_q) { fv.qx = ex.FQ.data(); fv.qy = ey.FQ.data(); fv.qj = FJQ.data(); }
        dfl_resume(st, d.s.n);
        const DflParseState st0 = st;
        const std::vector<uint16_t> lf0 = lf, df0 = df;
        DflCanon cn; cn.end = ec.end.data(); cn.code = ec.code.data(); cn.cum = ec.cum.data(); cn.n_sym = ec.n_sym;
        std::vector<uint32_t> accA(DFL_CUM_W), accB(DFL_CUM_W);
        const uint32_t strstart0 = st.strstart;
        DflCompactTrees *ct = new DflCompactTrees();
        const int how = dfl_pair_stream(d, fv, cfg, st, lf.data(), 1, df.data(), 1, *tr, use_canon ? &cn : nullptr, accA.data(), accB.data(), ct);
        delete ct;
        (void)strstart0;
        if (!how) {
            if (info) info[1] = 1;
            st = st0; lf = lf0; df = df0;
            dfl_pair_stream(d, fv, cfg, st, lf.data(), 1, df.data(), 1, *tr, (const DflCanon *)nullptr, accA.data(), accB.data());
        } else if (info) {
            info[0] = how == 2;
        }
        result = (int64_t)(st.bits >> 3);
    }
    delete tr; free(px); free(py);
    return result;
}

// The sequence alone through the chunked path (dfl_chunk_*_kernel + dfl_alone_kernel): scan / sync / count / emit
// with chunks of DFL_CHUNK bytes, cumulative rows, dfl_canon_blocks from symbol 0, serial tail.  Returns the raw
// deflate size, -1 when the stitched symbol stream differs from the serial parse's, -2 when the sizes differ, -3 when
// the sequence is too short for the path; info[0] = 1 when the chunked path stood (0: it gave the sequence up --
// chunks that never meet, or a block zlib might store), info[1] = chunks.
extern "C" int64_t emu_deflate_chunked(const uint8_t *x, uint32_t lx, int level, int32_t *info)
{
    const DflConfig cfg = dfl_config(level);
    info[0] = info[1] = 0;
    if (lx < DFL_PAR_MIN) return -3;
    uint8_t *px = padded_copy(x, lx);
    EmuIndex ix = emu_index(px, lx);
    DflStream dx; dx.s.x = px; dx.s.lx = lx; dx.s.y = px + lx; dx.s.n = lx; dx.pair = false;
    dx.ix.order = ix.order.data(); dx.ix.bstart = ix.bstart.data(); dx.iy = dx.ix;
    const bool want_q = level != 9;
    EmuSeq ex = emu_F_single(dx, cfg, want_q);
    const uint32_t *F = ex.F.data(), *Q = want_q ? ex.FQ.data() : nullptr;
    DflFView fv; fv.fx = fv.fy = fv.fj = F; fv.jx0 = fv.jend = fv.lx = lx; fv.qx = fv.qy = fv.qj = Q;
    const uint32_t cap = (lx / 4 + 1024 + DFL_CUM_G - 1) / DFL_CUM_G * DFL_CUM_G;
    // serial reference
    std::vector<uint32_t> rend(cap + 16); std::vector<uint16_t> rcode(cap + 16);
    DflRec rec{rend.data(), rcode.data(), cap, 0};
    DflTrees *tr = new DflTrees();
    std::vector<uint16_t> lf(DFL_L_CODES, 0), df(DFL_D_CODES, 0);
    DflParseState st; dfl_parse_fresh(st); lf[256] = 1;
    dfl_parse(dx, fv, cfg, st, lf.data(), 1, df.data(), 1, *tr, 0xffffffffu, &rec);
    const int64_t serial_size = (int64_t)(st.bits >> 3);
    int64_t result = serial_size;
    // chunked
    const uint32_t K = dfl_chunk_count(lx), E = dfl_chunk_end(lx);
    info[1] = (int32_t)K;
    std::vector<uint32_t> head((size_t)K * DFL_CHUNK_WORDS, 0), tail((size_t)K * DFL_CHUNK_WORDS, 0), ys(K, 0), cnt(K, 0), off(K, 0);
    for (uint32_t k = 0; k < K; ++k) {
        DflLite o; o.mode = 0; o.until = DFL_NONE; o.count = 0; o.end = nullptr; o.code = nullptr; o.cap = 0;
        o.head = head.data() + (size_t)k * DFL_CHUNK_WORDS; o.tail = tail.data() + (size_t)k * DFL_CHUNK_WORDS;
        const uint32_t s_k = k * DFL_CHUNK, s_next = s_k + DFL_CHUNK;
        const bool last = k + 1 == K;
        o.head_lo = k ? s_k : DFL_NOWIN; o.tail_lo = last ? DFL_NOWIN : s_next;
        const uint32_t stop = last ? (k ? tmin(s_k + DFL_CHUNK_OV, E) : 0u) : tmin(s_next + DFL_CHUNK_OV, E);
        dfl_lite_parse(dx, F, Q, cfg, s_k, stop, o);
    }
    bool ok = true;
    for (uint32_t k = 0; k < K; ++k) {
        ys[k] = k ? dfl_chunk_sync(tail.data() + (size_t)(k - 1) * DFL_CHUNK_WORDS, head.data() + (size_t)k * DFL_CHUNK_WORDS, k * DFL_CHUNK) : 0u;
        if (ys[k] == DFL_NONE) ok = false;
    }
    std::vector<uint32_t> cend(cap + 16, 0); std::vector<uint16_t> ccode(cap + 16, 0);
    uint32_t total = 0;
    if (ok) {
        for (int mode = 1; mode <= 2 && ok; ++mode) {
            uint32_t acc = 0;
            for (uint32_t k = 0; k < K; ++k) {
                DflLite o; o.mode = mode; o.until = k + 1 < K ? ys[k + 1] : DFL_NONE; o.count = 0;
                o.head = o.tail = nullptr; o.head_lo = o.tail_lo = DFL_NOWIN;
                o.end = cend.data() + off[k]; o.code = ccode.data() + off[k]; o.cap = mode == 2 ? cnt[k] : 0;
                if (!dfl_lite_parse(dx, F, Q, cfg, ys[k], E, o)) ok = false;
                if (mode == 1) { cnt[k] = o.count; off[k] = acc; acc += o.count; }
                else if (o.count != cnt[k]) ok = false;
            }
            if (mode == 1) { total = acc; if (total > cap) ok = false; }
        }
    }
    if (ok) {
        // the stitched stream is the serial one, symbol for symbol
        if (total > rec.n) result = -1;
        for (uint32_t k = 0; k < total && result >= 0; ++k) if (cend[k] != rend[k] || ccode[k] != rcode[k]) result = -1;
        // cumulative rows, then blocks from symbol 0 and the serial tail (dfl_alone_kernel)
        std::vector<uint32_t> cum((size_t)(cap / DFL_CUM_G + 1) * DFL_CUM_W, 0);
        for (uint32_t r = 1; r <= total / DFL_CUM_G; ++r) {
            uint32_t *row = cum.data() + (size_t)r * DFL_CUM_W;
            for (uint32_t c = 0; c < DFL_CUM_W; ++c) row[c] = (row - DFL_CUM_W)[c];
            for (uint32_t k = (r - 1) * DFL_CUM_G; k < r * DFL_CUM_G; ++k) {
                const uint32_t cd = ccode[k];
                row[cd & 511]++;
                if ((cd >> 9) != DFL_LIT) row[DFL_L_CODES + (cd >> 9)]++;
            }
        }
        DflCanon cn{cend.data(), ccode.data(), cum.data(), total};
        std::vector<uint32_t> accA(DFL_L_CODES + DFL_D_CODES), accB(DFL_L_CODES + DFL_D_CODES);
        std::fill(lf.begin(), lf.end(), 0); std::fill(df.begin(), df.end(), 0); lf[256] = 1;
        dfl_parse_fresh(st);
        uint32_t t_end = 0;
        const int b = dfl_canon_blocks(cn, 0, lx, DFL_NONE, st, lf.data(), 1, df.data(), 1, *tr, accA.data(), accB.data(),
                                       (DflCompactTrees *)nullptr, &t_end);
        if (b > 0) {
            DflRec rec2{cend.data(), ccode.data(), cap, t_end};
            if (dfl_parse(dx, fv, cfg, st, lf.data(), 1, df.data(), 1, *tr, dfl_jx0(lx), &rec2) == 0)
                dfl_parse(dx, fv, cfg, st, lf.data(), 1, df.data(), 1, *tr, 0xffffffffu, &rec2);
            if (result >= 0 && (int64_t)(st.bits >> 3) != serial_size) result = -2;
            if (result >= 0 && rec2.n != rec.n) result = -1;
            for (uint32_t k = 0; k < rec.n && result >= 0; ++k) if (cend[k] != rend[k] || ccode[k] != rcode[k]) result = -1;
            info[0] = 1;
        }
    }
    delete tr;
    free(px);
    return result;
}

extern "C" int64_t emu_deflate_size(const uint8_t *x, uint32_t lx, const uint8_t *y, int64_t ly_, int level)
{
    return emu_deflate_size_ex(x, lx, y, ly_, level, 1, nullptr);
}

// how many block costs the compacted trees have computed (and checked against the full ones) so far
extern "C" int64_t emu_compact_checked() { return dfl_compact_checked; }

// random symbol histograms: dfl_flush_block_compact against dfl_flush_block (bits, stored-candidate flag, last-block
// alignment).  Returns the number of disagreements; *applied = how many cases the compact version accepted.
extern "C" int64_t emu_flush_compact_fuzz(uint32_t seed, int32_t cases, int32_t *applied)
{
    uint64_t s = seed * 0x9E3779B97F4A7C15ull + 12345;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 11); };
    DflTrees *tr = new DflTrees();
    DflCompactTrees *ct = new DflCompactTrees();
    int64_t bad = 0; int32_t ok = 0;
    for (int32_t c = 0; c < cases; ++c) {
        uint16_t lf[DFL_L_CODES] = {0}, df[DFL_D_CODES] = {0};
        const int kl = 1 + (int)(rnd() % 60), kd = (int)(rnd() % 31), shape = (int)(rnd() % 4);
        uint32_t budget = 16383;
        auto draw = [&]() -> uint32_t {                     // skewed frequencies: ties, powers of two, long tails
            uint32_t f;
            switch (shape) {
                case 0: f = 1 + rnd() % 8; break;
                case 1: f = 1u << (rnd() % 12); break;
                case 2: f = 1 + rnd() % 2000; break;
                default: f = (rnd() % 4 == 0) ? 1 + rnd() % 4000 : 1 + rnd() % 3;
            }
            f = f > budget ? budget : f;
            budget -= f;
            return f;
        };
        for (int k = 0; k < kl && budget; ++k) lf[shape == 3 && k < 4 ? "ACGT"[k] : rnd() % DFL_L_CODES] += (uint16_t)draw();
        for (int k = 0; k < kd && budget; ++k) df[rnd() % DFL_D_CODES] += (uint16_t)draw();
        lf[256] = 1;
        const uint64_t stored = rnd() % 200000;
        const bool last = rnd() & 1, can = rnd() & 1;
        uint64_t b0 = rnd() % 64, b1 = b0; bool c0 = false, c1 = false;
        dfl_flush_block(*tr, lf, 1, df, 1, stored, can, last, b0, &c0);
        if (dfl_flush_block_compact(*ct, lf, 1, df, 1, stored, can, last, b1, &c1)) { ++ok; if (b0 != b1 || c0 != c1) ++bad; }
    }
    if (applied) *applied = ok;
    delete tr; delete ct;
    return bad;
}

// step counters of the packed LZ4 parse since the library was loaded: [0] general steps (pk_step), [1] steps inside the
// speculative loop (pk_turbo)
extern "C" void emu_lz4_step_counts(uint64_t *out) { out[0] = pk_general_steps; out[1] = pk_turbo_steps; out[2] = pk_batch_steps; }
