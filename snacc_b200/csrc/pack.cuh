// pack.cuh -- corpus alphabet detection and 2-bit packing (K0b of DESIGN.md), sm_100a.
//
// The reference hands the compressors one byte per base (pairwise_ncd.py:69).  Genomes use four byte
// values almost everywhere, so after upload the corpus alphabet is measured on the device (byte
// histogram), the four most frequent byte values become the codes 0..3, and every sequence made only of
// those bytes is additionally stored at 2 bits per base (base k in bits 2(k%32) of 64-bit word k/32).
// The packed LZ4 kernels (lz4_packed.cuh) run on that copy; sequences with any other byte keep to the
// byte-wise kernels.  Compressed sizes are those of the ORIGINAL bytes: the k-mer -> hash-bucket maps
// below are computed from the real byte values.
#pragma once
#include "common.cuh"
#include <vector>
#include <algorithm>

namespace snacc {

// ---- host: alphabet and code -> slot maps -----------------------------------------------------------
struct PkAlphabet {
    uint8_t sym[4];          // byte value of code 0..3 (ascending)
    uint8_t code_of[256];    // 0xff = not in the alphabet
};

inline PkAlphabet pk_choose_alphabet(const unsigned long long *hist)
{
    std::vector<int> order(256);
    for (int i = 0; i < 256; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return hist[a] > hist[b]; });
    std::vector<int> top(order.begin(), order.begin() + 4);
    std::sort(top.begin(), top.end());
    PkAlphabet a;
    for (int i = 0; i < 256; ++i) a.code_of[i] = 0xff;
    for (int i = 0; i < 4; ++i) { a.sym[i] = (uint8_t)top[i]; a.code_of[top[i]] = (uint8_t)i; }
    return a;
}

// LZ4 1.9.4 hash functions on the real bytes of a k-mer code (lz4.cuh: Lz4Table<>::hash)
inline uint32_t pk_bucket(const PkAlphabet &a, uint32_t code, bool u16)
{
    uint64_t seq = 0;
    const int k = u16 ? 4 : 5;
    for (int i = 0; i < k; ++i) seq |= (uint64_t)a.sym[(code >> (2 * i)) & 3] << (8 * i);
    if (u16) return ((uint32_t)seq * 2654435761u) >> (32 - 13);
    return (uint32_t)(((seq << 24) * 889523592379ull) >> (64 - 12));
}

// lut[c] = table slot of k-mer code c: one slot per distinct bucket, numbered in order of first use
inline uint32_t pk_slot_lut(const PkAlphabet &a, bool u16, uint16_t *lut)
{
    const uint32_t n = u16 ? 256 : 1024;
    std::vector<uint32_t> bucket;
    for (uint32_t c = 0; c < n; ++c) {
        const uint32_t b = pk_bucket(a, c, u16);
        size_t k = 0;
        while (k < bucket.size() && bucket[k] != b) ++k;
        if (k == bucket.size()) bucket.push_back(b);
        lut[c] = (uint16_t)k;
    }
    return (uint32_t)bucket.size();
}

#ifdef __CUDACC__
// ---- device -----------------------------------------------------------------------------------
__global__ void pk_hist_kernel(const uint8_t *__restrict__ corpus, const uint64_t *__restrict__ off,
                               const uint32_t *__restrict__ len, int32_t n_seqs, unsigned long long *__restrict__ hist)
{
    __shared__ unsigned int h[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) h[i] = 0;
    __syncthreads();
    // A genome uses four byte values almost everywhere, so counting straight into shared memory would serialise every
    // warp on four addresses.  Each thread keeps four (value, count) pairs in registers -- the first four distinct
    // values it meets -- and only the bytes that match none of them go to the shared counters.
    uint32_t cv0 = 0x100, cv1 = 0x100, cv2 = 0x100, cv3 = 0x100, cn0 = 0, cn1 = 0, cn2 = 0, cn3 = 0;
    // work item = (sequence, 64 KiB slice)
    for (int32_t s = blockIdx.y; s < n_seqs; s += gridDim.y) {
        const uint8_t *p = corpus + off[s];
        const uint32_t l = len[s];
        for (uint32_t base = blockIdx.x * 65536u; base < l; base += gridDim.x * 65536u) {
            const uint32_t end = tmin(l, base + 65536u);
            for (uint32_t i = base + threadIdx.x * 16; i < end; i += blockDim.x * 16) {
                const uint4 q = __ldg(reinterpret_cast<const uint4 *>(p + i));
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
                const uint32_t cnt = tmin(16u, end - i);
#pragma unroll
                for (uint32_t b = 0; b < 16; ++b) {
                    if (b >= cnt) break;
                    const uint32_t v = (w[b >> 2] >> (8 * (b & 3))) & 0xff;
                    if (v == cv0) ++cn0;
                    else if (v == cv1) ++cn1;
                    else if (v == cv2) ++cn2;
                    else if (v == cv3) ++cn3;
                    else if (cv0 == 0x100) { cv0 = v; cn0 = 1; }
                    else if (cv1 == 0x100) { cv1 = v; cn1 = 1; }
                    else if (cv2 == 0x100) { cv2 = v; cn2 = 1; }
                    else if (cv3 == 0x100) { cv3 = v; cn3 = 1; }
                    else atomicAdd(&h[v], 1u);
                }
            }
        }
    }
    if (cn0) atomicAdd(&h[cv0], cn0);
    if (cn1) atomicAdd(&h[cv1], cn1);
    if (cn2) atomicAdd(&h[cv2], cn2);
    if (cn3) atomicAdd(&h[cv3], cn3);
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (h[i]) atomicAdd(&hist[i], (unsigned long long)h[i]);
}

struct PkCodes { uint8_t c[256]; };   // byte -> code (0xff: outside the alphabet), passed by value

// one thread per packed word; nexc[s] counts the bytes of sequence s outside the alphabet.  Such a byte gets the filler
// code 0 and -- when `mask` is given (corpora that have any) -- its bit in the per-base mask (word j of a sequence's
// mask covers the same 32 bases as its packed word j; the mask is zeroed beforehand).
__global__ void pk_pack_kernel(const uint8_t *__restrict__ corpus, const uint64_t *__restrict__ off,
                               const uint32_t *__restrict__ len, const uint64_t *__restrict__ woff, int32_t n_seqs,
                               uint64_t *__restrict__ words, int32_t *__restrict__ nexc,
                               uint32_t *__restrict__ mask, const uint64_t *__restrict__ moff, const PkCodes codes)
{
    for (int32_t s = blockIdx.y; s < n_seqs; s += gridDim.y) {
        const uint8_t *p = corpus + off[s];
        const uint32_t l = len[s];
        const uint32_t nw = pk_words(l);
        uint64_t *dst = words + woff[s];
        int32_t n_bad = 0;
        for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nw; j += gridDim.x * blockDim.x) {
            uint64_t w = 0;
            uint32_t mw = 0;
            const uint32_t b0 = j * 32;
            if (b0 < l) {
                const uint4 q0 = __ldg(reinterpret_cast<const uint4 *>(p + b0));
                const uint4 q1 = __ldg(reinterpret_cast<const uint4 *>(p + b0 + 16));
                const uint32_t v[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                const uint32_t cnt = tmin(32u, l - b0);
                for (uint32_t b = 0; b < cnt; ++b) {
                    const uint32_t code = codes.c[(v[b >> 2] >> (8 * (b & 3))) & 0xff];
                    if (code > 3) mw |= 1u << b;
                    else w |= (uint64_t)code << (2 * b);
                }
            }
            dst[j] = w;
            if (mw) {
                n_bad += __popc(mw);
                if (mask) mask[moff[s] + j] = mw;
            }
        }
        if (n_bad) atomicAdd(&nexc[s], n_bad);
    }
}
#endif

}  // namespace snacc
