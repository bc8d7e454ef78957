// deflate.cuh -- raw-deflate compressed-size kernels (K3 of SURVEY.md 2.3), sm_100a.  [stub: filled next]
#pragma once
#include "common.cuh"
#include <string>

namespace snacc {

struct DeflateCorpus {
    const uint8_t *d_corpus; const uint64_t *d_off; const uint32_t *d_len;
    const uint64_t *h_off; const uint32_t *h_len; int32_t n_seqs;
};
struct DeflateState { int dummy = 0; };

static inline void deflate_free_corpus(DeflateState &) {}
static inline void deflate_free_work(DeflateState &) {}
static inline void deflate_invalidate(DeflateState &) {}
static inline int deflate_run(DeflateState &, const DeflateCorpus &, int, const int32_t *, const int32_t *,
                              const int32_t *, const int32_t *, int64_t, int64_t *, cudaStream_t, int64_t,
                              int64_t *, std::string &err)
{
    err = "deflate codecs are not built into this library yet";
    return -4;
}

}  // namespace snacc
