// api.cu -- C ABI of libsnacc_b200.so (see include/snacc_b200.h) and the host-side orchestration of
// the sm_100a kernels.  No torch, no C++ types across the boundary, no exceptions escape.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include <mutex>

#include "../../include/snacc_b200.h"
#include "common.cuh"
#include "lz4.cuh"
#include "deflate.cuh"

using namespace snacc;

#define SNACC_VERSION 100   /* 0.1.0 */

struct snacc_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    std::mutex mu;                 // serialises calls: the single-job shim may be hit from many threads

    // corpus (device)
    int32_t n_seqs = 0;
    uint8_t *d_corpus = nullptr;   // padded layout
    uint64_t corpus_bytes = 0;
    uint64_t *d_off = nullptr;     // n_seqs offsets into d_corpus
    uint32_t *d_len = nullptr;     // n_seqs lengths
    std::vector<uint64_t> h_off;
    std::vector<uint32_t> h_len;

    // LZ4 prefix checkpoints
    int32_t *d_slot_of = nullptr;          // per sequence: checkpoint slot or -1
    std::vector<int32_t> h_slot_of;
    std::vector<uint8_t> ckpt_done;        // per sequence
    int32_t n_slots = 0;
    uint8_t *d_ckpt_tab = nullptr;
    uint64_t *d_ckpt_total = nullptr;

    // working memory
    uint8_t *d_work = nullptr; size_t work_bytes = 0;
    unsigned long long *d_counter = nullptr;
    int32_t *d_jobx = nullptr, *d_joby = nullptr; int64_t job_cap = 0;
    int64_t *d_out = nullptr; int64_t out_cap = 0;

    // options / instrumentation
    int64_t streams_in_flight = 0;         // 0 = default
    double last_ms = 0.0; int64_t last_launches = 0;
    double last_main_ms = 0.0;              // dominant kernel only (lz4_stream_kernel / deflate pair kernel)
    cudaEvent_t evm0 = nullptr, evm1 = nullptr;

    DeflateState dfl;
};

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
        char b_[512]; snprintf(b_, sizeof b_, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
        ctx->err = b_; return SNACC_ERR_CUDA; } } while (0)
#define FAIL(code, msg) do { ctx->err = (msg); return (code); } while (0)

// ---------------------------------------------------------------------------------------------
// K0: scatter records into the padded corpus, optionally reverse-complementing each record
// (Biopython ambiguous-DNA table, case preserving; pairwise_ncd.py:33-34)
// ---------------------------------------------------------------------------------------------
__constant__ uint8_t c_complement[256];

static void build_complement_table(uint8_t *t)
{
    for (int i = 0; i < 256; ++i) t[i] = (uint8_t)i;
    const char *a = "ACGTMRWSYKVHDBXNU";
    const char *b = "TGCAKYWSRMBDHVXNA";
    for (int i = 0; a[i]; ++i) {
        t[(uint8_t)a[i]] = (uint8_t)b[i];
        t[(uint8_t)(a[i] | 0x20)] = (uint8_t)(b[i] | 0x20);
    }
}

__global__ void scatter_records_kernel(const uint8_t *__restrict__ src, uint64_t total,
                                       const uint64_t *__restrict__ rec_off, int64_t n_recs,
                                       const uint64_t *__restrict__ rec_dst, int rc,
                                       uint8_t *__restrict__ dst)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        // record containing source byte i: last r with rec_off[r] <= i
        int64_t lo = 0, hi = n_recs;
        while (hi - lo > 1) {
            int64_t mid = (lo + hi) >> 1;
            if (rec_off[mid] <= i) lo = mid; else hi = mid;
        }
        const uint64_t a = rec_off[lo], b = rec_off[lo + 1];
        const uint64_t k = i - a;
        uint8_t v = src[i];
        if (rc) dst[rec_dst[lo] + (b - a - 1 - k)] = c_complement[v];
        else    dst[rec_dst[lo] + k] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// K4: NCD epilogue in float64 (two correctly rounded divides and a min, like the reference)
// ---------------------------------------------------------------------------------------------
__global__ void ncd_kernel(const int64_t *__restrict__ C, const int64_t *__restrict__ S, int32_t n,
                           int formula, int64_t bias, double *__restrict__ D)
{
    const int64_t total = (int64_t)n * n;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total;
         k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t i = (int32_t)(k / n), j = (int32_t)(k % n);
        const int64_t x = C[i] + bias, y = C[j] + bias;
        const int64_t lo = x < y ? x : y, hi = x < y ? y : x;
        const int64_t cxy = S[(int64_t)i * n + j] + bias;
        const double d1 = __ddiv_rn((double)(cxy - lo), (double)hi);
        if (formula == SNACC_NCD_ONE_ORDER) { D[k] = d1; continue; }
        const int64_t cyx = S[(int64_t)j * n + i] + bias;
        const double d2 = __ddiv_rn((double)(cyx - lo), (double)hi);
        D[k] = d1 < d2 ? d1 : d2;
    }
}

// ---------------------------------------------------------------------------------------------
extern "C" int snacc_version(void) { return SNACC_VERSION; }

extern "C" const char *snacc_last_error(const snacc_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

static void free_corpus(snacc_ctx *ctx)
{
    cudaFree(ctx->d_corpus); ctx->d_corpus = nullptr;
    cudaFree(ctx->d_off); ctx->d_off = nullptr;
    cudaFree(ctx->d_len); ctx->d_len = nullptr;
    cudaFree(ctx->d_slot_of); ctx->d_slot_of = nullptr;
    cudaFree(ctx->d_ckpt_tab); ctx->d_ckpt_tab = nullptr;
    cudaFree(ctx->d_ckpt_total); ctx->d_ckpt_total = nullptr;
    ctx->n_seqs = 0; ctx->n_slots = 0;
    deflate_free_corpus(ctx->dfl);
}

extern "C" int snacc_ctx_create(int device_id, snacc_ctx **out)
{
    if (!out) return SNACC_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device_id < 0 || device_id >= n) return SNACC_ERR_CUDA;
    snacc_ctx *ctx = new (std::nothrow) snacc_ctx();
    if (!ctx) return SNACC_ERR_ARG;
    ctx->device = device_id;
    if (cudaSetDevice(device_id) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
        cudaEventCreate(&ctx->evm0) != cudaSuccess || cudaEventCreate(&ctx->evm1) != cudaSuccess ||
        cudaMalloc(&ctx->d_counter, 64) != cudaSuccess) {
        delete ctx;
        return SNACC_ERR_CUDA;
    }
    uint8_t tab[256];
    build_complement_table(tab);
    if (cudaMemcpyToSymbol(c_complement, tab, 256) != cudaSuccess) { delete ctx; return SNACC_ERR_CUDA; }
    *out = ctx;
    return SNACC_OK;
}

extern "C" void snacc_ctx_destroy(snacc_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_corpus(ctx);
    deflate_free_work(ctx->dfl);
    cudaFree(ctx->d_work); cudaFree(ctx->d_counter);
    cudaFree(ctx->d_jobx); cudaFree(ctx->d_joby); cudaFree(ctx->d_out);
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1);
    cudaEventDestroy(ctx->evm0); cudaEventDestroy(ctx->evm1);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

static int upload_impl(snacc_ctx *ctx, const void *bytes, bool on_device, const uint64_t *seq_offsets,
                       int32_t n_seqs, const uint64_t *rec_offsets, int64_t n_recs, int rc)
{
    if (!ctx) return SNACC_ERR_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->err.clear();
    if (!seq_offsets || n_seqs <= 0) FAIL(SNACC_ERR_ARG, "snacc_upload: no sequences");
    if (rc && (!rec_offsets || n_recs <= 0)) FAIL(SNACC_ERR_ARG, "snacc_upload: reverse complement needs record offsets");
    const uint64_t total = seq_offsets[n_seqs] - seq_offsets[0];
    if (!bytes && total) FAIL(SNACC_ERR_ARG, "snacc_upload: null bytes");
    for (int32_t i = 0; i < n_seqs; ++i) {
        if (seq_offsets[i + 1] < seq_offsets[i]) FAIL(SNACC_ERR_ARG, "snacc_upload: offsets not monotone");
        const uint64_t l = seq_offsets[i + 1] - seq_offsets[i];
        if (l == 0) {
            char b[128]; snprintf(b, sizeof b, "sequence %d is empty (no sequence extracted)", i);
            FAIL(SNACC_ERR_EMPTY, b);
        }
        if (l >= 0x7fffffffull) FAIL(SNACC_ERR_TOO_LARGE, "snacc_upload: sequence of 2 GiB or more");
    }
    CK(cudaSetDevice(ctx->device));
    free_corpus(ctx);

    // records default to whole sequences
    std::vector<uint64_t> recs;
    if (!rec_offsets) { recs.assign(seq_offsets, seq_offsets + n_seqs + 1); n_recs = n_seqs; }
    else recs.assign(rec_offsets, rec_offsets + n_recs + 1);
    if (recs.front() != seq_offsets[0] || recs.back() != seq_offsets[n_seqs])
        FAIL(SNACC_ERR_ARG, "snacc_upload: record offsets do not cover the sequences");

    ctx->h_off.resize(n_seqs); ctx->h_len.resize(n_seqs);
    uint64_t pos = 0;
    for (int32_t i = 0; i < n_seqs; ++i) {
        ctx->h_off[i] = pos;
        ctx->h_len[i] = (uint32_t)(seq_offsets[i + 1] - seq_offsets[i]);
        pos += ((uint64_t)ctx->h_len[i] + SEQ_PAD + SEQ_ALIGN - 1) & ~(uint64_t)(SEQ_ALIGN - 1);
    }
    ctx->corpus_bytes = pos + 64;
    // destination of every record
    std::vector<uint64_t> rec_dst((size_t)n_recs);
    {
        int32_t si = 0;
        for (int64_t r = 0; r < n_recs; ++r) {
            if (recs[r + 1] < recs[r]) FAIL(SNACC_ERR_ARG, "snacc_upload: record offsets not monotone");
            if (recs[r + 1] == recs[r]) { rec_dst[r] = 0; continue; }   // empty record: moves no bytes
            while (recs[r] >= seq_offsets[si + 1]) ++si;
            if (recs[r + 1] > seq_offsets[si + 1])
                FAIL(SNACC_ERR_ARG, "snacc_upload: a record straddles two sequences");
            rec_dst[r] = ctx->h_off[si] + (recs[r] - seq_offsets[si]);
        }
    }
    for (auto &v : recs) v -= seq_offsets[0];

    CK(cudaMalloc(&ctx->d_corpus, ctx->corpus_bytes));
    CK(cudaMemsetAsync(ctx->d_corpus, 0, ctx->corpus_bytes, ctx->stream));
    CK(cudaMalloc(&ctx->d_off, sizeof(uint64_t) * n_seqs));
    CK(cudaMalloc(&ctx->d_len, sizeof(uint32_t) * n_seqs));
    CK(cudaMemcpyAsync(ctx->d_off, ctx->h_off.data(), sizeof(uint64_t) * n_seqs, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_len, ctx->h_len.data(), sizeof(uint32_t) * n_seqs, cudaMemcpyHostToDevice, ctx->stream));

    uint8_t *d_src = nullptr; uint64_t *d_rec_off = nullptr, *d_rec_dst = nullptr;
    const uint8_t *src_dev = nullptr;
    if (on_device) src_dev = (const uint8_t *)bytes + 0;
    else {
        CK(cudaMalloc(&d_src, total ? total : 1));
        CK(cudaMemcpyAsync(d_src, (const uint8_t *)bytes + seq_offsets[0], total, cudaMemcpyHostToDevice, ctx->stream));
        src_dev = d_src;
    }
    if (on_device) src_dev += seq_offsets[0];
    CK(cudaMalloc(&d_rec_off, sizeof(uint64_t) * (n_recs + 1)));
    CK(cudaMalloc(&d_rec_dst, sizeof(uint64_t) * n_recs));
    CK(cudaMemcpyAsync(d_rec_off, recs.data(), sizeof(uint64_t) * (n_recs + 1), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_rec_dst, rec_dst.data(), sizeof(uint64_t) * n_recs, cudaMemcpyHostToDevice, ctx->stream));
    {
        const int threads = 256;
        const uint64_t want = (total + threads - 1) / threads;
        const int blocks = (int)std::min<uint64_t>(std::max<uint64_t>(want, 1), 148ull * 16);
        scatter_records_kernel<<<blocks, threads, 0, ctx->stream>>>(src_dev, total, d_rec_off, n_recs, d_rec_dst,
                                                                    rc ? 1 : 0, ctx->d_corpus);
        CK(cudaGetLastError());
    }

    // LZ4 checkpoint slots for sequences that own at least one full 64 KiB block
    ctx->h_slot_of.assign(n_seqs, -1);
    ctx->ckpt_done.assign(n_seqs, 0);
    ctx->n_slots = 0;
    for (int32_t i = 0; i < n_seqs; ++i)
        if (ctx->h_len[i] >= LZ4_BLOCK) ctx->h_slot_of[i] = ctx->n_slots++;
    CK(cudaMalloc(&ctx->d_slot_of, sizeof(int32_t) * n_seqs));
    CK(cudaMemcpyAsync(ctx->d_slot_of, ctx->h_slot_of.data(), sizeof(int32_t) * n_seqs, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->n_slots) {
        CK(cudaMalloc(&ctx->d_ckpt_tab, (size_t)ctx->n_slots * LZ4_TABLE_BYTES));
        CK(cudaMalloc(&ctx->d_ckpt_total, sizeof(uint64_t) * ctx->n_slots));
    }
    ctx->n_seqs = n_seqs;
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_src); cudaFree(d_rec_off); cudaFree(d_rec_dst);
    return SNACC_OK;
}

extern "C" int snacc_upload(snacc_ctx *ctx, const uint8_t *bytes, const uint64_t *seq_offsets, int32_t n_seqs,
                            const uint64_t *rec_offsets, int64_t n_recs, int reverse_complement)
{
    return upload_impl(ctx, bytes, false, seq_offsets, n_seqs, rec_offsets, n_recs, reverse_complement);
}

extern "C" int snacc_upload_device(snacc_ctx *ctx, const void *d_bytes, const uint64_t *seq_offsets, int32_t n_seqs,
                                   const uint64_t *rec_offsets, int64_t n_recs, int reverse_complement)
{
    return upload_impl(ctx, d_bytes, true, seq_offsets, n_seqs, rec_offsets, n_recs, reverse_complement);
}

extern "C" int snacc_download_sequence(snacc_ctx *ctx, int32_t i, uint8_t *out)
{
    if (!ctx || !out) return SNACC_ERR_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!ctx->n_seqs) FAIL(SNACC_ERR_STATE, "nothing uploaded");
    if (i < 0 || i >= ctx->n_seqs) FAIL(SNACC_ERR_ARG, "sequence index out of range");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(out, ctx->d_corpus + ctx->h_off[i], ctx->h_len[i], cudaMemcpyDeviceToHost));
    return SNACC_OK;
}

// ---------------------------------------------------------------------------------------------
static int ensure_job_buffers(snacc_ctx *ctx, int64_t n)
{
    if (n > ctx->job_cap) {
        cudaFree(ctx->d_jobx); cudaFree(ctx->d_joby); ctx->d_jobx = ctx->d_joby = nullptr; ctx->job_cap = 0;
        CK(cudaMalloc(&ctx->d_jobx, sizeof(int32_t) * n));
        CK(cudaMalloc(&ctx->d_joby, sizeof(int32_t) * n));
        ctx->job_cap = n;
    }
    if (n > ctx->out_cap) {
        cudaFree(ctx->d_out); ctx->d_out = nullptr; ctx->out_cap = 0;
        CK(cudaMalloc(&ctx->d_out, sizeof(int64_t) * n));
        ctx->out_cap = n;
    }
    return SNACC_OK;
}

static int ensure_work(snacc_ctx *ctx, size_t bytes)
{
    if (bytes > ctx->work_bytes) {
        cudaFree(ctx->d_work); ctx->d_work = nullptr; ctx->work_bytes = 0;
        CK(cudaMalloc(&ctx->d_work, bytes));
        ctx->work_bytes = bytes;
    }
    return SNACC_OK;
}

// LZ4: make sure every x that starts a linked-regime stream has its prefix checkpoint
static int lz4_prepare_prefixes(snacc_ctx *ctx, const int32_t *xs, int64_t n)
{
    std::vector<int32_t> todo;
    for (int64_t k = 0; k < n; ++k) {
        const int32_t x = xs[k];
        if (ctx->h_slot_of[x] >= 0 && !ctx->ckpt_done[x]) { ctx->ckpt_done[x] = 1; todo.push_back(x); }
    }
    if (todo.empty()) return SNACC_OK;
    int32_t *d_todo = nullptr;
    CK(cudaMalloc(&d_todo, sizeof(int32_t) * todo.size()));
    CK(cudaMemcpyAsync(d_todo, todo.data(), sizeof(int32_t) * todo.size(), cudaMemcpyHostToDevice, ctx->stream));
    const int threads = 32;
    const int blocks = (int)((todo.size() + threads - 1) / threads);
    lz4_prefix_kernel<<<blocks, threads, 0, ctx->stream>>>(ctx->d_corpus, ctx->d_off, ctx->d_len, d_todo,
                                                           (int32_t)todo.size(), ctx->d_slot_of, ctx->d_ckpt_tab,
                                                           ctx->d_ckpt_total);
    ctx->last_launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_todo);
    return SNACC_OK;
}

static int run_lz4(snacc_ctx *ctx, const int32_t *h_x, int64_t n_jobs, bool pairs)
{
    int r = lz4_prepare_prefixes(ctx, h_x, n_jobs);
    if (r) return r;
    // linked-regime streams each own a 16 KiB table that is hit at random: keep the set L2 resident
    // (one warp per SM); single-block streams are short and want many more threads in flight
    bool any_linked = false;
    for (int64_t k = 0; k < n_jobs && !any_linked; ++k) any_linked = ctx->h_len[h_x[k]] >= LZ4_BLOCK / 2;
    int64_t inflight = ctx->streams_in_flight > 0 ? ctx->streams_in_flight : (any_linked ? 148 * 32 : 148 * 1024);
    const int threads = 32;
    int64_t blocks = (std::min<int64_t>(inflight, n_jobs) + threads - 1) / threads;
    if (blocks < 1) blocks = 1;
    r = ensure_work(ctx, (size_t)blocks * threads * LZ4_TABLE_BYTES);
    if (r) return r;
    CK(cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned long long), ctx->stream));
    CK(cudaEventRecord(ctx->evm0, ctx->stream));
    lz4_stream_kernel<<<(unsigned)blocks, threads, 0, ctx->stream>>>(
        ctx->d_corpus, ctx->d_off, ctx->d_len, ctx->d_jobx, pairs ? ctx->d_joby : nullptr, n_jobs, ctx->d_slot_of,
        ctx->d_ckpt_tab, ctx->d_ckpt_total, ctx->d_work, ctx->d_counter, ctx->d_out);
    CK(cudaEventRecord(ctx->evm1, ctx->stream));
    ctx->last_launches++;
    CK(cudaGetLastError());
    return SNACC_OK;
}

static int64_t wrapper_bytes(int codec) { return codec == SNACC_GZIP9 ? 18 : codec == SNACC_ZLIB6 ? 6 : 0; }

// common driver: job lists on host -> sizes on host
static int sizes_impl(snacc_ctx *ctx, int codec, const int32_t *xs, const int32_t *ys, int64_t n_jobs, int64_t *out)
{
    if (!ctx) return SNACC_ERR_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->err.clear();
    if (!ctx->n_seqs) FAIL(SNACC_ERR_STATE, "nothing uploaded");
    if (codec != SNACC_LZ4F && codec != SNACC_GZIP9 && codec != SNACC_ZLIB6)
        FAIL(SNACC_ERR_CODEC, "codec not supported on the GPU path (supported: lz4, gzip, zlib)");
    if (n_jobs < 0 || (n_jobs && (!xs || !out))) FAIL(SNACC_ERR_ARG, "bad job list");
    if (n_jobs == 0) return SNACC_OK;
    for (int64_t k = 0; k < n_jobs; ++k) {
        if (xs[k] < 0 || xs[k] >= ctx->n_seqs || (ys && (ys[k] < 0 || ys[k] >= ctx->n_seqs)))
            FAIL(SNACC_ERR_ARG, "sequence index out of range");
        if (ys && (uint64_t)ctx->h_len[xs[k]] + ctx->h_len[ys[k]] >= 0x7fffffffull)
            FAIL(SNACC_ERR_TOO_LARGE, "pair stream of 2 GiB or more");
    }
    CK(cudaSetDevice(ctx->device));
    int r = ensure_job_buffers(ctx, n_jobs);
    if (r) return r;
    CK(cudaMemcpyAsync(ctx->d_jobx, xs, sizeof(int32_t) * n_jobs, cudaMemcpyHostToDevice, ctx->stream));
    if (ys) CK(cudaMemcpyAsync(ctx->d_joby, ys, sizeof(int32_t) * n_jobs, cudaMemcpyHostToDevice, ctx->stream));
    ctx->last_launches = 0;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    if (codec == SNACC_LZ4F) r = run_lz4(ctx, xs, n_jobs, ys != nullptr);
    else {
        DeflateCorpus dc{ctx->d_corpus, ctx->d_off, ctx->d_len, ctx->h_off.data(), ctx->h_len.data(), ctx->n_seqs};
        r = deflate_run(ctx->dfl, dc, codec == SNACC_GZIP9 ? 9 : 6, xs, ys, ctx->d_jobx, ys ? ctx->d_joby : nullptr,
                        n_jobs, ctx->d_out, ctx->stream, ctx->streams_in_flight, &ctx->last_launches, ctx->err);
    }
    if (r) return r;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaMemcpyAsync(out, ctx->d_out, sizeof(int64_t) * n_jobs, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    if (codec == SNACC_LZ4F) { CK(cudaEventElapsedTime(&ms, ctx->evm0, ctx->evm1)); ctx->last_main_ms = ms; }
    const int64_t wb = wrapper_bytes(codec);
    if (wb) for (int64_t k = 0; k < n_jobs; ++k) out[k] += wb;
    return SNACC_OK;
}

extern "C" int snacc_single_sizes(snacc_ctx *ctx, int codec, const int32_t *idx, int64_t n, int64_t *out)
{
    return sizes_impl(ctx, codec, idx, nullptr, n, out);
}

extern "C" int snacc_pair_sizes(snacc_ctx *ctx, int codec, const int32_t *xs, const int32_t *ys, int64_t n_jobs,
                                int64_t *out)
{
    if (n_jobs && !ys) return SNACC_ERR_ARG;
    return sizes_impl(ctx, codec, xs, ys, n_jobs, out);
}

extern "C" int snacc_tile_sizes(snacc_ctx *ctx, int codec, int32_t row0, int32_t n_rows, int32_t col0,
                                int32_t n_cols, int64_t *out)
{
    if (!ctx) return SNACC_ERR_ARG;
    if (n_rows < 0 || n_cols < 0 || row0 < 0 || col0 < 0) return SNACC_ERR_ARG;
    const int64_t n = (int64_t)n_rows * n_cols;
    std::vector<int32_t> xs((size_t)n), ys((size_t)n);
    for (int32_t r = 0; r < n_rows; ++r)
        for (int32_t c = 0; c < n_cols; ++c) {
            xs[(size_t)r * n_cols + c] = row0 + r;
            ys[(size_t)r * n_cols + c] = col0 + c;
        }
    return sizes_impl(ctx, codec, xs.data(), ys.data(), n, out);
}

extern "C" int snacc_ncd(snacc_ctx *ctx, const int64_t *C, const int64_t *S, int32_t n, int formula, int32_t bias,
                         double *D)
{
    if (!ctx) return SNACC_ERR_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->err.clear();
    if (n <= 0 || !C || !S || !D) FAIL(SNACC_ERR_ARG, "snacc_ncd: bad arguments");
    if (formula != SNACC_NCD_REFERENCE && formula != SNACC_NCD_ONE_ORDER) FAIL(SNACC_ERR_ARG, "snacc_ncd: bad formula");
    CK(cudaSetDevice(ctx->device));
    int64_t *dC = nullptr, *dS = nullptr; double *dD = nullptr;
    const size_t nn = (size_t)n * n;
    CK(cudaMalloc(&dC, sizeof(int64_t) * n));
    CK(cudaMalloc(&dS, sizeof(int64_t) * nn));
    CK(cudaMalloc(&dD, sizeof(double) * nn));
    CK(cudaMemcpyAsync(dC, C, sizeof(int64_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dS, S, sizeof(int64_t) * nn, cudaMemcpyHostToDevice, ctx->stream));
    const int threads = 256;
    const int blocks = (int)std::min<size_t>((nn + threads - 1) / threads, 148 * 8);
    ncd_kernel<<<blocks, threads, 0, ctx->stream>>>(dC, dS, n, formula, bias, dD);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(D, dD, sizeof(double) * nn, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(dC); cudaFree(dS); cudaFree(dD);
    return SNACC_OK;
}

extern "C" int snacc_last_kernel_ms(const snacc_ctx *ctx, double *ms, int64_t *launches)
{
    if (!ctx) return SNACC_ERR_ARG;
    if (ms) *ms = ctx->last_ms;
    if (launches) *launches = ctx->last_launches;
    return SNACC_OK;
}

extern "C" int snacc_get_stat(const snacc_ctx *ctx, const char *name, double *out)
{
    if (!ctx || !name || !out) return SNACC_ERR_ARG;
    if (!strcmp(name, "main_kernel_ms")) { *out = ctx->last_main_ms; return SNACC_OK; }
    if (!strcmp(name, "total_kernel_ms")) { *out = ctx->last_ms; return SNACC_OK; }
    if (!strcmp(name, "launches")) { *out = (double)ctx->last_launches; return SNACC_OK; }
    return SNACC_ERR_ARG;
}

extern "C" int snacc_set_option(snacc_ctx *ctx, const char *name, int64_t value)
{
    if (!ctx || !name) return SNACC_ERR_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!strcmp(name, "streams_in_flight")) { ctx->streams_in_flight = value; return SNACC_OK; }
    if (!strcmp(name, "invalidate_caches")) {
        // forget every per-sequence precomputation (prefix checkpoints ...) so the next sizes call
        // redoes the whole job; used by bench.py so that no step reuses work of an earlier step
        std::fill(ctx->ckpt_done.begin(), ctx->ckpt_done.end(), 0);
        deflate_invalidate(ctx->dfl);
        return SNACC_OK;
    }
    ctx->err = std::string("unknown option: ") + name;
    return SNACC_ERR_ARG;
}
