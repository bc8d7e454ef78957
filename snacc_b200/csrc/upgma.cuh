// upgma.cuh -- K5: the downstream step of the distance matrix on the device, sm_100a.
//   metrify      snacc/misc.py:20-25               D_sym = 0.5 (D + D^T), zero diagonal
//   hierarchical snacc/distmatrix_to_tree.py:9-15  scipy.cluster.hierarchy.linkage(squareform(D_sym), method='average')
// Output is scipy's linkage matrix Z ((n-1) x 4: cluster ids a < b, height, leaf count; new clusters are numbered n, n+1,
// ... in merge order).  UPGMA is a reducible linkage, so merging the globally closest pair each step gives the same
// hierarchy, in the same (non-decreasing height) order scipy sorts its nearest-neighbour-chain result into; heights use
// the same update d(k, a+b) = (|a| d(k,a) + |b| d(k,b)) / (|a| + |b|).  (Exactly tied heights may be ordered
// differently; real NCD matrices have none.)
//
// One CTA walks the n-1 merges (the chain of merges is serial); all the work inside a step is parallel over its 1024
// threads: a nearest-neighbour entry per row (closest active column to its right) makes the global minimum an O(n)
// reduction and the update O(n) plus the rescans of the few rows whose neighbour was merged away.  The matrix stays in
// HBM / L2 (n = 10 000: 800 MB); nothing here is a dense contraction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace snacc {

__global__ void metrify_kernel(const double *__restrict__ D, int32_t n, int do_metrify, double *__restrict__ M)
{
    const int64_t total = (int64_t)n * n;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t i = (int32_t)(k / n), j = (int32_t)(k % n);
        M[k] = !do_metrify ? D[k] : (i == j ? 0.0 : 0.5 * (D[k] + D[(int64_t)j * n + i]));
    }
}

constexpr int UPGMA_THREADS = 1024;

// closest active column j > k of row k (ties: smallest j), by one warp
__device__ __forceinline__ void upgma_rescan(const double *M, int32_t n, const uint8_t *active, int32_t k, int32_t lane,
                                             double *nn_val, int32_t *nn_idx)
{
    double bv = 1.0 / 0.0;
    int32_t bj = -1;
    for (int32_t j = k + 1 + lane; j < n; j += 32) {
        if (!active[j]) continue;
        const double v = M[(int64_t)k * n + j];
        if (v < bv) { bv = v; bj = j; }
    }
    for (int off = 16; off; off >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, bv, off);
        const int32_t oj = __shfl_down_sync(0xffffffffu, bj, off);
        if (oj >= 0 && (bj < 0 || ov < bv || (ov == bv && oj < bj))) { bv = ov; bj = oj; }
    }
    if (lane == 0) { nn_val[k] = bv; nn_idx[k] = bj; }
}

__global__ void __launch_bounds__(UPGMA_THREADS)
upgma_kernel(double *__restrict__ M, int32_t n, double *__restrict__ Z, int32_t *__restrict__ size, int32_t *__restrict__ label,
             int32_t *__restrict__ nn_idx, double *__restrict__ nn_val, uint8_t *__restrict__ active, int32_t *__restrict__ todo)
{
    __shared__ double s_val[UPGMA_THREADS / 32];
    __shared__ int32_t s_row[UPGMA_THREADS / 32];
    __shared__ int32_t s_a, s_b, s_ntodo;
    const int32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = UPGMA_THREADS / 32;
    for (int32_t i = tid; i < n; i += UPGMA_THREADS) { size[i] = 1; label[i] = i; active[i] = 1; }
    __syncthreads();
    for (int32_t k = warp; k < n; k += nwarp) upgma_rescan(M, n, active, k, lane, nn_val, nn_idx);
    __syncthreads();
    for (int32_t t = 0; t + 1 < n; ++t) {
        // (1) the closest pair: minimum over the rows' nearest neighbours (ties: smallest row)
        double bv = 1.0 / 0.0;
        int32_t bi = -1;
        for (int32_t i = tid; i < n; i += UPGMA_THREADS) {
            if (!active[i] || nn_idx[i] < 0) continue;
            const double v = nn_val[i];
            if (v < bv || (v == bv && (bi < 0 || i < bi))) { bv = v; bi = i; }
        }
        for (int off = 16; off; off >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, bv, off);
            const int32_t oi = __shfl_down_sync(0xffffffffu, bi, off);
            if (oi >= 0 && (bi < 0 || ov < bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
        }
        if (lane == 0) { s_val[warp] = bv; s_row[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            double v = s_val[0]; int32_t r = s_row[0];
            for (int w = 1; w < nwarp; ++w)
                if (s_row[w] >= 0 && (r < 0 || s_val[w] < v || (s_val[w] == v && s_row[w] < r))) { v = s_val[w]; r = s_row[w]; }
            const int32_t a = r, b = nn_idx[r];
            s_a = a; s_b = b; s_ntodo = 0;
            const int32_t la = label[a], lb = label[b];
            Z[4 * t + 0] = (double)(la < lb ? la : lb);
            Z[4 * t + 1] = (double)(la < lb ? lb : la);
            Z[4 * t + 2] = v;
            Z[4 * t + 3] = (double)(size[a] + size[b]);
        }
        __syncthreads();
        const int32_t a = s_a, b = s_b;                     // a < b: the merged cluster takes slot a
        const double sa = (double)size[a], sb = (double)size[b];
        // (2) distances of the merged cluster, and what they do to the rows' nearest neighbours
        for (int32_t k = tid; k < n; k += UPGMA_THREADS) {
            if (!active[k] || k == a || k == b) continue;
            const double dka = M[(int64_t)k * n + a], dkb = M[(int64_t)k * n + b];
            const double d = (sa * dka + sb * dkb) / (sa + sb);
            M[(int64_t)k * n + a] = d;
            M[(int64_t)a * n + k] = d;
            if (k < a) {
                if (nn_idx[k] == a || nn_idx[k] == b) todo[atomicAdd(&s_ntodo, 1)] = k;
                else if (d < nn_val[k] || (d == nn_val[k] && a < nn_idx[k])) { nn_val[k] = d; nn_idx[k] = a; }
            } else if (k < b) {
                if (nn_idx[k] == b) todo[atomicAdd(&s_ntodo, 1)] = k;
            }
        }
        __syncthreads();
        if (tid == 0) {
            size[a] += size[b]; label[a] = n + t; active[b] = 0;
            todo[s_ntodo++] = a;
        }
        __syncthreads();
        // (3) rows whose nearest neighbour was merged away (and the merged row itself): full rescan, one warp per row
        const int32_t nt = s_ntodo;
        for (int32_t q = warp; q < nt; q += nwarp) upgma_rescan(M, n, active, todo[q], lane, nn_val, nn_idx);
        __syncthreads();
    }
}

}  // namespace snacc
