// Device-wide exclusive prefix sum (three kernels: tile sums, scan of tile sums, tile scan + offset).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace bm25cu {

constexpr int kScanThreads = 1024;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

// Block-wide exclusive scan of one value per thread (blockDim.x <= 1024). Returns the exclusive prefix; total in *total.
__device__ __forceinline__ long long block_exclusive_scan(long long v, long long *warp_sums /*33*/, long long *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        long long w = (lane < nwarp) ? warp_sums[lane] : 0;
        long long s = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        if (lane < nwarp) warp_sums[lane] = s - w;  // exclusive warp offsets
        if (lane == 31) warp_sums[32] = s;
    }
    __syncthreads();
    long long res = warp_sums[warp] + x - v;
    if (total) *total = warp_sums[32];
    __syncthreads();
    return res;
}

template <typename In>
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const In *__restrict__ in, long long n,
                                                                      long long *tile_sums) {
    __shared__ long long ws[33];
    const long long base = (long long)blockIdx.x * kScanTile;
    long long s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        long long idx = base + (long long)i * kScanThreads + threadIdx.x;
        if (idx < n) s += (long long)in[idx];
    }
    long long total;
    block_exclusive_scan(s, ws, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

static __global__ void __launch_bounds__(kScanThreads) scan_tile_offsets_kernel(long long *tile_sums, long long n_tiles,
                                                                         long long *grand_total) {
    __shared__ long long ws[33];
    long long carry = 0;
    for (long long base = 0; base < n_tiles; base += kScanThreads) {
        long long idx = base + threadIdx.x;
        long long v = (idx < n_tiles) ? tile_sums[idx] : 0;
        long long total;
        long long ex = block_exclusive_scan(v, ws, &total);
        if (idx < n_tiles) tile_sums[idx] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0 && grand_total) *grand_total = carry;
}

// out[i] = sum of in[0..i) for i in [0, n); out[n] = total (out has n + 1 slots).
template <typename In>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const In *__restrict__ in, long long n,
                                                                  const long long *__restrict__ tile_offsets,
                                                                  long long *out) {
    __shared__ long long ws[33];
    const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
    long long v[kScanItems];
    long long s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        long long idx = base + i;
        v[i] = (idx < n) ? (long long)in[idx] : 0;
        s += v[i];
    }
    long long ex = block_exclusive_scan(s, ws, nullptr) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        long long idx = base + i;
        if (idx < n) out[idx] = ex;
        ex += v[i];
        if (idx == n - 1) out[n] = ex;
    }
}

// Host helper. `tmp` must hold ceil(n / kScanTile) + 1 long longs.
template <typename In>
inline cudaError_t exclusive_scan(const In *d_in, long long n, long long *d_out, long long *d_tmp, cudaStream_t stream) {
    if (n <= 0) return cudaMemsetAsync(d_out, 0, sizeof(long long), stream);
    const long long n_tiles = (n + kScanTile - 1) / kScanTile;
    scan_tile_sums_kernel<In><<<(unsigned)n_tiles, kScanThreads, 0, stream>>>(d_in, n, d_tmp);
    scan_tile_offsets_kernel<<<1, kScanThreads, 0, stream>>>(d_tmp, n_tiles, nullptr);
    scan_apply_kernel<In><<<(unsigned)n_tiles, kScanThreads, 0, stream>>>(d_in, n, d_tmp, d_out);
    return cudaGetLastError();
}

inline size_t scan_tmp_elems(long long n) { return (size_t)((n + kScanTile - 1) / kScanTile + 1); }

}  // namespace bm25cu
