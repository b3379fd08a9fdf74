// Stable LSD radix sort of (64-bit key, value) pairs in HBM, 8-bit digits: the ordering step of add_hci (first
// encounter order of the new determinants, hci.cu) and the transposition step of the incremental update (update.cu).
// Three kernels per digit: per-tile digit histograms, one exclusive scan over (digit, tile), and a scatter that ranks
// the keys of a tile in their original order (warp ballots group equal digits, so no atomics and the sort is stable).
#pragma once
#include "common.cuh"

int scan_counts(pyci_ctx *ctx, const int *cnt, long n, long *indptr, int *maxcnt);

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;                        // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;     // keys per CTA
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_WSLICE = 32 * RS_ITEMS;           // contiguous keys owned by one warp

// lanes of the warp holding the same 8-bit digit (valid lanes only)
__device__ __forceinline__ u32 rs_peers(u32 d, bool valid) {
    u32 peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
        const bool on = (d >> bit) & 1u;
        const u32 bal = __ballot_sync(0xffffffffu, on);
        peers &= on ? bal : ~bal;
    }
    return peers;
}

// hist[digit * ntiles + tile] = keys of the tile with that digit
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const u64 *__restrict__ keys, long n, int shift, long ntiles,
                                                             int *__restrict__ hist) {
    __shared__ int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const long base = (long)blockIdx.x * RS_TILE;
    for (int q = 0; q < RS_ITEMS; ++q) {
        const long i = base + q * RS_THREADS + threadIdx.x;
        if (i < n)
            atomicAdd(&h[(u32)(keys[i] >> shift) & 255u], 1);
    }
    __syncthreads();
    hist[(long)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

template<class V>
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const u64 *__restrict__ keys, const V *__restrict__ vals,
                                                                long n, int shift, long ntiles,
                                                                const long *__restrict__ offs, u64 *__restrict__ keys_out,
                                                                V *__restrict__ vals_out) {
    __shared__ u32 wh[RS_WARPS][256]; // per warp: count, then first output position, of every digit
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u32 lt = (1u << lane) - 1u;
    for (int t = threadIdx.x; t < RS_WARPS * 256; t += RS_THREADS)
        (&wh[0][0])[t] = 0u;
    __syncthreads();
    // warp w owns keys [base + w * RS_WSLICE, +RS_WSLICE), lane-strided chunks of 32 in their original order
    const long wbase = (long)blockIdx.x * RS_TILE + (long)w * RS_WSLICE;
    u64 k[RS_ITEMS];
    u32 peers[RS_ITEMS];
#pragma unroll
    for (int q = 0; q < RS_ITEMS; ++q) {
        const long i = wbase + q * 32 + lane;
        const bool valid = i < n;
        k[q] = valid ? keys[i] : 0ULL;
        const u32 d = (u32)(k[q] >> shift) & 255u;
        peers[q] = rs_peers(d, valid);
        if (valid && lane == __ffs(peers[q]) - 1)
            wh[w][d] += (u32)__popc(peers[q]); // one lane per digit and chunk: no atomics
        __syncwarp();
    }
    __syncthreads();
    // first position of (digit, warp): tile offset from the global scan, then the warps in order
    {
        const int d = threadIdx.x;
        long run = offs[(long)d * ntiles + blockIdx.x];
        for (int q = 0; q < RS_WARPS; ++q) {
            const u32 c = wh[q][d];
            wh[q][d] = (u32)(run - offs[(long)d * ntiles + blockIdx.x]); // relative to the tile's digit offset
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < RS_ITEMS; ++q) {
        const long i = wbase + q * 32 + lane;
        const bool valid = i < n;
        const u32 d = (u32)(k[q] >> shift) & 255u;
        if (valid) {
            const long pos = offs[(long)d * ntiles + blockIdx.x] + wh[w][d] + __popc(peers[q] & lt);
            keys_out[pos] = k[q];
            vals_out[pos] = vals[i];
        }
        __syncwarp();
        if (valid && lane == __ffs(peers[q]) - 1)
            wh[w][d] += (u32)__popc(peers[q]);
        __syncwarp();
    }
}

// Sorts (keys, vals)[0, n) by key bits [0, end_bit) ascending, stable.  keys_alt / vals_alt: scratch of the same size.
// On return *in_alt tells whether the result sits in the alt buffers (odd number of passes) or in the originals.
template<class V>
int radix_sort_pairs(pyci_ctx *ctx, u64 *keys, u64 *keys_alt, V *vals, V *vals_alt, long n, int end_bit, bool *in_alt) {
    *in_alt = false;
    if (n <= 1)
        return PYCI_OK;
    cudaStream_t st = ctx->stream;
    const long ntiles = (n + RS_TILE - 1) / RS_TILE;
    int *hist = nullptr;
    long *offs = nullptr;
    PYCI_CUDA(dev_malloc(&hist, sizeof(int) * (size_t)(256 * ntiles)));
    cudaError_t e = dev_malloc(&offs, sizeof(long) * (size_t)(256 * ntiles + 1));
    if (e != cudaSuccess) {
        dev_free(hist);
        PYCI_CUDA(e);
    }
    int rc = PYCI_OK;
    u64 *kin = keys, *kout = keys_alt;
    V *vin = vals, *vout = vals_alt;
    for (int shift = 0; shift < end_bit && rc == PYCI_OK; shift += 8) {
        rs_hist_kernel<<<(unsigned)ntiles, RS_THREADS, 0, st>>>(kin, n, shift, ntiles, hist);
        rc = scan_counts(ctx, hist, 256 * ntiles, offs, nullptr);
        if (rc != PYCI_OK)
            break;
        rs_scatter_kernel<V><<<(unsigned)ntiles, RS_THREADS, 0, st>>>(kin, vin, n, shift, ntiles, offs, kout, vout);
        ctx->launches += 2;
        std::swap(kin, kout);
        std::swap(vin, vout);
        *in_alt = !*in_alt;
    }
    if (rc == PYCI_OK && cudaGetLastError() != cudaSuccess) {
        pyci_set_error("CUDA error in the radix sort");
        rc = PYCI_ERR_CUDA;
    }
    dev_free(hist);
    dev_free(offs);
    return rc;
}

} // namespace
