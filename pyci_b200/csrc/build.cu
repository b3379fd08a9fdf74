// Construction of the CI Hamiltonian in CSR on the device: the B200 form of
// SparseOp::update / add_row / sort_row (/root/reference/pyci/src/sparseop.cpp:186-502).
//
//   index   : open-addressing hash of the determinant strings (replaces Wfn::index_det's
//             SpookyHash + flat_hash_map, pyci.h:121-128,157)
//   count   : one CTA per row, every thread enumerates excitations with bit tricks and probes the
//             hash; per-row hit counts                      (skipped when the space is complete)
//   scan    : int64 exclusive scan of the counts -> indptr
//   fill    : same enumeration; hits are appended (warp-aggregated) to a shared-memory row buffer as
//             (column << 32 | excitation code), sorted by column in shared memory (sort_row,
//             sparseop.cpp:214-218), then matrix elements are evaluated from the codes
//             (Slater-Condon rules with the reference's phases and summation order) and the row is
//             streamed to HBM with coalesced stores.
#include <algorithm>
#include <cstring>
#include <memory>
#include <vector>

#include "elements.cuh"
#include "join.cuh"

namespace {

// one thread per row: H_ii of this rank's rows (also the Davidson preconditioner)
template<int KIND>
__global__ void diag_kernel(BuildParams P) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.nloc)
        return;
    const long row = P.row0 + r;
    const u64 a = P.dets[row * P.nwords], b = (KIND == PYCI_FULLCI) ? P.dets[row * P.nwords + 1] : 0ULL;
    P.diag[r] = (KIND == PYCI_DOCI) ? diag_doci(P, a) : diag_twobody(P, a, b);
}

constexpr int UNROLL = 4;
constexpr int RANK_SORT_MAX = 320; // rows up to this many entries are ordered by a rank sort, longer ones by the radix sort
constexpr int RANK_SORT_MIN = 48;  // ... and rows of more than this many by the bucketed form of it
constexpr int SORT_BUCKETS = 512;  // buckets of the bucketed rank sort: a monotone map of [0, ncol) (two u32 arrays)

// ---- count pass ---------------------------------------------------------------------------------
// hitlist != nullptr: the hits of a row (candidate index, column) are also recorded, up to `cap` per row, so that
// the fill pass of a sparse (selected) space evaluates and sorts only those instead of enumerating and probing
// every excitation a second time; rows with more hits than `cap` are enumerated again by the fill pass.
template<int KIND, int KM>
__global__ void __launch_bounds__(256) count_kernel(BuildParams P, DetIndex<KM> index, u32 nSa, u32 nSb, uint2 *hitlist,
                                                    int cap, u32 pair_bytes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const RowTables T = carve_tables(smem_raw, nSa, nSb); // only the strings are allocated and used in this pass
    uchar2 *pairs = reinterpret_cast<uchar2 *>(smem_raw + table_strings_bytes(nSa, nSb));
    __shared__ RowShared rs;
    __shared__ int warp_sums[8];
    fill_pairs(pairs, P.npairs_dim);
    const int nspin = (KIND == PYCI_FULLCI) ? 2 : 1;
    if (KIND != PYCI_DOCI)
        pair_masks_carve(rs, P, smem_raw + table_strings_bytes(nSa, nSb) + pair_bytes, nspin);
    else
        pair_masks_none(rs);
    const int lane = threadIdx.x & 31;
    const u32 lt = (1u << lane) - 1u;
    for (long r = blockIdx.x; r < P.nloc; r += gridDim.x) {
        const long row = P.row0 + r;
        __syncthreads();
        row_setup(rs, P, row, nspin); // rs.count = 0: slot counter of the recorded hits
        __syncthreads();
        if (KIND != PYCI_DOCI) {
            build_tables<KIND, false>(P, rs, T, nSa, nSb);
            pair_masks_build(rs, P, pairs, nspin);
            __syncthreads();
        }
        uint2 *hrow = hitlist ? hitlist + (size_t)r * cap : nullptr;
        int cnt = 0;
        for (u32 base = 0; base < P.ncand; base += UNROLL * blockDim.x) {
            int hit[UNROLL];
            u64 A[UNROLL], B[UNROLL];
            bool want[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const u32 c = base + u * blockDim.x + threadIdx.x;
                want[u] = c < P.ncand;
                A[u] = B[u] = 0ULL;
                if (want[u]) {
                    double unused;
                    candidate<KIND, false>(P, rs, T, pairs, c, A[u], B[u], unused);
                }
            }
            find_batch<UNROLL>(index, A, B, want, hit);
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const bool keep = hit[u] >= 0 && hit[u] < P.ncol;
                cnt += keep;
                if (hrow) { // warp-aggregated append
                    const u32 msk = __ballot_sync(0xffffffffu, keep);
                    if (msk) {
                        int slot = 0;
                        const int leader = __ffs(msk) - 1;
                        if (lane == leader)
                            slot = atomicAdd(&rs.count, __popc(msk));
                        slot = __shfl_sync(0xffffffffu, slot, leader) + __popc(msk & lt);
                        if (keep && slot < cap)
                            hrow[slot] = make_uint2(base + u * blockDim.x + threadIdx.x, (u32)hit[u]);
                    }
                }
            }
        }
        // block reduce
        for (int o = 16; o > 0; o >>= 1)
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if ((threadIdx.x & 31) == 0)
            warp_sums[threadIdx.x >> 5] = cnt;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = (row < P.ncol) ? 1 : 0; // diagonal, sparseop.cpp:252-255
            for (int wq = 0; wq < (int)((blockDim.x + 31) >> 5); ++wq)
                tot += warp_sums[wq];
            P.rowcnt[r] = tot;
        }
    }
}

// ---- fill pass ----------------------------------------------------------------------------------

// lanes of the warp holding the same digit as this lane (valid lanes only): dbits ballots, independent
// of how many distinct digits the warp holds (match.any serialises over the distinct values)
__device__ __forceinline__ u32 digit_peers(u32 d, bool valid, int dbits) {
    u32 peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
        if (bit < dbits) {
            const bool on = (d >> bit) & 1u;
            const u32 bal = __ballot_sync(0xffffffffu, on);
            peers &= on ? bal : ~bal;
        }
    }
    return peers;
}

// Stable LSD radix sort of a[0,m) (ping-pong with b) on key bits [32, 32 + passes*dbits): the per-row
// sort_row of sparseop.cpp:214-218.  Each warp owns a contiguous slice of the row and a private counter
// row hist[w][bin]; lanes holding equal digits are grouped by ballots so that a counter is touched by one
// lane, which keeps the scatter stable without atomics.  Returns the buffer holding the result.
__device__ u64 *block_radix_sort(u64 *a, u64 *b, int m, int passes, int dbits, u32 *hist, u32 *tot) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const u32 lt = (1u << lane) - 1u;
    const int nbins = 1 << dbits;
    const u32 dmask = (u32)(nbins - 1);
    const int seg = (((m + nw - 1) / nw) + 31) & ~31;
    const int begin = min(m, w * seg), end = min(m, begin + seg);
    for (int pass = 0; pass < passes; ++pass) {
        const int shift = 32 + pass * dbits;
        for (int t = threadIdx.x; t < nw * nbins; t += blockDim.x)
            hist[t] = 0;
        __syncthreads();
        u32 *myhist = hist + w * nbins;
        // count: order does not matter, leaders add with shared-memory atomics, two chunks in flight
        for (int base = begin; base < end; base += 64) {
            const int i0 = base + lane, i1 = base + 32 + lane;
            const bool v0 = i0 < end, v1 = i1 < end;
            const u32 d0 = v0 ? (u32)(a[i0] >> shift) & dmask : 0u;
            const u32 d1 = v1 ? (u32)(a[i1] >> shift) & dmask : 0u;
            const u32 p0 = digit_peers(d0, v0, dbits);
            const u32 p1 = digit_peers(d1, v1, dbits);
            if (v0 && lane == __ffs(p0) - 1)
                atomicAdd(&myhist[d0], (u32)__popc(p0));
            if (v1 && lane == __ffs(p1) - 1)
                atomicAdd(&myhist[d1], (u32)__popc(p1));
        }
        __syncthreads();
        // exclusive scan in (bin, warp) order
        for (int bin = threadIdx.x; bin < nbins; bin += blockDim.x) {
            u32 run = 0;
            for (int q = 0; q < nw; ++q) {
                const u32 t = hist[q * nbins + bin];
                hist[q * nbins + bin] = run;
                run += t;
            }
            tot[bin] = run;
        }
        __syncthreads();
        if (w == 0) {
            const int per = nbins >> 5; // nbins >= 32
            u32 local = 0;
            for (int q = 0; q < per; ++q)
                local += tot[lane * per + q];
            u32 incl = local;
            for (int o = 1; o < 32; o <<= 1) {
                const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o)
                    incl += t;
            }
            u32 run = incl - local;
            for (int q = 0; q < per; ++q) {
                const u32 t = tot[lane * per + q];
                tot[lane * per + q] = run;
                run += t;
            }
        }
        __syncthreads();
        for (int t = threadIdx.x; t < nw * nbins; t += blockDim.x)
            hist[t] += tot[t & (nbins - 1)];
        __syncthreads();
        // scatter in slice order; the keys / ballots of the next chunk are prepared before the counter
        // read-modify-write of the current one
        {
            int idx = begin + lane;
            bool valid = idx < end;
            u64 k = valid ? a[idx] : 0ULL;
            u32 d = (u32)(k >> shift) & dmask;
            u32 peers = digit_peers(d, valid, dbits);
            for (int base = begin; base < end; base += 32) {
                const int nidx = base + 32 + lane;
                const bool nvalid = nidx < end;
                const u64 nk = nvalid ? a[nidx] : 0ULL;
                const u32 nd = (u32)(nk >> shift) & dmask;
                const u32 npeers = (base + 32 < end) ? digit_peers(nd, nvalid, dbits) : 0u;
                u32 pos = 0;
                if (valid)
                    pos = myhist[d] + __popc(peers & lt);
                __syncwarp();
                if (valid) {
                    b[pos] = k;
                    if (lane == __ffs(peers) - 1)
                        myhist[d] += __popc(peers);
                }
                __syncwarp();
                valid = nvalid;
                k = nk;
                d = nd;
                peers = npeers;
            }
        }
        __syncthreads();
        u64 *t = a;
        a = b;
        b = t;
    }
    return a;
}

// LAZY: evaluate the matrix element only for candidates that hit (selected spaces where most excitations
// leave the wave function); otherwise element loads are issued together with the probes.
// hitlist != nullptr: rows whose hits were recorded by the count pass (at most `cap`) are assembled from the record.
template<int KIND, int KM, bool LAZY>
__global__ void __launch_bounds__(256) fill_kernel(BuildParams P, DetIndex<KM> index, u32 nSa, u32 nSb,
                                                   const uint2 *__restrict__ hitlist, int cap, u64 *gscratch,
                                                   int short_cap, int staged) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: keys A | keys B | values | radix counters | excitation tables | pair table.  Rows of more than
    // `short_cap` entries (a few dominant determinants of a heat-bath space reach 10^4) do not fit shared memory:
    // a second launch (gscratch != nullptr) builds exactly those, with the three row buffers in a per-CTA slab of HBM.
    u64 *keyA = gscratch ? gscratch + (size_t)blockIdx.x * 3 * (size_t)P.maxrow : reinterpret_cast<u64 *>(smem_raw);
    u64 *keyB = keyA + P.maxrow;
    double *valbuf = reinterpret_cast<double *>(keyB + P.maxrow);
    u32 *hist = gscratch ? reinterpret_cast<u32 *>(smem_raw) : reinterpret_cast<u32 *>(valbuf + P.maxrow);
    const int nbins = 1 << P.sort_dbits, nw = blockDim.x >> 5;
    u32 *tot = hist + nw * nbins;
    // (the counters double as the two arrays of the bucketed rank sort: sort_counter_words on the host side)
    unsigned char *tbase = reinterpret_cast<unsigned char *>(hist + max((nw + 1) * nbins, 2 * SORT_BUCKETS));
    const RowTables T = carve_tables(tbase, nSa, nSb);
    uchar2 *pairs = reinterpret_cast<uchar2 *>(tbase + tables_bytes(nSa, nSb));
    __shared__ RowShared rs;
    __shared__ int low_count;
    pair_masks_none(rs);
    fill_pairs(pairs, P.npairs_dim);
    const int nspin = (KIND == PYCI_FULLCI) ? 2 : 1;
    const int lane = threadIdx.x & 31;
    const u32 lt = (1u << lane) - 1u;
    for (long r = blockIdx.x; r < P.nloc; r += gridDim.x) {
        const long row = P.row0 + r;
        {
            const long len = P.indptr[r + 1] - P.indptr[r];
            if (gscratch ? len <= (long)short_cap : len > (long)short_cap)
                continue; // the other launch's row (uniform over the CTA)
        }
        __syncthreads();
        row_setup(rs, P, row, nspin);
        if (threadIdx.x == 0)
            low_count = 0;
        __syncthreads();
        const int ndiag = (row < P.ncol) ? 1 : 0; // the diagonal takes slot 0 below
        const int nh = (int)(P.indptr[r + 1] - P.indptr[r]) - ndiag;
        const bool recorded = hitlist != nullptr && nh <= cap; // hits of this row were recorded by the count pass
        // ... or, when they overflowed the list, staged in the row's place in the CSR arrays by a second join pass
        const bool staged_row = staged != 0 && !recorded;
        if (KIND != PYCI_DOCI && !recorded && !staged_row)
            build_tables<KIND, true>(P, rs, T, nSa, nSb);
        if (threadIdx.x == 0 && row < P.ncol) { // diagonal, sparseop.cpp:252-255 / :421-424 / :496-499
            const int slot = atomicAdd(&rs.count, 1);
            keyA[slot] = ((u64)row << 32) | (u32)slot;
            valbuf[slot] = P.diag[r];
            atomicAdd(&low_count, 1);
        }
        __syncthreads();
        int nlow = 0;
        if (recorded || staged_row) {
            // sparse row recorded by the count pass: evaluate the hits only (no enumeration, no probes, no tables)
            const uint2 *hrow = hitlist + (size_t)r * cap;
            const long st0 = P.indptr[r] + ndiag;
            for (int k = threadIdx.x; k < nh; k += blockDim.x) {
                const uint2 h = recorded ? hrow[k]
                                         : make_uint2((u32)__double_as_longlong(P.vals[st0 + k]), (u32)P.cols[st0 + k]);
                const double v = hit_element<KIND>(P, rs, pairs, h.x);
                const int slot = ndiag + k;
                keyA[slot] = ((u64)h.y << 32) | (u32)slot;
                valbuf[slot] = v;
                nlow += ((long)h.y <= row);
            }
            if (threadIdx.x == 0)
                rs.count = ndiag + nh;
        } else {
            for (u32 base = 0; base < P.ncand; base += UNROLL * blockDim.x) {
                int hit[UNROLL];
                double val[UNROLL];
    #pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    const u32 c = base + u * blockDim.x + threadIdx.x;
                    hit[u] = -1;
                    val[u] = 0.0;
                    if (c < P.ncand) {
                        u64 A, B;
                        if (LAZY)
                            candidate<KIND, false>(P, rs, T, pairs, c, A, B, val[u]);
                        else
                            candidate<KIND, true>(P, rs, T, pairs, c, A, B, val[u]);
                        hit[u] = index.find(A, B);
                    }
                }
    #pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    // warp-aggregated append: one shared atomic per warp per step
                    const bool keep = hit[u] >= 0 && hit[u] < P.ncol;
                    if (LAZY && keep) {
                        u64 A, B;
                        candidate<KIND, true>(P, rs, T, pairs, base + u * blockDim.x + threadIdx.x, A, B, val[u]);
                    }
                    const u32 msk = __ballot_sync(0xffffffffu, keep);
                    if (msk) {
                        int slot = 0;
                        const int leader = __ffs(msk) - 1;
                        if (lane == leader)
                            slot = atomicAdd(&rs.count, __popc(msk));
                        slot = __shfl_sync(0xffffffffu, slot, leader) + __popc(msk & lt);
                        if (keep) {
                            keyA[slot] = ((u64)(u32)hit[u] << 32) | (u32)slot;
                            valbuf[slot] = val[u];
                            nlow += ((long)hit[u] <= row);
                        }
                    }
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1)
            nlow += __shfl_xor_sync(0xffffffffu, nlow, o);
        if (lane == 0 && nlow)
            atomicAdd(&low_count, nlow);
        __syncthreads();
        const int m = rs.count;
        const u64 *sorted;
        if (m > RANK_SORT_MIN && m <= RANK_SORT_MAX) {
            // Bucketed rank sort.  ncu of the plain rank sort below on a config-5-style space (profiles/r3q): two thirds
            // of the kernel's instructions, m / 2 shared-memory loads and m compares for every entry.  A column maps
            // monotonically to one of 512 buckets of [0, ncol); counts, one scan and a scatter group the row by bucket
            // (in bucket order = column order), and an entry is then ranked only against the few entries of its own
            // bucket -- the connected determinants of a selected space are spread over the whole list, runs of
            // neighbouring columns are a handful long.  A row whose entries all fall into one bucket costs what the
            // plain rank sort costs, never more (m <= 320).
            u32 *cnt = hist, *bstart = hist + SORT_BUCKETS;
            const u64 mul = ((u64)SORT_BUCKETS << 32) / (u64)max(P.ncol, 1L); // bucket = col * mul >> 32 < SORT_BUCKETS
            for (int b = threadIdx.x; b < SORT_BUCKETS; b += blockDim.x)
                cnt[b] = 0u;
            __syncthreads();
            for (int e = threadIdx.x; e < m; e += blockDim.x)
                atomicAdd(&cnt[(u32)(((keyA[e] >> 32) * mul) >> 32)], 1u);
            __syncthreads();
            if (threadIdx.x < 32) { // exclusive scan of the counts: sixteen consecutive buckets per lane
                constexpr int PER = SORT_BUCKETS / 32;
                u32 loc = 0u;
#pragma unroll
                for (int q = 0; q < PER; ++q)
                    loc += cnt[lane * PER + q];
                u32 run = loc;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const u32 up = __shfl_up_sync(0xffffffffu, run, o);
                    if (lane >= o)
                        run += up;
                }
                run -= loc;
#pragma unroll
                for (int q = 0; q < PER; ++q) {
                    const u32 c = cnt[lane * PER + q];
                    bstart[lane * PER + q] = run;
                    cnt[lane * PER + q] = run; // cursor of the scatter; ends up at the bucket's end
                    run += c;
                }
            }
            __syncthreads();
            for (int e = threadIdx.x; e < m; e += blockDim.x) {
                const u64 key = keyA[e];
                keyB[atomicAdd(&cnt[(u32)(((key >> 32) * mul) >> 32)], 1u)] = key;
            }
            __syncthreads();
            for (int e = threadIdx.x; e < m; e += blockDim.x) {
                const u64 key = keyB[e];
                const u32 col = (u32)(key >> 32), b = (u32)(((u64)col * mul) >> 32);
                const u32 s = bstart[b], t = cnt[b];
                u32 rank = s;
                for (u32 f = s; f < t; ++f)
                    rank += (u32)(keyB[f] >> 32) < col;
                keyA[rank] = key;
            }
            __syncthreads();
            sorted = keyA;
        } else if (m <= RANK_SORT_MAX) {
            // short row (the rows of a selected space hold ~10^2 entries): rank sort -- entry e goes to the number of
            // entries with a smaller key; every thread reads the same key at a time (shared-memory broadcast).  The
            // radix sort below costs ~10^4 warp instructions per row whatever its length (ncu r2d).
            // (the columns of a row are distinct, so the column word alone orders the keys: two keys per 16-byte
            // shared-memory load and 32-bit compares -- half the instructions of comparing whole keys one by one)
            const uint4 *kq = reinterpret_cast<const uint4 *>(keyA);
            const int mp = m >> 1;
            for (int e = threadIdx.x; e < m; e += blockDim.x) {
                const u64 key = keyA[e];
                const u32 col = (u32)(key >> 32);
                int rank = 0;
                for (int f = 0; f < mp; ++f) {
                    const uint4 q = kq[f]; // (slot, column) of keys 2f and 2f + 1
                    rank += (q.y < col) + (q.w < col);
                }
                if (m & 1)
                    rank += (u32)(keyA[m - 1] >> 32) < col;
                keyB[rank] = key;
            }
            __syncthreads();
            sorted = keyB;
        } else {
            sorted = block_radix_sort(keyA, keyB, m, P.sort_passes, P.sort_dbits, hist, tot);
        }
        // stream the row out: columns ascending, values gathered through the slot index
        const long out0 = P.indptr[r];
        for (int e = threadIdx.x; e < m; e += blockDim.x) {
            const u64 kv = sorted[e];
            P.cols[out0 + e] = (int)(kv >> 32);
            P.vals[out0 + e] = valbuf[(u32)kv];
        }
        if (threadIdx.x == 0)
            P.lowcnt[r] = low_count; // entries with col <= row: a prefix of the sorted row
    }
}

} // namespace

#include "build_complete.cuh"

namespace {

// ---- int64 exclusive scan of the row counts -------------------------------------------------------
constexpr int SCAN_BLOCK = 1024;

__global__ void scan_block_sums(const int *cnt, long n, long *blocksum) {
    __shared__ long ws[32];
    const long i = (long)blockIdx.x * SCAN_BLOCK + threadIdx.x;
    long v = (i < n) ? cnt[i] : 0;
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0)
        ws[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        long t = ws[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1)
            t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0)
            blocksum[blockIdx.x] = t;
    }
}

// single block: exclusive scan of blocksum[nb] in place, total -> blocksum[nb]
__global__ void scan_of_sums(long *blocksum, long nb) {
    __shared__ long ws[32];
    __shared__ long carry;
    if (threadIdx.x == 0)
        carry = 0;
    __syncthreads();
    for (long base = 0; base < nb; base += SCAN_BLOCK) {
        const long i = base + threadIdx.x;
        const long v = (i < nb) ? blocksum[i] : 0;
        long x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const long t = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o)
                x += t;
        }
        if ((threadIdx.x & 31) == 31)
            ws[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            long t = ws[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) {
                const long q = __shfl_up_sync(0xffffffffu, t, o);
                if (threadIdx.x >= o)
                    t += q;
            }
            ws[threadIdx.x] = t;
        }
        __syncthreads();
        const long woff = (threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : 0;
        const long incl = x + woff + carry;
        if (i < nb)
            blocksum[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == SCAN_BLOCK - 1)
            carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        blocksum[nb] = carry;
}

__global__ void scan_finish(const int *cnt, long n, const long *blocksum, long *indptr, int *maxcnt) {
    __shared__ long ws[32];
    const long i = (long)blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const long v = (i < n) ? cnt[i] : 0;
    long x = v;
    for (int o = 1; o < 32; o <<= 1) {
        const long t = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= o)
            x += t;
    }
    if ((threadIdx.x & 31) == 31)
        ws[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
        long t = ws[threadIdx.x];
        for (int o = 1; o < 32; o <<= 1) {
            const long q = __shfl_up_sync(0xffffffffu, t, o);
            if (threadIdx.x >= o)
                t += q;
        }
        ws[threadIdx.x] = t;
    }
    __syncthreads();
    const long woff = (threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : 0;
    if (i < n) {
        indptr[i] = blocksum[blockIdx.x] + woff + x - v;
        if (i == n - 1)
            indptr[n] = blocksum[blockIdx.x] + woff + x;
    }
    int mx = (int)v;
    for (int o = 16; o > 0; o >>= 1)
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0)
        atomicMax(maxcnt, mx);
}

__global__ void uniform_counts(int *cnt, long n, int value) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        cnt[i] = value;
}

// ---- hash index build -------------------------------------------------------------------------------

template<int KM>
__global__ void insert_kernel(typename SlotOf<KM>::type *slots, u32 mask, int shift, const u64 *dets,
                              int nwords, long ndet, u32 *bloom, u32 bmask) {
    const long idet = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idet >= ndet)
        return;
    const u64 a = dets[idet * nwords], b = (nwords == 2) ? dets[idet * nwords + 1] : 0ULL;
    DetIndex<KM> ix;
    ix.slots = slots;
    ix.mask = mask;
    ix.shift = shift;
    ix.bloom = nullptr;
    ix.bmask = 0;
    const u32 h = ix.hash(a, b);
    if (bloom)
        bloom_set(bloom, bmask, h);
    u32 p = h & mask;
    for (;;) {
        // claim an empty slot (val == -1); keys are written afterwards -- determinants are unique, so
        // no key comparison is needed while inserting (verified by verify_kernel)
        if (atomicCAS(&slots[p].val, -1, (int)idet) == -1) {
            if constexpr (KM == KEY32)
                slots[p].key = ix.key(a, b);
            else if constexpr (KM == KEY64)
                slots[p].key = ix.key(a, b);
            else {
                slots[p].k0 = a;
                slots[p].k1 = b;
            }
            return;
        }
        p = (p + 1) & mask;
    }
}

// every determinant is re-found at its own position (duplicates are not) and holds the declared number of
// electrons inside nbasis orbitals (Wfn::init / add_det preconditions); bad[0] = duplicates, bad[1] = first
// determinant with a wrong occupation (INT_MAX if none)
template<int KM>
__global__ void verify_kernel(DetIndex<KM> ix, const u64 *dets, int nwords, long ndet, u64 valid, int nocc_up,
                              int nocc_dn, int *bad) {
    const long idet = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idet >= ndet)
        return;
    const u64 a = dets[idet * nwords], b = (nwords == 2) ? dets[idet * nwords + 1] : 0ULL;
    if ((a & ~valid) || (b & ~valid) || __popcll(a) != nocc_up || (nwords == 2 && __popcll(b) != nocc_dn)) {
        atomicMin(bad + 1, (int)idet);
        return;
    }
    if (ix.find(a, b) != (int)idet)
        atomicAdd(bad, 1);
}

template<int KM>
__global__ void lookup_kernel(DetIndex<KM> ix, const u64 *dets, int nwords, long n, long *out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const u64 a = dets[i * nwords], b = (nwords == 2) ? dets[i * nwords + 1] : 0ULL;
    out[i] = ix.find(a, b);
}

// two-spin determinants: flags[0] |= 1 when (alpha, beta) does not strictly ascend (add_all_dets order,
// twospinwfn.cpp:195-218; strictly ascending also means no duplicates); flags[1] = first determinant whose strings
// do not hold the declared electrons inside nbasis orbitals (Wfn::init / add_det preconditions)
__global__ void sorted_check_kernel(const u64 *dets, long ndet, u64 valid, int nocc_up, int nocc_dn, int *flags) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ndet)
        return;
    const u64 a0 = dets[2 * i], b0 = dets[2 * i + 1];
    if ((a0 & ~valid) || (b0 & ~valid) || __popcll(a0) != nocc_up || __popcll(b0) != nocc_dn)
        atomicMin(flags + 1, (int)i);
    if (i + 1 < ndet) {
        const u64 a1 = dets[2 * i + 2], b1 = dets[2 * i + 3];
        if (!(a0 < a1 || (a0 == a1 && b0 < b1)))
            atomicOr(flags, 1);
    }
}

__global__ void linear_indptr_kernel(long *indptr, long n, long m) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n)
        indptr[i] = i * m;
}

size_t slot_bytes(int km) { return km == KEY32 ? sizeof(Slot32) : km == KEY64 ? sizeof(Slot64) : sizeof(Slot128); }

template<int KM>
int build_index_t(pyci_wfn *wfn) {
    pyci_ctx *ctx = wfn->ctx;
    typedef typename SlotOf<KM>::type slot_t;
    const long ndet = wfn->ndet;
    const int threads = 256;
    const unsigned blocks = (unsigned)((ndet + threads - 1) / threads);
    if (ndet > 0) {
        insert_kernel<KM><<<blocks, threads, 0, ctx->stream>>>(reinterpret_cast<slot_t *>(wfn->slots), wfn->mask,
                                                             (wfn->kind == PYCI_FULLCI) ? (int)wfn->nbasis : 0,
                                                             wfn->dets, wfn->nwords, ndet, wfn->bloom, wfn->bmask);
        ctx->launches++;
        int *bad = nullptr;
        const int init[2] = {0, 0x7fffffff};
        PYCI_CUDA(dev_malloc(&bad, 2 * sizeof(int)));
        PYCI_CUDA(cudaMemcpyAsync(bad, init, 2 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        const u64 valid = (wfn->nbasis >= 64) ? ~0ULL : ((1ULL << wfn->nbasis) - 1ULL);
        verify_kernel<KM><<<blocks, threads, 0, ctx->stream>>>(make_index<KM>(wfn), wfn->dets, wfn->nwords, ndet, valid,
                                                               (int)wfn->nocc_up, (int)wfn->nocc_dn, bad);
        ctx->launches++;
        int hbad[2] = {0, 0x7fffffff};
        PYCI_CUDA(cudaMemcpyAsync(hbad, bad, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        PYCI_CUDA(cudaStreamSynchronize(ctx->stream));
        dev_free(bad);
        if (hbad[1] != 0x7fffffff)
            PYCI_FAIL(PYCI_ERR_VALUE, "determinant %d does not have the declared occupation", hbad[1]);
        if (hbad[0])
            PYCI_FAIL(PYCI_ERR_VALUE, "wave function contains %d duplicate determinant(s)", hbad[0]);
    }
    return PYCI_OK;
}

double binom_d(long n, long k) {
    if (k < 0 || k > n)
        return 0.0;
    double b = 1.0;
    for (long d = 1; d <= k; ++d)
        b = b * (double)(n - k + d) / (double)d;
    return b;
}

// complete sorted two-spin space: string tables for both spins, then the segment-ordered fill (build_complete.cuh)
// one allocation per spin (T.cr heads it): fifteen stream-ordered allocations per spin cost more host time than the
// string-table kernels take on the device
int alloc_string_tables(StringTables &T, u32 N, u32 L, u32 L1, u32 nocc, u32 nS, u32 nD) {
    T.N = N;
    T.L = L;
    T.L1 = L1;
    T.nocc = nocc;
    T.nS = nS;
    T.nD = nD;
    const size_t e = (size_t)N * L, e1 = (size_t)N * L1, eS = std::max<size_t>((size_t)N * nS, 1),
                 eD = std::max<size_t>((size_t)N * nD, 1);
    const size_t sizes[15] = {4 * e, 8 * e, 8 * e1, 4 * e1, 8 * e1 * std::max<u32>(nocc, 1), 4 * (size_t)N, 4 * (size_t)N,
                              4 * eS, 4 * eS, 4 * eS, 8 * eS, 4 * eD, 4 * eD, 8 * eD, 4 * (size_t)N};
    size_t off[16];
    off[0] = 0;
    for (int q = 0; q < 15; ++q)
        off[q + 1] = off[q] + ((sizes[q] + 255) & ~(size_t)255);
    unsigned char *base = nullptr;
    PYCI_CUDA(dev_malloc(&base, off[15]));
    T.cr = reinterpret_cast<u32 *>(base + off[0]);
    T.dval = reinterpret_cast<double *>(base + off[1]);
    T.sub = reinterpret_cast<uint2 *>(base + off[2]);
    T.pos1 = reinterpret_cast<u32 *>(base + off[3]);
    T.terms = reinterpret_cast<double *>(base + off[4]);
    T.j1self = reinterpret_cast<u32 *>(base + off[5]);
    T.selfj = reinterpret_cast<u32 *>(base + off[6]);
    T.s_off = reinterpret_cast<u32 *>(base + off[7]);
    T.s_aux = reinterpret_cast<u32 *>(base + off[8]);
    T.s_cr = reinterpret_cast<u32 *>(base + off[9]);
    T.s_pre = reinterpret_cast<double *>(base + off[10]);
    T.d_off = reinterpret_cast<u32 *>(base + off[11]);
    T.d_cr = reinterpret_cast<u32 *>(base + off[12]);
    T.d_val = reinterpret_cast<double *>(base + off[13]);
    T.self_off = reinterpret_cast<u32 *>(base + off[14]);
    return PYCI_OK;
}

void free_string_tables(StringTables &T) {
    dev_free(T.cr); // heads the single allocation
    memset(&T, 0, sizeof(T));
}

int run_complete(pyci_ctx *ctx, const BuildParams &P, const SortedParams &S, size_t pair_bytes, int *used, bool *diag_done) {
    PYCI_NVTX("pyci:fill(complete: string tables + fill_complete_kernel)");
    cudaStream_t st = ctx->stream;
    *used = 0;
    const u32 Na = (u32)binom_d(P.n, P.nocc_a), Nb = S.Nb;
    const u32 L1a = 1 + P.nSa, L1b = 1 + P.nSb;
    const size_t table_bytes = ((size_t)Na * S.La + (size_t)Nb * S.Lb) * 32 + (size_t)Na * L1a * (32 + 8 * P.nocc_a) +
                               (size_t)Nb * L1b * (32 + 8 * P.nocc_b);
    // groups of 256 threads per CTA (one CTA per SM) and whether the two_mo slice of an alpha string fits beside them
    // slices of k <= l only when the integrals allow it (pyci_ham::kl_sym) -- PYCI_B200_NO_PACKED_SLICE keeps n^2
    const bool packed = P.kl_sym && !getenv("PYCI_B200_NO_PACKED_SLICE") && P.n * (P.n + 1) / 2 < 4096;
    const u32 nsl = packed ? (u32)(P.n * (P.n + 1) / 2) : (u32)(P.n * P.n);
    // threads per group: 256 (at most four groups per CTA), or 128 when shared memory has room for at least six row
    // buffers -- more rows in flight per SM (PYCI_B200_FILL_GT=128/256 overrides; the warp-specialised form has 256)
    int gt = 256;
    auto fit_gt = [&](bool sl, int gth) {
        for (int g = std::min(1024 / gth, 7); g >= 1; --g) // (two named barriers per group, sixteen per CTA)
            if ((long)complete_smem(P.nSa, P.nDa, (u32)P.n, S.M, g, sl, nsl).total <= (long)ctx->smem_optin)
                return g;
        return 0;
    };
    if (!getenv("PYCI_B200_FILL_WS")) {
        if (const char *e = getenv("PYCI_B200_FILL_GT"))
            gt = atoi(e) == 128 ? 128 : 256;
        else if (fit_gt(true, 128) >= 6 && L1b <= 128u)
            gt = 128;
    }
    auto fit = [&](bool sl) { return fit_gt(sl, gt); };
    bool with_slice = !getenv("PYCI_B200_NO_SLICE");
    int groups = with_slice ? fit(true) : 0;
    if (groups < 2) { // the slice leaves room for fewer than two row buffers: integrals through L1/L2 instead
        with_slice = false;
        groups = fit(false);
    }
    const size_t tsmem = std::max(string_table_smem(S.Wa, S.La, (u32)P.n, S.K1, pair_bytes),
                                  string_table_smem(S.Wb, S.Lb, (u32)P.n, S.K1, pair_bytes));
    if (!groups || S.La > 65535u || S.Lb > 65535u || P.n * P.n >= 4096 || S.M >= (1u << 30) ||
        table_bytes > ((size_t)4 << 30) || (long)tsmem > (long)ctx->smem_optin)
        return PYCI_OK; // the general sorted path takes it
    const size_t fsmem = complete_smem(P.nSa, P.nDa, (u32)P.n, S.M, groups, with_slice, nsl).total;
    CompleteParams C;
    memset(&C, 0, sizeof(C));
    int rc = alloc_string_tables(C.A, Na, S.La, L1a, (u32)P.nocc_a, P.nSa, P.nDa);
    if (rc == PYCI_OK)
        rc = alloc_string_tables(C.B, Nb, S.Lb, L1b, (u32)P.nocc_b, P.nSb, P.nDb);
    if (rc != PYCI_OK) {
        free_string_tables(C.A);
        free_string_tables(C.B);
        return rc;
    }
    C.M = S.M;
    C.Nb = Nb;
    C.dSb = P.dSb;
    C.nn = (u32)(P.n * P.n);
    C.nsl = nsl;
    C.packed = packed ? 1u : 0u;
    {
        const char *e = getenv("PYCI_B200_FILL_L2HINT");
        C.l2hint = (e && atoi(e) == 0) ? 0u : 1u;
    }
    C.GP = std::max(1u, (u32)gt / L1b);
    C.GPnn = C.GP * nsl;
    C.GPw = std::max(1u, 192u / L1b);
    C.GPwnn = C.GPw * nsl;
    C.dL1b = make_fastdiv(L1b);
    {
        PrepParams Q;
        Q.ga = std::min<u32>(Na, 16u * ctx->sm_count); // (a string is ~10 us of dependent phases: one CTA each)
        Q.gb = std::min<u32>(Nb, 16u * ctx->sm_count);
        Q.Wa = S.Wa, Q.Wb = S.Wb, Q.K1 = S.K1, Q.La = S.La, Q.Lb = S.Lb, Q.L1a = L1a, Q.L1b = L1b;
        Q.stride_a = (long)Nb;
        Q.binom = S.binom;
        Q.packed = (int)packed;
        const long gd = *diag_done ? 0 : (P.nloc + 127) / 128;
        PYCI_CUDA(cudaFuncSetAttribute(complete_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
        complete_prep_kernel<<<(unsigned)(Q.ga + Q.gb + gd), 128, tsmem, st>>>(P, C.A, C.B, Q);
        *diag_done = true;
        ctx->launches++;
    }
    const long grid = std::min<long>((P.nloc + groups - 1) / groups, (long)ctx->sm_count);
    PYCI_CUDA(cudaEventRecord(ctx->ev[4], st));
    // PYCI_B200_FILL_WS: the warp-specialised form of the kernel instead of the default one (every warp passes through
    // every segment); both write the same bytes.  Measured (profiles/r2c): equal at config 3 (5.10 ms both), 11 % slower
    // at config 4 on one GPU (27.8 vs 25.0 ms) -- the 192 alpha-beta threads make more trips than 256 do.
    const bool v1 = getenv("PYCI_B200_FILL_WS") == nullptr;
    void (*k)(BuildParams, CompleteParams, int);
    if (!v1)
        k = with_slice ? fill_complete_ws_kernel<true> : fill_complete_ws_kernel<false>;
    else if (gt == 128)
        k = with_slice ? fill_complete_kernel<true, 128> : fill_complete_kernel<false, 128>;
    else
        k = with_slice ? fill_complete_kernel<true, 256> : fill_complete_kernel<false, 256>;
    PYCI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
    k<<<(unsigned)grid, gt * groups, fsmem, st>>>(P, C, groups);
    PYCI_CUDA(cudaEventRecord(ctx->ev[5], st));
    ctx->launches++;
    PYCI_CUDA(cudaGetLastError());
    free_string_tables(C.A); // stream-ordered: released after the fill has run
    free_string_tables(C.B);
    *used = 1;
    return PYCI_OK;
}

template<int KIND, int KM>
int run_build(pyci_ctx *ctx, const pyci_wfn *wfn, pyci_op *op, BuildParams &P, int npairs_dim) {
    cudaStream_t st = ctx->stream;
    // The hash index of a complete sorted two-spin space is built lazily (wfn_build_index): its fill path takes the
    // columns from colex ranks and never probes.  Every other path needs it now.
    const bool may_skip_index = KIND == PYCI_FULLCI && wfn->complete && wfn->sorted2 && op->ncol == wfn->ndet &&
                                !getenv("PYCI_B200_NO_COMPLETE_PATH") && !getenv("PYCI_B200_NO_SORTED_PATH") &&
                                !getenv("PYCI_B200_FORCE_PROBE");
    if (!may_skip_index)
        PYCI_TRY(wfn_ensure_index(wfn));
    DetIndex<KM> ix = make_index<KM>(wfn);
    const long nloc = op->nloc;
    const size_t pair_bytes = pair_table_bytes(P);
    // single-excitation tables live in shared memory (two-body kinds only)
    // (enum_params_init clamps an empty beta-single list to 1 for its divisions: pass the real count, which does not
    // depend on whether alpha singles exist -- nocc_up == nbasis leaves nAB = 0 with beta singles still present)
    const u32 nSa = (KIND == PYCI_DOCI) ? 0u : P.nSa,
              nSb = (KIND == PYCI_FULLCI && P.nocc_b > 0 && P.n - P.nocc_b > 0) ? P.nSb : 0u;
    const size_t tab_bytes = tables_bytes(nSa, nSb);

    // block size from the amount of per-row work
    auto pick_block = [](long work) { return work <= 128 ? 32 : work <= 512 ? 64 : work <= 1024 ? 128 : 256; };

    PYCI_CUDA(cudaEventRecord(ctx->ev[0], st));
    auto *count_range = new PyciRange("pyci:count+scan");
    std::unique_ptr<PyciRange> count_guard(count_range);
    uint2 *hitlist = nullptr; // (candidate, column) of the hits found by the count pass, [nloc][hitcap]
    int hitcap = 0;
    int *rowcnt = nullptr;
    PYCI_CUDA(dev_malloc(&rowcnt, sizeof(int) * (size_t)(nloc + 1)));
    P.rowcnt = rowcnt;
    const bool analytic = wfn->complete && op->ncol == wfn->ndet;
    long nnz = 0;
    int maxrow = 0;
    bool joined = false; // the stored entries were found by the segment-pair join
    if (analytic) {
        // complete space: every excitation is in the wave function, so each row holds ncand + 1 entries and the row
        // pointer is a multiplication -- no count pass, no scan, nothing to read back
        linear_indptr_kernel<<<(unsigned)((nloc + 256) / 256), 256, 0, st>>>(op->indptr, nloc, (long)P.ncand + 1);
        ctx->launches++;
        nnz = nloc * ((long)P.ncand + 1);
        maxrow = nloc > 0 ? (int)P.ncand + 1 : 0;
        op->count_kernel = "analytic";
        PYCI_CUDA(cudaEventRecord(ctx->ev[1], st));
    } else {
        if (nloc > 0) {
            const int block = pick_block((long)P.ncand / 4);
            int per_sm = 1;
            const size_t csmem = table_strings_bytes(nSa, nSb) + ((pair_bytes + 7) & ~(size_t)7) + pair_mask_bytes(P, KIND);
            if ((long)csmem > (long)ctx->smem_optin)
                PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "excitation tables (%zu bytes) do not fit shared memory", csmem);
            PYCI_CUDA(cudaFuncSetAttribute(count_kernel<KIND, KM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
            PYCI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, count_kernel<KIND, KM>, block, csmem));
            const long grid = std::min<long>(nloc, (long)ctx->sm_count * std::max(per_sm, 1));
            // sparse (selected) spaces on the general fill path: record the hits of the count pass
            {
                bool sorted_path = false;
                if constexpr (KIND == PYCI_FULLCI)
                    sorted_path = wfn->sorted2 && P.nAB > 0 && binom_d(P.n, P.nocc_a) <= 65536.0 &&
                                  binom_d(P.n, P.nocc_b) <= 65536.0 && !getenv("PYCI_B200_NO_SORTED_PATH");
                size_t free_b = 0, total_b = 0;
                cudaMemGetInfo(&free_b, &total_b);
                const long budget = (long)(free_b / 4);
                long c = std::min<long>(4096, std::min<long>((long)P.ncand, budget / (8 * std::max<long>(nloc, 1))));
                long min_c = std::min<long>(32, (long)P.ncand);
                if (const char *e = getenv("PYCI_B200_HITCAP")) { // tests: a short list, so that rows overflow it
                    c = std::max(1L, std::min<long>(c, atol(e)));
                    min_c = 1;
                }
                const bool force_join = getenv("PYCI_B200_FORCE_JOIN") != nullptr;
                if (!sorted_path && (P.ncand >= 2048 || force_join) && c >= min_c && c > 0 &&
                    !getenv("PYCI_B200_NO_HITLIST")) {
                    hitcap = (int)c;
                    PYCI_CUDA(dev_malloc(&hitlist, sizeof(uint2) * (size_t)nloc * (size_t)hitcap));
                }
            }
            // Selected two-body space: find the stored entries by joining the determinant list with itself on
            // segment pairs (join.cuh) when that is predicted to take fewer bit tests than the enumeration takes
            // probes (a probe costs ~10 tests); else enumerate and probe.
            if constexpr (KIND != PYCI_DOCI) {
                // (below ~10^4 candidates per row the enumeration finishes before the join's sixty small launches do:
                // FullCI(16, 4a4b) thinned to 150 000 determinants, 3192 candidates per row: 4.7 ms against 22 ms)
                if (hitlist && !getenv("PYCI_B200_NO_JOIN") && (P.ncand >= 8192 || getenv("PYCI_B200_FORCE_JOIN"))) {
                    int used = 0;
                    const double budget = getenv("PYCI_B200_FORCE_JOIN") ? 1.0e300 : 10.0 * (double)P.ncand * (double)nloc;
                    PYCI_TRY((join_run<KIND, JOIN_HITLIST>(ctx, wfn, P, hitlist, hitcap, rowcnt, budget, &used, &op->join_tests)));
                    joined = used != 0;
                    op->joined = joined;
                    if (joined)
                        op->count_kernel = "join_rows_kernel";
                }
            }
            if (!joined) {
                op->count_kernel = "count_kernel";
                count_kernel<KIND, KM><<<(unsigned)grid, block, csmem, st>>>(P, ix, nSa, nSb, hitlist, hitcap,
                                                                        (u32)((pair_bytes + 7) & ~(size_t)7));
                ctx->launches++;
            }
        }
        // scan
        const long nb = (nloc + SCAN_BLOCK - 1) / SCAN_BLOCK;
        long *blocksum = nullptr;
        int *maxcnt = nullptr;
        PYCI_CUDA(dev_malloc(&blocksum, sizeof(long) * (size_t)(nb + 2)));
        PYCI_CUDA(dev_malloc(&maxcnt, sizeof(int)));
        PYCI_CUDA(cudaMemsetAsync(maxcnt, 0, sizeof(int), st));
        PYCI_CUDA(cudaMemsetAsync(op->indptr, 0, sizeof(long) * (size_t)(nloc + 1), st));
        if (nloc > 0) {
            scan_block_sums<<<(unsigned)nb, SCAN_BLOCK, 0, st>>>(rowcnt, nloc, blocksum);
            scan_of_sums<<<1, SCAN_BLOCK, 0, st>>>(blocksum, nb);
            scan_finish<<<(unsigned)nb, SCAN_BLOCK, 0, st>>>(rowcnt, nloc, blocksum, op->indptr, maxcnt);
            ctx->launches += 3;
        }
        PYCI_CUDA(cudaMemcpyAsync(&nnz, op->indptr + nloc, sizeof(long), cudaMemcpyDeviceToHost, st));
        PYCI_CUDA(cudaMemcpyAsync(&maxrow, maxcnt, sizeof(int), cudaMemcpyDeviceToHost, st));
        PYCI_CUDA(cudaEventRecord(ctx->ev[1], st));
        PYCI_CUDA(cudaStreamSynchronize(st));
        dev_free(blocksum);
        dev_free(maxcnt);
    }
    dev_free(rowcnt);
    P.rowcnt = nullptr;
    count_guard.reset();
    PYCI_NVTX("pyci:fill");

    op->nnz = nnz;
    // the larger array first: a pool that holds the blocks of a destroyed operator hands them back in matching sizes
    PYCI_CUDA(dev_malloc(&op->vals, sizeof(double) * (size_t)(nnz + 4))); // +4: 16-byte bulk reads may overrun the end
    PYCI_CUDA(dev_malloc(&op->cols, sizeof(int) * (size_t)(nnz + 4)));
    P.cols = op->cols;
    P.vals = op->vals;
    P.maxrow = (maxrow + 1) & ~1;
    // per-row LSD radix sort: digits of 5..8 bits covering the column index
    {
        int bits = 1;
        while ((1L << bits) < std::max<long>(op->ncol, 2))
            ++bits;
        P.sort_passes = (bits + 7) / 8;
        P.sort_dbits = std::max(5, (bits + P.sort_passes - 1) / P.sort_passes);
    }
    PYCI_CUDA(cudaEventRecord(ctx->ev[2], st)); // after the allocations: ev[2]..ev[3] brackets kernels only
    u64 *long_scratch = nullptr; // row buffers in HBM for rows too long for shared memory (freed after the fill)
    std::vector<u32> hb; // binomial table staged for the string-table pre-pass
    bool fill_timed = false;
    if (nloc > 0 && nnz > 0) {
        // H_ii of the rows: its own launch, except on the complete path, whose table launch computes it as well
        bool diag_done = false;
        auto ensure_diag = [&]() {
            if (!diag_done) {
                diag_kernel<KIND><<<(unsigned)((nloc + 127) / 128), 128, 0, st>>>(P);
                ctx->launches++;
                diag_done = true;
            }
        };
        bool done = false;
        if constexpr (KIND == PYCI_FULLCI) {
            // sorted two-spin wave function: slots in sorted order from two per-row rank lists (build_sorted.cuh)
            const double Ua = binom_d(P.n, P.nocc_a), Ub = binom_d(P.n, P.nocc_b);
            if (wfn->sorted2 && P.nocc_a > 0 && P.nocc_b > 0 && P.nvir_a > 0 && P.nvir_b > 0 && P.nAB > 0 &&
                Ua <= 65536.0 && Ub <= 65536.0 && !getenv("PYCI_B200_NO_SORTED_PATH")) {
                SortedParams S;
                S.La = 1 + P.nSa + P.nDa;
                S.Lb = 1 + P.nSb + P.nDb;
                S.Wa = ((u32)Ua + 31) / 32;
                S.Wb = ((u32)Ub + 31) / 32;
                S.K1 = (u32)std::max(P.nocc_a, P.nocc_b) + 1;
                S.M = P.ncand + 1;
                S.Nb = (u32)Ub;
                if (!ctx->binom_dev || ctx->binom_n != P.n || ctx->binom_k1 != (int)S.K1) {
                    hb.resize((size_t)P.n * S.K1);
                    for (int pp = 0; pp < P.n; ++pp)
                        for (u32 j = 0; j < S.K1; ++j)
                            hb[(size_t)pp * S.K1 + j] = (u32)std::min(binom_d(pp, j), 4294967295.0);
                    PYCI_CUDA(cudaStreamSynchronize(st)); // a build in flight may still read the table being replaced
                    if (ctx->binom_dev)
                        cudaFree(ctx->binom_dev);
                    ctx->binom_dev = nullptr;
                    PYCI_CUDA(cudaMalloc(&ctx->binom_dev, sizeof(u32) * hb.size()));
                    PYCI_CUDA(cudaMemcpyAsync(ctx->binom_dev, hb.data(), sizeof(u32) * hb.size(), cudaMemcpyHostToDevice, st));
                    PYCI_CUDA(cudaStreamSynchronize(st));
                    ctx->binom_n = P.n;
                    ctx->binom_k1 = (int)S.K1;
                }
                S.binom = ctx->binom_dev;
                const int block = pick_block((long)P.ncand / 4);
                // PYCI_B200_FORCE_PROBE: resolve columns through the hash index even in a complete space
                const bool direct = analytic && !getenv("PYCI_B200_FORCE_PROBE");
                const size_t smem = sorted_smem_bytes(S, nSa, nSb, (u32)P.n, direct, pair_bytes);
                if (direct && !getenv("PYCI_B200_NO_COMPLETE_PATH")) {
                    // complete space: per-string tables + slot-ordered fill (build_complete.cuh)
                    int used = 0;
                    PYCI_TRY(run_complete(ctx, P, S, pair_bytes, &used, &diag_done));
                    done = used != 0;
                    if (done) {
                        op->fill_kernel = "fill_complete_kernel";
                        fill_timed = true;
                    }
                }
                if (!done && (long)smem <= (long)ctx->smem_optin) {
                    ensure_diag();
                    if (!direct) {
                        PYCI_TRY(wfn_ensure_index(wfn));
                        ix = make_index<KM>(wfn);
                    }
                    int per_sm = 1;
                    long grid;
                    PYCI_CUDA(cudaEventRecord(ctx->ev[4], st));
                    fill_timed = true;
                    if (direct) {
                        PYCI_CUDA(cudaFuncSetAttribute(fill_sorted_kernel<KM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                        PYCI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fill_sorted_kernel<KM, true>, block, smem));
                        grid = std::min<long>(nloc, (long)ctx->sm_count * std::max(per_sm, 1));
                        fill_sorted_kernel<KM, true><<<(unsigned)grid, block, smem, st>>>(P, ix, S, nSa, nSb);
                    } else {
                        PYCI_CUDA(cudaFuncSetAttribute(fill_sorted_kernel<KM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                        PYCI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fill_sorted_kernel<KM, false>, block, smem));
                        grid = std::min<long>(nloc, (long)ctx->sm_count * std::max(per_sm, 1));
                        fill_sorted_kernel<KM, false><<<(unsigned)grid, block, smem, st>>>(P, ix, S, nSa, nSb);
                    }
                    PYCI_CUDA(cudaEventRecord(ctx->ev[5], st));
                    ctx->launches++;
                    done = true;
                    op->fill_kernel = "fill_sorted_kernel";
                }
            }
        }
        if (!done) {
            ensure_diag();
            PYCI_TRY(wfn_ensure_index(wfn));
            ix = make_index<KM>(wfn);
            int block = pick_block(std::max<long>((long)P.ncand / 4, maxrow));
            if (joined) // no row is enumerated by the fill pass: the threads of a CTA only share the row's hits
                block = pick_block(2 * std::max<long>(1, nnz / std::max<long>(nloc, 1)));
            // keys (ping-pong) + values + radix counters + tables + pair table
            size_t tab_used = tab_bytes; // excitation tables in shared memory (dropped below when no row is enumerated)
            auto smem_for = [&](int blk, long rows) {
                const size_t counters = std::max<size_t>(((size_t)(blk / 32) + 1) * (1u << P.sort_dbits), 2 * SORT_BUCKETS);
                return (size_t)24 * (size_t)rows + sizeof(u32) * counters + tab_used + pair_bytes;
            };
            const int full_rows = P.maxrow;
            while (block > 32 && (long)smem_for(block, full_rows) > (long)ctx->smem_optin)
                block >>= 1;
            int short_cap = full_rows; // longest row built in shared memory
            // PYCI_B200_SHORT_CAP=k: send rows of more than k entries through the HBM launch regardless (tests)
            const long forced_cap = getenv("PYCI_B200_SHORT_CAP") ? std::max(2L, atol(getenv("PYCI_B200_SHORT_CAP"))) : 0;
            const bool long_rows = (long)smem_for(block, full_rows) > (long)ctx->smem_optin ||
                                   (forced_cap > 0 && forced_cap < full_rows);
            if (long_rows) {
                // a few rows are too long for shared memory: those go through HBM in a second launch
                if (!joined) // (rows from recorded hits keep the block chosen for their mean length)
                    block = pick_block((long)P.ncand / 4);
                const long fit = ((long)ctx->smem_optin - (long)smem_for(block, 0) - 1024) / 24;
                if (fit < 64)
                    PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "excitation tables (%zu bytes) leave no room for a row buffer",
                              smem_for(block, 0));
                short_cap = (int)(std::min<long>(fit, forced_cap > 0 ? forced_cap : 4096) & ~1L);
            }
            size_t smem = smem_for(block, short_cap);
            // most candidates miss (selected space): evaluate elements for hits only
            const bool lazy = !analytic && (double)nnz < 0.25 * (double)nloc * ((double)P.ncand + 1.0);
            // rows with more hits than the recorded list holds: a second join pass stages their hits in the CSR arrays
            // (else the fill pass would enumerate and probe every candidate of those rows)
            int staged = 0;
            if constexpr (KIND != PYCI_DOCI) {
                if (joined && maxrow - 1 > hitcap) {
                    int *cursor = nullptr, used = 0;
                    PYCI_CUDA(dev_malloc(&cursor, sizeof(int) * (size_t)(nloc + 1)));
                    PYCI_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * (size_t)(nloc + 1), st));
                    const int rcj = join_run<KIND, JOIN_STAGE>(ctx, wfn, P, nullptr, hitcap, cursor, 1.0e300, &used, nullptr);
                    dev_free(cursor);
                    PYCI_TRY(rcj);
                    staged = used;
                }
            }
            // Every row comes from recorded or staged hits: the fill pass enumerates nothing, so the single-excitation
            // tables (14 KB of the 28 KB per CTA at 64 spin-orbitals / 20 electrons) stay out of shared memory and
            // twice as many rows are in flight per SM.
            u32 nSa_k = nSa, nSb_k = nSb;
            // (maxrow counts the diagonal where a row has one; rows beyond ncol of a rectangular operator do not, hence
            // the comparison without the "- 1" of the staging test above: conservative by one entry)
            if (joined && (maxrow <= hitcap || staged) && !getenv("PYCI_B200_FILL_KEEP_TABLES")) {
                tab_used = 0;
                nSa_k = nSb_k = 0;
                smem = smem_for(block, short_cap);
            }
            PYCI_CUDA(cudaEventRecord(ctx->ev[4], st));
            fill_timed = true;
            auto launch = [&](auto kern) -> int {
                int per_sm = 1;
                BuildParams Ps = P;
                Ps.maxrow = short_cap;
                PYCI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                PYCI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, smem));
                const long grid = std::min<long>(nloc, (long)ctx->sm_count * std::max(per_sm, 1));
                kern<<<(unsigned)grid, block, smem, st>>>(Ps, ix, nSa_k, nSb_k, hitlist, hitcap, nullptr, short_cap, staged);
                ctx->launches++;
                if (long_rows) {
                    const size_t smem2 = smem_for(block, 0);
                    PYCI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, smem2));
                    const long grid2 = std::min<long>(nloc, (long)ctx->sm_count * std::min(std::max(per_sm, 1), 2));
                    PYCI_CUDA(dev_malloc(&long_scratch, sizeof(u64) * 3 * (size_t)full_rows * (size_t)grid2));
                    kern<<<(unsigned)grid2, block, smem2, st>>>(P, ix, nSa_k, nSb_k, hitlist, hitcap, long_scratch, short_cap,
                                                                staged);
                    ctx->launches++;
                }
                return PYCI_OK;
            };
            if (lazy)
                PYCI_TRY(launch(fill_kernel<KIND, KM, true>));
            else
                PYCI_TRY(launch(fill_kernel<KIND, KM, false>));
            PYCI_CUDA(cudaEventRecord(ctx->ev[5], st));
            op->fill_kernel = "fill_kernel";
        }
    }
    PYCI_CUDA(cudaEventRecord(ctx->ev[3], st));
    dev_free(long_scratch);
    dev_free(hitlist);
    PYCI_CUDA(cudaStreamSynchronize(st));
    PYCI_CUDA(cudaGetLastError());
    float ms01 = 0, ms12 = 0, ms45 = 0;
    cudaEventElapsedTime(&ms01, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ms12, ctx->ev[2], ctx->ev[3]);
    if (fill_timed)
        cudaEventElapsedTime(&ms45, ctx->ev[4], ctx->ev[5]);
    op->times[0] = wfn->hash_seconds;
    op->times[1] = ms01 * 1e-3;
    op->times[2] = ms12 * 1e-3;
    op->times[3] = op->times[1] + op->times[2];
    op->fill_seconds = fill_timed ? ms45 * 1e-3 : op->times[2];
    return PYCI_OK;
}

template<int KIND>
int dispatch_key(pyci_ctx *ctx, const pyci_wfn *wfn, pyci_op *op, BuildParams &P, int npairs_dim) {
    switch (wfn->keymode) {
    case KEY32:
        return run_build<KIND, KEY32>(ctx, wfn, op, P, npairs_dim);
    case KEY64:
        return run_build<KIND, KEY64>(ctx, wfn, op, P, npairs_dim);
    default:
        if constexpr (KIND == PYCI_FULLCI)
            return run_build<KIND, KEY128>(ctx, wfn, op, P, npairs_dim);
        else
            PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "one-spin wave functions use 32- or 64-bit keys");
    }
}

__global__ void lowcnt_sum_kernel(const int *lowcnt, long n, unsigned long long *out) {
    long acc = 0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        acc += lowcnt[i];
    for (int o = 16; o > 0; o >>= 1)
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc)
        atomicAdd(out, (unsigned long long)acc);
}

} // namespace

long op_size_ref(pyci_op *op) {
    if (op->size_ref >= 0)
        return op->size_ref;
    pyci_ctx *ctx = op->ctx;
    if (ctx_activate(ctx) != PYCI_OK)
        return -1;
    unsigned long long *acc = nullptr, h = 0;
    if (dev_malloc(&acc, sizeof(unsigned long long)) != cudaSuccess)
        return -1;
    cudaMemsetAsync(acc, 0, sizeof(unsigned long long), ctx->stream);
    if (op->nloc > 0) {
        lowcnt_sum_kernel<<<ctx->sm_count, 256, 0, ctx->stream>>>(op->lowcnt, op->nloc, acc);
        ctx->launches++;
    }
    cudaMemcpyAsync(&h, acc, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    dev_free(acc);
    if (e != cudaSuccess)
        return -1;
    op->size_ref = (long)h;
    return op->size_ref;
}

// indptr[0..n] = exclusive int64 scan of cnt[0..n) (indptr[n] = total); *maxcnt (device, may be null) = max cnt
int scan_counts(pyci_ctx *ctx, const int *cnt, long n, long *indptr, int *maxcnt) {
    cudaStream_t st = ctx->stream;
    const long nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    long *blocksum = nullptr;
    int *mc = maxcnt;
    PYCI_CUDA(dev_malloc(&blocksum, sizeof(long) * (size_t)(nb + 2)));
    if (!mc)
        PYCI_CUDA(dev_malloc(&mc, sizeof(int)));
    PYCI_CUDA(cudaMemsetAsync(mc, 0, sizeof(int), st));
    PYCI_CUDA(cudaMemsetAsync(indptr, 0, sizeof(long) * (size_t)(n + 1), st));
    if (n > 0) {
        scan_block_sums<<<(unsigned)nb, SCAN_BLOCK, 0, st>>>(cnt, n, blocksum);
        scan_of_sums<<<1, SCAN_BLOCK, 0, st>>>(blocksum, nb);
        scan_finish<<<(unsigned)nb, SCAN_BLOCK, 0, st>>>(cnt, n, blocksum, indptr, mc);
        ctx->launches += 3;
    }
    PYCI_CUDA(cudaGetLastError());
    dev_free(blocksum);
    if (!maxcnt)
        dev_free(mc);
    return PYCI_OK;
}

// The hash index itself (slots + Bloom filter): insert, verify.
static int wfn_build_hash(pyci_wfn *wfn) {
    pyci_ctx *ctx = wfn->ctx;
    PYCI_NVTX("pyci:index(hash insert+verify)");
    // capacity: power of two with load factor in (0.25, 0.5]
    u64 cap = 16;
    while (cap < 2 * (u64)wfn->ndet)
        cap <<= 1;
    if (cap > (1ULL << 31))
        PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "too many determinants for the device index (%ld)", wfn->ndet);
    dev_free(wfn->slots);
    wfn->slots = nullptr;
    wfn->index_valid = false;
    wfn->mask = (u32)(cap - 1);
    const size_t bytes = slot_bytes(wfn->keymode) * (size_t)cap;
    PYCI_CUDA(dev_malloc(&wfn->slots, bytes));
    PYCI_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    PYCI_CUDA(cudaMemsetAsync(wfn->slots, 0xFF, bytes, ctx->stream));
    // Selected (incomplete) spaces: nearly every probe of the enumeration misses.  A blocked Bloom filter in front of
    // the slots -- 8-32 bits per determinant, false positives <= 0.4 % -- answers a miss with one 4-byte load from a
    // table 16-64x smaller than the slots: it stays L1/L2-resident when the slot table is L2-resident (count pass of
    // a 200 000-determinant GenCI space 334 -> 164 ms) and L2-resident when the slots are in HBM (5 M determinants:
    // 16.1 -> 4.0 s).  Complete spaces never miss and skip it.  PYCI_B200_BLOOM=0/1 forces it off / on.
    dev_free(wfn->bloom);
    wfn->bloom = nullptr;
    wfn->bmask = 0;
    {
        const char *e = getenv("PYCI_B200_BLOOM");
        const bool want = e ? atoi(e) != 0 : (!wfn->complete && wfn->ndet >= 4096);
        if (want && wfn->ndet > 0) {
            u64 words = 1024;
            const u64 per = getenv("PYCI_B200_BLOOM_DIV") ? (u64)atoi(getenv("PYCI_B200_BLOOM_DIV")) : 2;
            while (words < (u64)wfn->ndet / per) // >= 16 bits per determinant ...
                words <<= 1;
            while (words * 4 > ((u64)64 << 20) && words / 2 >= (u64)wfn->ndet / 4) // ... down to 8 above 64 MB
                words >>= 1;
            PYCI_CUDA(dev_malloc(&wfn->bloom, sizeof(u32) * (size_t)words));
            PYCI_CUDA(cudaMemsetAsync(wfn->bloom, 0, sizeof(u32) * (size_t)words, ctx->stream));
            wfn->bmask = (u32)(words - 1);
        }
    }
    int rc;
    switch (wfn->keymode) {
    case KEY32:
        rc = build_index_t<KEY32>(wfn);
        break;
    case KEY64:
        rc = build_index_t<KEY64>(wfn);
        break;
    default:
        rc = build_index_t<KEY128>(wfn);
        break;
    }
    PYCI_TRY(rc);
    PYCI_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    PYCI_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    wfn->hash_seconds += ms * 1e-3;
    wfn->index_valid = true;
    return PYCI_OK;
}

int wfn_ensure_index(const pyci_wfn *wfn) {
    if (wfn->index_valid)
        return PYCI_OK;
    if (wfn->keymode == KEY_MW)
        return mw_index_build(const_cast<pyci_wfn *>(wfn));
    return wfn_build_hash(const_cast<pyci_wfn *>(wfn)); // a cache: logically const
}

// Validation + index of a wave function whose determinants are resident.  Two-spin determinants are first checked for
// add_all_dets order and occupations in one pass; a COMPLETE space in that order needs nothing else -- strictly
// ascending strings are unique, and its construction path (build_complete.cuh) derives columns from colex ranks --
// so its hash index is deferred until something probes it (wfn_ensure_index: index_det, RDMs, add_hci, the general
// fill paths).  Every other wave function gets its hash index here (insert + verify: duplicates, occupations).
int wfn_build_index(pyci_wfn *wfn) {
    pyci_ctx *ctx = wfn->ctx;
    PYCI_NVTX("pyci:index(check)");
    if (wfn->keymode == KEY_MW) { // multi-word strings: index of determinant positions (multiword.cu)
        wfn->sorted2 = false;
        wfn->hash_seconds = 0.0;
        return mw_index_build(wfn);
    }
    if (wfn->generated && wfn->kind == PYCI_FULLCI && wfn->complete && !getenv("PYCI_B200_EAGER_INDEX")) {
        // unranked on the device in add_all_dets order (pyci_wfn_create_all_dets): sorted, valid and complete by
        // construction -- nothing to check, and the hash index stays deferred (or valid, if something built it)
        wfn->sorted2 = true;
        wfn->hash_seconds = 0.0;
        return PYCI_OK;
    }
    dev_free(wfn->slots);
    wfn->slots = nullptr;
    wfn->index_valid = false;
    wfn->sorted2 = false;
    wfn->hash_seconds = 0.0;
    if (wfn->kind == PYCI_FULLCI && wfn->ndet > 0) {
        int *flags = nullptr, h[2] = {1, 0x7fffffff};
        const int init[2] = {0, 0x7fffffff};
        PYCI_CUDA(dev_malloc(&flags, 2 * sizeof(int)));
        PYCI_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
        PYCI_CUDA(cudaMemcpyAsync(flags, init, 2 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        const u64 valid = (wfn->nbasis >= 64) ? ~0ULL : ((1ULL << wfn->nbasis) - 1ULL);
        sorted_check_kernel<<<(unsigned)((wfn->ndet + 255) / 256), 256, 0, ctx->stream>>>(
            wfn->dets, wfn->ndet, valid, (int)wfn->nocc_up, (int)wfn->nocc_dn, flags);
        ctx->launches++;
        PYCI_CUDA(cudaMemcpyAsync(h, flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        PYCI_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
        PYCI_CUDA(cudaStreamSynchronize(ctx->stream));
        dev_free(flags);
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
        wfn->hash_seconds = ms * 1e-3;
        if (h[1] != 0x7fffffff)
            PYCI_FAIL(PYCI_ERR_VALUE, "determinant %d does not have the declared occupation", h[1]);
        wfn->sorted2 = (h[0] == 0);
        if (wfn->complete && wfn->sorted2 && !getenv("PYCI_B200_EAGER_INDEX"))
            return PYCI_OK;
    }
    return wfn_build_hash(wfn);
}

int wfn_index_dets_impl(pyci_wfn *wfn, long n, const u64 *dets_dev, long *out_dev) {
    pyci_ctx *ctx = wfn->ctx;
    if (n <= 0)
        return PYCI_OK;
    PYCI_TRY(wfn_ensure_index(wfn));
    const unsigned blocks = (unsigned)((n + 255) / 256);
    switch (wfn->keymode) {
    case KEY32:
        lookup_kernel<KEY32><<<blocks, 256, 0, ctx->stream>>>(make_index<KEY32>(wfn), dets_dev, wfn->nwords, n, out_dev);
        break;
    case KEY64:
        lookup_kernel<KEY64><<<blocks, 256, 0, ctx->stream>>>(make_index<KEY64>(wfn), dets_dev, wfn->nwords, n, out_dev);
        break;
    default:
        lookup_kernel<KEY128><<<blocks, 256, 0, ctx->stream>>>(make_index<KEY128>(wfn), dets_dev, wfn->nwords, n, out_dev);
        break;
    }
    ctx->launches++;
    PYCI_CUDA(cudaGetLastError());
    return PYCI_OK;
}

int op_build_impl(pyci_ctx *ctx, const pyci_ham *ham, const pyci_wfn *wfn, pyci_op *op) {
    if (wfn->keymode == KEY_MW) {
        PYCI_TRY(wfn_ensure_index(wfn));
        return mw_op_build(ctx, ham, wfn, op);
    }
    BuildParams P;
    PYCI_TRY(enum_params_init(P, wfn));
    const int kind = wfn->kind;
    P.row0 = op->row0;
    P.nloc = op->nloc;
    P.ncol = op->ncol;
    P.one_mo = ham->one_mo;
    P.two_mo = ham->two_mo;
    P.kl_sym = ham->kl_sym ? 1 : 0;
    P.h = ham->h;
    P.v = ham->v;
    P.w = ham->w;
    P.indptr = op->indptr;
    P.lowcnt = op->lowcnt;
    P.diag = op->diag;
    const int npairs_dim = P.npairs_dim;

    int rc;
    if (kind == PYCI_DOCI)
        rc = dispatch_key<PYCI_DOCI>(ctx, wfn, op, P, npairs_dim);
    else if (kind == PYCI_FULLCI)
        rc = dispatch_key<PYCI_FULLCI>(ctx, wfn, op, P, npairs_dim);
    else
        rc = dispatch_key<PYCI_GENCI>(ctx, wfn, op, P, npairs_dim);
    PYCI_TRY(rc);

    // SparseOp::size in the reference's storage: summed from lowcnt on first use (op_size_ref)
    op->size_ref = op->symmetric ? -1 : op->nnz;
    return PYCI_OK;
}
