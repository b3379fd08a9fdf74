// Developer probe: what write bandwidth does the CSR output layout allow?  (not part of the product)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/write_bw.bin tools/write_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_flat(double2 *v, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        v[i] = make_double2(1.0, 2.0);
}
// MODE 0: cols+vals scalar, 1: vals only scalar, 2: cols only scalar
template<int MODE>
__global__ void k_rows(int *cols, double *vals, long nrow, int M) {
    const long per = (nrow + gridDim.x - 1) / gridDim.x;
    const long r0 = blockIdx.x * per, r1 = min(nrow, r0 + per);
    for (long r = r0; r < r1; ++r) {
        int *c = cols + r * M;
        double *v = vals + r * M;
        for (int e = threadIdx.x; e < M; e += blockDim.x) {
            if (MODE != 1) c[e] = e;
            if (MODE != 2) v[e] = (double)e;
        }
    }
}
// per-CTA contiguous flat range, aligned 16B stores
__global__ void k_cta_flat(int *cols, double *vals, long nrow, int M) {
    const long per = (nrow + gridDim.x - 1) / gridDim.x;
    const long r0 = blockIdx.x * per, r1 = min(nrow, r0 + per);
    if (r0 >= r1) return;
    long b = r0 * M, e = r1 * M;
    long b4 = (b + 3) & ~3L, e4 = e & ~3L;
    for (long i = b + threadIdx.x; i < b4; i += blockDim.x) { cols[i] = 1; vals[i] = 1.0; }
    for (long i = e4 + threadIdx.x; i < e; i += blockDim.x) { cols[i] = 1; vals[i] = 1.0; }
    for (long i = b4 / 4 + threadIdx.x; i < e4 / 4; i += blockDim.x) {
        reinterpret_cast<int4 *>(cols)[i] = make_int4(1, 2, 3, 4);
        reinterpret_cast<double2 *>(vals)[2 * i] = make_double2(1.0, 2.0);
        reinterpret_cast<double2 *>(vals)[2 * i + 1] = make_double2(1.0, 2.0);
    }
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// smem staging per row, then TMA bulk store (STAGE=1) or 16B vector stores from smem (STAGE=2); double buffered
template<int STAGE>
__global__ void k_staged(int *cols, double *vals, long nrow, int M) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int MP = (M + 8) & ~3;
    double *sval[2] = {reinterpret_cast<double *>(sm), reinterpret_cast<double *>(sm) + MP};
    int *scol[2] = {reinterpret_cast<int *>(sm + 16 * MP), reinterpret_cast<int *>(sm + 16 * MP) + MP};
    const long per = (nrow + gridDim.x - 1) / gridDim.x;
    const long r0 = blockIdx.x * per, r1 = min(nrow, r0 + per);
    int buf = 0;
    for (long r = r0; r < r1; ++r, buf ^= 1) {
        const long out0 = r * M;
        const int ov = (int)(out0 & 1), oc = (int)(out0 & 3);
        if (STAGE == 1) {
            // the bulk stores that read this buffer two rows ago must have finished reading
            if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncthreads();
        }
        for (int e = threadIdx.x; e < M; e += blockDim.x) {
            sval[buf][ov + e] = (double)e;
            scol[buf][oc + e] = e;
        }
        if (STAGE == 1) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            const int hv = ov ? 1 : 0, nv = (M - hv) & ~1;
            const int hc = (4 - oc) & 3, nc = (M - hc) & ~3;
            if (threadIdx.x == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(vals + out0 + hv),
                             "r"(smem_u32(sval[buf] + ov + hv)), "r"(nv * 8) : "memory");
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(cols + out0 + hc),
                             "r"(smem_u32(scol[buf] + oc + hc)), "r"(nc * 4) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            // heads and tails by scalar stores
            if (threadIdx.x < 32) {
                const int e = threadIdx.x;
                if (e < hv) vals[out0 + e] = sval[buf][ov + e];
                if (e >= 8 && e - 8 + hv + nv < M) vals[out0 + hv + nv + e - 8] = sval[buf][ov + hv + nv + e - 8];
                if (e >= 16 && e - 16 < hc) cols[out0 + e - 16] = scol[buf][oc + e - 16];
                if (e >= 24 && e - 24 + hc + nc < M) cols[out0 + hc + nc + e - 24] = scol[buf][oc + hc + nc + e - 24];
            }
        } else {
            __syncthreads();
            const int hv = ov ? 1 : 0, nv = (M - hv) >> 1;
            const int hc = (4 - oc) & 3, nc = (M - hc) >> 2;
            double2 *gv = reinterpret_cast<double2 *>(vals + out0 + hv);
            const double2 *sv = reinterpret_cast<const double2 *>(sval[buf] + ov + hv);
            for (int i = threadIdx.x; i < nv; i += blockDim.x) gv[i] = sv[i];
            int4 *gc = reinterpret_cast<int4 *>(cols + out0 + hc);
            const int4 *sc = reinterpret_cast<const int4 *>(scol[buf] + oc + hc);
            for (int i = threadIdx.x; i < nc; i += blockDim.x) gc[i] = sc[i];
            if (threadIdx.x < 32) {
                const int e = threadIdx.x;
                if (e < hv) vals[out0 + e] = sval[buf][ov + e];
                if (e >= 8 && e - 8 + hv + 2 * nv < M) vals[out0 + hv + 2 * nv + e - 8] = sval[buf][ov + hv + 2 * nv + e - 8];
                if (e >= 16 && e - 16 < hc) cols[out0 + e - 16] = scol[buf][oc + e - 16];
                if (e >= 24 && e - 24 + hc + 4 * nc < M) cols[out0 + hc + 4 * nc + e - 24] = scol[buf][oc + hc + 4 * nc + e - 24];
            }
        }
    }
    if (STAGE == 1 && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// variants of the TMA-staged row store: bulk stores cut into CHUNK-byte pieces (0 = one per array), an L2 evict-first
// hint, and a main body that starts and ends on 128-byte lines of global memory (ALIGN = 128; 16 = as the product does)
template<int CHUNK, int HINT, int ALIGN>
__global__ void k_staged_v(int *cols, double *vals, long nrow, int M) {
    extern __shared__ __align__(16) unsigned char sm[]; // (the base of dynamic shared memory is 1 KB aligned here: no static shared memory)
    const int MP = (M + 32 + 31) & ~31;
    double *sval[2] = {reinterpret_cast<double *>(sm), reinterpret_cast<double *>(sm) + MP};
    int *scol[2] = {reinterpret_cast<int *>(sm + 16 * MP), reinterpret_cast<int *>(sm + 16 * MP) + MP};
    const long per = (nrow + gridDim.x - 1) / gridDim.x;
    const long r0 = blockIdx.x * per, r1 = min(nrow, r0 + per);
    unsigned long long pol = 0;
    if (HINT) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    int buf = 0;
    const int AV = ALIGN / 8, AC = ALIGN / 4; // elements per alignment unit
    for (long r = r0; r < r1; ++r, buf ^= 1) {
        const long out0 = r * M;
        // shared-memory images keep the global address modulo ALIGN
        const int ov = (int)(out0 % AV), oc = (int)(out0 % AC);
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
        for (int e = threadIdx.x; e < M; e += blockDim.x) {
            sval[buf][ov + e] = (double)e;
            scol[buf][oc + e] = e;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        const int hv = (AV - ov) % AV, nv = (M - hv) / AV * AV;
        const int hc = (AC - oc) % AC, nc = (M - hc) / AC * AC;
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            const int cv = CHUNK ? CHUNK : nv * 8, cc = CHUNK ? CHUNK : nc * 4;
            const int nchv = (nv * 8 + cv - 1) / cv, nchc = (nc * 4 + cc - 1) / cc;
            for (int q = lane; q < nchv + nchc; q += 32) {
                const bool isv = q < nchv;
                const int k = isv ? q : q - nchv;
                const int off = k * (isv ? cv : cc);
                const int len = min(isv ? cv : cc, (isv ? nv * 8 : nc * 4) - off);
                const char *g = isv ? (const char *)(vals + out0 + hv) + off : (const char *)(cols + out0 + hc) + off;
                const unsigned sa = isv ? smem_u32(sval[buf] + ov + hv) + off : smem_u32(scol[buf] + oc + hc) + off;
                if (HINT)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(g),
                                 "r"(sa), "r"(len), "l"(pol) : "memory");
                else
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(sa), "r"(len) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // heads and tails by scalar stores (up to ALIGN bytes each)
            for (int e = lane; e < hv; e += 32) vals[out0 + e] = sval[buf][ov + e];
            for (int e = hv + nv + lane; e < M; e += 32) vals[out0 + e] = sval[buf][ov + e];
            for (int e = lane; e < hc; e += 32) cols[out0 + e] = scol[buf][oc + e];
            for (int e = hc + nc + lane; e < M; e += 32) cols[out0 + e] = scol[buf][oc + e];
        }
    }
    if (threadIdx.x < 32) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
template<int CHUNK, int HINT, int ALIGN>
void run_v(const char *label, int *cols, double *vals, long nrow, int M, cudaEvent_t a, cudaEvent_t b) {
    const int MP = (M + 32 + 31) & ~31;
    const size_t smem = 24 * (size_t)MP;
    cudaFuncSetAttribute(k_staged_v<CHUNK, HINT, ALIGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int g = 3; g <= 4; ++g) {
        float ms;
        cudaEventRecord(a);
        k_staged_v<CHUNK, HINT, ALIGN><<<148 * g, 256, smem>>>(cols, vals, nrow, M);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        printf("%-40s x%d %8.3f ms %8.1f GB/s  %s\n", label, g, ms, (double)nrow * M * 12 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
}
#define TIME(label, bytes, ...)                                                              \
    do {                                                                                     \
        cudaEventRecord(a); __VA_ARGS__; cudaEventRecord(b); cudaEventSynchronize(b);         \
        cudaEventElapsedTime(&ms, a, b);                                                     \
        printf("%-44s %8.3f ms %8.1f GB/s  %s\n", label, ms, (bytes) / ms / 1e6, cudaGetErrorString(cudaGetLastError())); \
    } while (0)
int main() {
    const long nrow = 1002001; const int M = 2221;
    const size_t nnz = (size_t)nrow * M;
    int *cols; double *vals;
    cudaMalloc(&cols, nnz * 4 + 64); cudaMalloc(&vals, nnz * 8 + 64);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float ms;
    const int MP = (M + 8) & ~3; const size_t smem = 24 * (size_t)MP;
    cudaFuncSetAttribute(k_staged<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_staged<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) {
        TIME("flat 16B grid-stride (vals)", nnz * 8, (k_flat<<<148 * 8, 256>>>((double2 *)vals, nnz / 2)));
        TIME("rows scalar cols+vals x4", nnz * 12, (k_rows<0><<<148 * 4, 256>>>(cols, vals, nrow, M)));
        TIME("rows scalar vals only x4", nnz * 8, (k_rows<1><<<148 * 4, 256>>>(cols, vals, nrow, M)));
        TIME("rows scalar cols only x4", nnz * 4, (k_rows<2><<<148 * 4, 256>>>(cols, vals, nrow, M)));
        TIME("rows scalar cols+vals x2", nnz * 12, (k_rows<0><<<148 * 2, 256>>>(cols, vals, nrow, M)));
        TIME("rows scalar cols+vals x1 (1024 thr)", nnz * 12, (k_rows<0><<<148, 1024>>>(cols, vals, nrow, M)));
        for (int g = 2; g <= 8; g *= 2) {
            char l[64];
            snprintf(l, 64, "per-CTA flat range, 16B stores x%d", g);
            TIME(l, nnz * 12, (k_cta_flat<<<148 * g, 256>>>(cols, vals, nrow, M)));
        }
        for (int g = 1; g <= 4; ++g) {
            char l[64];
            snprintf(l, 64, "smem staged + TMA bulk store x%d", g);
            TIME(l, nnz * 12, (k_staged<1><<<148 * g, 256, smem>>>(cols, vals, nrow, M)));
            snprintf(l, 64, "smem staged + 16B st.global x%d", g);
            TIME(l, nnz * 12, (k_staged<2><<<148 * g, 256, smem>>>(cols, vals, nrow, M)));
        }
    }
    for (int rep = 0; rep < 2; ++rep) {
        run_v<0, 0, 16>("TMA whole arrays, 16 B aligned", cols, vals, nrow, M, a, b);
        run_v<0, 1, 16>("TMA whole, evict-first hint", cols, vals, nrow, M, a, b);
        run_v<2048, 0, 16>("TMA 2 KB pieces", cols, vals, nrow, M, a, b);
        run_v<4096, 0, 16>("TMA 4 KB pieces", cols, vals, nrow, M, a, b);
        run_v<4096, 1, 16>("TMA 4 KB pieces, evict-first hint", cols, vals, nrow, M, a, b);
        // (main bodies on 128-byte lines of global memory: 5390 / 5820 GB/s at x3 / x4, below the 16-byte form)
    }
    // check the staged TMA variant wrote what the scalar variant writes
    k_rows<0><<<148 * 4, 256>>>(cols, vals, nrow, M);
    double *ref = new double[4 * M]; int *refc = new int[4 * M];
    double *got = new double[4 * M]; int *gotc = new int[4 * M];
    const long off = (nrow - 4) * M;
    cudaMemcpy(ref, vals + off, 4 * M * 8, cudaMemcpyDeviceToHost); cudaMemcpy(refc, cols + off, 4 * M * 4, cudaMemcpyDeviceToHost);
    cudaMemset(vals + off, 0, 4 * M * 8); cudaMemset(cols + off, 0, 4 * M * 4);
    k_staged<1><<<148 * 2, 256, smem>>>(cols, vals, nrow, M);
    cudaMemcpy(got, vals + off, 4 * M * 8, cudaMemcpyDeviceToHost); cudaMemcpy(gotc, cols + off, 4 * M * 4, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int i = 0; i < 4 * M; ++i) bad += (ref[i] != got[i]) + (refc[i] != gotc[i]);
    printf("TMA staged mismatches: %ld  %s\n", bad, cudaGetErrorString(cudaDeviceSynchronize()));
    {
        const int MP2 = (M + 32 + 31) & ~31;
        cudaMemset(vals + off, 0, 4 * M * 8); cudaMemset(cols + off, 0, 4 * M * 4);
        k_staged_v<4096, 1, 16><<<148 * 3, 256, 24 * (size_t)MP2>>>(cols, vals, nrow, M);
        cudaMemcpy(got, vals + off, 4 * M * 8, cudaMemcpyDeviceToHost); cudaMemcpy(gotc, cols + off, 4 * M * 4, cudaMemcpyDeviceToHost);
        bad = 0;
        for (int i = 0; i < 4 * M; ++i) bad += (ref[i] != got[i]) + (refc[i] != gotc[i]);
        printf("TMA pieces + hint mismatches: %ld  %s\n", bad, cudaGetErrorString(cudaDeviceSynchronize()));
    }
    return 0;
}
