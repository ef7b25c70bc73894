// Column-major FP64 building blocks for the small (p x p, r x r, r x k) algebra of the filter.
// These replace the Eigen dense products of the reference (filter.cpp:239-250, 275, 292-296, 327).
// They are deliberately plain CUDA-core kernels: the matrices are at most p x p (p <= a few
// thousand) and FP64; B200 has no FP64 tcgen05 path, and the N-scaled work lives elsewhere.
#include "common.cuh"

#include <algorithm>
#include <mutex>
#include <set>
#include <utility>

namespace nle {

thread_local long long g_launches = 0;

static void pool_init_once() {
    static thread_local int inited_dev = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    if (inited_dev == dev) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ULL;   // keep freed blocks cached in the pool across synchronisations
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    inited_dev = dev;
}

void* pool_alloc(size_t bytes) {
    pool_init_once();
    void* p = nullptr;
    NLE_CUDA(cudaMallocAsync(&p, bytes, (cudaStream_t) nullptr));
    return p;
}

void pool_free(void* p) {
    if (p) cudaFreeAsync(p, (cudaStream_t) nullptr);
}

void* Arena::alloc(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    total_used += bytes;
    while (cur < chunks.size()) {
        if (off + bytes <= chunks[cur].cap) {
            void* r = chunks[cur].p + off;
            off += bytes;
            return r;
        }
        ++cur;
        off = 0;
    }
    size_t cap = bytes > ((size_t)64 << 20) ? bytes : ((size_t)64 << 20);
    Chunk c{static_cast<char*>(pool_alloc(cap)), cap};
    chunks.push_back(c);
    cur = chunks.size() - 1;
    off = bytes;
    return c.p;
}

void Arena::release_all() {
    for (auto& c : chunks) pool_free(c.p);
    chunks.clear();
    cur = off = 0;
}

void Arena::reset() {
    // the previous call has synchronised its stream before returning, nothing is in flight
    if (chunks.size() > 1) {
        size_t want = total_used + (total_used >> 3);
        release_all();
        Chunk c{static_cast<char*>(pool_alloc(want)), want};
        chunks.push_back(c);
    }
    cur = off = total_used = 0;
}

Arena& thread_arena() {
    static thread_local Arena a;
    return a;
}

int sm_count() {
    static int cached[64] = {0};      // per device (a process may drive several GPUs, one per thread)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    int n = cached[dev];
    if (!n) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
        cached[dev] = n;
    }
    return n;
}

int smem_optin_max() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    int v = cached[dev];
    if (!v) {
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (v <= 0) v = 227 * 1024;
        cached[dev] = v;
    }
    return v;
}

void allow_max_dynamic_smem(const void* kernel) {
    static std::mutex mu;
    static std::set<std::pair<const void*, int>> done;
    int dev = 0;
    NLE_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (done.count({kernel, dev})) return;
    cudaFuncAttributes fa;
    NLE_CUDA(cudaFuncGetAttributes(&fa, kernel));
    NLE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin_max() - (int)fa.sharedSizeBytes));
    done.insert({kernel, dev});
}

// ------------------------------------------------------------------------------------------
// C(m x n) = alpha * op(A) * op(B) + beta * C, column-major, on the FP64 tensor pipe (mma.sync.m8n8k4.f64, SASS DMMA).
// CTA tile 64 x 64 x 16, 4 warps (2 x 2), warp tile 32 x 32 = 4 x 4 DMMA tiles (32 accumulators per lane); operand tiles
// are stored k-major in shared memory with a padded row (GLDS % 16 == 8: the four k-rows of a fragment fall into disjoint
// bank groups) and the next chunk is prefetched into registers while the current one is multiplied.
constexpr int GBM = 64, GBN = 64, GBK = 16, GLDS = 64 + 8;

template <bool TA, bool TB>
__global__ void __launch_bounds__(128)
dgemm_kernel(int M, int N, int Kfull, double alpha, const double* __restrict__ A, int lda,
             const double* __restrict__ B, int ldb, double beta, double* __restrict__ C, int ldc, int kchunk, size_t cz) {
    // split-K (dgemm_splitk): grid.z slices of kchunk; slice z writes its partial product to C + z * cz
    const int kbeg = blockIdx.z * kchunk;
    const int K = min(Kfull, kbeg + kchunk);
    C += (size_t)blockIdx.z * cz;
    __shared__ double As[GBK][GLDS];     // As[k][i] = op(A)[i0+i][k0+k]
    __shared__ double Bs[GBK][GLDS];     // Bs[k][j] = op(B)[k0+k][j0+j]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tq = lane & 3;
    const int wm = warp & 1, wn = warp >> 1;
    const int i0 = blockIdx.x * GBM, j0 = blockIdx.y * GBN;
    double acc[4][4][2];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v][0] = acc[u][v][1] = 0.0;
    double pa[8], pb[8];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            const int idx = tid + l * 128;
            int i, kk;
            if (TA) { kk = idx & 15; i = idx >> 4; } else { i = idx & 63; kk = idx >> 6; }
            const int gi = i0 + i, gk = k0 + kk;
            pa[l] = (gi < M && gk < K) ? (TA ? A[gk + (size_t)gi * lda] : A[gi + (size_t)gk * lda]) : 0.0;
            int j, kb;
            if (TB) { j = idx & 63; kb = idx >> 6; } else { kb = idx & 15; j = idx >> 4; }
            const int gj = j0 + j, gkb = k0 + kb;
            pb[l] = (gj < N && gkb < K) ? (TB ? B[gj + (size_t)gkb * ldb] : B[gkb + (size_t)gj * ldb]) : 0.0;
        }
    };
    auto commit = [&]() {
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            const int idx = tid + l * 128;
            int i, kk;
            if (TA) { kk = idx & 15; i = idx >> 4; } else { i = idx & 63; kk = idx >> 6; }
            As[kk][i] = pa[l];
            int j, kb;
            if (TB) { j = idx & 63; kb = idx >> 6; } else { kb = idx & 15; j = idx >> 4; }
            Bs[kb][j] = pb[l];
        }
    };
    fetch(kbeg);
    for (int k0 = kbeg; k0 < K; k0 += GBK) {
        __syncthreads();                   // previous chunk consumed
        commit();
        __syncthreads();
        if (k0 + GBK < K) fetch(k0 + GBK); // in flight during the DMMAs below
#pragma unroll
        for (int k4 = 0; k4 < GBK / 4; ++k4) {
            double a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] = As[k4 * 4 + tq][wm * 32 + u * 8 + g];
#pragma unroll
            for (int v = 0; v < 4; ++v) b[v] = Bs[k4 * 4 + tq][wn * 32 + v * 8 + g];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) dmma884(acc[u][v][0], acc[u][v][1], a[u], b[v]);
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int gi = i0 + wm * 32 + u * 8 + g;
        if (gi >= M) continue;
#pragma unroll
        for (int v = 0; v < 4; ++v)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gj = j0 + wn * 32 + v * 8 + 2 * tq + e;
                if (gj >= N) continue;
                double* c = C + gi + (size_t)gj * ldc;
                double r = alpha * acc[u][v][e];
                if (beta != 0.0) r = fma(beta, *c, r);
                *c = r;
            }
    }
}

void dgemm(bool transA, bool transB, int m, int n, int k, double alpha, const double* A, int lda,
           const double* B, int ldb, double beta, double* C, int ldc, cudaStream_t s) {
    if (m <= 0 || n <= 0) return;
    dim3 grid(cdiv(m, GBM), cdiv(n, GBN));
    if (!transA && !transB) dgemm_kernel<false, false><<<grid, 128, 0, s>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, k, 0);
    else if (transA && !transB) dgemm_kernel<true, false><<<grid, 128, 0, s>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, k, 0);
    else if (!transA && transB) dgemm_kernel<false, true><<<grid, 128, 0, s>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, k, 0);
    else dgemm_kernel<true, true><<<grid, 128, 0, s>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, k, 0);
    NLE_LAUNCH_CHECK();
}

// Split-K product for skinny outputs (the block eigensolver of eig_topk.cu: m x n is n x 64 or 64 x 64 with k in the
// thousands, i.e. a handful of 64 x 64 tiles that would otherwise run on a handful of SMs): slice z of the k range writes
// its partial product (m x n, ld m) to part + z*m*n; dgemm_combine sums the slices in index order (deterministic) and
// applies the epilogue  C = alpha * sum_z part_z + beta * P + gamma * Z.
int dgemm_splitk_slices(int m, int n, int k) {
    const int tiles = cdiv(m, GBM) * cdiv(n, GBN);
    int want = std::max(1, (2 * sm_count()) / tiles);            // ~two CTAs per SM
    int kchunk = std::max(GBK * 2, cdiv(cdiv(k, want), GBK) * GBK);
    return cdiv(k, kchunk);
}

void dgemm_splitk(bool transA, int m, int n, int k, const double* A, int lda, const double* B, int ldb, double* part,
                  cudaStream_t s) {
    if (m <= 0 || n <= 0) return;
    const int nz = dgemm_splitk_slices(m, n, k);
    const int kchunk = cdiv(cdiv(k, nz), GBK) * GBK;
    dim3 grid(cdiv(m, GBM), cdiv(n, GBN), cdiv(k, kchunk));
    if (!transA) dgemm_kernel<false, false><<<grid, 128, 0, s>>>(m, n, k, 1.0, A, lda, B, ldb, 0.0, part, m, kchunk, (size_t)m * n);
    else dgemm_kernel<true, false><<<grid, 128, 0, s>>>(m, n, k, 1.0, A, lda, B, ldb, 0.0, part, m, kchunk, (size_t)m * n);
    NLE_LAUNCH_CHECK();
}

__global__ void __launch_bounds__(256)
dgemm_combine_kernel(int m, int n, const double* __restrict__ part, int nz, double alpha, double beta,
                     const double* __restrict__ P, int ldp, double gamma, const double* __restrict__ Z, int ldz,
                     double* __restrict__ C, int ldc) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)m * n) return;
    const int i = (int)(e % m), j = (int)(e / m);
    double acc = 0.0;
    for (int z = 0; z < nz; ++z) acc += part[(size_t)z * m * n + e];
    double r = alpha * acc;
    if (beta != 0.0) r = fma(beta, P[i + (size_t)j * ldp], r);
    if (gamma != 0.0) r = fma(gamma, Z[i + (size_t)j * ldz], r);
    C[i + (size_t)j * ldc] = r;
}

void dgemm_combine(int m, int n, int k, const double* part, double alpha, double beta, const double* P, int ldp, double gamma,
                   const double* Z, int ldz, double* C, int ldc, cudaStream_t s) {
    if (m <= 0 || n <= 0) return;
    const int nz = dgemm_splitk_slices(m, n, k);
    const int kchunk = cdiv(cdiv(k, nz), GBK) * GBK;
    dgemm_combine_kernel<<<cdiv((long long)m * n, 256), 256, 0, s>>>(m, n, part, cdiv(k, kchunk), alpha, beta, P, ldp, gamma, Z, ldz, C, ldc);
    NLE_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------
// y = A x : CTA = 16 rows x 16 warps; lane pairs cover the 16 rows (128-byte row segments), the warps
// (and the two lane halves) stride over the columns, partial sums are combined in shared memory in
// a fixed order.  m, n <= a few thousand: the matrix is L2 resident, so this is latency bound and
// wants many loads in flight rather than few long dot products.
constexpr int GV_ROWS = 16, GV_WARPS = 16;
__global__ void __launch_bounds__(GV_WARPS * 32)
dgemv_n_kernel(int m, int n, const double* __restrict__ A, int lda,
               const double* __restrict__ x, double* __restrict__ y) {
    __shared__ double part[GV_WARPS * 2][GV_ROWS + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = lane & 15, half = lane >> 4;
    const int i = blockIdx.x * GV_ROWS + r;
    const int slot = warp * 2 + half;          // column phase 0..31
    double acc = 0.0;
    if (i < m) {
        const double* a = A + i;
        int j = slot;
        for (; j + 3 * 32 < n; j += 4 * 32) {
            const double a0 = a[(size_t)j * lda], a1 = a[(size_t)(j + 32) * lda];
            const double a2 = a[(size_t)(j + 64) * lda], a3 = a[(size_t)(j + 96) * lda];
            acc = fma(a0, x[j], acc);
            acc = fma(a1, x[j + 32], acc);
            acc = fma(a2, x[j + 64], acc);
            acc = fma(a3, x[j + 96], acc);
        }
        for (; j < n; j += 32) acc = fma(a[(size_t)j * lda], x[j], acc);
    }
    part[slot][r] = acc;
    __syncthreads();
    if (threadIdx.x < GV_ROWS) {
        const int ii = blockIdx.x * GV_ROWS + threadIdx.x;
        if (ii < m) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < GV_WARPS * 2; ++q) s += part[q][threadIdx.x];
            y[ii] = s;
        }
    }
}

// y = A^T x : one warp per column (coalesced along the column), fixed shuffle-tree order.
__global__ void dgemv_t_kernel(int m, int n, const double* __restrict__ A, int lda,
                               const double* __restrict__ x, double* __restrict__ y) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= n) return;
    const double* col = A + (size_t)warp * lda;
    double acc = 0.0;
    for (int i = lane; i < m; i += 32) acc = fma(col[i], x[i], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[warp] = acc;
}

// Sinkhorn half-step, sample side, one pass over U instead of two dgemv_n + two vector kernels:
//   w[i]  = sum_j U[i,j] t[j]                               (rest pixels use k_j^T w)
//   xs[i] = recip( sum_j U[i,j] (lam[j] t[j]) )             (samples: U[s,:] Lam t, inplaceReciprocal filter.cpp:42-54)
// U (13 MB at the bench config) does not survive in L2 between two half-iterations (each streams > 150 MB of level tables), so
// both kernels below are bound by the latency of a few dependent batches of loads, not by bandwidth: they are shaped for many
// CTAs and many loads in flight.  CTA = 8 rows x 16 warps; a quarter-warp covers the 8 rows (64-byte row segments), the 64 column
// phases (16 warps x 4 quarters) stride over the columns, partial sums are combined in shared memory in a fixed order.
constexpr int SS_ROWS = 8, SS_WARPS = 16, SS_PH = SS_WARPS * (32 / SS_ROWS);
__global__ void __launch_bounds__(SS_WARPS * 32)
sk_sample_step_kernel(int m, int n, const double* __restrict__ A, int lda, const double* __restrict__ t,
                      const double* __restrict__ lam, double eps, double* __restrict__ w, double* __restrict__ xs) {
    __shared__ double part[2][SS_PH][SS_ROWS + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = lane & (SS_ROWS - 1), sub = lane / SS_ROWS;
    const int i = blockIdx.x * SS_ROWS + r;
    const int slot = warp * (32 / SS_ROWS) + sub;
    double acc0 = 0.0, acc1 = 0.0;
    if (i < m) {
        const double* a = A + i;
        int j = slot;
        for (; j + 7 * SS_PH < n; j += 8 * SS_PH) {          // 8 column phases in flight per thread, ascending summation order
            double av[8], tv[8], lv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { av[u] = a[(size_t)(j + SS_PH * u) * lda]; tv[u] = t[j + SS_PH * u]; lv[u] = lam[j + SS_PH * u]; }
#pragma unroll
            for (int u = 0; u < 8; ++u) { acc0 = fma(av[u], tv[u], acc0); acc1 = fma(av[u], lv[u] * tv[u], acc1); }
        }
        for (; j < n; j += SS_PH) {
            const double a0 = a[(size_t)j * lda], t0 = t[j];
            acc0 = fma(a0, t0, acc0);
            acc1 = fma(a0, lam[j] * t0, acc1);
        }
    }
    part[0][slot][r] = acc0;
    part[1][slot][r] = acc1;
    __syncthreads();
    if (threadIdx.x < 2 * SS_ROWS) {
        const int which = threadIdx.x / SS_ROWS, rr = threadIdx.x % SS_ROWS;
        const int ii = blockIdx.x * SS_ROWS + rr;
        if (ii < m) {
            double s4[4] = {0.0, 0.0, 0.0, 0.0};                 // four interleaved chains, combined in a fixed order
#pragma unroll
            for (int q = 0; q < SS_PH; q += 4) {
                s4[0] += part[which][q][rr]; s4[1] += part[which][q + 1][rr];
                s4[2] += part[which][q + 2][rr]; s4[3] += part[which][q + 3][rr];
            }
            const double sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
            if (which == 0) w[ii] = sum;
            else xs[ii] = (fabs(sum) >= eps) ? 1.0 / sum : 0.0;
        }
    }
}

// t[j] = sum_i U[i,j] x[i] + inv_lam[j] * sum_i U[i,j] s[i]     (phi^T x in factor form): one pass over U instead of
// two dgemv_t + an axpy.  CTA = 8 warps = 2 columns x 4 row quarters (one warp per column left 8 warps per SM with seven
// dependent batches of loads each); the quarters are combined in a fixed order.
__global__ void __launch_bounds__(256)
sk_phiT_kernel(int m, int n, const double* __restrict__ A, int lda, const double* __restrict__ x,
               const double* __restrict__ sv, const double* __restrict__ inv_lam, double* __restrict__ t) {
    __shared__ double part[2][2][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cj = warp >> 2, qr = warp & 3;
    const int j = blockIdx.x * 2 + cj;
    const int mq = ((m + 3) / 4 + 31) & ~31;                 // rows per quarter, a multiple of 32
    double acc0 = 0.0, acc1 = 0.0;
    if (j < n) {
        const double* col = A + (size_t)j * lda;
        const int i1 = min(m, (qr + 1) * mq);
        int i = qr * mq + lane;
        for (; i + 7 * 32 < i1; i += 8 * 32) {               // 8 loads in flight per lane, ascending summation order
            double av[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) av[u] = col[i + 32 * u];
#pragma unroll
            for (int u = 0; u < 8; ++u) { acc0 = fma(av[u], x[i + 32 * u], acc0); acc1 = fma(av[u], sv[i + 32 * u], acc1); }
        }
        for (; i + 32 < i1; i += 2 * 32) {
            const double a0 = col[i], a1 = col[i + 32];
            acc0 = fma(a0, x[i], acc0); acc1 = fma(a0, sv[i], acc1);
            acc0 = fma(a1, x[i + 32], acc0); acc1 = fma(a1, sv[i + 32], acc1);
        }
        for (; i < i1; i += 32) {
            const double a = col[i];
            acc0 = fma(a, x[i], acc0);
            acc1 = fma(a, sv[i], acc1);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
        acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
    }
    if (lane == 0) { part[cj][0][qr] = acc0; part[cj][1][qr] = acc1; }
    __syncthreads();
    if (threadIdx.x < 2) {
        const int jj = blockIdx.x * 2 + threadIdx.x;
        if (jj < n) {
            const double* p0 = part[threadIdx.x][0];
            const double* p1 = part[threadIdx.x][1];
            const double a0 = (p0[0] + p0[1]) + (p0[2] + p0[3]);
            const double a1 = (p1[0] + p1[1]) + (p1[2] + p1[3]);
            t[jj] = fma(inv_lam[jj], a1, a0);
        }
    }
}

void sk_sample_step(int m, int n, const double* U, int ldu, const double* t, const double* lam, double eps, double* w,
                    double* xs, cudaStream_t s) {
    if (m <= 0) return;
    sk_sample_step_kernel<<<cdiv(m, SS_ROWS), SS_WARPS * 32, 0, s>>>(m, n, U, ldu, t, lam, eps, w, xs);
    NLE_LAUNCH_CHECK();
}

void sk_phiT(int m, int n, const double* U, int ldu, const double* x, const double* sv, const double* inv_lam, double* t,
             cudaStream_t s) {
    if (n <= 0) return;
    sk_phiT_kernel<<<cdiv(n, 2), 256, 0, s>>>(m, n, U, ldu, x, sv, inv_lam, t);
    NLE_LAUNCH_CHECK();
}

void dgemv_n(int m, int n, const double* A, int lda, const double* x, double* y, cudaStream_t s) {
    if (m <= 0) return;
    dgemv_n_kernel<<<cdiv(m, GV_ROWS), GV_WARPS * 32, 0, s>>>(m, n, A, lda, x, y);
    NLE_LAUNCH_CHECK();
}

void dgemv_t(int m, int n, const double* A, int lda, const double* x, double* y, cudaStream_t s) {
    if (n <= 0) return;
    dgemv_t_kernel<<<cdiv((long long)n * 32, 256), 256, 0, s>>>(m, n, A, lda, x, y);
    NLE_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------
__global__ void scale_rows_cols_kernel(int m, int n, const double* __restrict__ A, int lda,
                                       const double* __restrict__ rs, const double* __restrict__ cs,
                                       double* __restrict__ out, int ldo) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y;
    if (i >= m || j >= n) return;
    double v = A[i + (size_t)j * lda];
    if (rs) v *= rs[i];
    if (cs) v *= cs[j];
    out[i + (size_t)j * ldo] = v;
}

void scale_rows_cols(int m, int n, const double* A, int lda, const double* rowscale,
                     const double* colscale, double* out, int ldo, cudaStream_t s) {
    if (m <= 0 || n <= 0) return;
    dim3 grid(cdiv(m, 128), n);
    scale_rows_cols_kernel<<<grid, 128, 0, s>>>(m, n, A, lda, rowscale, colscale, out, ldo);
    NLE_LAUNCH_CHECK();
}

__global__ void guarded_reciprocal_kernel(double* v, int n, double eps) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        double x = v[i];
        v[i] = (fabs(x) >= eps) ? 1.0 / x : 0.0;
    }
}

__global__ void guarded_inv_sqrt_kernel(const double* v, double* out, int n, double eps) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        double x = v[i];
        out[i] = (fabs(x) >= eps) ? sqrt(1.0 / x) : 0.0;   // reciprocal then cwiseSqrt, as the reference
    }
}

void guarded_reciprocal(double* v, int n, double eps, cudaStream_t s) {
    if (n <= 0) return;
    guarded_reciprocal_kernel<<<cdiv(n, 256), 256, 0, s>>>(v, n, eps);
    NLE_LAUNCH_CHECK();
}

void guarded_inv_sqrt(const double* v, double* out, int n, double eps, cudaStream_t s) {
    if (n <= 0) return;
    guarded_inv_sqrt_kernel<<<cdiv(n, 256), 256, 0, s>>>(v, out, n, eps);
    NLE_LAUNCH_CHECK();
}

}  // namespace nle
