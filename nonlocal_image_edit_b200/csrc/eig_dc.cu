// Symmetric FP64 eigensolver, direct method: Householder tridiagonalisation + Cuppen divide & conquer
// + reflector back-transformation, all on the device (no LAPACK / cuSOLVER).
//
// Replaces the Eigen SelfAdjointEigenSolver calls of the reference (nle::eigenDecomposition,
// filter.cpp:204-228; call sites filter.cpp:262, 287, 313).  The block-Jacobi solver of eig.cu needs
// 30-50 sweeps on the Gaussian kernel matrices of this filter (measured) -- ~8 n^3
// flops per sweep -- whereas this path costs ~(4/3) n^3 for the reduction plus a D&C phase that
// deflates almost completely on such spectra (top merge k = 362 of 1600 for the bench Ka).
//
//   tridiag_kernel        persistent cooperative kernel; ONE pass over the trailing matrix and ONE
//                         grid.sync per Householder step: the rank-2 update of step j is fused with
//                         the symmetric matrix-vector product of step j+1 (the next reflector is
//                         derived redundantly by every CTA from the updated column j+1).  The matrix lives in L2,
//                         except for the highest columns of every CTA, which stay in shared memory.  Partial mode:
//                         stops after nstop reflectors and leaves the trailing block complete in global memory.
//   tridiag_cluster_kernel   the default for n >= 64: same arithmetic in the same order (bit-identical d, e,
//                         tau, reflectors), trailing matrix resident in shared memory, no grid barrier: p = A v and
//                         the next column travel as flagged 16-byte cells that land in both CTAs of a thread-block
//                         cluster by ONE TMA multicast copy per vector and step (profiles/r2zb_trd_multicast.md:
//                         n = 1600 in 7.3 ms against 10.2 ms for tridiag_kernel).
//                         A matrix whose columns do not fit in the shared memory of 148 SMs (n > ~1610) is reduced by
//                         tridiag_kernel (partial mode) until the trailing block fits and then handed over; alone,
//                         tridiag_kernel is the fallback and the cross-check of tests/test_gpu_eig_variants.py.
//   dc_leaf_kernel        implicit-shift QL on leaves of <= 32 rows, one warp per leaf.
//   dc_setup_kernel       per merge: z vector, rank sort, LAPACK dlaed2-style deflation.
//   dc_rotate_kernel      applies the deflation Givens rotations to the eigenvector columns.
//   dc_secular_kernel     one warp per root; safeguarded Newton + bisection on the BIT PATTERN of the offset from the
//                         nearer pole (<= 62 steps, converges to the last ulp, no safeguards needed).
//   dc_zhat_kernel        Gu-Eisenstat recomputed z (keeps eigenvectors orthogonal to rounding).
//   dc_smat_kernel        normalised eigenvectors of the rank-one-updated diagonal problem.
//   dc_gemm_kernel        Q_new = Q[:, non-deflated] * S  (gathered columns), all merges of a level.
//   backtransform_kernel  U = H_0 ... H_{n-3} Z, one warp per column (column resident in shared memory).
//
// scripts/proto_eig_dc.py is the NumPy prototype of exactly this structure (checked against LAPACK).
#include <cooperative_groups.h>
#include <cuda_pipeline.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <string>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace nle {

namespace {

constexpr double kUnitRoundoff = 1.1102230246251565e-16;
constexpr int kLeaf = 32;
constexpr int kTrdThreads = 512;
constexpr int kTrdWarps = kTrdThreads / 32;

__device__ __forceinline__ int node_start(int n, int depth, int i) { return (int)(((long long)i * n) >> depth); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_prod(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// deterministic block-wide sum for kTrdThreads threads; every thread returns the same value.
__device__ __forceinline__ double block_sum(double v, double* red /*>= 2*kTrdWarps doubles*/, int& phase) {
    v = warp_sum(v);
    double* slot = red + (phase & 1) * kTrdWarps;
    if ((threadIdx.x & 31) == 0) slot[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kTrdWarps; ++w) s += slot[w];
    ++phase;   // alternate slots so the next reduction cannot overwrite values still being read
    return s;
}

constexpr int kTrdRed = 2 * kTrdWarps + 4;     // block_sum slots + 4 scalars handed from their owner thread to the CTA

// w_i = tau p_i - kappa v_i and the updated entry of column j+1, in ONE fixed form (explicit roundings) so that both
// tridiagonalisation kernels -- and the shortcut w_{j+1} = tau p_{j+1} - kappa every thread computes for itself -- agree
// bit for bit.
__device__ __forceinline__ double trd_w(double tau, double p, double kappa, double v) { return fma(tau, p, -__dmul_rn(kappa, v)); }
__device__ __forceinline__ double trd_cn(double c, double v, double wj1, double w) { return __dsub_rn(fma(-v, wj1, c), w); }

// Steps (1b)-(2) of Householder step j, shared by tridiag_kernel and tridiag_cluster_kernel.
// On entry thread tid has written w[i] = p_i and cn[i] = A(i, j+1) for its rows i = j+1+tid, j+1+tid+kTrdThreads, ...,
// `part` = its share of p.v over the same rows, and the owner of row j+1 has stored p_{j+1} in sc[2]; NO barrier yet.
// Every loop below runs over the thread's own rows again, so the only barriers are the two inside the block sums and the
// final one: w = tau p - kappa v, cn = updated column j+1, its diagonal and sub-diagonal entries travel through sc[0..1];
// if `make`, cn(j+3:) is scaled into reflector j+1 (cn[j+2] = 1).  All threads return the same scalars.
__device__ __forceinline__ void trd_step_vectors(int j, int n, double tau_j, double part, double* __restrict__ w,
                                                 const double* __restrict__ v, double* __restrict__ cn, double* red, int& phase,
                                                 bool make, double& diag_next, double& beta_next, double& tau_next) {
    const int tid = threadIdx.x;
    double* sc = red + 2 * kTrdWarps;
    const double dot = block_sum(part, red, phase);
    const double kappa = 0.5 * tau_j * tau_j * dot;
    const double wj1 = trd_w(tau_j, sc[2], kappa, 1.0);      // = w[j+1]  (v[j+1] == 1)
    double part2 = 0.0;
    for (int i = j + 1 + tid; i < n; i += kTrdThreads) {
        const double vi = v[i];
        const double wi = trd_w(tau_j, w[i], kappa, vi);
        const double ci = trd_cn(cn[i], vi, wj1, wi);
        w[i] = wi;
        cn[i] = ci;
        if (i >= j + 3) part2 = fma(ci, ci, part2);
        else sc[i - (j + 1)] = ci;                           // rows j+1 (next diagonal) and j+2 (alpha)
    }
    tau_next = 0.0;
    if (!make) {
        __syncthreads();
        diag_next = sc[0];
        beta_next = (j + 2 < n) ? sc[1] : 0.0;               // j+1 == n-2: the last off-diagonal
        return;
    }
    const double xn2 = block_sum(part2, red, phase);         // its barrier also publishes w, cn and sc
    diag_next = sc[0];
    const double alpha = sc[1];
    if (xn2 == 0.0) {
        beta_next = alpha;
        if (tid == 1) cn[j + 2] = 1.0;                       // thread 1 owns row j+2
    } else {
        beta_next = -copysign(sqrt(fma(alpha, alpha, xn2)), alpha);
        tau_next = (beta_next - alpha) / beta_next;
        const double scal = 1.0 / (alpha - beta_next);
        for (int i = j + 1 + tid; i < n; i += kTrdThreads) {
            if (i >= j + 3) cn[i] *= scal;
            else if (i == j + 2) cn[i] = 1.0;
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Householder tridiagonalisation A = Q T Q^T of a symmetric matrix held in FULL storage (both
// triangles, column-major).  On exit: d (n), e (n-1), tau (n-1); reflector j (H_j = I - tau_j v v^T,
// v[j+1] = 1) is stored in A(j+1:n, j) including the explicit 1.
//
// nstop < n - 2: PARTIAL reduction.  Only the reflectors 0 .. nstop-1 are produced (d, e, tau [0, nstop) written) and the
// kernel leaves the trailing block A(nstop:n, nstop:n) fully updated in global memory, both triangles.  Householder
// tridiagonalisation of the rest is then an independent problem on that block: a second launch (of this kernel or of
// tridiag_cluster_kernel) on (A + nstop*lda + nstop, lda, n - nstop, d + nstop, e + nstop, tau + nstop) performs exactly the
// operations the uninterrupted kernel would have performed, in the same order -- the result is bit-identical.  This is how
// matrices too large for the shared memory of the SMs (n > ~1700) get the resident kernel for their last ~1700 columns.
__global__ void __launch_bounds__(kTrdThreads, 1)
tridiag_kernel(double* __restrict__ A, int lda, int n, double* __restrict__ d, double* __restrict__ e,
               double* __restrict__ tau, double* __restrict__ pbuf, int nstop, int qs) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double sm[];
    double* v = sm;            // current reflector, global row indexing
    double* w = sm + n;
    double* cn = sm + 2 * (size_t)n;   // updated next column -> next reflector
    double* red = sm + 3 * (size_t)n;  // kTrdRed
    double* cache = red + kTrdRed;     // qs columns of n rows
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, b = blockIdx.x;
    int phase = 0;
    // The trailing matrix lives in L2, and for large n a step is bound by reading and writing it once (~5 TB/s).  The qs
    // HIGHEST columns this CTA owns (the ones that stay in the trailing matrix longest) are therefore kept in shared memory for
    // the whole launch: their update + symv pass makes no global access.  A cached column goes back to global memory when
    // somebody else needs it -- in the step in which it becomes the next column (every CTA reads that one), and at the
    // hand-over of a partial reduction.  Same operations in the same order as for an uncached column.
    const int q_cnt = (b < n) ? (n - 1 - b) / G + 1 : 0;      // columns this CTA owns: slots q = 0 .. q_cnt-1
    const int q_c0 = max(0, q_cnt - qs);                      // slots q >= q_c0 are cached, cache slot q - q_c0

    if (n < 3) {
        if (b == 0 && tid == 0) {
            d[0] = A[0];
            if (n == 2) { e[0] = A[1]; d[1] = A[1 + (size_t)lda]; tau[0] = 0.0; }
        }
        return;
    }

    // make a reflector from cn[j0+1 .. n-1] in place (x -> v, v[j0+1] = 1); returns beta, sets tau_out.
    // cn[j0] (the diagonal) is left untouched.  All threads get the same result.
    auto make_reflector = [&](int j0, double& tau_out) -> double {
        const double alpha = cn[j0 + 1];
        double part = 0.0;
        // thread tid sums the rows j0 + tid, j0 + tid + kTrdThreads, ... (from j0 + 2 on): the row-to-thread map of
        // trd_step_vectors, so that a reduction that starts on a trailing block reproduces the uninterrupted one bit for bit
        for (int i = j0 + tid; i < n; i += kTrdThreads)
            if (i >= j0 + 2) part = fma(cn[i], cn[i], part);
        const double xn2 = block_sum(part, red, phase);
        double beta;
        if (xn2 == 0.0) {
            tau_out = 0.0;
            beta = alpha;
            __syncthreads();
            if (tid == 0) cn[j0 + 1] = 1.0;
        } else {
            beta = -copysign(sqrt(fma(alpha, alpha, xn2)), alpha);
            tau_out = (beta - alpha) / beta;
            const double scal = 1.0 / (alpha - beta);
            __syncthreads();
            for (int i = j0 + 2 + tid; i < n; i += kTrdThreads) cn[i] *= scal;
            if (tid == 0) cn[j0 + 1] = 1.0;
        }
        __syncthreads();
        return beta;
    };

    // symv for owned columns c >= c0: pout[c] = sum_{i >= c0} A[i,c] * vec[i]   (no update)
    // (used once, for the first reflector)
    for (int q = q_c0; q < q_cnt; ++q) {
        const double* src = A + (size_t)(b + G * q) * lda;
        double* dst = cache + (size_t)(q - q_c0) * n;
        for (int i = tid; i < n; i += kTrdThreads) dst[i] = src[i];
    }
    double tau_j, beta_j, diag_j;
    {
        for (int i = tid; i < n; i += kTrdThreads) cn[i] = A[i];
        __syncthreads();
        diag_j = cn[0];
        beta_j = make_reflector(0, tau_j);
        // v <- cn
        double* t = v; v = cn; cn = t;
        for (int q = warp; ; q += kTrdWarps) {
            const int c = b + G * q;
            if (c >= n) break;
            if (c < 1) continue;
            const double* col = A + (size_t)c * lda;
            double acc = 0.0;
            for (int i = 1 + lane; i < n; i += 32) acc = fma(col[i], v[i], acc);
            acc = warp_sum(acc);
            if (lane == 0) pbuf[c] = acc;
        }
        grid.sync();   // orders this step's global writes (stcg / plain stores) before every CTA's ldcg reads of the next step
    }

    for (int j = 0; j <= n - 3; ++j) {
        const double* p = pbuf + (size_t)(j & 1) * n;
        double* pn = pbuf + (size_t)((j + 1) & 1) * n;
        const bool last_partial = (j == nstop - 1);     // partial reduction: reflector j is the last one of this launch
        // (0) publish reflector j (column j of A is no longer read by anybody in this kernel)
        if (b == 0) {
            double* col = A + (size_t)j * lda;
            for (int i = j + 1 + tid; i < n; i += kTrdThreads) col[i] = v[i];
            if (tid == 0) { d[j] = diag_j; e[j] = beta_j; tau[j] = tau_j; }
        }
        // (1) w = tau*p - (tau^2/2)(p.v) v      (the loads of p and of column j+1 are issued together: both only
        //     depend on the grid barrier, and their L2 round trips are the longest links of the per-step chain)
        double part = 0.0;
        {
            const double* col = A + (size_t)(j + 1) * lda;
            for (int i = j + 1 + tid; i < n; i += kTrdThreads) {
                const double pi = __ldcg(p + i);
                const double ci = __ldcg(col + i);
                w[i] = pi;
                cn[i] = ci;
                part = fma(pi, v[i], part);
                if (i == j + 1) red[2 * kTrdWarps + 2] = pi;
            }
        }
        // (1b)-(2) w, the updated column j+1 (rows j+1..n-1), the next diagonal and the next reflector
        const bool has_next = (j + 1 <= n - 3) && !last_partial;
        double diag_next, tau_next, beta_next;
        trd_step_vectors(j, n, tau_j, part, w, v, cn, red, phase, has_next, diag_next, beta_next, tau_next);
        if (last_partial) {
            // hand-over: the updated column j+1 goes back to A unscaled (the next launch builds reflector j+1 from it);
            // its mirror image, row j+1 of the columns c >= j+2, is written by the owners of those columns below
            if (b == 0) {
                double* col = A + (size_t)(j + 1) * lda;
                for (int i = j + 1 + tid; i < n; i += kTrdThreads) col[i] = cn[i];
            }
        }
        // (3) rank-2 update of the owned columns c >= j+2 fused with the next symv
        for (int q = warp; ; q += kTrdWarps) {
            const int c = b + G * q;
            if (c >= n) break;
            if (c < j + 2) continue;
            double* col = A + (size_t)c * lda;
            const double wc = w[c], vc = v[c];
            double acc = 0.0;
            int i = j + 2 + lane;
            if (q >= q_c0) {
                // cached column: shared memory only; written through to global memory when it is the next column (c == j+2)
                // or the launch hands over.  Loads in front of the stores (same address space to the compiler).
                double* sc = cache + (size_t)(q - q_c0) * n;
                const bool wb = (c == j + 2) || last_partial;
                for (; i + 3 * 32 < n; i += 4 * 32) {
                    double a[4], wv[4], vv[4], cv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int ii = i + 32 * u;
                        a[u] = sc[ii]; wv[u] = w[ii]; vv[u] = v[ii]; cv[u] = cn[ii];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int ii = i + 32 * u;
                        a[u] = fma(-wv[u], vc, fma(-vv[u], wc, a[u]));
                        acc = fma(a[u], cv[u], acc);
                        sc[ii] = a[u];
                        if (wb) __stcg(col + ii, a[u]);
                    }
                }
                for (; i < n; i += 32) {
                    double a = sc[i];
                    a = fma(-w[i], vc, fma(-v[i], wc, a));
                    acc = fma(a, cn[i], acc);
                    sc[i] = a;
                    if (wb) __stcg(col + i, a);
                }
            } else {
            // batches of 16 independent loads to cover the L2 latency (the column is the longest link of the step)
            for (; i + 15 * 32 < n; i += 16 * 32) {
                double a[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) a[u] = __ldcg(col + i + 32 * u);
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int ii = i + 32 * u;
                    a[u] = fma(-w[ii], vc, fma(-v[ii], wc, a[u]));
                    acc = fma(a[u], cn[ii], acc);
                    __stcg(col + ii, a[u]);
                }
            }
            for (; i + 3 * 32 < n; i += 4 * 32) {
                double a[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) a[u] = __ldcg(col + i + 32 * u);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int ii = i + 32 * u;
                    a[u] = fma(-w[ii], vc, fma(-v[ii], wc, a[u]));
                    acc = fma(a[u], cn[ii], acc);
                    __stcg(col + ii, a[u]);
                }
            }
            for (; i < n; i += 32) {
                double a = __ldcg(col + i);
                a = fma(-w[i], vc, fma(-v[i], wc, a));
                acc = fma(a, cn[i], acc);
                __stcg(col + i, a);
            }
            }
            if (has_next) {
                acc = warp_sum(acc);
                if (lane == 0) pn[c] = acc;
            }
            if (last_partial && lane == 0) col[j + 1] = cn[c];
        }
        if (last_partial) return;      // d, e, tau [0, nstop) and the trailing block are complete
        // rotate state
        diag_j = diag_next; beta_j = beta_next; tau_j = tau_next;
        { double* t = v; v = cn; cn = t; }
        grid.sync();   // orders this step's global writes (stcg / plain stores) before every CTA's ldcg reads of the next step
    }
    // epilogue: j = n-2 entries and the last diagonal
    if (b == 0 && tid == 0) {
        d[n - 2] = diag_j;
        e[n - 2] = beta_j;
        tau[n - 2] = 0.0;
        d[n - 1] = __ldcg(A + (size_t)(n - 1) * lda + (n - 1));
    }
}

// ---------------------------------------------------------------------------------------------
// Shared-memory-resident tridiagonalisation (tridiag_cluster_kernel below).  Same arithmetic in the same order as
// tridiag_kernel (d, e, tau and the reflectors come out bit-identical); what changes is where the data lives and how
// the CTAs synchronise:
//   * CTA b keeps its columns c = b, b+G, ... of the trailing matrix in shared memory for the whole
//     factorisation (n = 1600 on 148 SMs: 11 columns x 12.8 KB), so the rank-2 update + symv pass of a
//     step makes no global loads or stores at all;
//   * the two vectors every CTA needs at the start of a step -- p = A v and the next column -- are
//     exchanged through global memory as flagged 16-byte cells {lo, tag, hi, tag} (two 8-byte halves, each
//     written atomically together with its tag; the low-latency protocol NCCL calls LL): there is no store-acknowledge
//     fence, no barrier atomic and no barrier poll; the consumers fetch all cells of a step with one TMA multicast copy
//     per vector and cluster, validate the tags, and re-poll the few cells that were not yet written;
//   * cells are double-buffered by the parity of the tag; a cluster leaves the kernel as soon as it owns no
//     column of the trailing matrix any more, so the CTAs that still exchange data are never more than
//     one step apart (each waits for cells written by all the others), which is what makes two buffers
//     enough.  The CTA that owns column n-1 lives to the end and writes d, e, tau and the reflectors.
constexpr unsigned kTrdSpinLimit = 1u << 21;   // poll rounds (~1 us each) before a waiting CTA gives up
constexpr int kResPer = 4;   // exchange cells per thread and vector: n <= kResPer * kTrdThreads

// relaxed, gpu scope, two 8-byte elements {tag:lo, tag:hi} (the volatile 4 x u32 form NCCL uses measured the same)
__device__ __forceinline__ void ll_store(uint4* cell, double x, unsigned tag) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(x);
    const unsigned long long t = (unsigned long long)tag << 32;
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(cell), "l"(t | (bits & 0xffffffffull)), "l"(t | (bits >> 32))
                 : "memory");
}
__device__ __forceinline__ uint4 ll_load(const uint4* cell) {
    uint4 r;
    unsigned long long a, b;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(cell) : "memory");
    r.x = (unsigned)a; r.y = (unsigned)(a >> 32); r.z = (unsigned)b; r.w = (unsigned)(b >> 32);
    return r;
}
__device__ __forceinline__ double ll_value(const uint4& c) {
    return __longlong_as_double((long long)(((unsigned long long)c.z << 32) | (unsigned long long)c.x));
}

// PROF: thread 0 of the CTA that owns column n-1 accumulates clock64() spans of the phases of a step
//       (developer diagnostic, NLE_B200_TRD_PROF=1; results in profiles/r1l_trd_phases.md).
// clock64 read that the compiler cannot hoist above the computation of `dep`
__device__ __forceinline__ long long clock_after(double dep) {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "d"(dep) : "memory");
    return t;
}
#define TRD_STAMP(k, dep)                                              \
    do {                                                               \
        if (PROF && tid == 0) {                                        \
            const long long t_ = clock_after(dep);                     \
            pacc[k] += t_ - tprev;                                     \
            tprev = t_;                                                \
        }                                                              \
    } while (0)

// ---- TMA multicast landing of the exchange cells (MC = true) ----------------------------------------------------
// One elected thread of cluster rank 0 issues, per Householder step, two 1-D bulk copies (the p cells and the next-column
// cells of the rows still alive) with .multicast::cluster: the TMA engine reads the cells from L2 ONCE per cluster and
// writes them into the landing buffers of every CTA of the cluster, completing a transaction count on each CTA's own
// mbarrier -- no thread issues a load for the bulk of the exchange and nothing is forwarded through distributed shared
// memory by threads.  What lands is validated cell by cell (tags); a cell that was not yet written when the copy passed
// is re-polled by its thread with the strong 16-byte load, so the copy never has to be repeated as a whole.
__device__ __forceinline__ uint32_t trd_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void trd_mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(trd_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void trd_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(trd_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void trd_mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(trd_smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void trd_tma_multicast(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar, unsigned short mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                     trd_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(trd_smem_u32(bar)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void trd_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void trd_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// Clusters: the CTAs of a thread-block cluster receive the cells of a step with ONE L2 read (TMA multicast).  Round 2
// history (profiles/r2a_trd_sweep.md, r2zb_trd_multicast.md): per-thread polling of all cells by every CTA (n = 1600:
// 9.5 ms); clusters of 2 whose CTAs poll every second cell with strong 256-bit loads and forward them through distributed
// shared memory, one cluster.sync per step (8.7 - 9.0 ms); the multicast landing below (7.9 ms; 7.3 with trd_step_vectors): the poll + forward phase
// of 6.6k cycles per step became 3.7k (copy in flight) + 1.5k (validation); the copy is issued column first because the p
// cells are the last thing a producer writes.  Unicast copies per CTA measure the same as the multicast (the L2 merges the
// concurrent reads of a cluster), larger clusters gain < 2 % at n <= 1041 and do not fit at n = 1600.
template <bool PROF>
__global__ void __launch_bounds__(kTrdThreads, 1)
tridiag_cluster_kernel(double* __restrict__ A, int lda, int n, double* __restrict__ d, double* __restrict__ e,
                        double* __restrict__ tau, uint4* __restrict__ ll /* 4n cells, zeroed before the launch */,
                        long long* __restrict__ prof /* 16 counters when PROF */) {
    cg::cluster_group cluster = cg::this_cluster();
    const int S = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    extern __shared__ __align__(16) double sm[];
    // Nobody but the TMA engine writes into this CTA's shared memory: the cells land in LP / LC (16 bytes per row and
    // vector), w is one vector, v / cn ping-pong between two.
    const int ne = (n + 1) & ~1;                 // even stride: every vector of cells starts on a 32-byte sector
    uint4* LP = reinterpret_cast<uint4*>(sm);                    // ne cells (p)
    uint4* LC = LP + ne;                                         // ne cells (next column)
    double* Wb = sm + 4 * (size_t)ne;                            // n
    double* Vb = Wb + (size_t)n;                                 // 2 n
    double* red = Vb + 2 * (size_t)n;                            // kTrdRed
    uint64_t* mbar = reinterpret_cast<uint64_t*>(red + kTrdRed); // one mbarrier (+ 8 bytes of padding)
    double* cols = red + kTrdRed + 2;                            // owned columns, slot q holds column b + G*q (all n rows)
    double* v = Vb;            // current reflector, global row indexing
    double* w = Wb;
    double* cn = Vb;           // updated next column -> next reflector (first reflector is built here, then becomes v)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // (Which CTA owns which column does not matter for speed: striding the logical CTA index over the clusters / GPCs changed
    // nothing, profiles/r2k_trd_experiments.md.)
    const int G = gridDim.x, b = blockIdx.x;
    const int q_last = (n - 1 - b) / G;          // b < G <= n
    const int c_last = b + G * q_last;           // the largest column this CTA owns
    const bool writer = (c_last == n - 1);
    int cl_last = 0;                             // the largest column any CTA of this cluster owns
    for (int r = 0; r < S; ++r) {
        const int br = b - rank + r;
        cl_last = max(cl_last, br + G * ((n - 1 - br) / G));
    }
    if (tid == 0) {
        trd_mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster.sync();            // no multicast reaches a CTA before its mbarrier exists
    trd_cluster_arrive();      // matches the wait in front of the first multicast
    int phase = 0;
    long long pacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = 0, rounds = 0;
    // cells of tag t: p at ll + (t&1)*2n, next column at ll + (t&1)*2n + n
    auto pcell = [&](unsigned t) { return ll + (size_t)(t & 1u) * 2 * ne; };
    auto ccell = [&](unsigned t) { return ll + (size_t)(t & 1u) * 2 * ne + ne; };

    auto make_reflector = [&](int j0, double& tau_out) -> double {     // identical to tridiag_kernel's
        const double alpha = cn[j0 + 1];
        double part = 0.0;
        // thread tid sums the rows j0 + tid, j0 + tid + kTrdThreads, ... (from j0 + 2 on): the row-to-thread map of
        // trd_step_vectors, so that a reduction that starts on a trailing block reproduces the uninterrupted one bit for bit
        for (int i = j0 + tid; i < n; i += kTrdThreads)
            if (i >= j0 + 2) part = fma(cn[i], cn[i], part);
        const double xn2 = block_sum(part, red, phase);
        double beta;
        if (xn2 == 0.0) {
            tau_out = 0.0;
            beta = alpha;
            __syncthreads();
            if (tid == 0) cn[j0 + 1] = 1.0;
        } else {
            beta = -copysign(sqrt(fma(alpha, alpha, xn2)), alpha);
            tau_out = (beta - alpha) / beta;
            const double scal = 1.0 / (alpha - beta);
            __syncthreads();
            for (int i = j0 + 2 + tid; i < n; i += kTrdThreads) cn[i] *= scal;
            if (tid == 0) cn[j0 + 1] = 1.0;
        }
        __syncthreads();
        return beta;
    };

    // owned columns -> shared memory; column 0 -> first reflector (every CTA, redundantly)
    for (int q = 0; q <= q_last; ++q) {
        const double* src = A + (size_t)(b + G * q) * lda;
        double* dst = cols + (size_t)q * n;
        for (int i = tid; i < n; i += kTrdThreads) dst[i] = src[i];
    }
    for (int i = tid; i < n; i += kTrdThreads) cn[i] = A[i];
    __syncthreads();
    double tau_j, beta_j, diag_j;
    diag_j = cn[0];
    beta_j = make_reflector(0, tau_j);
    v = Vb;                    // step j: v = Vb[j & 1], cn = Vb[(j + 1) & 1]
    // p = A v over the owned columns c >= 1 (tag 1); the owner of column 1 also publishes that column
    for (int q = warp; q <= q_last; q += kTrdWarps) {
        const int c = b + G * q;
        if (c < 1) continue;
        const double* col = cols + (size_t)q * n;
        double acc = 0.0;
        for (int i = 1 + lane; i < n; i += 32) acc = fma(col[i], v[i], acc);
        acc = warp_sum(acc);
        if (lane == 0) ll_store(pcell(1) + c, acc, 1u);
        if (c == 1)
            for (int i = 1 + lane; i < n; i += 32) ll_store(ccell(1) + i, col[i], 1u);
    }
    if (PROF && tid == 0) {
        tprev = clock64();
#pragma unroll
        for (int k = 0; k < 10; ++k) pacc[k] = 0;
    }

    for (int j = 0; j <= n - 3; ++j) {
        if (cl_last < j + 2) {            // cluster-uniform: the whole cluster leaves together, after its last barrier
            trd_cluster_wait();
            return;
        }
        v = Vb + (size_t)(j & 1) * n;
        cn = Vb + (size_t)((j + 1) & 1) * n;
        const unsigned T = (unsigned)(j + 1);
        const bool has_next = (j + 1 <= n - 3);
        double diag_next, tau_next, beta_next;
        // (1) the cells of rows j+1 .. n-1 land in every CTA of the cluster with one L2 read per cluster
        trd_cluster_wait();            // every CTA of the cluster has finished reading the landing buffers of step j-1
        if (tid == 0) {
            const unsigned bytes = (unsigned)(n - (j + 1)) * (unsigned)sizeof(uint4);
            trd_mbar_expect_tx(mbar, 2 * bytes);
            if (rank == 0) {
                const unsigned short mask = (unsigned short)((1u << S) - 1u);
                // column first: its cells were published while the producers were still updating; the p cells are the last
                // thing a producer writes in a step, so they are read last
                trd_tma_multicast(LC + (j + 1), ccell(T) + (j + 1), bytes, mbar, mask);
                trd_tma_multicast(LP + (j + 1), pcell(T) + (j + 1), bytes, mbar, mask);
            }
        }
        trd_mbar_wait(mbar, (unsigned)(j & 1));
        TRD_STAMP(0, __longlong_as_double((long long)LP[j + 1].x));      // slot 0 = barrier + copy in flight
        double part = 0.0;
        {
            unsigned spins = 0;
#pragma unroll
            for (int u = 0; u < kResPer; ++u) {
                const int i = j + 1 + tid + u * kTrdThreads;
                if (i < n) {
                    uint4 P = LP[i], C = LC[i];
                    while (P.y != T || P.w != T) {                 // not yet written when the copy passed: poll the cell itself
                        P = ll_load(pcell(T) + i);
                        if (PROF) ++rounds;
                        if (++spins > kTrdSpinLimit) __trap();     // a lost peer must not hang the device: abort the launch loudly
                    }
                    while (C.y != T || C.w != T) {
                        C = ll_load(ccell(T) + i);
                        if (PROF) ++rounds;
                        if (++spins > kTrdSpinLimit) __trap();
                    }
                    const double pi = ll_value(P);
                    w[i] = pi;
                    cn[i] = ll_value(C);
                    part = fma(pi, v[i], part);
                    if (i == j + 1) red[2 * kTrdWarps + 2] = pi;
                }
            }
        }
        __syncwarp();
        trd_cluster_arrive();          // this thread is done with the landing buffers
        TRD_STAMP(9, part);            // slot 9 = validation + re-polls
        // (1b)-(2) w = tau*p - (tau^2/2)(p.v) v, updated column j+1 -> next diagonal and next reflector
        trd_step_vectors(j, n, tau_j, part, w, v, cn, red, phase, has_next, diag_next, beta_next, tau_next);
        TRD_STAMP(6, beta_next);
        // (3) rank-2 update of the owned columns c >= j+2 (in shared memory) fused with the next symv;
        //     the owner of column j+2 publishes the updated column as it goes
        const unsigned Tn = (unsigned)(j + 2);
        for (int q = warp; q <= q_last; q += kTrdWarps) {
            const int c = b + G * q;
            if (c < j + 2) continue;
            double* col = cols + (size_t)q * n;
            const double wc = w[c], vc = v[c];
            const bool pub = has_next && c == j + 2;
            uint4* cc = ccell(Tn);
            double acc = 0.0;
            int i = j + 2 + lane;
            // batches with all shared-memory loads in front: col, w, v, cn are the same address space to the compiler,
            // so without this every iteration's loads wait behind the previous iteration's store (profiles/
            // r1l_trd_phases.md: ~140 cycles per 32-row iteration).  Same operations in the same order per lane.
            for (; i + 3 * 32 < n; i += 4 * 32) {
                double a[4], wv[4], vv[4], cv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int ii = i + 32 * u;
                    a[u] = col[ii]; wv[u] = w[ii]; vv[u] = v[ii]; cv[u] = cn[ii];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int ii = i + 32 * u;
                    a[u] = fma(-wv[u], vc, fma(-vv[u], wc, a[u]));
                    acc = fma(a[u], cv[u], acc);
                    col[ii] = a[u];
                    if (pub) ll_store(cc + ii, a[u], Tn);
                }
            }
            for (; i + 32 < n; i += 2 * 32) {
                double a[2], wv[2], vv[2], cv[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int ii = i + 32 * u;
                    a[u] = col[ii]; wv[u] = w[ii]; vv[u] = v[ii]; cv[u] = cn[ii];
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int ii = i + 32 * u;
                    a[u] = fma(-wv[u], vc, fma(-vv[u], wc, a[u]));
                    acc = fma(a[u], cv[u], acc);
                    col[ii] = a[u];
                    if (pub) ll_store(cc + ii, a[u], Tn);
                }
            }
            for (; i < n; i += 32) {
                double a = col[i];
                a = fma(-w[i], vc, fma(-v[i], wc, a));
                acc = fma(a, cn[i], acc);
                col[i] = a;
                if (pub) ll_store(cc + i, a, Tn);
            }
            if (has_next) {
                acc = warp_sum(acc);
                if (lane == 0) ll_store(pcell(Tn) + c, acc, Tn);
            }
        }
        TRD_STAMP(7, 0.0);
        // (0) reflector j and its scalars (off the critical path: nobody in this kernel reads them back)
        if (writer) {
            double* colj = A + (size_t)j * lda;
            for (int i = j + 1 + tid; i < n; i += kTrdThreads) colj[i] = v[i];
            if (tid == 0) { d[j] = diag_j; e[j] = beta_j; tau[j] = tau_j; }
        }
        diag_j = diag_next; beta_j = beta_next; tau_j = tau_next;
        __syncthreads();   // the vectors of this step are dead: the next landing may overwrite w / cn's partner
        TRD_STAMP(8, cn[j + 2]);
    }
    trd_cluster_wait();
    if (writer && tid == 0) {
        d[n - 2] = diag_j;
        e[n - 2] = beta_j;
        tau[n - 2] = 0.0;
        d[n - 1] = cols[(size_t)q_last * n + (n - 1)];
        if (PROF) {
#pragma unroll
            for (int k = 0; k < 10; ++k) prof[k] = pacc[k];
            prof[10] = rounds;
            prof[11] = n - 2;
        }
    }
}
#undef TRD_STAMP

// ---------------------------------------------------------------------------------------------
// Leaves: symmetric tridiagonal QL with implicit Wilkinson shift (EISPACK tql2 lineage), one warp
// per leaf; lane = row of the eigenvector matrix.  The scalar recurrence is executed redundantly by
// all lanes on a shared copy of (d, e) -- every lane writes identical values.
__global__ void __launch_bounds__(32)
dc_leaf_kernel(const double* __restrict__ d_in, const double* __restrict__ e_in, int n, int depth,
               double* __restrict__ d_out, double* __restrict__ Q, int ldq, int* __restrict__ fail) {
    const int leaf = blockIdx.x, lane = threadIdx.x;
    const int s0 = node_start(n, depth, leaf), s1 = node_start(n, depth, leaf + 1);
    const int m = s1 - s0;
    __shared__ double Zs[kLeaf][kLeaf + 1];
    __shared__ double ds[kLeaf + 1], es[kLeaf + 1];
    if (lane < m) {
        double val = d_in[s0 + lane];
        if (lane == 0 && s0 > 0) val -= fabs(e_in[s0 - 1]);          // rank-one tear on the left boundary
        if (lane == m - 1 && s1 < n) val -= fabs(e_in[s1 - 1]);      // ... and on the right boundary
        ds[lane] = val;
        es[lane] = (lane < m - 1) ? e_in[s0 + lane] : 0.0;
    }
    for (int i = 0; i < kLeaf; ++i) Zs[lane][i] = (i == lane) ? 1.0 : 0.0;
    __syncwarp();
    bool failed = false;
    for (int l = 0; l < m; ++l) {
        int iter = 0;
        while (true) {
            int mm = l;
            while (mm < m - 1) {
                const double dd = fabs(ds[mm]) + fabs(ds[mm + 1]);
                if (fabs(es[mm]) <= kUnitRoundoff * dd) break;
                ++mm;
            }
            if (mm == l) break;
            if (++iter > 80) { failed = true; break; }
            double g = (ds[l + 1] - ds[l]) / (2.0 * es[l]);
            double r = hypot(g, 1.0);
            g = ds[mm] - ds[l] + es[l] / (g + copysign(r, g));
            double s = 1.0, c = 1.0, p = 0.0;
            bool broke = false;
            // The scalar recurrence is one dependent chain per rotation; what is not on it is kept off it: d[i], e[i] are
            // fetched one rotation ahead (nothing in the loop writes them before they are read), d[i+1] and the eigenvector
            // entry Z[lane][i+1] are carried in registers, and r = sqrt(f^2 + g^2) with its reciprocal comes from one rsqrt
            // instead of hypot + two divisions (library hypot only outside the range where f^2 + g^2 is safely normal).
            double e_nx = es[mm - 1], d_nx = ds[mm - 1], d_ip1 = ds[mm];
            double zc = Zs[lane][mm];
            for (int i = mm - 1; i >= l; --i) {
                const double ei = e_nx, di = d_nx;
                if (i > l) { e_nx = es[i - 1]; d_nx = ds[i - 1]; }
                const double z0 = Zs[lane][i];
                const double f = s * ei;
                const double bb = c * ei;
                const double h2 = fma(f, f, g * g);
                double r, rinv;
                if (h2 > 1e-280 && h2 < 1e280) {
                    rinv = rsqrt(h2);
                    r = h2 * rinv;
                } else {
                    r = hypot(f, g);
                    rinv = (r == 0.0) ? 0.0 : 1.0 / r;
                }
                es[i + 1] = r;
                if (r == 0.0) {
                    ds[i + 1] = d_ip1 - p;
                    es[mm] = 0.0;
                    Zs[lane][i + 1] = zc;
                    broke = true;
                    break;
                }
                s = f * rinv;
                c = g * rinv;
                g = d_ip1 - p;
                r = (di - g) * s + 2.0 * c * bb;
                p = s * r;
                ds[i + 1] = g + p;
                g = c * r - bb;
                Zs[lane][i + 1] = s * z0 + c * zc;
                zc = c * z0 - s * zc;
                d_ip1 = di;
            }
            if (!broke) Zs[lane][l] = zc;
            if (broke) continue;
            ds[l] -= p;
            es[l] = g;
            es[mm] = 0.0;
        }
        if (failed) break;
    }
    __syncwarp();
    if (failed && lane == 0) atomicExch(fail, 1);
    if (lane < m) d_out[s0 + lane] = ds[lane];
    for (int i = 0; i < m; ++i)
        if (lane < m) Q[(size_t)(s0 + lane) + (size_t)(s0 + i) * ldq] = Zs[lane][i];
}

// ---------------------------------------------------------------------------------------------
struct DcArrays {
    double* dk;      // n  non-deflated poles (ascending) at [off, off+k)
    double* zk;      // n  their z components
    int* colmap;     // n  source column of output column off+j (non-deflated first, then deflated)
    int* rotA; int* rotB; double* rotC; double* rotS;   // n  deflation rotations at [off, off+nrot)
    int* org;        // n  origin pole of root j
    double* mu;      // n  offset of root j from its origin
    double* zh;      // n  Gu-Eisenstat z-hat
    int* kArr; int* nrotArr; double* rhoArr;            // per merge
};

// One CTA per merge.  Dynamic shared memory: 3*nm doubles + 2*nm ints.
__global__ void __launch_bounds__(1024)
dc_setup_kernel(int n, int depth, const double* __restrict__ e, const double* __restrict__ dcur,
                const double* __restrict__ Q, int ldq, double* __restrict__ dnew, DcArrays a) {
    const int node = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int off = node_start(n, depth, node), end = node_start(n, depth, node + 1);
    const int split = node_start(n, depth + 1, 2 * node + 1);
    const int nm = end - off, n1 = split - off;
    extern __shared__ double smd[];
    double* draw = smd;                 // unsorted d; later: deflated values
    double* dsrt = smd + nm;
    double* zsrt = smd + 2 * (size_t)nm;
    int* csrt = reinterpret_cast<int*>(smd + 3 * (size_t)nm);
    int* cdef = csrt + nm;
    __shared__ double redm[64];
    __shared__ int s_k, s_ndef;

    const double beta = e[split - 1];
    const double sgn = (beta >= 0.0) ? 1.0 : -1.0;
    const double rho = 2.0 * fabs(beta);
    for (int c = tid; c < nm; c += nt) draw[c] = dcur[off + c];
    __syncthreads();
    double dmax = 0.0, zmax = 0.0;
    for (int c = tid; c < nm; c += nt) {
        const double dc = draw[c];
        int rank = 0;
        for (int o = 0; o < nm; ++o) {
            const double dv = draw[o];
            rank += (dv < dc) || (dv == dc && o < c);
        }
        double z = (c < n1) ? Q[(size_t)(split - 1) + (size_t)(off + c) * ldq]
                            : sgn * Q[(size_t)split + (size_t)(off + c) * ldq];
        z *= 0.70710678118654752440;
        dsrt[rank] = dc;
        zsrt[rank] = z;
        csrt[rank] = off + c;
        dmax = fmax(dmax, fabs(dc));
        zmax = fmax(zmax, fabs(z));
    }
    dmax = warp_max(dmax);
    zmax = warp_max(zmax);
    if ((tid & 31) == 0) { redm[tid >> 5] = dmax; redm[32 + (tid >> 5)] = zmax; }
    __syncthreads();
    dmax = 0.0; zmax = 0.0;
    for (int wv = 0; wv < (nt >> 5); ++wv) { dmax = fmax(dmax, redm[wv]); zmax = fmax(zmax, redm[32 + wv]); }
    const double tol = 8.0 * kUnitRoundoff * fmax(dmax, zmax);
    __syncthreads();   // everyone has read draw[] before it is reused for the deflated list

    // Fast path (the common case on these spectra): if no pair of consecutive non-deflated poles passes the rotation
    // screen, the deflation is a pure compaction -- flags, a block scan and a parallel scatter -- and the serial scan
    // below (one thread, up to n dependent steps) is skipped.  The serial scan decides whenever a rotation is possible.
    __shared__ int s_fast, s_wcnt[32], s_wlast[32];
    if (tid == 0) s_fast = 1;
    __syncthreads();
    {
        const bool all_defl = rho * zmax <= tol;
        const int per = (nm + nt - 1) / nt;
        const int c0 = min(nm, tid * per), c1 = min(nm, c0 + per);
        int cnt = 0, last = -1;
        for (int i = c0; i < c1; ++i)
            if (!all_defl && !(rho * fabs(zsrt[i]) <= tol)) { ++cnt; last = i; }
        // block-wide exclusive scan: sum of cnt, max of last
        const int lane = tid & 31, wid = tid >> 5, nwarp = (nt + 31) >> 5;
        int icnt = cnt, ilast = last;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int vc = __shfl_up_sync(0xffffffffu, icnt, o);
            const int vl = __shfl_up_sync(0xffffffffu, ilast, o);
            if (lane >= o) { icnt += vc; ilast = max(ilast, vl); }
        }
        if (lane == 31) { s_wcnt[wid] = icnt; s_wlast[wid] = ilast; }
        __syncthreads();
        int bcnt = 0, blast = -1, total_k = 0;
        for (int w = 0; w < nwarp; ++w) {
            if (w < wid) { bcnt += s_wcnt[w]; blast = max(blast, s_wlast[w]); }
            total_k += s_wcnt[w];
        }
        const int pc = __shfl_up_sync(0xffffffffu, icnt, 1), pl = __shfl_up_sync(0xffffffffu, ilast, 1);
        const int excl_cnt = bcnt + (lane > 0 ? pc : 0);
        const int excl_last = max(blast, lane > 0 ? pl : -1);
        int prev = excl_last;
        bool maybe_rot = false;
        for (int i = c0; i < c1; ++i) {
            if (all_defl || rho * fabs(zsrt[i]) <= tol) continue;
            if (prev >= 0) {
                const double zp = zsrt[prev], zi = zsrt[i];
                const double t = dsrt[i] - dsrt[prev];
                if (fabs(t * zi * zp) <= 1.000001 * tol * fma(zi, zi, zp * zp)) maybe_rot = true;
            }
            prev = i;
        }
        if (maybe_rot) s_fast = 0;
        __syncthreads();
        if (s_fast) {
            int kpos = excl_cnt, dpos = c0 - excl_cnt;
            for (int i = c0; i < c1; ++i) {
                if (!all_defl && !(rho * fabs(zsrt[i]) <= tol)) {
                    a.dk[off + kpos] = dsrt[i]; a.zk[off + kpos] = zsrt[i]; a.colmap[off + kpos] = csrt[i];
                    ++kpos;
                } else {
                    draw[dpos] = dsrt[i]; cdef[dpos] = csrt[i];
                    ++dpos;
                }
            }
            if (tid == 0) {
                s_k = total_k; s_ndef = nm - total_k;
                a.kArr[node] = total_k; a.nrotArr[node] = 0; a.rhoArr[node] = rho;
            }
        }
    }
    if (tid == 0 && !s_fast) {
        int k = 0, ndef = 0, nrot = 0;
        if (rho * zmax <= tol) {
            for (int i = 0; i < nm; ++i) { draw[ndef] = dsrt[i]; cdef[ndef] = csrt[i]; ++ndef; }
        } else {
            int prev = -1;
            for (int i = 0; i < nm; ++i) {
                if (rho * fabs(zsrt[i]) <= tol) {
                    draw[ndef] = dsrt[i]; cdef[ndef] = csrt[i]; ++ndef;
                    continue;
                }
                if (prev >= 0) {
                    // dlaed2's test |t c s| <= tol with c = z_i/tt, s = -z_prev/tt.  This scan is serial (one thread,
                    // up to n steps per merge), so the hypot and the two divisions are only paid when the
                    // division-free form of the same inequality, |t z_i z_prev| <= tol (z_i^2 + z_prev^2), holds
                    // (with a hair of slack; the exact test below still decides).
                    const double zp = zsrt[prev], zi = zsrt[i];
                    const double t = dsrt[i] - dsrt[prev];
                    double sv = -zp, cv = zi, tt = 1.0;
                    bool rot = false;
                    if (fabs(t * zi * zp) <= 1.000001 * tol * fma(zi, zi, zp * zp)) {
                        tt = hypot(cv, sv);
                        cv /= tt; sv /= tt;
                        rot = fabs(t * cv * sv) <= tol;
                    }
                    if (rot) {
                        // rotate (prev, i): z_prev -> 0 (deflated), z_i -> tt
                        zsrt[i] = tt;
                        a.rotA[off + nrot] = csrt[prev]; a.rotB[off + nrot] = csrt[i];
                        a.rotC[off + nrot] = cv; a.rotS[off + nrot] = sv;
                        ++nrot;
                        const double dp = dsrt[prev] * cv * cv + dsrt[i] * sv * sv;
                        dsrt[i] = dsrt[prev] * sv * sv + dsrt[i] * cv * cv;
                        draw[ndef] = dp; cdef[ndef] = csrt[prev]; ++ndef;
                    } else {
                        a.dk[off + k] = dsrt[prev]; a.zk[off + k] = zsrt[prev]; a.colmap[off + k] = csrt[prev];
                        ++k;
                    }
                }
                prev = i;
            }
            if (prev >= 0) {
                a.dk[off + k] = dsrt[prev]; a.zk[off + k] = zsrt[prev]; a.colmap[off + k] = csrt[prev];
                ++k;
            }
        }
        s_k = k; s_ndef = ndef;
        a.kArr[node] = k; a.nrotArr[node] = nrot; a.rhoArr[node] = rho;
    }
    __syncthreads();
    const int k = s_k, ndef = s_ndef;
    for (int t = tid; t < ndef; t += nt) {
        dnew[off + k + t] = draw[t];
        a.colmap[off + k + t] = cdef[t];
    }
}

__global__ void __launch_bounds__(256)
dc_rotate_kernel(int n, int depth, double* __restrict__ Q, int ldq, DcArrays a) {
    const int node = blockIdx.y;
    const int off = node_start(n, depth, node), end = node_start(n, depth, node + 1);
    const int row = off + blockIdx.x * 256 + threadIdx.x;
    if (row >= end) return;
    const int nrot = a.nrotArr[node];
    for (int t = 0; t < nrot; ++t) {
        const int ca = a.rotA[off + t], cb = a.rotB[off + t];
        const double c = a.rotC[off + t], s = a.rotS[off + t];
        const double qa = Q[(size_t)row + (size_t)ca * ldq], qb = Q[(size_t)row + (size_t)cb * ldq];
        Q[(size_t)row + (size_t)ca * ldq] = c * qa + s * qb;
        Q[(size_t)row + (size_t)cb * ldq] = -s * qa + c * qb;
    }
}

// secular function  f(mu) = 1 + rho * sum_i z_i^2 / ((d_i - d_org) - mu), evaluated by one warp
__device__ __forceinline__ double secular_f(const double* __restrict__ dk, const double* __restrict__ zk, int k,
                                            double rho, double dorg, double mu, int lane) {
    double s = 0.0;
    for (int i = lane; i < k; i += 32) {
        const double z = zk[i];
        s += (z * z) / ((dk[i] - dorg) - mu);
    }
    s = warp_sum(s);
    return fma(rho, s, 1.0);
}

__global__ void __launch_bounds__(256)
dc_secular_kernel(int n, int depth, double* __restrict__ dnew, DcArrays a) {
    const int node = blockIdx.y, lane = threadIdx.x & 31;
    const int off = node_start(n, depth, node);
    const int k = a.kArr[node];
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j >= k) return;
    const double rho = a.rhoArr[node];
    const double* dk = a.dk + off;
    const double* zk = a.zk + off;
    int org;
    double sign, hi;
    if (j < k - 1) {
        const double half = 0.5 * (dk[j + 1] - dk[j]);
        const double fmid = secular_f(dk, zk, k, rho, dk[j], half, lane);
        if (fmid > 0.0) { org = j; sign = 1.0; } else { org = j + 1; sign = -1.0; }
        hi = half;
    } else {
        double s = 0.0;
        for (int i = lane; i < k; i += 32) s = fma(zk[i], zk[i], s);
        s = warp_sum(s);
        org = k - 1; sign = 1.0;
        hi = rho * s * (1.0 + 16.0 * kUnitRoundoff) + 1e-300;
    }
    const double dorg = dk[org];
    // Root of V(aa) = sign * f(sign * aa) on (0, hi], V increasing, V(0+) = -inf, V(hi) >= 0; the result is the bracket's upper end
    // once its two ends are ADJACENT bit patterns, as with plain bisection on the pattern (62 evaluations of k divisions each).
    // Newton steps from the last evaluated point (f' comes out of the same pass: one more multiply-add per pole) are taken
    // whenever they land inside the bracket, the pattern midpoint otherwise and on every third step -- the bracket closes in
    // 10-20 evaluations instead of 62 and the answer keeps the last-bit quality the Gu-Eisenstat z-hat needs.
    long long lo_i = 0, hi_i = __double_as_longlong(hi);
    double xv = 0.0, fv = 0.0, dv = 0.0;
    bool have = false;
    for (int it = 0; hi_i - lo_i > 1; ++it) {
        long long cand_i = lo_i + ((hi_i - lo_i) >> 1);       // no overflow: patterns of |mu| >= 2 exceed 2^62
        if (have && (it % 3) != 2) {
            const double xn = xv - fv / dv;
            if (xn > 0.0 && xn == xn) {
                const long long ni = __double_as_longlong(xn);
                cand_i = ni <= lo_i ? lo_i + 1 : (ni >= hi_i ? hi_i - 1 : ni);
            }
        }
        const double aa = __longlong_as_double(cand_i);
        double s0 = 0.0, s1 = 0.0;
        for (int i = lane; i < k; i += 32) {
            const double z = zk[i];
            const double q = z / ((dk[i] - dorg) - sign * aa);
            s0 = fma(z, q, s0);
            s1 = fma(q, q, s1);
        }
        s0 = warp_sum(s0);
        s1 = warp_sum(s1);
        xv = aa;
        fv = sign * fma(rho, s0, 1.0);
        dv = rho * s1;
        have = dv > 0.0 && dv == dv && fv == fv && fabs(fv) < 1e300;
        if (fv < 0.0) lo_i = cand_i; else hi_i = cand_i;
    }
    const double mu = sign * __longlong_as_double(hi_i);
    if (lane == 0) {
        a.org[off + j] = org;
        a.mu[off + j] = mu;
        dnew[off + j] = dorg + mu;
    }
}

__global__ void __launch_bounds__(256)
dc_zhat_kernel(int n, int depth, DcArrays a) {
    const int node = blockIdx.y, lane = threadIdx.x & 31;
    const int off = node_start(n, depth, node);
    const int k = a.kArr[node];
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= k) return;
    const double* dk = a.dk + off;
    const double* mu = a.mu + off;
    const int* org = a.org + off;
    const double di = dk[i];
    double prod = 1.0;
    for (int j = lane; j < k; j += 32) {
        const double num = (dk[org[j]] - di) + mu[j];     // lambda_j - d_i
        prod *= (j == i) ? num : num / (dk[j] - di);
    }
    prod = warp_prod(prod);
    if (lane == 0) a.zh[off + i] = copysign(sqrt(fabs(prod) / a.rhoArr[node]), a.zk[off + i]);
}

__global__ void __launch_bounds__(256)
dc_smat_kernel(int n, int depth, double* __restrict__ S, int lds, DcArrays a) {
    const int node = blockIdx.y, lane = threadIdx.x & 31;
    const int off = node_start(n, depth, node);
    const int k = a.kArr[node];
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j >= k) return;
    const double* dk = a.dk + off;
    const double* zh = a.zh + off;
    const double dorg = dk[a.org[off + j]], mu = a.mu[off + j];
    double* col = S + (size_t)off + (size_t)(off + j) * lds;
    double nrm = 0.0;
    for (int i = lane; i < k; i += 32) {
        const double val = zh[i] / ((dk[i] - dorg) - mu);     // zh_i / (d_i - lambda_j)
        col[i] = val;
        nrm = fma(val, val, nrm);
    }
    nrm = warp_sum(nrm);
    const double inv = 1.0 / sqrt(nrm);
    for (int i = lane; i < k; i += 32) col[i] *= inv;
}

// Qn[off+row, off+j] = sum_{i<k} Q[off+row, colmap[off+i]] * S[off+i, off+j]   (row < nm, j < k)
constexpr int DBM = 64, DBN = 64, DBK = 16;
__global__ void __launch_bounds__(256)
dc_gemm_kernel(int n, int depth, const double* __restrict__ Q, int ldq, const double* __restrict__ S, int lds,
               double* __restrict__ Qn, DcArrays a) {
    const int node = blockIdx.z;
    const int off = node_start(n, depth, node), end = node_start(n, depth, node + 1);
    const int nm = end - off;
    const int k = a.kArr[node];
    const int i0 = blockIdx.x * DBM, j0 = blockIdx.y * DBN;
    if (i0 >= nm || j0 >= k) return;
    // 64 x 64 tile on the FP64 tensor pipe: 8 warps as 4 (M) x 2 (N), warp tile 16 x 32 = 2 x 4 DMMA tiles
    constexpr int DLD = DBM + 4;         // k-major rows, stride % 16 == 4: the fragment loads of a half-warp are conflict-free
    __shared__ double As[DBK][DLD];
    __shared__ double Bs[DBK][DLD];
    __shared__ int cols[DBK];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tq = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;
    double acc[2][4][2];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int vv = 0; vv < 4; ++vv) acc[u][vv][0] = acc[u][vv][1] = 0.0;
    for (int k0 = 0; k0 < k; k0 += DBK) {
        if (tid < DBK) cols[tid] = (k0 + tid < k) ? a.colmap[off + k0 + tid] : -1;
        __syncthreads();
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int idx = tid + l * 256;
            const int i = idx & 63, kk = idx >> 6;
            const int gi = i0 + i, c = cols[kk];
            As[kk][i] = (gi < nm && c >= 0) ? Q[(size_t)(off + gi) + (size_t)c * ldq] : 0.0;
        }
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int idx = tid + l * 256;
            const int kk = idx & 15, j = idx >> 4;
            const int gj = j0 + j, gk = k0 + kk;
            Bs[kk][j] = (gj < k && gk < k) ? S[(size_t)(off + gk) + (size_t)(off + gj) * lds] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k4 = 0; k4 < DBK / 4; ++k4) {
            double av[2], bv[4];
#pragma unroll
            for (int u = 0; u < 2; ++u) av[u] = As[k4 * 4 + tq][wm * 16 + u * 8 + g];
#pragma unroll
            for (int vv = 0; vv < 4; ++vv) bv[vv] = Bs[k4 * 4 + tq][wn * 32 + vv * 8 + g];
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int vv = 0; vv < 4; ++vv) dmma884(acc[u][vv][0], acc[u][vv][1], av[u], bv[vv]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int gi = i0 + wm * 16 + u * 8 + g;
        if (gi >= nm) continue;
#pragma unroll
        for (int vv = 0; vv < 4; ++vv)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gj = j0 + wn * 32 + vv * 8 + 2 * tq + e;
                if (gj < k) Qn[(size_t)(off + gi) + (size_t)(off + gj) * ldq] = acc[u][vv][e];
            }
    }
}

// deflated columns are copied through: Qn[:, off+t] = Q[:, colmap[off+t]] for t in [k, nm)
__global__ void __launch_bounds__(128)
dc_copy_deflated_kernel(int n, int depth, const double* __restrict__ Q, int ldq, double* __restrict__ Qn, DcArrays a) {
    const int node = blockIdx.z;
    const int off = node_start(n, depth, node), end = node_start(n, depth, node + 1);
    const int nm = end - off;
    const int k = a.kArr[node];
    const int row = blockIdx.x * 128 + threadIdx.x;
    if (row >= nm) return;
    for (int t = k + blockIdx.y; t < nm; t += gridDim.y) {
        const int c = a.colmap[off + t];
        Qn[(size_t)(off + row) + (size_t)(off + t) * ldq] = Q[(size_t)(off + row) + (size_t)c * ldq];
    }
}

// ---------------------------------------------------------------------------------------------
// rank[j] = position of lam[j] in descending order (ties by index); *count = #(lam >= eps)  (zeroed by the caller)
__global__ void dc_rank_count_kernel(const double* __restrict__ lam, int n, double eps, int* __restrict__ order,
                                     int* __restrict__ count) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double lj = lam[j];
    int rank = 0;
    for (int l = 0; l < n; ++l) {
        const double ll = lam[l];
        rank += (ll > lj) || (ll == lj && l < j);
    }
    order[j] = rank;
    if (lj >= eps) atomicAdd(count, 1);
}

// gdot[j] = v_{j-1} . v_j  (rows j+1..n-1), j >= 1: lets the back-transformation apply two reflectors per pass
__global__ void reflector_dots_kernel(const double* __restrict__ A, int lda, int n, double* __restrict__ gdot) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (j < 1 || j > n - 3) return;
    const double* v0 = A + (size_t)(j - 1) * lda;
    const double* v1 = A + (size_t)j * lda;
    double acc = 0.0;
    for (int i = j + 1 + lane; i < n; i += 32) acc = fma(v0[i], v1[i], acc);
    acc = warp_sum(acc);
    if (lane == 0) gdot[j] = acc;
}

// collist[rank] = column whose eigenvalue has that rank (descending), for the back-transformation's work list
__global__ void dc_collist_kernel(const int* __restrict__ order, int n, int* __restrict__ collist) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) collist[order[j]] = j;
}

// U = H_0 H_1 ... H_{n-3} Z in place.  One warp per eigenvector column (the column lives in shared memory), two
// reflectors per pass:  z <- H_{j-1} H_j z  needs  d1 = v_j.z,  d2 = v_{j-1}.z - tau_j d1 (v_{j-1}.v_j)  -- both dots in
// one sweep over z, one shuffle reduction, one update sweep.  The reflector pair of a pass is staged ONCE per CTA in
// shared memory (cp.async, three buffers: the next two pairs stream in while this one is applied) and shared by all its
// warps; with every warp fetching the reflectors itself the kernel was bound by L2 traffic (columns x n^2 x 8 B =
// 21 GB at n = 1600).  Only the columns the caller needs are transformed -- the work list holds the columns by
// eigenvalue rank, the first min(vec_limit, *count) of them are taken.
__global__ void backtransform_kernel(const double* __restrict__ A, int lda, const double* __restrict__ tau,
                                     const double* __restrict__ gdot, int n, double* __restrict__ Z, int ldz,
                                     const int* __restrict__ collist, const int* __restrict__ count, int vec_limit) {
    extern __shared__ double zsm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int lim = vec_limit >= 0 ? min(vec_limit, *count) : n;
    if (blockIdx.x * wpb >= lim) return;                        // whole CTA idle (uniform)
    const int idx = blockIdx.x * wpb + warp;
    const bool active = idx < lim;
    const int col = active ? collist[idx] : 0;
    double* z = zsm + (size_t)warp * n;
    double* vb = zsm + (size_t)wpb * n;                           // [NBUF buffers][2 reflectors][n]
    double* zc = Z + (size_t)col * ldz;
    if (active)
        for (int i = lane; i < n; i += 32) z[i] = zc[i];
    // stage reflectors j (rows j+1..) and j-1 (rows j..) of a pass into buffer `buf`; always commits a group so that
    // the wait below can count groups
    auto stage = [&](int j, int buf) {
        if (j >= 1) {
            double* s1 = vb + (size_t)(2 * buf) * n;
            double* s0 = s1 + n;
            const double* v1 = A + (size_t)j * lda;
            const double* v0 = A + (size_t)(j - 1) * lda;
            for (int i = j + 1 + threadIdx.x; i < n; i += blockDim.x) {
                __pipeline_memcpy_async(s1 + i, v1 + i, sizeof(double));
                __pipeline_memcpy_async(s0 + i, v0 + i, sizeof(double));
            }
        }
        __pipeline_commit();
    };
    constexpr int NBUF = 3;                                       // passes in flight: this one + two being fetched
    int j = n - 3;
    int buf = 0;
    stage(j, 0);
    stage(j - 2, 1);
    for (; j >= 1; j -= 2) {
        stage(j - 4, (buf + 2) % NBUF);
        __pipeline_wait_prior(2);                                 // the group of pass j has landed
        __syncthreads();
        if (active) {
            const double t1 = tau[j], t0 = tau[j - 1];
            const double* v1 = vb + (size_t)(2 * buf) * n;       // rows j+1..
            const double* v0 = v1 + n;                           // rows j+1.. (v0[j] == 1 is implicit)
            double d1 = 0.0, d0 = 0.0;
            for (int i = j + 1 + lane; i < n; i += 32) {
                const double zi = z[i];
                d1 = fma(v1[i], zi, d1);
                d0 = fma(v0[i], zi, d0);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                d1 += __shfl_xor_sync(0xffffffffu, d1, o);
                d0 += __shfl_xor_sync(0xffffffffu, d0, o);
            }
            const double zj = z[j];
            const double s1 = t1 * d1;
            const double s0 = t0 * ((d0 + zj) - s1 * gdot[j]);
            for (int i = j + 1 + lane; i < n; i += 32) z[i] = fma(-s0, v0[i], fma(-s1, v1[i], z[i]));
            if (lane == 0) z[j] = zj - s0;
        }
        __syncthreads();                                          // buffer `buf` may be refilled by the next stage()
        buf = (buf + 1) % NBUF;
    }
    if (j == 0 && active) {
        const double t = tau[0];
        const double* v = A;
        double dot = 0.0;
        for (int i = 1 + lane; i < n; i += 32) dot = fma(v[i], z[i], dot);
        dot = warp_sum(dot) * t;
        for (int i = 1 + lane; i < n; i += 32) z[i] = fma(-dot, v[i], z[i]);
        __syncwarp();
    }
    if (active)
        for (int i = lane; i < n; i += 32) zc[i] = z[i];
}

// ---------------------------------------------------------------------------------------------
// Blocked back-transformation (compact WY, LAPACK dlarft/dlarfb forward-columnwise) on the FP64 tensor pipe.
// Reflectors are taken 32 at a time:  H_{j0} ... H_{j0+31} = I - V T V^T  with V the 32 reflector columns (unit lower
// trapezoidal, read straight from A with a mask) and T upper triangular,
//     T(i,i) = tau_i,   T(0:i, i) = -tau_i T(0:i,0:i) (V(:,0:i)^T v_i).
//   bt_tfactor_kernel   one CTA per block: S = V^T V, then the T recursion (one warp)
//   bt_wy_kernel        one CTA per 8 eigenvector columns (one DMMA n-tile), the columns live in shared memory;
//                       per block, last to first:  W = V^T Z (DMMA, K = rows, split over the warps),  W <- T W,
//                       Z -= V W (DMMA, K = 32).
// Against the per-reflector kernel above (bound by shared-memory traffic: 7 accesses per element and reflector pair)
// every reflector element is now used for 8 columns per load.
constexpr int kWyB = 32;

__global__ void __launch_bounds__(256)
bt_tfactor_kernel(const double* __restrict__ A, int lda, int n, const double* __restrict__ tau, int nrefl,
                  double* __restrict__ Tall) {
    __shared__ double Vc[64][kWyB + 1];
    __shared__ double Ss[kWyB][kWyB + 1];
    __shared__ double Ts[kWyB][kWyB + 1];
    const int kb = blockIdx.x, tid = threadIdx.x;
    const int j0 = kb * kWyB;
    const int nb = min(kWyB, nrefl - j0);
    const int ta = tid >> 3, tb0 = (tid & 7) * 4;           // thread -> S[ta][tb0 .. tb0+3]
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int r0 = j0 + 1; r0 < n; r0 += 64) {
        __syncthreads();
        for (int e = tid; e < 64 * kWyB; e += 256) {
            const int r = e & 63, c = e >> 6;
            const int row = r0 + r, j = j0 + c;
            Vc[r][c] = (row < n && c < nb && row >= j + 1) ? A[(size_t)row + (size_t)j * lda] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < 64; ++r) {
            const double va = Vc[r][ta];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] = fma(va, Vc[r][tb0 + q], acc[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) Ss[ta][tb0 + q] = acc[q];
    for (int e = tid; e < kWyB * (kWyB + 1); e += 256) (&Ts[0][0])[e] = 0.0;
    __syncthreads();
    if (tid < 32) {
        const int r = tid;
        for (int i = 0; i < nb; ++i) {
            const double ti = tau[j0 + i];
            double v = 0.0;
            if (r < i) {
                for (int q = r; q < i; ++q) v = fma(Ts[r][q], Ss[q][i], v);
                v *= -ti;
            } else if (r == i) {
                v = ti;
            }
            __syncwarp();
            if (r <= i) Ts[r][i] = v;
            __syncwarp();
        }
    }
    __syncthreads();
    double* out = Tall + (size_t)kb * kWyB * kWyB;
    for (int e = tid; e < kWyB * kWyB; e += 256) out[e] = Ts[e >> 5][e & 31];
}

__global__ void __launch_bounds__(256, 1)
bt_wy_kernel(const double* __restrict__ A, int lda, int n, const double* __restrict__ Tall, int nblocks, int nrefl,
             double* __restrict__ Z, int ldz, const int* __restrict__ collist, const int* __restrict__ count, int vec_limit) {
    extern __shared__ double wsm[];
    double* Zs = wsm;                         // n x 8, row-major
    double* Wp = Zs + (size_t)n * 8;          // 8 warps x 32 x 8 partials of V^T Z
    double* Ws = Wp + 8 * kWyB * 8;           // 32 x 8
    double* W2 = Ws + kWyB * 8;               // 32 x 8   (T W)
    double* Ts = W2 + kWyB * 8;               // 32 x 33
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tq = lane & 3;
    const int lim = vec_limit >= 0 ? min(vec_limit, *count) : n;
    if (blockIdx.x * 8 >= lim) return;
    int cols[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int idx = blockIdx.x * 8 + c;
        cols[c] = idx < lim ? collist[idx] : -1;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c)
        for (int row = tid; row < n; row += 256) Zs[(size_t)row * 8 + c] = cols[c] >= 0 ? Z[(size_t)row + (size_t)cols[c] * ldz] : 0.0;
    __syncthreads();
    for (int kb = nblocks - 1; kb >= 0; --kb) {
        const int j0 = kb * kWyB;
        const int nb = min(kWyB, nrefl - j0);
        const int r_lo = j0 + 1;                                   // first row any reflector of the block touches
        for (int e = tid; e < kWyB * kWyB; e += 256) Ts[(e >> 5) * (kWyB + 1) + (e & 31)] = Tall[(size_t)kb * kWyB * kWyB + e];
        // ---- phase 1: W = V^T Z, k-steps of 4 rows (aligned to 4) dealt round-robin to the warps
        {
            double acc[4][2];
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u][0] = acc[u][1] = 0.0;
            const int rs = r_lo & ~3;
#pragma unroll 4
            for (int rb = rs + 4 * warp; rb < n; rb += 32) {
                const int row = rb + tq;
                const bool rv = row < n && row >= r_lo;
                const double bfr = rv ? Zs[(size_t)row * 8 + g] : 0.0;
                double afr[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int c = 8 * u + g;
                    afr[u] = (rv && c < nb && row >= j0 + c + 1) ? A[(size_t)row + (size_t)(j0 + c) * lda] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) dmma884(acc[u][0], acc[u][1], afr[u], bfr);
            }
            double* wp = Wp + warp * (kWyB * 8);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                wp[(8 * u + g) * 8 + 2 * tq] = acc[u][0];
                wp[(8 * u + g) * 8 + 2 * tq + 1] = acc[u][1];
            }
        }
        __syncthreads();
        {
            double w = 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) w += Wp[q * (kWyB * 8) + tid];
            Ws[tid] = w;
        }
        __syncthreads();
        // ---- phase 2: W2 = -(T W)   (T upper triangular; the sign folds the subtraction of phase 3 into the DMMA)
        {
            const int r = tid >> 3, c = tid & 7;
            double v = 0.0;
            for (int q = r; q < nb; ++q) v = fma(Ts[r * (kWyB + 1) + q], Ws[q * 8 + c], v);
            W2[tid] = -v;
        }
        __syncthreads();
        // ---- phase 3: Z += V W2, m-tiles of 8 rows (aligned to 8) dealt round-robin to the warps
        {
            const int rs = r_lo & ~7;
#pragma unroll 4
            for (int rb = rs + 8 * warp; rb < n; rb += 64) {
                const int row = rb + g;
                const bool rv = row < n && row >= r_lo;
                double c0 = 0.0, c1 = 0.0;
                if (rv) { c0 = Zs[(size_t)row * 8 + 2 * tq]; c1 = Zs[(size_t)row * 8 + 2 * tq + 1]; }
#pragma unroll
                for (int kk = 0; kk < kWyB / 4; ++kk) {
                    const int c = 4 * kk + tq;
                    const double afr = (rv && c < nb && row >= j0 + c + 1) ? A[(size_t)row + (size_t)(j0 + c) * lda] : 0.0;
                    const double bfr = W2[c * 8 + g];
                    dmma884(c0, c1, afr, bfr);
                }
                if (rv) { Zs[(size_t)row * 8 + 2 * tq] = c0; Zs[(size_t)row * 8 + 2 * tq + 1] = c1; }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int c = 0; c < 8; ++c)
        if (cols[c] >= 0)
            for (int row = tid; row < n; row += 256) Z[(size_t)row + (size_t)cols[c] * ldz] = Zs[(size_t)row * 8 + c];
}

__global__ void dc_check_kernel(const double* __restrict__ U, int ldu, int n, int* __restrict__ fail) {
    // column norms must be 1 to ~1e-8 and finite (cheap sanity check of the whole pipeline)
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const double* c = U + (size_t)warp * ldu;
    double acc = 0.0;
    for (int i = lane; i < n; i += 32) acc = fma(c[i], c[i], acc);
    acc = warp_sum(acc);
    if (lane == 0 && !(fabs(acc - 1.0) < 1e-8)) atomicExch(fail, 2);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
void EigWorkspace::phase_mark(cudaStream_t s) {
    if (!phase_on) return;
    if (pev_used == (int)pev.size()) {
        cudaEvent_t e;
        NLE_CUDA(cudaEventCreate(&e));
        pev.push_back(e);
    }
    NLE_CUDA(cudaEventRecord(pev[pev_used++], s));
}
void EigWorkspace::phase_ms(double out[3]) {
    out[0] = out[1] = out[2] = 0.0;
    for (int i = 0; i + 3 < pev_used; i += 4)
        for (int k = 0; k < 3; ++k) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, pev[i + k], pev[i + k + 1]) == cudaSuccess) out[k] += ms; else cudaGetLastError();
        }
    pev_used = 0;
    phase_on = false;
}
EigWorkspace::~EigWorkspace() {
    for (cudaEvent_t e : pev) cudaEventDestroy(e);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (side) cudaStreamDestroy(side);
}

void EigWorkspace::side_init() {
    int dev = 0;
    NLE_CUDA(cudaGetDevice(&dev));
    if (side && side_dev == dev) return;
    if (ev_fork) { cudaEventDestroy(ev_fork); ev_fork = nullptr; }
    if (ev_join) { cudaEventDestroy(ev_join); ev_join = nullptr; }
    if (side) { cudaStreamDestroy(side); side = nullptr; }
    NLE_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));     // must not synchronise with the legacy default stream
    NLE_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    NLE_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    side_dev = dev;
}

void EigWorkspace::reserve_dc(int n) {
    if (n <= dc_cap) return;
    const size_t nn = (size_t)n * n;
    Qa.alloc(nn);
    Qb.alloc(nn);
    Sb.alloc(nn);
    dcd.alloc(16 * (size_t)n + 64);
    dci.alloc(8 * (size_t)n + 64);
    dc_cap = n;
}

// Eigen-decomposition of the symmetric matrix As (n x n, ld n, full storage; DESTROYED).
// Results: eigenvalues (unsorted) in *lam_out, eigenvectors in the columns of *vec_out (ld n); order[j] = rank
// of eigenvalue j in descending order, *count = #(eigenvalues >= eps).  vec_limit >= 0: only the eigenvectors
// ranked below min(vec_limit, *count) are back-transformed (the others are left in the tridiagonal basis).
// Returns false if a device-side sanity check failed (caller falls back to Jacobi).
bool sym_eig_dc_core(double* As, int n, double eps, int vec_limit, int* order, int* count, EigWorkspace& ws,
                     cudaStream_t s, double** lam_out, double** vec_out) {
    ws.reserve_dc(n);
    // developer timing (NLE_B200_EIG_PROF=1): wall-clock per phase with the stream drained
    const bool prof = getenv("NLE_B200_EIG_PROF") != nullptr;
    auto tnow = [&]() { if (prof) cudaStreamSynchronize(s); return std::chrono::steady_clock::now(); };
    auto t_start = tnow();
    ws.phase_mark(s);
    double* dd = ws.dcd.p;
    double* d0 = dd;                 // tridiagonal diagonal
    double* e0 = dd + n;             // off-diagonal
    double* tau = dd + 2 * (size_t)n;
    double* pbuf = dd + 3 * (size_t)n;   // 2n
    double* dA = dd + 5 * (size_t)n;     // eigenvalue ping
    double* dB = dd + 6 * (size_t)n;     // eigenvalue pong
    DcArrays a;
    a.dk = dd + 7 * (size_t)n;
    a.zk = dd + 8 * (size_t)n;
    a.rotC = dd + 9 * (size_t)n;
    a.rotS = dd + 10 * (size_t)n;
    a.mu = dd + 11 * (size_t)n;
    a.zh = dd + 12 * (size_t)n;
    a.rhoArr = dd + 13 * (size_t)n;
    int* di = ws.dci.p;
    a.colmap = di;
    a.rotA = di + n;
    a.rotB = di + 2 * (size_t)n;
    a.org = di + 3 * (size_t)n;
    a.kArr = di + 4 * (size_t)n;
    a.nrotArr = di + 5 * (size_t)n;
    int* fail = di + 6 * (size_t)n;
    NLE_CUDA(cudaMemsetAsync(fail, 0, sizeof(int), s));

    // ---- 1. tridiagonalisation
    // Default for n >= 64: tridiag_cluster_kernel, clusters of 2 CTAs, on the largest trailing block whose columns fit in
    // shared memory (the whole matrix up to n ~ 1700; before that, tridiag_kernel in partial mode); falls through to
    // tridiag_kernel when the cluster launch is refused.  NLE_B200_TRD=gridsync forces tridiag_kernel alone (the
    // bit-equality cross-check of tests/test_gpu_eig_variants.py), NLE_B200_TRD=nohybrid the round-2 rule (no hand-over:
    // tridiag_kernel for every matrix that does not fit; timing comparisons); NLE_B200_TRD_PROF=1 prints the per-phase
    // cycle counts of a step on stderr.
    static const bool force_gridsync = [] { const char* e = getenv("NLE_B200_TRD"); return e && std::string(e) == "gridsync"; }();
    int dev = 0, max_smem_trd = 0;
    NLE_CUDA(cudaGetDevice(&dev));
    NLE_CUDA(cudaDeviceGetAttribute(&max_smem_trd, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    constexpr int S = 2;
    // shared memory of tridiag_cluster_kernel for an m x m problem on a grid of Gt CTAs: landing cells (4 m), w, v, cn,
    // reduction slots + mbarrier, the resident columns
    auto cluster_smem = [&](int m, int Gt) {
        return (7 * (size_t)m + 8 + kTrdRed + 2 + (size_t)cdiv(m, Gt) * m) * sizeof(double);
    };
    // grid.sync kernel on the block (Ab, lda = n, m): reflectors [0, nstop) only when nstop < m - 2
    auto launch_gridsync = [&](double* Ab, int m, double* db, double* eb, double* taub, int nstop) {
        const void* kfn = (const void*)tridiag_kernel;
        allow_max_dynamic_smem((const void*)kfn);
        int grid = std::min(sm_count(), m);
        // as many of a CTA's columns as fit stay in shared memory for the whole launch (the highest ones)
        const size_t base = (3 * (size_t)m + kTrdRed) * sizeof(double);
        int qs = 0;
        if ((size_t)max_smem_trd > base) qs = (int)std::min<size_t>((size_t)cdiv(m, grid), ((size_t)max_smem_trd - base) / ((size_t)m * sizeof(double)));
        static const bool no_cache = getenv("NLE_B200_TRD_NOCACHE") != nullptr;     // timing comparisons
        if (no_cache || force_gridsync) qs = 0;       // the cross-check path is the plain L2-resident kernel
        size_t smem = base + (size_t)qs * m * sizeof(double);
        int per_sm = 0;
        NLE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, kTrdThreads, smem));
        if (per_sm < 1) throw Unsupported{"eigensolver: tridiagonalisation kernel does not fit on an SM (n=" + std::to_string(m) + ")"};
        int lda = n;
        void* args[] = {&Ab, &lda, &m, &db, &eb, &taub, &pbuf, &nstop, &qs};
        NLE_CUDA(cudaLaunchCooperativeKernel(kfn, dim3(grid), dim3(kTrdThreads), args, smem, s));
        ++g_launches;
    };
    // resident kernel on the block (Ab, lda = n, m); false if it does not fit or the cluster launch is refused
    auto launch_cluster = [&](double* Ab, int m, double* db, double* eb, double* taub) -> bool {
        if (m < 64 || m > kResPer * kTrdThreads) return false;
        const bool kprof = getenv("NLE_B200_TRD_PROF") != nullptr;
        const void* kfn = kprof ? (const void*)tridiag_cluster_kernel<true> : (const void*)tridiag_cluster_kernel<false>;
        // the grid depends on how many clusters fit, which depends on the shared memory, which depends on the grid:
        // start from one CTA per SM and shrink until the launch configuration is consistent.  The occupancy query is a
        // slow host call (the GPU idles meanwhile): its answer is cached per (device, m).
        struct Cfg { int dev, n, G; size_t smem; };
        static thread_local Cfg cache[8] = {};
        static thread_local int cache_next = 0;
        int G = 0;
        size_t smem = 0;
        for (const Cfg& c : cache)
            if (c.G > 0 && c.dev == dev && c.n == m) { G = c.G; smem = c.smem; }
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = S; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeCooperative;     // only for its co-residency guarantee
        at[1].val.cooperative = 1;
        cfg.blockDim = dim3(kTrdThreads);
        cfg.stream = s;
        cfg.attrs = at;
        if (G == 0) {
            int Gt = (std::min(sm_count(), m) / S) * S;
            for (int it = 0; it < 4 && Gt >= S; ++it) {
                const size_t sm = cluster_smem(m, Gt);
                if (sm > (size_t)max_smem_trd) break;
                allow_max_dynamic_smem((const void*)kfn);
                cfg.gridDim = dim3(Gt);
                cfg.dynamicSmemBytes = sm;
                cfg.numAttrs = 1;
                int ncl = 0;
                if (cudaOccupancyMaxActiveClusters(&ncl, kfn, &cfg) != cudaSuccess) { cudaGetLastError(); break; }
                if (ncl * S < Gt) { Gt = ncl * S; continue; }      // fewer clusters fit: retry with the smaller grid
                G = Gt;
                smem = sm;
                cache[cache_next] = Cfg{dev, m, G, smem};
                cache_next = (cache_next + 1) % 8;
                break;
            }
        }
        if (G < S) return false;
        allow_max_dynamic_smem((const void*)kfn);
        cfg.gridDim = dim3(G);
        cfg.dynamicSmemBytes = smem;
        cfg.numAttrs = 2;
        const size_t ne = ((size_t)m + 1) & ~(size_t)1;  // even stride per vector of cells (32-byte sector polls)
        const size_t cells = 4 * ne + 8;                 // + 8 cells = 16 profile counters
        if (ws.trdll.n < cells) ws.trdll.alloc(cells);
        NLE_CUDA(cudaMemsetAsync(ws.trdll.p, 0, cells * sizeof(uint4), s));   // tag 0 = never written
        int lda = n;
        uint4* ll = ws.trdll.p;
        long long* kp = reinterpret_cast<long long*>(ws.trdll.p + 4 * ne);
        void* args[] = {&Ab, &lda, &m, &db, &eb, &taub, &ll, &kp};
        // Without the co-residency guarantee of a cooperative launch the polling CTAs could wait for CTAs that are
        // not running: if the launch is refused, the caller falls through to the grid.sync kernel instead of launching anyway.
        if (cudaLaunchKernelExC(&cfg, kfn, args) != cudaSuccess) { cudaGetLastError(); return false; }
        ++g_launches;
        if (kprof) {
            long long h[16];
            NLE_CUDA(cudaMemcpyAsync(h, kp, sizeof(h), cudaMemcpyDeviceToHost, s));
            NLE_CUDA(cudaStreamSynchronize(s));
            const double st = (double)std::max(1LL, h[11]);
            double tot = 0;
            for (int k = 0; k < 10; ++k) tot += (double)h[k];
            fprintf(stderr, "[trd cluster S=%d G=%d n=%d] cycles/step: barrier + multicast in flight %.0f | validation + re-polls %.0f | "
                    "w, column, two reductions, reflector %.0f | update+symv %.0f | tail %.0f | total %.0f ; re-polls of thread 0 per step %.2f\n",
                    S, G, m, h[0] / st, h[9] / st, h[6] / st, h[7] / st, h[8] / st, tot / st, h[10] / st);
        }
        return true;
    };
    if (force_gridsync || n < 64) {
        launch_gridsync(As, n, d0, e0, tau, n);
    } else {
        // the largest trailing block whose columns fit in the shared memory of one CTA per SM
        const int Gfull = (sm_count() / S) * S;
        int m = std::min(n, kResPer * kTrdThreads);
        while (m >= 64 && cluster_smem(m, std::min(Gfull, (m / S) * S)) > (size_t)max_smem_trd) --m;
        static const bool no_hybrid = [] { const char* e = getenv("NLE_B200_TRD"); return e && std::string(e) == "nohybrid"; }();
        if (m == n) {
            if (!launch_cluster(As, n, d0, e0, tau)) launch_gridsync(As, n, d0, e0, tau, n);
        } else if (m < 64 || n - m > n - 3 || no_hybrid) {
            launch_gridsync(As, n, d0, e0, tau, n);
        } else {
            // n too large: the first n - m reflectors on the L2-resident matrix, the last m columns in shared memory
            const int s0 = n - m;
            launch_gridsync(As, n, d0, e0, tau, s0);
            double* Ab = As + (size_t)s0 * n + s0;
            if (!launch_cluster(Ab, m, d0 + s0, e0 + s0, tau + s0)) launch_gridsync(Ab, m, d0 + s0, e0 + s0, tau + s0, m);
        }
    }
    auto t_trd = tnow();
    ws.phase_mark(s);
    // T factors of the compact-WY back-transformation: they depend on the reflectors only -> side stream, beside the D&C
    const int nrefl = n - 2;
    const int nblocks = cdiv(std::max(nrefl, 1), kWyB);
    const size_t wy_smem = ((size_t)n * 8 + 8 * kWyB * 8 + 2 * kWyB * 8 + kWyB * (kWyB + 1)) * sizeof(double);
    const bool wy_path = n >= 64 && wy_smem <= (size_t)max_smem_trd;
    if (wy_path) {
        if (ws.wyT.n < (size_t)nblocks * kWyB * kWyB) ws.wyT.alloc((size_t)nblocks * kWyB * kWyB);
        ws.side_init();
        NLE_CUDA(cudaEventRecord(ws.ev_fork, s));
        NLE_CUDA(cudaStreamWaitEvent(ws.side, ws.ev_fork, 0));
        bt_tfactor_kernel<<<nblocks, 256, 0, ws.side>>>(As, n, n, tau, nrefl, ws.wyT.p);
        NLE_LAUNCH_CHECK();
        NLE_CUDA(cudaEventRecord(ws.ev_join, ws.side));
    }
    // ---- 2. divide & conquer on (d0, e0)
    int depth = 0;
    while (((n + (1 << depth) - 1) >> depth) > kLeaf) ++depth;
    const size_t nn = (size_t)n * n;
    NLE_CUDA(cudaMemsetAsync(ws.Qa.p, 0, nn * sizeof(double), s));
    NLE_CUDA(cudaMemsetAsync(ws.Qb.p, 0, nn * sizeof(double), s));
    double* Qc = ws.Qa.p;
    double* Qn = ws.Qb.p;
    double* dc = dA;
    double* dn = dB;
    dc_leaf_kernel<<<1 << depth, 32, 0, s>>>(d0, e0, n, depth, dc, Qc, n, fail);
    NLE_LAUNCH_CHECK();
    for (int t = depth - 1; t >= 0; --t) {
        const int nmerge = 1 << t;
        const int nm_max = (n + nmerge - 1) / nmerge;
        const size_t smem = (size_t)nm_max * (3 * sizeof(double) + 2 * sizeof(int));
        allow_max_dynamic_smem((const void*)dc_setup_kernel);
        const int threads = nm_max >= 1024 ? 1024 : (nm_max >= 256 ? 256 : 64);
        dc_setup_kernel<<<nmerge, threads, smem, s>>>(n, t, e0, dc, Qc, n, dn, a);
        NLE_LAUNCH_CHECK();
        dc_rotate_kernel<<<dim3(cdiv(nm_max, 256), nmerge), 256, 0, s>>>(n, t, Qc, n, a);
        NLE_LAUNCH_CHECK();
        dc_secular_kernel<<<dim3(cdiv(nm_max, 8), nmerge), 256, 0, s>>>(n, t, dn, a);
        NLE_LAUNCH_CHECK();
        dc_zhat_kernel<<<dim3(cdiv(nm_max, 8), nmerge), 256, 0, s>>>(n, t, a);
        NLE_LAUNCH_CHECK();
        dc_smat_kernel<<<dim3(cdiv(nm_max, 8), nmerge), 256, 0, s>>>(n, t, ws.Sb.p, n, a);
        NLE_LAUNCH_CHECK();
        dc_gemm_kernel<<<dim3(cdiv(nm_max, DBM), cdiv(nm_max, DBN), nmerge), 256, 0, s>>>(n, t, Qc, n, ws.Sb.p, n, Qn, a);
        NLE_LAUNCH_CHECK();
        int ycopies = nm_max < 64 ? nm_max : 64;
        dc_copy_deflated_kernel<<<dim3(cdiv(nm_max, 128), ycopies, nmerge), 128, 0, s>>>(n, t, Qc, n, Qn, a);
        NLE_LAUNCH_CHECK();
        std::swap(Qc, Qn);
        std::swap(dc, dn);
    }
    auto t_dc = tnow();
    ws.phase_mark(s);
    // ---- 3. back-transformation (in place in Qc)
    {
        int dev = 0, max_smem = 0;
        NLE_CUDA(cudaGetDevice(&dev));
        NLE_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        NLE_CUDA(cudaMemsetAsync(count, 0, sizeof(int), s));
        dc_rank_count_kernel<<<cdiv(n, 128), 128, 0, s>>>(dc, n, eps, order, count);
        NLE_LAUNCH_CHECK();
        // number of columns to transform (host copy: it sizes the CTAs so that one wave covers all SMs)
        int count_h = 0;
        NLE_CUDA(cudaMemcpyAsync(&count_h, count, sizeof(int), cudaMemcpyDeviceToHost, s));
        NLE_CUDA(cudaStreamSynchronize(s));
        const int m = vec_limit >= 0 ? std::min(vec_limit, count_h) : n;
        // shared memory: wpb eigenvector columns + three staged reflector pairs (6 columns)
        const int max_wpb = (int)((size_t)(max_smem - 1024) / ((size_t)n * sizeof(double))) - 6;
        if (max_wpb < 1) throw Unsupported{"eigensolver: back-transformation column does not fit in shared memory (n=" + std::to_string(n) + ")"};
        int wpb = std::max(2, cdiv(std::max(m, 1), sm_count()));
        wpb = std::min(wpb, std::min(max_wpb, 16));
        const size_t smem = (size_t)(wpb + 6) * n * sizeof(double);
        allow_max_dynamic_smem((const void*)backtransform_kernel);
        int* collist = a.colmap;  // the D&C column map is free again
        dc_collist_kernel<<<cdiv(n, 128), 128, 0, s>>>(order, n, collist);
        NLE_LAUNCH_CHECK();
        double* gdot = pbuf;      // the symv exchange buffer of the tridiagonalisation is free again
        // compact-WY blocks on DMMA when the 8-column tile fits in shared memory (else the per-reflector kernel)
        if (wy_path) NLE_CUDA(cudaStreamWaitEvent(s, ws.ev_join, 0));      // the T factors (side stream) are complete
        if (m > 0 && wy_path) {
            allow_max_dynamic_smem((const void*)bt_wy_kernel);
            bt_wy_kernel<<<cdiv(m, 8), 256, wy_smem, s>>>(As, n, n, ws.wyT.p, nblocks, nrefl, Qc, n, collist, count, vec_limit);
            NLE_LAUNCH_CHECK();
        } else if (m > 0) {
            if (n >= 4) {
                reflector_dots_kernel<<<cdiv((long long)n * 32, 256), 256, 0, s>>>(As, n, n, gdot);
                NLE_LAUNCH_CHECK();
            }
            backtransform_kernel<<<cdiv(m, wpb), wpb * 32, smem, s>>>(As, n, tau, gdot, n, Qc, n, collist, count, vec_limit);
            NLE_LAUNCH_CHECK();
        }
    }
    auto t_bt = tnow();
    ws.phase_mark(s);
    if (prof) {
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "[eig_dc n=%d] tridiag %.3f ms, divide&conquer %.3f ms, back-transform %.3f ms\n", n,
                ms(t_start, t_trd), ms(t_trd, t_dc), ms(t_dc, t_bt));
    }
    dc_check_kernel<<<cdiv((long long)n * 32, 256), 256, 0, s>>>(Qc, n, n, fail);
    NLE_LAUNCH_CHECK();
    int fail_h = 0;
    NLE_CUDA(cudaMemcpyAsync(&fail_h, fail, sizeof(int), cudaMemcpyDeviceToHost, s));
    NLE_CUDA(cudaStreamSynchronize(s));
    *lam_out = dc;
    *vec_out = Qc;
    if (fail_h != 0 && getenv("NLE_B200_EIG_DEBUG")) {
        fprintf(stderr, "[eig_dc n=%d] sanity check failed, code %d (1 = leaf QL, 2 = column norm)\n", n, fail_h);
        if (getenv("NLE_B200_EIG_NOCHECK")) return true;
    }
    return fail_h == 0;
}

}  // namespace nle
