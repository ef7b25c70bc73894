// Top-k eigenpairs of a symmetric positive semi-definite matrix: Chebyshev-filtered block subspace iteration with
// Cholesky-QR and Rayleigh-Ritz on the FP64 tensor pipe.
//
// Replaces the THIRD eigensolve of the pipeline (filter.cpp:311-316: eig(Q), of which only the nEigVectors largest pairs are
// used; the reference's USE_SPECTRA build does the same with a Lanczos solver, filter.cpp:169-200).  On the r2 x r2 block
// of Q (api.cu) a full tridiagonalisation + divide & conquer pays (4/3) r2^3 flop along a latency chain of r2 Householder
// steps for the sake of k = 50 ... 100 pairs; the block method needs ~50-100 products A * X with X r2 x m (m = k + guard),
// each spread over all SMs as a split-K DMMA GEMM.
//
// Algorithm (NumPy prototype: scripts/proto_topk.py):
//   X  <- orth(random n x m);  two power steps  X <- orth(A X)
//   repeat (at most kMaxBlocks times)
//       Rayleigh-Ritz: Y = A X, H = X^T Y, H = W diag(theta) W^T (sym_eig, m x m), X <- X W, Y <- Y W,
//                      res_j = ||Y_j - theta_j X_j||;   done when max_{j<k} res_j <= kResTol * theta_0
//       kRounds times:  X <- orth( T_d((A - c)/e) X ),  [lb, cut] = [-1e-3 theta_0, theta_m] mapped to [-1, 1], the degree d
//                       chosen so that the top Ritz value is amplified at most 1e6 times more than the k-th (Cholesky-QR
//                       stays well conditioned)
//   orth = Cholesky-QR: G = X^T X (split-K GEMM), G = L L^T and L^-1 in one CTA, X <- X L^-T (GEMM).
// Deterministic: fixed start block (integer hash), fixed summation orders, no atomics.  Anything unexpected (Cholesky pivot
// <= 0, no convergence, k-th Ritz value below eps) makes the caller fall back to the full solver.
#include "common.cuh"

#include <algorithm>
#include <cmath>

namespace nle {

namespace {

constexpr int kMaxM = 128;          // block width limit (L and L^-1 of the Cholesky-QR live in one CTA's shared memory)
constexpr int kMaxBlocks = 4;       // Rayleigh-Ritz steps
constexpr int kRounds = 8;          // filter + orthonormalisation rounds between two Rayleigh-Ritz steps
constexpr int kMaxDegree = 6;
constexpr double kMaxRatio = 1e6;   // amplification of theta_0 relative to theta_k per round
constexpr double kResTol = 2e-13;   // residual bound relative to theta_0

__global__ void topk_init_kernel(double* __restrict__ X, long long count) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    unsigned long long z = (unsigned long long)e * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;   // splitmix64
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    X[e] = (double)(long long)(z >> 11) * (1.0 / 4503599627370496.0) - 1.0;   // uniform in [-1, 1)
}

// G (m x m, ld m, symmetric, lower triangle read) -> Linv = L^-1 with G = L L^T, written to `Linv` (m x m, ld m, upper
// triangle zeroed); diagL (m) = diag(L).  One CTA, 8 threads per row.  fail[0] is set when a pivot is not positive.
//   Cholesky, row-owned left-looking:  L[i][j] = (G[i][j] - sum_{l<j} L[i][l] L[j][l]) / L[j][j]
//   inverse, column-owned:             X[i][j] = -(sum_{l=j}^{i-1} L[i][l] X[l][j]) / L[i][i],  X[j][j] = 1 / L[j][j]
// Both are chains of m dependent pivots; what sits on the chain is kept short: the pivot's reciprocal square root is one
// rsqrt (no sqrt + division), rows are divided by multiplying with it, and a dot product is split over 8 lanes.
constexpr int kCholPer = 8;     // threads per row
__global__ void __launch_bounds__(kCholPer * kMaxM)
topk_chol_inv_kernel(const double* __restrict__ G, int m, double* __restrict__ Linv, double* __restrict__ diagL, int* __restrict__ fail) {
    extern __shared__ double csm[];
    // packed triangles (m = 128: 2 x 66 KB):  L(i, l), i >= l, column by column -- the threads of consecutive rows i hit
    // consecutive banks;  Linv(l, j), l >= j, row by row -- the owners of consecutive columns j hit consecutive banks
    const int tri = m * (m + 1) / 2;
    double* L = csm;
    double* X = csm + tri;
    auto Lx = [&](int i, int l) -> double& { return L[l * m - (l * (l - 1)) / 2 + (i - l)]; };
    auto Xx = [&](int l, int j) -> double& { return X[(l * (l + 1)) / 2 + j]; };
    __shared__ double rdiag[kMaxM];        // 1 / L(j, j)
    __shared__ int bad;
    const int tid = threadIdx.x, row = tid / kCholPer, part = tid % kCholPer;
    if (tid == 0) bad = 0;
    for (int e = tid; e < m * m; e += blockDim.x) {
        const int i = e % m, j = e / m;
        if (i >= j) Lx(i, j) = G[i + (size_t)j * m];
    }
    __syncthreads();
    // the threads of a row form a shuffle group of their own: trip counts differ between the rows of a warp
    const unsigned gmask = ((1u << kCholPer) - 1u) << ((tid & 31) & ~(kCholPer - 1));
    for (int j = 0; j < m; ++j) {
        double s = 0.0;
        if (row >= j) {
            // L(row, l), L(j, l) for l = part, part + 8, ...: the offset of column l (minus l) advances by m - l - 1 per column
            // four columns per trip with all eight loads in front: the loads of a trip are in flight together
            int off = part * m - (part * (part - 1)) / 2 - part;
            for (int l = part; l < j; l += 4 * kCholPer) {
                double a[4], c[4];
                int o = off, ll = l;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const bool ok = ll < j;
                    a[u] = ok ? L[o + row] : 0.0;
                    c[u] = ok ? L[o + j] : 0.0;
                    o += kCholPer * (m - 1 - ll) - (kCholPer * (kCholPer - 1)) / 2;
                    ll += kCholPer;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) s = fma(a[u], c[u], s);
                off = o;
            }
#pragma unroll
            for (int o = 1; o < kCholPer; o <<= 1) s += __shfl_xor_sync(gmask, s, o);
            s = Lx(row, j) - s;
            if (row == j && part == 0) {
                if (!(s > 0.0)) { bad = 1; s = 1.0; }
                const double ri = rsqrt(s);
                rdiag[j] = ri;
                Lx(j, j) = s * ri;
            }
        }
        __syncthreads();
        if (row > j && part == 0) Lx(row, j) = s * rdiag[j];
        __syncthreads();
    }
    // inverse: thread group `row` owns COLUMN j = row of Linv; no dependency between columns
    {
        const int j = row;
        if (part == 0) Xx(j, j) = rdiag[j];
        __syncwarp(gmask);
        for (int i = j + 1; i < m; ++i) {
            double s = 0.0;
            const int l0 = j + part;
            int lo = l0 * m - (l0 * (l0 - 1)) / 2 - l0;     // column l of L, minus l
            int xo = (l0 * (l0 + 1)) / 2 + j;                 // row l of Linv, column j
            for (int l = l0; l < i; l += 4 * kCholPer) {
                double a[4], c[4];
                int ll = l;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const bool ok = ll < i;
                    a[u] = ok ? L[lo + i] : 0.0;
                    c[u] = ok ? X[xo] : 0.0;
                    lo += kCholPer * (m - 1 - ll) - (kCholPer * (kCholPer - 1)) / 2;
                    xo += kCholPer * ll + (kCholPer * (kCholPer + 1)) / 2;
                    ll += kCholPer;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) s = fma(a[u], c[u], s);
            }
#pragma unroll
            for (int o = 1; o < kCholPer; o <<= 1) s += __shfl_xor_sync(gmask, s, o);
            if (part == 0) Xx(i, j) = -s * rdiag[i];
            __syncwarp(gmask);
        }
    }
    __syncthreads();
    for (int e = tid; e < m * m; e += blockDim.x) {
        const int i = e % m, j = e / m;
        Linv[i + (size_t)j * m] = (i >= j) ? Xx(i, j) : 0.0;
    }
    if (tid < m) diagL[tid] = Lx(tid, tid);
    if (tid == 0 && bad) fail[0] = 1;
}

// res[j] = || Y_j - theta_j X_j ||_2, one CTA per column; fixed reduction order.
__global__ void __launch_bounds__(256)
topk_residual_kernel(const double* __restrict__ X, const double* __restrict__ Y, int n, const double* __restrict__ theta,
                     double* __restrict__ res) {
    __shared__ double red[8];
    const int j = blockIdx.x;
    const double th = theta[j];
    const double* x = X + (size_t)j * n;
    const double* y = Y + (size_t)j * n;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const double d = fma(-th, x[i], y[i]);
        acc = fma(d, d, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w];
        res[j] = sqrt(s);
    }
}

// count of the leading eigenvalues >= eps among the first k (filter.cpp:213-216 prefix rule) -> d_count; S <- theta[0..k)
__global__ void topk_finish_kernel(const double* __restrict__ theta, int k, double eps, double* __restrict__ S, int* __restrict__ d_count) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int r = 0;
        while (r < k && theta[r] >= eps) ++r;
        *d_count = r;
    }
    for (int j = threadIdx.x; j < k; j += blockDim.x) S[j] = theta[j];
}

struct Blocks {
    double *X, *Y, *T0, *T1, *T2, *G, *Linv, *W, *theta, *res, *diagL, *part;
    int* fail;
};

int choose_degree(double theta0, double thetak, double lb, double cut) {
    const double c = 0.5 * (cut + lb), e = 0.5 * (cut - lb);
    const double x0 = (theta0 - c) / e, xk = std::max((thetak - c) / e, 1.0 + 1e-9);
    int best = 2;
    for (int d = 2; d <= kMaxDegree; ++d) {
        const double ratio = std::cosh(d * std::acosh(x0)) / std::cosh(d * std::acosh(xk));
        if (ratio <= kMaxRatio) best = d;
    }
    return best;
}

}  // namespace

bool sym_eig_topk_supported(int n, int k) {
    if (k < 1) return false;
    // Measured on B200 (scripts/gpu_topk.py, profiles/r2p_topk.md): the fixed costs -- two or three m x m Rayleigh-Ritz solves
    // and a Cholesky-QR per filter round, all single-CTA latency chains -- equal the full solver's at n ~ 10 m
    // (m = 64: n ~ 600, m = 128: n ~ 1100); above that the block method wins and the margin grows with n.
    const int m = topk_block_width(k);
    return m <= kMaxM && n >= 10 * m && n >= 256;
}

// Policy of the training pipeline (api.cu): is the block method expected to be FASTER than the full solver for eig(Q)?
// Re-measured after the tridiagonalisation of the full solver got 20-25 % faster (scripts/gpu_topk_switch.py, one B200, the stage
// "Wa, WW, eig(Wa), Q, eig(Q)" with the block solver / with the full solver): block width 64 (k = 50): n = 855: 13.3 / 14.2 ms,
// n = 1357: 19.1 / 20.1 ms -- the block method keeps its lead; block width 128 (k = 100): n = 1833: 47.8 / 45.5 ms -- the two
// or three 128 x 128 Rayleigh-Ritz solves and the Cholesky-QR chains of 2 x 128 dependent pivots cost more than they save.
bool sym_eig_topk_preferred(int n, int k) {
    if (!sym_eig_topk_supported(n, k)) return false;
    const int m = topk_block_width(k);
    return m <= 64 || n >= 24 * m;
}

int topk_block_width(int k) {
    const int m = k + std::max(14, k / 4);
    return (m + 7) / 8 * 8;
}

// A: n x n symmetric positive semi-definite, FULL storage (both triangles), ld n.  Z (n x k, ld n) receives orthonormal
// eigenvectors of the k largest eigenvalues (descending), S (k) the eigenvalues, d_count the length of the prefix with
// S >= eps.  Returns false if the solver gave up (the caller then runs the full solver); `gemms` counts the A * X products.
bool sym_eig_topk(const double* A, int n, int k, double eps, double* Z, double* S, int* d_count, EigWorkspace& ws,
                  cudaStream_t s, int* gemms) {
    const int m = topk_block_width(k);
    const size_t nm = (size_t)n * m, mm = (size_t)m * m;
    const int nzA = dgemm_splitk_slices(n, m, n), nzG = dgemm_splitk_slices(m, m, n);
    const size_t part_doubles = std::max((size_t)nzA * nm, (size_t)nzG * mm);
    TmpBuf<double> buf(5 * nm + 3 * mm + 3 * (size_t)m + part_doubles + 64);
    TmpBuf<int> d_fail(4);
    Blocks b;
    b.X = buf.p; b.Y = b.X + nm; b.T0 = b.Y + nm; b.T1 = b.T0 + nm; b.T2 = b.T1 + nm;
    b.G = b.T2 + nm; b.Linv = b.G + mm; b.W = b.Linv + mm; b.theta = b.W + mm; b.res = b.theta + m; b.diagL = b.res + m; b.part = b.diagL + m + 32;
    b.fail = d_fail.p;
    NLE_CUDA(cudaMemsetAsync(b.fail, 0, 4 * sizeof(int), s));
    const size_t csm = (size_t)m * (m + 1) * sizeof(double);      // two packed triangles
    allow_max_dynamic_smem((const void*)topk_chol_inv_kernel);
    int ngemm = 0;

    auto ax = [&](const double* Xin, double* out, double alpha, double beta, const double* P, double gamma, const double* Zt) {
        dgemm_splitk(false, n, m, n, A, n, Xin, n, b.part, s);
        dgemm_combine(n, m, n, b.part, alpha, beta, P, n, gamma, Zt, n, out, n, s);
        ++ngemm;
    };
    // X <- Xin L^-T with Xin^T Xin = L L^T (Xin is overwritten as scratch)
    auto cholqr = [&](double* Xin, double* out) {
        dgemm_splitk(true, m, m, n, Xin, n, Xin, n, b.part, s);
        dgemm_combine(m, m, n, b.part, 1.0, 0.0, nullptr, 0, 0.0, nullptr, 0, b.G, m, s);
        topk_chol_inv_kernel<<<1, kCholPer * m, csm, s>>>(b.G, m, b.Linv, b.diagL, b.fail);
        NLE_LAUNCH_CHECK();
        dgemm(false, true, n, m, m, 1.0, Xin, n, b.Linv, m, 0.0, out, n, s);
    };

    topk_init_kernel<<<cdiv((long long)nm, 256), 256, 0, s>>>(b.T0, (long long)nm);
    NLE_LAUNCH_CHECK();
    cholqr(b.T0, b.T1);
    cholqr(b.T1, b.X);
    for (int it = 0; it < 2; ++it) {
        ax(b.X, b.T0, 1.0, 0.0, nullptr, 0.0, nullptr);
        cholqr(b.T0, b.X);
    }
    std::vector<double> h(2 * (size_t)m);
    int h_fail = 0;
    bool converged = false;
    for (int blk = 0; blk < kMaxBlocks; ++blk) {
        // Rayleigh-Ritz
        ax(b.X, b.Y, 1.0, 0.0, nullptr, 0.0, nullptr);
        dgemm_splitk(true, m, m, n, b.X, n, b.Y, n, b.part, s);
        dgemm_combine(m, m, n, b.part, 1.0, 0.0, nullptr, 0, 0.0, nullptr, 0, b.G, m, s);
        sym_eig(b.G, m, m, -1e300, false, b.W, b.theta, d_count, ws, s);       // all m Ritz pairs, descending
        dgemm(false, false, n, m, m, 1.0, b.X, n, b.W, m, 0.0, b.T0, n, s);
        dgemm(false, false, n, m, m, 1.0, b.Y, n, b.W, m, 0.0, b.T1, n, s);
        std::swap(b.X, b.T0);
        std::swap(b.Y, b.T1);
        topk_residual_kernel<<<m, 256, 0, s>>>(b.X, b.Y, n, b.theta, b.res);
        NLE_LAUNCH_CHECK();
        NLE_CUDA(cudaMemcpyAsync(h.data(), b.theta, 2 * (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, s));   // theta, res adjacent
        NLE_CUDA(cudaMemcpyAsync(&h_fail, b.fail, sizeof(int), cudaMemcpyDeviceToHost, s));
        NLE_CUDA(cudaStreamSynchronize(s));
        if (h_fail) break;
        const double* theta = h.data();
        const double* res = h.data() + m;
        bool finite = true;
        double worst = 0.0;
        for (int j = 0; j < m; ++j) finite = finite && std::isfinite(theta[j]) && std::isfinite(res[j]);
        for (int j = 0; j < k; ++j) worst = std::max(worst, res[j]);
        if (!finite || !(theta[0] > 0.0)) break;
        if (worst <= kResTol * theta[0]) { converged = true; break; }
        if (blk + 1 == kMaxBlocks) break;
        const double lb = -1e-3 * theta[0], cut = theta[m - 1];
        if (!(cut > lb) || !(theta[k - 1] > cut)) break;                          // no gap between the wanted part and the block's tail
        const int d = choose_degree(theta[0], theta[k - 1], lb, cut);
        const double c = 0.5 * (cut + lb), e = 0.5 * (cut - lb);
        for (int rd = 0; rd < kRounds; ++rd) {
            // T_d((A - c)/e) X by the three-term recurrence:  t1 = (A x - c x)/e,  t_{i+1} = 2 (A t_i - c t_i)/e - t_{i-1}
            double* t0 = b.X;
            double* t1 = b.T0;
            double* t2 = b.T1;
            ax(t0, t1, 1.0 / e, -c / e, t0, 0.0, nullptr);
            for (int i = 2; i <= d; ++i) {
                ax(t1, t2, 2.0 / e, -2.0 * c / e, t1, -1.0, t0);
                double* old = t0;
                t0 = t1; t1 = t2; t2 = (old == b.X) ? b.T2 : old;               // never recycle X itself before the round ends
            }
            cholqr(t1, b.X);
        }
    }
    if (gemms) *gemms = ngemm;
    if (!converged) return false;
    // eigenvalue prefix >= eps and the k-th pair must be a genuine member of the kept set
    if (!(h[k - 1] >= eps)) return false;
    NLE_CUDA(cudaMemcpyAsync(Z, b.X, (size_t)n * k * sizeof(double), cudaMemcpyDeviceToDevice, s));
    topk_finish_kernel<<<1, 128, 0, s>>>(b.theta, k, eps, S, d_count);
    NLE_LAUNCH_CHECK();
    return true;
}

}  // namespace nle
