// Internal helpers shared by the translation units of libnle_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace nle {

constexpr double kEps = 1e-10;  // nle::EPS, include/filter.hpp:14 of the reference

void set_error(const std::string& msg);
extern thread_local long long g_launches;  // kernels launched by this library on this thread
extern thread_local int g_eig_fallbacks;   // direct eigensolver -> block-Jacobi fallbacks on this thread (eig.cu)

struct CudaError {
    cudaError_t e;
    const char* what;
    const char* file;
    int line;
};

#define NLE_CUDA(call)                                                         \
    do {                                                                       \
        cudaError_t _e = (call);                                               \
        if (_e != cudaSuccess) throw ::nle::CudaError{_e, #call, __FILE__, __LINE__}; \
    } while (0)

#define NLE_LAUNCH_CHECK()                                                     \
    do {                                                                       \
        ++::nle::g_launches;                                                   \
        cudaError_t _e = cudaGetLastError();                                   \
        if (_e != cudaSuccess) throw ::nle::CudaError{_e, "kernel launch", __FILE__, __LINE__}; \
    } while (0)

struct InvalidArg { std::string msg; };
struct Unsupported { std::string msg; };
struct NoConvergence { std::string msg; };

// Device memory comes from the CUDA stream-ordered pool (cudaMallocAsync on the legacy stream, release
// threshold = unlimited, set once in dense.cu): after the first image every allocation of the
// training pipeline is a pool hit (microseconds) instead of a cudaMalloc/cudaFree pair.
void* pool_alloc(size_t bytes);
void pool_free(void* p);

// RAII device buffer (typed).
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        if (count) p = static_cast<T*>(pool_alloc(count * sizeof(T)));
    }
    void release() {
        if (p) pool_free(p);
        p = nullptr;
        n = 0;
    }
    void zero(cudaStream_t s) { if (n) NLE_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
    void upload(const T* h, size_t count, cudaStream_t s) {
        NLE_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void download(T* h, size_t count, cudaStream_t s) const {
        NLE_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    }
};

// Bump allocator for the temporaries of ONE training call (thread-local, reset at the start of the call).
// In steady state (same image shape) a training call performs no device allocation at all: the arena
// is a single chunk that is simply rewound.  If a call outgrows it, extra chunks are appended and the
// next reset() coalesces them into one chunk of the total size.
struct Arena {
    struct Chunk { char* p; size_t cap; };
    std::vector<Chunk> chunks;
    size_t cur = 0, off = 0, total_used = 0;
    void* alloc(size_t bytes);
    void reset();
    void release_all();
    ~Arena() { release_all(); }
};
Arena& thread_arena();

// Non-owning typed view of arena memory (same surface as DevBuf where train_core needs it).
template <typename T>
struct TmpBuf {
    T* p = nullptr;
    size_t n = 0;
    TmpBuf() = default;
    explicit TmpBuf(size_t count) { alloc(count); }
    void alloc(size_t count) {
        n = count;
        p = count ? static_cast<T*>(thread_arena().alloc(count * sizeof(T))) : nullptr;
    }
    void upload(const T* h, size_t count, cudaStream_t s) {
        NLE_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void download(T* h, size_t count, cudaStream_t s) const {
        NLE_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    }
};

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__
// D(8x8) += A(8x4, row) * B(4x8, col) on the FP64 tensor pipe (mma.sync.m8n8k4.f64, SASS: DMMA).
// lane = 4*g + t holds  a = A[g][t],  b = B[t][g],  d0/d1 = D[g][2t], D[g][2t+1].
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
#endif

int sm_count();
// Largest shared memory a CTA may opt into on the current device.
int smem_optin_max();
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) with the largest value the kernel can take (opt-in maximum minus its static
// shared memory), done once per (kernel, device).  Never the size of the launch at hand: the attribute is per function and
// device, and host threads that train different slabs on one device would otherwise lower it under each other's launches.
void allow_max_dynamic_smem(const void* kernel);

// ---- dense.cu : column-major FP64 building blocks ------------------------------------------
// C(m x n) = alpha * op(A) * op(B) + beta * C
void dgemm(bool transA, bool transB, int m, int n, int k, double alpha, const double* A, int lda,
           const double* B, int ldb, double beta, double* C, int ldc, cudaStream_t s);
// Split-K form for skinny outputs (few 64 x 64 tiles, long k): partial products per k slice, then
// C = alpha * sum_z part_z + beta * P + gamma * Z  (slices summed in index order).  part: slices * m * n doubles.
int dgemm_splitk_slices(int m, int n, int k);
void dgemm_splitk(bool transA, int m, int n, int k, const double* A, int lda, const double* B, int ldb, double* part,
                  cudaStream_t s);
void dgemm_combine(int m, int n, int k, const double* part, double alpha, double beta, const double* P, int ldp, double gamma,
                   const double* Z, int ldz, double* C, int ldc, cudaStream_t s);
// y(m) = A(m x n) x        /  y(n) = A(m x n)^T x
void dgemv_n(int m, int n, const double* A, int lda, const double* x, double* y, cudaStream_t s);
void dgemv_t(int m, int n, const double* A, int lda, const double* x, double* y, cudaStream_t s);
// Fused Sinkhorn sample-side steps (dense.cu): w = U t and xs = recip(U (lam o t)) in one pass; t = U^T x + inv_lam o (U^T sv).
void sk_sample_step(int m, int n, const double* U, int ldu, const double* t, const double* lam, double eps, double* w,
                    double* xs, cudaStream_t s);
void sk_phiT(int m, int n, const double* U, int ldu, const double* x, const double* sv, const double* inv_lam, double* t,
             cudaStream_t s);
// out(i,j) = rowscale[i] * A(i,j) * colscale[j]   (either scale may be null)
void scale_rows_cols(int m, int n, const double* A, int lda, const double* rowscale,
                     const double* colscale, double* out, int ldo, cudaStream_t s);
// v_i <- 1/v_i if |v_i| >= eps else 0        (inplaceReciprocal, filter.cpp:42-54)
void guarded_reciprocal(double* v, int n, double eps, cudaStream_t s);
// v_i <- 1/sqrt(v_i) if |v_i| >= eps else 0  (filter.cpp:289-291, 319-321)
void guarded_inv_sqrt(const double* v, double* out, int n, double eps, cudaStream_t s);

// ---- eig_dc.cu / eig.cu : symmetric eigensolver (FP64, no LAPACK) ----------------------------
// Default: Householder tridiagonalisation + divide & conquer + back-transformation (eig_dc.cu).
// Fallback (NLE_B200_EIG=jacobi, or if the direct solver's device-side sanity check fails): block
// one-sided Jacobi (eig.cu).
struct EigWorkspace {
    DevBuf<double> W, As, T, lam_unsorted;
    DevBuf<double> Qa, Qb, Sb, dcd;                     // divide & conquer buffers (eig_dc.cu)
    DevBuf<double> wyT;                                 // T factors of the blocked back-transformation
    DevBuf<uint4> trdll;                                // flagged exchange cells of tridiag_resident_kernel
    DevBuf<int> dci;
    int dc_cap = 0;
    void reserve_dc(int n);
    DevBuf<int> ctrl, order;
    DevBuf<long long> prof;
    long long prof_host[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // block-0 cycle counts: load, gram, inner, update, store, gridsync, steps, pairs
    int max_inner = 2;                                  // inner Jacobi sweeps per block-pair visit
    int cap = 0;
    void reserve(int n);
    // Phase marks of the direct solver (eig_dc.cu): between phase_reset() and phase_ms() every solve records four
    // events (start | tridiagonalised | divide & conquer done | back-transformed) WITHOUT synchronising; phase_ms adds
    // the three spans up over all solves since the reset (call it after the stream has been synchronised).
    std::vector<cudaEvent_t> pev;
    int pev_used = 0;
    bool phase_on = false;
    void phase_reset() { pev_used = 0; phase_on = true; }
    void phase_mark(cudaStream_t s);
    void phase_ms(double out[3]);
    // Side stream of the direct solver: the T factors of the blocked back-transformation depend on the reflectors only, so
    // bt_tfactor_kernel runs beside the divide & conquer phase (a chain of small kernels that leaves most SMs idle).
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int side_dev = -1;
    void side_init();
    ~EigWorkspace();
};
// Eigen-decomposition of the symmetric matrix defined by the LOWER triangle of M (n x n, ld ldm).
// U (n x n, ld n) receives eigenvectors sorted by descending eigenvalue, D (n) the eigenvalues;
// d_r (device int) receives the length of the prefix with D >= eps.  `psd_hint` selects a small
// spectral shift (inputs known to be positive semi-definite up to rounding).
// Returns the number of Jacobi sweeps used (0 when the direct solver was used).
bool sym_eig_dc_core(double* As, int n, double eps, int vec_limit, int* order, int* count, EigWorkspace& ws,
                     cudaStream_t s, double** lam_out, double** vec_out);
// vec_limit >= 0: the caller only uses the eigenvectors of the first min(vec_limit, *d_r) eigenvalues (the
// remaining columns of U are unspecified); < 0: all n columns are valid.
int sym_eig(const double* M, int ldm, int n, double eps, bool psd_hint, double* U, double* D,
            int* d_r, EigWorkspace& ws, cudaStream_t s, int vec_limit = -1);

// As(i,j) = As(j,i) = M(i,j), i >= j: full storage of the symmetric matrix the LOWER triangle of M defines (eig.cu).
void symmetrize_lower(const double* M, int ldm, int n, double* As, cudaStream_t s);

// ---- eig_topk.cu : top-k eigenpairs of a symmetric POSITIVE SEMI-DEFINITE matrix (third eigensolve, filter.cpp:311-316) ----
// Chebyshev-filtered block subspace iteration, Cholesky-QR, Rayleigh-Ritz; all products on the FP64 tensor pipe.
int topk_block_width(int k);                 // k + guard columns, multiple of 8
bool sym_eig_topk_preferred(int n, int k);   // ... and is expected to beat the full solver (the training pipeline's rule)
bool sym_eig_topk_supported(int n, int k);   // block fits one CTA's shared memory and n is large enough for the method to pay
// A: n x n, FULL storage (ld n).  Z: n x k (ld n) eigenvectors of the k largest eigenvalues, descending; S: k eigenvalues;
// d_count: length of the prefix with S >= eps.  false = gave up (Cholesky breakdown, no convergence, k-th eigenvalue < eps):
// the caller runs sym_eig instead.  gemms (host, optional): number of A * X block products used.
bool sym_eig_topk(const double* A, int n, int k, double eps, double* Z, double* S, int* d_count, EigWorkspace& ws,
                  cudaStream_t s, int* gemms);

}  // namespace nle
