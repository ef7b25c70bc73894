/*
 * nle_b200.h -- C ABI of the B200-native Nystrom spectral-filter library (libnle_b200.so).
 *
 * This is the drop-in boundary for the hot path of lightalchemist/nonlocal-image-edit: every entry
 * point below replaces one function of the reference's src/filter.cpp (cited as filter.cpp:LINE,
 * header include/filter.hpp as filter.hpp:LINE).  The reference has no FFI layer; its boundary is
 * the C++ header include/filter.hpp, so the binding a maintainer adds is a replacement translation
 * unit for src/filter.cpp that implements the unchanged filter.hpp on top of these calls
 * (see INTEGRATION.md).  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *   - All matrices are COLUMN-MAJOR doubles (Eigen's default, filter.hpp:10-11) unless stated.
 *   - Images/channels are row-major (raster) like cv::Mat (utils.hpp:11-14).
 *   - Pointers are HOST pointers unless the function name ends in _dev.
 *   - Every function returns 0 on success, a negative nle_b200_status otherwise, and never throws;
 *     nle_b200_last_error() returns the message for the calling thread.  The reference's
 *     std::runtime_error messages (filter.cpp:118,352,356,415,419,448) are reproduced verbatim there.
 *   - There is no CPU fallback: without a CUDA device every compute call fails with
 *     NLE_B200_ERR_CUDA.
 *   - A filter handle is not thread-safe; distinct handles are independent.
 */
#ifndef NLE_B200_H
#define NLE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    NLE_B200_OK = 0,
    NLE_B200_ERR_INVALID = -1,     /* bad argument (message mirrors the reference's runtime_error) */
    NLE_B200_ERR_CUDA = -2,        /* CUDA runtime failure / no device */
    NLE_B200_ERR_UNSUPPORTED = -3, /* outside the supported envelope (e.g. non-integer luminance) */
    NLE_B200_ERR_NOCONV = -4       /* eigensolver did not converge */
} nle_b200_status;

/* Opaque trained filter: the state of nle::NLEFilter (m_eigvecs N x k, m_eigvals k;
 * filter.hpp:52-53) kept resident in HBM. */
typedef struct nle_b200_filter nle_b200_filter;

typedef struct {
    int rows, cols;      /* full image size                                              */
    int row0, row1;      /* image rows [row0,row1) owned by this handle (sharding)       */
    int p;               /* number of Nystrom samples (selected.size(), filter.cpp:124)  */
    int r;               /* #eigenvalues of Ka >= 1e-10 (filter.cpp:213-216,262)         */
    int r2;              /* #eigenvalues of Wa >= 1e-10 (filter.cpp:287)                 */
    int k;               /* eigenvectors kept: min(nEigenVectors, #eig(Q)>=1e-10)        */
    int n_row_samples_eff, n_col_samples_eff; /* grid actually produced by samplePixels  */
    int eig_sweeps[3];   /* Jacobi sweeps used by the three eigensolves (0 = direct solver) */
    int eig_fallbacks;   /* eigensolves of this training call that fell back from the direct solver to block
                            Jacobi (also reported on stderr when it happens); 0 in normal operation */
    int topk_products;   /* block products A*X spent by the top-k solver on eig(Q) (filter.cpp:311-316); 0 = the block of Q
                            was too small for it and the full solver ran; < 0 = it gave up after that many and the full
                            solver ran */
} nle_b200_info;

/* Small all-reduce (sum) used by row-sharded training/apply.  `dev_buf` is a DEVICE pointer to
 * `count` doubles on the calling rank's GPU; the callee must sum it in place across ranks on
 * `cuda_stream` (a cudaStream_t) or synchronously.  The host language supplies it (the Python host
 * wraps torch.distributed/NCCL).  Only p-vectors, one p x p Gram and k-vectors ever go through it. */
typedef int (*nle_b200_allreduce_fn)(void* dev_buf, size_t count, void* cuda_stream, void* user);

/* NCCL communicator owned by the library (csrc/nccl_comm.cu; libnccl is resolved with dlopen at the first call, the
 * library itself links no NCCL).  Rank 0 calls nle_b200_comm_unique_id and the host ferries the 128 bytes to the other
 * ranks (any transport); then EVERY rank calls nle_b200_comm_create on its own device (collective).  Pass
 * nle_b200_comm_allreduce as `allreduce` and the communicator as `user` to the sharded training entry points: the
 * sums are then enqueued by the library on its own stream.  The p x p Gram is an ncclAllReduce; the latency-sized
 * messages (p-vectors, k-vectors: up to 8192 doubles) are ONE launch of the library's own peer_allreduce_kernel, which
 * stores flagged cells straight into every rank's inbox over NVLink peer memory (CUDA IPC, set up in
 * nle_b200_comm_create) and sums them in rank order -- bit-identical on every rank; where peers cannot map each other's
 * memory they fall back to ncclAllReduce.  nle_b200_comm_info reports which path is in use (*peer_path = 1: peer
 * memory), the reductions issued so far on the peer / NCCL path, and returns "ok" or the reason. */
typedef struct nle_b200_comm nle_b200_comm;
int nle_b200_comm_unique_id(unsigned char id[128]);
int nle_b200_comm_create(const unsigned char id[128], int rank, int nranks, nle_b200_comm** out);
int nle_b200_comm_allreduce(void* dev_buf, size_t count, void* cuda_stream, void* user /* nle_b200_comm* */);
const char* nle_b200_comm_info(nle_b200_comm* comm, int* peer_path, unsigned long long calls[2]);
void nle_b200_comm_destroy(nle_b200_comm* comm);

const char* nle_b200_last_error(void);
int nle_b200_version(void);
int nle_b200_device_count(void);

/* ---- (1) sample selection: samplePixels, filter.cpp:56-80; to1DIndex, utils.hpp:11-14 -------- */
/* Number of selected pixels p (can exceed nRowSamples*nColSamples). Pure host arithmetic. */
int nle_b200_sample_count(int rows, int cols, int nRowSamples, int nColSamples, int* p_out);
/* Raster indices of `selected` (p ints) and, if rest != NULL, of `rest` (rows*cols-p ints), both in
 * raster order, generated on the GPU.  Bit-exact with the reference (permutation of
 * filter.cpp:156-164 is [selected; rest]). */
int nle_b200_sample_indices(int rows, int cols, int nRowSamples, int nColSamples,
                            int32_t* selected, int32_t* rest);

/* ---- free functions of filter.hpp:20-33 (dense, for the reference's unit tests) ------------- */
/* computeKernel, filter.cpp:114-167.  channel: rows x cols doubles (integer-valued 0..255, as both
 * reference callers produce).  Ka: p x p.  Kab: p x (N-p) or NULL.  perm: N ints or NULL. */
int nle_b200_compute_kernel(const double* channel, int rows, int cols, int nRowSamples,
                            int nColSamples, double hx, double hy,
                            int32_t* perm, double* Ka, double* Kab);
/* eigenDecomposition, filter.cpp:204-228.  Reads the LOWER triangle of M (n x n) like Eigen's
 * SelfAdjointEigenSolver; eigenvalues descending; *r_out = length of the prefix with D >= eps.
 * U: n x n (all eigenvectors, first r are the reference's result), D: n. */
int nle_b200_eigen_decomposition(const double* M, int n, double eps, double* U, double* D,
                                 int* r_out);
/* topkEigenDecomposition, filter.cpp:169-200 (the reference's USE_SPECTRA build; its only caller passes Q, :311).
 * nev = min(nLargest, n-1) (:172) eigenpairs of largest magnitude in descending algebraic order; *r_out = length of the
 * prefix with D >= eps (:189-198).  U: n x nev, D: nev.  assume_psd != 0 promises a positive semi-definite M (true for Q)
 * and selects the Chebyshev-filtered block iteration that the training path uses for eig(Q) when n is large against nev;
 * otherwise, or when that solver gives up, the full eigensolver runs and the nev pairs are selected from it.
 * *products_out (optional): block products used by the fast solver (0: not used, < 0: gave up after that many). */
int nle_b200_topk_eigen_decomposition(const double* M, int n, int nLargest, double eps, int assume_psd,
                                      double* U, double* D, int* r_out, int* products_out);
/* nystromApproximation, filter.cpp:257-280.  eigvals: p, phi: (p+nrest) x p (first r columns
 * valid), *r_out = rank kept. */
int nle_b200_nystrom_approximation(const double* Ka, int p, const double* Kab, int nrest,
                                   double* eigvals, double* phi, int* r_out);
/* sinkhorn, filter.cpp:230-254.  phi: n x r, eigvals: r.  Wa: r x r, Wab: r x (n-r). */
int nle_b200_sinkhorn(const double* phi, int n, int r, const double* eigvals, int maxIter,
                      double* Wa, double* Wab);
/* orthogonalize, filter.cpp:282-331 (non-Spectra branch).  Wa: p x p, Wab: p x nrest.
 * V: (p+nrest) x nEigVectors (first *k_out columns valid), S: nEigVectors. */
int nle_b200_orthogonalize(const double* Wa, int p, const double* Wab, int nrest, int nEigVectors,
                           double eps, double* V, double* S, int* k_out);
/* transformEigenValues, filter.cpp:334-347 (host arithmetic, std::pow). */
int nle_b200_transform_eigenvalues(const double* eigvals, int k, const double* weights, int m,
                                   double* fS);

/* ---- NLEFilter::trainFilter, filter.cpp:480-502 ---------------------------------------------- */
/* channel: rows x cols luminance as produced by getLuminanceChannel (filter.cpp:460-469), i.e.
 * integer-valued doubles in [0,255].  Non-integer input -> NLE_B200_ERR_UNSUPPORTED. */
int nle_b200_train(const double* channel, int rows, int cols, int nRowSamples, int nColSamples,
                   double hx, double hy, int nSinkhornIter, int nEigenVectors,
                   nle_b200_filter** out);
/* Same, 8-bit luminance (what cvtColor(BGR2Lab) produced before convertTo(CV_64F)). */
int nle_b200_train_u8(const uint8_t* lum, int rows, int cols, int nRowSamples, int nColSamples,
                      double hx, double hy, int nSinkhornIter, int nEigenVectors,
                      nle_b200_filter** out);
/* Row-sharded training: this rank owns image rows [row0,row1) of the FULL host image `lum`
 * (every rank passes the same image; only the slab and the p sample values are uploaded).
 * allreduce may be NULL when the slab is the whole image. */
int nle_b200_train_u8_sharded(const uint8_t* lum, int rows, int cols, int row0, int row1,
                              int nRowSamples, int nColSamples, double hx, double hy,
                              int nSinkhornIter, int nEigenVectors,
                              nle_b200_allreduce_fn allreduce, void* user,
                              nle_b200_filter** out);
/* Device-resident variant: lum_slab_dev points to rows [row0,row1) (row-major u8) already in HBM;
 * sample_lum holds the p luminances of the selected pixels in raster order (HOST, may be NULL when
 * the slab is the whole image). */
int nle_b200_train_u8_dev(const uint8_t* lum_slab_dev, int rows, int cols, int row0, int row1,
                          const uint8_t* sample_lum, int nRowSamples, int nColSamples, double hx,
                          double hy, int nSinkhornIter, int nEigenVectors,
                          nle_b200_allreduce_fn allreduce, void* user, nle_b200_filter** out);

int nle_b200_filter_info(const nle_b200_filter* f, nle_b200_info* info);
/* m_eigvals (k doubles). */
int nle_b200_eigenvalues(const nle_b200_filter* f, double* S);
/* m_eigvecs after the un-permute of filter.cpp:502: (row1-row0)*cols x k, column-major, pixel
 * (raster) order of the owned slab. */
int nle_b200_eigenvectors(const nle_b200_filter* f, double* V);

/* ---- NLEFilter::apply, filter.cpp:445-458:  out = V diag(fS) V^T channel ---------------------- */
/* channel/out: the owned slab, (row1-row0)*cols doubles; n_values = channel.total(): a mismatch fails with the
 * reference's message "Number of values in channel must match that of training image." (filter.cpp:447-449). */
int nle_b200_apply(const nle_b200_filter* f, const double* channel, long long n_values, const double* fS,
                   double* out);
/* enhance on the L channel (filter.cpp:426-436): u8 -> transformEigenValues -> apply ->
 * max(.,0) -> min(.,255) -> convertTo(CV_8U) (round half to even), fused on the device. */
int nle_b200_enhance_luminance_u8(const nle_b200_filter* f, const uint8_t* lum,
                                  const double* weights, int m, uint8_t* out);
int nle_b200_enhance_luminance_u8_dev(const nle_b200_filter* f, const uint8_t* lum_slab_dev,
                                      const double* weights, int m, uint8_t* out_slab_dev);
/* denoise's per-channel step (filter.cpp:378-399): teig = pow(min(S,1),k); apply; clamp; round.  rows x cols = the
 * channel passed (owned slab); a size mismatch fails with the message of filter.cpp:355-357. */
int nle_b200_denoise_channel_u8(const nle_b200_filter* f, const uint8_t* chan, int rows, int cols, double k,
                                uint8_t* out);

/* ---- image-level entry points with the colour conversion on the device ------------------------
 * cv::cvtColor(COLOR_BGR2Lab / COLOR_Lab2BGR) on CV_8UC3 as used by the reference (filter.cpp:423,440,463,528),
 * byte-exact with OpenCV's fixed-point 8-bit path (csrc/lab.cu).  bgr/lab: interleaved, 3 bytes per pixel. */
int nle_b200_bgr_to_lab_u8(const uint8_t* bgr, long long npix, uint8_t* lab);
int nle_b200_lab_to_bgr_u8(const uint8_t* lab, long long npix, uint8_t* bgr);
/* NLEFilter::trainForEnhancement, filter.cpp:514-519: getLuminanceChannel (BGR2Lab, L) + trainFilter.  `bgr` is the
 * FULL rows x cols x 3 host image; this rank owns rows [row0,row1) (row0=0,row1=rows and allreduce=NULL: one GPU). */
int nle_b200_train_bgr_u8(const uint8_t* bgr, int rows, int cols, int row0, int row1, int nRowSamples,
                          int nColSamples, double hx, double hy, int nSinkhornIter, int nEigenVectors,
                          nle_b200_allreduce_fn allreduce, void* user, nle_b200_filter** out);
/* NLEFilter::enhance, filter.cpp:412-443, on the owned slab: BGR2Lab, enhance L (transformEigenValues, apply, clamp,
 * round), merge with the untouched a,b, Lab2BGR.  bgr_slab/out_slab: rows x cols x 3 host bytes, rows x cols being the
 * image (slab) passed: channels != 3 fails with "Can only enhance RGB image." (filter.cpp:414-416) and
 * rows*cols != (row1-row0)*cols of the handle with "Cannot apply filter on image with different size from the image
 * filter was trained on." (filter.cpp:418-420). */
int nle_b200_enhance_bgr_u8(const nle_b200_filter* f, const uint8_t* bgr_slab, int rows, int cols, int channels,
                            const double* weights, int m, uint8_t* out_slab);

/* ---- stage intermediates for parity tests (SURVEY.md 8b "test hooks") ------------------------ */
typedef enum {
    NLE_B200_STAGE_KA = 0,        /* p x p                                   */
    NLE_B200_STAGE_LAMBDA = 1,    /* r      eigenvalues of Ka kept           */
    NLE_B200_STAGE_RVEC_HEAD = 2, /* r      final Sinkhorn r on perm[0:r]    */
    NLE_B200_STAGE_C = 3,         /* (row1-row0)*cols, raster order of slab  */
    NLE_B200_STAGE_WA = 4,        /* r x r                                   */
    NLE_B200_STAGE_Q = 5,         /* r x r                                   */
    NLE_B200_STAGE_LA = 6,        /* r2     eigenvalues of Wa kept           */
    NLE_B200_STAGE_GRAM = 7,      /* p x p  sum_j c_j^2 k_j k_j^T (rest)     */
    NLE_B200_STAGE_TIMES_MS = 8   /* 16 doubles, device milliseconds of the training call (CUDA events recorded without
                                     host synchronisation): 0 tables+Ka | 1 eig(Ka) | 2 Sinkhorn | 3 Gram incl. all-reduce |
                                     4 small algebra + eig(Wa) + eig(Q) | 5 extension | 6 total | 7 Gram kernels only |
                                     8 tridiagonalisations | 9 divide & conquer | 10 back-transformations (sums over the
                                     three eigensolves) | 11..15 reserved */
} nle_b200_stage;
/* Copies min(cap, size) doubles; *size_out = full size.  Stages are kept only when the filter was
 * trained with nle_b200_set_keep_stages(1) (default 1; bench turns it off). */
int nle_b200_get_stage(const nle_b200_filter* f, int which, double* out, size_t cap,
                       size_t* size_out);
void nle_b200_set_keep_stages(int keep);
/* Number of kernel launches issued by this library on the calling thread since the last reset. */
long long nle_b200_launch_count(int reset);

/* FP64 FMA throughput of this GPU measured by a register-resident microbenchmark (TFLOP/s); the
 * roofline denominator bench.py uses for the DFMA-bound Gram kernel. */
double nle_b200_fp64_fma_peak_tflops(void);
/* Same for the FP64 tensor pipe: back-to-back mma.sync.m8n8k4.f64 (SASS DMMA) with register operands (TFLOP/s). */
double nle_b200_fp64_dmma_peak_tflops(void);
/* All measured ceilings bench.py quotes (csrc/peaks.cu), out[0..min(n,8)): FP64 FMA TFLOP/s, FP64 tensor pipe TFLOP/s,
 * FP32 FMA TFLOP/s, MUFU ex2 Gop/s, shared-memory load GB/s, L2 read GB/s (32 MB working set), HBM copy GB/s
 * (read + write bytes), shared-memory load bytes per SM clock.  Allocates 2 GiB of device memory for the duration of the call. */
int nle_b200_measured_peaks(double* out, int n);

void nle_b200_free(nle_b200_filter* f);
/* Releases what the calling thread keeps between calls (the parked eigenvector buffer of the last freed filter and the
 * arena of training temporaries).  Optional; a long-lived host calls it when it stops filtering. */
void nle_b200_release_cache(void);

#ifdef __cplusplus
}
#endif
#endif /* NLE_B200_H */
