// N-scaled kernels of the Nystrom spectral filter (sm_100a).
//
// The reference materialises Kab (p x (N-p) doubles, filter.cpp:126,139-145), phi (N x r, :274-275)
// and Wab (r x (N-r), :250).  Nothing of size O(p*N) exists here.  Every kernel below re-evaluates
// the affinity on the fly from three small FP64 tables, using two structural facts of the
// reference's own pipeline:
//   (a) the Nystrom samples form a product grid rows_sel x cols_sel (filter.cpp:68-70 is separable),
//   (b) the luminance is an 8-bit value (filter.cpp:463-466: 8-bit BGR2Lab, then convertTo CV_64F),
// so that for sample i=(a,b) and pixel j=(row,col,l)
//       K(i,j) = exp(-((row-Ra)^2+(col-Cb)^2)/hx^2 - (l-Yi)^2/hy^2) = Er[row][a]*Ec[col][b]*Gt[|l-Yi|].
// The Sinkhorn half-passes use this to contract over the photometric axis through a per-image-row
// table of 256 x nC entries (a "bilateral grid" in luminance), which removes the p*N exponentials
// AND most of the p*N multiply-adds; the Gram and extension kernels use it to generate FP64 operand
// tiles in shared memory with three table look-ups per element.
#include <cstdlib>

#include "kernels.cuh"

namespace nle {

// =============================================================================================
// (1) sample selection  -- filter.cpp:56-80, 156-164
__global__ void sample_indices_kernel(int rows, int cols, const int* __restrict__ rowa,
                                      const int* __restrict__ colb, const int* __restrict__ rowrank,
                                      const int* __restrict__ colrank, int nC,
                                      int32_t* __restrict__ selected, int32_t* __restrict__ rest) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)rows * cols) return;
    int row = (int)(idx / cols), col = (int)(idx - (long long)row * cols);
    int a = rowa[row], b = colb[col];
    // number of selected pixels strictly before (row,col) in raster order
    long long before = (long long)rowrank[row] * nC + (a >= 0 ? colrank[col] : 0);
    if (a >= 0 && b >= 0) selected[a * nC + b] = (int32_t)idx;       // to1DIndex, utils.hpp:11-14
    else if (rest) rest[idx - before] = (int32_t)idx;
}

void launch_sample_indices(int rows, int cols, const int* rowa, const int* colb,
                           const int* rowrank, const int* colrank, int nC,
                           int32_t* selected, int32_t* rest, cudaStream_t s) {
    long long n = (long long)rows * cols;
    sample_indices_kernel<<<cdiv(n, 256), 256, 0, s>>>(rows, cols, rowa, colb, rowrank, colrank, nC, selected, rest);
    NLE_LAUNCH_CHECK();
}

// =============================================================================================
// tables
__global__ void tables_kernel(int rows, int cols, int nR, int nC, const int* __restrict__ sel_rows,
                              const int* __restrict__ sel_cols, double sw, double pw,
                              double* __restrict__ Er, double* __restrict__ Ec,
                              double* __restrict__ EcT, double* __restrict__ Gt) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long nEr = (long long)rows * nR, nEc = (long long)cols * nC;
    if (idx < nEr) {
        int row = (int)(idx / nR), a = (int)(idx - (long long)row * nR);
        int d = row - sel_rows[a];
        Er[idx] = exp(-sw * (double)(d * d));
    } else if (idx < nEr + nEc) {
        long long e = idx - nEr;
        int col = (int)(e / nC), b = (int)(e - (long long)col * nC);
        int d = col - sel_cols[b];
        double v = exp(-sw * (double)(d * d));
        Ec[e] = v;
        EcT[(size_t)b * cols + col] = v;
    } else if (idx < nEr + nEc + 256) {
        int d = (int)(idx - nEr - nEc);
        Gt[d] = exp(-pw * (double)(d * d));
    }
}

void launch_tables(int rows, int cols, int nR, int nC, const int* sel_rows, const int* sel_cols,
                   double hx, double hy, double* Er, double* Ec, double* EcT, double* Gt,
                   cudaStream_t s) {
    double pw = 1.0 / (hy * hy), sw = 1.0 / (hx * hx);   // filter.cpp:128-129
    long long n = (long long)rows * nR + (long long)cols * nC + 256;
    tables_kernel<<<cdiv(n, 256), 256, 0, s>>>(rows, cols, nR, nC, sel_rows, sel_cols, sw, pw, Er, Ec, EcT, Gt);
    NLE_LAUNCH_CHECK();
}

// Ka with the reference's single exponential of the summed argument (filter.cpp:109-111,134-136,144)
__global__ void ka_kernel(int p, int nC, const int* __restrict__ sel_rows, const int* __restrict__ sel_cols,
                          const uint8_t* __restrict__ Ysel, double sw, double pw, double* __restrict__ Ka) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= p) return;
    int ai = i / nC, bi = i - ai * nC, aj = j / nC, bj = j - aj * nC;
    int dr = sel_rows[ai] - sel_rows[aj], dc = sel_cols[bi] - sel_cols[bj];
    double d2 = (double)(dr * dr + dc * dc);
    double dy = (double)Ysel[i] - (double)Ysel[j];
    Ka[i + (size_t)j * p] = exp(-sw * d2 - pw * (dy * dy));
}

void launch_ka(int p, int nC, const int* sel_rows, const int* sel_cols, const uint8_t* Ysel,
               double hx, double hy, double* Ka, cudaStream_t s) {
    double pw = 1.0 / (hy * hy), sw = 1.0 / (hx * hx);
    ka_kernel<<<dim3(cdiv(p, 128), p), 128, 0, s>>>(p, nC, sel_rows, sel_cols, Ysel, sw, pw, Ka);
    NLE_LAUNCH_CHECK();
}

// =============================================================================================
// Per-row level list: which of the 256 luminance levels occur in this image row.
// Requires blockDim.x == 256.  levidx[l] = compact index or -1; lev[li] = level; returns nlev.
__device__ __forceinline__ int build_levels(const uint8_t* __restrict__ Lrow, int W, int* flags,
                                            int* levidx, int* lev, int* wcount) {
    const int tid = threadIdx.x;
    flags[tid] = 0;
    __syncthreads();
    for (int c = tid; c < W; c += 256) flags[Lrow[c]] = 1;
    __syncthreads();
    unsigned m = __ballot_sync(0xffffffffu, flags[tid] != 0);
    if ((tid & 31) == 0) wcount[tid >> 5] = __popc(m);
    __syncthreads();
    int base = 0;
    for (int w = 0; w < (tid >> 5); ++w) base += wcount[w];
    int my = base + __popc(m & ((1u << (tid & 31)) - 1u));
    if (flags[tid]) { levidx[tid] = my; lev[my] = tid; } else levidx[tid] = -1;
    int total = 0;
    for (int w = 0; w < 8; ++w) total += wcount[w];
    __syncthreads();
    return total;
}

// ---------------------------------------------------------------------------------------------
// Sinkhorn "dot" half-pass, one CTA per image row (SURVEY App. A.4:  (K~x)_j = k_j^T w).
//   wp[a][b]   = Er[row][a] * w[a][b]
//   F[li][b]   = sum_a wp[a][b] * Gt[|lev[li] - Y[a][b]|]          (nlev*p multiply-adds per row)
//   y_col      = sum_b Ec[col][b] * F[li(col)][b]                   (W*nC per row)
//   x_col      = |y| >= eps ? 1/y : 0   (inplaceReciprocal, filter.cpp:42-54); 0 at sample pixels
__global__ void __launch_bounds__(256)
pass_dot_kernel(AffinityTables t, const double* __restrict__ w, double* __restrict__ x, int lev_cap) {
    extern __shared__ double smd[];
    const int p = t.p, nC = t.nC, nR = t.nR, W = t.cols;
    double* wp = smd;                       // p
    double* Gs = wp + p;                    // 256
    double* F = Gs + 256;                   // lev_cap * nC
    int* flags = reinterpret_cast<int*>(F + (size_t)lev_cap * nC);  // 256
    int* levidx = flags + 256;              // 256
    int* lev = levidx + 256;                // 256
    int* wcount = lev + 256;                // 8
    uint8_t* Ys = reinterpret_cast<uint8_t*>(wcount + 8);  // p
    uint8_t* Lrow = Ys + ((p + 15) / 16) * 16;             // W
    const int tid = threadIdx.x;
    for (int i = tid; i < p; i += 256) Ys[i] = t.Ysel[i];
    Gs[tid] = t.Gt[tid];
    for (int rl = blockIdx.x; rl < t.nrows; rl += gridDim.x) {
        const int row = t.row0 + rl;
        const uint8_t* Lg = t.lum + (size_t)rl * W;
        __syncthreads();
        for (int c = tid; c < W; c += 256) Lrow[c] = Lg[c];
        const double* er = t.Er + (size_t)row * nR;
        for (int i = tid; i < p; i += 256) wp[i] = er[i / nC] * w[i];
        __syncthreads();
        int nlev = build_levels(Lrow, W, flags, levidx, lev, wcount);
        for (int e = tid; e < nlev * nC; e += 256) {
            int li = e / nC, b = e - li * nC;
            int lvl = lev[li];
            double acc = 0.0;
            for (int a = 0; a < nR; ++a) {
                int i = a * nC + b;
                int d = lvl - (int)Ys[i];
                acc = fma(wp[i], Gs[d < 0 ? -d : d], acc);
            }
            F[e] = acc;
        }
        __syncthreads();
        const int a_row = t.rowa[row];
        double* xo = x + (size_t)rl * W;
        for (int c = tid; c < W; c += 256) {
            double r = 0.0;
            if (!(a_row >= 0 && t.colb[c] >= 0)) {
                const double* f = F + (size_t)levidx[Lrow[c]] * nC;
                double acc = 0.0;
                for (int b = 0; b < nC; ++b) acc = fma(t.EcT[(size_t)b * W + c], f[b], acc);
                r = (fabs(acc) >= kEps) ? 1.0 / acc : 0.0;
            }
            xo[c] = r;
        }
    }
}

static size_t pass_smem_bytes(const AffinityTables& t, bool reduce) {
    size_t b = 0;
    if (!reduce) b += (size_t)t.p * 8;                 // wp
    b += 256 * 8;                                      // Gs
    b += (size_t)256 * t.nC * 8;                       // F / Hh
    b += (256 * 3 + 8) * 4;                            // flags, levidx, lev, wcount
    b += ((t.p + 15) / 16) * 16;                       // Ys
    b += ((t.cols + 15) / 16) * 16;                    // Lrow
    if (reduce) b += (size_t)t.cols * 8 + ((t.cols + 15) / 16) * 16;   // xrow, lirow
    return b + 64;
}

void launch_pass_dot(const AffinityTables& t, const double* w, double* x, cudaStream_t s) {
    size_t smem = pass_smem_bytes(t, false);
    if (smem > 227 * 1024) throw Unsupported{"pass_dot: sample grid too wide for the per-row table (nC=" + std::to_string(t.nC) + ", p=" + std::to_string(t.p) + ")"};
    static size_t configured = 0;
    if (smem > configured) {
        NLE_CUDA(cudaFuncSetAttribute(pass_dot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    int grid = t.nrows;
    pass_dot_kernel<<<grid, 256, smem, s>>>(t, w, x, 256);
    NLE_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------
// Sinkhorn "reduce" half-pass, one CTA per image row (s = Kab x_rest).
//   Hh[li][b]  = sum_{col : level(col)=lev[li]} Ec[col][b] * x_col     (W*nC multiply-adds per row)
//   s_row[a,b] = Er[row][a] * sum_li Gt[|lev[li]-Y[a][b]|] * Hh[li][b] (nlev*p per row)
// Deterministic: each (li,b) bin is owned by one lane and filled in ascending column order.
__global__ void __launch_bounds__(256)
pass_reduce_kernel(AffinityTables t, const double* __restrict__ x, double* __restrict__ spart) {
    extern __shared__ double smd[];
    const int p = t.p, nC = t.nC, nR = t.nR, W = t.cols;
    double* Gs = smd;                        // 256
    double* Hh = Gs + 256;                   // 256 * nC
    double* xrow = Hh + (size_t)256 * nC;    // W
    int* flags = reinterpret_cast<int*>(xrow + W);
    int* levidx = flags + 256;
    int* lev = levidx + 256;
    int* wcount = lev + 256;
    uint8_t* Ys = reinterpret_cast<uint8_t*>(wcount + 8);
    uint8_t* Lrow = Ys + ((p + 15) / 16) * 16;
    uint8_t* lirow = Lrow + ((W + 15) / 16) * 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < p; i += 256) Ys[i] = t.Ysel[i];
    Gs[tid] = t.Gt[tid];
    for (int rl = blockIdx.x; rl < t.nrows; rl += gridDim.x) {
        const int row = t.row0 + rl;
        const uint8_t* Lg = t.lum + (size_t)rl * W;
        const double* xg = x + (size_t)rl * W;
        __syncthreads();
        for (int c = tid; c < W; c += 256) { Lrow[c] = Lg[c]; xrow[c] = xg[c]; }
        __syncthreads();
        int nlev = build_levels(Lrow, W, flags, levidx, lev, wcount);
        for (int c = tid; c < W; c += 256) lirow[c] = (uint8_t)levidx[Lrow[c]];
        for (int e = tid; e < nlev * nC; e += 256) Hh[e] = 0.0;
        __syncthreads();
        // warp `warp` owns the levels with (li & 7) == warp; lanes own b
        for (int c0 = 0; c0 < W; c0 += 32) {
            int c = c0 + lane;
            int li_l = (c < W) ? (int)lirow[c] : 0;
            bool mine = (c < W) && ((li_l & 7) == warp) && (xrow[c] != 0.0);
            unsigned m = __ballot_sync(0xffffffffu, mine);
            while (m) {
                int j = __ffs(m) - 1;
                m &= m - 1;
                int col = c0 + j;
                int li = __shfl_sync(0xffffffffu, li_l, j);
                double xv = xrow[col];
                const double* ec = t.Ec + (size_t)col * nC;
                double* h = Hh + (size_t)li * nC;
                for (int b = lane; b < nC; b += 32) h[b] = fma(ec[b], xv, h[b]);
            }
        }
        __syncthreads();
        const double* er = t.Er + (size_t)row * nR;
        double* so = spart + (size_t)rl * p;
        for (int i = tid; i < p; i += 256) {
            int a = i / nC, b = i - a * nC;
            int yi = (int)Ys[i];
            double acc = 0.0;
            for (int li = 0; li < nlev; ++li) {
                int d = lev[li] - yi;
                acc = fma(Gs[d < 0 ? -d : d], Hh[(size_t)li * nC + b], acc);
            }
            so[i] = er[a] * acc;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// One Sinkhorn half-iteration over the rest pixels in ONE kernel: the "dot" half (x = recip(k_j^T w))
// followed, for the same image row, by the "reduce" half (s += K(:,j) x_j).  Same arithmetic and
// summation order as pass_dot_kernel + pass_reduce_kernel (bit-identical results); the level list
// is built once, x never has to be re-read from HBM, and one launch replaces two.
__global__ void __launch_bounds__(256)
pass_fused_kernel(AffinityTables t, const double* __restrict__ w, double* __restrict__ x,
                  double* __restrict__ spart) {
    extern __shared__ double smd[];
    const int p = t.p, nC = t.nC, nR = t.nR, W = t.cols;
    double* wp = smd;                        // p          (dot half; reused as the Ec staging area in the reduce half)
    double* Gs = wp + p;                     // 256
    double* F = Gs + 256;                    // 256 * nC   (F table, then reused as the histogram Hh)
    double* xrow = F + (size_t)256 * nC;     // W
    int* flags = reinterpret_cast<int*>(xrow + W);
    int* levidx = flags + 256;
    int* lev = levidx + 256;
    int* wcount = lev + 256;
    uint8_t* Ys = reinterpret_cast<uint8_t*>(wcount + 8);
    uint8_t* Lrow = Ys + ((p + 15) / 16) * 16;
    uint8_t* lirow = Lrow + ((W + 15) / 16) * 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < p; i += 256) Ys[i] = t.Ysel[i];
    Gs[tid] = t.Gt[tid];
    const bool stage_ec = (32 * nC <= p);    // the 32 x nC slice of Ec fits in the wp area (nR >= 32)
    const int ngrp = 256 / nC;               // thread groups of the (column, group) decomposition; 0: generic path
    const int tb = tid % nC, tg = tid / nC;
    for (int rl = blockIdx.x; rl < t.nrows; rl += gridDim.x) {
        const int row = t.row0 + rl;
        const uint8_t* Lg = t.lum + (size_t)rl * W;
        __syncthreads();
        for (int c = tid; c < W; c += 256) Lrow[c] = Lg[c];
        const double* er = t.Er + (size_t)row * nR;
        for (int i = tid; i < p; i += 256) wp[i] = er[i / nC] * w[i];
        __syncthreads();
        const int nlev = build_levels(Lrow, W, flags, levidx, lev, wcount);
        // ---- dot half:  F[li][b] = sum_a wp[a][b] * Gt[|lev[li] - Y[a][b]|]  (ascending a).  Four outputs per
        // thread in flight: the loop is a chain of dependent shared-memory look-ups and one FMA, so a single
        // chain leaves the SM idle (ncu: 45 % short-scoreboard stalls, 24 % issue utilisation).
        const int nF = nlev * nC;
        if (ngrp > 0) {
            // thread = (grid column tb, group tg); 8 levels per thread: wp and Y are read once per 8 multiply-adds
            if (tg < ngrp) {
                for (int l0 = tg * 8; l0 < nlev; l0 += ngrp * 8) {
                    int lv8[8];
                    double acc[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) { lv8[u] = lev[min(l0 + u, nlev - 1)]; acc[u] = 0.0; }
                    for (int a = 0; a < nR; ++a) {
                        const int i = a * nC + tb;
                        const double wv = wp[i];
                        const int yv = (int)Ys[i];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int d = lv8[u] - yv;
                            acc[u] = fma(wv, Gs[d < 0 ? -d : d], acc[u]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (l0 + u < nlev) F[(l0 + u) * nC + tb] = acc[u];
                }
            }
        } else {
            for (int e0 = tid; e0 < nF; e0 += 1024) {
                int bb[4], lvl[4];
                double acc[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = min(e0 + 256 * u, nF - 1);
                    const int li = e / nC;
                    bb[u] = e - li * nC;
                    lvl[u] = lev[li];
                    acc[u] = 0.0;
                }
                for (int a = 0; a < nR; ++a) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = a * nC + bb[u];
                        const int d = lvl[u] - (int)Ys[i];
                        acc[u] = fma(wp[i], Gs[d < 0 ? -d : d], acc[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (e0 + 256 * u < nF) F[e0 + 256 * u] = acc[u];
            }
        }
        __syncthreads();
        const int a_row = t.rowa[row];
        double* xo = x + (size_t)rl * W;
        for (int c = tid; c < W; c += 256) {
            double r = 0.0;
            const int li = levidx[Lrow[c]];
            if (!(a_row >= 0 && t.colb[c] >= 0)) {
                const double* f = F + (size_t)li * nC;
                double acc = 0.0;
                for (int b = 0; b < nC; ++b) acc = fma(t.EcT[(size_t)b * W + c], f[b], acc);
                r = (fabs(acc) >= kEps) ? 1.0 / acc : 0.0;
            }
            xo[c] = r;
            xrow[c] = r;
            lirow[c] = (uint8_t)li;
        }
        __syncthreads();
        // ---- reduce half (F is now the histogram Hh)
        double* Hh = F;
        for (int e = tid; e < nlev * nC; e += 256) Hh[e] = 0.0;
        // Hh[li][b] += Ec[col][b] * x_col, each (li, b) bin owned by one lane and filled in ascending column order.
        // The Ec rows of 32 consecutive pixels are staged in shared memory (prefetched one chunk ahead in
        // registers): reading them straight from L2 inside the serial per-pixel loop cost 29 % of the kernel.
        if (stage_ec) {
            constexpr int NPF = 8;                       // staged elements per thread: 32 * nC <= 256 * NPF (nC <= 64)
            double* stage = wp;
            const int nst = 32 * nC;
            double pf[NPF];
            auto fetch = [&](int c0) {
#pragma unroll
                for (int q = 0; q < NPF; ++q) {
                    const int e = tid + 256 * q;
                    pf[q] = (e < nst && c0 * nC + e < W * nC) ? t.Ec[(size_t)c0 * nC + e] : 0.0;
                }
            };
            const bool wide = nst > 256 * NPF;
            if (!wide) fetch(0);
            for (int c0 = 0; c0 < W; c0 += 32) {
                __syncthreads();                          // previous chunk consumed (and Hh zeroed on the first pass)
                if (!wide) {
#pragma unroll
                    for (int q = 0; q < NPF; ++q) {
                        const int e = tid + 256 * q;
                        if (e < nst) stage[e] = pf[q];
                    }
                    if (c0 + 32 < W) fetch(c0 + 32);
                } else {
                    for (int e = tid; e < nst; e += 256) stage[e] = (c0 * nC + e < W * nC) ? t.Ec[(size_t)c0 * nC + e] : 0.0;
                }
                __syncthreads();
                const int c = c0 + lane;
                const int li_l = (c < W) ? (int)lirow[c] : 0;
                const bool mine = (c < W) && ((li_l & 7) == warp) && (xrow[c] != 0.0);
                unsigned m = __ballot_sync(0xffffffffu, mine);
                while (m) {
                    const int j = __ffs(m) - 1;
                    m &= m - 1;
                    const int li = __shfl_sync(0xffffffffu, li_l, j);
                    const double xv = xrow[c0 + j];
                    const double* ec = stage + j * nC;
                    double* h = Hh + (size_t)li * nC;
                    for (int b = lane; b < nC; b += 32) h[b] = fma(ec[b], xv, h[b]);
                }
            }
        } else {
            __syncthreads();
            for (int c0 = 0; c0 < W; c0 += 32) {
                int c = c0 + lane;
                int li_l = (c < W) ? (int)lirow[c] : 0;
                bool mine = (c < W) && ((li_l & 7) == warp) && (xrow[c] != 0.0);
                unsigned m = __ballot_sync(0xffffffffu, mine);
                while (m) {
                    int j = __ffs(m) - 1;
                    m &= m - 1;
                    int col = c0 + j;
                    int li = __shfl_sync(0xffffffffu, li_l, j);
                    double xv = xrow[col];
                    const double* ec = t.Ec + (size_t)col * nC;
                    double* h = Hh + (size_t)li * nC;
                    for (int b = lane; b < nC; b += 32) h[b] = fma(ec[b], xv, h[b]);
                }
            }
        }
        __syncthreads();
        // s[a][b] = er[a] * sum_li Gt[|lev[li] - Y[a][b]|] * Hh[li][b]  (ascending li), four outputs in flight
        double* so = spart + (size_t)rl * p;
        if (ngrp > 0) {
            // thread = (grid column tb, group tg); 8 grid rows per thread: Hh and the level are read once per 8 FMAs
            if (tg < ngrp) {
                for (int a0 = tg * 8; a0 < nR; a0 += ngrp * 8) {
                    int yy[8];
                    double acc[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) { yy[u] = (int)Ys[min(a0 + u, nR - 1) * nC + tb]; acc[u] = 0.0; }
                    for (int li = 0; li < nlev; ++li) {
                        const int lvl = lev[li];
                        const double h = Hh[(size_t)li * nC + tb];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int d = lvl - yy[u];
                            acc[u] = fma(Gs[d < 0 ? -d : d], h, acc[u]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int a = a0 + u;
                        if (a < nR) so[a * nC + tb] = er[a] * acc[u];
                    }
                }
            }
        } else {
            for (int i0 = tid; i0 < p; i0 += 1024) {
                int bb[4], yy[4];
                double acc[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = min(i0 + 256 * u, p - 1);
                    bb[u] = i % nC;
                    yy[u] = (int)Ys[i];
                    acc[u] = 0.0;
                }
                for (int li = 0; li < nlev; ++li) {
                    const int lvl = lev[li];
                    const double* hrow = Hh + (size_t)li * nC;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int d = lvl - yy[u];
                        acc[u] = fma(Gs[d < 0 ? -d : d], hrow[bb[u]], acc[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + 256 * u;
                    if (i < p) so[i] = er[i / nC] * acc[u];
                }
            }
        }
    }
}

// s[i] = sum over rows of spart[row][i]; two deterministic stages (row chunks, then chunks).
__global__ void colsum_stage_kernel(const double* __restrict__ in, int nrows, int p, int rows_per,
                                    double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int chunk = blockIdx.y;
    if (i >= p) return;
    int r0 = chunk * rows_per, r1 = min(nrows, r0 + rows_per);
    double acc = 0.0;
    for (int r = r0; r < r1; ++r) acc += in[(size_t)r * p + i];
    out[(size_t)chunk * p + i] = acc;
}

void launch_pass_reduce(const AffinityTables& t, const double* x, double* spart, double* s_out,
                        cudaStream_t s) {
    size_t smem = pass_smem_bytes(t, true);
    if (smem > 227 * 1024) throw Unsupported{"pass_reduce: sample grid / image too wide for the per-row table (nC=" + std::to_string(t.nC) + ", cols=" + std::to_string(t.cols) + ")"};
    static size_t configured = 0;
    if (smem > configured) {
        NLE_CUDA(cudaFuncSetAttribute(pass_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    pass_reduce_kernel<<<t.nrows, 256, smem, s>>>(t, x, spart);
    NLE_LAUNCH_CHECK();
    // stage 1: chunks of 32 rows written back into the head of spart's own rows is unsafe; use the
    // tail-free trick: chunk c's result goes to row c (c <= first row of the chunk, already consumed
    // by this same thread order) -- done in two separate launches to stay race-free.
    int rows_per = 32;
    int nchunks = cdiv(t.nrows, rows_per);
    double* stage = spart + (size_t)t.nrows * t.p;   // scratch tail: nchunks * p doubles
    colsum_stage_kernel<<<dim3(cdiv(t.p, 128), nchunks), 128, 0, s>>>(spart, t.nrows, t.p, rows_per, stage);
    NLE_LAUNCH_CHECK();
    colsum_stage_kernel<<<dim3(cdiv(t.p, 128), 1), 128, 0, s>>>(stage, nchunks, t.p, nchunks, s_out);
    NLE_LAUNCH_CHECK();
}

void launch_pass_fused(const AffinityTables& t, const double* w, double* x, double* spart, double* s_out,
                       cudaStream_t s) {
    size_t smem = pass_smem_bytes(t, true) + (size_t)t.p * 8;
    if (smem > 227 * 1024) {   // very wide grids: fall back to the two separate passes
        launch_pass_dot(t, w, x, s);
        launch_pass_reduce(t, x, spart, s_out, s);
        return;
    }
    static size_t configured = 0;
    if (smem > configured) {
        NLE_CUDA(cudaFuncSetAttribute(pass_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    pass_fused_kernel<<<t.nrows, 256, smem, s>>>(t, w, x, spart);
    NLE_LAUNCH_CHECK();
    int rows_per = 32;
    int nchunks = cdiv(t.nrows, rows_per);
    double* stage = spart + (size_t)t.nrows * t.p;
    colsum_stage_kernel<<<dim3(cdiv(t.p, 128), nchunks), 128, 0, s>>>(spart, t.nrows, t.p, rows_per, stage);
    NLE_LAUNCH_CHECK();
    colsum_stage_kernel<<<dim3(cdiv(t.p, 128), 1), 128, 0, s>>>(stage, nchunks, t.p, nchunks, s_out);
    NLE_LAUNCH_CHECK();
}

// =============================================================================================
// Weighted Gram  G = sum_j c_j^2 k_j k_j^T   (the "Wab Wab^T" of filter.cpp:296 in factor form,
// SURVEY App. A.5).  FP64 SYRK whose K dimension is the pixel axis; the operand tiles
// a_ij = c_j K(i,j) are generated in shared memory and never touch HBM.
// CTA = 128x128 output tile (upper-triangular tile pairs only) x one contiguous range of image rows.
constexpr int GT = 128;    // tile edge (samples)
constexpr int GKC = 16;    // pixels per chunk
constexpr int GLD = GT + 8;   // padded row: the 4 k-rows of a DMMA fragment fall into 2 disjoint bank groups

// CE[j][b] = c_j * Ec[col_j][b] for every slab pixel j (nloc x nC): takes the per-pixel Sinkhorn weight out of
// the tile producers' critical burst (see gram_kernel).
__global__ void gram_ce_kernel(AffinityTables t, const double* __restrict__ cvec, double* __restrict__ CE) {
    const long long total = (long long)t.nrows * t.cols * t.nC;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long j = e / t.nC;
        const int b = (int)(e - j * t.nC);
        const int col = (int)(j % t.cols);
        CE[e] = cvec[j] * t.Ec[(size_t)col * t.nC + b];
    }
}

// Warp-specialised: 16 warps = 8 PRODUCER warps (two per SM sub-core) that generate the operand tile
// a_ij = c_j K(i,j) in shared memory, and 8 CONSUMER warps (two per sub-core) that do nothing but
// LDS + DMMA.  Findings that shaped it (ncu captures summarised under profiles/):
//   * the FP64 tensor unit is per sub-core and one DMMA.8x8x4 occupies it for 16 cycles (ptxas pads the
//     DMMA stream with NOPs accordingly), so the two consumer warps of a sub-core can saturate it;
//   * an FP64 multiply issued while the DMMA stream is running queues behind it for 100-170 cycles
//     ("math pipe throttle" was 46-61 % of the producers' time in every variant that multiplied
//     concurrently, even with a single multiply per entry), which made the producers the bottleneck.
// So the producers do all their loads (CE, levels, table look-ups) while the consumers run the DMMAs of
// chunk c, keep the raw factors in registers, and multiply + store chunk c+1 in a short burst between
// the consumers' "done" (barrier 2) and "tile full" (barrier 1) -- the only time the FP64 pipe is shared.
// Entry:  a = (CE[j][b_i] * Gt[|l_j - Y_i|]) * Er[row][a_i]  with CE[j][b] = c_j * Ec[col_j][b], a global table
// written by gram_ce_kernel and streamed through L2.  (A per-row shared-memory table Er*Gt that brings the
// burst down to one multiply per entry was measured too: same speed, so the simpler form is kept.)
// Register file: setmaxnreg gives the producers 88 and the consumers 168 registers per thread
// (512 * 128 = 65536 at launch).
// Consumers: 4 (M) x 2 (N) warps, warp tile 32 x 64 = 4 x 8 DMMA tiles, 64 FP64 accumulators per lane;
// per 4-pixel step a warp issues 12 LDS.64 and 32 DMMA (8192 FMA).
constexpr int GRAM_THREADS = 512;
constexpr int GRAM_PRODUCERS = 256;

__global__ void __launch_bounds__(GRAM_THREADS, 1)
gram_kernel(AffinityTables t, const double* __restrict__ CE, int ntile, int nsplit,
            double* __restrict__ part) {
    extern __shared__ double gsm[];
    double (*As)[GKC][GLD] = reinterpret_cast<double (*)[GKC][GLD]>(gsm);                     // [2][GKC][GLD]
    double (*Bs)[GKC][GLD] = reinterpret_cast<double (*)[GKC][GLD]>(gsm + 2 * GKC * GLD);     // [2][GKC][GLD]
    double* Gs = gsm + 4 * GKC * GLD;                                                         // [256]

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int p = t.p, nC = t.nC, nR = t.nR, W = t.cols;
    // tile pair from linear upper-triangular index
    int tp = blockIdx.x, ti = 0;
    {
        int rem = tp;
        while (rem >= ntile - ti) { rem -= ntile - ti; ++ti; }
        tp = rem;
    }
    const int tj = ti + tp;
    const bool diag = (ti == tj);
    const int split = blockIdx.y;
    const int rb = (int)(((long long)t.nrows * split) / nsplit);
    const int re = (int)(((long long)t.nrows * (split + 1)) / nsplit);
    const int chunks_per_row = (W + GKC - 1) / GKC;
    const long long nchunks = (long long)(re - rb) * chunks_per_row;

    if (tid < 256) Gs[tid] = t.Gt[tid];
    __syncthreads();

    if (tid < GRAM_PRODUCERS) {
        // ------------------------------------------------------------------ producers
        asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
        // thread = (sample sIdx of each tile, pixel half ph): 8 (+8) tile entries per chunk
        const int sIdx = tid & 127, ph = tid >> 7;
        const int iA = ti * GT + sIdx, iB = tj * GT + sIdx;
        const bool vA = iA < p, vB = (iB < p) && !diag;
        const int aA = vA ? iA / nC : 0, bA = vA ? iA - aA * nC : 0;
        const int aB = vB ? iB / nC : 0, bB = vB ? iB - aB * nC : 0;
        const int yA = vA ? (int)t.Ysel[iA] : 0, yB = vB ? (int)t.Ysel[iB] : 0;
        const bool lum_vec = (W % 8 == 0) && ((reinterpret_cast<size_t>(t.lum) & 7) == 0);
        double erA = 0.0, erB = 0.0;
        double ceA[8], ceB[8], gA[8], gB[8];
        int prow = rb, pcol0 = 0;      // chunk cursor (no 64-bit divisions in the loop)
        // all loads of one chunk: CE values, pixel levels -> Gt look-ups, Er on a new row
        auto prep = [&]() {
            const size_t j0 = (size_t)prow * W + pcol0 + ph * 8;
            const double* ce = CE + j0 * nC;
            const uint8_t* lv = t.lum + j0;
            const int nok = min(8, W - (pcol0 + ph * 8));      // valid pixels among this thread's 8 (<= 0: none)
            // ---- issue every global load first (one 64-bit load for the 8 levels when the row allows it)
            unsigned long long lw = 0ull;
            if (lum_vec) {
                if (nok > 0) lw = *reinterpret_cast<const unsigned long long*>(lv);
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (q < nok) lw |= (unsigned long long)lv[q] << (8 * q);
            }
            if (pcol0 == 0) {
                const double* er = t.Er + (size_t)(t.row0 + prow) * nR;
                erA = vA ? er[aA] : 0.0;
                erB = vB ? er[aB] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                ceA[q] = (q < nok && vA) ? ce[(size_t)q * nC + bA] : 0.0;
                if (!diag) ceB[q] = (q < nok && vB) ? ce[(size_t)q * nC + bB] : 0.0;
            }
            // ---- then the Gt look-ups (shared memory)
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int l = (int)((lw >> (8 * q)) & 0xffull);
                const int dA = l - yA;
                gA[q] = Gs[dA < 0 ? -dA : dA];
                if (!diag) {
                    const int dB = l - yB;
                    gB[q] = Gs[dB < 0 ? -dB : dB];
                }
            }
            pcol0 += GKC;
            if (pcol0 >= W) { pcol0 = 0; ++prow; }
        };
        // the only FP64 arithmetic of the producers: 2 multiplies per entry, issued while the DMMA stream is idle
        auto burst = [&](int buf) {
#pragma unroll
            for (int q = 0; q < 8; ++q) As[buf][ph * 8 + q][sIdx] = (ceA[q] * gA[q]) * erA;
            if (!diag) {
#pragma unroll
                for (int q = 0; q < 8; ++q) Bs[buf][ph * 8 + q][sIdx] = (ceB[q] * gB[q]) * erB;
            }
        };
        if (nchunks > 0) {
            prep();
            burst(0);
            asm volatile("bar.arrive 1, 512;" ::: "memory");          // tile 0 full
        }
        for (long long ch = 0; ch < nchunks; ++ch) {
            const bool more = ch + 1 < nchunks;
            if (more) prep();                                          // overlaps the consumers' DMMAs of chunk ch
            // the consumers signal when only the last quarter of chunk ch is left: the burst (into the other tile
            // buffer) overlaps that tail instead of leaving the tensor pipe idle
            asm volatile("bar.sync 2, 512;" ::: "memory");
            if (more) {
                burst((int)((ch + 1) & 1));
                asm volatile("bar.arrive 1, 512;" ::: "memory");      // tile ch+1 full
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
    const int cw = warp - GRAM_PRODUCERS / 32;
    const int g = lane >> 2, tq = lane & 3;
    const int wm = cw & 3, wn = cw >> 2;
    double acc[4][8][2];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 8; ++v) acc[u][v][0] = acc[u][v][1] = 0.0;
    // A warp whose 32 x 64 sub-tile lies entirely in the padding (p is not a multiple of 128) or strictly
    // below the diagonal of a diagonal tile (only i <= j is ever read back) issues no DMMA.
    const bool wactive = (ti * GT + wm * 32 < p) && (tj * GT + wn * 64 < p) && !(diag && wm * 32 >= wn * 64 + 64);
    for (long long ch = 0; ch < nchunks; ++ch) {
        const int buf = (int)(ch & 1);
        const double (*Ap)[GLD] = As[buf];
        const double (*Bp)[GLD] = diag ? As[buf] : Bs[buf];
        asm volatile("bar.sync 1, 512;" ::: "memory");                 // tile ch full
#pragma unroll
        for (int k4 = 0; k4 < GKC / 4; ++k4) {
            if (k4 == GKC / 4 - 1) asm volatile("bar.arrive 2, 512;" ::: "memory");   // last quarter: producers may burst
                                                                                          // (signalling at half was measured slower)
            if (wactive) {
                const double* ar = &Ap[k4 * 4 + tq][wm * 32 + g];
                const double* br = &Bp[k4 * 4 + tq][wn * 64 + g];
                double a[4], b[8];
#pragma unroll
                for (int u = 0; u < 4; ++u) a[u] = ar[u * 8];
#pragma unroll
                for (int v = 0; v < 8; ++v) b[v] = br[v * 8];
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int v = 0; v < 8; ++v) dmma884(acc[u][v][0], acc[u][v][1], a[u], b[v]);
            }
        }
    }
    double* out = part + ((size_t)split * gridDim.x + blockIdx.x) * (GT * GT);
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            const int row = wm * 32 + u * 8 + g, col = wn * 64 + v * 8 + 2 * tq;
            *reinterpret_cast<double2*>(out + row * GT + col) = make_double2(acc[u][v][0], acc[u][v][1]);
        }
}

__global__ void gram_reduce_kernel(const double* __restrict__ part, int p, int ntile, int ntp,
                                   int nsplit, double* __restrict__ G) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= p) return;
    int ti = i / GT, tj = j / GT;
    int r = i % GT, c = j % GT;
    // only the upper triangle (tile pairs ti <= tj, and r <= c inside diagonal tiles) is computed
    if (ti > tj || (ti == tj && r > c)) { int tmp = ti; ti = tj; tj = tmp; tmp = r; r = c; c = tmp; }
    // linear index of (ti,tj), ti<=tj : sum_{q<ti} (ntile-q) + (tj-ti)
    int tp = ti * ntile - (ti * (ti - 1)) / 2 + (tj - ti);
    double acc = 0.0;
    for (int s = 0; s < nsplit; ++s) acc += part[((size_t)s * ntp + tp) * (GT * GT) + r * GT + c];
    G[i + (size_t)j * p] = acc;
}

static void gram_geometry(const AffinityTables& t, int& ntile, int& ntp, int& nsplit) {
    ntile = cdiv(t.p, GT);
    ntp = ntile * (ntile + 1) / 2;
    nsplit = (8 * sm_count()) / ntp;
    if (nsplit < 1) nsplit = 1;
    if (nsplit > t.nrows) nsplit = t.nrows;
    if (nsplit < 1) nsplit = 1;
}

// Default: the cell-contracted Gram (gram_cells.cu).  NLE_B200_GRAM=pixel selects the pixel-axis SYRK below
// (same result up to FP64 re-association; kept as a cross-check and for the parity tests).
static bool gram_pixel_path() {
    static const bool v = [] { const char* e = getenv("NLE_B200_GRAM"); return e && std::string(e) == "pixel"; }();
    return v;
}

size_t gram_scratch_doubles(const AffinityTables& t) {
    if (!gram_pixel_path()) return gram_cells_scratch_doubles(t);
    int ntile, ntp, nsplit;
    gram_geometry(t, ntile, ntp, nsplit);
    return (size_t)ntp * nsplit * GT * GT + (size_t)t.nrows * t.cols * t.nC;   // partial tiles + CE table
}

void launch_gram(const AffinityTables& t, const double* c, double* scratch, double* G, cudaStream_t s) {
    if (!gram_pixel_path()) { launch_gram_cells(t, c, scratch, G, s); return; }
    int ntile, ntp, nsplit;
    gram_geometry(t, ntile, ntp, nsplit);
    double* CE = scratch + (size_t)ntp * nsplit * GT * GT;
    gram_ce_kernel<<<sm_count() * 8, 256, 0, s>>>(t, c, CE);
    NLE_LAUNCH_CHECK();
    const size_t smem = (size_t)(4 * GKC * GLD + 256) * sizeof(double);
    static size_t configured = 0;
    if (smem > configured) {
        NLE_CUDA(cudaFuncSetAttribute(gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    gram_kernel<<<dim3(ntp, nsplit), GRAM_THREADS, smem, s>>>(t, CE, ntile, nsplit, scratch);
    NLE_LAUNCH_CHECK();
    gram_reduce_kernel<<<dim3(cdiv(t.p, 128), t.p), 128, 0, s>>>(scratch, t.p, ntile, ntp, nsplit, G);
    NLE_LAUNCH_CHECK();
}

// =============================================================================================
// Eigenvector extension for the non-sample pixels (filter.cpp:324-327 in factor form, App. A.6):
//   V_j = c_j * sum_i K(i,j) Y[i][:]
constexpr int EV = 32;    // eigenvector columns per CTA
constexpr int EI = 128;   // samples per shared-memory chunk
constexpr int EPX = 2;    // pixels per thread (register blocking over the broadcast Y loads)

__global__ void __launch_bounds__(256)
extension_kernel(AffinityTables t, const double* __restrict__ cvec, const double* __restrict__ Y,
                 int k, double* __restrict__ V) {
    __shared__ double2 Ys[EI][EV / 2 + 1];
    __shared__ double Gs[256];
    __shared__ int sa[EI], sb[EI], sy[EI];
    const int tid = threadIdx.x;
    const int p = t.p, nC = t.nC, nR = t.nR, W = t.cols;
    const long long nloc = (long long)t.nrows * W;
    const int v0 = blockIdx.y * EV;
    const int nv = min(EV, k - v0);
    long long jj[EPX];
    bool live[EPX];
    int lv[EPX];
    double cj[EPX];
    const double* er[EPX];
    const double* ec[EPX];
    bool any = false;
#pragma unroll
    for (int q = 0; q < EPX; ++q) {
        jj[q] = ((long long)blockIdx.x * EPX + q) * 256 + tid;
        live[q] = jj[q] < nloc;
        int rl = 0, col = 0;
        lv[q] = 0; cj[q] = 0.0;
        if (live[q]) {
            rl = (int)(jj[q] / W); col = (int)(jj[q] - (long long)rl * W);
            lv[q] = (int)t.lum[jj[q]];
            cj[q] = cvec[jj[q]];
            if (t.rowa[t.row0 + rl] >= 0 && t.colb[col] >= 0) live[q] = false;   // sample pixel: scattered separately
        }
        er[q] = t.Er + (size_t)(t.row0 + rl) * nR;
        ec[q] = t.Ec + (size_t)col * nC;
        any = any || (live[q] && cj[q] != 0.0);
    }
    Gs[tid] = t.Gt[tid];
    double acc[EPX][EV];
#pragma unroll
    for (int q = 0; q < EPX; ++q)
#pragma unroll
        for (int v = 0; v < EV; ++v) acc[q][v] = 0.0;
    double* Ysd = reinterpret_cast<double*>(&Ys[0][0]);
    constexpr int YLD = 2 * (EV / 2 + 1);
    for (int i0 = 0; i0 < p; i0 += EI) {
        __syncthreads();
        for (int e = tid; e < EI * EV; e += 256) {
            int sI = e & (EI - 1), v = e >> 7;          // EI == 128
            int i = i0 + sI;
            Ysd[sI * YLD + v] = (i < p && v < nv) ? Y[i + (size_t)(v0 + v) * p] : 0.0;
        }
        if (tid < EI) {
            int i = i0 + tid;
            if (i < p) { sa[tid] = i / nC; sb[tid] = i % nC; sy[tid] = (int)t.Ysel[i]; }
            else { sa[tid] = 0; sb[tid] = 0; sy[tid] = 0; }
        }
        __syncthreads();
        const int ni = min(EI, p - i0);
        if (any) {
            for (int sI = 0; sI < ni; ++sI) {
                double kv[EPX];
#pragma unroll
                for (int q = 0; q < EPX; ++q) {
                    int d = lv[q] - sy[sI];
                    kv[q] = er[q][sa[sI]] * ec[q][sb[sI]] * Gs[d < 0 ? -d : d];
                }
#pragma unroll
                for (int v2 = 0; v2 < EV / 2; ++v2) {
                    double2 y2 = Ys[sI][v2];
#pragma unroll
                    for (int q = 0; q < EPX; ++q) {
                        acc[q][2 * v2] = fma(kv[q], y2.x, acc[q][2 * v2]);
                        acc[q][2 * v2 + 1] = fma(kv[q], y2.y, acc[q][2 * v2 + 1]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < EPX; ++q) {
        if (!live[q]) continue;
        double* vo = V + (size_t)jj[q] * k + v0;
#pragma unroll
        for (int v = 0; v < EV; ++v)
            if (v < nv) vo[v] = cj[q] * acc[q][v];
    }
}

// ---------------------------------------------------------------------------------------------
// Extension on the FP64 tensor pipe:  V(pixels x k) = diag(c) * K_B^T(pixels x p) * Y(p x k).
// Same organisation as gram_kernel (see there for the measurements behind it): 8 producer warps generate the
// operand tile K(i,j) for 256 pixels x 16 samples -- loads and the first multiply (Ec*Er) while the consumers
// run the DMMAs of the previous chunk, the last multiply (*Gt) and the stores in a burst between two named
// barriers -- and stream the matching 16 x 56 slice of Y (repacked row-major, zero padded) into a double
// buffer; 8 consumer warps hold a 32 x 56 block of V each (4 x 7 DMMA tiles, 56 accumulators per lane).
constexpr int XM = 256;       // pixels per CTA
constexpr int XN = 56;        // eigenvector columns per CTA (7 DMMA tiles); 56 % 16 == 8 keeps B fragments conflict-free
constexpr int XK = 16;        // samples per chunk
constexpr int XLD = XM + 8;

__global__ void ext_pack_y_kernel(const double* __restrict__ Y, int p, int k, int kp, double* __restrict__ Yt) {
    // Yt[i][v] = Y[i + v*p] for v < k, 0 for k <= v < kp   (row-major, ld kp)
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)p * kp) return;
    const int i = (int)(e / kp), v = (int)(e - (long long)i * kp);
    Yt[e] = (v < k) ? Y[i + (size_t)v * p] : 0.0;
}

__global__ void __launch_bounds__(512, 1)
extension_dmma_kernel(AffinityTables t, const double* __restrict__ cvec, const double* __restrict__ Yt, int kp,
                      int k, double* __restrict__ V) {
    extern __shared__ double xsm[];
    double (*As)[XK][XLD] = reinterpret_cast<double (*)[XK][XLD]>(xsm);                     // [2][XK][XLD]
    double (*Bs)[XK][XN] = reinterpret_cast<double (*)[XK][XN]>(xsm + 2 * XK * XLD);        // [2][XK][XN]
    double* Gs = xsm + 2 * XK * XLD + 2 * XK * XN;                                          // [256]
    uint8_t* Ys = reinterpret_cast<uint8_t*>(Gs + 256);                                 // [p]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int p = t.p, nC = t.nC, nR = t.nR, W = t.cols;
    const long long nloc = (long long)t.nrows * W;
    const long long j0 = (long long)blockIdx.x * XM;
    const int v0 = blockIdx.y * XN;
    const int nchunks = (p + XK - 1) / XK;
    if (tid < 256) Gs[tid] = t.Gt[tid];
    for (int i = tid; i < p; i += 512) Ys[i] = t.Ysel[i];
    __syncthreads();

    if (tid < 256) {
        // ------------------------------------------------------------------ producers (thread = pixel)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
        const long long j = j0 + tid;
        const bool live = j < nloc;
        const int rl = live ? (int)(j / W) : 0, col = live ? (int)(j - (long long)rl * W) : 0;
        const int lv = live ? (int)t.lum[j] : 0;
        const double* er = t.Er + (size_t)(t.row0 + rl) * nR;
        const double* ec = t.Ec + (size_t)col * nC;
        double sp[XK], gg[XK];
        int pa = 0, pb = 0, pi = 0;        // grid coordinates / index of the first sample of the chunk being prepared
        auto prep = [&](int buf) {
            // Y slice of this chunk -> B double buffer (896 doubles, 3.5 per thread, coalesced rows of Yt)
            for (int e = tid; e < XK * XN; e += 256) {
                const int kq = e / XN, n = e - kq * XN;
                const int i = pi + kq;
                Bs[buf][kq][n] = (i < p) ? Yt[(size_t)i * kp + v0 + n] : 0.0;
            }
            int a = pa, b = pb;
#pragma unroll
            for (int q = 0; q < XK; ++q) {
                const bool ok = live && (pi + q < p);
                const double e1 = ok ? ec[b] : 0.0;
                const double e2 = ok ? er[a] : 0.0;
                int d = lv - (int)Ys[min(pi + q, p - 1)];
                d = d < 0 ? -d : d;
                gg[q] = Gs[d];
                sp[q] = e1 * e2;        // spatial factor: off the critical path (overlaps the consumers' DMMAs)
                if (++b == nC) { b = 0; ++a; }
            }
            pa = a; pb = b; pi += XK;
        };
        auto burst = [&](int buf) {
#pragma unroll
            for (int q = 0; q < XK; ++q) As[buf][q][tid] = sp[q] * gg[q];
        };
        if (nchunks > 0) {
            prep(0);
            burst(0);
            asm volatile("bar.arrive 1, 512;" ::: "memory");
        }
        for (int ch = 0; ch < nchunks; ++ch) {
            const bool more = ch + 1 < nchunks;
            if (more) prep((ch + 1) & 1);
            asm volatile("bar.sync 2, 512;" ::: "memory");      // consumers are in the last quarter of chunk ch
            if (more) {
                burst((ch + 1) & 1);
                asm volatile("bar.arrive 1, 512;" ::: "memory");
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
    const int cw = warp - 8;
    const int g = lane >> 2, tq = lane & 3;
    double acc[4][7][2];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 7; ++v) acc[u][v][0] = acc[u][v][1] = 0.0;
    for (int ch = 0; ch < nchunks; ++ch) {
        const int buf = ch & 1;
        asm volatile("bar.sync 1, 512;" ::: "memory");
#pragma unroll
        for (int k4 = 0; k4 < XK / 4; ++k4) {
            if (k4 == XK / 4 - 1) asm volatile("bar.arrive 2, 512;" ::: "memory");
            const double* ar = &As[buf][k4 * 4 + tq][cw * 32 + g];
            const double* br = &Bs[buf][k4 * 4 + tq][g];
            double a[4], b[7];
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] = ar[u * 8];
#pragma unroll
            for (int v = 0; v < 7; ++v) b[v] = br[v * 8];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 7; ++v) dmma884(acc[u][v][0], acc[u][v][1], a[u], b[v]);
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const long long j = j0 + cw * 32 + u * 8 + g;
        if (j >= nloc) continue;
        const int rl = (int)(j / W), col = (int)(j - (long long)rl * W);
        if (t.rowa[t.row0 + rl] >= 0 && t.colb[col] >= 0) continue;     // sample pixel: scattered separately
        const double cj = cvec[j];
        double* vo = V + (size_t)j * k;
#pragma unroll
        for (int v = 0; v < 7; ++v) {
            const int n = v0 + v * 8 + 2 * tq;
            if (n < k) vo[n] = cj * acc[u][v][0];
            if (n + 1 < k) vo[n + 1] = cj * acc[u][v][1];
        }
    }
}

void launch_extension(const AffinityTables& t, const double* c, const double* Y, int k, double* V,
                      cudaStream_t s) {
    long long nloc = (long long)t.nrows * t.cols;
    if (k <= 0 || nloc <= 0) return;
    static const bool legacy = getenv("NLE_B200_EXT_LEGACY") != nullptr;
    if (legacy) {
        extension_kernel<<<dim3(cdiv(nloc, 256 * EPX), cdiv(k, EV)), 256, 0, s>>>(t, c, Y, k, V);
        NLE_LAUNCH_CHECK();
        return;
    }
    const int nvb = cdiv(k, XN), kp = nvb * XN;
    DevBuf<double> Yt((size_t)t.p * kp);   // stream-ordered pool: freed after the kernel in stream order
    ext_pack_y_kernel<<<cdiv((long long)t.p * kp, 256), 256, 0, s>>>(Y, t.p, k, kp, Yt.p);
    NLE_LAUNCH_CHECK();
    const size_t smem = (size_t)(2 * XK * XLD + 2 * XK * XN + 256) * sizeof(double) + ((t.p + 15) / 16) * 16;
    static size_t configured = 0;
    if (smem > configured) {
        NLE_CUDA(cudaFuncSetAttribute(extension_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    extension_dmma_kernel<<<dim3(cdiv(nloc, XM), nvb), 512, smem, s>>>(t, c, Yt.p, kp, k, V);
    NLE_LAUNCH_CHECK();
}

__global__ void scatter_rows_kernel(AffinityTables t, const int32_t* __restrict__ sel, int i0, int n,
                                    const double* __restrict__ src, int ld, int k, double* __restrict__ V) {
    int i = blockIdx.x;          // sample within [0,n)
    long long pix = sel[i0 + i];
    int row = (int)(pix / t.cols);
    if (row < t.row0 || row >= t.row0 + t.nrows) return;
    long long loc = pix - (long long)t.row0 * t.cols;
    for (int v = threadIdx.x; v < k; v += blockDim.x) V[(size_t)loc * k + v] = src[i + (size_t)v * ld];
}

void launch_scatter_rows(const AffinityTables& t, const int32_t* sel, int i0, int n, const double* src,
                         int ld, int k, double* V, cudaStream_t s) {
    if (n <= 0 || k <= 0) return;
    scatter_rows_kernel<<<n, 64, 0, s>>>(t, sel, i0, n, src, ld, k, V);
    NLE_LAUNCH_CHECK();
}

// =============================================================================================
// apply (filter.cpp:445-458):  out = V (g o (V^T z)),  V row-major (N x k).
constexpr int AP_PIX = 1024;   // pixels per CTA in the V^T z pass

int apply_blocks(long long nloc) { return cdiv(nloc, AP_PIX); }

__global__ void __launch_bounds__(256)
vtz_kernel(long long nloc, int k, const double* __restrict__ V, const uint8_t* __restrict__ z8,
           const double* __restrict__ z64, double* __restrict__ partial) {
    // thread v-lane layout: 256 threads = 8 pixel-lanes x 32 v-lanes; each reads V rows coalesced.
    extern __shared__ double red[];   // 8 * kpad
    const int vl = threadIdx.x & 31, pl = threadIdx.x >> 5;
    const long long base = (long long)blockIdx.x * AP_PIX;
    const int kpad = ((k + 31) / 32) * 32;
    for (int vb = 0; vb < k; vb += 32) {
        int v = vb + vl;
        double acc = 0.0;
        if (v < k) {
            for (int q = pl; q < AP_PIX; q += 8) {
                long long j = base + q;
                if (j >= nloc) break;
                double z = z8 ? (double)z8[j] : z64[j];
                acc = fma(V[(size_t)j * k + v], z, acc);
            }
        }
        red[pl * kpad + vb + vl] = acc;
    }
    __syncthreads();
    for (int v = threadIdx.x; v < k; v += 256) {
        double s = 0.0;
        for (int q = 0; q < 8; ++q) s += red[q * kpad + v];
        partial[(size_t)blockIdx.x * k + v] = s;
    }
}

__global__ void vtz_final_kernel(const double* __restrict__ partial, int nblocks, int k,
                                 double* __restrict__ tout) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= k) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * k + v];
    tout[v] = s;
}

void launch_vtz(long long nloc, int k, const double* V, const uint8_t* z_u8, const double* z_f64,
                double* scratch, double* t_out, cudaStream_t s) {
    int nb = apply_blocks(nloc);
    int kpad = ((k + 31) / 32) * 32;
    vtz_kernel<<<nb, 256, (size_t)8 * kpad * sizeof(double), s>>>(nloc, k, V, z_u8, z_f64, scratch);
    NLE_LAUNCH_CHECK();
    vtz_final_kernel<<<cdiv(k, 64), 64, 0, s>>>(scratch, nb, k, t_out);
    NLE_LAUNCH_CHECK();
}

__global__ void __launch_bounds__(256)
recompose_kernel(long long nloc, int k, const double* __restrict__ V, const double* __restrict__ g,
                 double* __restrict__ out64, uint8_t* __restrict__ out8) {
    // one warp per group of pixels; lanes stride over k so V rows are read coalesced
    extern __shared__ double gs[];
    for (int v = threadIdx.x; v < k; v += 256) gs[v] = g[v];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    long long warp = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;
    long long nwarps = ((long long)gridDim.x * 256) >> 5;
    for (long long j = warp; j < nloc; j += nwarps) {
        const double* vr = V + (size_t)j * k;
        double acc = 0.0;
        for (int v = lane; v < k; v += 32) acc = fma(vr[v], gs[v], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            if (out64) out64[j] = acc;
            if (out8) {
                // cv::max(.,0), cv::min(.,255), convertTo(CV_8U) = cvRound = round half to even
                double c = fmin(fmax(acc, 0.0), 255.0);
                out8[j] = (uint8_t)__double2int_rn(c);
            }
        }
    }
}

void launch_recompose(long long nloc, int k, const double* V, const double* g, double* out_f64,
                      uint8_t* out_u8, cudaStream_t s) {
    if (nloc <= 0) return;
    long long warps_needed = nloc;
    int grid = (int)std::min<long long>((warps_needed + 7) / 8, (long long)sm_count() * 16);
    if (grid < 1) grid = 1;
    recompose_kernel<<<grid, 256, (size_t)(k > 0 ? k : 1) * sizeof(double), s>>>(nloc, k, V, g, out_f64, out_u8);
    NLE_LAUNCH_CHECK();
}

// =============================================================================================
__global__ void fill_kernel(double* p, long long n, double v) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
void launch_fill(double* p, long long n, double v, cudaStream_t s) {
    if (n <= 0) return;
    fill_kernel<<<cdiv(n, 256), 256, 0, s>>>(p, n, v);
    NLE_LAUNCH_CHECK();
}

__global__ void mask_samples_kernel(AffinityTables t, double* x) {
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= (long long)t.nrows * t.cols) return;
    int rl = (int)(j / t.cols), col = (int)(j - (long long)rl * t.cols);
    if (t.rowa[t.row0 + rl] >= 0 && t.colb[col] >= 0) x[j] = 0.0;
}
void launch_mask_samples(const AffinityTables& t, double* x, cudaStream_t s) {
    long long n = (long long)t.nrows * t.cols;
    if (n <= 0) return;
    mask_samples_kernel<<<cdiv(n, 256), 256, 0, s>>>(t, x);
    NLE_LAUNCH_CHECK();
}

__global__ void u8_from_f64_kernel(const double* __restrict__ in, long long n, uint8_t* __restrict__ out,
                                   int* __restrict__ bad) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = in[i];
    double r = rint(v);
    if (!(v >= 0.0 && v <= 255.0) || r != v) atomicExch(bad, 1);
    out[i] = (uint8_t)(int)fmin(fmax(r, 0.0), 255.0);
}
void launch_u8_from_f64(const double* in, long long n, uint8_t* out, int* bad_flag, cudaStream_t s) {
    if (n <= 0) return;
    u8_from_f64_kernel<<<cdiv(n, 256), 256, 0, s>>>(in, n, out, bad_flag);
    NLE_LAUNCH_CHECK();
}

__global__ void gather_c_sel_kernel(AffinityTables t, const int32_t* __restrict__ sel,
                                    const double* __restrict__ c_sel, double* __restrict__ c_full) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t.p) return;
    long long pix = sel[i];
    int row = (int)(pix / t.cols);
    if (row < t.row0 || row >= t.row0 + t.nrows) return;
    c_full[pix - (long long)t.row0 * t.cols] = c_sel[i];
}
void launch_gather_c_sel(const AffinityTables& t, const int32_t* sel, const double* c_sel, double* c_full,
                         cudaStream_t s) {
    gather_c_sel_kernel<<<cdiv(t.p, 128), 128, 0, s>>>(t, sel, c_sel, c_full);
    NLE_LAUNCH_CHECK();
}

}  // namespace nle
