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
// This file holds the table / Ka / sampling kernels, the un-permute scatter, the two HBM-bound apply passes and
// small element-wise helpers; the N-scaled contractions (Sinkhorn passes, Gram, extension) are the (image row,
// luminance level) cell kernels of sinkhorn_cells.cu and cell_kernels.cu.
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
// V[pixel(sel[i0+i])][:] = src(i,:) for the samples that fall in the slab (un-permute of filter.cpp:502).
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
__global__ void fill_kernel(double* p, long long n, double v) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
void launch_fill(double* p, long long n, double v, cudaStream_t s) {
    if (n <= 0) return;
    fill_kernel<<<cdiv(n, 256), 256, 0, s>>>(p, n, v);
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
