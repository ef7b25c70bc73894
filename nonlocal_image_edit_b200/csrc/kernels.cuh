// Launchers of the N-scaled kernels (filter_kernels.cu, cell_kernels.cu, sinkhorn_cells.cu).  All device pointers.
#pragma once
#include "common.cuh"

namespace nle {

// Everything the N-scaled kernels need to evaluate K(i,j) for sample i=(a,b) and pixel j=(row,col,l):
//   K = Er[row][a] * Ec[col][b] * Gt[|l - Ysel[i]|]      (filter.cpp:104-112,128-129,144-145)
// exact because samples form a product grid (filter.cpp:68-70) and L is 8-bit (filter.cpp:463-466).
struct AffinityTables {
    int rows, cols;          // full image
    int row0, nrows;         // slab owned by this rank: image rows [row0, row0+nrows)
    int nR, nC, p;           // sample grid and p = nR*nC
    const uint8_t* lum;      // slab luminance, nrows x cols
    const double* Er;        // rows x nR   (row-major)
    const double* Ec;        // cols x nC   (row-major)
    const double* EcT;       // nC x cols
    const double* Gt;        // 256
    const uint8_t* Ysel;     // p   luminance of sample i (raster order: i = a*nC + b)
    const int* rowa;         // rows : a if the image row is a sample row else -1
    const int* colb;         // cols : b or -1
};

void launch_sample_indices(int rows, int cols, const int* rowa, const int* colb,
                           const int* rowrank, const int* colrank, int nC,
                           int32_t* selected, int32_t* rest, cudaStream_t s);
void launch_tables(int rows, int cols, int nR, int nC, const int* sel_rows, const int* sel_cols,
                   double hx, double hy, double* Er, double* Ec, double* EcT, double* Gt,
                   cudaStream_t s);
void launch_ka(int p, int nC, const int* sel_rows, const int* sel_cols, const uint8_t* Ysel,
               double hx, double hy, double* Ka, cudaStream_t s);

// Cell index of a slab (cell_kernels.cu): the non-empty (image row, luminance level) cells, padded to a multiple of 4
// per row, and the columns of every row ordered by (level, column).  All device pointers into `scratch`.
struct CellIndex {
    const int* koff;         // nrows+1: first cell of each row; koff[nrows] = number of (padded) cells
    const uint8_t* lev;      // level of each cell (0 for padding cells)
    const int* row;          // local image row of each cell
    const int* pstart;       // first position of the cell in `sorted`
    const int* pcount;       // pixels in the cell (0 for padding cells)
    const int* sorted;       // nrows*cols: column indices, row by row, ordered by (level, column)
    int cap_cells;           // upper bound of the number of cells (host-side grid sizing)
};
size_t cell_index_scratch_doubles(const AffinityTables& t);
CellIndex build_cell_index(const AffinityTables& t, double* scratch, cudaStream_t s);

// One Sinkhorn half-iteration with the sample-axis contractions as level-table GEMMs on the FP64 tensor pipe
// (sinkhorn_cells.cu):  x = recip(k_j^T w) on the rest pixels (w == nullptr: x = 1), then s_out = Kab x.
bool sinkhorn_cells_supported(const AffinityTables& t);
size_t sinkhorn_cells_scratch_doubles(const AffinityTables& t);
void sinkhorn_cells_prepare(const AffinityTables& t, double* scratch, cudaStream_t s);
void launch_sinkhorn_cells(const AffinityTables& t, const CellIndex* ci, const double* w, double* x, double* scratch,
                           double* s_out, cudaStream_t s);

// Weighted Gram  G = sum_j c_j^2 k_j k_j^T  (p x p, column-major, both triangles) over the slab, contracted over the
// (image row, luminance level) cells (cell_kernels.cu): K_cells*p*(p+1) flop instead of N*p*(p+1).
size_t gram_cells_scratch_doubles(const AffinityTables& t);
// ci (optional): the slab's cell index (build_cell_index); with it the per-cell histograms Hh are built on the tensor pipe from
// that index instead of sorting every image row again.
void launch_gram_cells(const AffinityTables& t, const double* c, double* scratch, double* G, cudaStream_t s,
                       const CellIndex* ci = nullptr);

// Extension  V_j = c_j * k_j^T Y  for the non-sample slab pixels.  Y: p x k (column-major), V: (nrows*cols) x k ROW-major
// (k fastest).  Through the (image row, luminance level) cells (cell_kernels.cu): K_cells*p*k + N*nC*k multiply-adds.
size_t extension_cells_scratch_doubles(const AffinityTables& t, int k);
// ci (optional): the slab's cell index; used instead of a second per-row sort when FX fits in one batch of rows.
void launch_extension_cells(const AffinityTables& t, const double* c, const double* Y, int k, double* scratch, double* V,
                            cudaStream_t s, const CellIndex* ci = nullptr);
// V[pixel(sel[i0+i]) - slab offset][:] = src(i, :)  for samples that fall in the slab.  src: n x k col-major (ld).
void launch_scatter_rows(const AffinityTables& t, const int32_t* sel, int i0, int n, const double* src,
                         int ld, int k, double* V, cudaStream_t s);

// The HBM-bound output path (apply_kernels.cu): both passes stream V (nloc x k, row-major) through shared memory with TMA
// bulk copies; the colour conversions of NLEFilter::enhance are fused into them.
//   launch_vtz        t = V^T z (and g = fS o t when fS != nullptr); exactly one of z_u8 / z_f64 / z_bgr non-null (z_bgr:
//                     interleaved 8-bit BGR, z = L of its Lab conversion).  scratch: apply_blocks(nloc, k) * k doubles.
//   launch_scale_t    g = fS o t (after the all-reduce of t in the sharded path)
//   launch_recompose  out = V g; epilogue by the non-null output: out_f64 | out_u8 (clamp, round half to even) | out_bgr
//                     (clamp, round, Lab2BGR with the a, b channels recomputed from bgr_in)
bool apply_tma_supported(int k);
int apply_blocks(long long nloc, int k);
void launch_vtz(long long nloc, int k, const double* V, const uint8_t* z_u8, const double* z_f64, const uint8_t* z_bgr,
                const double* fS, double* scratch, double* t_out, double* g_out, cudaStream_t s);
void launch_scale_t(int k, const double* t, const double* fS, double* g, cudaStream_t s);
void launch_recompose(long long nloc, int k, const double* V, const double* g, double* out_f64, uint8_t* out_u8,
                      uint8_t* out_bgr, const uint8_t* bgr_in, cudaStream_t s);

// 8-bit BGR <-> Lab, byte-exact with cv::cvtColor on CV_8UC3 (apply_kernels.cu).  bgr: npix x 3 interleaved; L: npix; ab: npix x 2.
void launch_bgr2lab(const uint8_t* bgr, long long npix, uint8_t* L, uint8_t* ab, cudaStream_t s);
void launch_lab2bgr(const uint8_t* L, const uint8_t* ab, long long npix, uint8_t* bgr, cudaStream_t s);

// misc elementwise
void launch_fill(double* p, long long n, double v, cudaStream_t s);
void launch_u8_from_f64(const double* in, long long n, uint8_t* out, int* bad_flag, cudaStream_t s);
void launch_gather_c_sel(const AffinityTables& t, const int32_t* sel, const double* c_sel, double* c_full,
                         cudaStream_t s);

}  // namespace nle
