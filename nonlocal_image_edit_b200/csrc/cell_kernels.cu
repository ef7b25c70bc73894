// Weighted Gram  G = sum_j c_j^2 k_j k_j^T  over the rest pixels (the "Wab Wab^T" of filter.cpp:296 in
// factor form), contracted through (image row, luminance level) CELLS.
//
// With K(i,j) = Er[row][a] * Ec[col][b] * Gt[|l - Y_ab|]  (kernels.cuh) the Gram separates as
//
//   G[(a,b),(a',b')] = sum_{row,l}  (Er[row][a] Gt[|l-Y_ab|]) * Hh[row,l][b,b'] * (Er[row][a'] Gt[|l-Y_a'b'|])
//   Hh[row,l][b,b']  = sum_{col : lum(row,col) = l}  c_j^2 Ec[col][b] Ec[col][b']
//
// i.e. the pixel axis (N terms) collapses onto the non-empty cells (row, l) -- at most 256 per image row,
// K_cells of them in total -- and for every pair of grid columns (b <= b') the nR x nR block of G is a small
// GEMM whose K dimension is the cell axis.  Work drops from N*p*(p+1) flop (the pixel-axis SYRK of
// gram_kernel, filter_kernels.cu) to K_cells*p*(p+1): a factor W/nlev (5.9 on the 1024^2 bench image, >= 16
// on a 4096-wide one), and nothing about it is approximate -- it is a re-association of the same FP64 sum.
//
// Kernels:
//   cell_count_kernel  per image row: number of distinct levels, padded to a multiple of 4
//   cell_scan_kernel   exclusive scan -> koff[row] (first cell of each row), koff[nrows] = K_pad
//   cell_hist_kernel   per image row: level list, stable counting sort of the columns by level, Hh for every
//                      pair (b <= b') accumulated in ascending column order (deterministic), cell levels
//   gram_cells_kernel  one WARP per (pair, a-block, a'-block, K split): 8*MT x 8*MT FP64 accumulators on the
//                      FP64 tensor pipe (DMMA m8n8k4); operand fragments are generated straight into
//                      registers (one table look-up and one or two multiplies per element), no shared-memory
//                      tiles, no block barriers in the main loop
//   gram_cells_reduce_kernel  fixed-order sum over the K splits, mirrored into both triangles of G
//
// The eigenvector extension (filter.cpp:324-327 in factor form) uses the same cells, see the second half of
// this file:  V_j = c_j sum_b Ec[col_j][b] FX[cell(j)][b][:],   FX[cell][b][:] = sum_a Er[row][a] Gt[|l-Y_ab|] Y[(a,b),:].
#include <algorithm>
#include <cstdlib>

#include "kernels.cuh"

namespace nle {

namespace {

// Distinct luminance levels of one image row (blockDim.x == 256).  levidx[l] = compact index (ascending
// level) or -1, lev[li] = level.  Returns the number of levels.
__device__ __forceinline__ int row_levels(const uint8_t* __restrict__ Lrow, int W, int* flags, int* levidx,
                                          int* lev, int* wcount) {
    const int tid = threadIdx.x;
    flags[tid] = 0;
    __syncthreads();
    for (int c = tid; c < W; c += 256) flags[Lrow[c]] = 1;
    __syncthreads();
    const unsigned m = __ballot_sync(0xffffffffu, flags[tid] != 0);
    if ((tid & 31) == 0) wcount[tid >> 5] = __popc(m);
    __syncthreads();
    int base = 0, total = 0;
    for (int w = 0; w < 8; ++w) {
        if (w < (tid >> 5)) base += wcount[w];
        total += wcount[w];
    }
    const int my = base + __popc(m & ((1u << (tid & 31)) - 1u));
    if (flags[tid]) { levidx[tid] = my; lev[my] = tid; } else levidx[tid] = -1;
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(256)
cell_count_kernel(const uint8_t* __restrict__ lum, int nrows, int W, int* __restrict__ cnt) {
    __shared__ int flags[256];
    __shared__ int wcount[8];
    const int tid = threadIdx.x;
    for (int rl = blockIdx.x; rl < nrows; rl += gridDim.x) {
        __syncthreads();
        flags[tid] = 0;
        __syncthreads();
        const uint8_t* L = lum + (size_t)rl * W;
        for (int c = tid; c < W; c += 256) flags[L[c]] = 1;
        __syncthreads();
        const unsigned m = __ballot_sync(0xffffffffu, flags[tid] != 0);
        if ((tid & 31) == 0) wcount[tid >> 5] = __popc(m);
        __syncthreads();
        if (tid == 0) {
            int total = 0;
            for (int w = 0; w < 8; ++w) total += wcount[w];
            cnt[rl] = (total + 3) & ~3;
        }
    }
}

// koff[0] = 0, koff[i+1] = koff[i] + cnt[i]   (single CTA of 1024 threads; nrows is a few thousand)
__global__ void __launch_bounds__(1024)
cell_scan_kernel(const int* __restrict__ cnt, int nrows, int* __restrict__ koff) {
    __shared__ int part[1024];
    const int tid = threadIdx.x;
    const int per = (nrows + 1023) / 1024;
    const int r0 = min(nrows, tid * per), r1 = min(nrows, r0 + per);
    int s = 0;
    for (int r = r0; r < r1; ++r) s += cnt[r];
    part[tid] = s;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int i = 0; i < 1024; ++i) { const int v = part[i]; part[i] = run; run += v; }
        koff[nrows] = run;
    }
    __syncthreads();
    int run = part[tid];
    for (int r = r0; r < r1; ++r) { koff[r] = run; run += cnt[r]; }
}

// linear index of the pair (b <= b') in the upper triangle of an nC x nC matrix, row by row
__host__ __device__ __forceinline__ int pair_start(int b, int nC) { return b * nC - (b * (b - 1)) / 2; }
__device__ __forceinline__ void pair_decode(int q, int nC, int& b, int& bp) {
    int lo = 0;
    while (lo + 1 < nC && pair_start(lo + 1, nC) <= q) ++lo;
    b = lo;
    bp = lo + (q - pair_start(lo, nC));
}

constexpr int HS = 32;      // sorted pixels staged per step
constexpr int HPT = 4;      // pairs per thread and batch

// One CTA per image row.  Hh[(koff[row] + li) * ld + q] for every pair q; cell_lev[koff[row] + li].
__global__ void __launch_bounds__(256)
cell_hist_kernel(AffinityTables t, const double* __restrict__ cvec, const int* __restrict__ koff, int npairs,
                 int ld, uint8_t* __restrict__ cell_lev, double* __restrict__ Hh) {
    extern __shared__ double hsm[];
    const int nC = t.nC, W = t.cols;
    double* stA = hsm;                          // HS * nC   c_j^2 * Ec[col][.]
    double* stB = stA + HS * nC;                // HS * nC   Ec[col][.]
    int* flags = reinterpret_cast<int*>(stB + HS * nC);
    int* levidx = flags + 256;
    int* lev = levidx + 256;
    int* cstart = lev + 256;                    // 257: first sorted position of each cell
    int* wcount = cstart + 260;
    int* sorted = wcount + 8;                   // W: columns ordered by (level, column)
    uint8_t* Lrow = reinterpret_cast<uint8_t*>(sorted + W);           // W
    uint8_t* lis = Lrow + ((W + 15) / 16) * 16;                       // W: cell index of each sorted position
    const int tid = threadIdx.x;
    for (int rl = blockIdx.x; rl < t.nrows; rl += gridDim.x) {
        __syncthreads();
        const uint8_t* Lg = t.lum + (size_t)rl * W;
        for (int c = tid; c < W; c += 256) Lrow[c] = Lg[c];
        __syncthreads();
        const int nlev = row_levels(Lrow, W, flags, levidx, lev, wcount);
        const int npad = (nlev + 3) & ~3;
        const int k0 = koff[rl];
        // stable counting sort of the columns by level: thread li scans the row (ascending column)
        if (tid < nlev) {
            const int lv = lev[tid];
            int n = 0;
            for (int c = 0; c < W; ++c) n += (Lrow[c] == lv);
            flags[tid] = n;                       // flags is free again: per-cell pixel count
        }
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            for (int li = 0; li < nlev; ++li) { cstart[li] = run; run += flags[li]; }
            cstart[nlev] = run;
        }
        __syncthreads();
        if (tid < nlev) {
            const int lv = lev[tid];
            int pos = cstart[tid];
            for (int c = 0; c < W; ++c)
                if (Lrow[c] == lv) { sorted[pos] = c; lis[pos] = (uint8_t)tid; ++pos; }
        }
        if (tid < npad) cell_lev[k0 + tid] = (uint8_t)(tid < nlev ? lev[tid] : 0);
        __syncthreads();
        const double* cg = cvec + (size_t)rl * W;
        for (int q0 = 0; q0 < npairs; q0 += 256 * HPT) {
            int qb[HPT], qbp[HPT];
            double acc[HPT];
#pragma unroll
            for (int i = 0; i < HPT; ++i) {
                const int q = q0 + tid + 256 * i;
                if (q < npairs) pair_decode(q, nC, qb[i], qbp[i]); else { qb[i] = 0; qbp[i] = 0; }
                acc[i] = 0.0;
            }
            int cur = 0;
            auto flush = [&](int li) {
                double* h = Hh + (size_t)(k0 + li) * ld;
#pragma unroll
                for (int i = 0; i < HPT; ++i) {
                    const int q = q0 + tid + 256 * i;
                    if (q < npairs) h[q] = acc[i];
                    acc[i] = 0.0;
                }
            };
            for (int s0 = 0; s0 < W; s0 += HS) {
                __syncthreads();
                const int ns = min(HS, W - s0);
                for (int e = tid; e < ns * nC; e += 256) {
                    const int sI = e / nC, b = e - sI * nC;
                    const int col = sorted[s0 + sI];
                    const double cj = cg[col];
                    const double ev = t.Ec[(size_t)col * nC + b];
                    stA[e] = (cj * cj) * ev;
                    stB[e] = ev;
                }
                __syncthreads();
                for (int sI = 0; sI < ns; ++sI) {
                    const int li = (int)lis[s0 + sI];
                    if (li != cur) { flush(cur); cur = li; }
                    const double* sa = stA + sI * nC;
                    const double* sb = stB + sI * nC;
#pragma unroll
                    for (int i = 0; i < HPT; ++i) acc[i] = fma(sa[qb[i]], sb[qbp[i]], acc[i]);
                }
            }
            flush(cur);
            // padding cells carry zero weight
            for (int li = nlev; li < npad; ++li) {
                double* h = Hh + (size_t)(k0 + li) * ld;
#pragma unroll
                for (int i = 0; i < HPT; ++i) {
                    const int q = q0 + tid + 256 * i;
                    if (q < npairs) h[q] = 0.0;
                }
            }
        }
    }
}


// Hh[cell][q(b, b')] = sum_{col in cell} (c_j^2 Ec[col][b]) * Ec[col][b'],  b <= b', on the FP64 tensor pipe: one warp per cell of
// an existing cell index (the one the Sinkhorn passes use), the cell's pixels four at a time as the K dimension, the sample
// column on both the M and the N axis -- the same Ec values serve as A (scaled by c_j^2) and as B fragment, loaded once as
// 64-byte runs; only the tiles on and above the diagonal are issued.  (cell_hist_kernel above does the same sums with two
// shared-memory look-ups per multiply-add and is bound by shared-memory bandwidth: 0.75 ms per 1024 x 1024 image against 0.86 G
// multiply-adds; it remains for slabs whose Hh table has to be built in batches of rows.)  Pixels are taken in ascending
// column order, as there.
template <int NB>
__global__ void __launch_bounds__(256)
cell_hist_dmma_kernel(AffinityTables t, CellIndex ci, const double* __restrict__ cvec, int ld, double* __restrict__ Hh) {
    const int nC = t.nC, W = t.cols;
    const int lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const int K = ci.koff[t.nrows];
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    for (int cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); cell < K; cell += nwarps) {
        const int np = ci.pcount[cell], rl = ci.row[cell];
        const int* pix = ci.sorted + ci.pstart[cell];
        const double* cg = cvec + (size_t)rl * W;
        double acc[NB][NB][2];
#pragma unroll
        for (int u = 0; u < NB; ++u)
#pragma unroll
            for (int v = u; v < NB; ++v) acc[u][v][0] = acc[u][v][1] = 0.0;
        for (int p0 = 0; p0 < np; p0 += 4) {
            const bool ok = p0 + tq < np;
            const int col = ok ? pix[p0 + tq] : 0;
            const double cj = ok ? cg[col] : 0.0;
            const double c2 = cj * cj;
            const double* ecr = t.Ec + (size_t)col * nC + g;
            double ev[NB], av[NB];
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                ev[u] = (ok && 8 * u + g < nC) ? ecr[8 * u] : 0.0;
                av[u] = c2 * ev[u];
            }
#pragma unroll
            for (int u = 0; u < NB; ++u)
#pragma unroll
                for (int v = u; v < NB; ++v) dmma884(acc[u][v][0], acc[u][v][1], av[u], ev[v]);
        }
        double* h = Hh + (size_t)cell * ld;
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int b = 8 * u + g;
            if (b >= nC) continue;
            double* hb = h + pair_start(b, nC) - b;
#pragma unroll
            for (int v = u; v < NB; ++v)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int bp = 8 * v + 2 * tq + e;
                    if (bp >= b && bp < nC) hb[bp] = acc[u][v][e];
                }
        }
    }
}

// One warp per task = (pair (b,b'), a-block, a'-block) x K split.  Tile: T x T with T = 8*MT grid rows.
//   A[k][a ] = Er[row_k][a ] * Gt[|lev_k - Y[a ][b ]|]
//   B[k][a'] = Er[row_k][a'] * Gt[|lev_k - Y[a'][b']|] * Hh[k][pair]
// DMMA fragments (lane = 4g + tq): A elem (m = g, k = tq), B elem (k = tq, n = g), D elems (g, 2tq), (g, 2tq+1).
// The 256-entry Gt table is replicated 16x in shared memory, interleaved so that lane L always reads bank pair
// (L & 15): the 64-bit look-ups of a warp (two 16-lane phases) are conflict-free whatever the levels are.
// WARPS per CTA: 12 (three warps per sub-core) whenever the tile fits 168 registers without spilling -- up to 5 x 5
// (ptxas: 168 registers, 0 spills); the 6 x 4 / 7 x 4 tiles of nR > 40 keep 8 warps (224 registers).
template <int MT, int NT, bool SKIP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
gram_cells_kernel(AffinityTables t, const int* __restrict__ koff, const uint8_t* __restrict__ cell_lev,
                  const double* __restrict__ Hh, int ld, int nabA, int nabB, int ntasks, int nsplit, int accumulate,
                  double* __restrict__ part) {
    __shared__ double Gs16[256 * 16];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < 256 * 16; e += WARPS * 32) Gs16[e] = t.Gt[e >> 4];
    __syncthreads();
    const int task = blockIdx.x * WARPS + warp;
    if (task >= ntasks) return;
    const int nC = t.nC, nR = t.nR;
    constexpr int TA = 8 * MT, TB = 8 * NT;
    const int pair = task / (nabA * nabB);
    const int rem = task - pair * nabA * nabB;
    const int ab = rem / nabB, abp = rem - ab * nabB;
    int b, bp;
    pair_decode(pair, nC, b, bp);
    const int g = lane >> 2, tq = lane & 3;
    const int split = blockIdx.y;

    int yA[MT], yB[NT], aA[MT], aB[NT];
    bool okA[MT], okB[NT];          // SKIP: DMMA tiles that lie entirely in the padding of a blocked shape are not issued
#pragma unroll
    for (int u = 0; u < MT; ++u) {
        aA[u] = ab * TA + 8 * u + g;
        okA[u] = ab * TA + 8 * u < nR;
        yA[u] = aA[u] < nR ? (int)t.Ysel[aA[u] * nC + b] : 0;
    }
#pragma unroll
    for (int v = 0; v < NT; ++v) {
        aB[v] = abp * TB + 8 * v + g;
        okB[v] = abp * TB + 8 * v < nR;
        yB[v] = aB[v] < nR ? (int)t.Ysel[aB[v] * nC + bp] : 0;
    }
    double acc[MT][NT][2];
#pragma unroll
    for (int u = 0; u < MT; ++u)
#pragma unroll
        for (int v = 0; v < NT; ++v) acc[u][v][0] = acc[u][v][1] = 0.0;

    const int nsteps = koff[t.nrows] >> 2;
    int k = (int)(((long long)nsteps * split) / nsplit) * 4;
    const int kend = (int)(((long long)nsteps * (split + 1)) / nsplit) * 4;
    // image row of the first cell
    int row = 0;
    {
        int lo = 0, hi = t.nrows - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (koff[mid] <= k) lo = mid; else hi = mid - 1;
        }
        row = lo;
    }
    int krow_end = koff[row + 1];
    double erA[MT], erB[NT];
    auto load_er = [&]() {
        const double* er = t.Er + (size_t)(t.row0 + row) * nR;
#pragma unroll
        for (int u = 0; u < MT; ++u) erA[u] = aA[u] < nR ? er[aA[u]] : 0.0;
#pragma unroll
        for (int v = 0; v < NT; ++v) erB[v] = aB[v] < nR ? er[aB[v]] : 0.0;
    };
    load_er();
    const double* gs = Gs16 + (lane & 15);
    const double* hp = Hh + pair;

    // levels and Hh values of 32 cells per lane-load, fetched one block ahead of their use (the HBM round trip of
    // the Hh stream is ~3 blocks of DMMA work for the two warps of a sub-core)
    int lv_next = 0;
    double hh_next = 0.0;
    if (k + lane < kend) {
        lv_next = (int)cell_lev[k + lane];
        hh_next = hp[(size_t)(k + lane) * ld];
    }
    while (k < kend) {
        const int lv32 = lv_next;
        const double hh32 = hh_next;
        {
            const int kn = k + 32 + lane;
            lv_next = 0;
            hh_next = 0.0;
            if (kn < kend) {
                lv_next = (int)cell_lev[kn];
                hh_next = hp[(size_t)kn * ld];
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int ks = k + 4 * i;
            if (ks >= kend) break;
            if (ks >= krow_end) {
                do { ++row; krow_end = koff[row + 1]; } while (ks >= krow_end);
                load_er();
            }
            const int lv = __shfl_sync(0xffffffffu, lv32, 4 * i + tq);
            const double hh = __shfl_sync(0xffffffffu, hh32, 4 * i + tq);
            double a[MT], bf[NT];
#pragma unroll
            for (int u = 0; u < MT; ++u) {
                const int dA = lv - yA[u];
                a[u] = erA[u] * gs[(dA < 0 ? -dA : dA) << 4];
            }
#pragma unroll
            for (int v = 0; v < NT; ++v) {
                const int dB = lv - yB[v];
                bf[v] = (erB[v] * hh) * gs[(dB < 0 ? -dB : dB) << 4];
            }
#pragma unroll
            for (int u = 0; u < MT; ++u)
#pragma unroll
                for (int v = 0; v < NT; ++v)
                    if (!SKIP || (okA[u] && okB[v])) dmma884(acc[u][v][0], acc[u][v][1], a[u], bf[v]);
        }
        k += 32;
    }

    double* out = part + ((size_t)split * ntasks + task) * (TA * TB);
#pragma unroll
    for (int u = 0; u < MT; ++u)
#pragma unroll
        for (int v = 0; v < NT; ++v) {
            double2* o = reinterpret_cast<double2*>(out + (8 * u + g) * TB + 8 * v + 2 * tq);
            double2 val = make_double2(acc[u][v][0], acc[u][v][1]);
            if (accumulate) { const double2 old = *o; val.x += old.x; val.y += old.y; }
            *o = val;
        }
}

__global__ void gram_cells_reduce_kernel(const double* __restrict__ part, int p, int nR, int nC, int TA, int TB, int nabA,
                                         int nabB, int ntasks, int nsplit, double* __restrict__ G) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= p) return;
    int a = i / nC, b = i - a * nC, ap = j / nC, bp = j - ap * nC;
    if (b > bp || (b == bp && a > ap)) { int tmp = a; a = ap; ap = tmp; tmp = b; b = bp; bp = tmp; }
    const int pair = pair_start(b, nC) + (bp - b);
    const int task = pair * nabA * nabB + (a / TA) * nabB + (ap / TB);
    const size_t off = (size_t)task * (TA * TB) + (size_t)(a % TA) * TB + (ap % TB);
    double acc = 0.0;
    for (int s = 0; s < nsplit; ++s) acc += part[(size_t)s * ntasks * (TA * TB) + off];
    G[i + (size_t)j * p] = acc;
}

struct CellGeom {
    int npairs, ld, MT, NT, TA, TB, nabA, nabB, ntasks, warps, ncta, nsplit, capc, rows_batch;
    size_t hh_doubles, part_doubles, lev_bytes, int_count;
};

CellGeom cell_geometry(const AffinityTables& t) {
    CellGeom g;
    g.npairs = t.nC * (t.nC + 1) / 2;
    g.ld = (g.npairs + 3) & ~3;
    // warp tile: 8*MT grid rows (A side) x 8*NT grid rows (B side).  Up to 5 DMMA tiles per side: square; 6-7 (nR = 50 at
    // p = 2500): 7 x 4 (28 tiles; 7 x 7 would need 196 accumulator registers); beyond: blocks of at most 7 x 4.
    const int tiles = cdiv(t.nR, 8);
    if (tiles <= 5) {
        g.MT = g.NT = tiles; g.nabA = g.nabB = 1;
    } else {
        g.nabA = cdiv(tiles, 7); g.MT = cdiv(tiles, g.nabA);
        g.NT = 4; g.nabB = cdiv(tiles, 4);
    }
    g.TA = 8 * g.MT; g.TB = 8 * g.NT;
    g.ntasks = g.npairs * g.nabA * g.nabB;
    g.warps = 12;
    g.ncta = cdiv(g.ntasks, g.warps);
    g.nsplit = std::max(1, (7 * sm_count()) / g.ncta);
    g.nsplit = std::min(g.nsplit, std::max(1, t.nrows));
    g.capc = std::min(256, (t.cols + 3) & ~3);
    // Hh is sized for the worst case (every row holds capc cells): bound it to ~2.5 GB per batch of rows
    const size_t per_row = (size_t)g.capc * g.ld;
    const size_t budget = (size_t)320 << 20;   // doubles
    g.rows_batch = (int)std::max<size_t>(1, std::min<size_t>((size_t)t.nrows, budget / per_row));
    g.hh_doubles = per_row * g.rows_batch;
    g.part_doubles = (size_t)g.nsplit * g.ntasks * g.TA * g.TB;
    g.lev_bytes = (size_t)g.capc * g.rows_batch;
    g.int_count = 2 * (size_t)g.rows_batch + 8;
    return g;
}

// =============================================================================================
// Eigenvector extension through cells.
//
//   ext_index_kernel   per image row: level list, stable counting sort of the columns by level; writes for every
//                      cell its level, image row, first position in the row's sorted column list and pixel count
//   ext_fx_kernel      FX[cell][b][m] = sum_a (Er[row][a] Gt[|l-Y_ab|]) * Y[(a,b)][m]      DMMA, K = nR
//                      (cells x nR) * (nR x 56) per grid column b and 56-column block of Y
//   ext_pix_kernel     one warp per (cell, column block): V[pixel][m] = c_j * sum_b Ec[col_j][b] FX[cell][b][m]
//                      DMMA with the cell's pixels (8 at a time) as the M tile, K = nC; FX is streamed once
// Work: K_cells*p*k' + N*nC*k' multiply-adds instead of N*p*k' (extension_dmma_kernel).
constexpr int XC_N = 56;          // eigenvector columns per block (7 DMMA n-tiles)
constexpr int XC_WARPS = 12;       // three warps per sub-core: the 4 x 7 tile fits 168 registers
constexpr int XC_THREADS = XC_WARPS * 32;
constexpr int XC_CELLS = XC_WARPS * 32;   // cells per sub-tile of ext_fx_kernel (XC_WARPS warps x 4 m-tiles)
constexpr int XC_SUB = 16;        // sub-tiles per CTA (the Y slice and the Er rows are staged once for all of them)
constexpr int XC_ERROWS = 256;    // image rows whose Er rows are staged at most (fewer when the sample grid leaves less shared memory)

__global__ void __launch_bounds__(256)
ext_index_kernel(const uint8_t* __restrict__ lum, int nrows, int W, const int* __restrict__ koff,
                 uint8_t* __restrict__ cell_lev, int* __restrict__ cell_row, int* __restrict__ cell_pstart,
                 int* __restrict__ cell_pcount, int* __restrict__ sorted) {
    extern __shared__ int ism[];
    int* flags = ism;              // 256
    int* levidx = flags + 256;     // 256
    int* lev = levidx + 256;       // 256
    int* cstart = lev + 256;       // 260
    int* wcount = cstart + 260;    // 8
    uint8_t* Lrow = reinterpret_cast<uint8_t*>(wcount + 8);   // W
    const int tid = threadIdx.x;
    for (int rl = blockIdx.x; rl < nrows; rl += gridDim.x) {
        __syncthreads();
        const uint8_t* Lg = lum + (size_t)rl * W;
        for (int c = tid; c < W; c += 256) Lrow[c] = Lg[c];
        __syncthreads();
        const int nlev = row_levels(Lrow, W, flags, levidx, lev, wcount);
        const int npad = (nlev + 3) & ~3;
        const int k0 = koff[rl];
        if (tid < nlev) {
            const int lv = lev[tid];
            int n = 0;
            for (int c = 0; c < W; ++c) n += (Lrow[c] == lv);
            flags[tid] = n;
        }
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            for (int li = 0; li < nlev; ++li) { cstart[li] = run; run += flags[li]; }
            cstart[nlev] = run;
        }
        __syncthreads();
        if (tid < nlev) {
            const int lv = lev[tid];
            int pos = cstart[tid];
            int* so = sorted + (size_t)rl * W;
            for (int c = 0; c < W; ++c)
                if (Lrow[c] == lv) so[pos++] = c;
        }
        if (tid < npad) {
            const bool real = tid < nlev;
            cell_lev[k0 + tid] = (uint8_t)(real ? lev[tid] : 0);
            cell_row[k0 + tid] = rl;
            cell_pstart[k0 + tid] = rl * W + (real ? cstart[tid] : 0);
            cell_pcount[k0 + tid] = real ? flags[tid] : 0;
        }
    }
}

// grid (ceil(cap_cells / (XC_CELLS * XC_SUB)), nC, column blocks); warp = 32 cells (4 m-tiles) x 56 columns (7 n-tiles).
__global__ void __launch_bounds__(XC_THREADS, 1)
ext_fx_kernel(AffinityTables t, const int* __restrict__ koff, const uint8_t* __restrict__ cell_lev,
              const int* __restrict__ cell_row, const double* __restrict__ Yt, int kp, double* __restrict__ FX, int erows) {
    extern __shared__ double xsm[];
    const int nR = t.nR, nC = t.nC;
    const int nR4 = (nR + 3) & ~3;
    double* Ys = xsm;                                  // nR4 * XC_N   slice of Y for grid column b
    double* Gs = Ys + (size_t)nR4 * XC_N;              // 256
    int* ysl = reinterpret_cast<int*>(Gs + 256);       // nR4          sample luminances of grid column b
    double* ErS = reinterpret_cast<double*>(ysl + nR4 + (nR4 & 1));   // erows * nR4: Er rows of the image rows this CTA's cells span
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tq = lane & 3;
    const int b = blockIdx.y, vb = blockIdx.z;
    const int K = koff[t.nrows];
    const int k0 = blockIdx.x * (XC_CELLS * XC_SUB);
    if (k0 >= K) return;
    if (tid < 256) Gs[tid] = t.Gt[tid];
    for (int e = tid; e < nR4 * XC_N; e += XC_THREADS) {
        const int a = e / XC_N, m = e - a * XC_N;
        Ys[e] = a < nR ? Yt[(size_t)(a * nC + b) * kp + vb * XC_N + m] : 0.0;
    }
    for (int a = tid; a < nR4; a += XC_THREADS) ysl[a] = a < nR ? (int)t.Ysel[a * nC + b] : 0;
    // cells are ordered by image row: the CTA's cells span rows [row_first, row_last]; their Er rows are staged in shared
    // memory when there are at most `erows` of them (ncu: 33 % of the samples sat on the Er look-up in global memory)
    const int row_first = cell_row[k0];
    const int row_last = cell_row[min(K, k0 + XC_CELLS * XC_SUB) - 1];
    const bool er_smem = row_last - row_first < erows;
    if (er_smem)
        for (int e = tid; e < (row_last - row_first + 1) * nR4; e += XC_THREADS) {
            const int r = e / nR4, a = e - r * nR4;
            ErS[e] = a < nR ? t.Er[(size_t)(t.row0 + row_first + r) * nR + a] : 0.0;
        }
    __syncthreads();
    for (int sub = 0; sub < XC_SUB; ++sub) {
        const int kw = k0 + sub * XC_CELLS + warp * 32;
        if (kw >= K) break;
        int crow[4], clev[4];
        bool cok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int cell = kw + 8 * u + g;
            cok[u] = cell < K;
            crow[u] = cok[u] ? cell_row[cell] : row_first;
            clev[u] = cok[u] ? (int)cell_lev[cell] : 0;
        }
        double acc[4][7][2];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 7; ++v) acc[u][v][0] = acc[u][v][1] = 0.0;
        for (int kk = 0; kk < nR4; kk += 4) {
            const int a = kk + tq;
            const int ya = ysl[a];
            double af[4], bf[7];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int d = clev[u] - ya;
                double er = 0.0;
                if (cok[u] && a < nR) er = er_smem ? ErS[(crow[u] - row_first) * nR4 + a] : t.Er[(size_t)(t.row0 + crow[u]) * nR + a];
                af[u] = er * Gs[d < 0 ? -d : d];
            }
#pragma unroll
            for (int v = 0; v < 7; ++v) bf[v] = Ys[a * XC_N + 8 * v + g];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 7; ++v) dmma884(acc[u][v][0], acc[u][v][1], af[u], bf[v]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int cell = kw + 8 * u + g;
            if (cell >= K) continue;
            double* o = FX + ((size_t)cell * nC + b) * kp + vb * XC_N + 2 * tq;
#pragma unroll
            for (int v = 0; v < 7; ++v) *reinterpret_cast<double2*>(o + 8 * v) = make_double2(acc[u][v][0], acc[u][v][1]);
        }
    }
}

// One warp per (cell, column block), grid-stride.
__global__ void __launch_bounds__(256)
ext_pix_kernel(AffinityTables t, const int* __restrict__ koff, const int* __restrict__ cell_row,
               const int* __restrict__ cell_pstart, const int* __restrict__ cell_pcount,
               const int* __restrict__ sorted, const double* __restrict__ cvec, const double* __restrict__ FX,
               int kp, int nvb, int k, double* __restrict__ V) {
    const int nC = t.nC, W = t.cols;
    const int nC4 = (nC + 3) & ~3;
    const int lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const long long K = koff[t.nrows];
    const long long nwork = K * nvb;
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long wi = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); wi < nwork; wi += wstride) {
        const int cell = (int)(wi / nvb), vb = (int)(wi - (long long)cell * nvb);
        const int np = cell_pcount[cell];
        if (np == 0) continue;
        const int rl = cell_row[cell];
        const int* pix = sorted + cell_pstart[cell];
        const double* fx = FX + (size_t)cell * nC * kp + vb * XC_N + g;
        for (int p0 = 0; p0 < np; p0 += 8) {
            const bool pok = p0 + g < np;
            const int col = pok ? pix[p0 + g] : 0;
            const double* ec = t.Ec + (size_t)col * nC;
            double acc[7][2];
#pragma unroll
            for (int v = 0; v < 7; ++v) acc[v][0] = acc[v][1] = 0.0;
            for (int kk = 0; kk < nC4; kk += 4) {
                const int b = kk + tq;
                const bool bok = b < nC;
                const double af = (pok && bok) ? ec[b] : 0.0;
                double bf[7];
                const double* f = fx + (size_t)b * kp;
#pragma unroll
                for (int v = 0; v < 7; ++v) bf[v] = bok ? f[8 * v] : 0.0;
#pragma unroll
                for (int v = 0; v < 7; ++v) dmma884(acc[v][0], acc[v][1], af, bf[v]);
            }
            if (!pok) continue;
            if (t.rowa[t.row0 + rl] >= 0 && t.colb[col] >= 0) continue;      // sample pixel: scattered separately
            const size_t j = (size_t)rl * W + col;
            const double cj = cvec[j];
            double* vo = V + j * k;
#pragma unroll
            for (int v = 0; v < 7; ++v) {
                const int m = vb * XC_N + 8 * v + 2 * tq;
                if (m < k) vo[m] = cj * acc[v][0];
                if (m + 1 < k) vo[m + 1] = cj * acc[v][1];
            }
        }
    }
}

__global__ void ext_pack_yt_kernel(const double* __restrict__ Y, int p, int k, int kp, double* __restrict__ Yt) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)p * kp) return;
    const int i = (int)(e / kp), v = (int)(e - (long long)i * kp);
    Yt[e] = (v < k) ? Y[i + (size_t)v * p] : 0.0;
}

struct ExtGeom {
    int nvb, kp, capc, rows_batch;
    size_t fx_doubles, yt_doubles, int_count, lev_bytes;
};

ExtGeom ext_geometry(const AffinityTables& t, int k) {
    ExtGeom g;
    g.nvb = cdiv(k, XC_N);
    g.kp = g.nvb * XC_N;
    g.capc = std::min(256, (t.cols + 3) & ~3);
    const size_t per_row = (size_t)g.capc * t.nC * g.kp;                 // worst case: capc cells in every row
    const size_t budget = (size_t)1 << 30;                               // doubles (8 GB) per batch of rows
    g.rows_batch = (int)std::max<size_t>(1, std::min<size_t>((size_t)t.nrows, budget / per_row));
    g.fx_doubles = per_row * g.rows_batch;
    g.yt_doubles = (size_t)t.p * g.kp;
    const size_t cells = (size_t)g.capc * g.rows_batch;
    g.int_count = 2 * (size_t)g.rows_batch + 8 + 3 * cells + (size_t)g.rows_batch * t.cols;
    g.lev_bytes = cells;
    return g;
}

}  // namespace

size_t gram_cells_scratch_doubles(const AffinityTables& t) {
    const CellGeom g = cell_geometry(t);
    return g.hh_doubles + g.part_doubles + (g.lev_bytes + 7) / 8 + (g.int_count * 4 + 7) / 8 + 8;
}

template <int MT, int NT>
static void launch_gc(const AffinityTables& tb, const CellGeom& g, const int* koff, const uint8_t* cell_lev,
                      const double* Hh, int accumulate, double* part, cudaStream_t s) {
    constexpr int WARPS = 12;      // must match cell_geometry's g.warps
    gram_cells_kernel<MT, NT, (MT != NT), WARPS><<<dim3(g.ncta, g.nsplit), WARPS * 32, 0, s>>>(tb, koff, cell_lev, Hh, g.ld, g.nabA, g.nabB,
                                                                                g.ntasks, g.nsplit, accumulate, part);
    NLE_LAUNCH_CHECK();
}

template <int NB>
static void launch_hist_dmma(const AffinityTables& t, const CellIndex& ci, const double* c, int ld, double* Hh, cudaStream_t s) {
    cell_hist_dmma_kernel<NB><<<sm_count() * 8, 256, 0, s>>>(t, ci, c, ld, Hh);
    NLE_LAUNCH_CHECK();
}

void launch_gram_cells(const AffinityTables& t, const double* c, double* scratch, double* G, cudaStream_t s, const CellIndex* ci) {
    const CellGeom g = cell_geometry(t);
    double* Hh = scratch;
    double* part = Hh + g.hh_doubles;
    if (ci != nullptr && g.rows_batch >= t.nrows && t.nC <= 64) {
        // the slab's cell index exists already and Hh fits in one batch: histograms on the tensor pipe, no second sort
        switch (cdiv(t.nC, 8)) {
            case 1: launch_hist_dmma<1>(t, *ci, c, g.ld, Hh, s); break;
            case 2: launch_hist_dmma<2>(t, *ci, c, g.ld, Hh, s); break;
            case 3: launch_hist_dmma<3>(t, *ci, c, g.ld, Hh, s); break;
            case 4: launch_hist_dmma<4>(t, *ci, c, g.ld, Hh, s); break;
            case 5: launch_hist_dmma<5>(t, *ci, c, g.ld, Hh, s); break;
            case 6: launch_hist_dmma<6>(t, *ci, c, g.ld, Hh, s); break;
            case 7: launch_hist_dmma<7>(t, *ci, c, g.ld, Hh, s); break;
            default: launch_hist_dmma<8>(t, *ci, c, g.ld, Hh, s); break;
        }
        switch (g.MT * 10 + g.NT) {
            case 11: launch_gc<1, 1>(t, g, ci->koff, ci->lev, Hh, 0, part, s); break;
            case 22: launch_gc<2, 2>(t, g, ci->koff, ci->lev, Hh, 0, part, s); break;
            case 33: launch_gc<3, 3>(t, g, ci->koff, ci->lev, Hh, 0, part, s); break;
            case 44: launch_gc<4, 4>(t, g, ci->koff, ci->lev, Hh, 0, part, s); break;
            case 55: launch_gc<5, 5>(t, g, ci->koff, ci->lev, Hh, 0, part, s); break;
            case 64: launch_gc<6, 4>(t, g, ci->koff, ci->lev, Hh, 0, part, s); break;
            case 74: launch_gc<7, 4>(t, g, ci->koff, ci->lev, Hh, 0, part, s); break;
            default: throw Unsupported{"gram: no kernel instance for the warp tile " + std::to_string(g.MT) + " x " + std::to_string(g.NT)};
        }
        gram_cells_reduce_kernel<<<dim3(cdiv(t.p, 128), t.p), 128, 0, s>>>(part, t.p, t.nR, t.nC, g.TA, g.TB, g.nabA, g.nabB,
                                                                           g.ntasks, g.nsplit, G);
        NLE_LAUNCH_CHECK();
        return;
    }
    uint8_t* cell_lev = reinterpret_cast<uint8_t*>(part + g.part_doubles);
    int* cnt = reinterpret_cast<int*>(cell_lev + ((g.lev_bytes + 7) / 8) * 8);
    int* koff = cnt + g.rows_batch + 4;
    const size_t hsm = (size_t)2 * HS * t.nC * sizeof(double) + (size_t)(256 * 3 + 260 + 8 + t.cols) * sizeof(int) +
                       2 * (size_t)((t.cols + 15) / 16) * 16 + 64;
    if (hsm > 227 * 1024) throw Unsupported{"gram: image too wide for the per-row cell sort (cols=" + std::to_string(t.cols) + ")"};
    allow_max_dynamic_smem((const void*)cell_hist_kernel);
    for (int r0 = 0, batch = 0; r0 < t.nrows; r0 += g.rows_batch, ++batch) {
        AffinityTables tb = t;
        tb.row0 = t.row0 + r0;
        tb.nrows = std::min(g.rows_batch, t.nrows - r0);
        tb.lum = t.lum + (size_t)r0 * t.cols;
        const double* cb = c + (size_t)r0 * t.cols;
        cell_count_kernel<<<std::min(tb.nrows, sm_count() * 8), 256, 0, s>>>(tb.lum, tb.nrows, tb.cols, cnt);
        NLE_LAUNCH_CHECK();
        cell_scan_kernel<<<1, 1024, 0, s>>>(cnt, tb.nrows, koff);
        NLE_LAUNCH_CHECK();
        cell_hist_kernel<<<tb.nrows, 256, hsm, s>>>(tb, cb, koff, g.npairs, g.ld, cell_lev, Hh);
        NLE_LAUNCH_CHECK();
        switch (g.MT * 10 + g.NT) {
            case 11: launch_gc<1, 1>(tb, g, koff, cell_lev, Hh, batch > 0, part, s); break;
            case 22: launch_gc<2, 2>(tb, g, koff, cell_lev, Hh, batch > 0, part, s); break;
            case 33: launch_gc<3, 3>(tb, g, koff, cell_lev, Hh, batch > 0, part, s); break;
            case 44: launch_gc<4, 4>(tb, g, koff, cell_lev, Hh, batch > 0, part, s); break;
            case 55: launch_gc<5, 5>(tb, g, koff, cell_lev, Hh, batch > 0, part, s); break;
            case 64: launch_gc<6, 4>(tb, g, koff, cell_lev, Hh, batch > 0, part, s); break;
            case 74: launch_gc<7, 4>(tb, g, koff, cell_lev, Hh, batch > 0, part, s); break;
            default: throw Unsupported{"gram: no kernel instance for the warp tile " + std::to_string(g.MT) + " x " + std::to_string(g.NT)};
        }
    }
    gram_cells_reduce_kernel<<<dim3(cdiv(t.p, 128), t.p), 128, 0, s>>>(part, t.p, t.nR, t.nC, g.TA, g.TB, g.nabA, g.nabB,
                                                                       g.ntasks, g.nsplit, G);
    NLE_LAUNCH_CHECK();
}

size_t extension_cells_scratch_doubles(const AffinityTables& t, int k) {
    const ExtGeom g = ext_geometry(t, k);
    return g.fx_doubles + g.yt_doubles + (g.lev_bytes + 7) / 8 + (g.int_count * 4 + 7) / 8 + 16;
}

// V_j = c_j k_j^T Y for the non-sample slab pixels (contract in kernels.cuh).
void launch_extension_cells(const AffinityTables& t, const double* c, const double* Y, int k, double* scratch, double* V,
                            cudaStream_t s, const CellIndex* ci) {
    if (k <= 0 || t.nrows <= 0) return;
    const ExtGeom g = ext_geometry(t, k);
    double* FX = scratch;
    double* Yt = FX + g.fx_doubles;
    uint8_t* cell_lev = reinterpret_cast<uint8_t*>(Yt + g.yt_doubles);
    int* cnt = reinterpret_cast<int*>(cell_lev + ((g.lev_bytes + 7) / 8) * 8);
    int* koff = cnt + g.rows_batch + 4;
    const size_t cells = (size_t)g.capc * g.rows_batch;
    int* cell_row = koff + g.rows_batch + 4;
    int* cell_pstart = cell_row + cells;
    int* cell_pcount = cell_pstart + cells;
    int* sorted = cell_pcount + cells;
    ext_pack_yt_kernel<<<cdiv((long long)t.p * g.kp, 256), 256, 0, s>>>(Y, t.p, k, g.kp, Yt);
    NLE_LAUNCH_CHECK();
    const size_t ism = (size_t)(256 * 3 + 260 + 8) * sizeof(int) + ((t.cols + 15) / 16) * 16 + 16;
    const int nR4 = (t.nR + 3) & ~3;
    // Er staging: a CTA's 6144 cells span 20-40 image rows on a noisy image and > 100 on a smooth one (few levels per row); stage up to
    // XC_ERROWS of them, as many as the sample grid leaves room for (beyond that the kernel reads Er from global memory)
    const size_t fbase = ((size_t)nR4 * XC_N + 256) * sizeof(double) + (size_t)(nR4 + 2) * sizeof(int) + 16;
    if (ism > 227 * 1024 || fbase + 8 * (size_t)nR4 * sizeof(double) > 227 * 1024) throw Unsupported{"extension: grid/image too large for the cell kernels"};
    int erows = (int)std::min<size_t>(XC_ERROWS, (227 * 1024 - fbase) / ((size_t)nR4 * sizeof(double)));
    const size_t fsm = fbase + (size_t)erows * nR4 * sizeof(double);
    allow_max_dynamic_smem((const void*)ext_index_kernel);
    allow_max_dynamic_smem((const void*)ext_fx_kernel);
    if (ci != nullptr && g.rows_batch >= t.nrows) {
        // one batch and the slab's cell index exists already: no second per-row sort
        const int cap_cells = g.capc * t.nrows;
        ext_fx_kernel<<<dim3(cdiv(cap_cells, XC_CELLS * XC_SUB), t.nC, g.nvb), XC_THREADS, fsm, s>>>(t, ci->koff, ci->lev, ci->row, Yt, g.kp, FX, erows);
        NLE_LAUNCH_CHECK();
        ext_pix_kernel<<<sm_count() * 8, 256, 0, s>>>(t, ci->koff, ci->row, ci->pstart, ci->pcount, ci->sorted, c, FX, g.kp, g.nvb, k, V);
        NLE_LAUNCH_CHECK();
        return;
    }
    for (int r0 = 0; r0 < t.nrows; r0 += g.rows_batch) {
        AffinityTables tb = t;
        tb.row0 = t.row0 + r0;
        tb.nrows = std::min(g.rows_batch, t.nrows - r0);
        tb.lum = t.lum + (size_t)r0 * t.cols;
        const double* cb = c + (size_t)r0 * t.cols;
        double* Vb = V + (size_t)r0 * t.cols * k;
        cell_count_kernel<<<std::min(tb.nrows, sm_count() * 8), 256, 0, s>>>(tb.lum, tb.nrows, tb.cols, cnt);
        NLE_LAUNCH_CHECK();
        cell_scan_kernel<<<1, 1024, 0, s>>>(cnt, tb.nrows, koff);
        NLE_LAUNCH_CHECK();
        ext_index_kernel<<<std::min(tb.nrows, sm_count() * 8), 256, ism, s>>>(tb.lum, tb.nrows, tb.cols, koff, cell_lev, cell_row,
                                                                             cell_pstart, cell_pcount, sorted);
        NLE_LAUNCH_CHECK();
        const int cap_cells = g.capc * tb.nrows;
        ext_fx_kernel<<<dim3(cdiv(cap_cells, XC_CELLS * XC_SUB), t.nC, g.nvb), XC_THREADS, fsm, s>>>(tb, koff, cell_lev, cell_row, Yt, g.kp, FX, erows);
        NLE_LAUNCH_CHECK();
        ext_pix_kernel<<<sm_count() * 8, 256, 0, s>>>(tb, koff, cell_row, cell_pstart, cell_pcount, sorted, cb, FX, g.kp, g.nvb, k, Vb);
        NLE_LAUNCH_CHECK();
    }
}

// ---- cell index of a whole slab (used by the Sinkhorn pixel pass; the Gram and the extension build their own per batch)
static void cell_index_layout(const AffinityTables& t, size_t& cells, size_t& lev_bytes, size_t& ints) {
    const int capc = std::min(256, (t.cols + 3) & ~3);
    cells = (size_t)capc * t.nrows;
    lev_bytes = (cells + 7) / 8 * 8;
    ints = 2 * ((size_t)t.nrows + 4) + 3 * cells + (size_t)t.nrows * t.cols;
}

size_t cell_index_scratch_doubles(const AffinityTables& t) {
    size_t cells, lev_bytes, ints;
    cell_index_layout(t, cells, lev_bytes, ints);
    return lev_bytes / 8 + (ints * 4 + 7) / 8 + 8;
}

CellIndex build_cell_index(const AffinityTables& t, double* scratch, cudaStream_t s) {
    size_t cells, lev_bytes, ints;
    cell_index_layout(t, cells, lev_bytes, ints);
    uint8_t* cell_lev = reinterpret_cast<uint8_t*>(scratch);
    int* cnt = reinterpret_cast<int*>(cell_lev + lev_bytes);
    int* koff = cnt + t.nrows + 4;
    int* cell_row = koff + t.nrows + 4;
    int* cell_pstart = cell_row + cells;
    int* cell_pcount = cell_pstart + cells;
    int* sorted = cell_pcount + cells;
    const size_t ism = (size_t)(256 * 3 + 260 + 8) * sizeof(int) + ((t.cols + 15) / 16) * 16 + 16;
    if (ism > 227 * 1024) throw Unsupported{"cell index: image too wide (cols=" + std::to_string(t.cols) + ")"};
    allow_max_dynamic_smem((const void*)ext_index_kernel);
    cell_count_kernel<<<std::min(t.nrows, sm_count() * 8), 256, 0, s>>>(t.lum, t.nrows, t.cols, cnt);
    NLE_LAUNCH_CHECK();
    cell_scan_kernel<<<1, 1024, 0, s>>>(cnt, t.nrows, koff);
    NLE_LAUNCH_CHECK();
    ext_index_kernel<<<std::min(t.nrows, sm_count() * 8), 256, ism, s>>>(t.lum, t.nrows, t.cols, koff, cell_lev, cell_row, cell_pstart,
                                                                        cell_pcount, sorted);
    NLE_LAUNCH_CHECK();
    CellIndex ci;
    ci.koff = koff; ci.lev = cell_lev; ci.row = cell_row; ci.pstart = cell_pstart; ci.pcount = cell_pcount; ci.sorted = sorted;
    ci.cap_cells = (int)cells;
    return ci;
}

}  // namespace nle
