// Sinkhorn half-iteration over the rest pixels (filter.cpp:238-245 in factor form) with the
// two sample-axis contractions on the FP64 tensor pipe.
//
// With K(i,j) = Er[row][a] * Ec[col][b] * Gt[|l - Y_ab|]  (kernels.cuh) one half-iteration is
//
//   dot:     y_j  = sum_b Ec[col_j][b] * F[row_j][b][l_j],     F[row][b][l] = sum_a Er[row][a] * (w_ab Gt[|l-Y_ab|])
//            x_j  = |y_j| >= eps ? 1/y_j : 0                    (inplaceReciprocal, filter.cpp:42-54)
//   reduce:  s_ab = sum_l Gt[|l-Y_ab|] * M[b][a][l],            M[b][a][l]   = sum_row Er[row][a] * Hx[row][b][l]
//            Hx[row][b][l] = sum_{col : lum(row,col) = l} Ec[col][b] * x_j
//
// F (for every grid column b an (image rows x nR) * (nR x 256) product) and M (its transpose) are plain dense
// GEMMs over ALL 256 luminance levels: they share the level-by-sample table across image rows (the first
// design, one table look-up per (row, level, sample) in a per-row kernel, was bound by shared-memory look-ups:
// profiles/k7_pass_fused_kernel_full.md).
//
//   sk_dot_gemm_kernel      F  = Er * B_b,  B operand generated in registers (one look-up per 8 DMMAs)
//   sk_pix_cells_kernel     the pixel pass: one warp per (image row, level) cell of the cell index, y, x and the
//                           cell's histogram bins Hx in registers (sample grids up to 64 columns)
//   sk_pix_kernel           the same pass for wider sample grids, one CTA per image row (deterministic: every
//                           (level, b) bin is owned by one lane and filled in ascending column order);
//                           Hx overwrites F in place
//   sk_reduce_gemm_kernel   M partials = Er^T * Hx over row splits
//   sk_reduce_final_kernel  s_ab = sum_l Gt * (sum over splits), fixed order
#include <algorithm>
#include <cstdlib>

#include "kernels.cuh"

namespace nle {

namespace {

constexpr int NL = 256;        // luminance levels
constexpr int DG_ROWS = 32;    // image rows per CTA of the dot GEMM
constexpr int DG_MT = DG_ROWS / 8;   // DMMA m-tiles per warp

// F[row][b][l] = sum_a Er[row][a] * w_ab * Gt[|l - Y_ab|].   grid (ceil(nrows/DG_ROWS), nC), 256 threads.
// Warp w owns levels [32w, 32w+32) (4 n-tiles) for all DG_ROWS rows.
__global__ void __launch_bounds__(256, 2)
sk_dot_gemm_kernel(AffinityTables t, const double* __restrict__ w, double* __restrict__ F) {
    extern __shared__ double dsm[];
    const int nR = t.nR, nC = t.nC;
    const int nRp = (nR + 3) & ~3;
    const int lda = nRp + ((12 - (nRp & 15)) & 15);          // lda % 16 == 12: conflict-free A fragments
    double* As = dsm;                                        // DG_ROWS * lda
    double* Gs = As + DG_ROWS * lda;                         // 256
    double* wc = Gs + NL;                                    // nRp   w_ab of this grid column
    int* yc = reinterpret_cast<int*>(wc + nRp);              // nRp   Y_ab
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tq = lane & 3;
    const int b = blockIdx.y;
    const int r0 = blockIdx.x * DG_ROWS;
    Gs[tid] = t.Gt[tid];
    for (int a = tid; a < nRp; a += 256) {
        wc[a] = a < nR ? w[a * nC + b] : 0.0;
        yc[a] = a < nR ? (int)t.Ysel[a * nC + b] : 0;
    }
    for (int e = tid; e < DG_ROWS * nRp; e += 256) {
        const int r = e / nRp, a = e - r * nRp;
        As[r * lda + a] = (r0 + r < t.nrows && a < nR) ? t.Er[(size_t)(t.row0 + r0 + r) * nR + a] : 0.0;
    }
    __syncthreads();
    double acc[DG_MT][4][2];
#pragma unroll
    for (int u = 0; u < DG_MT; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v][0] = acc[u][v][1] = 0.0;
    const int l0 = warp * 32 + g;
    for (int kk = 0; kk < nRp; kk += 4) {
        const int a = kk + tq;
        const double wv = wc[a];
        const int yv = yc[a];
        double bf[4], af[DG_MT];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int d = l0 + 8 * v - yv;
            bf[v] = wv * Gs[d < 0 ? -d : d];
        }
#pragma unroll
        for (int u = 0; u < DG_MT; ++u) af[u] = As[(8 * u + g) * lda + a];
#pragma unroll
        for (int u = 0; u < DG_MT; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) dmma884(acc[u][v][0], acc[u][v][1], af[u], bf[v]);
    }
#pragma unroll
    for (int u = 0; u < DG_MT; ++u) {
        const int r = r0 + 8 * u + g;
        if (r >= t.nrows) continue;
        double* o = F + ((size_t)r * nC + b) * NL + warp * 32 + 2 * tq;          // F[row][b][l]
#pragma unroll
        for (int v = 0; v < 4; ++v) *reinterpret_cast<double2*>(o + 8 * v) = make_double2(acc[u][v][0], acc[u][v][1]);
    }
}

// ---- level-major tables for the cell pass ----------------------------------------------------------------------------
// sk_pix_cells_kernel reads, per cell, the nC values F[row][.][l] and writes the nC bins Hx[row][.][l]: with the sample column b
// as the fastest index (F[row][l][ldb], ldb = nC rounded up to 4) these are 10 full 32-byte sectors per cell instead of 40
// sectors used for 8 bytes each (ncu, profiles/r2f: the pass is bound by L1 sector throughput; 40 % of its sectors were those;
// measured: 141 -> 98 us per pass).  The two level GEMMs put the sample column on the N axis of the DMMA tile, so that the
// accumulator layout (a lane holds two adjacent N columns, four lanes hold eight) is b-contiguous by itself: no transposition,
// every global access a run of 64 bytes.  (Keeping the level on the N axis and only switching the layout made the dot GEMM 4x
// slower -- scattered 8-byte stores --, transposing through shared memory 2x: 81 and 52 us against 37 and 36.)

// F[row][l][b] = sum_a Er[row][a] * (w_ab * Gt[|l - Y_ab|]):  for a fixed level an (image rows x nR) * (nR x nC) product whose B
// operand is generated in registers from the sample tables.   grid (ceil(nrows/(8 RT)), 256/(8 LW)), 256 threads: warp = LW
// levels, all RT row tiles of the CTA (a generated B fragment -- one look-up and one multiply -- feeds RT DMMAs), LQ levels at a time.
template <int NT, int LQ, int RT, int LW>
__global__ void __launch_bounds__(256, 2)
sk_dot_gemm_nb_kernel(AffinityTables t, const double* __restrict__ w, int ldb, double* __restrict__ F) {
    extern __shared__ double dsm[];
    const int nR = t.nR, nC = t.nC;
    const int nRp = (nR + 3) & ~3;
    const int lda = nRp + ((12 - (nRp & 15)) & 15);          // lda % 16 == 12: conflict-free A fragments
    constexpr int ST = 8 * NT + 4;                           // table row stride, % 16 == 4 or 12: ditto for the B look-ups
    double* As = dsm;                                        // 8 RT * lda
    double* Gs16 = As + 8 * RT * lda;                        // 256 * 16: Gt replicated so that lane L reads bank pair L & 15 -- the
                                                             // look-ups of a warp (32 different samples) are conflict-free
    double* wS = Gs16 + NL * 16;                             // nRp * ST   w_ab, 0 outside the grid
    int* yS = reinterpret_cast<int*>(wS + (size_t)nRp * ST); // nRp * ST   Y_ab
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tq = lane & 3;
    const int r0 = blockIdx.x * (8 * RT);
    {
        const double gv = t.Gt[tid];
#pragma unroll
        for (int k = 0; k < 16; ++k) Gs16[tid * 16 + ((k + tid) & 15)] = gv;      // rotated: the lanes of a store hit different banks
    }
    for (int e = tid; e < nRp * ST; e += 256) {
        const int a = e / ST, b = e - a * ST;
        const bool ok = a < nR && b < nC;
        wS[e] = ok ? w[a * nC + b] : 0.0;
        yS[e] = ok ? (int)t.Ysel[a * nC + b] : 0;
    }
    for (int e = tid; e < 8 * RT * nRp; e += 256) {
        const int r = e / nRp, a = e - r * nRp;
        As[r * lda + a] = (r0 + r < t.nrows && a < nR) ? t.Er[(size_t)(t.row0 + r0 + r) * nR + a] : 0.0;
    }
    __syncthreads();
    const int lbase = 8 * LW * blockIdx.y + LW * warp;
    const double* Gl = Gs16 + (lane & 15);
    for (int lg = 0; lg < LW; lg += LQ) {
        double acc[RT][LQ][NT][2];
#pragma unroll
        for (int u = 0; u < RT; ++u)
#pragma unroll
            for (int j = 0; j < LQ; ++j)
#pragma unroll
                for (int v = 0; v < NT; ++v) acc[u][j][v][0] = acc[u][j][v][1] = 0.0;
        for (int kk = 0; kk < nRp; kk += 4) {
            const int a = kk + tq;
            double af[RT];
#pragma unroll
            for (int u = 0; u < RT; ++u) af[u] = As[(8 * u + g) * lda + a];
            double wv[NT];
            int yv[NT];
#pragma unroll
            for (int v = 0; v < NT; ++v) {
                wv[v] = wS[a * ST + 8 * v + g];
                yv[v] = yS[a * ST + 8 * v + g];
            }
#pragma unroll
            for (int j = 0; j < LQ; ++j) {
                const int l = lbase + lg + j;
#pragma unroll
                for (int v = 0; v < NT; ++v) {
                    const int d = l - yv[v];
                    const double bf = wv[v] * Gl[(d < 0 ? -d : d) << 4];
#pragma unroll
                    for (int u = 0; u < RT; ++u) dmma884(acc[u][j][v][0], acc[u][j][v][1], af[u], bf);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < RT; ++u) {
            const int row = r0 + 8 * u + g;
            if (row >= t.nrows) continue;
#pragma unroll
            for (int j = 0; j < LQ; ++j) {
                double* o = F + ((size_t)row * NL + lbase + lg + j) * ldb + 2 * tq;
#pragma unroll
                for (int v = 0; v < NT; ++v)
                    if (8 * v + 2 * tq < ldb) *reinterpret_cast<double2*>(o + 8 * v) = make_double2(acc[u][j][v][0], acc[u][j][v][1]);
            }
        }
    }
}

// P[ks][level group][i] = sum over the group's 8 levels of  Gt[|l - Y_i|] * sum_{row in split ks} Er[row][a] * Hx[row][l][b],
// i = a * nC + b: for a fixed level an (nR x rows) * (rows x nC) product with both operands read as 64-byte runs; the level
// weight is applied to the finished tile and the 8 warps (= 8 levels) of the CTA are summed in warp order through shared
// memory.  s = sum over (ks, group) of P (sk_sum_partials_kernel).   grid (ceil(256/WARPS), nks, nab), WARPS warps, warp = level.
constexpr int RB_ST = 6;          // ring stages per warp (4 image rows each)
template <int MT, int NT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
sk_reduce_gemm_nb_kernel(AffinityTables t, const double* __restrict__ Hx, int ldb, int nks, int ring_doubles, double* __restrict__ P) {
    extern __shared__ double rsm[];          // 8 warps x ring_doubles (cp.async rings, later the weighted tiles) | Er slice
    const int nR = t.nR, nC = t.nC;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tq = lane & 3;
    const int l = min(WARPS * (int)blockIdx.x + warp, NL - 1), ks = blockIdx.y, a0 = blockIdx.z * (8 * MT);
    const bool live = WARPS * (int)blockIdx.x + warp < NL;      // the last CTA of a level split that does not divide 256
    const int rb = (int)(((long long)t.nrows * ks) / nks);
    const int re = (int)(((long long)t.nrows * (ks + 1)) / nks);
    double acc[MT][NT][2];
#pragma unroll
    for (int u = 0; u < MT; ++u)
#pragma unroll
        for (int v = 0; v < NT; ++v) acc[u][v][0] = acc[u][v][1] = 0.0;
    // Er rows of the CTA's split (shared by its 8 warps): rows x 8 MT values, stride % 16 == 4 or 12 (conflict-free A fragments)
    constexpr int AS = 8 * MT + 4;
    double* ErS = rsm + (size_t)WARPS * ring_doubles;
    for (int e = tid; e < (re - rb) * (8 * MT); e += WARPS * 32) {
        const int rr = e / (8 * MT), a = e - rr * (8 * MT);
        ErS[rr * AS + a] = (a0 + a < nR) ? t.Er[(size_t)(t.row0 + rb + rr) * nR + a0 + a] : 0.0;
    }
    // Per-warp ring of RB_ST stages x 4 image rows of the warp's level, filled with 16-byte cp.async copies (each row of Hx[.][l][.]
    // is one contiguous run of ldb doubles): the loads run RB_ST - 1 DMMA rounds ahead of their use without holding registers.
    // Row stride % 16 == 4: the B fragments (4 rows x 8 columns per half warp) fall into 16 different banks.
    const int RS = ldb + ((20 - (ldb & 15)) & 15);
    double* ring = rsm + (size_t)warp * ring_doubles;
    const int nvec = ldb / 2;                                   // 16-byte pieces per row; 4 * nvec <= 128 pieces per stage
    int p_row[4], p_dst[4], p_src[4];                           // the pieces this lane copies: row in the stage, offsets (doubles)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int e = lane + 32 * i;
        const int rr = e / nvec, c = e - rr * nvec;
        p_row[i] = (e < 4 * nvec) ? rr : 1 << 20;
        p_dst[i] = rr * RS + 2 * c;
        p_src[i] = rr * NL * ldb + 2 * c;
    }
    const unsigned ring_sa = (unsigned)__cvta_generic_to_shared(ring);
    const double* hx_l = Hx + (size_t)l * ldb;
    auto issue = [&](int r, int stage) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (r + p_row[i] < re)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring_sa + 8u * (unsigned)(stage * 4 * RS + p_dst[i])),
                             "l"(hx_l + (size_t)r * NL * ldb + p_src[i])
                             : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int nsteps = live ? (re - rb + 3) / 4 : 0;
#pragma unroll
    for (int i = 0; i < RB_ST - 1; ++i) issue(rb + 4 * i, i);      // empty groups past the end keep the group count uniform
    __syncthreads();                                               // ErS complete
    int stage = 0;
    for (int it = 0; it < nsteps; ++it) {
        const int r = rb + 4 * it;
        issue(r + 4 * (RB_ST - 1), stage == 0 ? RB_ST - 1 : stage - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(RB_ST - 1) : "memory");
        __syncwarp();
        const double* cur = ring + (size_t)stage * 4 * RS;
        const bool ok = r + tq < re;
        double af[MT], bf[NT];
        const double* er = ErS + (size_t)(r - rb + tq) * AS + g;
#pragma unroll
        for (int u = 0; u < MT; ++u) af[u] = ok ? er[8 * u] : 0.0;
#pragma unroll
        for (int v = 0; v < NT; ++v) bf[v] = (ok && 8 * v + g < ldb) ? cur[tq * RS + 8 * v + g] : 0.0;
#pragma unroll
        for (int u = 0; u < MT; ++u)
#pragma unroll
            for (int v = 0; v < NT; ++v) dmma884(acc[u][v][0], acc[u][v][1], af[u], bf[v]);
        __syncwarp();                                              // the stage is refilled RB_ST - 1 rounds later
        stage = (stage + 1 == RB_ST) ? 0 : stage + 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                               // the rings are reused for the weighted tiles below
    const int tile = 8 * MT * nC;             // <= ring_doubles
    double* mine = rsm + (size_t)warp * tile;
#pragma unroll
    for (int u = 0; u < MT; ++u) {
        const int a = a0 + 8 * u + g;
#pragma unroll
        for (int v = 0; v < NT; ++v)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int b = 8 * v + 2 * tq + e;
                if (a < nR && b < nC) {
                    const int d = l - (int)t.Ysel[a * nC + b];
                    mine[(8 * u + g) * nC + b] = live ? t.Gt[d < 0 ? -d : d] * acc[u][v][e] : 0.0;
                }
            }
    }
    __syncthreads();
    const int na = min(8 * MT, nR - a0);
    double* out = P + ((size_t)ks * gridDim.x + blockIdx.x) * t.p + (size_t)a0 * nC;
    for (int e = tid; e < na * nC; e += WARPS * 32) {
        double sum = 0.0;
#pragma unroll
        for (int wv = 0; wv < WARPS; ++wv) sum += rsm[(size_t)wv * tile + e];
        out[e] = sum;
    }
}

// s[i] = sum_q P[q][i], q = 0 .. nq-1: eight interleaved chains per sample (warp c of the CTA takes q = c, c + 8, ...; a warp reads 32
// consecutive samples), eight loads in flight per thread, chains combined in a fixed order.  The partials (nq ~ 130 vectors) come
// straight from the reduce GEMM, i.e. from L2: the kernel is a chain of dependent load batches between two dependent kernels.
__global__ void __launch_bounds__(256)
sk_sum_partials_kernel(const double* __restrict__ P, int nq, int p, double* __restrict__ s_out) {
    __shared__ double part[8][32];
    const int lane = threadIdx.x & 31, c = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    double acc = 0.0;
    if (i < p) {
        int q = c;
        for (; q + 56 < nq; q += 64) {
            double x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) x[u] = P[(size_t)(q + 8 * u) * p + i];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += x[u];
        }
        double x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = (q + 8 * u < nq) ? P[(size_t)(q + 8 * u) * p + i] : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += x[u];
    }
    part[c][lane] = acc;
    __syncthreads();
    if (threadIdx.x < 32 && i < p)
        s_out[i] = ((part[0][lane] + part[1][lane]) + (part[2][lane] + part[3][lane])) +
                   ((part[4][lane] + part[5][lane]) + (part[6][lane] + part[7][lane]));
}

// One CTA per image row.  FH row block: [nC][256] in global memory; on entry F (ignored when w_given == 0:
// the initial pass with x = 1), on exit Hx.  x: slab vector (0 at sample pixels).
__global__ void __launch_bounds__(256)
sk_pix_kernel(AffinityTables t, int w_given, double* __restrict__ x, double* __restrict__ FH) {
    extern __shared__ double psm[];
    const int nC = t.nC, W = t.cols;
    const int nCp = nC | 1;                                  // odd stride: transposed copies are conflict-free
    double* Ts = psm;                                        // NL * nCp   F, then the histogram
    double* stage = Ts + (size_t)NL * nCp;                   // 32 * nC    Ec rows of 32 consecutive pixels
    double* xrow = stage + 32 * nC;                          // W
    uint8_t* Lrow = reinterpret_cast<uint8_t*>(xrow + W);    // W
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int rl = blockIdx.x; rl < t.nrows; rl += gridDim.x) {
        const int row = t.row0 + rl;
        double* fh = FH + (size_t)rl * nC * NL;
        const uint8_t* Lg = t.lum + (size_t)rl * W;
        __syncthreads();
        for (int c = tid; c < W; c += 256) Lrow[c] = Lg[c];
        const int a_row = t.rowa[row];
        double* xo = x + (size_t)rl * W;
        if (w_given) {
            for (int e = tid; e < nC * NL; e += 256) Ts[(e & (NL - 1)) * nCp + (e >> 8)] = fh[e];
            __syncthreads();
            for (int c = tid; c < W; c += 256) {
                double r = 0.0;
                if (!(a_row >= 0 && t.colb[c] >= 0)) {
                    const double* f = Ts + (size_t)Lrow[c] * nCp;
                    double acc = 0.0;
                    for (int b = 0; b < nC; ++b) acc = fma(t.EcT[(size_t)b * W + c], f[b], acc);
                    r = (fabs(acc) >= kEps) ? 1.0 / acc : 0.0;
                }
                xo[c] = r;
                xrow[c] = r;
            }
        } else {
            __syncthreads();
            for (int c = tid; c < W; c += 256) {
                const double r = (a_row >= 0 && t.colb[c] >= 0) ? 0.0 : 1.0;
                xo[c] = r;
                xrow[c] = r;
            }
        }
        __syncthreads();
        for (int e = tid; e < NL * nCp; e += 256) Ts[e] = 0.0;
        // Hx[l][b] += Ec[col][b] * x_col: warp `warp` owns the levels with (l & 7) == warp, lanes own b; pixels are
        // visited in ascending column order, 32 at a time, their Ec rows staged in shared memory
        for (int c0 = 0; c0 < W; c0 += 32) {
            __syncthreads();
            const int nst = min(32, W - c0) * nC;
            for (int e = tid; e < nst; e += 256) stage[e] = t.Ec[(size_t)c0 * nC + e];
            __syncthreads();
            const int c = c0 + lane;
            const int lv_l = (c < W) ? (int)Lrow[c] : 0;
            const bool mine = (c < W) && ((lv_l & 7) == warp) && (xrow[c] != 0.0);
            unsigned m = __ballot_sync(0xffffffffu, mine);
            while (m) {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                const int lv = __shfl_sync(0xffffffffu, lv_l, j);
                const double xv = xrow[c0 + j];
                const double* ec = stage + j * nC;
                double* h = Ts + (size_t)lv * nCp;
                for (int b = lane; b < nC; b += 32) h[b] = fma(ec[b], xv, h[b]);
            }
        }
        __syncthreads();
        for (int e = tid; e < nC * NL; e += 256) fh[e] = Ts[(e & (NL - 1)) * nCp + (e >> 8)];
    }
}

// Row pass over the cell index, one WARP per cell and no shared memory or block barrier at all: the pixels of a cell
// share the level, so its F row (nC values) lives in registers and its
// histogram bins are register accumulators that are stored once.  Lanes = 4 pixel slots x 8 b-lanes (b = sub + 8i):
// every Ec row is loaded once and serves both the dot (xor-shuffle tree inside the 8-lane group) and the histogram.
//   F: [row][l][ldb] level-major (read: the nC values of a cell are contiguous), H: same layout (written for the non-empty
//   cells only; the rest was zeroed once per training call and is never touched), x: slab vector.
template <int NB>
__global__ void __launch_bounds__(256)
sk_pix_cells_kernel(AffinityTables t, CellIndex ci, int w_given, const double* __restrict__ F, int ldb, double* __restrict__ x,
                    double* __restrict__ H) {
    const int nC = t.nC, W = t.cols;
    const int lane = threadIdx.x & 31, q = lane >> 3, sub = lane & 7;
    const int K = ci.koff[t.nrows];
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    // the metadata of the warp's NEXT cell is fetched while the current one is processed (the per-cell chain
    // metadata -> columns -> Ec rows is three dependent L2 round trips)
    int cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int m_np = 0, m_rl = 0, m_lev = 0, m_ps = 0;
    if (cell < K) { m_np = ci.pcount[cell]; m_rl = ci.row[cell]; m_lev = (int)ci.lev[cell]; m_ps = ci.pstart[cell]; }
    for (; cell < K; cell += nwarps) {
        const int np = m_np, rl = m_rl, lev = m_lev;
        const int* pix = ci.sorted + m_ps;
        {
            const int nx = cell + nwarps;
            if (nx < K) { m_np = ci.pcount[nx]; m_rl = ci.row[nx]; m_lev = (int)ci.lev[nx]; m_ps = ci.pstart[nx]; }
        }
        if (np == 0) continue;
        const int a_row = t.rowa[t.row0 + rl];
        double f[NB], acc[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int b = sub + 8 * i;
            f[i] = (w_given && b < nC) ? F[((size_t)rl * NL + lev) * ldb + b] : 0.0;      // F[row][l][b]: the cell's row is contiguous
            acc[i] = 0.0;
        }
        for (int p0 = 0; p0 < np; p0 += 4) {
            const bool ok = p0 + q < np;
            const int col = ok ? pix[p0 + q] : 0;
            const double* ecr = t.Ec + (size_t)col * nC;
            double e[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const int b = sub + 8 * i;
                e[i] = (ok && b < nC) ? ecr[b] : 0.0;
            }
            double xv = 1.0;
            if (w_given) {
                double v = 0.0;
#pragma unroll
                for (int i = 0; i < NB; ++i) v = fma(e[i], f[i], v);
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                xv = (fabs(v) >= kEps) ? 1.0 / v : 0.0;
            }
            if (!ok || (a_row >= 0 && t.colb[col] >= 0)) xv = 0.0;
            if (ok && sub == 0) x[(size_t)rl * W + col] = xv;
#pragma unroll
            for (int i = 0; i < NB; ++i) acc[i] = fma(e[i], xv, acc[i]);
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
            acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
        }
        if (q == 0) {
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const int b = sub + 8 * i;
                if (b < nC) H[((size_t)rl * NL + lev) * ldb + b] = acc[i];
            }
        }
    }
}

// Mpart[ks][b][a][l] = sum_{row in split ks} Er[row][a] * Hx[row][b][l].
// grid (nC, nks, nab), 256 threads; warp w owns levels [32w, 32w+32) (4 n-tiles) for MT m-tiles of grid rows.
template <int MT>
__global__ void __launch_bounds__(256, 2)
sk_reduce_gemm_kernel(AffinityTables t, const double* __restrict__ Hx, int nks, int nRp, double* __restrict__ Mpart) {
    const int nR = t.nR, nC = t.nC;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tq = lane & 3;
    const int b = blockIdx.x, ks = blockIdx.y, a0 = blockIdx.z * (8 * MT);
    const int rb = (int)(((long long)t.nrows * ks) / nks);
    const int re = (int)(((long long)t.nrows * (ks + 1)) / nks);
    double acc[MT][4][2];
#pragma unroll
    for (int u = 0; u < MT; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v][0] = acc[u][v][1] = 0.0;
    const double* hb = Hx + (size_t)b * NL + warp * 32 + g;
    auto load = [&](int r, double* af, double* bf) {
        const int rr = r + tq;
        const bool ok = rr < re;
        const double* er = t.Er + (size_t)(t.row0 + rr) * nR;
        const double* h = hb + (size_t)rr * nC * NL;
#pragma unroll
        for (int u = 0; u < MT; ++u) {
            const int a = a0 + 8 * u + g;
            af[u] = (ok && a < nR) ? er[a] : 0.0;
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) bf[v] = ok ? h[8 * v] : 0.0;
    };
    double af0[MT], bf0[4], af1[MT], bf1[4];
    if (rb < re) load(rb, af0, bf0);
    for (int r = rb; r < re; r += 8) {
        if (r + 4 < re) load(r + 4, af1, bf1);
#pragma unroll
        for (int u = 0; u < MT; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) dmma884(acc[u][v][0], acc[u][v][1], af0[u], bf0[v]);
        if (r + 4 >= re) break;
        if (r + 8 < re) load(r + 8, af0, bf0);
#pragma unroll
        for (int u = 0; u < MT; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) dmma884(acc[u][v][0], acc[u][v][1], af1[u], bf1[v]);
    }
    double* out = Mpart + (((size_t)ks * nC + b) * nRp) * NL;
#pragma unroll
    for (int u = 0; u < MT; ++u) {
        const int a = a0 + 8 * u + g;
        if (a >= nRp) continue;
        double* o = out + (size_t)a * NL + warp * 32 + 2 * tq;
#pragma unroll
        for (int v = 0; v < 4; ++v) *reinterpret_cast<double2*>(o + 8 * v) = make_double2(acc[u][v][0], acc[u][v][1]);
    }
}

// s[a*nC+b] = sum_l Gt[|l - Y_ab|] * sum_ks Mpart[ks][b][a][l].   One warp per sample, fixed summation order.
__global__ void __launch_bounds__(256)
sk_reduce_final_kernel(AffinityTables t, const double* __restrict__ Mpart, int nks, int nRp, double* __restrict__ s_out) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= t.p) return;
    const int a = i / t.nC, b = i - a * t.nC;
    const int y = (int)t.Ysel[i];
    // the 8 level groups of a lane are independent: their loads are issued together (one L2 round trip per split)
    double m[NL / 32];
#pragma unroll
    for (int q = 0; q < NL / 32; ++q) m[q] = 0.0;
    for (int ks = 0; ks < nks; ++ks) {
        const double* src = Mpart + (((size_t)ks * t.nC + b) * nRp + a) * NL + lane;
#pragma unroll
        for (int q = 0; q < NL / 32; ++q) m[q] += src[32 * q];
    }
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < NL / 32; ++q) {
        const int d = lane + 32 * q - y;
        acc = fma(t.Gt[d < 0 ? -d : d], m[q], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_out[i] = acc;
}

struct SkGeom {
    int nRp, MT, nab, nks, ldb;
    int NTb, MTb, nabb, nksb, ringb, wrb, lgb;   // level-major GEMMs (wrb warps = levels per CTA of the reduce GEMM, lgb level groups) of the cell path: n-tiles of sample columns, m-tiles of grid rows per CTA
    size_t fh_doubles, mpart_doubles, pix_smem, dot_smem, dot_nb_smem, red_nb_smem;
};

SkGeom sk_geometry(const AffinityTables& t) {
    SkGeom g;
    g.nRp = (t.nR + 7) & ~7;
    const int tiles = g.nRp / 8;
    g.nab = cdiv(tiles, 5);
    g.MT = cdiv(tiles, g.nab);
    g.nRp = g.nab * g.MT * 8;
    g.nks = std::max(1, std::min(std::max(1, t.nrows / 32), (2 * sm_count()) / std::max(1, t.nC * g.nab)));
    g.ldb = (t.nC + 3) & ~3;                                  // level-major tables of the cell path: F[row][l][ldb]
    g.fh_doubles = (size_t)t.nrows * g.ldb * NL;
    g.mpart_doubles = (size_t)g.nks * t.nC * g.nRp * NL;
    const int nCp = t.nC | 1;
    g.pix_smem = ((size_t)NL * nCp + 32 * (size_t)t.nC + t.cols) * sizeof(double) + ((t.cols + 15) / 16) * 16 + 64;
    const int nR4 = (t.nR + 3) & ~3;
    const int lda = nR4 + ((12 - (nR4 & 15)) & 15);
    g.dot_smem = ((size_t)DG_ROWS * lda + NL + nR4) * sizeof(double) + (size_t)nR4 * sizeof(int) + 64;
    g.NTb = std::max(1, std::min(8, cdiv(t.nC, 8)));
    static const int mt_for_nt[9] = {0, 7, 7, 7, 7, 5, 4, 4, 3};        // MTb * NTb * 2 <= 56 accumulators per lane
    g.MTb = mt_for_nt[g.NTb];
    g.nabb = cdiv(cdiv(t.nR, 8), g.MTb);
    g.wrb = (g.MTb * g.NTb <= 25) ? 12 : 8;                   // three warps per sub-core where the tile leaves the registers for it
    g.lgb = cdiv(NL, g.wrb);
    const int st = 8 * g.NTb + 4;
    g.dot_nb_smem = ((size_t)32 * lda + 16 * NL + (size_t)nR4 * st) * sizeof(double) + (size_t)nR4 * st * sizeof(int) + 64;
    g.ringb = (int)std::max((size_t)(8 * g.MTb) * t.nC, (size_t)RB_ST * 4 * (g.ldb + 16));
    g.ringb = (g.ringb + 1) & ~1;
    // row splits: one wave of one CTA per SM; the Er slice of a split sits in shared memory next to the rings
    const long long room = 227 * 1024 - 64 - (long long)g.wrb * g.ringb * 8;
    const int cap_rows = (int)std::max<long long>(8, room / ((8 * g.MTb + 4) * 8) - 1);
    g.nksb = std::max({1, std::min(std::max(1, t.nrows / 32), sm_count() / (g.lgb * g.nabb)), cdiv(t.nrows, cap_rows)});
    const int rows_split = cdiv(t.nrows, g.nksb) + 1;
    g.red_nb_smem = ((size_t)g.wrb * g.ringb + (size_t)rows_split * (8 * g.MTb + 4)) * sizeof(double) + 64;
    return g;
}

}  // namespace

bool sinkhorn_cells_supported(const AffinityTables& t) {
    const SkGeom g = sk_geometry(t);
    return t.nC <= 64 ? (g.dot_nb_smem <= 227 * 1024 && g.red_nb_smem <= 227 * 1024) : (g.pix_smem <= 227 * 1024 && g.dot_smem <= 227 * 1024);
}

size_t sinkhorn_cells_scratch_doubles(const AffinityTables& t) {
    const SkGeom g = sk_geometry(t);
    return 2 * g.fh_doubles + std::max(g.mpart_doubles, (size_t)g.nksb * g.lgb * t.p) + 8;
}

// Zeroes the histogram table once per training call (the cell pass only ever writes the non-empty cells).
void sinkhorn_cells_prepare(const AffinityTables& t, double* scratch, cudaStream_t s) {
    const SkGeom g = sk_geometry(t);
    NLE_CUDA(cudaMemsetAsync(scratch + g.fh_doubles, 0, g.fh_doubles * sizeof(double), s));
}

template <int NB>
static void launch_pc(const AffinityTables& t, const CellIndex& ci, int w_given, const double* F, int ldb, double* x, double* H,
                      cudaStream_t s) {
    sk_pix_cells_kernel<NB><<<sm_count() * 8, 256, 0, s>>>(t, ci, w_given, F, ldb, x, H);
    NLE_LAUNCH_CHECK();
}

template <int MT>
static void launch_rg(const AffinityTables& t, const SkGeom& g, const double* FH, double* Mpart, cudaStream_t s) {
    sk_reduce_gemm_kernel<MT><<<dim3(t.nC, g.nks, g.nab), 256, 0, s>>>(t, FH, g.nks, g.nRp, Mpart);
    NLE_LAUNCH_CHECK();
}

template <int NT, int LQ, int RT>
static void launch_dot_nb(const AffinityTables& t, const SkGeom& g, const double* w, double* F, cudaStream_t s) {
    constexpr int LW = (RT == 4) ? 4 : 8;       // 8 RT rows x 8 LW levels per CTA: 32 x 32 or 16 x 64 -> the same number of CTAs
    // the device maximum, not this call's size: host threads that train different slabs on one device share the attribute
    allow_max_dynamic_smem((const void*)sk_dot_gemm_nb_kernel<NT, LQ, RT, LW>);
    sk_dot_gemm_nb_kernel<NT, LQ, RT, LW><<<dim3(cdiv(t.nrows, 8 * RT), NL / (8 * LW)), 256, g.dot_nb_smem, s>>>(t, w, g.ldb, F);
    NLE_LAUNCH_CHECK();
}

template <int MT, int NT, int WARPS>
static void launch_red_nb(const AffinityTables& t, const SkGeom& g, const double* Hx, double* P, cudaStream_t s) {
    allow_max_dynamic_smem((const void*)sk_reduce_gemm_nb_kernel<MT, NT, WARPS>);
    sk_reduce_gemm_nb_kernel<MT, NT, WARPS><<<dim3(g.lgb, g.nksb, g.nabb), WARPS * 32, g.red_nb_smem, s>>>(t, Hx, g.ldb, g.nksb, g.ringb, P);
    NLE_LAUNCH_CHECK();
}

// One half-iteration:  x = recip(k_j^T w) on the rest pixels (w == nullptr: x = 1), then s = Kab x.
// ci != nullptr: pixel pass over the cell index (sk_pix_cells_kernel; sinkhorn_cells_prepare must have run on this
// scratch); otherwise the staged per-row kernels.
void launch_sinkhorn_cells(const AffinityTables& t, const CellIndex* ci, const double* w, double* x, double* scratch,
                           double* s_out, cudaStream_t s) {
    const SkGeom g = sk_geometry(t);
    const bool cells = ci != nullptr && t.nC <= 64;    // wider sample grids: the per-row staged kernel
    double* FH = scratch;                              // staged path: F, overwritten in place by the histogram
    double* Hc = scratch + g.fh_doubles;               // cell path: separate histogram table
    double* Mpart = scratch + 2 * g.fh_doubles;
    if (!cells) allow_max_dynamic_smem((const void*)sk_pix_kernel);
    if (w) {
        if (cells) {
            switch (g.NTb) {
                case 1: launch_dot_nb<1, 1, 4>(t, g, w, FH, s); break;
                case 2: launch_dot_nb<2, 1, 4>(t, g, w, FH, s); break;
                case 3: launch_dot_nb<3, 1, 4>(t, g, w, FH, s); break;
                case 4: launch_dot_nb<4, 1, 4>(t, g, w, FH, s); break;
                case 5: launch_dot_nb<5, 1, 4>(t, g, w, FH, s); break;
                case 6: launch_dot_nb<6, 1, 2>(t, g, w, FH, s); break;
                case 7: launch_dot_nb<7, 1, 2>(t, g, w, FH, s); break;
                default: launch_dot_nb<8, 1, 2>(t, g, w, FH, s); break;
            }
        } else {
            allow_max_dynamic_smem((const void*)sk_dot_gemm_kernel);
            sk_dot_gemm_kernel<<<dim3(cdiv(t.nrows, DG_ROWS), t.nC), 256, g.dot_smem, s>>>(t, w, FH);
            NLE_LAUNCH_CHECK();
        }
    }
    const double* Hx = FH;
    if (cells) {
        const int nb = cdiv(t.nC, 8);
        switch (nb) {
            case 1: launch_pc<1>(t, *ci, w ? 1 : 0, FH, g.ldb, x, Hc, s); break;
            case 2: launch_pc<2>(t, *ci, w ? 1 : 0, FH, g.ldb, x, Hc, s); break;
            case 3: launch_pc<3>(t, *ci, w ? 1 : 0, FH, g.ldb, x, Hc, s); break;
            case 4: launch_pc<4>(t, *ci, w ? 1 : 0, FH, g.ldb, x, Hc, s); break;
            case 5: launch_pc<5>(t, *ci, w ? 1 : 0, FH, g.ldb, x, Hc, s); break;
            case 6: launch_pc<6>(t, *ci, w ? 1 : 0, FH, g.ldb, x, Hc, s); break;
            case 7: launch_pc<7>(t, *ci, w ? 1 : 0, FH, g.ldb, x, Hc, s); break;
            default: launch_pc<8>(t, *ci, w ? 1 : 0, FH, g.ldb, x, Hc, s); break;
        }
        Hx = Hc;
    } else {
        sk_pix_kernel<<<t.nrows, 256, g.pix_smem, s>>>(t, w ? 1 : 0, x, FH);
        NLE_LAUNCH_CHECK();
    }
    if (cells) {
        switch (g.NTb) {
            case 1: launch_red_nb<7, 1, 12>(t, g, Hx, Mpart, s); break;
            case 2: launch_red_nb<7, 2, 12>(t, g, Hx, Mpart, s); break;
            case 3: launch_red_nb<7, 3, 12>(t, g, Hx, Mpart, s); break;
            case 4: launch_red_nb<7, 4, 8>(t, g, Hx, Mpart, s); break;
            case 5: launch_red_nb<5, 5, 12>(t, g, Hx, Mpart, s); break;
            case 6: launch_red_nb<4, 6, 12>(t, g, Hx, Mpart, s); break;
            case 7: launch_red_nb<4, 7, 8>(t, g, Hx, Mpart, s); break;
            default: launch_red_nb<3, 8, 12>(t, g, Hx, Mpart, s); break;
        }
        sk_sum_partials_kernel<<<cdiv(t.p, 32), 256, 0, s>>>(Mpart, g.nksb * g.lgb, t.p, s_out);
        NLE_LAUNCH_CHECK();
        return;
    }
    switch (g.MT) {
        case 1: launch_rg<1>(t, g, Hx, Mpart, s); break;
        case 2: launch_rg<2>(t, g, Hx, Mpart, s); break;
        case 3: launch_rg<3>(t, g, Hx, Mpart, s); break;
        case 4: launch_rg<4>(t, g, Hx, Mpart, s); break;
        default: launch_rg<5>(t, g, Hx, Mpart, s); break;
    }
    sk_reduce_final_kernel<<<cdiv(t.p, 8), 256, 0, s>>>(t, Mpart, g.nks, g.nRp, s_out);
    NLE_LAUNCH_CHECK();
}

}  // namespace nle
