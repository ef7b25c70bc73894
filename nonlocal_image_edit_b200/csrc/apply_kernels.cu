// The HBM-bound output path (SURVEY.md kernel K5): NLEFilter::apply  out = V (fS o (V^T z))  (filter.cpp:445-458), fused
// with what surrounds it in NLEFilter::enhance (filter.cpp:422-440): 8-bit BGR -> Lab on the way in, max(.,0), min(.,255),
// convertTo(CV_8U) (round half to even) and Lab -> BGR on the way out.
//
// V is nloc x k doubles, row-major and dense, so any range of rows is ONE contiguous byte range: both passes stream it
// through shared memory with the TMA engine's 1-D bulk copy (cp.async.bulk.shared::cluster.global + mbarrier
// complete_tx; SASS UBLKCP / SYNCS.ARRIVE.TRANS64) in a ring of up to 4 stages of ~50 KB, one persistent CTA per SM.
// No thread issues a global load for V; the 256 threads only read shared memory (four independent partial sums per
// thread: with 8 warps per SM the FP64 FMA latency of a single dependent chain was the bound, profiles/r2c):
//   vtz_tma_kernel        t = V^T z.  Thread = (column v, row group): consecutive lanes read consecutive doubles of a row.
//   recompose_tma_kernel  out_j = V_j . g, g = fS o t.  Thread = row, columns rotated by the lane so that the 16 lanes of
//                         a half-warp fall into 16 different bank pairs although consecutive rows are k doubles apart.
//                         Tiles are visited in DESCENDING order: the tail of V that vtz just streamed is still in the
//                         126 MB L2.
// The colour conversions are byte-exact with cv::cvtColor(COLOR_BGR2Lab / COLOR_Lab2BGR) on CV_8UC3.  OpenCV is an
// un-vendored, unpinned dependency of the reference (CMakeLists.txt:34); its 8-bit Lab path is fixed-point arithmetic
// over small tables (imgproc color_lab.cpp: RGB2Lab_b, Lab2RGBinteger).  scripts/make_lab_tables.py restates the tables,
// verifies the whole integer pipeline against cv2 on all 2^24 BGR and all 2^24 Lab triples, and writes lab_tables.inc;
// tests/test_gpu_lab.py repeats the exhaustive comparison through these kernels.
#include <algorithm>

#include "kernels.cuh"

namespace nle {

namespace {

#define NLE_LAB_TAB __device__ static const
#include "lab_tables.inc"
#undef NLE_LAB_TAB

constexpr int kGammaN = 256, kCbrtN = 3072, kInvGammaN = 4096;
constexpr int kLabShift = 12, kLabShift2 = 15, kBase = 1 << 14;

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
__device__ __forceinline__ int sat_u8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// abToXZ_b of color_lab.cpp as pure integer arithmetic (C division truncates toward zero)
__device__ __forceinline__ int ab_to_xz(int i) {
    if (i <= 3390) return i * 108 / 841 - kBase * 16 / 116 * 108 / 841;
    const long long ii = (long long)i * i / kBase;
    return (int)(ii * i / kBase);
}

// Shared-memory copies of the conversion tables (13.7 KB when both directions are needed).
struct LabTables {
    unsigned short gam[kGammaN];
    unsigned short cbr[kCbrtN];
    unsigned short ytab[256], fytab[256];
    unsigned char ig[kInvGammaN];
};

__device__ __forceinline__ void lab_tables_load(LabTables& T, bool fwd, bool inv, int tid, int nthreads) {
    if (fwd) {
        for (int i = tid; i < kGammaN; i += nthreads) T.gam[i] = kLabGammaTab[i];
        for (int i = tid; i < kCbrtN; i += nthreads) T.cbr[i] = kLabCbrtTab[i];
    }
    if (inv) {
        for (int i = tid; i < 256; i += nthreads) { T.ytab[i] = kLabYTab[i]; T.fytab[i] = kLabFyTab[i]; }
        for (int i = tid; i < kInvGammaN; i += nthreads) T.ig[i] = kLabInvGammaTab[i];
    }
}

// RGB2Lab_b: one BGR pixel -> (L, a, b)
__device__ __forceinline__ void bgr_to_lab_px(const LabTables& T, int b8, int g8, int r8, int& L, int& a, int& b) {
    const int B = T.gam[b8], G = T.gam[g8], R = T.gam[r8];
    const int fX = T.cbr[descale(R * kLabFwdCoef[0] + G * kLabFwdCoef[1] + B * kLabFwdCoef[2], kLabShift)];
    const int fY = T.cbr[descale(R * kLabFwdCoef[3] + G * kLabFwdCoef[4] + B * kLabFwdCoef[5], kLabShift)];
    const int fZ = T.cbr[descale(R * kLabFwdCoef[6] + G * kLabFwdCoef[7] + B * kLabFwdCoef[8], kLabShift)];
    const int Lscale = (116 * 255 + 50) / 100;
    const int Lshift = -((16 * 255 * (1 << kLabShift2) + 50) / 100);
    L = sat_u8(descale(Lscale * fY + Lshift, kLabShift2));
    a = sat_u8(descale(500 * (fX - fY) + 128 * (1 << kLabShift2), kLabShift2));
    b = sat_u8(descale(200 * (fY - fZ) + 128 * (1 << kLabShift2), kLabShift2));
}

// Lab2RGBinteger: (L, a, b) -> one BGR pixel
__device__ __forceinline__ void lab_to_bgr_px(const LabTables& T, int l, int a, int b, uint8_t* bgr) {
    const int y = T.ytab[l], ify = T.fytab[l];
    const int adiv = ((5 * a * 53687 + (1 << 7)) >> 13) - 128 * kBase / 500;
    const int bdiv = ((b * 41943 + (1 << 4)) >> 9) - 128 * kBase / 200 + 1;
    const int x = ab_to_xz(ify + adiv), z = ab_to_xz(ify - bdiv);
    int rgb[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        // |C x| < 2^14 * 2^17: the sum needs 64 bits before the shift
        const long long acc = (long long)kLabInvCoef[3 * r] * x + (long long)kLabInvCoef[3 * r + 1] * y + (long long)kLabInvCoef[3 * r + 2] * z;
        int v = (int)((acc + (1 << 13)) >> 14);
        v = v < 0 ? 0 : (v > kInvGammaN - 1 ? kInvGammaN - 1 : v);
        rgb[r] = T.ig[v];
    }
    bgr[0] = (uint8_t)rgb[2];
    bgr[1] = (uint8_t)rgb[1];
    bgr[2] = (uint8_t)rgb[0];
}

// bgr: npix x 3 interleaved.  Writes L (npix) and, if ab != nullptr, ab (npix x 2 interleaved).
__global__ void __launch_bounds__(256)
bgr2lab_kernel(const uint8_t* __restrict__ bgr, long long npix, uint8_t* __restrict__ L, uint8_t* __restrict__ ab) {
    __shared__ LabTables T;
    lab_tables_load(T, true, false, threadIdx.x, 256);
    __syncthreads();
    for (long long j = (long long)blockIdx.x * 256 + threadIdx.x; j < npix; j += (long long)gridDim.x * 256) {
        int l, a, b;
        bgr_to_lab_px(T, bgr[3 * j], bgr[3 * j + 1], bgr[3 * j + 2], l, a, b);
        L[j] = (uint8_t)l;
        if (ab) { ab[2 * j] = (uint8_t)a; ab[2 * j + 1] = (uint8_t)b; }
    }
}

__global__ void __launch_bounds__(256)
lab2bgr_kernel(const uint8_t* __restrict__ L, const uint8_t* __restrict__ ab, long long npix, uint8_t* __restrict__ bgr) {
    __shared__ LabTables T;
    lab_tables_load(T, false, true, threadIdx.x, 256);
    __syncthreads();
    for (long long j = (long long)blockIdx.x * 256 + threadIdx.x; j < npix; j += (long long)gridDim.x * 256)
        lab_to_bgr_px(T, L[j], ab[2 * j], ab[2 * j + 1], bgr + 3 * j);
}

// ---------------------------------------------------------------------------------------------
// TMA ring: 1-D bulk copies global -> shared, completion on an mbarrier per stage.
constexpr int AP_THREADS = 256;
constexpr int AP_MAXSTAGES = 4;
constexpr int AP_KMAX = 400;            // widest V row the shared-memory ring holds (32 rows x 400 doubles = 100 KB x 2 stages)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct ApplyGeom {
    int TR;          // rows per tile (32, 64 or 128)
    int stages;
    size_t stage_doubles;
    long long ntiles;
    int grid;
};

ApplyGeom apply_geometry(long long nloc, int k) {
    ApplyGeom g;
    g.TR = 128;
    while (g.TR > 32 && (size_t)g.TR * k * sizeof(double) > 52 * 1024) g.TR >>= 1;
    g.stage_doubles = (size_t)g.TR * k;
    g.stages = (int)std::min<size_t>(AP_MAXSTAGES, (size_t)(200 * 1024) / (g.stage_doubles * sizeof(double)));
    g.ntiles = (nloc + g.TR - 1) / g.TR;
    g.grid = (int)std::min<long long>(g.ntiles, sm_count());
    return g;
}

// Loads rows [row0, row0 + nrows) of V into `dst` and arms `bar`.  The bulk copy moves multiples of 16 bytes; when
// nrows * k is odd (only the last tile of an odd-k filter) the final double is copied by this thread before it arms
// the barrier (its arrive has release semantics for the waiting threads).
__device__ __forceinline__ void issue_tile(double* dst, const double* V, long long row0, int nrows, int k, uint64_t* bar) {
    const size_t n = (size_t)nrows * k;
    const double* src = V + (size_t)row0 * k;
    const unsigned bytes = (unsigned)((n & ~(size_t)1) * sizeof(double));
    if (n & 1) dst[n - 1] = src[n - 1];
    mbar_expect_tx(bar, bytes);
    if (bytes) tma_load_1d(dst, src, bytes, bar);
}

// Z source of the V^T z pass: 8-bit channel, double channel, or the L channel of an interleaved 8-bit BGR image.
enum ZMode { Z_U8 = 0, Z_F64 = 1, Z_BGR = 2 };

// partial[blockIdx.x][v] = sum over this CTA's rows of V[row][v] * z[row]   (fixed order: deterministic)
template <int NV>
__global__ void __launch_bounds__(AP_THREADS, 1)
vtz_tma_kernel(long long nloc, int k, const double* __restrict__ V, const uint8_t* __restrict__ z8, const double* __restrict__ z64,
               int zmode, ApplyGeom geo, double* __restrict__ partial) {
    extern __shared__ __align__(128) unsigned char apsm[];
    double* ring = reinterpret_cast<double*>(apsm);
    double* zs = ring + (size_t)geo.stages * geo.stage_doubles;            // stages x TR
    uint64_t* bars = reinterpret_cast<uint64_t*>(zs + (size_t)geo.stages * geo.TR);
    double* red = reinterpret_cast<double*>(bars + AP_MAXSTAGES);          // AP_THREADS * NV
    LabTables* T = reinterpret_cast<LabTables*>(red + AP_THREADS * NV);
    const int tid = threadIdx.x, TR = geo.TR, S = geo.stages;
    const long long G = gridDim.x;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(bars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (zmode == Z_BGR) lab_tables_load(*T, true, false, tid, AP_THREADS);
    __syncthreads();
    auto tile_rows = [&](long long tile) { return (int)min((long long)TR, nloc - tile * TR); };
    if (tid == 0)
        for (int s = 0; s < S; ++s) {
            const long long tile = blockIdx.x + s * G;
            if (tile < geo.ntiles) issue_tile(ring + (size_t)s * geo.stage_doubles, V, tile * TR, tile_rows(tile), k, bars + s);
        }
    // thread -> (column, row group): k <= 256: ng = 256 / k groups of k threads; k > 256: one group, NV columns per thread
    const int ng = (k <= AP_THREADS) ? AP_THREADS / k : 1;
    const int grp = (k <= AP_THREADS) ? tid / k : 0;
    const int v0 = (k <= AP_THREADS) ? tid - grp * k : tid;
    const bool active = grp < ng;
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.0;
    long long n = 0;
    for (long long tile = blockIdx.x; tile < geo.ntiles; tile += G, ++n) {
        const int s = (int)(n % S);
        const unsigned parity = (unsigned)((n / S) & 1);
        const int nrows = tile_rows(tile);
        const long long row0 = tile * TR;
        double* zt = zs + (size_t)s * TR;
        for (int r = tid; r < nrows; r += AP_THREADS) {
            double z;
            if (zmode == Z_U8) z = (double)z8[row0 + r];
            else if (zmode == Z_F64) z = z64[row0 + r];
            else {
                const uint8_t* px = z8 + 3 * (row0 + r);
                int l, a, b;
                bgr_to_lab_px(*T, px[0], px[1], px[2], l, a, b);                  // filter.cpp:422-426
                z = (double)l;
            }
            zt[r] = z;
        }
        __syncthreads();
        mbar_wait(bars + s, parity);
        const double* tileV = ring + (size_t)s * geo.stage_doubles;
        if (active) {
            // rows grp, grp + ng, ...: four rows in flight per step (independent accumulators, fixed combination order)
            double a1[NV], a2[NV], a3[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) a1[i] = a2[i] = a3[i] = 0.0;
            int r = grp;
            for (; r + 3 * ng < nrows; r += 4 * ng) {
                const double z0 = zt[r], z1 = zt[r + ng], z2 = zt[r + 2 * ng], z3 = zt[r + 3 * ng];
                const double* row = tileV + (size_t)r * k;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const int v = v0 + i * AP_THREADS;
                    if (NV == 1 || v < k) {
                        acc[i] = fma(row[v], z0, acc[i]);
                        a1[i] = fma(row[(size_t)ng * k + v], z1, a1[i]);
                        a2[i] = fma(row[(size_t)2 * ng * k + v], z2, a2[i]);
                        a3[i] = fma(row[(size_t)3 * ng * k + v], z3, a3[i]);
                    }
                }
            }
            for (; r < nrows; r += ng) {
                const double z = zt[r];
                const double* row = tileV + (size_t)r * k;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const int v = v0 + i * AP_THREADS;
                    if (NV == 1 || v < k) acc[i] = fma(row[v], z, acc[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < NV; ++i) acc[i] += (a1[i] + a2[i]) + a3[i];
        }
        __syncthreads();                       // every thread is done with stage s (V tile and z tile)
        const long long next = tile + (long long)S * G;
        if (tid == 0 && next < geo.ntiles) issue_tile(ring + (size_t)s * geo.stage_doubles, V, next * TR, tile_rows(next), k, bars + s);
    }
    // groups -> one value per column, ascending group order
#pragma unroll
    for (int i = 0; i < NV; ++i) red[i * AP_THREADS + tid] = acc[i];
    __syncthreads();
    for (int v = tid; v < k; v += AP_THREADS) {
        double sum = 0.0;
        if (k <= AP_THREADS) {
            for (int q = 0; q < ng; ++q) sum += red[q * k + v];
        } else {
            sum = red[(v / AP_THREADS) * AP_THREADS + (v % AP_THREADS)];
        }
        partial[(size_t)blockIdx.x * k + v] = sum;
    }
}

// t[v] = sum_b partial[b][v] (ascending b), then g[v] = fS[v] * t[v] if fS != nullptr (the all-reduce of a sharded filter
// sits between the two, so the sharded path calls this kernel twice: first with fS == nullptr, then on t alone).
__global__ void vtz_final_kernel(const double* __restrict__ partial, int nblocks, int k, const double* __restrict__ fS,
                                 double* __restrict__ tout, double* __restrict__ gout) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= k) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * k + v];
    if (tout) tout[v] = s;
    if (gout) gout[v] = fS[v] * s;
}

// Output of the recompose pass.
enum OutMode { OUT_F64 = 0, OUT_U8 = 1, OUT_BGR = 2 };

__global__ void __launch_bounds__(AP_THREADS, 1)
recompose_tma_kernel(long long nloc, int k, const double* __restrict__ V, const double* __restrict__ gvec, int omode,
                     double* __restrict__ out64, uint8_t* __restrict__ out8, const uint8_t* __restrict__ bgr_in, ApplyGeom geo) {
    extern __shared__ __align__(128) unsigned char apsm[];
    double* ring = reinterpret_cast<double*>(apsm);
    double* gs = ring + (size_t)geo.stages * geo.stage_doubles;            // k (+ pad)
    uint64_t* bars = reinterpret_cast<uint64_t*>(gs + ((k + 1) & ~1));
    double* part = reinterpret_cast<double*>(bars + AP_MAXSTAGES);         // AP_THREADS
    LabTables* T = reinterpret_cast<LabTables*>(part + AP_THREADS);
    const int tid = threadIdx.x, TR = geo.TR, S = geo.stages;
    const long long G = gridDim.x;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(bars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int v = tid; v < k; v += AP_THREADS) gs[v] = gvec[v];
    if (omode == OUT_BGR) lab_tables_load(*T, true, true, tid, AP_THREADS);
    __syncthreads();
    auto tile_rows = [&](long long tile) { return (int)min((long long)TR, nloc - tile * TR); };
    // descending tile order: the rows the V^T z pass touched last are the ones most likely still in L2
    const long long last = geo.ntiles - 1 - blockIdx.x;
    if (tid == 0)
        for (int s = 0; s < S; ++s) {
            const long long tile = last - s * G;
            if (tile >= 0) issue_tile(ring + (size_t)s * geo.stage_doubles, V, tile * TR, tile_rows(tile), k, bars + s);
        }
    // thread -> (row, column part): nparts = 256 / TR threads share a row and split its columns
    const int nparts = AP_THREADS / TR;
    const int rloc = tid % TR, prt = tid / TR;
    const int rot = (k & 1) ? 2 : 1;            // (k + rot) odd: the 16 lanes of a half-warp hit 16 different bank pairs
    long long n = 0;
    for (long long tile = last; tile >= 0; tile -= G, ++n) {
        const int s = (int)(n % S);
        const unsigned parity = (unsigned)((n / S) & 1);
        const int nrows = tile_rows(tile);
        mbar_wait(bars + s, parity);
        const double* row = ring + (size_t)s * geo.stage_doubles + (size_t)rloc * k;
        double acc = 0.0;
        if (rloc < nrows) {
            // this thread's columns: (rot * rloc + prt + m * nparts) mod k, m = 0, 1, ...; four partial sums in flight
            double b1 = 0.0, b2 = 0.0, b3 = 0.0;
            int v = (rot * rloc + prt) % k;
            auto step = [&](int& vv) { const int cur = vv; vv += nparts; if (vv >= k) vv -= k; return cur; };
            int i = prt;
            for (; i + 3 * nparts < k; i += 4 * nparts) {
                const int c0 = step(v), c1 = step(v), c2 = step(v), c3 = step(v);
                acc = fma(row[c0], gs[c0], acc);
                b1 = fma(row[c1], gs[c1], b1);
                b2 = fma(row[c2], gs[c2], b2);
                b3 = fma(row[c3], gs[c3], b3);
            }
            for (; i < k; i += nparts) {
                const int c0 = step(v);
                acc = fma(row[c0], gs[c0], acc);
            }
            acc += (b1 + b2) + b3;
        }
        if (nparts > 1) {
            part[tid] = acc;
            __syncthreads();
            if (prt == 0)
                for (int q = 1; q < nparts; ++q) acc += part[q * TR + rloc];
        }
        if (prt == 0 && rloc < nrows) {
            const long long j = tile * TR + rloc;
            if (omode == OUT_F64) {
                out64[j] = acc;
            } else {
                // cv::max(.,0), cv::min(.,255), convertTo(CV_8U) = cvRound = round half to even (filter.cpp:434-436)
                const int lq = __double2int_rn(fmin(fmax(acc, 0.0), 255.0));
                if (omode == OUT_U8) {
                    out8[j] = (uint8_t)lq;
                } else {
                    const uint8_t* px = bgr_in + 3 * j;
                    int l0, a, b;
                    bgr_to_lab_px(*T, px[0], px[1], px[2], l0, a, b);             // the untouched a, b (filter.cpp:438)
                    lab_to_bgr_px(*T, lq, a, b, out8 + 3 * j);                    // filter.cpp:440
                }
            }
        }
        __syncthreads();                       // every thread is done with stage s
        const long long next = tile - (long long)S * G;
        if (tid == 0 && next >= 0) issue_tile(ring + (size_t)s * geo.stage_doubles, V, next * TR, tile_rows(next), k, bars + s);
    }
}

// ---- plain-load fall-back of the two passes for k > AP_KMAX (a V row no longer fits the ring) ---------------------
constexpr int AP_PIX = 1024;
__global__ void __launch_bounds__(256)
vtz_wide_kernel(long long nloc, int k, const double* __restrict__ V, const uint8_t* __restrict__ z8,
                const double* __restrict__ z64, double* __restrict__ partial) {
    extern __shared__ double red[];   // 8 * kpad
    const int vl = threadIdx.x & 31, pl = threadIdx.x >> 5;
    const long long base = (long long)blockIdx.x * AP_PIX;
    const int kpad = ((k + 31) / 32) * 32;
    for (int vb = 0; vb < k; vb += 32) {
        const int v = vb + vl;
        double acc = 0.0;
        if (v < k)
            for (int q = pl; q < AP_PIX; q += 8) {
                const long long j = base + q;
                if (j >= nloc) break;
                acc = fma(V[(size_t)j * k + v], z8 ? (double)z8[j] : z64[j], acc);
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
__global__ void __launch_bounds__(256)
recompose_wide_kernel(long long nloc, int k, const double* __restrict__ V, const double* __restrict__ g,
                      double* __restrict__ out64, uint8_t* __restrict__ out8) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * 256) >> 5;
    for (long long j = warp; j < nloc; j += nwarps) {
        const double* vr = V + (size_t)j * k;
        double acc = 0.0;
        for (int v = lane; v < k; v += 32) acc = fma(vr[v], g[v], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            if (out64) out64[j] = acc;
            if (out8) out8[j] = (uint8_t)__double2int_rn(fmin(fmax(acc, 0.0), 255.0));
        }
    }
}

size_t vtz_smem(const ApplyGeom& g, int nv) {
    return ((size_t)g.stages * g.stage_doubles + (size_t)g.stages * g.TR + AP_MAXSTAGES + (size_t)AP_THREADS * nv) * sizeof(double) +
           sizeof(LabTables) + 16;
}
size_t recompose_smem(const ApplyGeom& g, int k) {
    return ((size_t)g.stages * g.stage_doubles + ((k + 1) & ~1) + AP_MAXSTAGES + AP_THREADS) * sizeof(double) + sizeof(LabTables) + 16;
}

}  // namespace

bool apply_tma_supported(int k) { return k >= 1 && k <= AP_KMAX; }

int apply_blocks(long long nloc, int k) {
    if (!apply_tma_supported(k)) return cdiv(nloc, AP_PIX);
    return apply_geometry(nloc, k).grid;
}

// t (and, if fS != nullptr, g = fS o t) from V^T z.  Exactly one of z_u8 / z_f64 / z_bgr is non-null.
void launch_vtz(long long nloc, int k, const double* V, const uint8_t* z_u8, const double* z_f64, const uint8_t* z_bgr,
                const double* fS, double* scratch, double* t_out, double* g_out, cudaStream_t s) {
    if (nloc <= 0 || k <= 0) return;
    int nb;
    if (apply_tma_supported(k)) {
        const ApplyGeom g = apply_geometry(nloc, k);
        nb = g.grid;
        const int nv = k <= AP_THREADS ? 1 : 2;                 // k <= AP_KMAX = 400 < 2 * AP_THREADS
        const size_t smem = vtz_smem(g, nv);
        const int zmode = z_bgr ? Z_BGR : (z_u8 ? Z_U8 : Z_F64);
        const uint8_t* z8 = z_bgr ? z_bgr : z_u8;
        auto go = [&](auto kern) {
            allow_max_dynamic_smem((const void*)kern);
            kern<<<g.grid, AP_THREADS, smem, s>>>(nloc, k, V, z8, z_f64, zmode, g, scratch);
        };
        if (nv == 1) go(vtz_tma_kernel<1>); else go(vtz_tma_kernel<2>);
        NLE_LAUNCH_CHECK();
    } else {
        if (z_bgr) throw Unsupported{"fused BGR apply needs k <= " + std::to_string(AP_KMAX)};
        nb = cdiv(nloc, AP_PIX);
        const int kpad = ((k + 31) / 32) * 32;
        vtz_wide_kernel<<<nb, 256, (size_t)8 * kpad * sizeof(double), s>>>(nloc, k, V, z_u8, z_f64, scratch);
        NLE_LAUNCH_CHECK();
    }
    vtz_final_kernel<<<cdiv(k, 64), 64, 0, s>>>(scratch, nb, k, fS, t_out, fS ? g_out : nullptr);
    NLE_LAUNCH_CHECK();
}

// g = fS o t on its own (after the all-reduce of t in the sharded path)
void launch_scale_t(int k, const double* t, const double* fS, double* g, cudaStream_t s) {
    vtz_final_kernel<<<cdiv(k, 64), 64, 0, s>>>(t, 1, k, fS, nullptr, g);
    NLE_LAUNCH_CHECK();
}

// out = V g with the epilogue selected by the non-null output: out_f64 | out_u8 (clamp + round) | out_bgr (clamp + round +
// Lab2BGR with the a, b of bgr_in).
void launch_recompose(long long nloc, int k, const double* V, const double* g, double* out_f64, uint8_t* out_u8,
                      uint8_t* out_bgr, const uint8_t* bgr_in, cudaStream_t s) {
    if (nloc <= 0 || k <= 0) return;
    if (apply_tma_supported(k)) {
        const ApplyGeom geo = apply_geometry(nloc, k);
        const size_t smem = recompose_smem(geo, k);
        const int omode = out_bgr ? OUT_BGR : (out_u8 ? OUT_U8 : OUT_F64);
        allow_max_dynamic_smem((const void*)recompose_tma_kernel);
        recompose_tma_kernel<<<geo.grid, AP_THREADS, smem, s>>>(nloc, k, V, g, omode, out_f64, out_bgr ? out_bgr : out_u8, bgr_in, geo);
        NLE_LAUNCH_CHECK();
    } else {
        if (out_bgr) throw Unsupported{"fused BGR apply needs k <= " + std::to_string(AP_KMAX)};
        const int grid = (int)std::min<long long>((nloc + 7) / 8, (long long)sm_count() * 16);
        recompose_wide_kernel<<<std::max(grid, 1), 256, 0, s>>>(nloc, k, V, g, out_f64, out_u8);
        NLE_LAUNCH_CHECK();
    }
}

void launch_bgr2lab(const uint8_t* bgr, long long npix, uint8_t* L, uint8_t* ab, cudaStream_t s) {
    if (npix <= 0) return;
    const int grid = (int)std::min<long long>((npix + 255) / 256, (long long)sm_count() * 8);
    bgr2lab_kernel<<<grid, 256, 0, s>>>(bgr, npix, L, ab);
    NLE_LAUNCH_CHECK();
}

void launch_lab2bgr(const uint8_t* L, const uint8_t* ab, long long npix, uint8_t* bgr, cudaStream_t s) {
    if (npix <= 0) return;
    const int grid = (int)std::min<long long>((npix + 255) / 256, (long long)sm_count() * 8);
    lab2bgr_kernel<<<grid, 256, 0, s>>>(L, ab, npix, bgr);
    NLE_LAUNCH_CHECK();
}

}  // namespace nle
