// 8-bit BGR <-> CIE Lab on the device, byte-exact with cv::cvtColor(COLOR_BGR2Lab / COLOR_Lab2BGR) on CV_8UC3 --
// the conversions NLEFilter::trainForEnhancement / enhance perform around the spectral filter
// (reference filter.cpp:422-426, 438-440, 463-466).  OpenCV is an un-vendored, unpinned dependency of the reference
// (CMakeLists.txt:34); its 8-bit Lab path is fixed-point arithmetic over small tables (imgproc color_lab.cpp:
// RGB2Lab_b, Lab2RGBinteger).  scripts/make_lab_tables.py restates the tables, verifies the whole integer pipeline
// against cv2 on all 2^24 BGR and all 2^24 Lab triples, and writes lab_tables.inc; tests/test_gpu_lab.py repeats the
// exhaustive comparison through these kernels.
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

// bgr: npix x 3 interleaved.  Writes L (npix) and ab (npix x 2 interleaved); either may alias nothing else.
__global__ void __launch_bounds__(256)
bgr2lab_kernel(const uint8_t* __restrict__ bgr, long long npix, uint8_t* __restrict__ L, uint8_t* __restrict__ ab) {
    __shared__ unsigned short gam[kGammaN];
    __shared__ unsigned short cbr[kCbrtN];
    for (int i = threadIdx.x; i < kGammaN; i += 256) gam[i] = kLabGammaTab[i];
    for (int i = threadIdx.x; i < kCbrtN; i += 256) cbr[i] = kLabCbrtTab[i];
    __syncthreads();
    const int C0 = kLabFwdCoef[0], C1 = kLabFwdCoef[1], C2 = kLabFwdCoef[2], C3 = kLabFwdCoef[3], C4 = kLabFwdCoef[4],
              C5 = kLabFwdCoef[5], C6 = kLabFwdCoef[6], C7 = kLabFwdCoef[7], C8 = kLabFwdCoef[8];
    const int Lscale = (116 * 255 + 50) / 100;
    const int Lshift = -((16 * 255 * (1 << kLabShift2) + 50) / 100);
    for (long long j = (long long)blockIdx.x * 256 + threadIdx.x; j < npix; j += (long long)gridDim.x * 256) {
        const int B = gam[bgr[3 * j]], G = gam[bgr[3 * j + 1]], R = gam[bgr[3 * j + 2]];
        const int fX = cbr[descale(R * C0 + G * C1 + B * C2, kLabShift)];
        const int fY = cbr[descale(R * C3 + G * C4 + B * C5, kLabShift)];
        const int fZ = cbr[descale(R * C6 + G * C7 + B * C8, kLabShift)];
        L[j] = (uint8_t)sat_u8(descale(Lscale * fY + Lshift, kLabShift2));
        if (ab) {
            ab[2 * j] = (uint8_t)sat_u8(descale(500 * (fX - fY) + 128 * (1 << kLabShift2), kLabShift2));
            ab[2 * j + 1] = (uint8_t)sat_u8(descale(200 * (fY - fZ) + 128 * (1 << kLabShift2), kLabShift2));
        }
    }
}

__global__ void __launch_bounds__(256)
lab2bgr_kernel(const uint8_t* __restrict__ L, const uint8_t* __restrict__ ab, long long npix, uint8_t* __restrict__ bgr) {
    __shared__ unsigned short ytab[256], fytab[256];
    __shared__ unsigned char ig[kInvGammaN];
    for (int i = threadIdx.x; i < 256; i += 256) { ytab[i] = kLabYTab[i]; fytab[i] = kLabFyTab[i]; }
    for (int i = threadIdx.x; i < kInvGammaN; i += 256) ig[i] = kLabInvGammaTab[i];
    __syncthreads();
    int C[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) C[q] = kLabInvCoef[q];
    for (long long j = (long long)blockIdx.x * 256 + threadIdx.x; j < npix; j += (long long)gridDim.x * 256) {
        const int l = L[j], a = ab[2 * j], b = ab[2 * j + 1];
        const int y = ytab[l], ify = fytab[l];
        const int adiv = ((5 * a * 53687 + (1 << 7)) >> 13) - 128 * kBase / 500;
        const int bdiv = ((b * 41943 + (1 << 4)) >> 9) - 128 * kBase / 200 + 1;
        const int x = ab_to_xz(ify + adiv), z = ab_to_xz(ify - bdiv);
        int rgb[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            // |C x| < 2^14 * 2^17: the sum needs 64 bits before the shift
            const long long acc = (long long)C[3 * r] * x + (long long)C[3 * r + 1] * y + (long long)C[3 * r + 2] * z;
            int v = (int)((acc + (1 << 13)) >> 14);
            v = v < 0 ? 0 : (v > kInvGammaN - 1 ? kInvGammaN - 1 : v);
            rgb[r] = ig[v];
        }
        bgr[3 * j] = (uint8_t)rgb[2];
        bgr[3 * j + 1] = (uint8_t)rgb[1];
        bgr[3 * j + 2] = (uint8_t)rgb[0];
    }
}

}  // namespace

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
