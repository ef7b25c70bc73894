// Measured ceilings of this GPU for the roofline lines of bench.py (SURVEY.md 8d asks for measured FP64 / FP32 / MUFU /
// shared-memory peaks; MEASURED_PEAKS.json carries HBM copy and dense bf16 only).  Register- or cache-resident
// microbenchmarks, best of five launches each, timed with CUDA events on the launching stream.
//   [0] FP64 FMA (DFMA)  TFLOP/s      [1] FP64 tensor pipe (mma.sync.m8n8k4.f64, DMMA)  TFLOP/s   (api.cu)
//   [2] FP32 FMA  TFLOP/s             [3] MUFU ex2.approx.f32  Gop/s
//   [4] shared-memory loads (16 B per lane, conflict-free)  GB/s
//   [5] L2 reads (32 MB working set, 16 B per lane)  GB/s     [6] HBM copy (1 GiB -> 1 GiB, read + write bytes)  GB/s
//   [7] shared-memory loads in bytes per SM clock (clock64 span of the slowest SM, same kernel as [4])
#include "common.cuh"

#include <algorithm>

namespace nle {

double measure_fp64_peak_tflops();
double measure_dmma_peak_tflops();

namespace {

struct EvTimer {
    cudaEvent_t a, b;
    cudaStream_t s;
    explicit EvTimer(cudaStream_t st) : s(st) { NLE_CUDA(cudaEventCreate(&a)); NLE_CUDA(cudaEventCreate(&b)); NLE_CUDA(cudaEventRecord(a, s)); }
    double stop() {
        float ms = 0.f;
        NLE_CUDA(cudaEventRecord(b, s));
        NLE_CUDA(cudaEventSynchronize(b));
        NLE_CUDA(cudaEventElapsedTime(&ms, a, b));
        return ms;
    }
    ~EvTimer() { cudaEventDestroy(a); cudaEventDestroy(b); }
};

__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters, float a, float b) {
    float x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) x[u] = threadIdx.x + u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int u = 0; u < 8; ++u) x[u] = fmaf(x[u], a, b);
    }
    float r = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) r += x[u];
    if (r == 123.456f) out[0] = r;
}

__global__ void __launch_bounds__(256) mufu_peak_kernel(float* out, int iters) {
    float x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) x[u] = -1.0f - 0.001f * (threadIdx.x + u);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int u = 0; u < 8; ++u) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[u]));
    }
    float r = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) r += x[u];
    if (r == 123.456f) out[0] = r;
}

// One CTA of 1024 threads per SM (128 KB of dynamic shared memory keeps a second one out); every warp reads 512 contiguous
// bytes per instruction (LDS.128, conflict-free) from a window that moves through the buffer.  cyc[b] = clock64 span of block b.
__global__ void __launch_bounds__(1024) smem_peak_kernel(double* out, int iters, long long* cyc) {
    extern __shared__ double2 sbuf[];
    constexpr int kWords = 128 * 1024 / 16;
    for (int e = threadIdx.x; e < kWords; e += 1024) sbuf[e] = make_double2(e, 1.0);
    __syncthreads();
    unsigned long long ax = 0, ay = 0;                                  // integer accumulation: the FP64 pipe stays out of the way
    const unsigned base = (unsigned)__cvta_generic_to_shared(sbuf);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            unsigned long long vx, vy;
            const unsigned addr = base + 16u * ((threadIdx.x + 1024u * (unsigned)(u + i)) & (kWords - 1));
            asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(vx), "=l"(vy) : "r"(addr));   // volatile: not hoisted out of the loop
            ax ^= vx;
            ay ^= vy;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
    if ((ax ^ ay) == 0x123456789abcdefull) out[0] = (double)ax;
}

__global__ void __launch_bounds__(256) l2_read_kernel(const double2* __restrict__ src, size_t count, int passes, double* out) {
    double2 acc = make_double2(0.0, 0.0);
    const size_t stride = (size_t)gridDim.x * 256;
    for (int p = 0; p < passes; ++p)
        for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < count; e += 4 * stride) {
            double2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const size_t ee = e + u * stride;
                v[u] = ee < count ? __ldcg(src + ee) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { acc.x += v[u].x; acc.y += v[u].y; }
        }
    if (acc.x + acc.y == 123.456) out[0] = acc.x;
}

__global__ void __launch_bounds__(256) hbm_copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst, size_t count) {
    const size_t stride = (size_t)gridDim.x * 256;
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < count; e += 4 * stride) {
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const size_t ee = e + u * stride;
            if (ee < count) v[u] = __ldcs(src + ee);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const size_t ee = e + u * stride;
            if (ee < count) __stcs(dst + ee, v[u]);
        }
    }
}

template <typename F>
double best_of(int reps, F&& run) {
    double best = 1e300;
    for (int r = 0; r < reps; ++r) best = std::min(best, run());
    return best;
}

}  // namespace

void measure_peaks(double* out, int n) {
    cudaStream_t s = nullptr;
    double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    v[0] = measure_fp64_peak_tflops();
    v[1] = measure_dmma_peak_tflops();
    DevBuf<float> df(1);
    DevBuf<double> dd(1);
    const int blocks = sm_count() * 8, iters = 2048;
    double ms = best_of(5, [&] { EvTimer t(s); fp32_peak_kernel<<<blocks, 256, 0, s>>>(df.p, iters, 0.999999f, 1e-9f); NLE_LAUNCH_CHECK(); return t.stop(); });
    v[2] = 2.0 * 64.0 * iters * 256.0 * blocks / (ms * 1e-3) * 1e-12;
    ms = best_of(5, [&] { EvTimer t(s); mufu_peak_kernel<<<blocks, 256, 0, s>>>(df.p, iters); NLE_LAUNCH_CHECK(); return t.stop(); });
    v[3] = 64.0 * iters * 256.0 * blocks / (ms * 1e-3) * 1e-9;
    {
        const int nb = sm_count();
        DevBuf<long long> cyc(nb);
        NLE_CUDA(cudaFuncSetAttribute(smem_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
        ms = best_of(5, [&] { EvTimer t(s); smem_peak_kernel<<<nb, 1024, 128 * 1024, s>>>(dd.p, iters, cyc.p); NLE_LAUNCH_CHECK(); return t.stop(); });
        v[4] = 16.0 * 8.0 * iters * 1024.0 * nb / (ms * 1e-3) * 1e-9;
        std::vector<long long> h(nb);
        cyc.download(h.data(), nb, s);
        NLE_CUDA(cudaStreamSynchronize(s));
        long long worst = 1;
        for (long long c : h) worst = std::max(worst, c);
        v[7] = 16.0 * 8.0 * iters * 1024.0 / (double)worst;              // bytes per SM clock of the slowest SM
    }
    {
        const size_t count = ((size_t)32 << 20) / sizeof(double2);
        DevBuf<double2> buf(count);
        NLE_CUDA(cudaMemsetAsync(buf.p, 0, count * sizeof(double2), s));
        const int passes = 32;
        ms = best_of(5, [&] { EvTimer t(s); l2_read_kernel<<<sm_count() * 8, 256, 0, s>>>(buf.p, count, passes, dd.p); NLE_LAUNCH_CHECK(); return t.stop(); });
        v[5] = (double)passes * count * sizeof(double2) / (ms * 1e-3) * 1e-9;
    }
    {
        const size_t count = ((size_t)1 << 30) / sizeof(double2);
        DevBuf<double2> a(count), b(count);
        NLE_CUDA(cudaMemsetAsync(a.p, 0, count * sizeof(double2), s));
        ms = best_of(5, [&] { EvTimer t(s); hbm_copy_kernel<<<sm_count() * 16, 256, 0, s>>>(a.p, b.p, count); NLE_LAUNCH_CHECK(); return t.stop(); });
        v[6] = 2.0 * count * sizeof(double2) / (ms * 1e-3) * 1e-9;
    }
    for (int i = 0; i < n && i < 8; ++i) out[i] = v[i];
}

}  // namespace nle
