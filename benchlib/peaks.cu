// benchlib/peaks.cu -- measurement-only microbenchmarks (libsmb_peaks.so), loaded by bench.py and the sweep
// scripts through ctypes.  NOT part of libstereo_b200.so and not declared in include/stereo_b200.h.
//
// INT32 issue-rate microbenchmark: the roofline denominator of the hot path.
//
// MEASURED_PEAKS.json carries HBM GB/s and bf16 TFLOP/s only; this path is bound by the
// integer ALU issue rate, so bench.py measures that peak on the box with
// this kernel: long chains of independent 3-input integer instructions, operands rotated
// so that ptxas can neither fold nor strength-reduce them.
//   mode 0: IADD3            (alu pipe)
//   mode 1: LOP3             (alu pipe)
//   mode 2: IADD3 + IMAD     (alu + fma pipes together -- the dual-issue ceiling)
//   mode 3: LOP3 + IMAD
// Result: thread-instructions per second (one instruction = one "integer op").
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace {

#define PK_CUDA(call)                                                                                  \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            fprintf(stderr, "%s:%d: %s -> %s\n", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return -2;                                                                                 \
        }                                                                                              \
    } while (0)


constexpr int PEAK_REGS = 8;
constexpr int PEAK_UNROLL = 16;  // rounds per loop iteration; one round = PEAK_REGS instructions

template <int MODE>
__global__ void __launch_bounds__(256) k_int_peak(uint32_t *out, int iters, uint32_t seed)
{
    uint32_t a[PEAK_REGS];
#pragma unroll
    for (int k = 0; k < PEAK_REGS; k++) a[k] = seed * (k + 1) + threadIdx.x * 2654435761u + blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < PEAK_UNROLL; u++) {
#pragma unroll
            for (int k = 0; k < PEAK_REGS; k++) {
                uint32_t x = a[k], y = a[(k + 1) % PEAK_REGS], z = a[(k + 3) % PEAK_REGS];
                bool second = (k & 1) != 0;
                if (MODE == 0 || (MODE == 2 && !second)) {
                    asm volatile("{ .reg .u32 t; add.u32 t, %1, %2; add.u32 %0, t, %3; }"
                                 : "=r"(a[k]) : "r"(x), "r"(y), "r"(z));
                } else if (MODE == 1 || (MODE == 3 && !second)) {
                    asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a[k]) : "r"(x), "r"(y), "r"(z));
                } else {
                    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a[k]) : "r"(y), "r"(z), "r"(x));
                }
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < PEAK_REGS; k++) r ^= a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

}  // namespace

extern "C" int smb_measure_int_peak(int device, int mode, double *gops_per_s)
{
    if (!gops_per_s || mode < 0 || mode > 3) return -1;
    int prev = -1;
    cudaGetDevice(&prev);
    PK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PK_CUDA(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 2000;
    uint32_t *out = nullptr;
    PK_CUDA(cudaMalloc(&out, (size_t)blocks * threads * sizeof(uint32_t)));
    cudaEvent_t e0, e1;
    PK_CUDA(cudaEventCreate(&e0));
    PK_CUDA(cudaEventCreate(&e1));
    float best_ms = 1e30f;
    for (int rep = 0; rep < 5; rep++) {  // rep 0 is the warm-up
        PK_CUDA(cudaEventRecord(e0));
        switch (mode) {
        case 0: k_int_peak<0><<<blocks, threads>>>(out, iters, 17u + rep); break;
        case 1: k_int_peak<1><<<blocks, threads>>>(out, iters, 17u + rep); break;
        case 2: k_int_peak<2><<<blocks, threads>>>(out, iters, 17u + rep); break;
        default: k_int_peak<3><<<blocks, threads>>>(out, iters, 17u + rep); break;
        }
        PK_CUDA(cudaEventRecord(e1));
        PK_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        PK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    double instr = (double)blocks * threads * iters * PEAK_UNROLL * PEAK_REGS;
    *gops_per_s = instr / (best_ms * 1e-3) / 1e9;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (prev >= 0) cudaSetDevice(prev);
    return 0;
}

// Host <-> device copy bandwidth of this GPU's link from pinned memory: the ceiling of the
// end-to-end (host buffers) figure.  mode 0: H2D alone, 1: D2H alone, 2: both directions at
// once (two streams), reported per direction (H2D in gbs[0], D2H in gbs[1]).
// Three calls so that several ranks can measure AT THE SAME TIME (set up, barrier in the caller, measure):
// with eight GPUs on one host the per-GPU figure under contention is what bounds the end-to-end run.
namespace {
struct CopyPeak {
    int device = -1, prev = -1;
    void *h[2] = {nullptr, nullptr}, *d[2] = {nullptr, nullptr};
    cudaStream_t st[2] = {nullptr, nullptr};
    cudaEvent_t e0[2] = {nullptr, nullptr}, e1[2] = {nullptr, nullptr};
} g_cp;
const size_t kCopyBytes = (size_t)256 << 20;
}  // namespace

extern "C" int smb_copy_peak_setup(int device)
{
    if (g_cp.device >= 0) return -1;
    cudaGetDevice(&g_cp.prev);
    PK_CUDA(cudaSetDevice(device));
    for (int k = 0; k < 2; k++) {
        PK_CUDA(cudaHostAlloc(&g_cp.h[k], kCopyBytes, cudaHostAllocDefault));
        memset(g_cp.h[k], k + 1, kCopyBytes);
        PK_CUDA(cudaMalloc(&g_cp.d[k], kCopyBytes));
        PK_CUDA(cudaStreamCreateWithFlags(&g_cp.st[k], cudaStreamNonBlocking));
        PK_CUDA(cudaEventCreate(&g_cp.e0[k]));
        PK_CUDA(cudaEventCreate(&g_cp.e1[k]));
    }
    g_cp.device = device;
    return 0;
}

// reps timed repetitions after one warm-up; best != 0: the fastest repetition, else the mean rate over all of them
extern "C" int smb_copy_peak_measure(int mode, int reps, int best, double *gbs)
{
    if (!gbs || mode < 0 || mode > 2 || reps < 1 || g_cp.device < 0) return -1;
    gbs[0] = gbs[1] = 0.0;
    double total_ms[2] = {0.0, 0.0};
    for (int rep = 0; rep <= reps; rep++) {  // rep 0 is the warm-up
        for (int k = 0; k < 2; k++) {
            if (mode != 2 && k != mode) continue;
            PK_CUDA(cudaEventRecord(g_cp.e0[k], g_cp.st[k]));
            if (k == 0)
                PK_CUDA(cudaMemcpyAsync(g_cp.d[0], g_cp.h[0], kCopyBytes, cudaMemcpyHostToDevice, g_cp.st[0]));
            else
                PK_CUDA(cudaMemcpyAsync(g_cp.h[1], g_cp.d[1], kCopyBytes, cudaMemcpyDeviceToHost, g_cp.st[1]));
            PK_CUDA(cudaEventRecord(g_cp.e1[k], g_cp.st[k]));
        }
        for (int k = 0; k < 2; k++) {
            if (mode != 2 && k != mode) continue;
            PK_CUDA(cudaEventSynchronize(g_cp.e1[k]));
            float ms = 0;
            PK_CUDA(cudaEventElapsedTime(&ms, g_cp.e0[k], g_cp.e1[k]));
            if (rep == 0) continue;
            total_ms[k] += ms;
            const double g = (double)kCopyBytes / (ms * 1e-3) / 1e9;
            if (best && g > gbs[k]) gbs[k] = g;
        }
    }
    if (!best)
        for (int k = 0; k < 2; k++)
            if (total_ms[k] > 0) gbs[k] = (double)kCopyBytes * reps / (total_ms[k] * 1e-3) / 1e9;
    return 0;
}

// The link under the traffic MIX of a call: `reps` rounds of h2d_bytes up and d2h_bytes down (each at most 256 MB),
// the two directions on their own streams and free-running; seconds per round by the host clock.
extern "C" int smb_copy_peak_mix(size_t h2d_bytes, size_t d2h_bytes, int reps, double *seconds)
{
    if (!seconds || reps < 1 || g_cp.device < 0 || h2d_bytes > kCopyBytes || d2h_bytes > kCopyBytes) return -1;
    for (int rep = -1; rep < reps; rep++) {  // rep -1 is the warm-up
        if (rep == 0) {
            PK_CUDA(cudaStreamSynchronize(g_cp.st[0]));
            PK_CUDA(cudaStreamSynchronize(g_cp.st[1]));
            PK_CUDA(cudaEventRecord(g_cp.e0[0], g_cp.st[0]));
            PK_CUDA(cudaEventRecord(g_cp.e0[1], g_cp.st[1]));
        }
        if (h2d_bytes) PK_CUDA(cudaMemcpyAsync(g_cp.d[0], g_cp.h[0], h2d_bytes, cudaMemcpyHostToDevice, g_cp.st[0]));
        if (d2h_bytes) PK_CUDA(cudaMemcpyAsync(g_cp.h[1], g_cp.d[1], d2h_bytes, cudaMemcpyDeviceToHost, g_cp.st[1]));
    }
    PK_CUDA(cudaEventRecord(g_cp.e1[0], g_cp.st[0]));
    PK_CUDA(cudaEventRecord(g_cp.e1[1], g_cp.st[1]));
    float ms0 = 0, ms1 = 0;
    PK_CUDA(cudaEventSynchronize(g_cp.e1[0]));
    PK_CUDA(cudaEventSynchronize(g_cp.e1[1]));
    PK_CUDA(cudaEventElapsedTime(&ms0, g_cp.e0[0], g_cp.e1[0]));
    PK_CUDA(cudaEventElapsedTime(&ms1, g_cp.e0[1], g_cp.e1[1]));
    *seconds = (double)(ms0 > ms1 ? ms0 : ms1) * 1e-3 / reps;
    return 0;
}

extern "C" int smb_copy_peak_teardown(void)
{
    if (g_cp.device < 0) return 0;
    for (int k = 0; k < 2; k++) {
        cudaEventDestroy(g_cp.e0[k]);
        cudaEventDestroy(g_cp.e1[k]);
        cudaStreamDestroy(g_cp.st[k]);
        cudaFree(g_cp.d[k]);
        cudaFreeHost(g_cp.h[k]);
    }
    if (g_cp.prev >= 0) cudaSetDevice(g_cp.prev);
    g_cp = CopyPeak();
    return 0;
}

extern "C" int smb_measure_copy_peak(int device, int mode, double *gbs)
{
    if (!gbs || mode < 0 || mode > 2) return -1;
    if (smb_copy_peak_setup(device)) return -1;
    int rc = smb_copy_peak_measure(mode, 3, 1, gbs);
    smb_copy_peak_teardown();
    return rc;
}
