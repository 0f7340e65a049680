// k_edges.cu -- step 1, find_all_edges on the device.
//
// Replaces find_all_edges<<<>>> of the reference (stereo.cu:83-92 with the four
// detectors stereo.cu:17-81; ghost twin stereo-ghost.cu:84-93).  One thread per pixel,
// FP64 with the reference's exact operation order -- ((a+b)+c)/3.0, (l+r)/2.0,
// fabs(l-r) > min(max(thr*overall,0),1) (stereo.c:16-28) -- written with the
// round-to-nearest intrinsics so that nothing is contracted or reassociated.
// Input is either the 8-bit pixels (converted as image.c:13 does, v/256.0, exact) or
// the reference's own double layout.
#include "sm_common.cuh"

namespace smb {

template <typename T>
__device__ __forceinline__ double to_bright(T v);
template <>
__device__ __forceinline__ double to_bright<uint8_t>(uint8_t v)
{
    return __ddiv_rn((double)v, 256.0);
}
template <>
__device__ __forceinline__ double to_bright<double>(double v)
{
    return v;
}

__device__ __forceinline__ int detect(double a0, double a1, double a2, double b0, double b1,
                                      double b2, double thr)
{
    double l = __ddiv_rn(__dadd_rn(__dadd_rn(a0, a1), a2), 3.0);
    double r = __ddiv_rn(__dadd_rn(__dadd_rn(b0, b1), b2), 3.0);
    double ov = __ddiv_rn(__dadd_rn(l, r), 2.0);
    double lim = __dmul_rn(thr, ov);
    lim = lim > 0.0 ? lim : 0.0;  // CLAMP = MIN(MAX(x, 0), 1), util.h:24-26
    lim = lim < 1.0 ? lim : 1.0;
    return fabs(__dsub_rn(l, r)) > lim;
}

// Rows handled: frame rows ystart .. ystart+nrows-1.  WRAP: taken mod FH (a band's halo
// rows wrap around the frame); GHOST: rows outside the frame are skipped (their edge
// cells are ghost zeros that the pack kernel supplies).
template <typename T, int VARIANT>
__global__ void __launch_bounds__(256) k_edges(const T *__restrict__ img, int W, int FH, int ystart,
                                               int nrows, double thr, uint8_t *__restrict__ edges)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int r = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || r >= nrows) return;
    int y = ystart + r;
    if (VARIANT == SM_WRAP) {
        y %= FH;
        if (y < 0) y += FH;
    } else if (y < 0 || y >= FH) {
        return;
    }
    double b[3][3];
#pragma unroll
    for (int dy = -1; dy <= 1; dy++) {
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
            int xx = x + dx, yy = y + dy;
            double v;
            if (VARIANT == SM_WRAP) {
                xx = xx < 0 ? xx + W : (xx >= W ? xx - W : xx);
                yy = yy < 0 ? yy + FH : (yy >= FH ? yy - FH : yy);
                v = to_bright<T>(img[(size_t)yy * W + xx]);
            } else {
                // brightness ghost cell = 128.0 (stereo-ghost.c:384-385)
                bool in = xx >= 0 && xx < W && yy >= 0 && yy < FH;
                v = in ? to_bright<T>(img[(size_t)yy * W + xx]) : 128.0;
            }
            b[dy + 1][dx + 1] = v;
        }
    }
#define B(dx, dy) b[(dy) + 1][(dx) + 1]
    int e = detect(B(-1, -1), B(-1, 0), B(-1, 1), B(1, -1), B(1, 0), B(1, 1), thr)       // left_right
            | detect(B(-1, -1), B(0, -1), B(1, -1), B(-1, 1), B(0, 1), B(1, 1), thr)     // top_bottom
            | detect(B(-1, -1), B(0, -1), B(-1, 0), B(1, 0), B(0, 1), B(1, 1), thr)      // upleft_downright
            | detect(B(-1, 1), B(0, 1), B(-1, 0), B(0, -1), B(1, -1), B(1, 0), thr);     // downleft_upright
#undef B
    edges[(size_t)y * W + x] = (uint8_t)e;
}

template <typename T>
int launch_edges(const T *img, int W, int FH, int ystart, int nrows, int variant, double threshold,
                 uint8_t *edges, cudaStream_t s)
{
    dim3 block(64, 4);
    dim3 grid((W + block.x - 1) / block.x, (nrows + block.y - 1) / block.y);
    if (variant == SM_WRAP)
        k_edges<T, SM_WRAP><<<grid, block, 0, s>>>(img, W, FH, ystart, nrows, threshold, edges);
    else
        k_edges<T, SM_GHOST><<<grid, block, 0, s>>>(img, W, FH, ystart, nrows, threshold, edges);
    SM_CUDA(cudaGetLastError());
    return 1;
}

template int launch_edges<uint8_t>(const uint8_t *, int, int, int, int, int, double, uint8_t *,
                                   cudaStream_t);
template int launch_edges<double>(const double *, int, int, int, int, int, double, uint8_t *,
                                  cudaStream_t);

}  // namespace smb
