// k_bitslice.cu -- the fast hot-path kernel (SM_KERNEL_BITSLICE).
//
// What it replaces: fillup_matches + 30x{memset, addup_pixels_in_square, record_score} +
// find_highest_scoring_shifts of the reference (stereo.cu:127-225), i.e. per pixel and per
// shift a (2*half+1)^2 box sum of the 0/1 match image, masked by the centre match, then the
// arg-max over shifts with ties to the highest shift.
//
// Formulation.  Shifts are the SIMD axis: one 32-bit word holds, for one pixel, the 1-bit
// match flags of 32 consecutive shifts ("lanes").  Counts are bit-sliced: a k-bit count for
// 32 shifts is k words (planes).  For a pixel u of a row the match word is
//        M(u) = (R[u..u+31] & A(u)) | (~R[u..u+31] & B(u))
// where R[..] is a 32-bit funnel-shifted window of the packed right-edge row and A/B are
// the left pixel's LA/LB bits splatted to words (sm_common.cuh).  The box sum is built
// incrementally, exactly as an integer running sum would be (so the counts are identical
// to the reference's direct sum):
//   pass A (horizontal): a walker slides along x keeping H(x) = sum_{|sx|<=half} M(x+sx) as a
//           KH-plane up/down counter: one new word enters, one leaves, 2*KH-1 LOP3.
//   pass B (vertical):   one thread per pixel column keeps V(x,y) = sum_{|sy|<=half} H(x,y+sy)
//           in PV planes; per row it ripple-adds the entering H row and ripple-subtracts
//           the leaving one.
//   WTA: bit-serial max over the 32*NW lanes, from the top plane down:
//           t = cand & V_p & M; if (t != 0) { cand = t; best |= 1 << p; }
//        The survivors are the shifts whose masked score equals the maximum; the highest
//        surviving lane is the reference's tie rule (last i with scores[i] == best,
//        stereo.c:212-219), and with no survivor plane at all (every score 0) cand is still
//        "all shifts", giving web = num_shifts as the reference does.
// Work per pixel x shift is ~2 ALU instructions instead of the reference's (2*half+1)^2 taps.
//
// Decomposition.  A CTA owns a strip of TW = 64 pixel columns and a run of output rows, and
// streams down the rows in blocks of RB rows: stage the block's packed rows into shared
// memory (R words + A/B splat table), pass A by one warp of walkers into a ring of H rows,
// pass B by all threads.  The ring keeps RB + 2*half + 1 rows so that the row leaving the
// vertical window is still there.  More than 32*NW shifts are processed as successive
// chunks over the same rows, merging (best, web) in place (a later chunk holds higher
// shifts, so it wins ties).
#include "sm_common.cuh"

namespace smb {

namespace {

__host__ __device__ constexpr int bits_for(int v)
{
    int b = 0;
    while ((1 << b) <= v) b++;
    return b;
}

template <int HALF, int NW>
struct BS {
    static constexpr int N = 2 * HALF + 1;       // window side
    static constexpr int KH = bits_for(N);       // planes of a horizontal count (<= N)
    static constexpr int PV = bits_for(N * N);   // planes of a box count (<= N*N)
    static constexpr int TW = 64;                // pixel columns per CTA == threads per CTA
    static constexpr int SEG = 32;               // columns per pass-A walker
    static constexpr int NSEG = TW / SEG;
    static constexpr int RB = 32 / (NW * NSEG);  // rows per block: the walkers fill one warp
    static constexpr int NR = RB + N;            // ring rows
    static constexpr int HROW = TW + 1;          // uint4 per (ring row, word); +1 staggers banks
    static constexpr int ABN = TW + 2 * HALF;    // A/B entries per ring row
    static constexpr int ABROW = ABN | 1;        // odd stride: conflict-free LDS.64 across rows
    static constexpr int RW = (TW / 32 + NW + 3) | 1;  // R words per ring row
    static constexpr int H5N = KH > 4 ? ((NR * NW * HROW + 3) & ~3) : 0;  // words, 16-byte multiple
    static constexpr size_t SMEM = (size_t)NR * NW * HROW * 16 + (size_t)H5N * 4 + (size_t)NR * ABROW * 8 +
                                   (size_t)NR * RW * 4;
};

struct BitsliceArgs {
    HotArgs h;
    int rows_per_seg;  // output rows per CTA
};

// V (PV planes) += H (KH planes), ripple carry; the sum always fits PV planes.
template <int PV, int KH>
__device__ __forceinline__ void planes_add(uint32_t (&V)[PV], const uint32_t (&H)[5])
{
    uint32_t c = V[0] & H[0];
    V[0] ^= H[0];
#pragma unroll
    for (int k = 1; k < PV; k++) {
        if (k < KH) {
            uint32_t s = V[k] ^ H[k] ^ c;
            c = (V[k] & H[k]) | (c & (V[k] ^ H[k]));
            V[k] = s;
        } else {
            uint32_t s = V[k] ^ c;
            c = V[k] & c;
            V[k] = s;
        }
    }
}

// V -= H, ripple borrow; the caller guarantees V >= H lane-wise (H was added before).
template <int PV, int KH>
__device__ __forceinline__ void planes_sub(uint32_t (&V)[PV], const uint32_t (&H)[5])
{
    uint32_t b = ~V[0] & H[0];
    V[0] ^= H[0];
#pragma unroll
    for (int k = 1; k < PV; k++) {
        if (k < KH) {
            uint32_t d = V[k] ^ H[k] ^ b;
            b = (~V[k] & (H[k] | b)) | (H[k] & b);
            V[k] = d;
        } else {
            uint32_t d = V[k] ^ b;
            b = ~V[k] & b;
            V[k] = d;
        }
    }
}

template <int HALF, int NW>
__global__ void __launch_bounds__(64) k_bitslice(BitsliceArgs a)
{
    using C = BS<HALF, NW>;
    constexpr int N = C::N, KH = C::KH, PV = C::PV, TW = C::TW, SEG = C::SEG, RB = C::RB, NR = C::NR;
    constexpr int HROW = C::HROW, ABN = C::ABN, ABROW = C::ABROW, RW = C::RW;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4 *Hq = reinterpret_cast<uint4 *>(smem_raw);                              // [NR][NW][HROW]
    uint32_t *H5 = reinterpret_cast<uint32_t *>(Hq + NR * NW * HROW);             // [NR][NW][HROW] (KH == 5)
    uint2 *ABt = reinterpret_cast<uint2 *>(H5 + C::H5N);                          // [NR][ABROW]
    uint32_t *Rs = reinterpret_cast<uint32_t *>(ABt + NR * ABROW);                // [NR][RW]

    const PackedGeom &g = a.h.g;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW;
    const int ja = blockIdx.y * a.rows_per_seg;
    const int jb = min(g.BH, ja + a.rows_per_seg);
    if (ja >= jb) return;
    const int ximg = x0 + tid;
    const bool store_ok = ximg < g.W;
    const int nchunks = (g.D + 32 * NW - 1) / (32 * NW);
    const int last_pr = jb + 2 * HALF;  // padded rows [ja, last_pr) feed this CTA

    for (int chunk = 0; chunk < nchunks; chunk++) {
        const int wg0 = chunk * NW;                      // first 32-shift word of this chunk
        const int kbase = (PADL + x0) / 32 - 1 + wg0;    // global word index of Rs[.][0]
        uint32_t valid[NW];
#pragma unroll
        for (int w = 0; w < NW; w++) {
            int lanes = g.D - 32 * (wg0 + w);
            valid[w] = lanes >= 32 ? 0xFFFFFFFFu : (lanes <= 0 ? 0u : ((1u << lanes) - 1u));
        }
        uint32_t V[NW][PV];
#pragma unroll
        for (int w = 0; w < NW; w++)
#pragma unroll
            for (int p = 0; p < PV; p++) V[w][p] = 0;

        int slot0 = 0;  // ring slot of padded row p0
        for (int p0 = ja; p0 < last_pr; p0 += RB) {
            const int nrows = min(RB, last_pr - p0);

            // ---- stage: packed R words and the A/B splat table of rows p0 .. p0+nrows-1 ----
            for (int i = tid; i < nrows * RW; i += TW) {
                int r = i / RW, k = i - r * RW;
                int slot = slot0 + r;
                slot = slot >= NR ? slot - NR : slot;
                Rs[slot * RW + k] = __ldg(a.h.RB + (size_t)(p0 + r) * g.WPR + kbase + k);
            }
            for (int i = tid; i < nrows * ABN; i += TW) {
                int r = i / ABN, t = i - r * ABN;
                int slot = slot0 + r;
                slot = slot >= NR ? slot - NR : slot;
                int c = PADL + x0 - HALF + t;
                size_t o = (size_t)(p0 + r) * g.WPR + (c >> 5);
                uint32_t la = __ldg(a.h.LA + o), lb = __ldg(a.h.LB + o);
                uint2 e;
                e.x = 0u - ((la >> (c & 31)) & 1u);
                e.y = 0u - ((lb >> (c & 31)) & 1u);
                ABt[slot * ABROW + t] = e;
            }
            __syncthreads();

            // ---- pass A: one warp of walkers, lane -> (segment, row, word) ----
            if (tid < 32) {
                const int s = tid / (NW * RB), wl = tid - s * (NW * RB);
                const int r = wl / NW, w = wl - r * NW;
                if (r < nrows) {
                    int slot = slot0 + r;
                    slot = slot >= NR ? slot - NR : slot;
                    const uint32_t *rw = Rs + slot * RW + 1 + w + s * (SEG / 32);
                    const uint2 *ab = ABt + slot * ABROW + s * SEG;
                    uint4 *hq = Hq + (slot * NW + w) * HROW + s * SEG;
                    uint32_t *h5 = H5 + (slot * NW + w) * HROW + s * SEG;
                    uint32_t P[5] = {0, 0, 0, 0, 0};
                    uint32_t m[SEG + 2 * HALF];
#pragma unroll
                    for (int t = 0; t < SEG + 2 * HALF; t++) {
                        const int d = t - HALF;            // bit offset of pixel u inside this walker's words
                        const int wi = d >> 5, sh = d & 31;
                        uint32_t rwin = __funnelshift_r(rw[wi], rw[wi + 1], sh);
                        uint2 e = ab[t];
                        m[t] = (rwin & e.x) | (~rwin & e.y);
                        const uint32_t in = m[t], out = t >= N ? m[t - N] : 0u;
                        // up/down counter: +1 where in & ~out, -1 where out & ~in
                        uint32_t c = (in ^ out) & (P[0] ^ out);
                        P[0] ^= in ^ out;
#pragma unroll
                        for (int k = 1; k < KH; k++) {
                            uint32_t cn = c & (P[k] ^ out);
                            P[k] ^= c;
                            c = cn;
                        }
                        if (t >= 2 * HALF) {
                            hq[t - 2 * HALF] = make_uint4(P[0], P[1], P[2], P[3]);
                            if (KH > 4) h5[t - 2 * HALF] = P[4];
                        }
                    }
                }
            }
            __syncthreads();

            // ---- pass B: vertical running sum + winner-take-all, one pixel column per thread ----
            for (int r = 0; r < nrows; r++) {
                const int pr = p0 + r;
                int slot_n = slot0 + r;
                slot_n = slot_n >= NR ? slot_n - NR : slot_n;
#pragma unroll
                for (int w = 0; w < NW; w++) {
                    uint4 q = Hq[(slot_n * NW + w) * HROW + tid];
                    uint32_t h[5] = {q.x, q.y, q.z, q.w, 0};
                    if (KH > 4) h[4] = H5[(slot_n * NW + w) * HROW + tid];
                    planes_add<PV, KH>(V[w], h);
                }
                const int j = pr - 2 * HALF;  // output row whose window is now complete
                if (j < ja) continue;
                if (j > ja) {
                    int slot_o = slot_n - N;  // padded row j-1 left the window
                    slot_o = slot_o < 0 ? slot_o + NR : slot_o;
#pragma unroll
                    for (int w = 0; w < NW; w++) {
                        uint4 q = Hq[(slot_o * NW + w) * HROW + tid];
                        uint32_t h[5] = {q.x, q.y, q.z, q.w, 0};
                        if (KH > 4) h[4] = H5[(slot_o * NW + w) * HROW + tid];
                        planes_sub<PV, KH>(V[w], h);
                    }
                }
                // centre match word of this pixel (padded row j + HALF)
                int slot_c = slot_n - HALF;
                slot_c = slot_c < 0 ? slot_c + NR : slot_c;
                const uint2 e = ABt[slot_c * ABROW + HALF + tid];
                uint32_t M[NW], cand[NW];
#pragma unroll
                for (int w = 0; w < NW; w++) {
                    const uint32_t *rw = Rs + slot_c * RW + 1 + w + (tid >> 5);
                    uint32_t rwin = __funnelshift_r(rw[0], rw[1], tid & 31);
                    M[w] = (rwin & e.x) | (~rwin & e.y);
                    cand[w] = valid[w];
                }
                int best = 0;
#pragma unroll
                for (int p = PV - 1; p >= 0; p--) {
                    uint32_t t[NW], any = 0;
#pragma unroll
                    for (int w = 0; w < NW; w++) {
                        t[w] = cand[w] & V[w][p] & M[w];
                        any |= t[w];
                    }
                    if (any) {
#pragma unroll
                        for (int w = 0; w < NW; w++) cand[w] = t[w];
                        best |= 1 << p;
                    }
                }
                int idx = 0;
#pragma unroll
                for (int w = 0; w < NW; w++)
                    if (cand[w]) idx = 32 * w + 31 - __clz(cand[w]);  // later words overwrite: highest lane
                int web = 32 * wg0 + idx + 1;
                if (store_ok) {
                    size_t o = (size_t)(a.h.row0 + j) * g.W + ximg;
                    if (chunk == 0 || best >= a.h.best[o]) {
                        a.h.best[o] = best;
                        a.h.web[o] = web;
                    }
                }
            }
            __syncthreads();
            slot0 += nrows;
            slot0 = slot0 >= NR ? slot0 - NR : slot0;
        }
    }
}

template <int HALF, int NW>
int launch_one(const HotArgs &h, int num_sms, cudaStream_t s)
{
    using C = BS<HALF, NW>;
    auto kern = k_bitslice<HALF, NW>;
    static int occ_of_device[64] = {0};  // per instantiation and per device
    int dev = 0;
    SM_CUDA(cudaGetDevice(&dev));
    dev &= 63;
    if (occ_of_device[dev] == 0) {
        SM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        int occ = 0;
        SM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::TW, C::SMEM));
        occ_of_device[dev] = occ > 0 ? occ : 1;
    }
    const int blocks_per_sm = occ_of_device[dev];
    BitsliceArgs a;
    a.h = h;
    const int strips = (h.g.W + C::TW - 1) / C::TW;
    // aim at one full wave of resident CTAs; keep runs long enough to amortise the 2*half warm-up rows
    int slots = num_sms * blocks_per_sm;
    int segs = slots / strips;
    if (segs < 1) segs = 1;
    int min_rows = 4 * C::N > 32 ? 4 * C::N : 32;
    int max_segs = (h.g.BH + min_rows - 1) / min_rows;
    if (segs > max_segs) segs = max_segs;
    a.rows_per_seg = (h.g.BH + segs - 1) / segs;
    segs = (h.g.BH + a.rows_per_seg - 1) / a.rows_per_seg;
    dim3 grid(strips, segs);
    kern<<<grid, C::TW, C::SMEM, s>>>(a);
    SM_CUDA(cudaGetLastError());
    return 1;
}

template <int NW>
int dispatch_half(int half, const HotArgs &h, int num_sms, cudaStream_t s)
{
    switch (half) {
    case 0: return launch_one<0, NW>(h, num_sms, s);
    case 1: return launch_one<1, NW>(h, num_sms, s);
    case 2: return launch_one<2, NW>(h, num_sms, s);
    case 3: return launch_one<3, NW>(h, num_sms, s);
    case 4: return launch_one<4, NW>(h, num_sms, s);
    case 5: return launch_one<5, NW>(h, num_sms, s);
    case 6: return launch_one<6, NW>(h, num_sms, s);
    case 7: return launch_one<7, NW>(h, num_sms, s);
    case 8: return launch_one<8, NW>(h, num_sms, s);
    case 9: return launch_one<9, NW>(h, num_sms, s);
    case 10: return launch_one<10, NW>(h, num_sms, s);
    default: set_error("bit-sliced kernel: window half %d not instantiated", half); return SM_ERR_ARG;
    }
}

}  // namespace

// square_width up to 21 (the reference default, stereo.c:8); wider windows take the direct kernel.
bool bitslice_supports(int half, int D) { return half >= 0 && half <= 10 && D >= 1 && D <= 512; }

int launch_bitslice(const HotArgs &h, int num_sms, cudaStream_t s)
{
    if (h.g.D <= 32) return dispatch_half<1>(h.g.half, h, num_sms, s);
    return dispatch_half<2>(h.g.half, h, num_sms, s);
}

}  // namespace smb
