// stereo_b200.cu -- the C ABI of include/stereo_b200.h: context, transfers, orchestration.
//
// Stands in for the host side of the reference's CUDA programs: main()'s
// MAKE_GPU_COPY uploads (stereo.cu:402-403), algorithm()'s allocations, launch
// sequence and frees (stereo.cu:296-347), write_gpu_image's D2H (image.cu:15-23) and
// the min/max helpers (util.cu:34-42).  No kernel lives here (k_*.cu).
#include <stdarg.h>
#include <stdlib.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "sm_common.cuh"

namespace smb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

}  // namespace smb

using namespace smb;

struct sm_ctx {
    int device = 0;
    int W = 0, FH = 0;         // frame geometry
    int row0 = 0, row1 = 0;    // output rows owned by this context
    int D = 0, sw = 0, half = 0, variant = 0;
    int kernel = SM_KERNEL_AUTO;
    int num_sms = 148;
    PackedGeom g{};

    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    // sm_profile_*: one event triple per hot-path call
    cudaEvent_t *prof_ev = nullptr;
    int prof_cap = 0, prof_n = 0;
    int last_launches = 0;
    int tuned_segs = 0;  // SM_OPT_ROW_RUNS: row runs per strip of a one-pair launch (0: the kernel's cost model)
    int occ = 0;         // resident warps per SM of the bit-sliced kernel of this geometry (prepare_bitslice)

    // frame-sized device arrays (a band context touches only the rows it needs)
    uint8_t *img_u8[2] = {nullptr, nullptr};
    double *img_f64[2] = {nullptr, nullptr};
    int img_kind = 0;  // 0 none, 1 u8, 2 f64
    uint8_t *edges[2] = {nullptr, nullptr};
    bool have_edges = false;
    bool planes_valid = false;  // LA/LB/RB already hold the packed planes of edges[] (sm_edges wrote them itself)
    uint32_t *edge_lut = nullptr;  // detector decisions for all (L, R) sum pairs at lut_threshold
    double lut_threshold = -1.0;
    bool edges_fp64_only = false;  // SM_OPT_EDGES_FP64: always take the FP64 kernel (cross-check)
    int pipe_group_opt = 0;        // SM_OPT_PIPE_GROUP: pairs per sm_run_batch stage (0: by frame size)
    int32_t *best = nullptr, *web = nullptr;
    bool have_web = false;
    int32_t *web2 = nullptr, *tmp = nullptr;  // step 3 ping-pong
    const int32_t *web_filled = nullptr;      // result of fill_web_holes (may alias web)
    bool have_web2 = false;
    bool web_may_have_holes = false;          // only a web from sm_set_web can contain zeros
    uint8_t *out = nullptr;
    bool have_out = false;
    int32_t *minmax = nullptr;       // two (min, max) slots, used alternately (k_step3.cu)
    int32_t *minmax_host = nullptr;  // pinned landing place of a slot
    int minmax_cur = 0;
    // band-sized packed planes
    uint32_t *LA = nullptr, *LB = nullptr, *RB = nullptr;
    // sm_match_wta_dev_batch: a second stream packs pair k+1.. while the main kernel of
    // pair k runs; NPB packed-plane sets rotate between the two
    static constexpr int NPB = 3;
    cudaStream_t pack_stream = nullptr, main2_stream = nullptr;  // main kernels alternate stream / main2_stream
    cudaEvent_t ev_join = nullptr;
    int batch_group = 1;      // pairs per launch in sm_match_wta_dev_batch (1 for the literal kernel)
    int batch_group_cap = 0;  // ... of the bit-sliced kernel; the plane sets are sized for it
    bool batch_ready = false;
    uint32_t *pLA[NPB] = {nullptr, nullptr, nullptr}, *pLB[NPB] = {nullptr, nullptr, nullptr},
             *pRB[NPB] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_packed[NPB] = {nullptr, nullptr, nullptr}, ev_used[NPB] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr;
    // scratch for on-demand debug planes / u8 web
    uint8_t *scratch_u8 = nullptr;
    int32_t *scratch_i32 = nullptr;

    // sm_run_batch: a three-stage pipeline (H2D | edges + hot path | D2H) over groups of pairs;
    // NPIPE group buffer sets rotate between the stages
    static constexpr int NPIPE = 3;
    struct Pipe {
        int group = 0;  // pairs per stage
        bool ready = false;
        cudaStream_t up = nullptr, down = nullptr;
        uint8_t *img[NPIPE] = {nullptr, nullptr, nullptr};    // [2*group][npix]: first images, then second images
        uint8_t *edg[NPIPE] = {nullptr, nullptr, nullptr};    // same layout
        int32_t *web[NPIPE] = {nullptr, nullptr, nullptr};    // [group][npix]
        int32_t *best[NPIPE] = {nullptr, nullptr, nullptr};   // [group][npix]
        uint8_t *web8[NPIPE] = {nullptr, nullptr, nullptr};   // [group][npix], only for the u8 result
        cudaEvent_t ev_up[NPIPE] = {nullptr, nullptr, nullptr}, ev_comp[NPIPE] = {nullptr, nullptr, nullptr},
                    ev_down[NPIPE] = {nullptr, nullptr, nullptr};
    } pipe;

    size_t npix() const { return (size_t)W * FH; }
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

#define SM_ENTER(ctx)                                                   \
    if (!(ctx)) {                                                       \
        set_error("%s: NULL context", __func__);                        \
        return SM_ERR_ARG;                                              \
    }                                                                   \
    DeviceGuard guard__((ctx)->device);                                 \
    if (!guard__.ok) {                                                  \
        set_error("%s: cudaSetDevice(%d) failed", __func__, (ctx)->device); \
        return SM_ERR_CUDA;                                             \
    }

template <typename T>
int dev_alloc(T **p, size_t count)
{
    if (*p) return SM_OK;
    SM_CUDA(cudaMalloc((void **)p, count * sizeof(T)));
    return SM_OK;
}

// Frame rows [lo, hi) as up to a few contiguous [start, start+count) runs inside the
// frame: wrapped mod FH for WRAP (util.h:42-47), clipped for GHOST.
struct RowRuns {
    int n = 0;
    int start[4], count[4];
};

RowRuns row_runs(int lo, int hi, int FH, bool wrap)
{
    RowRuns r;
    if (hi - lo >= FH) {
        r.n = 1;
        r.start[0] = 0;
        r.count[0] = FH;
        return r;
    }
    if (!wrap) {
        lo = lo < 0 ? 0 : lo;
        hi = hi > FH ? FH : hi;
        if (hi > lo) {
            r.n = 1;
            r.start[0] = lo;
            r.count[0] = hi - lo;
        }
        return r;
    }
    int a = ((lo % FH) + FH) % FH, len = hi - lo;
    while (len > 0 && r.n < 4) {
        int c = len < FH - a ? len : FH - a;
        r.start[r.n] = a;
        r.count[r.n] = c;
        r.n++;
        len -= c;
        a = 0;
    }
    return r;
}

template <typename T>
int copy_rows_h2d(sm_ctx *c, T *dst, const T *src, int lo, int hi)
{
    RowRuns rr = row_runs(lo, hi, c->FH, c->variant == SM_WRAP);
    for (int k = 0; k < rr.n; k++) {
        size_t off = (size_t)rr.start[k] * c->W;
        SM_CUDA(cudaMemcpyAsync(dst + off, src + off, (size_t)rr.count[k] * c->W * sizeof(T),
                                cudaMemcpyHostToDevice, c->stream));
    }
    return SM_OK;
}

template <typename T>
int copy_band_d2h(sm_ctx *c, T *host, const T *dev)
{
    size_t off = (size_t)c->row0 * c->W;
    SM_CUDA(cudaMemcpyAsync(host + off, dev + off, (size_t)(c->row1 - c->row0) * c->W * sizeof(T),
                            cudaMemcpyDeviceToHost, c->stream));
    return SM_OK;
}

HotArgs hot_args(sm_ctx *c, int32_t *best, int32_t *web)
{
    HotArgs a;
    a.g = c->g;
    a.LA = c->LA;
    a.LB = c->LB;
    a.RB = c->RB;
    a.best = best;
    a.web = web;
    a.row0 = c->row0;
    a.force_segs = c->tuned_segs;
    a.blocks_per_sm = c->occ;
    return a;
}

int launch_main(sm_ctx *c, const HotArgs &a, cudaStream_t st)
{
    int k = c->kernel;
    if (k == SM_KERNEL_AUTO) k = bitslice_supports(c->half, c->D) ? SM_KERNEL_BITSLICE : SM_KERNEL_DIRECT;
    if (k == SM_KERNEL_BITSLICE) {
        if (!bitslice_supports(c->half, c->D)) {
            set_error("bit-sliced kernel does not cover square_width %d / num_shifts %d", c->sw, c->D);
            return SM_ERR_ARG;
        }
        return launch_bitslice(a, c->num_sms, st);
    }
    return launch_direct(a, st);
}

int run_hot(sm_ctx *c, const uint8_t *e1, const uint8_t *e2, int32_t *best, int32_t *web)
{
    int launches = 0, rc;
    cudaEvent_t *pe = (c->prof_ev && c->prof_n < c->prof_cap) ? c->prof_ev + 3 * c->prof_n : nullptr;
    SM_CUDA(cudaEventRecord(c->ev0, c->stream));
    if (pe) SM_CUDA(cudaEventRecord(pe[0], c->stream));
    // sm_edges writes the packed planes itself (k_edges_planes): no pack launch then
    const bool packed = c->planes_valid && e1 == c->edges[0] && e2 == c->edges[1];
    if (!packed) {
        rc = launch_pack(e1, e2, c->FH, c->row0, c->variant, c->g, c->LA, c->LB, c->RB, c->stream);
        if (rc < 0) return rc;
        launches += rc;
    }
    if (pe) SM_CUDA(cudaEventRecord(pe[1], c->stream));
    HotArgs a = hot_args(c, best, web);
#ifdef SMB_DEV
    static const bool no_pdl = getenv("SMB_NO_PDL") && atoi(getenv("SMB_NO_PDL"));  // experiment hook, development build only
#else
    constexpr bool no_pdl = false;
#endif
    // the producer of the planes (pack or edge kernel) is the launch just before on this stream unless the
    // per-kernel profile put an event in between: the main kernel may start as its programmatic dependent.
    // (It waits for every earlier grid of the stream to complete before it reads anything, whatever that grid is.)
    a.after_pack = pe == nullptr && !no_pdl;
    rc = launch_main(c, a, c->stream);
    if (rc < 0) return rc;
    launches += rc;
    SM_CUDA(cudaEventRecord(c->ev1, c->stream));
    if (pe) {
        SM_CUDA(cudaEventRecord(pe[2], c->stream));
        c->prof_n++;
    }
    c->timed = true;
    c->last_launches = launches;
    return SM_OK;
}

void profile_free(sm_ctx *c)
{
    for (int k = 0; k < 3 * c->prof_cap; k++)
        if (c->prof_ev[k]) cudaEventDestroy(c->prof_ev[k]);
    free(c->prof_ev);
    c->prof_ev = nullptr;
    c->prof_cap = c->prof_n = 0;
}

}  // namespace

// ---- library-level ---------------------------------------------------------------

extern "C" const char *sm_last_error(void) { return g_err; }

// 1.1: sm_multi_*, sm_bands_*; 1.2: sm_set_option / sm_get_info, square_width 0, the microbenchmarks moved out
extern "C" int sm_version(void) { return (1 << 16) | 2; }

extern "C" int sm_device_count(void)
{
    int n = 0;
    SM_CUDA(cudaGetDeviceCount(&n));
    return n;
}

extern "C" int sm_host_alloc(void **ptr, size_t bytes)
{
    SM_REQUIRE(ptr, "sm_host_alloc: NULL ptr");
    SM_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return SM_OK;
}

extern "C" int sm_host_free(void *ptr)
{
    if (ptr) SM_CUDA(cudaFreeHost(ptr));
    return SM_OK;
}

extern "C" int sm_band_rows(int height, int n_bands, int band, int *row0, int *row1)
{
    SM_REQUIRE(height > 0 && n_bands > 0 && band >= 0 && band < n_bands && row0 && row1,
               "sm_band_rows: bad arguments");
    // as even as possible: the first (height % n_bands) bands get one extra row
    int q = height / n_bands, r = height % n_bands;
    *row0 = band * q + (band < r ? band : r);
    *row1 = *row0 + q + (band < r ? 1 : 0);
    return SM_OK;
}

// ---- context -----------------------------------------------------------------------

extern "C" int sm_create_band(sm_ctx **out, int device, int width, int frame_height, int row0,
                              int row1, int num_shifts, int square_width, int variant)
{
    SM_REQUIRE(out, "sm_create: NULL ctx pointer");
    *out = nullptr;
    SM_REQUIRE(width > 0 && frame_height > 0, "sm_create: width/height must be positive");
    SM_REQUIRE((size_t)width * frame_height < ((size_t)1 << 31), "sm_create: frame too large");
    SM_REQUIRE(num_shifts >= 1 && num_shifts <= 512, "sm_create: num_shifts must be in [1, 512]");
    // the reference takes any square_width up to the frame size (stereo.cu:395-398); its window is
    // half = square_width / 2 either side, so 0 (and -1) mean the 1x1 window.  Wider than 63 is not built here.
    SM_REQUIRE(square_width >= -1 && square_width <= 63, "sm_create: square_width must be in [0, 63]");
    // same check as the reference driver (stereo.cu:395-398)
    SM_REQUIRE(square_width <= width && square_width <= frame_height,
               "sm_create: square width must not be higher than image width/height");
    SM_REQUIRE(variant == SM_WRAP || variant == SM_GHOST, "sm_create: unknown variant");
    SM_REQUIRE(row0 >= 0 && row1 > row0 && row1 <= frame_height, "sm_create: bad row band");
    int ndev = 0;
    SM_CUDA(cudaGetDeviceCount(&ndev));
    SM_REQUIRE(device >= 0 && device < ndev, "sm_create: device %d of %d", device, ndev);

    sm_ctx *c = new (std::nothrow) sm_ctx;
    if (!c) {
        set_error("sm_create: out of memory");
        return SM_ERR_NOMEM;
    }
    c->device = device;
    DeviceGuard guard(device);
    c->W = width;
    c->FH = frame_height;
    c->row0 = row0;
    c->row1 = row1;
    c->D = num_shifts;
    c->sw = square_width;
    c->half = square_width / 2;
    c->variant = variant;
    c->g.W = width;
    c->g.BH = row1 - row0;
    c->g.half = c->half;
    c->g.ER = c->g.BH + 2 * c->half;
    c->g.D = num_shifts;
    c->g.WPR = packed_words_per_row(width, c->half, num_shifts);

    int rc = SM_OK;
    auto fail = [&](int code) {
        sm_destroy(c);
        return code;
    };
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) {
        set_error("sm_create: stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(SM_ERR_CUDA);
    }
    c->own_stream = true;
    size_t n = c->npix(), pw = (size_t)c->g.ER * c->g.WPR;
    // both edge maps (and, in sm_upload_u8, both images) live in ONE allocation each, second right after first,
    // so that sm_edges detects the pair in a single launch (grid z = image)
    if ((rc = dev_alloc(&c->edges[0], 2 * n)) ||
        (rc = dev_alloc(&c->best, n)) || (rc = dev_alloc(&c->web, n)) ||
        (rc = dev_alloc(&c->LA, pw)) || (rc = dev_alloc(&c->LB, pw)) || (rc = dev_alloc(&c->RB, pw)) ||
        (rc = dev_alloc(&c->minmax, 4)))
        return fail(rc);
    if (cudaHostAlloc((void **)&c->minmax_host, 2 * sizeof(int32_t), cudaHostAllocDefault) != cudaSuccess) {
        set_error("sm_create: pinned allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(SM_ERR_NOMEM);
    }
    if ((rc = launch_minmax_arm(c->minmax, c->stream)) < 0) return fail(rc);
    c->edges[1] = c->edges[0] + n;
    // whole-frame contexts run step 3 as well: its buffers belong to the untimed set-up, like
    // the allocations at the top of the reference's algorithm() (stereo.cu:299-306)
    if (row0 == 0 && row1 == frame_height && ((rc = dev_alloc(&c->tmp, n)) || (rc = dev_alloc(&c->out, n))))
        return fail(rc);
    // load the kernels this geometry uses now, not inside the first timed call
    warm_edges(variant);
    warm_pack(variant);
    warm_direct();
    warm_step3();
    if (bitslice_supports(c->half, c->D)) {
        HotArgs a = hot_args(c, c->best, c->web);
        if ((rc = prepare_bitslice(a, c->num_sms)) < 0) return fail(rc);
        c->occ = rc;
    }
    // the detector's decision table depends on the threshold alone: build it for the reference's default
    // (DEFAULT_THRESHOLD 0.15, stereo.cu:7) here, so that sm_edges at that threshold is one launch; another threshold
    // rebuilds the table on first use
    if ((rc = dev_alloc(&c->edge_lut, edge_lut_words()))) return fail(rc);
    if ((rc = launch_edge_lut(0.15, c->edge_lut, c->stream)) < 0) return fail(rc);
    c->lut_threshold = 0.15;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) {
        set_error("sm_create: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(SM_ERR_CUDA);
    }
    *out = c;
    return SM_OK;
}

extern "C" int sm_create(sm_ctx **ctx, int device, int width, int height, int num_shifts,
                         int square_width, int variant)
{
    return sm_create_band(ctx, device, width, height, 0, height, num_shifts, square_width, variant);
}

extern "C" int sm_destroy(sm_ctx *c)
{
    if (!c) return SM_OK;
    DeviceGuard guard(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (cudaStream_t st : {c->pipe.up, c->pipe.down})
        if (st) {
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
    for (int k = 0; k < sm_ctx::NPIPE; k++) {
        void *pp[] = {c->pipe.img[k], c->pipe.edg[k], c->pipe.web[k], c->pipe.best[k], c->pipe.web8[k]};
        for (void *p : pp)
            if (p) cudaFree(p);
        for (cudaEvent_t e : {c->pipe.ev_up[k], c->pipe.ev_comp[k], c->pipe.ev_down[k]})
            if (e) cudaEventDestroy(e);
    }
    if (c->edge_lut) cudaFree(c->edge_lut);
    if (c->minmax_host) cudaFreeHost(c->minmax_host);
    void *ptrs[] = {c->img_u8[0], nullptr,      c->img_f64[0], c->img_f64[1], c->edges[0], nullptr,  // [1] = [0] + npix
                    c->best,      c->web,       c->web2,       c->tmp,        c->out,      c->minmax,
                    c->LA,        c->LB,        c->RB,         c->scratch_u8, c->scratch_i32};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (c->prof_ev) profile_free(c);
    if (c->pack_stream) {
        cudaStreamSynchronize(c->pack_stream);
        cudaStreamDestroy(c->pack_stream);
    }
    if (c->main2_stream) {
        cudaStreamSynchronize(c->main2_stream);
        cudaStreamDestroy(c->main2_stream);
    }
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    for (int k = 0; k < sm_ctx::NPB; k++) {
        if (c->pLA[k]) cudaFree(c->pLA[k]);
        if (c->pLB[k]) cudaFree(c->pLB[k]);
        if (c->pRB[k]) cudaFree(c->pRB[k]);
        if (c->ev_packed[k]) cudaEventDestroy(c->ev_packed[k]);
        if (c->ev_used[k]) cudaEventDestroy(c->ev_used[k]);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return SM_OK;
}

extern "C" int sm_set_stream(sm_ctx *c, void *cuda_stream)
{
    SM_ENTER(c);
    if (c->own_stream && c->stream) {
        SM_CUDA(cudaStreamSynchronize(c->stream));
        SM_CUDA(cudaStreamDestroy(c->stream));
    }
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
    c->timed = false;
    return SM_OK;
}

extern "C" int sm_set_kernel(sm_ctx *c, int kernel)
{
    SM_ENTER(c);
    SM_REQUIRE(kernel >= SM_KERNEL_AUTO && kernel <= SM_KERNEL_BITSLICE, "sm_set_kernel: unknown kernel");
    c->kernel = kernel;
    return SM_OK;
}

extern "C" int sm_set_option(sm_ctx *c, int option, int value)
{
    SM_ENTER(c);
    switch (option) {
    case SM_OPT_EDGES_FP64: c->edges_fp64_only = value != 0; return SM_OK;
    case SM_OPT_PIPE_GROUP:
        SM_REQUIRE(value >= 0 && value <= 1024, "sm_set_option: pairs per stage must be in [0, 1024]");
        SM_REQUIRE(!c->pipe.ready, "sm_set_option: the batch pipeline of this context is already set up");
        c->pipe_group_opt = value;
        return SM_OK;
    case SM_OPT_ROW_RUNS:
        SM_REQUIRE(value >= 0, "sm_set_option: row runs must be >= 0");
        c->tuned_segs = value;
        return SM_OK;
    default: set_error("sm_set_option: unknown option %d", option); return SM_ERR_ARG;
    }
}

extern "C" int sm_get_info(sm_ctx *c, int what)
{
    SM_ENTER(c);
    switch (what) {
    case SM_INFO_WARPS_PER_SM: return c->occ;
    case SM_INFO_PAIRS_PER_LAUNCH: return c->batch_group_cap;
    case SM_INFO_TMEM_COLUMNS: return bitslice_supports(c->half, c->D) ? bitslice_tmem_columns(hot_args(c, c->best, c->web)) : 0;
    case SM_INFO_EDGE_THRESHOLDS: {
        if (!c->edge_lut || c->lut_threshold < 0.0) return -1;
        uint32_t flag = 0;
        SM_CUDA(cudaMemcpyAsync(&flag, c->edge_lut + edge_lut_words() - 1, sizeof flag, cudaMemcpyDeviceToHost, c->stream));
        SM_CUDA(cudaStreamSynchronize(c->stream));
        return flag != 0u;
    }
    default: set_error("sm_get_info: unknown item %d", what); return SM_ERR_ARG;
    }
}

extern "C" int sm_synchronize(sm_ctx *c)
{
    SM_ENTER(c);
    SM_CUDA(cudaStreamSynchronize(c->stream));
    return SM_OK;
}

// ---- step 0: upload ------------------------------------------------------------------

// rows of the brightness images this context reads: its output rows, the window halo,
// and one more row for the 3x3 edge stencil
static void image_rows(const sm_ctx *c, int *lo, int *hi)
{
    *lo = c->row0 - c->half - 1;
    *hi = c->row1 + c->half + 1;
}

extern "C" int sm_upload_u8(sm_ctx *c, const uint8_t *first, const uint8_t *second)
{
    SM_ENTER(c);
    SM_REQUIRE(first && second, "sm_upload_u8: NULL image");
    int rc, lo, hi;
    if ((rc = dev_alloc(&c->img_u8[0], 2 * c->npix()))) return rc;
    c->img_u8[1] = c->img_u8[0] + c->npix();
    image_rows(c, &lo, &hi);
    if ((rc = copy_rows_h2d(c, c->img_u8[0], first, lo, hi)) ||
        (rc = copy_rows_h2d(c, c->img_u8[1], second, lo, hi)))
        return rc;
    c->img_kind = 1;
    return SM_OK;
}

extern "C" int sm_upload_f64(sm_ctx *c, const double *first, const double *second)
{
    SM_ENTER(c);
    SM_REQUIRE(first && second, "sm_upload_f64: NULL image");
    int rc, lo, hi;
    if ((rc = dev_alloc(&c->img_f64[0], c->npix())) || (rc = dev_alloc(&c->img_f64[1], c->npix()))) return rc;
    image_rows(c, &lo, &hi);
    if ((rc = copy_rows_h2d(c, c->img_f64[0], first, lo, hi)) ||
        (rc = copy_rows_h2d(c, c->img_f64[1], second, lo, hi)))
        return rc;
    c->img_kind = 2;
    return SM_OK;
}

// ---- step 1: edges -------------------------------------------------------------------

extern "C" int sm_edges(sm_ctx *c, double threshold)
{
    SM_ENTER(c);
    if (c->img_kind == 0) {
        set_error("sm_edges: no image uploaded");
        return SM_ERR_STATE;
    }
    // same range check as the reference driver (stereo.cu:391-394)
    SM_REQUIRE(threshold >= 0.0 && threshold <= 1.0, "sm_edges: threshold must be between 0 and 1");
    int ystart = c->row0 - c->half, nrows = c->g.ER;
    if (nrows >= c->FH) {
        ystart = 0;
        nrows = c->FH;
    }
    const bool use_lut = c->img_kind == 1 && !c->edges_fp64_only;
    if (use_lut && c->lut_threshold != threshold) {
        int rc;
        if ((rc = dev_alloc(&c->edge_lut, edge_lut_words()))) return rc;
        if ((rc = launch_edge_lut(threshold, c->edge_lut, c->stream)) < 0) return rc;
        c->lut_threshold = threshold;
    }
    c->planes_valid = false;
    if (use_lut) {
        // both images in one launch, straight into the packed planes of the hot path; the byte maps are written
        // as well (sm_download(SM_EDGES*), debug planes)
        int rc = launch_edges_planes(c->img_u8[0], c->img_u8[1], c->FH, c->row0, c->variant, c->g, threshold, c->edge_lut,
                                     c->LA, c->LB, c->RB, c->edges[0], c->edges[1], c->stream);
        if (rc < 0) return rc;
        c->planes_valid = true;
    } else {
        for (int k = 0; k < 2; k++) {
            int rc;
            if (c->img_kind == 1)
                rc = launch_edges<uint8_t>(c->img_u8[k], c->W, c->FH, ystart, nrows, c->variant, threshold,
                                           c->edges[k], c->stream);
            else
                rc = launch_edges<double>(c->img_f64[k], c->W, c->FH, ystart, nrows, c->variant, threshold,
                                          c->edges[k], c->stream);
            if (rc < 0) return rc;
        }
    }
    c->have_edges = true;
    return SM_OK;
}

extern "C" int sm_set_edges(sm_ctx *c, const uint8_t *first_edges, const uint8_t *second_edges)
{
    SM_ENTER(c);
    SM_REQUIRE(first_edges && second_edges, "sm_set_edges: NULL edge map");
    int rc;
    if ((rc = copy_rows_h2d(c, c->edges[0], first_edges, c->row0 - c->half, c->row1 + c->half)) ||
        (rc = copy_rows_h2d(c, c->edges[1], second_edges, c->row0 - c->half, c->row1 + c->half)))
        return rc;
    c->planes_valid = false;
    c->have_edges = true;
    return SM_OK;
}

// ---- step 2: the hot path --------------------------------------------------------------

extern "C" int sm_match_wta(sm_ctx *c)
{
    SM_ENTER(c);
    if (!c->have_edges) {
        set_error("sm_match_wta: no edges (call sm_edges or sm_set_edges first)");
        return SM_ERR_STATE;
    }
    int rc = run_hot(c, c->edges[0], c->edges[1], c->best, c->web);
    if (rc) return rc;
    c->have_web = true;
    c->web_may_have_holes = false;
    c->have_web2 = c->have_out = false;
    return SM_OK;
}

extern "C" int sm_match_wta_dev(sm_ctx *c, const uint8_t *d_first_edges, const uint8_t *d_second_edges,
                                int32_t *d_best, int32_t *d_web)
{
    SM_ENTER(c);
    SM_REQUIRE(d_first_edges && d_second_edges && d_best && d_web, "sm_match_wta_dev: NULL pointer");
    return run_hot(c, d_first_edges, d_second_edges, d_best, d_web);
}

// The batched hot path on device-resident inputs: either byte edge maps (packed by k_pack) or, with
// from_images, the 8-bit images themselves (edges detected straight into the packed planes by k_edges_planes;
// the byte maps never exist).  Producer launches run on a second stream one group ahead of the main kernels.
static int batch_core(sm_ctx *c, int n_pairs, bool from_images, double threshold, const uint8_t *d_first_edges,
                      const uint8_t *d_second_edges, size_t edge_stride, int32_t *d_best, int32_t *d_web,
                      size_t out_stride)
{
    constexpr int NPB = sm_ctx::NPB;
    const size_t pw = (size_t)c->g.ER * c->g.WPR;
    int kk = c->kernel;
    if (kk == SM_KERNEL_AUTO) kk = bitslice_supports(c->half, c->D) ? SM_KERNEL_BITSLICE : SM_KERNEL_DIRECT;
    if (!c->batch_ready) {
        // Set-up of the batch path.  Every object is created only if it does not exist yet and the ready flag is
        // set last, so that a call that failed half way (out of memory on the plane sets, say) is simply resumed
        // by the next one instead of launching with missing buffers.
        // Pairs per launch: enough that every warp's run of rows is long against its warm-up rows
        if (c->batch_group_cap == 0) {
            c->batch_group_cap = 1;
            if (bitslice_supports(c->half, c->D)) {
                HotArgs a = hot_args(c, d_best, d_web);
                int pmax = 16;
#ifdef SMB_DEV
                if (getenv("SMB_PMAX")) pmax = atoi(getenv("SMB_PMAX"));  // experiment hook, development build only
#endif
                c->batch_group_cap = bitslice_pairs_per_launch(a, c->num_sms, pmax);
            }
        }
        if (!c->pack_stream) SM_CUDA(cudaStreamCreateWithFlags(&c->pack_stream, cudaStreamNonBlocking));
        if (!c->main2_stream) SM_CUDA(cudaStreamCreateWithFlags(&c->main2_stream, cudaStreamNonBlocking));
        if (!c->ev_fork) SM_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        if (!c->ev_join) SM_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
        for (int k = 0; k < NPB; k++) {
            int rc;
            if ((rc = dev_alloc(&c->pLA[k], pw * c->batch_group_cap)) || (rc = dev_alloc(&c->pLB[k], pw * c->batch_group_cap)) ||
                (rc = dev_alloc(&c->pRB[k], pw * c->batch_group_cap)))
                return rc;
            if (!c->ev_packed[k]) SM_CUDA(cudaEventCreateWithFlags(&c->ev_packed[k], cudaEventDisableTiming));
            if (!c->ev_used[k]) SM_CUDA(cudaEventCreateWithFlags(&c->ev_used[k], cudaEventDisableTiming));
        }
        c->batch_ready = true;
    }
    // the kernel may have been switched since the last call (sm_set_kernel): the literal kernel takes one pair
    // per launch, the bit-sliced one a group
    c->batch_group = kk == SM_KERNEL_BITSLICE ? c->batch_group_cap : 1;
    const int G = c->batch_group;
    // everything queued on the context's stream so far (the producers of the edge maps) comes first
    SM_CUDA(cudaEventRecord(c->ev0, c->stream));
    SM_CUDA(cudaEventRecord(c->ev_fork, c->stream));
    SM_CUDA(cudaStreamWaitEvent(c->pack_stream, c->ev_fork, 0));
    SM_CUDA(cudaStreamWaitEvent(c->main2_stream, c->ev_fork, 0));
    int launches = 0, group = 0;
    for (int k = 0; k < n_pairs; k += G, group++) {
        const int np = n_pairs - k < G ? n_pairs - k : G;  // pairs in this launch
        const int b = group % NPB;
        int rc;
        if (group >= NPB) SM_CUDA(cudaStreamWaitEvent(c->pack_stream, c->ev_used[b], 0));  // set b is free again
        if (from_images)
            rc = launch_edges_planes(d_first_edges + (size_t)k * edge_stride, d_second_edges + (size_t)k * edge_stride,
                                     c->FH, c->row0, c->variant, c->g, threshold, c->edge_lut, c->pLA[b], c->pLB[b],
                                     c->pRB[b], nullptr, nullptr, c->pack_stream, np, edge_stride, pw);
        else
            rc = launch_pack(d_first_edges + (size_t)k * edge_stride, d_second_edges + (size_t)k * edge_stride, c->FH,
                             c->row0, c->variant, c->g, c->pLA[b], c->pLB[b], c->pRB[b], c->pack_stream, np,
                             edge_stride, pw);
        if (rc < 0) return rc;
        launches += rc;
        SM_CUDA(cudaEventRecord(c->ev_packed[b], c->pack_stream));
        // consecutive launches are independent: alternate two streams so that the next main
        // kernel's warps fill the SM slots the previous one frees (no tail / ramp between them)
        cudaStream_t ms = (group & 1) ? c->main2_stream : c->stream;
        SM_CUDA(cudaStreamWaitEvent(ms, c->ev_packed[b], 0));
        cudaEvent_t *pe = (c->prof_ev && c->prof_n < c->prof_cap) ? c->prof_ev + 3 * c->prof_n : nullptr;
        if (pe) {
            SM_CUDA(cudaEventRecord(pe[0], ms));  // pack runs on the other stream: not timed here
            SM_CUDA(cudaEventRecord(pe[1], ms));
        }
        HotArgs a = hot_args(c, d_best + (size_t)k * out_stride, d_web + (size_t)k * out_stride);
        a.LA = c->pLA[b];
        a.LB = c->pLB[b];
        a.RB = c->pRB[b];
        a.npairs = np;
        a.plane_stride = pw;
        a.out_stride = out_stride;
        rc = launch_main(c, a, ms);
        if (rc < 0) return rc;
        launches += rc;
        if (pe) {
            SM_CUDA(cudaEventRecord(pe[2], ms));
            c->prof_n++;
        }
        SM_CUDA(cudaEventRecord(c->ev_used[b], ms));
    }
    // join: the context's stream is complete only when both main streams are
    SM_CUDA(cudaEventRecord(c->ev_join, c->main2_stream));
    SM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    SM_CUDA(cudaEventRecord(c->ev1, c->stream));
    c->timed = true;
    c->last_launches = launches;
    return SM_OK;
}

extern "C" int sm_match_wta_dev_batch(sm_ctx *c, int n_pairs, const uint8_t *d_first_edges,
                                      const uint8_t *d_second_edges, size_t edge_stride, int32_t *d_best,
                                      int32_t *d_web, size_t out_stride)
{
    SM_ENTER(c);
    SM_REQUIRE(n_pairs >= 0 && d_first_edges && d_second_edges && d_best && d_web,
               "sm_match_wta_dev_batch: bad arguments");
    SM_REQUIRE(edge_stride >= c->npix() && out_stride >= c->npix(), "sm_match_wta_dev_batch: stride below frame size");
    return batch_core(c, n_pairs, false, 0.0, d_first_edges, d_second_edges, edge_stride, d_best, d_web, out_stride);
}

extern "C" int sm_elapsed_ms(sm_ctx *c, float *ms)
{
    SM_ENTER(c);
    SM_REQUIRE(ms, "sm_elapsed_ms: NULL");
    if (!c->timed) {
        set_error("sm_elapsed_ms: no hot-path call has been made on this context");
        return SM_ERR_STATE;
    }
    SM_CUDA(cudaEventSynchronize(c->ev1));
    SM_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return SM_OK;
}

extern "C" int sm_last_launches(sm_ctx *c) { return c ? c->last_launches : SM_ERR_ARG; }

extern "C" int sm_profile_begin(sm_ctx *c, int max_calls)
{
    SM_ENTER(c);
    SM_REQUIRE(max_calls >= 0 && max_calls <= (1 << 20), "sm_profile_begin: bad max_calls");
    SM_CUDA(cudaStreamSynchronize(c->stream));
    if (c->prof_ev) profile_free(c);
    if (max_calls == 0) return SM_OK;
    c->prof_ev = (cudaEvent_t *)calloc((size_t)3 * max_calls, sizeof(cudaEvent_t));
    if (!c->prof_ev) {
        set_error("sm_profile_begin: out of memory");
        return SM_ERR_NOMEM;
    }
    c->prof_cap = max_calls;
    for (int k = 0; k < 3 * max_calls; k++) SM_CUDA(cudaEventCreate(&c->prof_ev[k]));
    return SM_OK;
}

extern "C" int sm_profile_read(sm_ctx *c, int *n_calls, double *pack_ms_total, double *main_ms_total)
{
    SM_ENTER(c);
    SM_REQUIRE(n_calls && pack_ms_total && main_ms_total, "sm_profile_read: NULL");
    SM_CUDA(cudaStreamSynchronize(c->stream));
    double p = 0, m = 0;
    for (int k = 0; k < c->prof_n; k++) {
        float a = 0, b = 0;
        SM_CUDA(cudaEventElapsedTime(&a, c->prof_ev[3 * k], c->prof_ev[3 * k + 1]));
        SM_CUDA(cudaEventElapsedTime(&b, c->prof_ev[3 * k + 1], c->prof_ev[3 * k + 2]));
        p += a;
        m += b;
    }
    *n_calls = c->prof_n;
    *pack_ms_total = p;
    *main_ms_total = m;
    c->prof_n = 0;
    return SM_OK;
}

// ---- step 3 ----------------------------------------------------------------------------

extern "C" int sm_fill_web_holes(sm_ctx *c, int times)
{
    SM_ENTER(c);
    if (!c->have_web) {
        set_error("sm_fill_web_holes: no web (call sm_match_wta first)");
        return SM_ERR_STATE;
    }
    SM_REQUIRE(times >= 0, "sm_fill_web_holes: times must be >= 0");
    SM_REQUIRE(c->row0 == 0 && c->row1 == c->FH, "sm_fill_web_holes: whole-frame contexts only");
    if (!c->web_may_have_holes) {
        // fill_web_holes only ever writes where the web is 0 (stereo.cu:238), and the web that
        // sm_match_wta produces is >= 1 everywhere (some shift always equals the maximum,
        // stereo.c:212-219): every one of the `times` passes is the identity, so
        // the filled web IS the web.  No kernel, no copy.
        c->web_filled = c->web;
        c->have_web2 = true;
        return SM_OK;
    }
    int rc;
    size_t n = c->npix();
    if ((rc = dev_alloc(&c->web2, n)) || (rc = dev_alloc(&c->tmp, n))) return rc;
    // web2 <- web, tmp <- web (stereo.cu:328), then ping-pong (stereo.cu:247-259)
    SM_CUDA(cudaMemcpyAsync(c->web2, c->web, n * 4, cudaMemcpyDeviceToDevice, c->stream));
    SM_CUDA(cudaMemcpyAsync(c->tmp, c->web, n * 4, cudaMemcpyDeviceToDevice, c->stream));
    int32_t *a = c->web2, *b = c->tmp;
    for (int t = 0; t < times; t++) {
        if ((rc = launch_fill_web_holes_step(b, a, c->W, c->FH, c->stream)) < 0) return rc;
        int32_t *s = a;
        a = b;
        b = s;
    }
    c->web2 = a;  // whichever buffer the reference would return as `web`
    c->tmp = b;
    c->web_filled = c->web2;
    c->have_web2 = true;
    return SM_OK;
}

// Step 3 on a caller-supplied web (host, i32): the only way a web with holes (zeros) can
// enter the library; makes fill_web_holes / draw_contour_map usable and testable on their own.
extern "C" int sm_set_web(sm_ctx *c, const int32_t *web)
{
    SM_ENTER(c);
    SM_REQUIRE(web, "sm_set_web: NULL web");
    SM_REQUIRE(c->row0 == 0 && c->row1 == c->FH, "sm_set_web: whole-frame contexts only");
    SM_CUDA(cudaMemcpyAsync(c->web, web, c->npix() * 4, cudaMemcpyHostToDevice, c->stream));
    SM_CUDA(cudaMemsetAsync(c->best, 0, c->npix() * 4, c->stream));
    c->have_web = true;
    c->web_may_have_holes = true;
    c->have_web2 = c->have_out = false;
    return SM_OK;
}

extern "C" int sm_draw_contour_map(sm_ctx *c, int lines, int32_t *web_min, int32_t *web_max)
{
    SM_ENTER(c);
    if (!c->have_web) {
        set_error("sm_draw_contour_map: no web (call sm_match_wta first)");
        return SM_ERR_STATE;
    }
    SM_REQUIRE(c->row0 == 0 && c->row1 == c->FH, "sm_draw_contour_map: whole-frame contexts only");
    const int32_t *web = c->have_web2 ? c->web_filled : c->web;
    int rc;
    if ((rc = dev_alloc(&c->out, c->npix()))) return rc;
    // min/max, its copy to the host and the contour kernel are queued back to back: the kernel takes min and max
    // from the device slot, the host only needs them for the caller and for the degenerate case
    const int cur = c->minmax_cur;
    c->minmax_cur ^= 1;
    const int32_t *slot = c->minmax + 2 * cur;
    if ((rc = launch_minmax(web, c->npix(), c->minmax, cur, c->stream)) < 0) return rc;
    SM_CUDA(cudaMemcpyAsync(c->minmax_host, slot, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    if (lines != 0 && (rc = launch_contour(web, c->npix(), slot, lines, c->out, c->stream)) < 0) return rc;
    SM_CUDA(cudaStreamSynchronize(c->stream));
    const int32_t mm[2] = {c->minmax_host[0], c->minmax_host[1]};
    if (web_min) *web_min = mm[0];
    if (web_max) *web_max = mm[1];
    // interval = (max - min) / num_lines (stereo.cu:276-285 -> stereo.c:265-266)
    if (lines == 0 || (mm[1] - mm[0]) / lines == 0) {
        set_error("sm_draw_contour_map: (max %d - min %d) / lines %d is zero; the reference divides by "
                  "zero here", mm[1], mm[0], lines);
        return SM_ERR_DEGENERATE;
    }
    c->have_out = true;
    return SM_OK;
}

// ---- download ----------------------------------------------------------------------------

extern "C" int sm_download(sm_ctx *c, int which, int shift, void *host)
{
    SM_ENTER(c);
    SM_REQUIRE(host, "sm_download: NULL host pointer");
    int rc = SM_OK;
    switch (which) {
    case SM_EDGES1:
    case SM_EDGES2:
        if (!c->have_edges) goto state;
        rc = copy_band_d2h(c, (uint8_t *)host, c->edges[which - SM_EDGES1]);
        break;
    case SM_MATCH:
    case SM_SCORE_ALL:
    case SM_SCORE: {
        if (!c->have_edges) goto state;
        SM_REQUIRE(shift >= 0 && shift < c->D, "sm_download: shift %d out of [0, %d)", shift, c->D);
        if ((rc = dev_alloc(&c->scratch_u8, c->npix())) || (rc = dev_alloc(&c->scratch_i32, c->npix())))
            return rc;
        rc = launch_pack(c->edges[0], c->edges[1], c->FH, c->row0, c->variant, c->g, c->LA, c->LB, c->RB,
                         c->stream);
        if (rc < 0) return rc;
        HotArgs a = hot_args(c, nullptr, nullptr);
        rc = launch_planes(a, shift, which == SM_MATCH ? c->scratch_u8 : nullptr,
                           which == SM_SCORE_ALL ? c->scratch_i32 : nullptr,
                           which == SM_SCORE ? c->scratch_i32 : nullptr, c->stream);
        if (rc < 0) return rc;
        rc = which == SM_MATCH ? copy_band_d2h(c, (uint8_t *)host, c->scratch_u8)
                               : copy_band_d2h(c, (int32_t *)host, c->scratch_i32);
        break;
    }
    case SM_BEST:
        if (!c->have_web) goto state;
        rc = copy_band_d2h(c, (int32_t *)host, c->best);
        break;
    case SM_WEB:
        if (!c->have_web) goto state;
        rc = copy_band_d2h(c, (int32_t *)host, c->web);
        break;
    case SM_WEB_FILLED:
        if (!c->have_web2) goto state;
        rc = copy_band_d2h(c, (int32_t *)host, c->web_filled);
        break;
    case SM_OUTPUT:
        if (!c->have_out) goto state;
        rc = copy_band_d2h(c, (uint8_t *)host, c->out);
        break;
    default:
        set_error("sm_download: unknown plane %d", which);
        return SM_ERR_ARG;
    }
    if (rc) return rc;
    SM_CUDA(cudaStreamSynchronize(c->stream));
    return SM_OK;
state:
    set_error("sm_download: plane %d has not been computed yet", which);
    return SM_ERR_STATE;
}

static int queue_web_u8(sm_ctx *c, uint8_t *host)
{
    int rc;
    if ((rc = dev_alloc(&c->scratch_u8, c->npix()))) return rc;
    size_t off = (size_t)c->row0 * c->W, n = (size_t)(c->row1 - c->row0) * c->W;
    if ((rc = launch_i32_to_u8(c->web + off, c->scratch_u8 + off, n, c->stream)) < 0) return rc;
    return copy_band_d2h(c, host, c->scratch_u8);
}

extern "C" int sm_download_web_u8(sm_ctx *c, uint8_t *host)
{
    SM_ENTER(c);
    SM_REQUIRE(host, "sm_download_web_u8: NULL host pointer");
    SM_REQUIRE(c->D <= 255, "sm_download_web_u8: num_shifts %d does not fit 8 bits", c->D);
    if (!c->have_web) {
        set_error("sm_download_web_u8: no web (call sm_match_wta first)");
        return SM_ERR_STATE;
    }
    int rc = queue_web_u8(c, host);
    if (rc) return rc;
    SM_CUDA(cudaStreamSynchronize(c->stream));
    return SM_OK;
}

// ---- whole pairs, batched ----------------------------------------------------------------

extern "C" int sm_run_batch(sm_ctx *c, int n_pairs, const uint8_t *first, const uint8_t *second,
                            double threshold, void *web_out, int web_u8, int32_t *best_out)
{
    SM_ENTER(c);
    SM_REQUIRE(n_pairs >= 0 && first && second && web_out, "sm_run_batch: bad arguments");
    SM_REQUIRE(c->row0 == 0 && c->row1 == c->FH, "sm_run_batch: whole-frame contexts only");
    SM_REQUIRE(!web_u8 || c->D <= 255, "sm_run_batch: u8 web needs num_shifts <= 255");
    SM_REQUIRE(threshold >= 0.0 && threshold <= 1.0, "sm_run_batch: threshold must be between 0 and 1");
    constexpr int NP = sm_ctx::NPIPE;
    sm_ctx::Pipe &P = c->pipe;
    const size_t n = c->npix();
    int rc;
    if (!P.ready) {  // resumable set-up, ready flag last (see sm_match_wta_dev_batch)
        // pairs per stage: enough for the batched hot path to run in throughput mode, bounded
        // so that the three buffer sets stay small against HBM (16 B per pixel and pair)
        if (P.group == 0) {
            P.group = n <= ((size_t)1 << 22) ? 16 : (n <= ((size_t)1 << 24) ? 8 : 2);
            if (c->pipe_group_opt > 0) P.group = c->pipe_group_opt;
        }
        if (!P.up) SM_CUDA(cudaStreamCreateWithFlags(&P.up, cudaStreamNonBlocking));
        if (!P.down) SM_CUDA(cudaStreamCreateWithFlags(&P.down, cudaStreamNonBlocking));
        for (int k = 0; k < NP; k++) {
            if ((rc = dev_alloc(&P.img[k], 2 * P.group * n)) || (rc = dev_alloc(&P.web[k], P.group * n)) ||
                (rc = dev_alloc(&P.best[k], P.group * n)))
                return rc;
            if (!P.ev_up[k]) SM_CUDA(cudaEventCreateWithFlags(&P.ev_up[k], cudaEventDisableTiming));
            if (!P.ev_comp[k]) SM_CUDA(cudaEventCreateWithFlags(&P.ev_comp[k], cudaEventDisableTiming));
            if (!P.ev_down[k]) SM_CUDA(cudaEventCreateWithFlags(&P.ev_down[k], cudaEventDisableTiming));
        }
        P.ready = true;
    }
    if (web_u8)
        for (int k = 0; k < NP; k++)
            if ((rc = dev_alloc(&P.web8[k], P.group * n))) return rc;
    const bool use_lut = !c->edges_fp64_only;
    if (!use_lut)  // byte edge maps exist only on the FP64 cross-check path
        for (int k = 0; k < NP; k++)
            if ((rc = dev_alloc(&P.edg[k], 2 * P.group * n))) return rc;
    if (use_lut && c->lut_threshold != threshold) {
        if ((rc = dev_alloc(&c->edge_lut, edge_lut_words()))) return rc;
        if ((rc = launch_edge_lut(threshold, c->edge_lut, c->stream)) < 0) return rc;
        c->lut_threshold = threshold;
    }
    const int G = P.group;
    int stage = 0, launches = 0;
    for (int k = 0; k < n_pairs; k += G, stage++) {
        const int np = n_pairs - k < G ? n_pairs - k : G;
        const int b = stage % NP;
        // ---- stage 1, upload stream: the group's images (contiguous in the caller's arrays).
        // Buffer set b is free once the download of the group that used it last has finished.
        if (stage >= NP) SM_CUDA(cudaStreamWaitEvent(P.up, P.ev_down[b], 0));
        SM_CUDA(cudaMemcpyAsync(P.img[b], first + (size_t)k * n, (size_t)np * n, cudaMemcpyHostToDevice, P.up));
        SM_CUDA(cudaMemcpyAsync(P.img[b] + (size_t)G * n, second + (size_t)k * n, (size_t)np * n,
                                cudaMemcpyHostToDevice, P.up));
        SM_CUDA(cudaEventRecord(P.ev_up[b], P.up));
        // ---- stage 2, the context's stream: edges of all 2*np images, then the batched hot path
        SM_CUDA(cudaStreamWaitEvent(c->stream, P.ev_up[b], 0));
        if (use_lut) {
            // edges of all 2*np images straight into the packed planes, then the hot path: two kernels per group
            if ((rc = batch_core(c, np, true, threshold, P.img[b], P.img[b] + (size_t)G * n, n, P.best[b], P.web[b], n)))
                return rc;
        } else {
            for (int j = 0; j < 2 * G; j++) {
                if (j % G >= np) continue;
                rc = launch_edges<uint8_t>(P.img[b] + (size_t)j * n, c->W, c->FH, 0, c->FH, c->variant, threshold,
                                           P.edg[b] + (size_t)j * n, c->stream);
                if (rc < 0) return rc;
                launches += rc;
            }
            if ((rc = batch_core(c, np, false, 0.0, P.edg[b], P.edg[b] + (size_t)G * n, n, P.best[b], P.web[b], n)))
                return rc;
        }
        launches += c->last_launches;
        if (web_u8) {
            if ((rc = launch_i32_to_u8(P.web[b], P.web8[b], (size_t)np * n, c->stream)) < 0) return rc;
            launches += rc;
        }
        SM_CUDA(cudaEventRecord(P.ev_comp[b], c->stream));
        // ---- stage 3, download stream
        SM_CUDA(cudaStreamWaitEvent(P.down, P.ev_comp[b], 0));
        if (web_u8)
            SM_CUDA(cudaMemcpyAsync((uint8_t *)web_out + (size_t)k * n, P.web8[b], (size_t)np * n,
                                    cudaMemcpyDeviceToHost, P.down));
        else
            SM_CUDA(cudaMemcpyAsync((int32_t *)web_out + (size_t)k * n, P.web[b], (size_t)np * n * sizeof(int32_t),
                                    cudaMemcpyDeviceToHost, P.down));
        if (best_out)
            SM_CUDA(cudaMemcpyAsync(best_out + (size_t)k * n, P.best[b], (size_t)np * n * sizeof(int32_t),
                                    cudaMemcpyDeviceToHost, P.down));
        SM_CUDA(cudaEventRecord(P.ev_down[b], P.down));
    }
    SM_CUDA(cudaStreamSynchronize(P.up));
    SM_CUDA(cudaStreamSynchronize(c->stream));
    SM_CUDA(cudaStreamSynchronize(P.down));
    c->last_launches = launches;
    return SM_OK;
}

// ---- whole pairs over several GPUs of one box ------------------------------------------------
//
// pairs shard with no exchange step.  One context and one host thread per device; device d
// takes a contiguous range of the batch (the caller's arrays are contiguous, so every stage of its
// sm_run_batch pipeline stays a single copy).  No cross-device traffic, no collective.

// One persistent host thread per device slot (sm_multi_*, sm_bands_*): created with the object, parked on a
// condition variable between calls, joined at destroy -- no thread is created or joined inside a run call.
namespace {

class SlotWorkers {
public:
    explicit SlotWorkers(int n) : jobs_((size_t)n), state_((size_t)n, 0)
    {
        for (int k = 0; k < n; k++) {
            try {
                threads_.emplace_back([this, k] { loop(k); });
            } catch (...) {  // no thread to be had: that slot's work runs on the calling thread
                break;
            }
        }
    }
    ~SlotWorkers()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            quit_ = true;
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    // run job(k) for every slot k and wait for all of them
    void run_all(const std::function<void(int)> &job)
    {
        const int n = (int)jobs_.size(), nt = (int)threads_.size();
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (int k = 0; k < nt; k++) {
                jobs_[(size_t)k] = job;
                state_[(size_t)k] = 1;
            }
        }
        cv_.notify_all();
        for (int k = nt; k < n; k++) job(k);
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] {
            for (int k = 0; k < nt; k++)
                if (state_[(size_t)k] != 0) return false;
            return true;
        });
    }

private:
    void loop(int k)
    {
        for (;;) {
            std::function<void(int)> job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return quit_ || state_[(size_t)k] == 1; });
                if (quit_) return;
                job = jobs_[(size_t)k];
                state_[(size_t)k] = 2;
            }
            job(k);
            {
                std::lock_guard<std::mutex> lk(mu_);
                state_[(size_t)k] = 0;
            }
            done_.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_, done_;
    std::vector<std::function<void(int)>> jobs_;
    std::vector<int> state_;  // 0 idle, 1 job posted, 2 running
    std::vector<std::thread> threads_;
    bool quit_ = false;
};

}  // namespace

struct sm_multi {
    int n = 0;
    sm_ctx **ctx = nullptr;
    SlotWorkers *workers = nullptr;
};

extern "C" int sm_multi_create(sm_multi **out, const int *devices, int n_devices, int width, int height,
                               int num_shifts, int square_width, int variant)
{
    SM_REQUIRE(out && devices && n_devices >= 1 && n_devices <= 64, "sm_multi_create: bad arguments");
    *out = nullptr;
    sm_multi *m = new (std::nothrow) sm_multi;
    if (!m) {
        set_error("sm_multi_create: out of memory");
        return SM_ERR_NOMEM;
    }
    m->ctx = (sm_ctx **)calloc((size_t)n_devices, sizeof(sm_ctx *));
    if (!m->ctx) {
        delete m;
        set_error("sm_multi_create: out of memory");
        return SM_ERR_NOMEM;
    }
    m->n = n_devices;
    for (int d = 0; d < n_devices; d++) {
        int rc = sm_create(&m->ctx[d], devices[d], width, height, num_shifts, square_width, variant);
        if (rc) {
            for (int k = 0; k < d; k++) sm_destroy(m->ctx[k]);
            free(m->ctx);
            delete m;
            return rc;
        }
    }
    m->workers = new (std::nothrow) SlotWorkers(n_devices);
    *out = m;
    return SM_OK;
}

extern "C" int sm_multi_destroy(sm_multi *m)
{
    if (!m) return SM_OK;
    delete m->workers;
    for (int d = 0; d < m->n; d++) sm_destroy(m->ctx[d]);
    free(m->ctx);
    delete m;
    return SM_OK;
}

extern "C" int sm_multi_device_count(const sm_multi *m) { return m ? m->n : SM_ERR_ARG; }

extern "C" int sm_multi_run_batch(sm_multi *m, int n_pairs, const uint8_t *first, const uint8_t *second,
                                  double threshold, void *web_out, int web_u8, int32_t *best_out)
{
    SM_REQUIRE(m && n_pairs >= 0 && first && second && web_out, "sm_multi_run_batch: bad arguments");
    const int N = m->n;
    std::vector<int> rcs((size_t)N, SM_OK);
    std::vector<std::string> errs((size_t)N);
    const size_t n = m->ctx[0]->npix();
    auto work = [&](int d) {
        const int k0 = (int)((long long)n_pairs * d / N), k1 = (int)((long long)n_pairs * (d + 1) / N);
        if (k1 <= k0) return;
        uint8_t *w8 = (uint8_t *)web_out;
        void *wout = web_u8 ? (void *)(w8 + (size_t)k0 * n) : (void *)((int32_t *)web_out + (size_t)k0 * n);
        rcs[(size_t)d] = sm_run_batch(m->ctx[d], k1 - k0, first + (size_t)k0 * n, second + (size_t)k0 * n, threshold, wout,
                                      web_u8, best_out ? best_out + (size_t)k0 * n : nullptr);
        if (rcs[(size_t)d]) errs[(size_t)d] = sm_last_error();  // the message is per thread: carry it to the caller's
    };
    if (m->workers)
        m->workers->run_all(work);
    else
        for (int d = 0; d < N; d++) work(d);
    for (int d = 0; d < N; d++)
        if (rcs[d]) {
            set_error("sm_multi_run_batch: device slot %d: %s", d, errs[d].c_str());
            return rcs[d];
        }
    return SM_OK;
}

// ---- one pair, row bands over several GPUs of one box ----------------------------------------------
//
// SURVEY 8e, BASELINE config 3: GPU g owns output rows sm_band_rows(height, N, g); it uploads those rows of
// both images plus half + 1 halo rows per side (wrapped for WRAP, clipped for GHOST), detects the edges of its
// rows plus half, runs the hot path on its band and writes only its own rows of web / best into the caller's
// frame-sized host arrays.  One band context and one host thread per device slot; no exchange step.

struct sm_bands {
    int n = 0;
    sm_ctx **ctx = nullptr;
    SlotWorkers *workers = nullptr;
};

extern "C" int sm_bands_destroy(sm_bands *b)
{
    if (!b) return SM_OK;
    delete b->workers;
    for (int d = 0; d < b->n; d++) sm_destroy(b->ctx[d]);
    free(b->ctx);
    delete b;
    return SM_OK;
}

extern "C" int sm_bands_create(sm_bands **out, const int *devices, int n_devices, int width, int height,
                               int num_shifts, int square_width, int variant)
{
    SM_REQUIRE(out && devices && n_devices >= 1 && n_devices <= 64, "sm_bands_create: bad arguments");
    SM_REQUIRE(height >= n_devices, "sm_bands_create: fewer rows than bands");
    *out = nullptr;
    sm_bands *b = new (std::nothrow) sm_bands;
    if (!b) {
        set_error("sm_bands_create: out of memory");
        return SM_ERR_NOMEM;
    }
    b->ctx = (sm_ctx **)calloc((size_t)n_devices, sizeof(sm_ctx *));
    if (!b->ctx) {
        delete b;
        set_error("sm_bands_create: out of memory");
        return SM_ERR_NOMEM;
    }
    b->n = n_devices;
    for (int d = 0; d < n_devices; d++) {
        int r0 = 0, r1 = 0;
        int rc = sm_band_rows(height, n_devices, d, &r0, &r1);
        if (!rc) rc = sm_create_band(&b->ctx[d], devices[d], width, height, r0, r1, num_shifts, square_width, variant);
        if (rc) {
            sm_bands_destroy(b);  // sm_destroy(NULL) is a no-op for the slots not created yet
            return rc;
        }
    }
    b->workers = new (std::nothrow) SlotWorkers(n_devices);
    *out = b;
    return SM_OK;
}

extern "C" int sm_bands_run(sm_bands *b, const uint8_t *first, const uint8_t *second, double threshold,
                            int32_t *web_out, int32_t *best_out)
{
    SM_REQUIRE(b && first && second && web_out, "sm_bands_run: bad arguments");
    const int N = b->n;
    std::vector<int> rcs((size_t)N, SM_OK);
    std::vector<std::string> errs((size_t)N);
    auto work = [&](int d) {
        sm_ctx *c = b->ctx[d];
        int rc = sm_upload_u8(c, first, second);  // a band context copies only the rows it needs
        if (!rc) rc = sm_edges(c, threshold);
        if (!rc) rc = sm_match_wta(c);
        if (!rc) rc = sm_download(c, SM_WEB, 0, web_out);  // ... and writes only its own rows
        if (!rc && best_out) rc = sm_download(c, SM_BEST, 0, best_out);
        rcs[(size_t)d] = rc;
        if (rc) errs[(size_t)d] = sm_last_error();
    };
    if (b->workers)
        b->workers->run_all(work);
    else
        for (int d = 0; d < N; d++) work(d);
    for (int d = 0; d < N; d++)
        if (rcs[d]) {
            set_error("sm_bands_run: band %d: %s", d, errs[d].c_str());
            return rcs[d];
        }
    return SM_OK;
}
