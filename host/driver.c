/*
 * driver.c -- stereopar / stereopar-ghost: the reference's CUDA programs with the CUDA part
 * replaced by libstereo_b200.so.  Plain C over the C ABI (include/stereo_b200.h).
 *
 * It keeps the observable contract of src/stereo.cu and src/stereo-ghost.cu:
 *   - argv: image1 image2 [threshold] [square_width] [times] [lines], the same defaults
 *     and the same validation messages and exit codes (stereo.cu:350-398);
 *   - stdout: "width = %d, height = %d, t1 = %f, t2 = %f, elapsed = %f\n" (stereo.cu:336),
 *     field 15 of which test/time.sh reads;
 *   - unless built with -DNO_WRITES, the 96 PPM files per run that the reference writes
 *     (edges-{1,2}, matches-i, score_all-i, scores-i, score_best-0, web-{1,2}, output-0), into
 *     par/ or pargh/ under -DDEBUG, so that test/diff.sh can compare them with the serial
 *     programs' ser/ and sergh/;
 *   - allocation / CUDA failures print a message and exit(1) (util.h:49-58,
 *     helper_cuda.h:890-905): the library only returns codes, the policy lives here.
 *
 * Build: -DSM_VARIANT=0 -> stereopar (wrap-around), -DSM_VARIANT=1 -> stereopar-ghost.
 * -DREF_HOST (the recipe lives with the test infrastructure, target refhost): the host side is the REFERENCE'S OWN -- src/image.c with the
 * vendored stb_image.h loader, Image.data as double, its write_image and util.h -- compiled where it lies;
 * the images then go up through sm_upload_f64 in the reference's double layout.  That build proves the
 * boundary of INTEGRATION.md section 2 with the reference's loader and writer instead of host/hostimage.c.
 * NUM_SHIFTS is the reference's compile-time constant (stereo.cu:6); here it can also be
 * set at run time with the environment variable STEREO_NUM_SHIFTS.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifdef REF_HOST
#include "image.h" /* the reference's (-I<reference>/src): Image, read_image, make_filename, write_image; brings util.h */
typedef Image HostImage;
#define WRITE_IMAGE(data, w, h, type, file) write_image((void *)(data), (w), (h), 0, (type), (file))
#define UPLOAD(ctx, a, b) sm_upload_f64((ctx), (a).data, (b).data) /* Image.data: double = u8/256.0, image.c:13 */
#else
#include "hostimage.h"
typedef Image8 HostImage;
#define WRITE_IMAGE(data, w, h, type, file) write_image((data), (w), (h), (type), (file))
#define UPLOAD(ctx, a, b) sm_upload_u8((ctx), (a).data, (b).data)
#endif
#include "stereo_b200.h"

#ifndef SM_VARIANT
#define SM_VARIANT 0
#endif
#ifndef NUM_SHIFTS
#define NUM_SHIFTS 30
#endif
#define DEFAULT_THRESHOLD 0.15
#define DEFAULT_SQUARE_WIDTH 21
#define DEFAULT_TIMES 32
#define DEFAULT_LINES 10

#define PROGRAM_TYPE (SM_VARIANT == 0 ? PAR : PARGHOST)

#ifndef REF_HOST /* util.h of the reference has its own get_time, xmalloc and parse_* (util.h:49-75,104-109) */
static double get_time(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + (double)ts.tv_nsec / 1e9;
}

#ifndef NO_WRITES
static void *xmalloc(size_t size)
{
    void *p = malloc(size ? size : 1);
    if (!p) {
        fprintf(stderr, "error: out of memory\n");
        exit(1);
    }
    memset(p, 0, size);
    return p;
}
#endif

/* strtod/strtol with the reference's "is it a number" rule (util.h:63-75) */
static int parse_double(const char *s, double *n)
{
    char *end;
    *n = strtod(s, &end);
    return *n == 0 && end == s;
}

static int parse_int(const char *s, int *n)
{
    char *end;
    *n = (int)strtol(s, &end, 0);
    return *n == 0 && end == s;
}
#endif /* !REF_HOST */

/* checkCudaErrors' policy (helper_cuda.h:890-905): report and stop */
#define CHECK(call)                                                                            \
    do {                                                                                       \
        int rc__ = (call);                                                                     \
        if (rc__ != SM_OK) {                                                                   \
            fprintf(stderr, "error: %s failed (%d): %s\n", #call, rc__, sm_last_error());      \
            exit(EXIT_FAILURE);                                                                \
        }                                                                                      \
    } while (0)

typedef struct AlgorithmParams {
    double threshold;
    int square_width;
    int times;
    int lines_to_draw;
} AlgorithmParams;

#ifndef NO_WRITES
static void dump(sm_ctx *ctx, int which, int shift, void *host, int w, int h, ImageType type, const char *name,
                 int number)
{
    CHECK(sm_download(ctx, which, shift, host));
    WRITE_IMAGE(host, w, h, type, make_filename(name, PROGRAM_TYPE, number));
}
#endif

/* algorithm() of the reference (stereo.cu:296-347), stage for stage */
static void algorithm(sm_ctx *ctx, int width, int height, int num_shifts, AlgorithmParams params)
{
#ifndef NO_WRITES
    uint8_t *h8 = (uint8_t *)xmalloc((size_t)width * height);
    int32_t *h32 = (int32_t *)xmalloc(sizeof(int32_t) * (size_t)width * height);
#else
    (void)num_shifts;
#endif
    double t1 = get_time();

    /* first step: find edges in both images */
    CHECK(sm_edges(ctx, params.threshold));
#ifndef NO_WRITES
    dump(ctx, SM_EDGES1, 0, h8, width, height, IMTYPE_BINARY, "edges", 1);
    dump(ctx, SM_EDGES2, 0, h8, width, height, IMTYPE_BINARY, "edges", 2);
#endif

    /* second step: match edges between images.  The library never materialises matches[] /
     * scores[]; the debug planes are produced on demand for the dumps only. */
    CHECK(sm_match_wta(ctx));
#ifndef NO_WRITES
    for (int i = 0; i < num_shifts; i++) dump(ctx, SM_MATCH, i, h8, width, height, IMTYPE_BINARY, "matches", i);
    for (int i = 0; i < num_shifts; i++) dump(ctx, SM_SCORE_ALL, i, h32, width, height, IMTYPE_GRAY_INT, "score_all", i);
    for (int i = 0; i < num_shifts; i++) dump(ctx, SM_SCORE, i, h32, width, height, IMTYPE_GRAY_INT, "scores", i);
    dump(ctx, SM_BEST, 0, h32, width, height, IMTYPE_GRAY_INT, "score_best", 0);
    dump(ctx, SM_WEB, 0, h32, width, height, IMTYPE_GRAY_INT, "web", 1);
#endif

    /* third step: draw contour lines */
    CHECK(sm_fill_web_holes(ctx, params.times));
#ifndef NO_WRITES
    dump(ctx, SM_WEB_FILLED, 0, h32, width, height, IMTYPE_GRAY_INT, "web", 2);
#endif
    int rc = sm_draw_contour_map(ctx, params.lines_to_draw, NULL, NULL);
    if (rc == SM_ERR_DEGENERATE) {
        /* the reference divides by zero here (stereo.c:265-272); say so instead of trapping */
        fprintf(stderr, "error: %s\n", sm_last_error());
        exit(EXIT_FAILURE);
    }
    CHECK(rc);
#ifndef NO_WRITES
    dump(ctx, SM_OUTPUT, 0, h8, width, height, IMTYPE_BINARY, "output", 0);
#endif

    CHECK(sm_synchronize(ctx));
    double t2 = get_time();
    double elapsed = t2 - t1;
    printf("width = %d, height = %d, t1 = %f, t2 = %f, elapsed = %f\n", width, height, t1, t2, elapsed);
#ifndef NO_WRITES
    free(h8);
    free(h32);
#endif
}

int main(int argc, char *argv[])
{
    if (argc < 3) {
        fprintf(stderr,
                "usage: stereomatch [image 1] [image 2] [threshold = %g] "
                "[square_width = %d] [times = %d] [lines = %d]\n",
                DEFAULT_THRESHOLD, DEFAULT_SQUARE_WIDTH, DEFAULT_TIMES, DEFAULT_LINES);
        return 1;
    }

    HostImage first, second;
    if (read_image(argv[1], &first)) return 1;
    if (read_image(argv[2], &second)) return 1;
    if (first.width != second.width || first.height != second.height) {
        fprintf(stderr, "error: the two images must have equal width and height\n");
        return 1;
    }

    AlgorithmParams params = {.threshold = DEFAULT_THRESHOLD,
                              .square_width = DEFAULT_SQUARE_WIDTH,
                              .times = DEFAULT_TIMES,
                              .lines_to_draw = DEFAULT_LINES};

    if (argc >= 4 && parse_double(argv[3], &params.threshold)) {
        fprintf(stderr, "error: threshold must be a number\n");
        return 1;
    }
    if (argc >= 5 && parse_int(argv[4], &params.square_width)) {
        fprintf(stderr, "error: square_width must be a number\n");
        return 1;
    }
    if (argc >= 6 && parse_int(argv[5], &params.times)) {
        fprintf(stderr, "error: times must be a number\n");
        return 1;
    }
    if (argc >= 7 && parse_int(argv[6], &params.lines_to_draw)) {
        fprintf(stderr, "error: lines must be a number\n");
        return 1;
    }

    if (params.threshold < 0.0 || params.threshold > 1.0) {
        fprintf(stderr, "error: threshold must be between 0 and 1\n");
        return 1;
    }
    if (params.square_width > first.width || params.square_width > first.height) {
        fprintf(stderr, "error: square width must not be higher than image width/height\n");
        return 1;
    }

    int num_shifts = NUM_SHIFTS;
    const char *env = getenv("STEREO_NUM_SHIFTS");
    if (env && parse_int(env, &num_shifts)) {
        fprintf(stderr, "error: STEREO_NUM_SHIFTS must be a number\n");
        return 1;
    }

    /* MAKE_GPU_COPY x2 + the allocations at the top of algorithm() (stereo.cu:299-306,402-403):
     * outside the timed region, as in the reference */
    sm_ctx *ctx = NULL;
    CHECK(sm_create(&ctx, 0, first.width, first.height, num_shifts, params.square_width,
                    SM_VARIANT == 0 ? SM_WRAP : SM_GHOST));
    CHECK(UPLOAD(ctx, first, second));
    CHECK(sm_synchronize(ctx));

    algorithm(ctx, first.width, first.height, num_shifts, params);

    CHECK(sm_destroy(ctx));
    free(first.data);
    free(second.data);
    return 0;
}
