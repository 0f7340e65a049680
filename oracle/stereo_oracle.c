/*
 * stereo_oracle.c -- CPU restatement of chrg127/stereomatching's algorithm.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product path (stereomatching_b200/, host/) never links or
 * calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here
 * against (a) the CRC32 table recorded from the unmodified reference
 * (tests/golden/golden.json) and (b) the reference itself,
 * compiled where it lies by oracle/Makefile into oracle/_ref/.
 *
 * All file:line citations are relative to the reference checkout
 * (/root/reference).  Nothing here is copied from it; each function restates
 * what the cited lines compute.
 *
 * variant: 0 = WRAP  (src/stereo.c, toroidal idx(), src/util.h:42-47)
 *          1 = GHOST (src/stereo-ghost.c, padded arrays, src/ghost.h:54-55)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORACLE_WRAP 0
#define ORACLE_GHOST 1

/* ------------------------------------------------------------------ */
/* step 1: edge detection                                              */
/* ------------------------------------------------------------------ */

/* Brightness sample as the reference sees it.
 * WRAP : brightness[idx(x,y,w,h)], idx wraps both coordinates (util.h:42-47).
 * GHOST: the image is padded by one cell of 128.0 (stereo-ghost.c:384-385),
 *        so any coordinate outside [0,w)x[0,h) reads 128.0.
 * Pixel values are u8/256.0 (image.c:9-15). */
static double bright(const uint8_t *img, int w, int h, int x, int y, int variant)
{
    if (variant == ORACLE_WRAP) {
        x = (x + w) % w;
        y = (y + h) % h;
    } else if (x < 0 || y < 0 || x >= w || y >= h) {
        return 128.0;
    }
    return img[(size_t)y * w + x] / 256.0;
}

/* One directional detector: two 3-pixel averages, their mean, and the
 * thresholded absolute difference (stereo.c:16-28; the other three detectors
 * stereo.c:30-70 differ only in which six neighbours they read).
 * CLAMP(t*overall, 0, 1) is MIN(MAX(x,0),1) (util.h:24-26).
 * Operation order is kept: ((a+b)+c)/3.0, (l+r)/2.0. */
static int detect(double a0, double a1, double a2, double b0, double b1, double b2,
                  double threshold)
{
    double avg_left = (a0 + a1 + a2) / 3.0;
    double avg_right = (b0 + b1 + b2) / 3.0;
    double overall = (avg_left + avg_right) / 2.0;
    double lim = threshold * overall;
    lim = lim > 0.0 ? lim : 0.0;
    lim = lim < 1.0 ? lim : 1.0;
    return fabs(avg_left - avg_right) > lim;
}

/* find_all_edges (stereo.c:72-84, stereo-ghost.c:74-85): OR of the four
 * detectors.  `edges` is the plain w*h array (the GHOST variant's padding of
 * NUM_SHIFTS zero cells, stereo-ghost.c:286-287, is implied: readers treat
 * out-of-image edge cells as 0). */
void oracle_edges(const uint8_t *img, int w, int h, double threshold, int variant,
                  uint8_t *edges)
{
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
#define B(dx, dy) bright(img, w, h, x + (dx), y + (dy), variant)
            int e =
                /* left_right, stereo.c:16-28 */
                detect(B(-1, -1), B(-1, 0), B(-1, 1), B(1, -1), B(1, 0), B(1, 1), threshold)
                /* top_bottom, stereo.c:30-42 */
                || detect(B(-1, -1), B(0, -1), B(1, -1), B(-1, 1), B(0, 1), B(1, 1), threshold)
                /* upleft_downright, stereo.c:44-56 */
                || detect(B(-1, -1), B(0, -1), B(-1, 0), B(1, 0), B(0, 1), B(1, 1), threshold)
                /* downleft_upright, stereo.c:58-70 */
                || detect(B(-1, 1), B(0, 1), B(-1, 0), B(0, -1), B(1, -1), B(1, 0), threshold);
#undef B
            edges[(size_t)y * w + x] = (uint8_t)e;
        }
    }
}

/* ------------------------------------------------------------------ */
/* step 2: matches -> window scores -> winner-take-all                 */
/* ------------------------------------------------------------------ */

/* matches[i][x,y] (stereo.c:113-127; stereo-ghost.c:113-126).
 * WRAP : right index x+i wraps mod w.
 * GHOST: right edge map is padded with zeros, so x+i >= w reads 0. */
static inline int match_at(const uint8_t *le, const uint8_t *re, int w, int x, int y, int i,
                           int variant)
{
    int xr = x + i;
    int r;
    if (variant == ORACLE_WRAP)
        r = re[(size_t)y * w + (xr % w)];
    else
        r = xr < w ? re[(size_t)y * w + xr] : 0;
    return le[(size_t)y * w + x] == r;
}

void oracle_match_plane(const uint8_t *le, const uint8_t *re, int w, int h, int i, int variant,
                        uint8_t *m)
{
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            m[(size_t)y * w + x] = (uint8_t)match_at(le, re, w, x, y, i, variant);
}

/* addup_pixels_in_square, literal form (stereo.c:132-148; ghost:131-147):
 * (2*half+1)^2 taps, half = square_width/2.  WRAP wraps both coordinates;
 * GHOST taps outside the image read the zero padding of the match image
 * (stereo-ghost.c:93-97), i.e. contribute nothing. */
void oracle_box_direct(const uint8_t *m, int w, int h, int sw, int variant, int32_t *total)
{
    int half = sw / 2;
    memset(total, 0, sizeof(int32_t) * (size_t)w * h);
    for (int sy = -half; sy <= half; sy++)
        for (int sx = -half; sx <= half; sx++)
            for (int y = 0; y < h; y++)
                for (int x = 0; x < w; x++) {
                    int xx = x + sx, yy = y + sy;
                    if (variant == ORACLE_WRAP) {
                        xx = (xx + w) % w;
                        yy = (yy + h) % h;
                    } else if (xx < 0 || yy < 0 || xx >= w || yy >= h) {
                        continue;
                    }
                    total[(size_t)y * w + x] += m[(size_t)yy * w + xx];
                }
}

/* Same sums through separable passes (rows, then columns).  All-integer, so
 * identical to oracle_box_direct; tests/test_oracle.py asserts that. */
void oracle_box_fast(const uint8_t *m, int w, int h, int sw, int variant, int32_t *total)
{
    int half = sw / 2;
    int32_t *rows = (int32_t *)malloc(sizeof(int32_t) * (size_t)w * h);
    int n = w + 2 * half;
    int32_t *p = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 1));
    int32_t *acc = (int32_t *)calloc((size_t)w, sizeof(int32_t));
    /* horizontal window sums through a prefix sum over the extended row */
    for (int y = 0; y < h; y++) {
        const uint8_t *r = m + (size_t)y * w;
        int32_t *o = rows + (size_t)y * w;
        p[0] = 0;
        for (int k = 0; k < n; k++) {
            int xx = k - half, v;
            if (variant == ORACLE_WRAP)
                v = r[((xx % w) + w) % w];
            else
                v = (xx < 0 || xx >= w) ? 0 : r[xx];
            p[k + 1] = p[k] + v;
        }
        for (int x = 0; x < w; x++)
            o[x] = p[x + 2 * half + 1] - p[x];
    }
    /* vertical window sums: acc holds rows y-half .. y+half */
#define ROWPTR(yy)                                                                   \
    (variant == ORACLE_WRAP ? rows + (size_t)((((yy) % h) + h) % h) * w              \
                            : ((yy) < 0 || (yy) >= h ? NULL : rows + (size_t)(yy)*w))
    for (int yy = -half; yy <= half; yy++) {
        const int32_t *r = ROWPTR(yy);
        if (r)
            for (int x = 0; x < w; x++) acc[x] += r[x];
    }
    for (int y = 0; y < h; y++) {
        int32_t *o = total + (size_t)y * w;
        for (int x = 0; x < w; x++) o[x] = acc[x];
        const int32_t *out = ROWPTR(y - half), *in = ROWPTR(y + half + 1);
        if (out)
            for (int x = 0; x < w; x++) acc[x] -= out[x];
        if (in)
            for (int x = 0; x < w; x++) acc[x] += in[x];
    }
#undef ROWPTR
    free(acc);
    free(p);
    free(rows);
}

/* One shift's three debug planes: matches-i, score_all-i (the box sum before
 * masking, stereo.c:188-189) and scores-i (record_score, stereo.c:172-182:
 * the sum where the centre matched, else the xmalloc zero, util.h:56). */
void oracle_shift_planes(const uint8_t *le, const uint8_t *re, int w, int h, int sw, int i,
                         int variant, int direct, uint8_t *m, int32_t *score_all, int32_t *score)
{
    oracle_match_plane(le, re, w, h, i, variant, m);
    if (direct)
        oracle_box_direct(m, w, h, sw, variant, score_all);
    else
        oracle_box_fast(m, w, h, sw, variant, score_all);
    for (size_t p = 0; p < (size_t)w * h; p++)
        score[p] = m[p] == 1 ? score_all[p] : 0;
}

/* The whole hot path: fillup_matches + fillup_scores +
 * find_highest_scoring_shifts (stereo.c:306-312).
 * best  = max over shifts of the masked score, starting from the memset 0
 *         (stereo.c:311, 203-209);
 * web   = i+1 for the LAST i whose score equals best (stereo.c:212-219:
 *         ascending i, unconditional overwrite on equality), so ties go to
 *         the highest shift and an all-zero pixel ends at num_shifts.
 * direct != 0 uses the literal sw*sw tap loop. */
void oracle_match_wta(const uint8_t *le, const uint8_t *re, int w, int h, int num_shifts, int sw,
                      int variant, int direct, int32_t *best, int32_t *web)
{
    size_t n = (size_t)w * h;
    uint8_t *m = (uint8_t *)malloc(n);
    int32_t *all = (int32_t *)malloc(sizeof(int32_t) * n);
    int32_t *sc = (int32_t *)malloc(sizeof(int32_t) * n);
    memset(best, 0, sizeof(int32_t) * n);
    memset(web, 0, sizeof(int32_t) * n);
    for (int i = 0; i < num_shifts; i++) {
        oracle_shift_planes(le, re, w, h, sw, i, variant, direct, m, all, sc);
        for (size_t p = 0; p < n; p++) {
            if (sc[p] >= best[p]) { /* equivalent single pass, SURVEY 0 */
                best[p] = sc[p];
                web[p] = i + 1;
            }
        }
    }
    free(m);
    free(all);
    free(sc);
}

/* ------------------------------------------------------------------ */
/* step 3 (next rows n3): hole filling and contour drawing             */
/* ------------------------------------------------------------------ */

/* fill_web_holes (stereo.c:230-252).  web is never 0 after step 2, so the
 * body never fires; the oracle keeps the loop but guards the unwrapped
 * neighbour reads (stereo.c:239-242 would read out of bounds at borders). */
void oracle_fill_web_holes(int32_t *web, int w, int h, int times)
{
    size_t n = (size_t)w * h;
    int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * n);
    int32_t *a = web, *b = tmp;
    memcpy(tmp, web, sizeof(int32_t) * n);
    for (int t = 0; t < times; t++) {
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                size_t p = (size_t)y * w + x;
                if (b[p] == 0) {
                    /* the reference indexes with the UNWRAPPED IDX(x+-1, y) (stereo.c:237-243): at a row end the
                     * "right" neighbour is the first pixel of the next row and vice versa.  Only the reads that
                     * leave the array altogether (first/last pixel, rows above the first and below the last) are
                     * undefined there; they count as 0 here. */
                    int32_t r = p + 1 < n ? b[p + 1] : 0, l = p > 0 ? b[p - 1] : 0;
                    int32_t u = y + 1 < h ? b[p + w] : 0, d = y > 0 ? b[p - w] : 0;
                    a[p] = (r + u + l + d) / 4;
                }
            }
        int32_t *s = a;
        a = b;
        b = s;
    }
    /* the reference returns whichever buffer is "web" after the swaps
     * (stereo.c:246-251); with an even/odd count the data are equal anyway
     * because nothing is ever written. */
    if (a != web)
        memcpy(web, a, sizeof(int32_t) * n);
    free(tmp);
}

/* draw_contour_map (stereo.c:256-274): out = ((web-min) % interval) == 0,
 * interval = (max-min)/num_lines.  Returns 1 instead of dividing by zero when
 * interval == 0 (the reference would trap). */
int oracle_draw_contour_map(const int32_t *web, int w, int h, int num_lines, uint8_t *out)
{
    size_t n = (size_t)w * h;
    int32_t mx = INT32_MIN, mn = INT32_MAX;
    for (size_t p = 0; p < n; p++) {
        if (web[p] > mx) mx = web[p];
        if (web[p] < mn) mn = web[p];
    }
    if (num_lines == 0) return 1;
    int32_t interval = (mx - mn) / num_lines;
    if (interval == 0) return 1;
    for (size_t p = 0; p < n; p++)
        out[p] = ((web[p] - mn) % interval) == 0;
    return 0;
}

/* ------------------------------------------------------------------ */
/* synthetic pair generator                              */
/* ------------------------------------------------------------------ */

static uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

#define SYNTH_K 0xD6E8FEB86659FD93ull

static uint8_t synth_left(uint64_t seed, int x, int y)
{
    uint64_t hk = splitmix64(seed * SYNTH_K + ((uint64_t)y << 20) + (uint64_t)x);
    return ((hk >> 8) & 7) == 0 ? (uint8_t)(hk & 0xFF) : 128;
}

/* left: flat gray with 1-in-8 random speckle; disparity constant on
 * TW x TH tiles; right(x,y) = left((x - d) mod W, y).  disp may be NULL. */
void oracle_synth_pair(uint64_t seed, int w, int h, int num_shifts, uint8_t *left, uint8_t *right,
                       int32_t *disp)
{
    int tw = 4 * num_shifts > 240 ? 4 * num_shifts : 240, th = 120;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint64_t hd =
                splitmix64((seed + 1) * SYNTH_K + (uint64_t)(y / th) * 4096 + (uint64_t)(x / tw));
            int d = (int)(hd % (uint64_t)num_shifts);
            int xs = ((x - d) % w + w) % w;
            left[(size_t)y * w + x] = synth_left(seed, x, y);
            right[(size_t)y * w + x] = synth_left(seed, xs, y);
            if (disp) disp[(size_t)y * w + x] = d;
        }
}

/* ------------------------------------------------------------------ */
/* CRC32 (zlib polynomial) so C harnesses can print the golden pins     */
/* ------------------------------------------------------------------ */
uint32_t oracle_crc32(const void *data, size_t n)
{
    static uint32_t table[256];
    static int init = 0;
    if (!init) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = c & 1 ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = 1;
    }
    uint32_t c = 0xFFFFFFFFu;
    const uint8_t *p = (const uint8_t *)data;
    for (size_t i = 0; i < n; i++) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}
