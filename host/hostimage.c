/* hostimage.c -- see hostimage.h.  Plain C, links zlib for the PNG inflate. */
#include "hostimage.h"

#include <errno.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

/* ------------------------------------------------------------------------- */
/* PNG reader (non-interlaced and Adam7 are both rare here; interlace is rejected) */
/* ------------------------------------------------------------------------- */

static uint32_t be32(const uint8_t *p)
{
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}

static int paeth(int a, int b, int c)
{
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    if (pa <= pb && pa <= pc) return a;
    return pb <= pc ? b : c;
}

static int fail_read(const char *name, const char *why)
{
    /* the reference prints "error reading image %s:" followed by perror("") (image.c:22-25) */
    fprintf(stderr, "error reading image %s:", name);
    if (why)
        fprintf(stderr, " %s\n", why);
    else
        perror("");
    return 1;
}

int read_image(const char *name, Image8 *out)
{
    FILE *f = fopen(name, "rb");
    if (!f) return fail_read(name, NULL);
    fseek(f, 0, SEEK_END);
    long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *buf = (uint8_t *)malloc(size > 0 ? (size_t)size : 1);
    if (!buf || fread(buf, 1, (size_t)size, f) != (size_t)size) {
        fclose(f);
        free(buf);
        return fail_read(name, "short read");
    }
    fclose(f);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (size < 8 || memcmp(buf, sig, 8) != 0) {
        free(buf);
        return fail_read(name, "not a PNG file");
    }
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0, have_trns = 0, have_ihdr = 0;
    uint8_t *idat = (uint8_t *)malloc((size_t)size);
    size_t idat_len = 0;
    for (long pos = 8; pos + 12 <= size;) {
        uint32_t len = be32(buf + pos);
        const uint8_t *type = buf + pos + 4, *data = buf + pos + 8;
        if ((long)len > size - pos - 12) break;
        if (!memcmp(type, "IHDR", 4) && len >= 13) {
            w = be32(data);
            h = be32(data + 4);
            depth = data[8];
            ctype = data[9];
            interlace = data[12];
            have_ihdr = 1;
        } else if (!memcmp(type, "tRNS", 4)) {
            have_trns = 1;
        } else if (!memcmp(type, "IDAT", 4)) {
            memcpy(idat + idat_len, data, len);
            idat_len += len;
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (long)len;
    }
    free(buf);
    if (!have_ihdr || w == 0 || h == 0 || w > (1u << 24) || h > (1u << 24)) {
        free(idat);
        return fail_read(name, "bad PNG header");
    }
    /* channel count as stb_image would report it with req_comp = 0 (image.c:21) */
    int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 3 : ctype == 4 ? 2 : 4;
    if (have_trns && (ctype == 0 || ctype == 2 || ctype == 3)) channels += 1;
    if (channels != 1) {
        free(idat);
        /* verbatim reference message, no newline (image.c:27-31) */
        fprintf(stderr,
                "error reading image %s: wrong number of channels (%d) "
                "(image must be grayscale)",
                name, channels);
        return 1;
    }
    if (interlace != 0 || !(depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) {
        free(idat);
        return fail_read(name, "unsupported PNG (interlaced or odd bit depth)");
    }
    size_t stride = ((size_t)w * depth + 7) / 8, bpp = depth == 16 ? 2 : 1;
    uLongf raw_len = (uLongf)((stride + 1) * h);
    uint8_t *raw = (uint8_t *)malloc(raw_len);
    if (!raw || uncompress(raw, &raw_len, idat, (uLong)idat_len) != Z_OK || raw_len != (stride + 1) * h) {
        free(idat);
        free(raw);
        return fail_read(name, "corrupt PNG data");
    }
    free(idat);
    /* undo the per-row filters in place */
    for (uint32_t y = 0; y < h; y++) {
        uint8_t *row = raw + (stride + 1) * y + 1;
        const uint8_t *prev = y ? row - (stride + 1) : NULL;
        int ft = row[-1];
        for (size_t i = 0; i < stride; i++) {
            int a = i >= bpp ? row[i - bpp] : 0, b = prev ? prev[i] : 0, c = (prev && i >= bpp) ? prev[i - bpp] : 0;
            int x = row[i];
            switch (ft) {
            case 0: break;
            case 1: x += a; break;
            case 2: x += b; break;
            case 3: x += (a + b) >> 1; break;
            case 4: x += paeth(a, b, c); break;
            default:
                free(raw);
                return fail_read(name, "corrupt PNG filter");
            }
            row[i] = (uint8_t)x;
        }
    }
    uint8_t *pix = (uint8_t *)malloc((size_t)w * h);
    if (!pix) {
        free(raw);
        fprintf(stderr, "error: out of memory\n");
        exit(1);
    }
    /* to 8 bits the way stb_image does: 16 -> high byte, 1/2/4 -> scaled to 0..255 */
    static const int scale[9] = {0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01};
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t *row = raw + (stride + 1) * y + 1;
        uint8_t *o = pix + (size_t)y * w;
        if (depth == 8) {
            memcpy(o, row, w);
        } else if (depth == 16) {
            for (uint32_t x = 0; x < w; x++) o[x] = row[2 * x];
        } else {
            for (uint32_t x = 0; x < w; x++) {
                int per = 8 / depth, shift = (per - 1 - (int)(x % per)) * depth;
                o[x] = (uint8_t)(((row[x / per] >> shift) & ((1 << depth) - 1)) * scale[depth]);
            }
        }
    }
    free(raw);
    out->data = pix;
    out->width = (int)w;
    out->height = (int)h;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* PPM writer                                                                 */
/* ------------------------------------------------------------------------- */

char *make_filename(const char *name, ImageProgramType type, int number)
{
    char *filename = (char *)calloc(1024, 1);
    if (!filename) {
        fprintf(stderr, "error: out of memory\n");
        exit(1);
    }
#ifdef DEBUG
    /* in debug builds every program writes into its own directory so that test/diff.sh can
     * compare them file by file (image.c:55-63) */
    static const char *dirs[] = {"ser", "par", "sergh", "pargh"};
    snprintf(filename, 1024, "%s/%s-%d.ppm", dirs[type], name, number);
#else
    (void)type;
    snprintf(filename, 1024, "%s-%d.ppm", name, number);
#endif
    return filename;
}

#ifndef NO_WRITES
/* map() of the reference (image.c:37-40): truncating long arithmetic.  The reference
 * divides by zero for a constant image; that case writes 0 here instead of trapping. */
static long map_range(long x, long in_min, long in_max, long out_min, long out_max)
{
    if (in_max == in_min) return out_min;
    return (x - in_min) * (out_max - out_min) / (in_max - in_min) + out_min;
}
#endif

void write_image(const void *data, int width, int height, ImageType type, char *filename)
{
#ifndef NO_WRITES
    FILE *f = fopen(filename, "w");
    free(filename);
    if (!f) return;
    const size_t n = (size_t)width * height;
    int mn = 0, mx = 0;
    if (type == IMTYPE_GRAY_INT) {
        const int32_t *p = (const int32_t *)data;
        mn = INT_MAX;
        mx = INT_MIN;
        for (size_t i = 0; i < n; i++) {
            if (p[i] < mn) mn = p[i];
            if (p[i] > mx) mx = p[i];
        }
    }
    /* the file is large (3 numbers per pixel as text): format into a buffer per row */
    fprintf(f, "P3\n%d %d\n255\n", width, height);
    char *line = (char *)malloc((size_t)width * 16 + 16);
    for (int y = 0; y < height; y++) {
        char *o = line;
        for (int x = 0; x < width; x++) {
            size_t i = (size_t)y * width + x;
            int v = type == IMTYPE_BINARY ? (((const uint8_t *)data)[i] == 1 ? 0 : 255)
                                          : (int)map_range(((const int32_t *)data)[i], mn, mx, 0, 255);
            o += sprintf(o, "%d %d %d\n", v, v, v);
        }
        fwrite(line, 1, (size_t)(o - line), f);
    }
    free(line);
    fclose(f);
#else
    (void)data, (void)width, (void)height, (void)type;
    free(filename);
#endif
}
