/*
 * hostimage.h -- host-side image I/O of the drivers: the contract of the reference's
 * src/image.h / src/image.c (read_image, make_filename, write_image), re-implemented.
 *
 * The reference decodes PNGs with the vendored stb_image.h and converts to double
 * (image.c:9-35); that file is third-party code and is not copied here.  read_image()
 * below is a small PNG reader on top of zlib for the files the reference accepts
 * (1-channel / grayscale; anything else is rejected with the reference's message,
 * image.c:27-31) and keeps the 8-bit pixels: sm_upload_u8() is exactly equivalent to
 * uploading u8/256.0 doubles.
 *
 * write_image() reproduces the reference's ASCII P3 writer byte for byte
 * (image.c:37-47,71-88): "P3\n%d %d\n255\n" then "%d %d %d\n" per pixel, binary
 * images as 1 -> 0 / else 255, integer images min/max-normalised with the same
 * truncating long arithmetic.  test/diff.sh compares these files.
 */
#ifndef HOSTIMAGE_H_INCLUDED
#define HOSTIMAGE_H_INCLUDED

#include <stdint.h>

typedef struct {
    uint8_t *data; /* width*height 8-bit pixels, row-major; caller frees */
    int width, height;
} Image8;

typedef enum ImageType {
    IMTYPE_BINARY,   /* u8, 1 = black, anything else white   (image.h:15) */
    IMTYPE_GRAY_INT  /* i32, normalised to 0..255 on writing (image.h:17) */
} ImageType;

/* which program writes: selects the -DDEBUG output directory (image.h:20-22, image.c:57-63) */
typedef enum ImageProgramType { SER = 0, PAR, SERGHOST, PARGHOST } ImageProgramType;

/* 0 on success; 1 after printing the reference's message to stderr (image.c:18-35). */
int read_image(const char *name, Image8 *out);

/* malloc'd "name-number.ppm" (or "<dir>/name-number.ppm" under -DDEBUG); write_image frees it. */
char *make_filename(const char *name, ImageProgramType type, int number);

/* Writes width*height pixels of `data` (u8 for IMTYPE_BINARY, int32_t for IMTYPE_GRAY_INT)
 * as ASCII PPM and frees `filename`.  Compiles to a no-op under -DNO_WRITES (image.c:73). */
void write_image(const void *data, int width, int height, ImageType type, char *filename);

#endif
