/* pwn_image.h -- frame dumps for visual diffing (the display tail without SDL). */
#ifndef PWN_IMAGE_H
#define PWN_IMAGE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/*
 * Write the top-left w x h of a frame (stride rwidth words, 0xAARRGGBB little-endian as the
 * reference hands it to SDL, main.c:388 / util.h:48-59) as a binary PPM (P6).  0 = ok.
 */
int pwn_save_ppm(const char *path, const uint32_t *sbuf, int w, int h, int rwidth);

/* |a - b| per colour byte (alpha ignored) amplified by `gain` and saturated, as a frame: a visual diff */
int pwn_frame_diff(const uint32_t *a, const uint32_t *b, uint32_t *out, int w, int h, int rwidth, int gain);

#ifdef __cplusplus
}
#endif
#endif
