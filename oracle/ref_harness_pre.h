/*
 * ref_harness_pre.h -- TEST INFRASTRUCTURE (oracle/_ref build), not product code.
 *
 * Prelude that is concatenated IN FRONT of the reference's own src/qpsk.c (streamed
 * from /root/reference through sed, never copied into this repository) so that one
 * translation unit can (a) rename the reference's main(), (b) capture the DEBUG2
 * printf at /root/reference/src/qpsk.c:198 instead of printing it, and (c) reach and
 * reset the file-scope statics at /root/reference/src/qpsk.c:34-60 from the harness
 * functions that are concatenated BEHIND it (ref_harness_post.c).
 */
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <stdint.h>

static int ref_capture_printf(const char *fmt, ...);

#define printf ref_capture_printf
#define main ref_main
