/*
 * cue_storm.c -- does a stream that changes cues all the time hold up the frames of the
 * others? Plain C over the C ABI (a Python probe measures the GIL instead).
 *
 * Main thread: batches of 32 device-resident 4K NV12 frames through submit_many for 2 s.
 * Storm threads: overlay_set of a whole 4K image (33 MB from pageable memory) on other
 * streams, 1 ms apart. overlay_set uploads with the context unlocked.
 *
 *   gcc -O2 -std=c99 -D_POSIX_C_SOURCE=200809L -Iinclude tools/cue_storm.c \
 *       -Lflu-plugins-oss_b200/csrc -lfluc_ttmlblend -Wl,-rpath,$PWD/flu-plugins-oss_b200/csrc -lpthread -o build/cue_storm
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "fluc_ttmlblend.h"

enum { W = 3840, H = 2160, N = 32 };

static FlucTtmlBlend *ctx;
static uint8_t *image;
static volatile int stop;
static long changes[16];

static double
now (void)
{
  struct timespec t;
  clock_gettime (CLOCK_MONOTONIC, &t);
  return t.tv_sec + 1e-9 * t.tv_nsec;
}

static void *
storm (void *arg)
{
  const long i = (long) arg;
  while (!stop) {
    if (fluc_ttmlblend_overlay_set (ctx, 100 + (uint32_t) i, image, W, H, W * 4, NULL, 0) != 0)
      break;
    changes[i]++;
    {
      /* a pause between cue changes: without one an unfair mutex lets a library that uploads
       * under its lock starve the frame thread for good */
      const struct timespec ms = { 0, 1000000 };
      nanosleep (&ms, NULL);
    }
  }
  return NULL;
}

int
main (void)
{
  FlucTtmlBlendFrame src[N], dst[N];
  uint32_t streams[N];
  const FlucTtmlBlendRect boxes[2] = { { 0, 1728, W, 360 }, { 0, 72, W, 144 } };
  int k, y, x, storms;
  if (fluc_ttmlblend_new (0, &ctx) != 0) {
    fprintf (stderr, "no CUDA device\n");
    return 1;
  }
  /* a cue image like config 3: two filled boxes, premultiplied, the rest transparent */
  image = (uint8_t *) calloc ((size_t) W * H, 4);
  for (k = 0; k < 2; k++)
    for (y = boxes[k].y; y < boxes[k].y + boxes[k].h; y++)
      for (x = 0; x < W; x++) {
        uint8_t *p = image + ((size_t) y * W + x) * 4;
        p[3] = 191;
        p[0] = p[1] = p[2] = (uint8_t) ((x ^ y) & 127);
      }
  fluc_ttmlblend_overlay_set (ctx, 1, image, W, H, W * 4, boxes, 2);
  fluc_ttmlblend_set_batch (ctx, N, 0);
  for (k = 0; k < N; k++) {
    fluc_ttmlblend_frame_pool_acquire (ctx, FLUC_TTMLBLEND_FORMAT_NV12, W, H, 0, &src[k]);
    fluc_ttmlblend_frame_pool_acquire (ctx, FLUC_TTMLBLEND_FORMAT_NV12, W, H, 0, &dst[k]);
    streams[k] = 1;
  }
  for (storms = 0; storms <= 4; storms = storms ? storms * 2 : 1) {
    pthread_t th[16];
    double t0, t1;
    long frames = 0, total = 0;
    long i;
    stop = 0;
    memset (changes, 0, sizeof changes);
    for (i = 0; i < storms; i++)
      pthread_create (&th[i], NULL, storm, (void *) i);
    fluc_ttmlblend_sync (ctx);
    t0 = now ();
    do {
      for (k = 0; k < 20; k++)
        fluc_ttmlblend_submit_many (ctx, N, streams, FLUC_TTMLBLEND_FORMAT_NV12, W, H, 0, src, dst, NULL);
      fluc_ttmlblend_sync (ctx);
      frames += 20 * N;
      t1 = now ();
    } while (t1 - t0 < 2.0);
    stop = 1;
    for (i = 0; i < storms; i++) {
      pthread_join (th[i], NULL);
      total += changes[i];
    }
    printf ("%d storm thread(s): %.0f frames/s, %.0f cue changes/s (whole 4K image each)\n", storms,
        frames / (t1 - t0), total / (t1 - t0));
    fflush (stdout);
  }
  fluc_ttmlblend_free (ctx);
  free (image);
  return 0;
}
