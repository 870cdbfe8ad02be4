/*
 * ttmlblend_demo.c -- the C ABI driven from plain C, the way an element does
 * (INTEGRATION.md section 2): one cue image, a batch of device-resident NV12
 * frames, then the host-frame drop-in call, with frames/s for both.
 *
 *   gcc -O2 -Iinclude examples/ttmlblend_demo.c -Lflu-plugins-oss_b200/csrc \
 *       -lfluc_ttmlblend -Wl,-rpath,$PWD/flu-plugins-oss_b200/csrc -o build/ttmlblend_demo
 *   ./build/ttmlblend_demo [width height frames_per_batch batches]
 *
 * Only include/fluc_ttmlblend.h is used: no CUDA, GLib or Python.
 */
#define _POSIX_C_SOURCE 200809L
#include "fluc_ttmlblend.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define CHECK(call) do { int rc_ = (call); if (rc_ != FLUC_TTMLBLEND_OK) {            \
    fprintf (stderr, "%s: %s (%s)\n", #call, fluc_ttmlblend_strerror (rc_),          \
        ctx ? fluc_ttmlblend_last_cuda_error (ctx) : ""); return 1; } } while (0)

static double
now (void)
{
  struct timespec t;
  clock_gettime (CLOCK_MONOTONIC, &t);
  return t.tv_sec + 1e-9 * t.tv_nsec;
}

int
main (int argc, char **argv)
{
  const int W = argc > 1 ? atoi (argv[1]) : 1920, H = argc > 2 ? atoi (argv[2]) : 1080;
  const int n = argc > 3 ? atoi (argv[3]) : 32, batches = argc > 4 ? atoi (argv[4]) : 200;
  const FlucTtmlBlendFormat fmt = FLUC_TTMLBLEND_FORMAT_NV12;
  FlucTtmlBlend *ctx = NULL;
  FlucTtmlBlendFrame *src, *dst, *host;
  uint32_t *streams;
  uint64_t *tickets;
  uint8_t *cue;
  FlucTtmlBlendRect region;
  FlucTtmlBlendStats st;
  double t0, dt;
  int i, b, x, y;

  if (n < 1 || n > 1024 || W < 16 || H < 16)
    return 2;
  CHECK (fluc_ttmlblend_new (-1, &ctx));
  printf ("%s, %d device(s)\n", fluc_ttmlblend_version (), fluc_ttmlblend_device_count ());

  /* what ttmlrender hands over: W*H premultiplied BGRA, cleared, one region box with a
   * translucent background and a few opaque "glyph" bars */
  cue = (uint8_t *) calloc ((size_t) W * H, 4);
  region.x = W / 10; region.y = H * 8 / 10; region.w = W * 8 / 10; region.h = H * 15 / 100;
  for (y = region.y; y < region.y + region.h; y++)
    for (x = region.x; x < region.x + region.w; x++) {
      uint8_t *p = cue + ((size_t) y * W + x) * 4;
      const int glyph = ((x / 12) % 3 == 0) && ((y - region.y) % 40 > 8) && ((y - region.y) % 40 < 32);
      p[3] = glyph ? 255 : 160;                         /* alpha */
      p[0] = p[1] = p[2] = glyph ? 255 : 0;             /* premultiplied colour <= alpha */
    }
  CHECK (fluc_ttmlblend_overlay_set (ctx, 1, cue, W, H, W * 4, &region, 1));

  src = (FlucTtmlBlendFrame *) calloc (n, sizeof *src);
  dst = (FlucTtmlBlendFrame *) calloc (n, sizeof *dst);
  host = (FlucTtmlBlendFrame *) calloc (n, sizeof *host);
  streams = (uint32_t *) calloc (n, sizeof *streams);
  tickets = (uint64_t *) calloc (n, sizeof *tickets);
  for (i = 0; i < n; i++) {
    int pl;
    streams[i] = 1;
    CHECK (fluc_ttmlblend_frame_pool_acquire (ctx, fmt, W, H, 0, &src[i]));
    CHECK (fluc_ttmlblend_frame_pool_acquire (ctx, fmt, W, H, 0, &dst[i]));
    CHECK (fluc_ttmlblend_frame_pool_acquire (ctx, fmt, W, H, 1, &host[i]));
    for (pl = 0; pl < fluc_ttmlblend_format_planes (fmt); pl++)
      memset (host[i].plane[pl], 64 + 16 * pl + i, (size_t) host[i].stride[pl] *
          fluc_ttmlblend_plane_rows (fmt, pl, H));
    CHECK (fluc_ttmlblend_frame_upload (ctx, fmt, W, H, &host[i], &src[i]));
  }
  CHECK (fluc_ttmlblend_set_batch (ctx, (uint32_t) n, 0));

  /* device-resident frames: one submit_many per batch, the batch limit launches it */
  for (b = 0; b < 5; b++)
    CHECK (fluc_ttmlblend_submit_many (ctx, n, streams, fmt, W, H, 0, src, dst, tickets));
  CHECK (fluc_ttmlblend_sync (ctx));
  t0 = now ();
  for (b = 0; b < batches; b++)
    CHECK (fluc_ttmlblend_submit_many (ctx, n, streams, fmt, W, H, 0, src, dst, tickets));
  CHECK (fluc_ttmlblend_sync (ctx));
  dt = now () - t0;
  printf ("device frames : %dx%d NV12, %d per launch: %.0f frames/s (%.1f us per launch)\n", W, H, n,
      (double) n * batches / dt, dt / batches * 1e6);

  /* host frames, in place: gst_video_overlay_composition_blend's job */
  for (b = 0; b < 3; b++) {
    for (i = 0; i < n; i++)
      CHECK (fluc_ttmlblend_blend_host (ctx, 1, fmt, W, H, 0, &host[i], &tickets[i]));
    CHECK (fluc_ttmlblend_wait (ctx, tickets[n - 1]));
  }
  t0 = now ();
  for (b = 0; b < batches / 4 + 1; b++) {
    for (i = 0; i < n; i++)
      CHECK (fluc_ttmlblend_blend_host (ctx, 1, fmt, W, H, 0, &host[i], &tickets[i]));
    CHECK (fluc_ttmlblend_wait (ctx, tickets[n - 1]));
  }
  dt = now () - t0;
  printf ("host frames   : %.0f frames/s through fluc_ttmlblend_blend_host (pinned, zero copy)\n",
      (double) n * (batches / 4 + 1) / dt);

  /* single frame latency: submit + wait */
  CHECK (fluc_ttmlblend_set_batch (ctx, 1, 0));
  t0 = now ();
  for (b = 0; b < 2000; b++) {
    CHECK (fluc_ttmlblend_submit (ctx, 1, fmt, W, H, 0, &src[0], &dst[0], &tickets[0]));
    CHECK (fluc_ttmlblend_wait (ctx, tickets[0]));
  }
  dt = now () - t0;
  printf ("one frame     : %.1f us submit -> wait\n", dt / 2000 * 1e6);

  fluc_ttmlblend_stats_copy (ctx, &st);
  printf ("stats: %llu frames, %llu launches (%llu group), %.1f GB algorithmic\n",
      (unsigned long long) st.frames_blended, (unsigned long long) st.launches,
      (unsigned long long) st.group_launches, st.algorithmic_bytes / 1e9);
  fluc_ttmlblend_free (ctx);
  free (cue);
  return 0;
}
