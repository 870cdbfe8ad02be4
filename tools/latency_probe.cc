// tools/latency_probe.cc -- single-frame launches through the C ABI (no Python in the loop):
// BASELINE config 2 (1920x1080 NV12, 3 regions) and config 1 (1280x720 I420, one cue), one frame
// per launch, device-resident, out of place. Host us per submit (issue loop), device us per
// frame (device timer around the loop), and the latency of submit + wait.
#include "../include/fluc_ttmlblend.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define OK(x) do { int rc_ = (x); if (rc_) { fprintf (stderr, "%s: %s\n", #x, fluc_ttmlblend_strerror (rc_)); return 1; } } while (0)
static double now () { return std::chrono::duration<double> (std::chrono::steady_clock::now ().time_since_epoch ()).count (); }

static int
run (FlucTtmlBlend *ctx, const char *name, FlucTtmlBlendFormat fmt, int W, int H, const FlucTtmlBlendRect *regions, int n_regions,
    uint32_t stream)
{
  std::vector<uint8_t> img ((size_t) W * H * 4, 0);
  for (int r = 0; r < n_regions; r++)
    for (int y = regions[r].y; y < regions[r].y + regions[r].h; y++)
      for (int x = regions[r].x; x < regions[r].x + regions[r].w; x++) {
        uint8_t *p = &img[((size_t) y * W + x) * 4];
        const bool text = ((x / 24) & 1) && ((y / 12) % 3 == 1);
        p[0] = p[1] = p[2] = text ? 255 : 0;
        p[3] = text ? 255 : 160;
      }
  OK (fluc_ttmlblend_overlay_set (ctx, stream, img.data (), W, H, W * 4, regions, (uint32_t) n_regions));
  OK (fluc_ttmlblend_set_batch (ctx, 1, 0));
  const int n_buf = 8;
  FlucTtmlBlendFrame src[n_buf], dst[n_buf];
  for (int i = 0; i < n_buf; i++) {
    OK (fluc_ttmlblend_frame_pool_acquire (ctx, fmt, W, H, 0, &src[i]));
    OK (fluc_ttmlblend_frame_pool_acquire (ctx, fmt, W, H, 0, &dst[i]));
  }
  uint64_t t = 0;
  for (int i = 0; i < 200; i++)
    OK (fluc_ttmlblend_submit (ctx, stream, fmt, W, H, 0, &src[i % n_buf], &dst[i % n_buf], &t));
  OK (fluc_ttmlblend_sync (ctx));
  const int n = 20000;
  OK (fluc_ttmlblend_timer_begin (ctx));
  const double t0 = now ();
  for (int i = 0; i < n; i++)
    OK (fluc_ttmlblend_submit (ctx, stream, fmt, W, H, 0, &src[i % n_buf], &dst[i % n_buf], &t));
  const double issue = now () - t0;
  double ms = 0;
  OK (fluc_ttmlblend_timer_end (ctx, &ms));
  OK (fluc_ttmlblend_sync (ctx));
  const int m = 5000;
  const double t1 = now ();
  for (int i = 0; i < m; i++) {
    OK (fluc_ttmlblend_submit (ctx, stream, fmt, W, H, 0, &src[i % n_buf], &dst[i % n_buf], &t));
    OK (fluc_ttmlblend_wait (ctx, t));
  }
  const double lat = (now () - t1) / m;
  printf ("%-28s host %6.2f us/submit   device %6.2f us/frame (%8.0f frames/s)   submit+wait %6.2f us\n", name,
      issue / n * 1e6, ms / n * 1e3, n / (ms * 1e-3), lat * 1e6);
  for (int i = 0; i < n_buf; i++) {
    fluc_ttmlblend_frame_pool_release (ctx, &src[i]);
    fluc_ttmlblend_frame_pool_release (ctx, &dst[i]);
  }
  return 0;
}

int
main ()
{
  FlucTtmlBlend *ctx = nullptr;
  OK (fluc_ttmlblend_new (0, &ctx));
  const FlucTtmlBlendRect c2[3] = { { 192, 54, 1536, 108 }, { 96, 486, 672, 162 }, { 192, 864, 1536, 162 } };
  const FlucTtmlBlendRect c1[1] = { { 128, 576, 1024, 108 } };
  int rc = run (ctx, "cfg 2: 1080p NV12, 3 regions", FLUC_TTMLBLEND_FORMAT_NV12, 1920, 1080, c2, 3, 2);
  rc |= run (ctx, "cfg 1: 720p I420, 1 cue", FLUC_TTMLBLEND_FORMAT_I420, 1280, 720, c1, 1, 1);
  fluc_ttmlblend_free (ctx);
  return rc;
}
