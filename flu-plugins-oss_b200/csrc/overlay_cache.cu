/*
 * overlay_cache.cu -- the per-stream overlay cache: upload of the cue rectangles, auto-crop to
 * non-transparent row runs, the once-per-cue prepare launches, and stream-ordered deferred frees.
 * Producer side in the reference: gst_ttmlrender_gen_buffer,
 * /root/reference/plugins/ttml/gstttmlrender.c:1427-1478.
 */
#include "ttmlblend_internal.h"

namespace tbh {

/* Frees device memory once everything already queued on any of the
 * context's streams has run: the reaper stream waits for a fence event on
 * each of them, then frees in stream order. */
void
free_deferred (Ctx *c, const std::vector<void *> &ptrs)
{
  if (ptrs.empty ())
    return;
  cudaStream_t all[kLanes + 2];
  int n = 0;
  all[n++] = c->blend_stream;
  all[n++] = c->up_stream;
  for (int i = 0; i < kLanes; i++)
    all[n++] = c->lanes[i].stream;
  for (int i = 0; i < n; i++) {
    if (!all[i])
      continue;
    cudaEventRecord (c->ev_fence[i], all[i]);
    cudaStreamWaitEvent (c->reaper, c->ev_fence[i], 0);
  }
  for (void *p : ptrs)
    cudaFreeAsync (p, c->reaper);
}

Overlay::~Overlay ()
{
  std::vector<void *> ptrs;
  for (void *a : raw_allocs)
    ptrs.push_back (a);
  for (auto &p : prepared) {
    for (void *a : p->allocs)
      ptrs.push_back (a);
    if (p->ready)
      cudaEventDestroy (p->ready);
  }
  if (ctx)
    free_deferred (ctx, ptrs);
}

/* ---------------------------------------------------------------------- */
/* prepare: raw BGRA rectangle -> per-plane prepared overlay              */

int
dev_alloc (Ctx *c, Prepared *p, size_t bytes, uint8_t **out)
{
  void *ptr = nullptr;
  CU (c, cudaMallocAsync (&ptr, std::max<size_t> (bytes, 16), c->up_stream));
  p->allocs.push_back (ptr);
  *out = static_cast<uint8_t *> (ptr);
  return 0;
}

int
prepare_overlay (Ctx *c, Overlay *ov, int format, int W, int H, Prepared **out)
{
  for (auto &p : ov->prepared)
    if (p->format == format && p->W == W && p->H == H && p->chroma_average == c->chroma_average) {
      *out = p.get ();
      return 0;
    }

  NvtxRange nvtx ("ttmlblend.prepare_overlay");
  TBLOG (2, "prepare overlay: format %d, %dx%d, %zu rectangle(s)", format, W, H, ov->rects.size ());
  std::unique_ptr<Prepared> P (new Prepared ());
  P->format = format;
  P->W = W;
  P->H = H;
  P->chroma_average = c->chroma_average;
  const int kind = plane_kind (format);
  const int n_planes = format_planes (format);

  for (const FlucTtmlBlendRect &d : ov->declared) {
    const int w = std::min (d.x + d.w, W) - d.x, h = std::min (d.y + d.h, H) - d.y;
    if (w > 0 && h > 0)
      P->overlay_px += (uint64_t) w * (uint64_t) h;
  }
  for (const RawRect &rr : ov->rects) {
    /* gst_video_blend clipping: rr is already clipped at the left/top */
    const int cx0 = rr.x, cy0 = rr.y;
    const int cx1 = std::min (rr.x + rr.w, W), cy1 = std::min (rr.y + rr.h, H);
    if (cx1 <= cx0 || cy1 <= cy0)
      continue;
    if (rr.ga == 0)
      continue;                 /* asrc == 0 everywhere: blends nothing */

    PrepareParams pp = {};
    pp.raw = rr.dev;
    pp.raw_pitch = rr.pitch;
    pp.raw_w = rr.w;
    pp.raw_h = rr.h;
    pp.fx = rr.x;
    pp.fy = rr.y;
    pp.cx0 = cx0; pp.cy0 = cy0; pp.cx1 = cx1; pp.cy1 = cy1;
    pp.ga = rr.ga;
    pp.premul = rr.premul ? 1 : 0;

    if (kind == PK_PLANE8) {
      /* luma plane: byte == pixel */
      {
        RectRef ref = {};
        ref.v0 = cx0 / 16;
        ref.v1 = ceil_div (cx1, 16);
        ref.y0 = cy0;
        ref.y1 = cy1;
        ref.pitch = (ref.v1 - ref.v0) * 16;
        ref.ga = 255;
        const size_t bytes = (size_t) ref.pitch * (cy1 - cy0);
        uint8_t *a, *y;
        int rc;
        if ((rc = dev_alloc (c, P.get (), bytes, &a)) || (rc = dev_alloc (c, P.get (), bytes, &y)))
          return rc;
        ref.a = a;
        ref.c = y;
        pp.mode = PM_LUMA;
        pp.out_a = a; pp.out_c = y; pp.out_c2 = nullptr;
        pp.out_pitch = ref.pitch;
        pp.v0 = ref.v0;
        pp.row0 = cy0;
        pp.rows = cy1 - cy0;
        CU (c, launch_prepare (pp, ref.pitch, c->up_stream));
        c->stats.prepare_launches++;
        P->h_rects[0].push_back (ref);
      }
      /* chroma: the samples sited on even x / even y (parity); with the non-parity 2x2
       * average every sample with at least one covered pixel */
      pp.chroma_average = P->chroma_average ? 1 : 0;
      const int bx0 = P->chroma_average ? cx0 / 2 : ceil_div (cx0, 2), bx1 = ceil_div (cx1, 2);
      const int by0 = P->chroma_average ? cy0 / 2 : ceil_div (cy0, 2), by1 = ceil_div (cy1, 2);
      if (bx1 > bx0 && by1 > by0) {
        if (n_planes == 3) {
          RectRef ref = {};
          ref.v0 = bx0 / 16;
          ref.v1 = ceil_div (bx1, 16);
          ref.y0 = by0;
          ref.y1 = by1;
          ref.pitch = (ref.v1 - ref.v0) * 16;
          ref.ga = 255;
          const size_t bytes = (size_t) ref.pitch * (by1 - by0);
          uint8_t *a, *u, *v;
          int rc;
          if ((rc = dev_alloc (c, P.get (), bytes, &a)) || (rc = dev_alloc (c, P.get (), bytes, &u))
              || (rc = dev_alloc (c, P.get (), bytes, &v)))
            return rc;
          pp.mode = PM_CHROMA_PLANAR;
          pp.out_a = a; pp.out_c = u; pp.out_c2 = v;
          pp.out_pitch = ref.pitch;
          pp.v0 = ref.v0;
          pp.row0 = by0;
          pp.rows = by1 - by0;
          CU (c, launch_prepare (pp, ref.pitch, c->up_stream));
          c->stats.prepare_launches++;
          const int pu = format == FLUC_TTMLBLEND_FORMAT_I420 ? 1 : 2;
          const int pv = 3 - pu;
          ref.a = a;
          ref.c = u;
          P->h_rects[pu].push_back (ref);
          ref.c = v;
          P->h_rects[pv].push_back (ref);
        } else {
          RectRef ref = {};
          ref.v0 = (2 * bx0) / 16;
          ref.v1 = ceil_div (2 * bx1, 16);
          ref.y0 = by0;
          ref.y1 = by1;
          ref.pitch = (ref.v1 - ref.v0) * 16;
          ref.ga = 255;
          const size_t bytes = (size_t) ref.pitch * (by1 - by0);
          uint8_t *a, *uv;
          int rc;
          if ((rc = dev_alloc (c, P.get (), bytes, &a)) || (rc = dev_alloc (c, P.get (), bytes, &uv)))
            return rc;
          pp.mode = format == FLUC_TTMLBLEND_FORMAT_NV12 ? PM_CHROMA_UV : PM_CHROMA_VU;
          pp.out_a = a; pp.out_c = uv; pp.out_c2 = nullptr;
          pp.out_pitch = ref.pitch;
          pp.v0 = ref.v0;
          pp.row0 = by0;
          pp.rows = by1 - by0;
          CU (c, launch_prepare (pp, ref.pitch / 2, c->up_stream));
          c->stats.prepare_launches++;
          ref.a = a;
          ref.c = uv;
          P->h_rects[1].push_back (ref);
        }
      }
    } else {
      RectRef ref = {};
      ref.v0 = cx0 / 4;
      ref.v1 = ceil_div (cx1, 4);
      ref.y0 = cy0;
      ref.y1 = cy1;
      ref.pitch = (ref.v1 - ref.v0) * 16;
      ref.ga = rr.ga;
      const bool yuv = format == FLUC_TTMLBLEND_FORMAT_AYUV;
      ref.src_premul = (!yuv && rr.premul) ? 1 : 0;
      const size_t bytes = (size_t) ref.pitch * (cy1 - cy0);
      uint8_t *w;
      int rc;
      if ((rc = dev_alloc (c, P.get (), bytes, &w)))
        return rc;
      switch (format) {
        case FLUC_TTMLBLEND_FORMAT_AYUV: pp.mode = PM_PACKED_AYUV; break;
        case FLUC_TTMLBLEND_FORMAT_ARGB: pp.mode = PM_PACKED_ARGB; break;
        case FLUC_TTMLBLEND_FORMAT_ABGR: pp.mode = PM_PACKED_ABGR; break;
        case FLUC_TTMLBLEND_FORMAT_RGBA: pp.mode = PM_PACKED_RGBA; break;
        default: pp.mode = PM_PACKED_BGRA; break;
      }
      pp.out_a = w; pp.out_c = nullptr; pp.out_c2 = nullptr;
      pp.out_pitch = ref.pitch;
      pp.v0 = ref.v0;
      pp.row0 = cy0;
      pp.rows = cy1 - cy0;
      CU (c, launch_prepare (pp, ref.pitch / 4, c->up_stream));
      c->stats.prepare_launches++;
      ref.a = w;
      ref.c = nullptr;
      P->h_rects[0].push_back (ref);
    }
  }

  /* rectangle tables: one contiguous device array, plane after plane */
  {
    size_t total = 0;
    for (int pl = 0; pl < 3; pl++) {
      if (P->h_rects[pl].size () > FLUC_TTMLBLEND_MAX_RECTANGLES)
        return FLUC_TTMLBLEND_ERROR_TOO_MANY_RECTANGLES;
      P->rect_off[pl] = (int32_t) total;
      total += P->h_rects[pl].size ();
    }
    if (total) {
      P->h_rects_all.clear ();
      for (int pl = 0; pl < 3; pl++)
        P->h_rects_all.insert (P->h_rects_all.end (), P->h_rects[pl].begin (), P->h_rects[pl].end ());
      uint8_t *d;
      int rc;
      if ((rc = dev_alloc (c, P.get (), total * sizeof (RectRef), &d)))
        return rc;
      /* h_rects_all lives as long as the Prepared: safe source for the async copy */
      CU (c, cudaMemcpyAsync (d, P->h_rects_all.data (), total * sizeof (RectRef),
              cudaMemcpyHostToDevice, c->up_stream));
      P->d_rects_all = reinterpret_cast<RectRef *> (d);
      for (int pl = 0; pl < 3; pl++)
        if (!P->h_rects[pl].empty ())
          P->d_rects[pl] = P->d_rects_all + P->rect_off[pl];
    }
  }
  CU (c, cudaEventCreateWithFlags (&P->ready, cudaEventDisableTiming));
  CU (c, cudaEventRecord (P->ready, c->up_stream));
  *out = P.get ();
  ov->prepared.push_back (std::move (P));
  return 0;
}

/* ---------------------------------------------------------------------- */
/* plane jobs                                                             */

/* Decomposes possibly overlapping region rectangles into disjoint ones that
 * cover the same pixels, so that every pixel of the ttmlrender image is
 * blended exactly once. */
std::vector<FlucTtmlBlendRect>
disjoint_cover (const std::vector<FlucTtmlBlendRect> &in)
{
  std::vector<int> ys;
  for (auto &r : in) {
    ys.push_back (r.y);
    ys.push_back (r.y + r.h);
  }
  std::sort (ys.begin (), ys.end ());
  ys.erase (std::unique (ys.begin (), ys.end ()), ys.end ());
  std::vector<FlucTtmlBlendRect> out;
  for (size_t i = 0; i + 1 < ys.size (); i++) {
    const int y0 = ys[i], y1 = ys[i + 1];
    std::vector<std::pair<int, int>> xs;
    for (auto &r : in)
      if (r.y <= y0 && r.y + r.h >= y1)
        xs.push_back ({ r.x, r.x + r.w });
    std::sort (xs.begin (), xs.end ());
    std::vector<std::pair<int, int>> merged;
    for (auto &x : xs) {
      if (!merged.empty () && x.first <= merged.back ().second)
        merged.back ().second = std::max (merged.back ().second, x.second);
      else
        merged.push_back (x);
    }
    for (auto &m : merged) {
      bool grown = false;
      for (auto &o : out)
        if (o.x == m.first && o.w == m.second - m.first && o.y + o.h == y0) {
          o.h += y1 - y0;
          grown = true;
          break;
        }
      if (!grown)
        out.push_back ({ m.first, y0, m.second - m.first, y1 - y0 });
    }
  }
  return out;
}

/* Cuts a rectangle down to where it is not transparent: runs of non-empty
 * rows (text lines, boxes) become separate sub-rectangles, each as wide as its
 * outermost non-transparent pixels. Pixels with alpha 0 never change the frame
 * (BLENDSPEC section 2, `continue`), so dropping them is exact; what it buys is
 * that ttmlrender's frame-sized, mostly empty image costs overlay reads, ALU
 * work and -- for host frames -- PCIe traffic only where there is a cue. */
void
crop_runs (const std::vector<int2> &spans, int min_gap, size_t max_runs, std::vector<FlucTtmlBlendRect> &out)
{
  struct Run { int y0, y1, x0, x1; };
  std::vector<Run> runs;
  for (int y = 0; y < (int) spans.size (); y++) {
    if (spans[y].y < spans[y].x)
      continue;
    if (!runs.empty () && y - runs.back ().y1 < min_gap) {
      Run &r = runs.back ();
      r.y1 = y + 1;
      r.x0 = std::min (r.x0, spans[y].x);
      r.x1 = std::max (r.x1, spans[y].y + 1);
    } else {
      runs.push_back ({ y, y + 1, spans[y].x, spans[y].y + 1 });
    }
  }
  while (runs.size () > max_runs) {
    size_t best = 0;
    for (size_t i = 1; i + 1 < runs.size (); i++)
      if (runs[i + 1].y0 - runs[i].y1 < runs[best + 1].y0 - runs[best].y1)
        best = i;
    runs[best].y1 = runs[best + 1].y1;
    runs[best].x0 = std::min (runs[best].x0, runs[best + 1].x0);
    runs[best].x1 = std::max (runs[best].x1, runs[best + 1].x1);
    runs.erase (runs.begin () + best + 1);
  }
  for (const Run &r : runs)
    out.push_back ({ r.x0, r.y0, r.x1 - r.x0, r.y1 - r.y0 });
}

int
overlay_install (Ctx *c, uint32_t stream, const FlucTtmlBlendRectangle *rects, uint32_t n)
{
  NvtxRange nvtx ("ttmlblend.overlay_set");
  std::shared_ptr<Overlay> ov (new Overlay ());
  ov->ctx = c;
  struct Up { RawRect rr; int2 *d_spans; std::vector<int2> spans; };
  std::vector<Up> ups;
  for (uint32_t i = 0; i < n; i++) {
    const FlucTtmlBlendRectangle &r = rects[i];
    if (!r.pixels || r.width <= 0 || r.height <= 0 || r.stride < r.width * 4)
      return FLUC_TTMLBLEND_ERROR_INVALID_ARGUMENT;
    /* gst_video_blend: negative offsets skip source columns / rows */
    const int xoff = r.x < 0 ? -r.x : 0, yoff = r.y < 0 ? -r.y : 0;
    if (xoff >= r.width || yoff >= r.height)
      continue;
    Up u;
    RawRect &rr = u.rr;
    rr.w = r.width - xoff;
    rr.h = r.height - yoff;
    rr.x = r.x + xoff;
    rr.y = r.y + yoff;
    rr.ga = (int) (255.0 * r.global_alpha);
    rr.ga = std::max (0, std::min (255, rr.ga));
    rr.premul = (r.flags & FLUC_TTMLBLEND_FLAG_PREMULTIPLIED_ALPHA) != 0;
    rr.pitch = (int) align_up ((size_t) rr.w * 4, 256);
    ov->declared.push_back ({ rr.x, rr.y, rr.w, rr.h });
    void *d = nullptr;
    CU (c, cudaMallocAsync (&d, (size_t) rr.pitch * rr.h, c->up_stream));
    ov->raw_allocs.push_back (d);
    rr.dev = static_cast<uint8_t *> (d);
    CU (c, cudaMemcpy2DAsync (rr.dev, rr.pitch, r.pixels + (size_t) yoff * r.stride + (size_t) xoff * 4,
            r.stride, (size_t) rr.w * 4, rr.h, cudaMemcpyHostToDevice, c->up_stream));
    c->stats.h2d_bytes += (uint64_t) rr.w * 4 * rr.h;
    u.d_spans = nullptr;
    if (c->autocrop) {
      void *sp = nullptr;
      CU (c, cudaMallocAsync (&sp, (size_t) rr.h * sizeof (int2), c->up_stream));
      u.d_spans = static_cast<int2 *> (sp);
      u.spans.resize (rr.h);
      CU (c, launch_rowspan (rr.dev, rr.pitch, rr.w, rr.h, u.d_spans, c->up_stream));
      CU (c, cudaMemcpyAsync (u.spans.data (), u.d_spans, (size_t) rr.h * sizeof (int2),
              cudaMemcpyDeviceToHost, c->up_stream));
      CU (c, cudaFreeAsync (sp, c->up_stream));
    }
    ups.push_back (std::move (u));
  }
  /* the caller's pixels must be consumed (and the row spans back) before we return */
  CU (c, cudaStreamSynchronize (c->up_stream));
  for (Up &u : ups) {
    if (!c->autocrop) {
      ov->rects.push_back (u.rr);
      continue;
    }
    std::vector<FlucTtmlBlendRect> subs;
    /* at most 8 runs per rectangle, and never more sub-rectangles than the 64-bit band masks hold */
    crop_runs (u.spans, 16, std::max<size_t> (1, std::min<size_t> (8, FLUC_TTMLBLEND_MAX_RECTANGLES / ups.size ())), subs);
    for (const FlucTtmlBlendRect &s : subs) {
      RawRect q = u.rr;
      q.dev = u.rr.dev + (size_t) s.y * u.rr.pitch + (size_t) s.x * 4;
      q.x = u.rr.x + s.x;
      q.y = u.rr.y + s.y;
      q.w = s.w;
      q.h = s.h;
      ov->rects.push_back (q);
    }
  }
  if (ov->rects.size () > FLUC_TTMLBLEND_MAX_RECTANGLES)
    return FLUC_TTMLBLEND_ERROR_TOO_MANY_RECTANGLES;
  c->overlays[stream] = ov;       /* frames already queued keep the old one */
  c->stats.overlays_set++;
  return 0;
}

}  // namespace tbh
