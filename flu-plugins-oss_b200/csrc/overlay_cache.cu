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

DevBlock::~DevBlock ()
{
  if (ctx && ptr)
    free_deferred (ctx, { ptr });
}

/* Blocks nobody else holds are freed in one go (one fence per stream instead of one per block);
 * blocks a newer overlay of the stream shares (overlay_update) live on with it. */
Overlay::~Overlay ()
{
  std::vector<void *> ptrs;
  auto take = [&ptrs](std::shared_ptr<DevBlock> &b) {
    if (b && b.use_count () == 1 && b->ptr) {
      ptrs.push_back (b->ptr);
      b->ptr = nullptr;
    }
  };
  for (RawRect &r : rects)
    take (r.block);
  for (auto &p : prepared) {
    for (PreparedRect &pr : p->per_rect)
      for (auto &b : pr.blocks)
        take (b);
    take (p->table_block);
    if (p->ready)
      cudaEventDestroy (p->ready);
  }
  if (ctx)
    free_deferred (ctx, ptrs);
}

/* ---------------------------------------------------------------------- */
/* prepare: raw BGRA rectangle -> per-plane prepared overlay              */

/* a stream-ordered allocation on the upload stream */
static int
dev_block (Ctx *c, size_t bytes, std::shared_ptr<DevBlock> *out)
{
  void *ptr = nullptr;
  CU (c, cudaMallocFromPoolAsync (&ptr, std::max<size_t> (bytes, 16), c->mem_pool, c->up_stream));
  std::shared_ptr<DevBlock> b (new DevBlock ());
  b->ctx = c;
  b->ptr = ptr;
  *out = std::move (b);
  return 0;
}

static int
dev_alloc (Ctx *c, PreparedRect *pr, size_t bytes, uint8_t **out)
{
  std::shared_ptr<DevBlock> b;
  int rc = dev_block (c, bytes, &b);
  if (rc)
    return rc;
  *out = static_cast<uint8_t *> (b->ptr);
  pr->blocks.push_back (std::move (b));
  return 0;
}

static int prepare_build (Ctx *c, Overlay *ov, int format, int W, int H, std::unique_ptr<Prepared> &P);

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
  const int rc = prepare_build (c, ov, format, W, H, P);
  if (rc) {
    /* half-built (out of memory, too many rectangles): its blocks go back with it */
    if (P->ready)
      cudaEventDestroy (P->ready);
    return rc;
  }
  *out = P.get ();
  ov->prepared.push_back (std::move (P));
  return 0;
}

/* Prepares one rectangle for one destination format: allocates its planes, runs the prepare
 * kernel(s) on the upload stream, and notes how each destination plane sees it. */
static int
prepare_rect (Ctx *c, const RawRect &rr, int format, int W, int H, bool chroma_average, PreparedRect &out)
{
  const int kind = plane_kind (format);
  const int n_planes = format_planes (format);
  /* gst_video_blend clipping: rr is already clipped at the left/top */
  const int cx0 = rr.x, cy0 = rr.y;
  const int cx1 = std::min (rr.x + rr.w, W), cy1 = std::min (rr.y + rr.h, H);
  if (cx1 <= cx0 || cy1 <= cy0)
    return 0;
  if (rr.ga == 0)
    return 0;                 /* asrc == 0 everywhere: blends nothing */

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

  pp.sub_x = format_sub_x (format);
  pp.sub_y = format_sub_y (format);
  if (kind == PK_PLANE8 && format_packed_422 (format)) {
    /* YUY2 / UYVY: one plane of macropixels, every byte gets its own alpha + colour */
    RectRef ref = {};
    ref.v0 = (4 * (cx0 / 2)) / 16;
    ref.v1 = ceil_div (4 * ceil_div (cx1, 2), 16);
    ref.y0 = cy0;
    ref.y1 = cy1;
    ref.pitch = (ref.v1 - ref.v0) * 16;
    ref.ga = 255;
    const size_t bytes = (size_t) ref.pitch * (cy1 - cy0);
    uint8_t *a, *col;
    int rc;
    if ((rc = dev_alloc (c, &out, bytes, &a)) || (rc = dev_alloc (c, &out, bytes, &col)))
      return rc;
    ref.a = a;
    ref.c = col;
    pp.mode = format == FLUC_TTMLBLEND_FORMAT_YUY2 ? PM_YUY2 : format == FLUC_TTMLBLEND_FORMAT_UYVY ? PM_UYVY :
        format == FLUC_TTMLBLEND_FORMAT_YVYU ? PM_YVYU : PM_VYUY;
    pp.out_a = a; pp.out_c = col; pp.out_c2 = nullptr;
    pp.out_pitch = ref.pitch;
    pp.v0 = ref.v0;
    pp.row0 = cy0;
    pp.rows = cy1 - cy0;
    CU (c, launch_prepare (pp, ref.pitch / 4, c->up_stream));
    c->stats.prepare_launches++;
    out.refs[out.n_refs++] = { 0, ref };
  } else if (kind == PK_PLANE8_RGB || (kind == PK_PLANE8 && format_packed_444_3 (format))) {
    /* v308 / IYU2 / RGB / BGR: one plane, three bytes per pixel, every byte its own alpha +
     * colour; the RGB pair keeps the source colours and the rectangle's flags for the blend */
    RectRef ref = {};
    ref.v0 = (3 * cx0) / 16;
    ref.v1 = ceil_div (3 * cx1, 16);
    ref.y0 = cy0;
    ref.y1 = cy1;
    ref.pitch = (ref.v1 - ref.v0) * 16;
    ref.ga = 255;
    const size_t bytes = (size_t) ref.pitch * (cy1 - cy0);
    uint8_t *a, *col;
    int rc;
    if ((rc = dev_alloc (c, &out, bytes, &a)) || (rc = dev_alloc (c, &out, bytes, &col)))
      return rc;
    ref.a = a;
    ref.c = col;
    if (kind == PK_PLANE8_RGB) {
      ref.ga = rr.ga;
      ref.src_premul = rr.premul ? 1 : 0;
    }
    pp.mode = format == FLUC_TTMLBLEND_FORMAT_v308 ? PM_V308 : format == FLUC_TTMLBLEND_FORMAT_IYU2 ? PM_IYU2 :
        format == FLUC_TTMLBLEND_FORMAT_RGB ? PM_RGB24 : PM_BGR24;
    pp.out_a = a; pp.out_c = col; pp.out_c2 = nullptr;
    pp.out_pitch = ref.pitch;
    pp.v0 = ref.v0;
    pp.row0 = cy0;
    pp.rows = cy1 - cy0;
    CU (c, launch_prepare (pp, ref.pitch, c->up_stream));
    c->stats.prepare_launches++;
    out.refs[out.n_refs++] = { 0, ref };
  } else if (kind == PK_PLANE8) {
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
      if ((rc = dev_alloc (c, &out, bytes, &a)) || (rc = dev_alloc (c, &out, bytes, &y)))
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
      out.refs[out.n_refs++] = { 0, ref };
    }
    /* chroma: the samples sited on the even pixel of each pair / the even line of each
     * line pair, as far as the format subsamples (parity); with the non-parity 2x2 average
     * of 4:2:0 every sample with at least one covered pixel */
    const int sx = pp.sub_x, sy = pp.sub_y;
    const bool avg = chroma_average && sx == 2 && sy == 2;
    pp.chroma_average = avg ? 1 : 0;
    const int bx0 = avg ? cx0 / 2 : ceil_div (cx0, sx), bx1 = ceil_div (cx1, sx);
    const int by0 = avg ? cy0 / 2 : ceil_div (cy0, sy), by1 = ceil_div (cy1, sy);
    if (n_planes >= 2 && bx1 > bx0 && by1 > by0) {
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
        if ((rc = dev_alloc (c, &out, bytes, &a)) || (rc = dev_alloc (c, &out, bytes, &u))
            || (rc = dev_alloc (c, &out, bytes, &v)))
          return rc;
        pp.mode = PM_CHROMA_PLANAR;
        pp.out_a = a; pp.out_c = u; pp.out_c2 = v;
        pp.out_pitch = ref.pitch;
        pp.v0 = ref.v0;
        pp.row0 = by0;
        pp.rows = by1 - by0;
        CU (c, launch_prepare (pp, ref.pitch, c->up_stream));
        c->stats.prepare_launches++;
        const int pu = format == FLUC_TTMLBLEND_FORMAT_YV12 ? 2 : 1;
        const int pv = 3 - pu;
        ref.a = a;
        ref.c = u;
        out.refs[out.n_refs++] = { pu, ref };
        ref.c = v;
        out.refs[out.n_refs++] = { pv, ref };
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
        if ((rc = dev_alloc (c, &out, bytes, &a)) || (rc = dev_alloc (c, &out, bytes, &uv)))
          return rc;
        pp.mode = (format == FLUC_TTMLBLEND_FORMAT_NV21 || format == FLUC_TTMLBLEND_FORMAT_NV61) ?
            PM_CHROMA_VU : PM_CHROMA_UV;
        pp.out_a = a; pp.out_c = uv; pp.out_c2 = nullptr;
        pp.out_pitch = ref.pitch;
        pp.v0 = ref.v0;
        pp.row0 = by0;
        pp.rows = by1 - by0;
        CU (c, launch_prepare (pp, ref.pitch / 2, c->up_stream));
        c->stats.prepare_launches++;
        ref.a = a;
        ref.c = uv;
        out.refs[out.n_refs++] = { 1, ref };
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
    if ((rc = dev_alloc (c, &out, bytes, &w)))
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
    out.refs[out.n_refs++] = { 0, ref };
  }
  return 0;
}

static int
prepare_assemble (Ctx *c, Prepared *P)
{
  for (int pl = 0; pl < 3; pl++)
    P->h_rects[pl].clear ();
  for (const PreparedRect &pr : P->per_rect)
    for (int k = 0; k < pr.n_refs; k++)
      P->h_rects[pr.refs[k].plane].push_back (pr.refs[k].ref);
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
      int rc;
      if ((rc = dev_block (c, total * sizeof (RectRef), &P->table_block)))
        return rc;
      uint8_t *d = static_cast<uint8_t *> (P->table_block->ptr);
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
  return 0;
}

static int
prepare_build (Ctx *c, Overlay *ov, int format, int W, int H, std::unique_ptr<Prepared> &P)
{
  P->format = format;
  P->W = W;
  P->H = H;
  P->chroma_average = c->chroma_average;
  for (const FlucTtmlBlendRect &d : ov->declared) {
    const int w = std::min (d.x + d.w, W) - d.x, h = std::min (d.y + d.h, H) - d.y;
    if (w > 0 && h > 0)
      P->overlay_px += (uint64_t) w * (uint64_t) h;
  }
  P->per_rect.resize (ov->rects.size ());
  for (size_t i = 0; i < ov->rects.size (); i++) {
    const int rc = prepare_rect (c, ov->rects[i], format, W, H, P->chroma_average, P->per_rect[i]);
    if (rc)
      return rc;
  }
  return prepare_assemble (c, P.get ());
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

static std::shared_ptr<DevBlock>
make_block (Ctx *c, void *ptr)
{
  std::shared_ptr<DevBlock> b (new DevBlock ());
  b->ctx = c;
  b->ptr = ptr;
  return b;
}

/* A rectangle whose pixels already sit in device memory, on its way into an overlay. */
struct Up {
  RawRect rr;
  std::vector<int2> spans;
  std::vector<int> groups;      /* per row: 16-pixel groups with any alpha */
  std::vector<int> raw;         /* both as they come back from the GPU in one copy: h x int2, then h x int */
};

/* Is this pointer device memory (a cue produced on the GPU, or left there by a previous stage)? */
static bool
on_device (const void *p)
{
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes (&attr, p) != cudaSuccess) {
    cudaGetLastError ();
    return false;
  }
  return attr.type == cudaMemoryTypeDevice;
}

/* What an install does to the context's counters, added once the lock is held again. */
struct InstallStats {
  uint64_t h2d_bytes = 0;
  uint32_t launches = 0;
};

/* CU for the parts of an install that run without the context lock: the lock is taken back
 * before the error is recorded and reported. */
#define CUU(ctx, lk, call) do {                                              \
    cudaError_t eu_ = (call);                                                \
    if (eu_ != cudaSuccess) {                                                \
      if (!(lk).owns_lock ())                                                \
        (lk).lock ();                                                        \
      (ctx)->cuda_error = std::string (#call) + ": " + cudaGetErrorString (eu_); \
      if (eu_ == cudaErrorMemoryAllocation) {                                \
        cudaGetLastError ();                                                 \
        return FLUC_TTMLBLEND_ERROR_OUT_OF_MEMORY;                           \
      }                                                                      \
      (ctx)->sticky = FLUC_TTMLBLEND_ERROR_CUDA;                             \
      return FLUC_TTMLBLEND_ERROR_CUDA;                                      \
    }                                                                        \
  } while (0)

/* Queues the row-span scan of one device-resident rectangle on the upload stream. */
static int
scan_rows (Ctx *c, std::unique_lock<std::mutex> &lk, Up &u)
{
  if (!c->autocrop)
    return 0;
  void *sp = nullptr;
  const size_t nb = (size_t) u.rr.h * (sizeof (int2) + sizeof (int));
  CUU (c, lk, cudaMallocFromPoolAsync (&sp, nb, c->mem_pool, c->up_stream));
  u.raw.resize ((size_t) u.rr.h * 3);
  int2 *d_spans = static_cast<int2 *> (sp);
  int *d_groups = reinterpret_cast<int *> (d_spans + u.rr.h);
  CUU (c, lk, launch_rowspan (u.rr.dev, u.rr.pitch, u.rr.w, u.rr.h, d_spans, d_groups, c->up_stream));
  CUU (c, lk, cudaMemcpyAsync (u.raw.data (), sp, nb, cudaMemcpyDeviceToHost, c->up_stream));   /* one copy for both */
  CUU (c, lk, cudaFreeAsync (sp, c->up_stream));
  return 0;
}

/* after the upload stream has been waited for: the scan's two arrays apart */
static void
unpack_scan (Up &u)
{
  if (u.raw.empty ())
    return;
  const size_t h = (size_t) u.rr.h;
  u.spans.resize (h);
  u.groups.resize (h);
  memcpy (u.spans.data (), u.raw.data (), h * sizeof (int2));
  memcpy (u.groups.data (), u.raw.data () + 2 * h, h * sizeof (int));
  u.raw.clear ();
}

/* One uploaded rectangle becomes a box of the overlay: cropped to its non-transparent row runs
 * (or kept whole without auto-crop), with the sparsity figures of what remains. */
static void
append_box (Ctx *c, Overlay *ov, Up &u, size_t n_boxes)
{
  unpack_scan (u);
  OverlayBox box;
  box.declared = { u.rr.x, u.rr.y, u.rr.w, u.rr.h };
  box.first_rect = (uint32_t) ov->rects.size ();
  if (!c->autocrop) {
    ov->rects.push_back (u.rr);
  } else {
    std::vector<FlucTtmlBlendRect> subs;
    /* at most 8 runs per rectangle, and never more sub-rectangles than the 64-bit band masks hold */
    crop_runs (u.spans, 16, std::max<size_t> (1, std::min<size_t> (8, FLUC_TTMLBLEND_MAX_RECTANGLES / std::max<size_t> (1, n_boxes))), subs);
    for (const FlucTtmlBlendRect &s : subs) {
      /* how sparse is what remains after the crop: 16-pixel groups with some alpha against
       * all groups of the kept sub-rectangles */
      box.groups_all += (uint64_t) ceil_div (s.w, 16) * (uint64_t) s.h;
      for (int y = s.y; y < s.y + s.h; y++) {
        box.groups_on += (uint64_t) (u.groups[y] & 0xffff);
        if (u.rr.ga == 255)
          box.groups_opaque += (uint64_t) (u.groups[y] >> 16);
      }
      RawRect q = u.rr;
      q.dev = u.rr.dev + (size_t) s.y * u.rr.pitch + (size_t) s.x * 4;
      q.x = u.rr.x + s.x;
      q.y = u.rr.y + s.y;
      q.w = s.w;
      q.h = s.h;
      ov->rects.push_back (q);
    }
  }
  box.n_rects = (uint32_t) ov->rects.size () - box.first_rect;
  ov->boxes.push_back (box);
}

/* Text without a background box leaves most 16-byte vectors under the cue untouched, and
 * under an opaque box (alpha 255) the result does not depend on the frame at all. In place
 * (dst == src, host frames over PCIe) it then pays to look at the overlay before touching
 * the frame: transparent vectors are skipped, opaque ones are written without being read.
 * Under a translucent box it costs latency for nothing (tools/lazy_probe.py).
 * FLUC_TTMLBLEND_LAZY=0/1 forces it. */
static void
decide_lazy (Overlay *ov)
{
  static const char *lazy_env = getenv ("FLUC_TTMLBLEND_LAZY");
  uint64_t groups_all = 0, groups_on = 0, groups_opaque = 0;
  for (const OverlayBox &b : ov->boxes) {
    groups_all += b.groups_all;
    groups_on += b.groups_on;
    groups_opaque += b.groups_opaque;
  }
  ov->transparent_fraction = groups_all ? 1.0 - (double) groups_on / (double) groups_all : 0.0;
  ov->opaque_fraction = groups_all ? (double) groups_opaque / (double) groups_all : 0.0;
  ov->lazy_inplace = lazy_env ? atoi (lazy_env) != 0 : ov->transparent_fraction + ov->opaque_fraction >= 0.3;
  /* out of place every vector is written anyway; what can be saved is the read under an opaque
   * box (FLUC_TTMLBLEND_OPAQUE_SKIP=0/1 forces it) */
  static const char *skip_env = getenv ("FLUC_TTMLBLEND_OPAQUE_SKIP");
  ov->opaque_skip = skip_env ? atoi (skip_env) != 0 : ov->opaque_fraction >= 0.3;
}

/* Waits for uploads and scans, crops every rectangle to its non-transparent row runs and
 * swaps the stream's overlay. */
static int
finish_install (Ctx *c, std::unique_lock<std::mutex> &lk, uint32_t stream, std::shared_ptr<Overlay> ov,
    std::vector<Up> &ups, const InstallStats &st)
{
  /* the caller's pixels must be consumed (and the row spans back) before we return; the
   * context stays unlocked while that takes its time */
  CUU (c, lk, cudaStreamSynchronize (c->up_stream));
  if (!lk.owns_lock ())
    lk.lock ();
  if (c->sticky)
    return c->sticky;
  c->stats.h2d_bytes += st.h2d_bytes;
  c->stats.prepare_launches += st.launches;
  for (Up &u : ups)
    append_box (c, ov.get (), u, ups.size ());
  if (ov->rects.size () > FLUC_TTMLBLEND_MAX_RECTANGLES)
    return FLUC_TTMLBLEND_ERROR_TOO_MANY_RECTANGLES;
  decide_lazy (ov.get ());
  /* The stream's frames will very likely keep the format and size they had: prepare the new
   * cue for them now, on the upload stream, so that the first frame after a cue change only
   * has an event to wait for instead of the prepare launches in front of it. */
  auto prev = c->overlays.find (stream);
  if (prev != c->overlays.end () && c->eager_prepare) {
    for (auto &p : prev->second->prepared) {
      if (!p->used || p->chroma_average != c->chroma_average)
        continue;
      Prepared *unused = nullptr;
      const int rc = prepare_overlay (c, ov.get (), p->format, p->W, p->H, &unused);
      if (rc)
        return rc;
    }
  }
  c->overlays[stream] = ov;       /* frames already queued keep the old one */
  c->stats.overlays_set++;
  return 0;
}

/* gst_video_blend_scale_linear_RGBA's vertical pass, row by row: upstream keeps the
 * horizontally resampled source lines in a two-line cache (slot = row & 1) and tracks with
 * `y1` how far it has resampled. Simulating that bookkeeping here gives, for every
 * destination row, the two source rows and the 8-bit weight the kernel merges -- including
 * the rows where the cache holds something else than rows j and j+1 (docs/BLENDSPEC.md
 * section 10). */
std::vector<int4>
scale_row_plan (int src_h, int dst_h)
{
  const int y_inc = dst_h == 1 ? 0 : ((src_h - 1) << 16) / (dst_h - 1) - 1;
  int slot[2] = { 0, 0 };
  int y1 = 0, acc = 0;
  std::vector<int4> plan ((size_t) dst_h);
  for (int i = 0; i < dst_h; i++) {
    const int j = acc >> 16, x = acc & 0xffff;
    if (x == 0) {
      plan[i] = make_int4 (slot[j & 1], slot[j & 1], 0, 0);       /* memcpy of LINE (j) */
    } else {
      if (j > y1) {
        slot[j & 1] = j;
        y1++;
      }
      if (j >= y1) {
        slot[(j + 1) & 1] = j + 1;
        y1++;
      }
      plan[i] = make_int4 (slot[j & 1], slot[(j + 1) & 1], x >> 8, 0);
    }
    acc += y_inc;
  }
  return plan;
}

int
overlay_install (Ctx *c, std::unique_lock<std::mutex> &lk, uint32_t stream, const FlucTtmlBlendRectangle *rects,
    uint32_t n, int image_w, int image_h)
{
  NvtxRange nvtx ("ttmlblend.overlay_set");
  for (uint32_t i = 0; i < n; i++) {
    const FlucTtmlBlendRectangle &r = rects[i];
    if (!r.pixels || r.width <= 0 || r.height <= 0 || r.stride < r.width * 4 || r.render_width < 0 ||
        r.render_height < 0 || r.render_width > 32768 || r.render_height > 32768)
      return FLUC_TTMLBLEND_ERROR_INVALID_ARGUMENT;
    const int rw = r.render_width ? r.render_width : r.width, rh = r.render_height ? r.render_height : r.height;
    if ((rw != r.width || rh != r.height) && (r.width < 2 || r.height < 2))
      return FLUC_TTMLBLEND_ERROR_INVALID_ARGUMENT;   /* upstream reads outside the image here */
  }
  std::shared_ptr<Overlay> ov (new Overlay ());
  ov->ctx = c;
  /* image_w > 0: the rectangles are the (disjoint) region boxes of one image of that size at
   * the frame's origin -- the ttmlrender form, which overlay_update can patch later */
  ov->image_w = image_w;
  ov->image_h = image_h;
  ov->updatable = image_w > 0 && image_h > 0;
  std::vector<Up> ups;
  std::vector<std::vector<int4>> plans;     /* host side of async uploads: alive until finish_install */
  InstallStats st;
  /* Uploading (a copy from pageable memory holds the caller for its whole length), scaling and
   * scanning touch only this overlay-to-be and the upload stream: the context is unlocked
   * meanwhile, so a cue change of one stream does not stall the frames of all the others. Any
   * failure takes the lock back before it reports (CUU). */
  lk.unlock ();
  for (uint32_t i = 0; i < n; i++) {
    const FlucTtmlBlendRectangle &r = rects[i];
    /* gst_video_overlay_rectangle_needs_scaling: render size != pixel size */
    const int rw = r.render_width ? r.render_width : r.width, rh = r.render_height ? r.render_height : r.height;
    const bool scaled = rw != r.width || rh != r.height;
    /* gst_video_blend: negative offsets skip source columns / rows */
    const int xoff = r.x < 0 ? -r.x : 0, yoff = r.y < 0 ? -r.y : 0;
    if (xoff >= rw || yoff >= rh)
      continue;
    Up u;
    RawRect &rr = u.rr;
    rr.w = rw - xoff;
    rr.h = rh - yoff;
    rr.x = r.x + xoff;
    rr.y = r.y + yoff;
    rr.ga = (int) (255.0 * r.global_alpha);
    rr.ga = std::max (0, std::min (255, rr.ga));
    rr.premul = (r.flags & FLUC_TTMLBLEND_FLAG_PREMULTIPLIED_ALPHA) != 0;
    ov->declared.push_back ({ rr.x, rr.y, rr.w, rr.h });
    if (!scaled) {
      rr.pitch = (int) align_up ((size_t) rr.w * 4, 256);
      void *d = nullptr;
      CUU (c, lk, cudaMallocFromPoolAsync (&d, (size_t) rr.pitch * rr.h, c->mem_pool, c->up_stream));
      rr.block = make_block (c, d);
      rr.dev = static_cast<uint8_t *> (d);
      /* cudaMemcpyDefault: the pixels may just as well be in device memory already (UVA tells) */
      CUU (c, lk, cudaMemcpy2DAsync (rr.dev, rr.pitch, r.pixels + (size_t) yoff * r.stride + (size_t) xoff * 4,
              r.stride, (size_t) rr.w * 4, rr.h, cudaMemcpyDefault, c->up_stream));
      if (!on_device (r.pixels))
        st.h2d_bytes += (uint64_t) rr.w * 4 * rr.h;
    } else {
      /* the whole source goes up, is scaled to the render size on the GPU, and the clipped
       * part of the scaled image is what gets blended */
      const int sp = (int) align_up ((size_t) r.width * 4, 256);
      void *s = nullptr, *d = nullptr, *p = nullptr;
      CUU (c, lk, cudaMallocFromPoolAsync (&s, (size_t) sp * r.height, c->mem_pool, c->up_stream));
      CUU (c, lk, cudaMemcpy2DAsync (s, sp, r.pixels, r.stride, (size_t) r.width * 4, r.height,
              cudaMemcpyDefault, c->up_stream));
      if (!on_device (r.pixels))
        st.h2d_bytes += (uint64_t) r.width * 4 * r.height;
      plans.push_back (scale_row_plan (r.height, rh));
      CUU (c, lk, cudaMallocFromPoolAsync (&p, (size_t) rh * sizeof (int4), c->mem_pool, c->up_stream));
      CUU (c, lk, cudaMemcpyAsync (p, plans.back ().data (), (size_t) rh * sizeof (int4), cudaMemcpyHostToDevice,
              c->up_stream));
      rr.pitch = (int) align_up ((size_t) rw * 4, 256);
      CUU (c, lk, cudaMallocFromPoolAsync (&d, (size_t) rr.pitch * rh, c->mem_pool, c->up_stream));
      rr.block = make_block (c, d);
      const int x_inc = rw == 1 ? 0 : ((r.width - 1) << 16) / (rw - 1) - 1;
      CUU (c, lk, launch_scale (static_cast<const uint8_t *> (s), sp, static_cast<const int4 *> (p), x_inc,
              static_cast<uint8_t *> (d), rr.pitch, rw, rh, c->up_stream));
      st.launches++;
      CUU (c, lk, cudaFreeAsync (s, c->up_stream));
      CUU (c, lk, cudaFreeAsync (p, c->up_stream));
      rr.dev = static_cast<uint8_t *> (d) + (size_t) yoff * rr.pitch + (size_t) xoff * 4;
    }
    int rc = scan_rows (c, lk, u);
    if (rc)
      return rc;
    ups.push_back (std::move (u));
  }
  return finish_install (c, lk, stream, ov, ups, st);
}

/* overlay_update: the stream's cue changed inside `changed` only (a <set> animation step, a
 * roll-up line: the producer re-renders its whole image for every timeline event,
 * /root/reference/plugins/ttml/gstttmlrender.c:1442-1452, /root/reference/plugins/ttml/gstttmlevent.c:208-233).
 * Boxes of the current overlay that the changed rectangles do not touch keep their device
 * pixels, their crop and their prepared planes -- the new overlay shares those blocks with the
 * old one -- and only the touched boxes are uploaded, scanned and prepared again. Returns
 * FLUC_TTMLBLEND_ERROR_NOT_FOUND when there is nothing to patch (no overlay from overlay_set of
 * the same image size, or a change outside every box): the caller installs the image whole. */
int
overlay_update_image (Ctx *c, std::unique_lock<std::mutex> &lk, uint32_t stream, const uint8_t *bgra, int w, int h,
    int stride, const FlucTtmlBlendRect *changed, uint32_t n_changed)
{
  NvtxRange nvtx ("ttmlblend.overlay_update");
  auto it = c->overlays.find (stream);
  if (it == c->overlays.end () || !it->second->updatable || it->second->image_w != w || it->second->image_h != h)
    return FLUC_TTMLBLEND_ERROR_NOT_FOUND;
  std::shared_ptr<Overlay> old = it->second;
  std::vector<bool> touched (old->boxes.size (), false);
  size_t n_touched = 0;
  for (uint32_t k = 0; k < n_changed; k++) {
    const int x0 = std::max (changed[k].x, 0), y0 = std::max (changed[k].y, 0);
    const int x1 = std::min (changed[k].x + changed[k].w, w), y1 = std::min (changed[k].y + changed[k].h, h);
    if (x1 <= x0 || y1 <= y0)
      continue;
    /* the boxes are disjoint: the change lies inside them iff the intersections add up to it */
    uint64_t covered = 0;
    for (size_t b = 0; b < old->boxes.size (); b++) {
      const FlucTtmlBlendRect &d = old->boxes[b].declared;
      const int ix0 = std::max (x0, d.x), iy0 = std::max (y0, d.y);
      const int ix1 = std::min (x1, d.x + d.w), iy1 = std::min (y1, d.y + d.h);
      if (ix1 > ix0 && iy1 > iy0) {
        covered += (uint64_t) (ix1 - ix0) * (uint64_t) (iy1 - iy0);
        if (!touched[b]) {
          touched[b] = true;
          n_touched++;
        }
      }
    }
    if (covered != (uint64_t) (x1 - x0) * (uint64_t) (y1 - y0))
      return FLUC_TTMLBLEND_ERROR_NOT_FOUND;
  }
  if (n_touched == 0)
    return 0;                   /* nothing changed */

  /* upload + scan of the touched boxes, context unlocked (as in overlay_install) */
  std::vector<Up> ups (old->boxes.size ());
  InstallStats st;
  lk.unlock ();
  for (size_t b = 0; b < old->boxes.size (); b++) {
    if (!touched[b])
      continue;
    const FlucTtmlBlendRect &d = old->boxes[b].declared;
    RawRect &rr = ups[b].rr;
    rr.x = d.x; rr.y = d.y; rr.w = d.w; rr.h = d.h;
    rr.ga = 255;
    rr.premul = true;
    rr.pitch = (int) align_up ((size_t) rr.w * 4, 256);
    void *dev = nullptr;
    CUU (c, lk, cudaMallocFromPoolAsync (&dev, (size_t) rr.pitch * rr.h, c->mem_pool, c->up_stream));
    rr.block = make_block (c, dev);
    rr.dev = static_cast<uint8_t *> (dev);
    CUU (c, lk, cudaMemcpy2DAsync (rr.dev, rr.pitch, bgra + (size_t) d.y * stride + (size_t) d.x * 4, stride,
            (size_t) rr.w * 4, rr.h, cudaMemcpyDefault, c->up_stream));
    if (!on_device (bgra))
      st.h2d_bytes += (uint64_t) rr.w * 4 * rr.h;
    int rc = scan_rows (c, lk, ups[b]);
    if (rc)
      return rc;
  }
  CUU (c, lk, cudaStreamSynchronize (c->up_stream));
  if (!lk.owns_lock ())
    lk.lock ();
  if (c->sticky)
    return c->sticky;
  auto now = c->overlays.find (stream);
  if (now == c->overlays.end () || now->second != old)
    return FLUC_TTMLBLEND_ERROR_NOT_FOUND;      /* replaced meanwhile by another thread */
  c->stats.h2d_bytes += st.h2d_bytes;

  std::shared_ptr<Overlay> nv (new Overlay ());
  nv->ctx = c;
  nv->image_w = w;
  nv->image_h = h;
  nv->updatable = true;
  nv->declared = old->declared;
  std::vector<int> from_old;    /* new rectangle -> the old overlay's rectangle it is, or -1 */
  for (size_t b = 0; b < old->boxes.size (); b++) {
    if (touched[b]) {
      append_box (c, nv.get (), ups[b], old->boxes.size ());
      from_old.resize (nv->rects.size (), -1);
    } else {
      OverlayBox box = old->boxes[b];
      box.first_rect = (uint32_t) nv->rects.size ();
      for (uint32_t k = 0; k < old->boxes[b].n_rects; k++) {
        nv->rects.push_back (old->rects[old->boxes[b].first_rect + k]);
        from_old.push_back ((int) (old->boxes[b].first_rect + k));
      }
      nv->boxes.push_back (box);
    }
  }
  if (nv->rects.size () > FLUC_TTMLBLEND_MAX_RECTANGLES)
    return FLUC_TTMLBLEND_ERROR_TOO_MANY_RECTANGLES;
  decide_lazy (nv.get ());
  /* prepared for the formats / sizes the stream's frames have been using: kept rectangles bring
   * their prepared planes along, the others are prepared now */
  for (auto &p : old->prepared) {
    if (!p->used || p->chroma_average != c->chroma_average || p->per_rect.size () != old->rects.size ())
      continue;
    std::unique_ptr<Prepared> P (new Prepared ());
    P->format = p->format;
    P->W = p->W;
    P->H = p->H;
    P->chroma_average = p->chroma_average;
    P->overlay_px = p->overlay_px;
    P->used = true;
    P->per_rect.resize (nv->rects.size ());
    int rc = 0;
    for (size_t i = 0; i < nv->rects.size () && !rc; i++) {
      if (from_old[i] >= 0)
        P->per_rect[i] = p->per_rect[(size_t) from_old[i]];
      else
        rc = prepare_rect (c, nv->rects[i], P->format, P->W, P->H, P->chroma_average, P->per_rect[i]);
    }
    if (!rc)
      rc = prepare_assemble (c, P.get ());
    if (rc) {
      if (P->ready)
        cudaEventDestroy (P->ready);
      return rc;
    }
    nv->prepared.push_back (std::move (P));
  }
  c->overlays[stream] = nv;       /* frames already queued keep the old one */
  c->stats.overlays_set++;
  c->stats.overlays_updated++;
  return 0;
}

/* Cairo's colour conversion: double components -> premultiplied 16-bit shorts
 * (_cairo_color_compute_shorts: d * 65535.0 + 0.5) -> pixman a8r8g8b8 (short >> 8). */
static uint32_t
cairo_solid_pixel (uint32_t rgba8888)
{
  /* GET_CAIRO_COMP, /root/reference/plugins/ttml/gstttmlrender.c:1178 */
  const double r = ((rgba8888 >> 24) & 255) / 255.0, g = ((rgba8888 >> 16) & 255) / 255.0;
  const double b = ((rgba8888 >> 8) & 255) / 255.0, a = (rgba8888 & 255) / 255.0;
  const uint32_t as = (uint16_t) (a * 65535.0 + 0.5), rs = (uint16_t) (r * a * 65535.0 + 0.5);
  const uint32_t gs = (uint16_t) (g * a * 65535.0 + 0.5), bs = (uint16_t) (b * a * 65535.0 + 0.5);
  return ((as >> 8) << 24) | ((rs >> 8) << 16) | ((gs >> 8) << 8) | (bs >> 8);
}

/* The overlay composed on the GPU from region descriptors:
 * a cleared frame-sized canvas, every region drawn onto it in list order by
 * ttmlblend_region_kernel, then installed like an uploaded image cropped to the boxes. */
int
overlay_install_regions (Ctx *c, std::unique_lock<std::mutex> &lk, uint32_t stream, int W, int H,
    const FlucTtmlBlendRegion *regions, uint32_t n)
{
  NvtxRange nvtx ("ttmlblend.overlay_set_regions");
  for (uint32_t i = 0; i < n; i++)
    if (regions[i].w <= 0 || regions[i].h <= 0 || !(regions[i].opacity >= 0.0 && regions[i].opacity <= 1.0) ||
        (regions[i].layer && regions[i].layer_stride < 4 * regions[i].w))
      return FLUC_TTMLBLEND_ERROR_INVALID_ARGUMENT;
  {
    std::vector<FlucTtmlBlendRect> boxes;
    for (uint32_t i = 0; i < n; i++) {
      const FlucTtmlBlendRegion &r = regions[i];
      const int x0 = std::max (r.x, 0), y0 = std::max (r.y, 0);
      const int x1 = std::min (r.x + r.w, W), y1 = std::min (r.y + r.h, H);
      if (x1 > x0 && y1 > y0 && (r.background_color || r.layer))
        boxes.push_back ({ x0, y0, x1 - x0, y1 - y0 });
    }
    if (disjoint_cover (boxes).size () > FLUC_TTMLBLEND_MAX_RECTANGLES)
      return FLUC_TTMLBLEND_ERROR_TOO_MANY_RECTANGLES;
  }
  std::shared_ptr<Overlay> ov (new Overlay ());
  ov->ctx = c;
  InstallStats st;
  lk.unlock ();                 /* as in overlay_install: only this overlay and the upload stream from here */
  const int pitch = (int) align_up ((size_t) W * 4, 256);
  void *canvas = nullptr;
  CUU (c, lk, cudaMallocFromPoolAsync (&canvas, (size_t) pitch * H, c->mem_pool, c->up_stream));
  std::shared_ptr<DevBlock> canvas_block = make_block (c, canvas);
  CUU (c, lk, cudaMemsetAsync (canvas, 0, (size_t) pitch * H, c->up_stream));   /* CAIRO_OPERATOR_CLEAR */
  std::vector<void *> layers;
  std::vector<FlucTtmlBlendRect> boxes;
  for (uint32_t i = 0; i < n; i++) {
    const FlucTtmlBlendRegion &r = regions[i];
    const int x0 = std::max (r.x, 0), y0 = std::max (r.y, 0);
    const int x1 = std::min (r.x + r.w, W), y1 = std::min (r.y + r.h, H);
    if (x1 <= x0 || y1 <= y0)
      continue;
    RegionParams p = {};
    p.canvas = static_cast<uint8_t *> (canvas);
    p.canvas_pitch = pitch;
    p.x = x0; p.y = y0; p.w = x1 - x0; p.h = y1 - y0;
    p.lx = x0 - r.x; p.ly = y0 - r.y;
    p.bg = r.background_color ? cairo_solid_pixel (r.background_color) : 0u;
    p.m8 = r.opacity < 1.0 ? ((uint32_t) (uint16_t) (r.opacity * 65535.0 + 0.5)) >> 8 : 255u;
    if (r.layer) {
      const int lp = (int) align_up ((size_t) r.w * 4, 256);
      void *d = nullptr;
      CUU (c, lk, cudaMallocFromPoolAsync (&d, (size_t) lp * r.h, c->mem_pool, c->up_stream));
      layers.push_back (d);
      CUU (c, lk, cudaMemcpy2DAsync (d, lp, r.layer, r.layer_stride, (size_t) r.w * 4, r.h,
              cudaMemcpyDefault, c->up_stream));
      if (!on_device (r.layer))
        st.h2d_bytes += (uint64_t) r.w * 4 * r.h;
      p.layer = static_cast<uint8_t *> (d);
      p.layer_pitch = lp;
    }
    if (!p.bg && !p.layer)
      continue;                 /* nothing to draw */
    CUU (c, lk, launch_region (p, c->up_stream));
    st.launches++;
    boxes.push_back ({ x0, y0, x1 - x0, y1 - y0 });
  }
  for (void *d : layers)
    CUU (c, lk, cudaFreeAsync (d, c->up_stream));
  std::vector<Up> ups;
  for (const FlucTtmlBlendRect &b : disjoint_cover (boxes)) {
    Up u;
    u.rr.block = canvas_block;
    u.rr.dev = static_cast<uint8_t *> (canvas) + (size_t) b.y * pitch + (size_t) b.x * 4;
    u.rr.pitch = pitch;
    u.rr.w = b.w; u.rr.h = b.h; u.rr.x = b.x; u.rr.y = b.y;
    u.rr.ga = 255;
    u.rr.premul = true;
    ov->declared.push_back (b);
    int rc = scan_rows (c, lk, u);
    if (rc)
      return rc;
    ups.push_back (std::move (u));
  }
  return finish_install (c, lk, stream, ov, ups, st);
}

}  // namespace tbh
