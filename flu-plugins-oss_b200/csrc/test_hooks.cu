/*
 * test_hooks.cu -- C entry points into the HOST logic of the runtime (no GPU work), built
 * into a separate libfluc_ttmlblend_testhooks.so for the CPU test-suite only: the job
 * builder (bands, windows, classes, groups), the region-box decomposition, the
 * row-run crop and the row plan of the rectangle scaler. Never linked into libfluc_ttmlblend.so.
 */
#include "ttmlblend_internal.h"

using namespace tbh;

extern "C" {

/* One prepared rectangle as the job builder sees it on one plane. */
struct TbHookRect { int32_t v0, v1, y0, y1; };

/* One job back: window, class, mask and flags. */
struct TbHookJob {
  int32_t plane, cls, flags;
  int32_t win_v0, win_nv, win_y0, win_rows;
  uint32_t n_chunks, div_magic;
  uint64_t rect_mask;
  int32_t one_rect;
  int32_t grouped;              /* ended up in the band list of a group launch */
};

/* Runs build_jobs (+ make_groupable) for a frame whose planes sit at fake, 16-byte aligned
 * (or, with `misalign`, odd) addresses. rects[p] / n_rects[p]: prepared rectangles of plane p.
 * Returns the number of jobs written (<= max_jobs), or -1. *algo_bytes gets the byte count. */
__attribute__ ((visibility ("default"))) int
tb_hook_build_jobs (int format, int W, int H, int windowed, int misalign,
    const TbHookRect *rects0, int n0, const TbHookRect *rects1, int n1, const TbHookRect *rects2, int n2,
    TbHookJob *out, int max_jobs, uint64_t *algo_bytes, uint32_t *chunks_per_frame)
{
  if (!format_valid (format))
    return -1;
  format = format_canon (format);
  Prepared prep;
  prep.format = format;
  prep.W = W;
  prep.H = H;
  const TbHookRect *rr[3] = { rects0, rects1, rects2 };
  const int nn[3] = { n0, n1, n2 };
  static RectRef fake_table[3 * 64];
  for (int pl = 0; pl < 3; pl++) {
    for (int i = 0; i < nn[pl]; i++) {
      RectRef r = {};
      r.v0 = rr[pl][i].v0; r.v1 = rr[pl][i].v1; r.y0 = rr[pl][i].y0; r.y1 = rr[pl][i].y1;
      r.pitch = (r.v1 - r.v0) * 16;
      r.ga = 255;
      prep.h_rects[pl].push_back (r);
    }
    prep.d_rects[pl] = fake_table + 64 * pl;
    prep.rect_off[pl] = 64 * pl;
  }
  prep.d_rects_all = fake_table;
  FlucTtmlBlendFrame src = {}, dst = {};
  uintptr_t base = 0x10000000u + (misalign ? 4 : 0);
  for (int pl = 0; pl < format_planes (format); pl++) {
    const int stride = (int) align_up ((size_t) plane_row_bytes (format, pl, W), 256) + (misalign ? 4 : 0);
    src.plane[pl] = (void *) base;
    src.stride[pl] = stride;
    dst.plane[pl] = (void *) (windowed ? base : base + 0x40000000u);
    dst.stride[pl] = stride;
    base += (uintptr_t) stride * plane_rows (format, pl, H) + 4096;
  }
  Layout L;
  L.kind = plane_kind (format);
  L.algo_bytes = build_jobs (format, W, H, 0, &src, &dst, &prep, windowed != 0, L.jobs);
  std::vector<PlaneJob> all = L.jobs;
  make_groupable (L, false);
  if (algo_bytes)
    *algo_bytes = L.algo_bytes;
  if (chunks_per_frame)
    *chunks_per_frame = L.grouped ? L.chunks_per_frame : 0;
  int n = 0;
  for (const PlaneJob &j : all) {
    if (n >= max_jobs)
      return -1;
    TbHookJob &o = out[n++];
    o.plane = j.plane; o.cls = j.cls; o.flags = j.flags;
    o.win_v0 = j.win_v0; o.win_nv = j.win_nv; o.win_y0 = j.win_y0; o.win_rows = j.win_rows;
    o.n_chunks = j.n_chunks; o.div_magic = j.div_magic;
    o.rect_mask = j.rect_mask; o.one_rect = j.one_rect;
    o.grouped = (L.grouped && (j.flags & JF_FAST)) ? 1 : 0;
  }
  return n;
}

__attribute__ ((visibility ("default"))) int
tb_hook_disjoint_cover (const FlucTtmlBlendRect *in, int n, FlucTtmlBlendRect *out, int max_out)
{
  std::vector<FlucTtmlBlendRect> v (in, in + n);
  std::vector<FlucTtmlBlendRect> r = disjoint_cover (v);
  if ((int) r.size () > max_out)
    return -1;
  for (size_t i = 0; i < r.size (); i++)
    out[i] = r[i];
  return (int) r.size ();
}

/* spans: (first, last) non-transparent x per row, (w, -1) for an empty row */
__attribute__ ((visibility ("default"))) int
tb_hook_crop_runs (const int32_t *first, const int32_t *last, int rows, int min_gap, int max_runs,
    FlucTtmlBlendRect *out, int max_out)
{
  std::vector<int2> spans (rows);
  for (int i = 0; i < rows; i++)
    spans[i] = make_int2 (first[i], last[i]);
  std::vector<FlucTtmlBlendRect> r;
  crop_runs (spans, min_gap, (size_t) max_runs, r);
  if ((int) r.size () > max_out)
    return -1;
  for (size_t i = 0; i < r.size (); i++)
    out[i] = r[i];
  return (int) r.size ();
}

/* Layout cache, canonical ids and multi-layout packing, without a GPU: `n_frames` GRAY8 frames
 * of W x H; frame i belongs to "overlay" ov[i] (frames of one overlay share a Prepared) whose
 * single rectangle is rects[ov[i]], and has stride[i]. Out per frame: the layout's id and which
 * multi-layout launch (index) it was packed into; returns the number of launches, and in
 * *n_band_lists the distinct band lists the launches carry in total. */
__attribute__ ((visibility ("default"))) int
tb_hook_pack_layouts (int W, int H, int n_frames, const int32_t *ov, int n_overlays, const TbHookRect *rects,
    const int32_t *stride, uint64_t *layout_id, int32_t *launch_of, int32_t *n_band_lists)
{
  const int format = FLUC_TTMLBLEND_FORMAT_GRAY8;
  Ctx c;                                /* no CUDA object is created or touched */
  static RectRef fake_table[64];
  std::vector<std::unique_ptr<Prepared>> preps;
  for (int k = 0; k < n_overlays; k++) {
    std::unique_ptr<Prepared> p (new Prepared ());
    p->format = format;
    p->W = W;
    p->H = H;
    RectRef r = {};
    r.v0 = rects[k].v0; r.v1 = rects[k].v1; r.y0 = rects[k].y0; r.y1 = rects[k].y1;
    r.pitch = (r.v1 - r.v0) * 16;
    r.ga = 255;
    p->h_rects[0].push_back (r);
    p->d_rects[0] = fake_table;
    p->d_rects_all = fake_table;
    preps.push_back (std::move (p));
  }
  std::vector<PendingFrame> frames ((size_t) n_frames);
  std::vector<std::unique_ptr<MultiGroup>> multis;
  for (int i = 0; i < n_frames; i++) {
    if (ov[i] < 0 || ov[i] >= n_overlays || stride[i] < W)
      return -1;
    FlucTtmlBlendFrame src = {}, dst = {};
    src.plane[0] = (void *) (uintptr_t) (0x10000000u + 0x100000u * (unsigned) i);
    dst.plane[0] = (void *) (uintptr_t) (0x50000000u + 0x100000u * (unsigned) i);
    src.stride[0] = dst.stride[0] = stride[i];
    PendingFrame &f = frames[(size_t) i];
    f.prep = preps[(size_t) ov[i]].get ();
    f.layout = find_layout (&c, f.prep, false, format, W, H, 0, &src, &dst, false);
    f.src[0] = static_cast<const uint8_t *> (src.plane[0]);
    f.dst[0] = static_cast<uint8_t *> (dst.plane[0]);
    layout_id[i] = f.layout->id;
    launch_of[i] = -1;
    if (!f.layout->grouped)
      continue;
    for (size_t k = 0; k < multis.size () && launch_of[i] < 0; k++)
      if (multi_add (*multis[k], f))
        launch_of[i] = (int32_t) k;
    if (launch_of[i] < 0) {
      multis.emplace_back (new MultiGroup ());
      multi_start (*multis.back (), f);
      if (multi_add (*multis.back (), f))
        launch_of[i] = (int32_t) multis.size () - 1;
    }
  }
  int lists = 0;
  for (auto &m : multis) {
    lists += (int) m->layouts.size ();
    /* the parameters must be self-consistent: frames partition the chunks, band lists in range */
    const MultiParams &P = m->P;
    for (uint32_t k = 0; k < P.n_frames; k++) {
      if (P.frame_begin[k + 1] <= P.frame_begin[k] || P.frame_band0[k] + P.frame_nbands[k] > m->n_bands)
        return -2;
      uint32_t chunks = 0;
      for (uint32_t b = 0; b < P.frame_nbands[k]; b++) {
        if (P.bands[P.frame_band0[k] + b].chunk_begin != chunks)
          return -3;
        chunks += P.bands[P.frame_band0[k] + b].n_chunks;
      }
      if (chunks != P.frame_begin[k + 1] - P.frame_begin[k])
        return -4;
    }
  }
  if (n_band_lists)
    *n_band_lists = lists;
  return (int) multis.size ();
}

/* rows[3*i..] = (source row a, source row b, weight) of destination row i */
__attribute__ ((visibility ("default"))) int
tb_hook_scale_row_plan (int src_h, int dst_h, int32_t *rows)
{
  if (src_h < 2 || dst_h < 1)
    return -1;
  const std::vector<int4> plan = scale_row_plan (src_h, dst_h);
  for (size_t i = 0; i < plan.size (); i++) {
    rows[3 * i] = plan[i].x;
    rows[3 * i + 1] = plan[i].y;
    rows[3 * i + 2] = plan[i].z;
  }
  return (int) plan.size ();
}

/* The hazard tracker's range set: ops[i] = (kind, lo, hi) with kind 0 = add, 1 = query; every
 * query's answer goes to out[] in order. Returns the number of ranges held at the end, which
 * are written to ranges[2*k], ranges[2*k+1] (up to max_ranges). */
__attribute__ ((visibility ("default"))) int
tb_hook_interval_set (const uint64_t *ops, int n_ops, int32_t *out, uint64_t *ranges, int max_ranges)
{
  IntervalSet s;
  int q = 0;
  for (int i = 0; i < n_ops; i++) {
    const uint64_t kind = ops[3 * i], lo = ops[3 * i + 1], hi = ops[3 * i + 2];
    if (kind == 0)
      s.add ((uintptr_t) lo, (uintptr_t) hi);
    else
      out[q++] = s.overlaps ((uintptr_t) lo, (uintptr_t) hi) ? 1 : 0;
  }
  for (size_t k = 0; k < s.v.size () && (int) k < max_ranges; k++) {
    ranges[2 * k] = s.v[k].first;
    ranges[2 * k + 1] = s.v[k].second;
  }
  return (int) s.v.size ();
}

}  /* extern "C" */
