/*
 * jobs.cu -- turns one frame into kernel work: planes are cut into bands at the rectangles'
 * edges and into windows left to right, each window gets its class (copy / one / bulk /
 * general); frames of equal geometry are collected into groups for the parameter-table kernel.
 */
#include "ttmlblend_internal.h"

namespace tbh {

int
check_frame (int fmt, int W, int H, const FlucTtmlBlendFrame *f)
{
  if (!f)
    return FLUC_TTMLBLEND_ERROR_INVALID_ARGUMENT;
  if (!format_valid (fmt))
    return FLUC_TTMLBLEND_ERROR_UNSUPPORTED_FORMAT;
  if (W <= 0 || H <= 0 || W > 32768 || H > 32768)
    return FLUC_TTMLBLEND_ERROR_INVALID_ARGUMENT;
  for (int pl = 0; pl < format_planes (fmt); pl++)
    if (!f->plane[pl] || f->stride[pl] < plane_row_bytes (fmt, pl, W))
      return FLUC_TTMLBLEND_ERROR_INVALID_ARGUMENT;
  return 0;
}

void
push_window (std::vector<PlaneJob> &jobs, PlaneJob base, int v0, int v1, int y0, int y1)
{
  if (v1 <= v0 || y1 <= y0)
    return;
  const uint32_t nv = (uint32_t) (v1 - v0);
  /* magic division exactness: item * e < 2^32 with e = magic*nv - 2^32 < nv */
  const uint64_t magic = ((1ull << 32) + nv - 1) / nv;
  const uint64_t e = magic * nv - (1ull << 32);
  uint64_t max_items = e ? ((1ull << 32) - 1) / e : (1ull << 31);
  max_items = std::min<uint64_t> (max_items, 1ull << 31);
  int max_rows = (int) std::max<uint64_t> (1, std::min<uint64_t> (max_items / nv, 1 << 30));
  for (int r0 = y0; r0 < y1; r0 += max_rows) {
    PlaneJob j = base;
    j.win_v0 = v0;
    j.win_nv = (int32_t) nv;
    j.win_y0 = r0;
    j.win_rows = std::min (max_rows, y1 - r0);
    j.div_magic = (uint32_t) magic;     /* nv == 1 -> 2^32 truncates to 0; kernel special-cases it */
    const uint64_t items = (uint64_t) nv * (uint64_t) j.win_rows;
    j.n_chunks = (uint32_t) ((items + kItemsPerChunk - 1) / kItemsPerChunk);
    jobs.push_back (j);
  }
}

/* A window goes to the fast kernel where whole 16-byte vectors can be moved
 * (aligned frame, vector inside row_bytes); a ragged last vector column and
 * unaligned frames go to the byte-granular variant. */
void
push_split (std::vector<PlaneJob> &jobs, PlaneJob b, bool aligned, int v0, int v1, int y0, int y1)
{
  const int nv_full = b.row_bytes / 16;
  if (!aligned) {
    b.flags &= ~(JF_VECTOR | JF_FAST);
    push_window (jobs, b, v0, v1, y0, y1);
    return;
  }
  PlaneJob f = b;
  f.flags |= JF_VECTOR | JF_FAST;
  push_window (jobs, f, v0, std::min (v1, nv_full), y0, y1);
  if (v1 > nv_full) {
    PlaneJob t = b;
    t.flags = (t.flags | JF_VECTOR) & ~JF_FAST;
    push_window (jobs, t, std::max (v0, nv_full), v1, y0, y1);
  }
}

/* Builds the jobs of one frame: every plane is cut into bands of rows at the
 * top and bottom edges of the prepared rectangles, so that each band sees a
 * fixed set of rectangles and its class (copy / one rectangle / general) is
 * decided here, once, instead of per vector on the GPU. Returns the
 * algorithmic bytes moved. */
uint64_t
build_jobs (int format, int W, int H, uint32_t frame_flags, const FlucTtmlBlendFrame *src,
    const FlucTtmlBlendFrame *dst, const Prepared *prep, bool windowed, std::vector<PlaneJob> &jobs)
{
  const int n_planes = format_planes (format);
  static const bool use_bulk = !getenv ("FLUC_TTMLBLEND_BULK") || atoi (getenv ("FLUC_TTMLBLEND_BULK")) != 0;
  /* windowed: only the vectors a rectangle covers are read and written (in
   * place, or host frames where untouched bytes never cross PCIe) */
  const bool inplace = windowed;
  uint64_t bytes = 0;
  std::vector<int> ys;
  for (int pl = 0; pl < n_planes; pl++) {
    PlaneJob b = {};
    b.src = static_cast<const uint8_t *> (src->plane[pl]);
    b.dst = static_cast<uint8_t *> (dst->plane[pl]);
    b.src_pitch = src->stride[pl];
    b.dst_pitch = dst->stride[pl];
    b.row_bytes = plane_row_bytes (format, pl, W);
    b.kind = plane_kind (format);
    b.plane = pl;
    const int rows = plane_rows (format, pl, H);
    const bool aligned = (((uintptr_t) b.src | (uintptr_t) b.dst | (uintptr_t) b.src_pitch |
            (uintptr_t) b.dst_pitch) & 15u) == 0;
    b.flags = (inplace ? JF_INPLACE : 0) |
        ((frame_flags & FLUC_TTMLBLEND_FLAG_PREMULTIPLIED_ALPHA) ? JF_DST_PREMUL : 0);
    static const std::vector<RectRef> none;
    const std::vector<RectRef> &rects = prep ? prep->h_rects[pl] : none;
    b.rects = prep ? prep->d_rects[pl] : nullptr;
    const int nv_row = ceil_div (b.row_bytes, 16);

    ys.clear ();
    ys.push_back (0);
    ys.push_back (rows);
    for (const RectRef &r : rects) {
      ys.push_back (std::max (0, std::min (rows, r.y0)));
      ys.push_back (std::max (0, std::min (rows, r.y1)));
    }
    std::sort (ys.begin (), ys.end ());
    ys.erase (std::unique (ys.begin (), ys.end ()), ys.end ());

    for (size_t bi = 0; bi + 1 < ys.size (); bi++) {
      const int ya = ys[bi], yb = ys[bi + 1];
      /* rectangles over this band, by first column; then cut the band into
       * windows left to right: gaps copy, a rectangle alone in its columns is
       * JC_ONE (no per-vector tests), rectangles sharing columns form one
       * JC_GENERAL window that applies them in blend order */
      struct InBand { int idx, v0, v1; };
      InBand in_band[FLUC_TTMLBLEND_MAX_RECTANGLES];
      int n_in = 0;
      for (size_t i = 0; i < rects.size (); i++) {
        const RectRef &r = rects[i];
        if (r.y0 <= ya && r.y1 >= yb && r.v0 < nv_row && r.v1 > 0)
          in_band[n_in++] = { (int) i, std::max (r.v0, 0), std::min (r.v1, nv_row) };
      }
      std::sort (in_band, in_band + n_in, [](const InBand &a, const InBand &c) { return a.v0 < c.v0; });
      auto emit = [&](int cls, unsigned long long mask, int one, int v0, int v1) {
        if (v1 <= v0)
          return;
        PlaneJob j = b;
        j.rect_mask = mask;
        j.one_rect = one < 0 ? 0 : one;
        j.cls = cls;
        /* prepared rows packed over exactly the window's columns: the overlay bytes of any
         * run of vectors of the band are contiguous -> TMA bulk staging in the group kernel */
        if (cls == JC_ONE && use_bulk && rects[one].v0 == v0 && rects[one].v1 == v1 &&
            rects[one].pitch == (v1 - v0) * 16 && (b.row_bytes & 15) == 0)
          j.cls = JC_ONE_BULK;
        push_split (jobs, j, aligned, v0, v1, ya, yb);
        bytes += 2ull * (uint64_t) std::min ((v1 - v0) * 16, b.row_bytes - v0 * 16) * (uint64_t) (yb - ya);
      };
      int cursor = 0;
      for (int i = 0; i < n_in;) {
        unsigned long long mask = 1ull << in_band[i].idx;
        int c0 = in_band[i].v0, c1 = in_band[i].v1, k = i + 1;
        while (k < n_in && in_band[k].v0 < c1) {      /* shares columns with the cluster */
          mask |= 1ull << in_band[k].idx;
          c1 = std::max (c1, in_band[k].v1);
          k++;
        }
        if (!inplace)
          emit (JC_COPY, 0, -1, cursor, c0);
        emit (k - i == 1 ? JC_ONE : JC_GENERAL, mask, in_band[i].idx, c0, c1);
        cursor = c1;
        i = k;
      }
      if (!inplace)
        emit (JC_COPY, 0, -1, cursor, nv_row);
    }
  }
  if (prep)
    bytes += 4ull * prep->overlay_px;
  return bytes;
}

/* Moves the fast jobs of a frame into a band list for the group kernel. */
void
make_groupable (PendingFrame &f, const FlucTtmlBlendFrame *src, const FlucTtmlBlendFrame *dst)
{
  size_t n_fast = 0;
  for (const PlaneJob &j : f.jobs)
    n_fast += (j.flags & JF_FAST) ? 1 : 0;
  if (n_fast == 0 || n_fast > (size_t) kMaxGroupBands)
    return;
  std::vector<PlaneJob> rest;
  uint32_t total = 0;
  int gflags = -1;
  for (const PlaneJob &j : f.jobs) {
    if (!(j.flags & JF_FAST)) {
      rest.push_back (j);
      continue;
    }
    BandDesc b = {};
    b.chunk_begin = total;
    b.plane = j.plane;
    b.win_v0 = j.win_v0;
    b.win_nv = j.win_nv;
    b.win_y0 = j.win_y0;
    b.win_rows = j.win_rows;
    b.div_magic = j.div_magic;
    b.cls = j.cls;
    b.one_rect = j.one_rect;
    b.rect_mask_lo = (uint32_t) j.rect_mask;
    b.rect_mask_hi = (uint32_t) (j.rect_mask >> 32);
    b.n_chunks = j.n_chunks;
    total += j.n_chunks;
    f.bands.push_back (b);
    gflags = j.flags & (JF_INPLACE | JF_DST_PREMUL);
  }
  /* frame = umulhi (chunk, ceil (2^32 / cpf)) must be exact for every chunk of a full group */
  const uint64_t magic = ((1ull << 32) + total - 1) / total;
  const uint64_t e = magic * total - (1ull << 32);
  if ((uint64_t) kMaxGroupFrames * total * e >= (1ull << 32) || (uint64_t) kMaxGroupFrames * total >= (1ull << 26)) {
    f.bands.clear ();
    return;
  }
  if ((gflags & JF_INPLACE) && f.overlay && f.overlay->lazy_inplace)
    gflags |= JF_LAZY;
  for (const PlaneJob &j : f.jobs)
    if (j.flags & JF_FAST)
      f.gjobs.push_back (j);
  f.jobs.swap (rest);
  f.grouped = true;
  f.chunks_per_frame = total;
  f.gflags = gflags;
  for (int pl = 0; pl < 3; pl++) {
    f.ptrs.src[pl] = static_cast<const uint8_t *> (src->plane[pl]);
    f.ptrs.dst[pl] = static_cast<uint8_t *> (dst->plane[pl]);
    f.src_pitch[pl] = src->stride[pl];
    f.dst_pitch[pl] = dst->stride[pl];
    f.rect_off[pl] = f.prep ? f.prep->rect_off[pl] : 0;
  }
  f.ptrs.rects = f.prep ? f.prep->d_rects_all : nullptr;
  f.ptrs.pad_ = 0;
}

bool
group_accepts (const Group &g, const PendingFrame &f)
{
  const GroupParams &P = g.P;
  if (g.kind != f.kind || P.n_frames >= (uint32_t) kMaxGroupFrames || P.n_bands != f.bands.size () ||
      P.chunks_per_frame != f.chunks_per_frame || P.flags != f.gflags)
    return false;
  if (memcmp (P.src_pitch, f.src_pitch, sizeof P.src_pitch) || memcmp (P.dst_pitch, f.dst_pitch, sizeof P.dst_pitch) ||
      memcmp (P.rect_off, f.rect_off, sizeof P.rect_off))
    return false;
  return memcmp (P.bands, f.bands.data (), f.bands.size () * sizeof (BandDesc)) == 0;
}

void
group_start (Group &g, const PendingFrame &f)
{
  memset (&g.P, 0, sizeof g.P);
  g.kind = f.kind;
  g.P.n_bands = (uint32_t) f.bands.size ();
  g.P.chunks_per_frame = f.chunks_per_frame;
  g.P.cpf_magic = (uint32_t) (((1ull << 32) + f.chunks_per_frame - 1) / f.chunks_per_frame);
  g.P.flags = f.gflags;
  memcpy (g.P.src_pitch, f.src_pitch, sizeof g.P.src_pitch);
  memcpy (g.P.dst_pitch, f.dst_pitch, sizeof g.P.dst_pitch);
  memcpy (g.P.rect_off, f.rect_off, sizeof g.P.rect_off);
  memcpy (g.P.bands, f.bands.data (), f.bands.size () * sizeof (BandDesc));
}

void
multi_start (MultiGroup &m, const PendingFrame &f)
{
  m.kind = f.kind;
  m.n_bands = 0;
  m.layouts.clear ();
  /* only the header: frames and bands are written as they are added */
  m.P.n_frames = 0;
  m.P.flags = f.gflags;
  memcpy (m.P.src_pitch, f.src_pitch, sizeof m.P.src_pitch);
  memcpy (m.P.dst_pitch, f.dst_pitch, sizeof m.P.dst_pitch);
  m.P.frame_begin[0] = 0;
}

/* Adds a groupable frame to a multi-layout launch if it fits: same kind, pitches and flags,
 * room for the frame and -- unless an equal band list is already there -- for its bands. */
bool
multi_add (MultiGroup &m, const PendingFrame &f)
{
  MultiParams &P = m.P;
  if (m.kind != f.kind || P.flags != f.gflags || P.n_frames >= (uint32_t) kMaxGroupFrames ||
      memcmp (P.src_pitch, f.src_pitch, sizeof P.src_pitch) || memcmp (P.dst_pitch, f.dst_pitch, sizeof P.dst_pitch))
    return false;
  if ((uint64_t) P.frame_begin[P.n_frames] + f.chunks_per_frame >= (1ull << 26))
    return false;
  for (int pl = 0; pl < 3; pl++)
    if (f.rect_off[pl] < 0 || f.rect_off[pl] > 0xffff)
      return false;
  const size_t nb = f.bands.size ();
  int layout = -1;
  for (size_t i = 0; i < m.layouts.size (); i++)
    if (m.layouts[i].second == nb && memcmp (&P.bands[m.layouts[i].first], f.bands.data (), nb * sizeof (BandDesc)) == 0) {
      layout = (int) i;
      break;
    }
  if (layout < 0) {
    if (m.n_bands + nb > (size_t) kMaxMultiBands)
      return false;
    memcpy (&P.bands[m.n_bands], f.bands.data (), nb * sizeof (BandDesc));
    m.layouts.push_back ({ (uint16_t) m.n_bands, (uint16_t) nb });
    layout = (int) m.layouts.size () - 1;
    m.n_bands += (uint32_t) nb;
  }
  const uint32_t k = P.n_frames++;
  P.frame_band0[k] = m.layouts[layout].first;
  P.frame_nbands[k] = m.layouts[layout].second;
  P.frames[k] = f.ptrs;
  P.frames[k].pad_ = (uint64_t) f.rect_off[0] | ((uint64_t) f.rect_off[1] << 16) | ((uint64_t) f.rect_off[2] << 32);
  P.frame_begin[k + 1] = P.frame_begin[k] + f.chunks_per_frame;
  return true;
}

}  // namespace tbh
