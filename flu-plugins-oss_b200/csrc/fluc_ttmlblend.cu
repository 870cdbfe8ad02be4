/*
 * fluc_ttmlblend.cu -- host side of libfluc_ttmlblend.so: the C ABI declared in
 * include/fluc_ttmlblend.h. Context (one per GPU), overlay cache, frame pool,
 * multi-stream batch scheduler and the host-frame (PCIe) path.
 *
 * This is the native runtime around the kernels in ttmlblend_kernels.cu. It
 * stands where the reference leaves compositing to GStreamer
 * (/root/reference/plugins/ttml/README.md:45-48) and is shaped after the
 * reference's own helper libraries: opaque object + monitor
 * (/root/reference/libs/fluc/flu-codec-sdk/fluc/threads/fluc_monitor.c:15-70),
 * worker thread draining a queue under that monitor
 * (/root/reference/libs/flu/downloader/lib/fludownloader.c:490-532), stats
 * copied out under the lock (.../bwmeter/fluc_bwmeter.c:71-76).
 *
 * No CPU fallback exists anywhere in this file: without a usable CUDA device
 * every entry point fails with FLUC_TTMLBLEND_ERROR_NO_DEVICE / _CUDA.
 */
#include "../../include/fluc_ttmlblend.h"
#include "ttmlblend_kernels.cuh"

#include <nvtx3/nvToolsExt.h>       /* header-only: ranges show up in nsys / ncu timelines */

#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

using namespace tb;

namespace {

/* ---------------------------------------------------------------------- */
/* format geometry                                                        */

enum FormatClass { FC_I420, FC_YV12, FC_NV12, FC_NV21, FC_AYUV, FC_ARGB, FC_ABGR, FC_RGBA, FC_BGRA, FC_COUNT };

inline bool
format_valid (int f)
{
  return f >= 0 && f < FLUC_TTMLBLEND_FORMAT_COUNT;
}

inline int
format_planes (int f)
{
  switch (f) {
    case FLUC_TTMLBLEND_FORMAT_I420:
    case FLUC_TTMLBLEND_FORMAT_YV12:
      return 3;
    case FLUC_TTMLBLEND_FORMAT_NV12:
    case FLUC_TTMLBLEND_FORMAT_NV21:
      return 2;
    default:
      return 1;
  }
}

inline int
plane_row_bytes (int f, int plane, int w)
{
  switch (f) {
    case FLUC_TTMLBLEND_FORMAT_I420:
    case FLUC_TTMLBLEND_FORMAT_YV12:
      return plane == 0 ? w : (w + 1) / 2;
    case FLUC_TTMLBLEND_FORMAT_NV12:
    case FLUC_TTMLBLEND_FORMAT_NV21:
      return plane == 0 ? w : 2 * ((w + 1) / 2);
    default:
      return 4 * w;
  }
}

inline int
plane_rows (int f, int plane, int h)
{
  return (format_planes (f) > 1 && plane > 0) ? (h + 1) / 2 : h;
}

inline int
plane_kind (int f)
{
  switch (f) {
    case FLUC_TTMLBLEND_FORMAT_AYUV:
    case FLUC_TTMLBLEND_FORMAT_ARGB:
    case FLUC_TTMLBLEND_FORMAT_ABGR:
      return PK_PACKED_A0;
    case FLUC_TTMLBLEND_FORMAT_RGBA:
    case FLUC_TTMLBLEND_FORMAT_BGRA:
      return PK_PACKED_A3;
    default:
      return PK_PLANE8;
  }
}

inline int ceil_div (int a, int b) { return (a + b - 1) / b; }
inline size_t align_up (size_t v, size_t a) { return (v + a - 1) / a * a; }

/* ---------------------------------------------------------------------- */
/* overlay cache                                                          */

struct Ctx;

/* device copy of one rectangle's BGRA pixels (left/top clipped at 0) */
struct RawRect {
  uint8_t *dev = nullptr;
  int pitch = 0, w = 0, h = 0;
  int x = 0, y = 0;
  int ga = 255;
  bool premul = true;
};

/* everything frame-independent, for one (format, W, H) */
struct Prepared {
  int format = -1, W = 0, H = 0;
  std::vector<void *> allocs;
  std::vector<RectRef> h_rects[3];     /* per plane, host copy */
  std::vector<RectRef> h_rects_all;
  RectRef *d_rects[3] = { nullptr, nullptr, nullptr };
  RectRef *d_rects_all = nullptr;      /* the three tables, contiguous */
  int32_t rect_off[3] = { 0, 0, 0 };   /* first entry of plane p in d_rects_all */
  uint64_t overlay_px = 0;             /* sum of clipped w*h */
  cudaEvent_t ready = nullptr;
  bool blend_waited = false;           /* blend stream already ordered after `ready` */
};

struct Overlay {
  Ctx *ctx = nullptr;
  std::vector<RawRect> rects;          /* what gets prepared: cropped to non-transparent pixels */
  std::vector<void *> raw_allocs;      /* device copies the rects point into */
  std::vector<FlucTtmlBlendRect> declared;   /* rectangles as handed in (algorithmic bytes) */
  std::vector<std::unique_ptr<Prepared>> prepared;
  ~Overlay ();
};

struct PendingFrame {
  uint64_t ticket;
  std::shared_ptr<Overlay> overlay;
  Prepared *prep;
  int kind;
  std::vector<PlaneJob> jobs;          /* generic-kernel jobs (byte-granular parts, odd frames) */
  uint64_t algo_bytes;
  /* group launch: the fast windows as a band list + this frame's pointers */
  bool grouped = false;
  std::vector<BandDesc> bands;
  FramePtrs ptrs;
  int32_t src_pitch[3], dst_pitch[3], rect_off[3], gflags;
  uint32_t chunks_per_frame = 0;
  const void *dst0 = nullptr;
};

/* frames that can share one launch: everything but the pointers is equal */
struct Group {
  int kind;
  GroupParams P;
};

struct Batch {
  uint64_t last_ticket;
  cudaEvent_t done;
  cudaEvent_t t0, t1;                  /* profiling pair (may be null) */
  std::vector<std::shared_ptr<Overlay>> keep;
};

struct TableSlot {
  PlaneJob *h_jobs = nullptr, *d_jobs = nullptr;
  uint32_t *h_begin = nullptr, *d_begin = nullptr;
  size_t cap = 0;
  cudaEvent_t copied = nullptr;        /* last kernel that read the slot has finished */
  cudaEvent_t uploaded = nullptr;      /* table copy has landed */
};

struct PoolEntry {
  void *base;
  size_t bytes;
  int fmt, W, H, on_host;
  FlucTtmlBlendFrame frame;
};

struct Lane {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  uint64_t ticket = 0;
  bool busy = false;
  uint8_t *dev = nullptr;
  size_t dev_bytes = 0;
  TableSlot table[2];
  std::shared_ptr<Overlay> keep;
};

/* How blend_host moves a device-accessible (pinned) host frame. */
enum HostMode { HM_STAGED = 0, HM_ZEROCOPY = 1, HM_WRITEBACK = 2 };

constexpr int kLanes = 4;
constexpr int kTableSlots = 8;

struct Ctx {
  int device = 0;
  std::mutex mu;
  std::condition_variable cv;
  int sticky = 0;
  std::string cuda_error;

  cudaStream_t blend_stream = nullptr, up_stream = nullptr, reaper = nullptr, table_stream = nullptr;
  cudaEvent_t ev_fence[kLanes + 2] = {};
  cudaEvent_t timer0 = nullptr, timer1 = nullptr;

  std::unordered_map<uint32_t, std::shared_ptr<Overlay>> overlays;

  std::vector<PendingFrame> pending;
  std::vector<Group> groups;           /* scratch of launch_pending */
  std::unordered_set<const void *> pending_dst;   /* destination buffers queued in `pending` */
  std::vector<cudaEvent_t> timing_pool;
  std::chrono::steady_clock::time_point oldest_pending;
  uint64_t next_ticket = 0;
  std::deque<Batch> batches;
  std::vector<cudaEvent_t> event_pool;
  TableSlot slots[kTableSlots];
  int next_slot = 0;

  uint32_t max_batch = 32, linger_us = 200;
  int host_mode = HM_ZEROCOPY;
  bool autocrop = true;                /* FLUC_TTMLBLEND_AUTOCROP=0: blend rectangles as handed in */
  bool use_groups = true;              /* FLUC_TTMLBLEND_GROUPS=0: generic table kernel only */
  bool profiling = false;
  uint32_t profile_every = 1, profile_seq = 0;   /* FLUC_TTMLBLEND_PROFILE_EVERY */
  std::thread sched;
  bool quit = false;

  Lane lanes[kLanes];
  int next_lane = 0;
  std::map<uint64_t, int> lane_tickets;

  std::vector<PoolEntry> pool_free, pool_used;
  uint8_t *scrub = nullptr;
  size_t scrub_bytes = 0;

  FlucTtmlBlendStats stats = {};
};

#define CU(ctx, call) do {                                                   \
    cudaError_t e_ = (call);                                                 \
    if (e_ != cudaSuccess) {                                                 \
      (ctx)->sticky = FLUC_TTMLBLEND_ERROR_CUDA;                             \
      (ctx)->cuda_error = std::string (#call) + ": " + cudaGetErrorString (e_); \
      return e_ == cudaErrorMemoryAllocation ?                               \
          FLUC_TTMLBLEND_ERROR_OUT_OF_MEMORY : FLUC_TTMLBLEND_ERROR_CUDA;    \
    }                                                                        \
  } while (0)

int
log_level ()
{
  static int lvl = -1;
  if (lvl < 0) {
    const char *e = getenv ("FLUC_TTMLBLEND_DEBUG");
    lvl = e ? atoi (e) : 0;
  }
  return lvl;
}

/* NVTX range for the current scope */
struct NvtxRange {
  explicit NvtxRange (const char *name) { nvtxRangePushA (name); }
  ~NvtxRange () { nvtxRangePop (); }
};

#define TBLOG(n, ...) do { if (log_level () >= (n)) { fprintf (stderr, "ttmlblend: " __VA_ARGS__); fputc ('\n', stderr); } } while (0)

cudaEvent_t
event_get (Ctx *c)
{
  if (!c->event_pool.empty ()) {
    cudaEvent_t e = c->event_pool.back ();
    c->event_pool.pop_back ();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreateWithFlags (&e, cudaEventDisableTiming);
  return e;
}

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
    if (p->format == format && p->W == W && p->H == H) {
      *out = p.get ();
      return 0;
    }

  NvtxRange nvtx ("ttmlblend.prepare_overlay");
  TBLOG (2, "prepare overlay: format %d, %dx%d, %zu rectangle(s)", format, W, H, ov->rects.size ());
  std::unique_ptr<Prepared> P (new Prepared ());
  P->format = format;
  P->W = W;
  P->H = H;
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
      /* chroma: the samples sited on even x / even y */
      const int bx0 = ceil_div (cx0, 2), bx1 = ceil_div (cx1, 2);
      const int by0 = ceil_div (cy0, 2), by1 = ceil_div (cy1, 2);
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

int
slot_reserve (Ctx *c, TableSlot &s, size_t n)
{
  if (s.cap >= n)
    return 0;
  const size_t cap = std::max<size_t> (256, n * 2);
  if (s.h_jobs) cudaFreeHost (s.h_jobs);
  if (s.h_begin) cudaFreeHost (s.h_begin);
  if (s.d_jobs) cudaFree (s.d_jobs);
  if (s.d_begin) cudaFree (s.d_begin);
  s.cap = 0;
  CU (c, cudaHostAlloc ((void **) &s.h_jobs, cap * sizeof (PlaneJob), cudaHostAllocDefault));
  CU (c, cudaHostAlloc ((void **) &s.h_begin, cap * sizeof (uint32_t), cudaHostAllocDefault));
  CU (c, cudaMalloc ((void **) &s.d_jobs, cap * sizeof (PlaneJob)));
  CU (c, cudaMalloc ((void **) &s.d_begin, cap * sizeof (uint32_t)));
  if (!s.copied)
    CU (c, cudaEventCreateWithFlags (&s.copied, cudaEventDisableTiming));
  if (!s.uploaded)
    CU (c, cudaEventCreateWithFlags (&s.uploaded, cudaEventDisableTiming));
  s.cap = cap;
  return 0;
}

/* Copies `jobs` (one PlaneKind) into a table slot and launches the kernel. */
int
launch_jobs (Ctx *c, TableSlot &s, const PlaneJob *jobs, size_t n, int kind, bool fast,
    cudaStream_t stream)
{
  if (n == 0)
    return 0;
  if (s.copied && s.cap)
    CU (c, cudaEventSynchronize (s.copied));
  int rc = slot_reserve (c, s, n);
  if (rc)
    return rc;
  uint32_t total = 0;
  for (size_t i = 0; i < n; i++) {
    s.h_jobs[i] = jobs[i];
    s.h_begin[i] = total;
    total += jobs[i].n_chunks;
  }
  /* the table goes up on the copy stream, so that it overlaps the kernel still
   * running on `stream`; the slot is free (its last kernel waited on above) */
  CU (c, cudaMemcpyAsync (s.d_jobs, s.h_jobs, n * sizeof (PlaneJob), cudaMemcpyHostToDevice, c->table_stream));
  CU (c, cudaMemcpyAsync (s.d_begin, s.h_begin, n * sizeof (uint32_t), cudaMemcpyHostToDevice, c->table_stream));
  CU (c, cudaEventRecord (s.uploaded, c->table_stream));
  CU (c, cudaStreamWaitEvent (stream, s.uploaded, 0));
  CU (c, launch_blend (s.d_jobs, s.d_begin, (int) n, total, kind, fast, stream));
  CU (c, cudaEventRecord (s.copied, stream));    /* slot busy until this kernel is done */
  c->stats.launches++;
  return 0;
}

void
reap_batches (Ctx *c)
{
  while (!c->batches.empty ()) {
    Batch &b = c->batches.front ();
    if (cudaEventQuery (b.done) != cudaSuccess)
      break;
    if (b.t0 && b.t1) {
      float ms = 0.f;
      if (cudaEventElapsedTime (&ms, b.t0, b.t1) == cudaSuccess) {
        c->stats.kernel_ms += ms;
        c->stats.kernel_ms_launches++;
      }
      c->timing_pool.push_back (b.t0);
      c->timing_pool.push_back (b.t1);
    }
    c->event_pool.push_back (b.done);
    c->batches.pop_front ();
  }
}

/* Launches everything pending as one batch (per plane kind). mu held. */
int
launch_pending (Ctx *c)
{
  if (c->pending.empty ())
    return 0;
  NvtxRange nvtx ("ttmlblend.launch_batch");
  reap_batches (c);
  Batch b = {};
  b.last_ticket = c->pending.back ().ticket;
  std::vector<PlaneJob> by_kind[6];     /* PlaneKind x {byte-granular, fast} */
  std::vector<Group> &groups = c->groups;
  groups.clear ();
  for (PendingFrame &f : c->pending) {
    for (const PlaneJob &j : f.jobs)
      by_kind[f.kind * 2 + ((j.flags & JF_FAST) ? 1 : 0)].push_back (j);
    if (f.grouped) {
      Group *g = nullptr;
      for (Group &o : groups)
        if (group_accepts (o, f)) {
          g = &o;
          break;
        }
      if (!g) {
        groups.emplace_back ();
        g = &groups.back ();
        group_start (*g, f);
      }
      g->P.frames[g->P.n_frames++] = f.ptrs;
    }
    if (f.overlay)
      b.keep.push_back (f.overlay);
    if (f.prep && !f.prep->blend_waited) {
      /* once per prepared overlay: later launches follow in stream order */
      f.prep->blend_waited = true;
      CU (c, cudaStreamWaitEvent (c->blend_stream, f.prep->ready, 0));
    }
    c->stats.frames_blended++;
    c->stats.algorithmic_bytes += f.algo_bytes;
  }
  c->pending.clear ();
  c->pending_dst.clear ();
  size_t n_launches = groups.size ();
  for (int k = 0; k < 6; k++)
    n_launches += !by_kind[k].empty ();
  /* an event pair keeps the batch from overlapping its neighbours (~2 us of stream time):
   * sample every profile_every-th batch. A batch of several launches (several groups /
   * kinds) is timed from before its first to after its last launch. */
  if (c->profiling && n_launches >= 1 && (c->profile_seq++ % c->profile_every) == 0) {
    for (cudaEvent_t *e : { &b.t0, &b.t1 }) {
      if (!c->timing_pool.empty ()) {
        *e = c->timing_pool.back ();
        c->timing_pool.pop_back ();
      } else {
        CU (c, cudaEventCreate (e));
      }
    }
  }
  if (b.t0)
    CU (c, cudaEventRecord (b.t0, c->blend_stream));
  for (Group &g : groups) {
    CU (c, launch_group (g.P, g.kind, c->blend_stream));
    c->stats.launches++;
    c->stats.group_launches++;
  }
  for (int k = 0; k < 6; k++) {
    if (by_kind[k].empty ())
      continue;
    TableSlot &s = c->slots[c->next_slot];
    c->next_slot = (c->next_slot + 1) % kTableSlots;
    int rc = launch_jobs (c, s, by_kind[k].data (), by_kind[k].size (), k / 2, (k & 1) != 0,
        c->blend_stream);
    if (rc)
      return rc;
  }
  if (b.t1)
    CU (c, cudaEventRecord (b.t1, c->blend_stream));
  b.done = event_get (c);
  CU (c, cudaEventRecord (b.done, c->blend_stream));
  c->batches.push_back (std::move (b));
  return 0;
}

void
scheduler_main (Ctx *c)
{
  cudaSetDevice (c->device);
  std::unique_lock<std::mutex> lk (c->mu);
  while (!c->quit) {
    if (c->pending.empty () || c->linger_us == 0) {
      c->cv.wait (lk);
      continue;
    }
    const auto deadline = c->oldest_pending + std::chrono::microseconds (c->linger_us);
    if (std::chrono::steady_clock::now () >= deadline) {
      if (!c->sticky)
        launch_pending (c);
      else {
        c->pending.clear ();
        c->pending_dst.clear ();
      }
    } else {
      c->cv.wait_until (lk, deadline);
    }
  }
}

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

int
lane_reserve (Ctx *c, Lane &l, size_t bytes)
{
  if (l.dev_bytes >= bytes)
    return 0;
  if (l.dev)
    CU (c, cudaFree (l.dev));
  l.dev = nullptr;
  l.dev_bytes = 0;
  CU (c, cudaMalloc ((void **) &l.dev, bytes));
  l.dev_bytes = bytes;
  return 0;
}

}  // namespace

struct _FlucTtmlBlend {
  Ctx c;
};

#define ENTER(thiz)                                                          \
  if (!(thiz)) return FLUC_TTMLBLEND_ERROR_INVALID_ARGUMENT;                 \
  Ctx *c = &(thiz)->c;                                                       \
  std::unique_lock<std::mutex> lk (c->mu);                                   \
  if (c->sticky) return c->sticky;                                           \
  cudaSetDevice (c->device)

extern "C" {

const char *
fluc_ttmlblend_version (void)
{
  return "fluc_ttmlblend 0.1 (sm_100a)";
}

const char *
fluc_ttmlblend_strerror (int err)
{
  switch (err) {
    case FLUC_TTMLBLEND_OK: return "ok";
    case FLUC_TTMLBLEND_ERROR_INVALID_ARGUMENT: return "invalid argument";
    case FLUC_TTMLBLEND_ERROR_NO_DEVICE: return "no usable CUDA device (there is no CPU fallback)";
    case FLUC_TTMLBLEND_ERROR_CUDA: return "CUDA error (context unusable)";
    case FLUC_TTMLBLEND_ERROR_OUT_OF_MEMORY: return "out of memory";
    case FLUC_TTMLBLEND_ERROR_UNSUPPORTED_FORMAT: return "unsupported video format";
    case FLUC_TTMLBLEND_ERROR_NOT_FOUND: return "not found";
    case FLUC_TTMLBLEND_ERROR_TOO_MANY_RECTANGLES: return "too many rectangles";
    default: return "unknown error";
  }
}

int
fluc_ttmlblend_device_count (void)
{
  int n = 0;
  if (cudaGetDeviceCount (&n) != cudaSuccess) {
    cudaGetLastError ();
    return 0;
  }
  return n;
}

int
fluc_ttmlblend_new (int device, FlucTtmlBlend **out)
{
  if (!out)
    return FLUC_TTMLBLEND_ERROR_INVALID_ARGUMENT;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount (&n) != cudaSuccess || n <= 0) {
    cudaGetLastError ();
    return FLUC_TTMLBLEND_ERROR_NO_DEVICE;
  }
  if (device < 0) {
    const char *e = getenv ("FLUC_TTMLBLEND_DEVICE");
    device = e ? atoi (e) : 0;
  }
  if (device >= n)
    return FLUC_TTMLBLEND_ERROR_NO_DEVICE;
  if (cudaSetDevice (device) != cudaSuccess) {
    cudaGetLastError ();
    return FLUC_TTMLBLEND_ERROR_NO_DEVICE;
  }
  FlucTtmlBlend *t = new (std::nothrow) FlucTtmlBlend ();
  if (!t)
    return FLUC_TTMLBLEND_ERROR_OUT_OF_MEMORY;
  Ctx *c = &t->c;
  c->device = device;
  bool ok = true;
  ok &= cudaStreamCreateWithFlags (&c->blend_stream, cudaStreamNonBlocking) == cudaSuccess;
  ok &= cudaStreamCreateWithFlags (&c->up_stream, cudaStreamNonBlocking) == cudaSuccess;
  ok &= cudaStreamCreateWithFlags (&c->reaper, cudaStreamNonBlocking) == cudaSuccess;
  ok &= cudaStreamCreateWithFlags (&c->table_stream, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < kLanes + 2 && ok; i++)
    ok &= cudaEventCreateWithFlags (&c->ev_fence[i], cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; i < kLanes && ok; i++) {
    ok &= cudaStreamCreateWithFlags (&c->lanes[i].stream, cudaStreamNonBlocking) == cudaSuccess;
    ok &= cudaEventCreateWithFlags (&c->lanes[i].done, cudaEventDisableTiming) == cudaSuccess;
  }
  ok &= cudaEventCreate (&c->timer0) == cudaSuccess;
  ok &= cudaEventCreate (&c->timer1) == cudaSuccess;
  if (ok) {
    /* keep freed overlay memory in the pool instead of returning it to the OS */
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool (&pool, device) == cudaSuccess) {
      uint64_t thr = ~0ull;
      cudaMemPoolSetAttribute (pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
  if (!ok) {
    cudaGetLastError ();
    delete t;
    return FLUC_TTMLBLEND_ERROR_NO_DEVICE;
  }
  const char *e;
  if ((e = getenv ("FLUC_TTMLBLEND_BATCH")))
    c->max_batch = (uint32_t) std::max (1, std::min (1024, atoi (e)));
  if ((e = getenv ("FLUC_TTMLBLEND_PROFILE_EVERY")))
    c->profile_every = (uint32_t) std::max (1, atoi (e));
  if ((e = getenv ("FLUC_TTMLBLEND_AUTOCROP")))
    c->autocrop = atoi (e) != 0;
  if ((e = getenv ("FLUC_TTMLBLEND_GROUPS")))
    c->use_groups = atoi (e) != 0;
  if ((e = getenv ("FLUC_TTMLBLEND_HOST_MODE")))
    c->host_mode = std::max (0, std::min (2, atoi (e)));
  if ((e = getenv ("FLUC_TTMLBLEND_LINGER_US")))
    c->linger_us = (uint32_t) std::max (0, atoi (e));
  c->sched = std::thread (scheduler_main, c);
  *out = t;
  TBLOG (1, "context on device %d", device);
  return 0;
}

void
fluc_ttmlblend_free (FlucTtmlBlend *thiz)
{
  if (!thiz)
    return;
  Ctx *c = &thiz->c;
  {
    std::unique_lock<std::mutex> lk (c->mu);
    c->quit = true;
    c->cv.notify_all ();
  }
  if (c->sched.joinable ())
    c->sched.join ();
  cudaSetDevice (c->device);
  cudaDeviceSynchronize ();
  {
    std::unique_lock<std::mutex> lk (c->mu);
    c->pending.clear ();
    for (auto &b : c->batches) {
      if (b.t0) cudaEventDestroy (b.t0);
      if (b.t1) cudaEventDestroy (b.t1);
      cudaEventDestroy (b.done);
    }
    c->batches.clear ();
    for (int i = 0; i < kLanes; i++)
      c->lanes[i].keep.reset ();
    c->overlays.clear ();       /* frees through the reaper stream */
    cudaStreamSynchronize (c->reaper);
    for (auto e : c->event_pool)
      cudaEventDestroy (e);
    for (auto e : c->timing_pool)
      cudaEventDestroy (e);
    auto free_slot = [](TableSlot &s) {
      if (s.h_jobs) cudaFreeHost (s.h_jobs);
      if (s.h_begin) cudaFreeHost (s.h_begin);
      if (s.d_jobs) cudaFree (s.d_jobs);
      if (s.d_begin) cudaFree (s.d_begin);
      if (s.copied) cudaEventDestroy (s.copied);
      if (s.uploaded) cudaEventDestroy (s.uploaded);
    };
    for (auto &s : c->slots)
      free_slot (s);
    for (int i = 0; i < kLanes; i++) {
      free_slot (c->lanes[i].table[0]);
      free_slot (c->lanes[i].table[1]);
      if (c->lanes[i].dev) cudaFree (c->lanes[i].dev);
      if (c->lanes[i].done) cudaEventDestroy (c->lanes[i].done);
      if (c->lanes[i].stream) cudaStreamDestroy (c->lanes[i].stream);
    }
    for (auto &p : c->pool_free) {
      if (p.on_host) cudaFreeHost (p.base); else cudaFree (p.base);
    }
    for (auto &p : c->pool_used) {
      if (p.on_host) cudaFreeHost (p.base); else cudaFree (p.base);
    }
    if (c->scrub) cudaFree (c->scrub);
    for (auto e : c->ev_fence)
      if (e) cudaEventDestroy (e);
    if (c->timer0) cudaEventDestroy (c->timer0);
    if (c->timer1) cudaEventDestroy (c->timer1);
    cudaStreamDestroy (c->blend_stream);
    cudaStreamDestroy (c->up_stream);
    cudaStreamDestroy (c->reaper);
    cudaStreamDestroy (c->table_stream);
  }
  delete thiz;
}

const char *
fluc_ttmlblend_last_cuda_error (FlucTtmlBlend *thiz)
{
  if (!thiz)
    return "";
  std::unique_lock<std::mutex> lk (thiz->c.mu);
  return thiz->c.cuda_error.c_str ();
}

/* ---- overlay --------------------------------------------------------- */

int
fluc_ttmlblend_overlay_set_rectangles (FlucTtmlBlend *thiz, uint32_t stream,
    const FlucTtmlBlendRectangle *rects, uint32_t n_rects)
{
  ENTER (thiz);
  if (n_rects && !rects)
    return FLUC_TTMLBLEND_ERROR_INVALID_ARGUMENT;
  if (n_rects > FLUC_TTMLBLEND_MAX_RECTANGLES)
    return FLUC_TTMLBLEND_ERROR_TOO_MANY_RECTANGLES;
  return overlay_install (c, stream, rects, n_rects);
}

int
fluc_ttmlblend_overlay_set (FlucTtmlBlend *thiz, uint32_t stream, const uint8_t *bgra,
    int32_t w, int32_t h, int32_t stride, const FlucTtmlBlendRect *rects, uint32_t n_rects)
{
  ENTER (thiz);
  if (!bgra || w <= 0 || h <= 0 || stride < 4 * w || (n_rects && !rects))
    return FLUC_TTMLBLEND_ERROR_INVALID_ARGUMENT;
  std::vector<FlucTtmlBlendRect> in;
  if (n_rects == 0) {
    in.push_back ({ 0, 0, w, h });
  } else {
    for (uint32_t i = 0; i < n_rects; i++) {
      /* region boxes clipped to the image ttmlrender drew them into */
      const int x0 = std::max (rects[i].x, 0), y0 = std::max (rects[i].y, 0);
      const int x1 = std::min (rects[i].x + rects[i].w, w), y1 = std::min (rects[i].y + rects[i].h, h);
      if (x1 > x0 && y1 > y0)
        in.push_back ({ x0, y0, x1 - x0, y1 - y0 });
    }
    in = disjoint_cover (in);
  }
  if (in.size () > FLUC_TTMLBLEND_MAX_RECTANGLES)
    return FLUC_TTMLBLEND_ERROR_TOO_MANY_RECTANGLES;
  std::vector<FlucTtmlBlendRectangle> rr;
  for (auto &r : in) {
    FlucTtmlBlendRectangle q;
    q.pixels = bgra + (size_t) r.y * stride + (size_t) r.x * 4;
    q.width = r.w;
    q.height = r.h;
    q.stride = stride;
    q.x = r.x;
    q.y = r.y;
    q.global_alpha = 1.0f;
    q.flags = FLUC_TTMLBLEND_FLAG_PREMULTIPLIED_ALPHA;   /* Cairo ARGB32, gstttmlrender.c:1446 */
    rr.push_back (q);
  }
  return overlay_install (c, stream, rr.data (), (uint32_t) rr.size ());
}

int
fluc_ttmlblend_overlay_clear (FlucTtmlBlend *thiz, uint32_t stream)
{
  ENTER (thiz);
  c->overlays.erase (stream);
  return 0;
}

/* ---- device-resident frames ------------------------------------------ */

static int
submit_locked (Ctx *c, uint32_t stream, int fmt, int32_t W, int32_t H, uint32_t frame_flags,
    const FlucTtmlBlendFrame *src, const FlucTtmlBlendFrame *dst, uint64_t *ticket)
{
  int rc;
  if ((rc = check_frame (fmt, W, H, src)) || (rc = check_frame (fmt, W, H, dst)))
    return rc;
  PendingFrame f;
  f.kind = plane_kind (fmt);
  f.prep = nullptr;
  auto it = c->overlays.find (stream);
  if (it != c->overlays.end ()) {
    f.overlay = it->second;
    if ((rc = prepare_overlay (c, f.overlay.get (), fmt, W, H, &f.prep)))
      return rc;
  }
  /* a buffer written twice in one batch would race: launch what is queued first */
  if (c->pending_dst.count (dst->plane[0]) && (rc = launch_pending (c)))
    return rc;
  f.algo_bytes = build_jobs (fmt, W, H, frame_flags, src, dst, f.prep,
      src->plane[0] == dst->plane[0], f.jobs);
  f.dst0 = dst->plane[0];
  if (c->use_groups)
    make_groupable (f, src, dst);
  f.ticket = ++c->next_ticket;
  if (ticket)
    *ticket = f.ticket;
  if (c->pending.empty ())
    c->oldest_pending = std::chrono::steady_clock::now ();
  c->pending_dst.insert (f.dst0);
  c->pending.push_back (std::move (f));
  if (c->pending.size () >= c->max_batch)
    return launch_pending (c);
  return 0;
}

int
fluc_ttmlblend_submit (FlucTtmlBlend *thiz, uint32_t stream, FlucTtmlBlendFormat fmt,
    int32_t W, int32_t H, uint32_t frame_flags, const FlucTtmlBlendFrame *src,
    const FlucTtmlBlendFrame *dst, uint64_t *ticket)
{
  ENTER (thiz);
  int rc = submit_locked (c, stream, fmt, W, H, frame_flags, src, dst, ticket);
  if (rc == 0 && c->linger_us && !c->pending.empty ())
    c->cv.notify_all ();
  return rc;
}

int
fluc_ttmlblend_submit_many (FlucTtmlBlend *thiz, uint32_t n, const uint32_t *streams,
    FlucTtmlBlendFormat fmt, int32_t W, int32_t H, uint32_t frame_flags,
    const FlucTtmlBlendFrame *srcs, const FlucTtmlBlendFrame *dsts, uint64_t *tickets)
{
  ENTER (thiz);
  if (n && (!streams || !srcs || !dsts))
    return FLUC_TTMLBLEND_ERROR_INVALID_ARGUMENT;
  for (uint32_t i = 0; i < n; i++) {
    int rc = submit_locked (c, streams[i], fmt, W, H, frame_flags, &srcs[i], &dsts[i],
        tickets ? &tickets[i] : nullptr);
    if (rc)
      return rc;
  }
  if (c->linger_us && !c->pending.empty ())
    c->cv.notify_all ();
  return 0;
}

int
fluc_ttmlblend_flush (FlucTtmlBlend *thiz)
{
  ENTER (thiz);
  return launch_pending (c);
}

int
fluc_ttmlblend_wait (FlucTtmlBlend *thiz, uint64_t ticket)
{
  ENTER (thiz);
  if (ticket == 0 || ticket > c->next_ticket)
    return FLUC_TTMLBLEND_ERROR_NOT_FOUND;
  auto lt = c->lane_tickets.find (ticket);
  if (lt != c->lane_tickets.end ()) {
    Lane &l = c->lanes[lt->second];
    cudaEvent_t ev = l.done;
    lk.unlock ();
    cudaError_t e = cudaEventSynchronize (ev);
    lk.lock ();
    if (e != cudaSuccess) {
      c->sticky = FLUC_TTMLBLEND_ERROR_CUDA;
      c->cuda_error = std::string ("wait: ") + cudaGetErrorString (e);
      return c->sticky;
    }
    auto again = c->lane_tickets.find (ticket);
    if (again != c->lane_tickets.end ()) {
      Lane &l2 = c->lanes[again->second];
      if (l2.ticket == ticket) {
        l2.busy = false;
        l2.keep.reset ();
      }
      c->lane_tickets.erase (again);
    }
    return 0;
  }
  if (!c->pending.empty () && ticket >= c->pending.front ().ticket) {
    int rc = launch_pending (c);
    if (rc)
      return rc;
  }
  cudaEvent_t ev = nullptr;
  for (auto &b : c->batches)
    if (b.last_ticket >= ticket) {
      ev = b.done;
      break;
    }
  if (!ev) {
    return 0;                   /* already reaped: finished */
  }
  /* the event stays valid while we wait: batches are only reaped under mu,
   * and a reaped event goes back to the pool, not destroyed */
  lk.unlock ();
  cudaError_t e = cudaEventSynchronize (ev);
  lk.lock ();
  if (e != cudaSuccess) {
    c->sticky = FLUC_TTMLBLEND_ERROR_CUDA;
    c->cuda_error = std::string ("wait: ") + cudaGetErrorString (e);
    return c->sticky;
  }
  reap_batches (c);
  return 0;
}

int
fluc_ttmlblend_sync (FlucTtmlBlend *thiz)
{
  ENTER (thiz);
  int rc = launch_pending (c);
  if (rc)
    return rc;
  CU (c, cudaStreamSynchronize (c->blend_stream));
  for (int i = 0; i < kLanes; i++) {
    CU (c, cudaStreamSynchronize (c->lanes[i].stream));
    c->lanes[i].busy = false;
    c->lanes[i].keep.reset ();
  }
  c->lane_tickets.clear ();
  CU (c, cudaStreamSynchronize (c->up_stream));
  reap_batches (c);
  return 0;
}

int
fluc_ttmlblend_set_batch (FlucTtmlBlend *thiz, uint32_t max_frames, uint32_t linger_us)
{
  ENTER (thiz);
  if (max_frames < 1 || max_frames > 1024)
    return FLUC_TTMLBLEND_ERROR_INVALID_ARGUMENT;
  c->max_batch = max_frames;
  c->linger_us = linger_us;
  c->cv.notify_all ();
  return 0;
}

/* ---- host-resident frames -------------------------------------------- */

int
fluc_ttmlblend_blend_host (FlucTtmlBlend *thiz, uint32_t stream, FlucTtmlBlendFormat fmt,
    int32_t W, int32_t H, uint32_t frame_flags, const FlucTtmlBlendFrame *hf, uint64_t *ticket)
{
  ENTER (thiz);
  int rc;
  if ((rc = check_frame (fmt, W, H, hf)))
    return rc;
  const uint64_t tk = ++c->next_ticket;
  if (ticket)
    *ticket = tk;
  auto it = c->overlays.find (stream);
  if (it == c->overlays.end ())
    return 0;                   /* no overlay: the frame passes through untouched */
  std::shared_ptr<Overlay> ov = it->second;
  Prepared *prep = nullptr;
  if ((rc = prepare_overlay (c, ov.get (), fmt, W, H, &prep)))
    return rc;

  /* Is the host frame device-accessible (pool frame / host_register)? Then the
   * kernel can reach it over PCIe itself. */
  const int n_planes = format_planes (fmt);
  FlucTtmlBlendFrame zf = {};
  bool mapped = c->host_mode != HM_STAGED;
  for (int pl = 0; pl < n_planes && mapped; pl++) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes (&attr, hf->plane[pl]) != cudaSuccess ||
        attr.type != cudaMemoryTypeHost || !attr.devicePointer) {
      cudaGetLastError ();
      mapped = false;
    } else {
      zf.plane[pl] = attr.devicePointer;
      zf.stride[pl] = hf->stride[pl];
      /* byte-granular accesses to host memory would each cross PCIe: frames that are not
       * 16-byte aligned go through the staging lanes (DMA, then the vector kernel) instead */
      if ((((uintptr_t) zf.plane[pl] | (uintptr_t) zf.stride[pl]) & 15u) != 0)
        mapped = false;
    }
  }

  if (mapped && c->host_mode == HM_ZEROCOPY) {
    /* zero copy: the frame joins the batch; the blend kernel reads the rows
     * under the cue from host memory and writes them back, all over PCIe,
     * one launch for every queued frame */
    PendingFrame f;
    f.kind = plane_kind (fmt);
    f.overlay = ov;
    f.prep = prep;
    f.ticket = tk;
    if (c->pending_dst.count (zf.plane[0]) && (rc = launch_pending (c)))
      return rc;
    f.algo_bytes = build_jobs (fmt, W, H, frame_flags, &zf, &zf, prep, true, f.jobs);
    f.dst0 = zf.plane[0];
    c->pending_dst.insert (f.dst0);
    for (const PlaneJob &j : f.jobs) {
      const uint64_t nb = (uint64_t) std::min (j.win_nv * 16, j.row_bytes - j.win_v0 * 16) * j.win_rows;
      c->stats.h2d_bytes += nb;
      c->stats.d2h_bytes += nb;
    }
    if (c->use_groups)
      make_groupable (f, &zf, &zf);
    if (c->pending.empty ())
      c->oldest_pending = std::chrono::steady_clock::now ();
    c->pending.push_back (std::move (f));
    if (c->pending.size () >= c->max_batch)
      return launch_pending (c);
    if (c->linger_us)
      c->cv.notify_all ();
    return 0;
  }

  Lane &l = c->lanes[c->next_lane];
  const int lane_idx = c->next_lane;
  c->next_lane = (c->next_lane + 1) % kLanes;
  if (l.busy) {
    CU (c, cudaEventSynchronize (l.done));
    c->lane_tickets.erase (l.ticket);
    l.busy = false;
    l.keep.reset ();
  }

  /* device staging frame: same strides as a pool frame */
  FlucTtmlBlendFrame df = {};
  size_t off = 0;
  size_t plane_off[3];
  for (int pl = 0; pl < n_planes; pl++) {
    df.stride[pl] = (int32_t) align_up ((size_t) plane_row_bytes (fmt, pl, W), 256);
    plane_off[pl] = off;
    off += (size_t) df.stride[pl] * plane_rows (fmt, pl, H);
  }
  if ((rc = lane_reserve (c, l, off)))
    return rc;
  for (int pl = 0; pl < n_planes; pl++)
    df.plane[pl] = l.dev + plane_off[pl];

  /* HM_WRITEBACK: the copy engine brings the rows in, the kernel stores the
   * result straight into the host frame (posted PCIe writes), no copy back */
  const bool writeback = mapped && c->host_mode == HM_WRITEBACK;
  std::vector<PlaneJob> jobs;
  const uint64_t algo = build_jobs (fmt, W, H, frame_flags, &df, writeback ? &zf : &df, prep, true, jobs);
  if (jobs.empty ())
    return 0;
  CU (c, cudaStreamWaitEvent (l.stream, prep->ready, 0));

  /* rows each window touches: host -> device. Bands that are neighbours in y
   * with the same byte range merge into one copy. */
  struct Span { int pl, b0, nb, y0, rows; };
  std::vector<Span> spans;
  for (const PlaneJob &j : jobs) {
    Span s;
    s.pl = j.plane;
    s.b0 = j.win_v0 * 16;
    s.nb = std::min (j.win_nv * 16, j.row_bytes - s.b0);
    s.y0 = j.win_y0;
    s.rows = j.win_rows;
    bool merged = false;
    for (Span &o : spans)
      if (o.pl == s.pl && o.b0 == s.b0 && o.nb == s.nb && o.y0 + o.rows == s.y0) {
        o.rows += s.rows;
        merged = true;
        break;
      }
    if (!merged)
      spans.push_back (s);
  }
  auto copy_span = [&](const Span &s, bool to_device) -> cudaError_t {
    uint8_t *d = static_cast<uint8_t *> (df.plane[s.pl]) + (size_t) s.y0 * df.stride[s.pl] + s.b0;
    uint8_t *h = static_cast<uint8_t *> (hf->plane[s.pl]) + (size_t) s.y0 * hf->stride[s.pl] + s.b0;
    if (s.nb == df.stride[s.pl] && s.nb == hf->stride[s.pl]) {
      /* full rows, equal strides: one linear copy */
      return to_device ?
          cudaMemcpyAsync (d, h, (size_t) s.nb * s.rows, cudaMemcpyHostToDevice, l.stream) :
          cudaMemcpyAsync (h, d, (size_t) s.nb * s.rows, cudaMemcpyDeviceToHost, l.stream);
    }
    return to_device ?
        cudaMemcpy2DAsync (d, df.stride[s.pl], h, hf->stride[s.pl], s.nb, s.rows,
            cudaMemcpyHostToDevice, l.stream) :
        cudaMemcpy2DAsync (h, hf->stride[s.pl], d, df.stride[s.pl], s.nb, s.rows,
            cudaMemcpyDeviceToHost, l.stream);
  };
  for (const Span &s : spans) {
    CU (c, copy_span (s, true));
    c->stats.h2d_bytes += (uint64_t) s.nb * s.rows;
  }
  {
    std::vector<PlaneJob> grp[2];
    for (const PlaneJob &j : jobs)
      grp[(j.flags & JF_FAST) ? 1 : 0].push_back (j);
    for (int g = 0; g < 2; g++)
      if (!grp[g].empty ()) {
        if ((rc = launch_jobs (c, l.table[g], grp[g].data (), grp[g].size (), plane_kind (fmt),
                    g == 1, l.stream)))
          return rc;
      }
  }
  for (const Span &s : spans) {
    if (!writeback)
      CU (c, copy_span (s, false));
    c->stats.d2h_bytes += (uint64_t) s.nb * s.rows;
  }
  CU (c, cudaEventRecord (l.done, l.stream));
  l.busy = true;
  l.ticket = tk;
  l.keep = ov;
  c->lane_tickets[tk] = lane_idx;
  c->stats.frames_blended++;
  c->stats.algorithmic_bytes += algo;
  return 0;
}

int
fluc_ttmlblend_host_register (FlucTtmlBlend *thiz, void *ptr, size_t bytes)
{
  ENTER (thiz);
  if (!ptr || !bytes)
    return FLUC_TTMLBLEND_ERROR_INVALID_ARGUMENT;
  cudaError_t e = cudaHostRegister (ptr, bytes, cudaHostRegisterDefault);
  if (e != cudaSuccess) {
    cudaGetLastError ();
    return FLUC_TTMLBLEND_ERROR_INVALID_ARGUMENT;
  }
  return 0;
}

int
fluc_ttmlblend_host_unregister (FlucTtmlBlend *thiz, void *ptr)
{
  ENTER (thiz);
  cudaError_t e = cudaHostUnregister (ptr);
  if (e != cudaSuccess) {
    cudaGetLastError ();
    return FLUC_TTMLBLEND_ERROR_NOT_FOUND;
  }
  return 0;
}

/* ---- frame pool ------------------------------------------------------ */

int
fluc_ttmlblend_format_planes (FlucTtmlBlendFormat fmt)
{
  return format_valid (fmt) ? format_planes (fmt) : 0;
}

int
fluc_ttmlblend_plane_row_bytes (FlucTtmlBlendFormat fmt, int plane, int32_t width)
{
  if (!format_valid (fmt) || plane < 0 || plane >= format_planes (fmt))
    return 0;
  return plane_row_bytes (fmt, plane, width);
}

int
fluc_ttmlblend_plane_rows (FlucTtmlBlendFormat fmt, int plane, int32_t height)
{
  if (!format_valid (fmt) || plane < 0 || plane >= format_planes (fmt))
    return 0;
  return plane_rows (fmt, plane, height);
}

int
fluc_ttmlblend_frame_pool_acquire (FlucTtmlBlend *thiz, FlucTtmlBlendFormat fmt, int32_t W,
    int32_t H, int on_host, FlucTtmlBlendFrame *out)
{
  ENTER (thiz);
  if (!out || W <= 0 || H <= 0)
    return FLUC_TTMLBLEND_ERROR_INVALID_ARGUMENT;
  if (!format_valid (fmt))
    return FLUC_TTMLBLEND_ERROR_UNSUPPORTED_FORMAT;
  on_host = on_host ? 1 : 0;
  for (size_t i = 0; i < c->pool_free.size (); i++) {
    PoolEntry &p = c->pool_free[i];
    if (p.fmt == fmt && p.W == W && p.H == H && p.on_host == on_host) {
      *out = p.frame;
      c->pool_used.push_back (p);
      c->pool_free.erase (c->pool_free.begin () + i);
      return 0;
    }
  }
  PoolEntry p = {};
  p.fmt = fmt; p.W = W; p.H = H; p.on_host = on_host;
  size_t off = 0, plane_off[3] = { 0, 0, 0 };
  const int n_planes = format_planes (fmt);
  for (int pl = 0; pl < n_planes; pl++) {
    p.frame.stride[pl] = (int32_t) align_up ((size_t) plane_row_bytes (fmt, pl, W), 256);
    plane_off[pl] = off;
    off += (size_t) p.frame.stride[pl] * plane_rows (fmt, pl, H);
  }
  p.bytes = off;
  if (on_host)
    CU (c, cudaHostAlloc (&p.base, off, cudaHostAllocDefault));
  else
    CU (c, cudaMalloc (&p.base, off));
  for (int pl = 0; pl < n_planes; pl++)
    p.frame.plane[pl] = static_cast<uint8_t *> (p.base) + plane_off[pl];
  *out = p.frame;
  c->pool_used.push_back (p);
  return 0;
}

int
fluc_ttmlblend_frame_pool_release (FlucTtmlBlend *thiz, const FlucTtmlBlendFrame *frame)
{
  ENTER (thiz);
  if (!frame)
    return FLUC_TTMLBLEND_ERROR_INVALID_ARGUMENT;
  for (size_t i = 0; i < c->pool_used.size (); i++)
    if (c->pool_used[i].frame.plane[0] == frame->plane[0]) {
      c->pool_free.push_back (c->pool_used[i]);
      c->pool_used.erase (c->pool_used.begin () + i);
      return 0;
    }
  return FLUC_TTMLBLEND_ERROR_NOT_FOUND;
}

static int
frame_copy (Ctx *c, int fmt, int W, int H, const FlucTtmlBlendFrame *s, const FlucTtmlBlendFrame *d,
    cudaMemcpyKind kind)
{
  int rc;
  if ((rc = check_frame (fmt, W, H, s)) || (rc = check_frame (fmt, W, H, d)))
    return rc;
  for (int pl = 0; pl < format_planes (fmt); pl++) {
    const size_t rb = (size_t) plane_row_bytes (fmt, pl, W);
    CU (c, cudaMemcpy2DAsync (d->plane[pl], d->stride[pl], s->plane[pl], s->stride[pl], rb,
            plane_rows (fmt, pl, H), kind, c->blend_stream));
    if (kind == cudaMemcpyHostToDevice)
      c->stats.h2d_bytes += rb * plane_rows (fmt, pl, H);
    else
      c->stats.d2h_bytes += rb * plane_rows (fmt, pl, H);
  }
  CU (c, cudaStreamSynchronize (c->blend_stream));
  return 0;
}

int
fluc_ttmlblend_frame_upload (FlucTtmlBlend *thiz, FlucTtmlBlendFormat fmt, int32_t W, int32_t H,
    const FlucTtmlBlendFrame *host_src, const FlucTtmlBlendFrame *dev_dst)
{
  ENTER (thiz);
  return frame_copy (c, fmt, W, H, host_src, dev_dst, cudaMemcpyHostToDevice);
}

int
fluc_ttmlblend_frame_download (FlucTtmlBlend *thiz, FlucTtmlBlendFormat fmt, int32_t W, int32_t H,
    const FlucTtmlBlendFrame *dev_src, const FlucTtmlBlendFrame *host_dst)
{
  ENTER (thiz);
  int rc = launch_pending (c);
  if (rc)
    return rc;
  return frame_copy (c, fmt, W, H, dev_src, host_dst, cudaMemcpyDeviceToHost);
}

/* ---- outline blur (per cue, producer side) --------------------------- */

int
fluc_ttmlblend_blur_argb32 (FlucTtmlBlend *thiz, const uint8_t *src, int32_t w, int32_t h,
    int32_t stride, int32_t radius, double sigma, uint8_t *dst, int32_t dst_stride)
{
  ENTER (thiz);
  if (!src || !dst || w <= 0 || h <= 0 || stride < 4 * w || dst_stride < 4 * w || radius < 0 ||
      radius > 64 || !(sigma > 0.0))
    return FLUC_TTMLBLEND_ERROR_INVALID_ARGUMENT;
  /* gst_ttml_blur_create_gaussian_kernel, /root/reference/plugins/ttml/gstttmlblur.c:28-67:
   * G(x,y) = exp (-(x^2 + y^2) / (2 sigma^2)) / (2 pi sigma^2), normalised, 16.16 fixed */
  const int size = 2 * radius + 1, n = size * size;
  std::vector<double> tmp (n);
  std::vector<int32_t> taps (n);
  const double scale2 = 2.0 * sigma * sigma;
  const double scale1 = 1.0 / (3.14159265358979323846 * scale2);
  double sum = 0;
  int i = 0;
  for (int x = -radius; x <= radius; ++x)
    for (int y = -radius; y <= radius; ++y, ++i) {
      const double u = x * x, v = y * y;
      tmp[i] = scale1 * exp (-(u + v) / scale2);
      sum += tmp[i];
    }
  for (i = 0; i < n; ++i)
    taps[i] = (int32_t) ((tmp[i] / sum) * 65536.0);       /* pixman_double_to_fixed */

  const size_t pitch = align_up ((size_t) w * 4, 256);
  void *d_src = nullptr, *d_dst = nullptr, *d_taps = nullptr;
  CU (c, cudaMallocAsync (&d_src, pitch * h, c->up_stream));
  CU (c, cudaMallocAsync (&d_dst, pitch * h, c->up_stream));
  CU (c, cudaMallocAsync (&d_taps, (size_t) n * 4, c->up_stream));
  CU (c, cudaMemcpy2DAsync (d_src, pitch, src, stride, (size_t) w * 4, h, cudaMemcpyHostToDevice, c->up_stream));
  CU (c, cudaMemcpyAsync (d_taps, taps.data (), (size_t) n * 4, cudaMemcpyHostToDevice, c->up_stream));
  CU (c, launch_blur (static_cast<uint8_t *> (d_src), w, h, (int) pitch, static_cast<int32_t *> (d_taps),
          radius, static_cast<uint8_t *> (d_dst), (int) pitch, c->up_stream));
  CU (c, cudaMemcpy2DAsync (dst, dst_stride, d_dst, pitch, (size_t) w * 4, h, cudaMemcpyDeviceToHost, c->up_stream));
  CU (c, cudaFreeAsync (d_src, c->up_stream));
  CU (c, cudaFreeAsync (d_dst, c->up_stream));
  CU (c, cudaFreeAsync (d_taps, c->up_stream));
  CU (c, cudaStreamSynchronize (c->up_stream));
  c->stats.h2d_bytes += (uint64_t) w * 4 * h;
  c->stats.d2h_bytes += (uint64_t) w * 4 * h;
  return 0;
}

/* ---- observability --------------------------------------------------- */

void
fluc_ttmlblend_stats_copy (FlucTtmlBlend *thiz, FlucTtmlBlendStats *out)
{
  if (!thiz || !out)
    return;
  std::unique_lock<std::mutex> lk (thiz->c.mu);
  cudaSetDevice (thiz->c.device);
  reap_batches (&thiz->c);
  *out = thiz->c.stats;
}

void
fluc_ttmlblend_stats_reset (FlucTtmlBlend *thiz)
{
  if (!thiz)
    return;
  std::unique_lock<std::mutex> lk (thiz->c.mu);
  cudaSetDevice (thiz->c.device);
  reap_batches (&thiz->c);
  thiz->c.stats = FlucTtmlBlendStats ();
}

int
fluc_ttmlblend_set_profiling (FlucTtmlBlend *thiz, int enabled)
{
  ENTER (thiz);
  c->profiling = enabled != 0;
  if (enabled > 0)
    c->profile_every = (uint32_t) enabled;
  c->profile_seq = 0;
  return 0;
}

int
fluc_ttmlblend_timer_begin (FlucTtmlBlend *thiz)
{
  ENTER (thiz);
  int rc = launch_pending (c);
  if (rc)
    return rc;
  CU (c, cudaEventRecord (c->timer0, c->blend_stream));
  return 0;
}

int
fluc_ttmlblend_timer_end (FlucTtmlBlend *thiz, double *ms)
{
  ENTER (thiz);
  int rc = launch_pending (c);
  if (rc)
    return rc;
  CU (c, cudaEventRecord (c->timer1, c->blend_stream));
  CU (c, cudaEventSynchronize (c->timer1));
  float f = 0.f;
  CU (c, cudaEventElapsedTime (&f, c->timer0, c->timer1));
  if (ms)
    *ms = f;
  reap_batches (c);
  return 0;
}

int
fluc_ttmlblend_scrub_l2 (FlucTtmlBlend *thiz, size_t bytes)
{
  ENTER (thiz);
  if (bytes > c->scrub_bytes) {
    if (c->scrub)
      CU (c, cudaFree (c->scrub));
    c->scrub = nullptr;
    c->scrub_bytes = 0;
    CU (c, cudaMalloc ((void **) &c->scrub, bytes));
    c->scrub_bytes = bytes;
  }
  CU (c, launch_scrub (c->scrub, bytes, c->blend_stream));
  return 0;
}

void *
fluc_ttmlblend_stream_handle (FlucTtmlBlend *thiz)
{
  return thiz ? (void *) thiz->c.blend_stream : nullptr;
}

}  /* extern "C" */
