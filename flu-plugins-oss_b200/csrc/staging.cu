/*
 * staging.cu -- host frames the GPU cannot reach (ordinary pageable memory, which is what an
 * arbitrary upstream GstBuffer is until the pinned allocator has been negotiated, or frames that
 * are not 16-byte aligned) on their way through the zero-copy path.
 *
 * gst_video_overlay_composition_blend (comp, frame) touches only the rows under the cue, and so
 * does this: a pool of worker threads copies those rows into a pinned staging frame (memcpy runs
 * at memory speed on every core; a cudaMemcpy from pageable memory is staged by the driver on one
 * thread), the staging frame joins the batch like any pinned frame -- the blend kernel reads and
 * rewrites it over PCIe -- and once its batch has finished a worker copies the rows back into
 * the caller's frame. Copy-in of the next frames, the blend of the current ones and copy-out of
 * the previous ones overlap. Nothing is registered or pinned behind the caller's back.
 *
 *   blend_host ()  --->  stage_in  --workers: copy in-->  pending batch  --GPU-->  stage_gpu
 *        wait (ticket)  <--- DONE <--workers: copy out--  stage_out  <--completer: batch event--
 *
 * Shape after the reference's worker-thread-under-a-monitor
 * (/root/reference/libs/flu/downloader/lib/fludownloader.c:490-532).
 */
#include "ttmlblend_internal.h"

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace tbh {

/* memcpy whose stores bypass the cache. Neither side of a staging copy is read again by this
 * core: the staging frame is read by the GPU over PCIe, the caller's frame by whoever comes next
 * in the pipeline, much later. Ordinary stores would first fetch every destination line
 * (read-for-ownership): three memory transfers per byte instead of two, on a path that is
 * bound by host memory bandwidth once a few threads copy at the same time. */
#if defined(__x86_64__)
__attribute__ ((target ("avx2"))) static void
copy_stream_avx2 (uint8_t *d, const uint8_t *s, size_t n)
{
  const size_t head = (32 - ((uintptr_t) d & 31)) & 31;
  if (head >= n) {
    memcpy (d, s, n);
    return;
  }
  memcpy (d, s, head);
  d += head;
  s += head;
  n -= head;
  size_t i = 0;
  for (; i + 128 <= n; i += 128) {
    const __m256i a = _mm256_loadu_si256 ((const __m256i *) (s + i));
    const __m256i b = _mm256_loadu_si256 ((const __m256i *) (s + i + 32));
    const __m256i c = _mm256_loadu_si256 ((const __m256i *) (s + i + 64));
    const __m256i e = _mm256_loadu_si256 ((const __m256i *) (s + i + 96));
    _mm256_stream_si256 ((__m256i *) (d + i), a);
    _mm256_stream_si256 ((__m256i *) (d + i + 32), b);
    _mm256_stream_si256 ((__m256i *) (d + i + 64), c);
    _mm256_stream_si256 ((__m256i *) (d + i + 96), e);
  }
  for (; i + 32 <= n; i += 32)
    _mm256_stream_si256 ((__m256i *) (d + i), _mm256_loadu_si256 ((const __m256i *) (s + i)));
  _mm_sfence ();
  memcpy (d + i, s + i, n - i);
}
#endif

static void
copy_stream (uint8_t *d, const uint8_t *s, size_t n)
{
#if defined(__x86_64__)
  static const bool avx2 = __builtin_cpu_supports ("avx2") &&
      !(getenv ("FLUC_TTMLBLEND_STAGE_NT") && atoi (getenv ("FLUC_TTMLBLEND_STAGE_NT")) == 0);
  if (avx2 && n >= 4096) {
    copy_stream_avx2 (d, s, n);
    return;
  }
#endif
  memcpy (d, s, n);
}

static void
copy_spans (const StageJob &job, bool in)
{
  for (const StageSpan &s : job.spans) {
    const int us = job.user.stride[s.plane], ss = job.slot.frame.stride[s.plane];
    uint8_t *u = static_cast<uint8_t *> (job.user.plane[s.plane]) + (size_t) s.y0 * us + s.b0;
    uint8_t *p = static_cast<uint8_t *> (job.slot.frame.plane[s.plane]) + (size_t) s.y0 * ss + s.b0;
    if (us == ss && s.nb == us) {
      /* full rows, equal strides: one run */
      if (in)
        copy_stream (p, u, (size_t) s.nb * s.rows);
      else
        copy_stream (u, p, (size_t) s.nb * s.rows);
      continue;
    }
    for (int r = 0; r < s.rows; r++) {
      if (in)
        memcpy (p + (size_t) r * ss, u + (size_t) r * us, (size_t) s.nb);
      else
        memcpy (u + (size_t) r * us, p + (size_t) r * ss, (size_t) s.nb);
    }
  }
}

/* a pinned staging frame of this geometry: a recycled one or a new one (context locked) */
static int
slot_acquire (Ctx *c, int fmt, int W, int H, PoolEntry *out)
{
  for (size_t i = 0; i < c->stage_slots_free.size (); i++) {
    PoolEntry &p = c->stage_slots_free[i];
    if (p.fmt == fmt && p.W == W && p.H == H) {
      *out = p;
      c->stage_slots_free.erase (c->stage_slots_free.begin () + i);
      return 0;
    }
  }
  /* other geometries' slots give way when the pool is at its limit */
  while (c->stage_slots_total >= c->stage_slots_max && !c->stage_slots_free.empty ()) {
    PoolEntry &old = c->stage_slots_free.back ();
    for (int pl = 0; pl < 3; pl++)
      if (old.frame.plane[pl])
        c->pinned_planes.erase (old.frame.plane[pl]);
    /* a slot is only ever free once its batch has finished and its rows have been copied out */
    cudaFreeHost (old.base);
    c->stage_slots_free.pop_back ();
    c->stage_slots_total--;
  }
  PoolEntry p = {};
  p.fmt = fmt; p.W = W; p.H = H; p.on_host = 1;
  size_t off = 0, plane_off[3] = { 0, 0, 0 };
  const int n_planes = format_planes (fmt);
  for (int pl = 0; pl < n_planes; pl++) {
    p.frame.stride[pl] = (int32_t) align_up ((size_t) plane_row_bytes (fmt, pl, W), 256);
    plane_off[pl] = off;
    off += (size_t) p.frame.stride[pl] * plane_rows (fmt, pl, H);
  }
  p.bytes = off;
  CU (c, cudaHostAlloc (&p.base, off, cudaHostAllocDefault));
  for (int pl = 0; pl < n_planes; pl++) {
    p.frame.plane[pl] = static_cast<uint8_t *> (p.base) + plane_off[pl];
    c->pinned_planes.insert (p.frame.plane[pl]);
  }
  c->stage_slots_total++;
  *out = p;
  return 0;
}

static void
job_finish (Ctx *c, const std::shared_ptr<StageJob> &job, int rc)
{
  job->rc = rc;
  job->state = StageJob::DONE;
  job->ov.reset ();
  c->stage_slots_free.push_back (job->slot);
  c->stage_active--;
  c->stage_done_cv.notify_all ();
}

/* Worker: copy-outs first (they free staging frames and complete tickets), then copy-ins. */
static void
stage_worker_main (Ctx *c)
{
  cudaSetDevice (c->device);
  std::unique_lock<std::mutex> lk (c->mu);
  while (!c->quit) {
    if (!c->stage_out.empty ()) {
      std::shared_ptr<StageJob> job = c->stage_out.front ();
      c->stage_out.pop_front ();
      lk.unlock ();
      copy_spans (*job, false);
      lk.lock ();
      job_finish (c, job, 0);
      continue;
    }
    if (!c->stage_in.empty ()) {
      std::shared_ptr<StageJob> job = c->stage_in.front ();
      c->stage_in.pop_front ();
      c->stage_copying++;
      lk.unlock ();
      copy_spans (*job, true);
      lk.lock ();
      c->stage_copying--;
      int rc = c->sticky;
      if (!rc) {
        /* the staging frame joins the batch under a ticket of its own: tickets are handed out
         * in the order frames are queued, which for staged frames is now, not at blend_host */
        job->gpu_ticket = ++c->next_ticket;
        rc = queue_mapped_frame (c, job->gpu_ticket, job->stream, job->ov, job->prep, job->fmt, job->W, job->H,
            job->frame_flags, &job->slot.frame, &job->slot.frame);
      }
      if (rc) {
        job_finish (c, job, rc);
        continue;
      }
      job->state = StageJob::ON_GPU;
      c->stage_gpu.push_back (job);
      /* nobody else is about to join the batch: launch it now instead of leaving it to the
       * linger timer */
      if (c->stage_in.empty () && c->stage_copying == 0) {
        if (!c->pending.empty ())
          launch_pending (c);
        c->stage_done_cv.notify_all ();   /* sync () waits for "every copy-in has been queued" */
      } else if (c->pending.size () >= (size_t) std::max (4, c->stage_threads)) {
        /* a wave of copy-ins is in: off it goes, so that its trip over PCIe overlaps the
         * copy-in of the next wave instead of waiting for the whole burst */
        launch_pending (c);
      }
      c->stage_gpu_cv.notify_all ();
      continue;
    }
    c->stage_cv.wait (lk);
  }
}

/* Completer: waits for the batch of the oldest staged frame on the GPU, then hands every
 * staged frame of that batch (and of earlier ones) to the workers for the copy back. */
static void
stage_completer_main (Ctx *c)
{
  cudaSetDevice (c->device);
  std::unique_lock<std::mutex> lk (c->mu);
  while (!c->quit) {
    if (c->stage_gpu.empty ()) {
      c->stage_gpu_cv.wait (lk);
      continue;
    }
    const uint64_t want = c->stage_gpu.front ()->gpu_ticket;
    if (!c->pending.empty () && want >= c->pending.front ().ticket) {
      /* still queued: the worker that finishes the last copy-in of the burst launches the batch
       * (or it fills up, or the linger timer fires); look again shortly */
      if (c->stage_in.empty () && c->stage_copying == 0)
        launch_pending (c);
      else
        c->stage_gpu_cv.wait_for (lk, std::chrono::microseconds (100));
      continue;
    }
    cudaEvent_t ev = nullptr;
    uint64_t upto = want;
    if (want > c->retired_through)
      for (auto &b : c->batches)
        if (b.last_ticket >= want) {
          ev = b.done;
          upto = b.last_ticket;
          break;
        }
    int rc = 0;
    if (ev) {
      /* the event stays valid while we wait (reaped events go back to the pool, and a pooled
       * event that is recorded again only makes us wait a little longer) */
      lk.unlock ();
      const cudaError_t e = cudaEventSynchronize (ev);
      lk.lock ();
      if (e != cudaSuccess) {
        c->sticky = FLUC_TTMLBLEND_ERROR_CUDA;
        c->cuda_error = std::string ("staging: ") + cudaGetErrorString (e);
        rc = c->sticky;
      }
    }
    for (const Ctx::FailedRange &fr : c->failed_ranges)
      if (want >= fr.first && want <= fr.last)
        rc = rc ? rc : fr.rc;
    while (!c->stage_gpu.empty () && c->stage_gpu.front ()->gpu_ticket <= upto) {
      std::shared_ptr<StageJob> job = c->stage_gpu.front ();
      c->stage_gpu.pop_front ();
      if (rc) {
        job_finish (c, job, rc);
      } else {
        job->state = StageJob::COPY_OUT;
        c->stage_out.push_back (job);
      }
    }
    c->stage_cv.notify_all ();
  }
}

static void
stage_start_threads (Ctx *c)
{
  if (!c->stage_workers.empty ())
    return;
  for (int i = 0; i < c->stage_threads; i++)
    c->stage_workers.emplace_back (stage_worker_main, c);
  c->stage_completer = std::thread (stage_completer_main, c);
}

/* blend_host () of a frame the GPU cannot reach. Context locked (lk); may block -- unlocked --
 * while every staging frame is in use or the same host buffer is still on its way. */
int
stage_frame (Ctx *c, std::unique_lock<std::mutex> &lk, uint64_t tk, uint32_t stream, const std::shared_ptr<Overlay> &ov,
    Prepared *prep, int fmt, int W, int H, uint32_t frame_flags, const FlucTtmlBlendFrame *hf)
{
  stage_start_threads (c);
  const FrameExtent xh (fmt, W, H, hf);
  const uintptr_t lo = xh.hull_lo (), hi = xh.hull_hi ();
  /* the same buffer twice (two streams' cues on one frame): the second copy-in must see the
   * first one's result, so it waits for it */
  for (;;) {
    bool busy = false;
    for (auto &kv : c->stage_jobs)
      if (kv.second->state != StageJob::DONE && lo < kv.second->user_hi && kv.second->user_lo < hi)
        busy = true;
    if (!busy || c->quit || c->sticky)
      break;
    c->stage_done_cv.wait (lk);
  }
  /* ... or still queued / running as a zero-copy frame (it was device-accessible a moment ago) */
  if (xh.hits (c->pending_dst) || xh.hits (c->inflight_host)) {
    int rc = launch_pending (c);
    if (rc)
      return rc;
    if (!c->batches.empty ()) {
      cudaEvent_t ev = c->batches.back ().done;
      lk.unlock ();
      cudaEventSynchronize (ev);
      lk.lock ();
    }
  }
  while (c->stage_active >= c->stage_slots_max && !c->quit && !c->sticky)
    c->stage_done_cv.wait (lk);
  if (c->sticky)
    return c->sticky;
  if (c->quit)
    return FLUC_TTMLBLEND_ERROR_INVALID_ARGUMENT;
  /* the overlay may have been replaced while we waited: the frame keeps the cue it was
   * submitted under (ov / prep are held by the job) */
  std::shared_ptr<StageJob> job (new StageJob ());
  int rc = slot_acquire (c, fmt, W, H, &job->slot);
  if (rc)
    return rc;
  job->ticket = tk;
  job->stream = stream;
  job->fmt = fmt;
  job->W = W;
  job->H = H;
  job->frame_flags = frame_flags;
  job->user = *hf;
  job->user_lo = lo;
  job->user_hi = hi;
  job->ov = ov;
  job->prep = prep;
  const Layout *L = find_layout (c, prep, ov->lazy_inplace, fmt, W, H, frame_flags, &job->slot.frame,
      &job->slot.frame, true);
  layout_spans (L, job->spans);
  if (job->spans.empty ()) {
    c->stage_slots_free.push_back (job->slot);      /* the cue does not touch this frame */
    return 0;
  }
  job->state = StageJob::COPY_IN;
  c->stage_active++;
  c->stats.staged_frames++;
  /* finished jobs nobody waited for do not pile up */
  if (c->stage_jobs.size () > 4096)
    for (auto it = c->stage_jobs.begin (); it != c->stage_jobs.end () && c->stage_jobs.size () > 2048;)
      it = it->second->state == StageJob::DONE ? c->stage_jobs.erase (it) : std::next (it);
  c->stage_jobs[tk] = job;
  c->stage_in.push_back (job);
  c->stage_cv.notify_one ();
  return 0;
}

/* wait (ticket) for a staged frame: 1 = it was one (rc holds its result), 0 = not a staged ticket */
int
stage_wait (Ctx *c, std::unique_lock<std::mutex> &lk, uint64_t ticket, int *rc)
{
  auto it = c->stage_jobs.find (ticket);
  if (it == c->stage_jobs.end ())
    return 0;
  std::shared_ptr<StageJob> job = it->second;
  c->stage_done_cv.wait (lk, [&] { return job->state == StageJob::DONE || c->quit; });
  *rc = job->state == StageJob::DONE ? job->rc : FLUC_TTMLBLEND_ERROR_INVALID_ARGUMENT;
  c->stage_jobs.erase (ticket);
  return 1;
}

/* Every staged frame has reached the pending batch (phase 0) / has been copied back (phase 1). */
void
stage_drain (Ctx *c, std::unique_lock<std::mutex> &lk, int phase)
{
  if (c->stage_workers.empty ())
    return;
  if (phase == 0)
    c->stage_done_cv.wait (lk, [&] { return c->quit || c->sticky || (c->stage_in.empty () && c->stage_copying == 0); });
  else
    c->stage_done_cv.wait (lk, [&] { return c->quit || c->sticky || c->stage_active == 0; });
}

/* fluc_ttmlblend_free: threads first (quit is set), then the staging frames */
void
stage_shutdown (Ctx *c)
{
  {
    std::unique_lock<std::mutex> lk (c->mu);
    c->stage_cv.notify_all ();
    c->stage_gpu_cv.notify_all ();
    c->stage_done_cv.notify_all ();
  }
  for (std::thread &t : c->stage_workers)
    if (t.joinable ())
      t.join ();
  c->stage_workers.clear ();
  if (c->stage_completer.joinable ())
    c->stage_completer.join ();
}

void
stage_free_slots (Ctx *c)
{
  for (auto &kv : c->stage_jobs)
    if (kv.second->state != StageJob::DONE && kv.second->slot.base)
      cudaFreeHost (kv.second->slot.base);
  c->stage_jobs.clear ();
  c->stage_in.clear ();
  c->stage_out.clear ();
  c->stage_gpu.clear ();
  for (PoolEntry &p : c->stage_slots_free)
    cudaFreeHost (p.base);
  c->stage_slots_free.clear ();
}

}  // namespace tbh
