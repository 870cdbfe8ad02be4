/*
 * scheduler.cu -- the multi-stream batch scheduler: pending frames -> group launches and table
 * launches, table slots uploaded on a copy stream, batch events / tickets, the linger thread.
 * Shape after the reference's worker-thread-under-a-monitor
 * (/root/reference/libs/flu/downloader/lib/fludownloader.c:490-532).
 */
#include "ttmlblend_internal.h"

namespace tbh {

cudaEvent_t
event_get (Ctx *c)
{
  if (!c->event_pool.empty ()) {
    cudaEvent_t e = c->event_pool.back ();
    c->event_pool.pop_back ();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreateWithFlags (&e, cudaEventDisableTiming | (c->blocking_sync ? cudaEventBlockingSync : 0));
  return e;
}

int
slot_reserve (Ctx *c, TableSlot &s, size_t n, size_t words)
{
  if (s.cap >= n && s.cap_words >= words)
    return 0;
  const size_t cap = std::max<size_t> (256, n * 2), cap_words = std::max<size_t> (4096, words * 2);
  if (s.h_jobs) cudaFreeHost (s.h_jobs);
  if (s.h_begin) cudaFreeHost (s.h_begin);
  if (s.d_jobs) cudaFree (s.d_jobs);
  if (s.d_begin) cudaFree (s.d_begin);
  s.h_jobs = nullptr; s.h_begin = nullptr; s.d_jobs = nullptr; s.d_begin = nullptr;
  s.cap = s.cap_words = 0;
  CU (c, cudaHostAlloc ((void **) &s.h_jobs, cap * sizeof (PlaneJob), cudaHostAllocDefault));
  CU (c, cudaHostAlloc ((void **) &s.h_begin, cap_words * sizeof (uint32_t), cudaHostAllocDefault));
  CU (c, cudaMalloc ((void **) &s.d_jobs, cap * sizeof (PlaneJob)));
  CU (c, cudaMalloc ((void **) &s.d_begin, cap_words * sizeof (uint32_t)));
  if (!s.copied)
    CU (c, cudaEventCreateWithFlags (&s.copied, cudaEventDisableTiming));
  if (!s.uploaded)
    CU (c, cudaEventCreateWithFlags (&s.uploaded, cudaEventDisableTiming));
  s.cap = cap;
  s.cap_words = cap_words;
  return 0;
}

/* Copies `jobs` (one PlaneKind) into a table slot and launches the kernel. The slot's word
 * array holds the first chunk of every job, then the coarse index (job of every
 * 2^kCoarseShift-th chunk) the kernel starts its search from. */
int
launch_jobs (Ctx *c, TableSlot &s, const PlaneJob *jobs, size_t n, int kind, bool fast, int sync,
    cudaStream_t stream)
{
  if (n == 0)
    return 0;
  if (s.copied && s.cap)
    CU (c, cudaEventSynchronize (s.copied));
  uint64_t total64 = 0;
  for (size_t i = 0; i < n; i++)
    total64 += jobs[i].n_chunks;
  if (total64 >= (1ull << 31))
    return FLUC_TTMLBLEND_ERROR_INVALID_ARGUMENT;
  const size_t n_coarse = (size_t) ((total64 + (1u << kCoarseShift) - 1) >> kCoarseShift);
  int rc = slot_reserve (c, s, n, n + n_coarse);
  if (rc)
    return rc;
  uint32_t total = 0;
  uint32_t *coarse = s.h_begin + n;
  size_t next = 0;
  for (size_t i = 0; i < n; i++) {
    s.h_jobs[i] = jobs[i];
    s.h_begin[i] = total;
    total += jobs[i].n_chunks;
    for (; next < n_coarse && ((uint32_t) next << kCoarseShift) < total; next++)
      coarse[next] = (uint32_t) i;
  }
  /* the table goes up on the copy stream, so that it overlaps the kernel still
   * running on `stream`; the slot is free (its last kernel waited on above) */
  CU (c, cudaMemcpyAsync (s.d_jobs, s.h_jobs, n * sizeof (PlaneJob), cudaMemcpyHostToDevice, c->table_stream));
  CU (c, cudaMemcpyAsync (s.d_begin, s.h_begin, (n + n_coarse) * sizeof (uint32_t), cudaMemcpyHostToDevice,
          c->table_stream));
  CU (c, cudaEventRecord (s.uploaded, c->table_stream));
  CU (c, cudaStreamWaitEvent (stream, s.uploaded, 0));
  CU (c, launch_blend (s.d_jobs, s.d_begin, s.d_begin + n, (int) n, total, kind, fast, sync, stream));
  CU (c, cudaEventRecord (s.copied, stream));    /* slot busy until this kernel is done */
  c->stats.launches++;
  return 0;
}

/* Retires the batches that have finished. Batches complete in launch order (one stream), so
 * the boundary between finished and running ones is found by bisection: a handful of event
 * queries however many batches are in flight, instead of one per batch. */
void
reap_batches (Ctx *c)
{
  size_t n = c->batches.size ();
  if (n == 0)
    return;
  size_t done = 0;              /* batches [0, done) have finished */
  if (cudaEventQuery (c->batches[n - 1].done) == cudaSuccess) {
    done = n;
  } else {
    cudaGetLastError ();
    size_t lo = 0, hi = n - 1;  /* batch hi is running; find the first running one */
    while (lo < hi) {
      const size_t mid = (lo + hi) / 2;
      if (cudaEventQuery (c->batches[mid].done) == cudaSuccess)
        lo = mid + 1;
      else {
        cudaGetLastError ();
        hi = mid;
      }
    }
    done = lo;
  }
  for (size_t i = 0; i < done; i++) {
    Batch &b = c->batches.front ();
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
    b.keep.clear ();            /* drops the overlay references, keeps the capacity */
    c->keep_pool.push_back (std::move (b.keep));
    c->batches.pop_front ();
  }
  if (c->batches.empty ()) {
    c->inflight_host.clear ();
    /* nothing is running any more: the next launch depends on nothing */
    c->inflight_dst.clear ();
    c->inflight_src.clear ();
  }
}

/* Launches everything pending as one batch (per plane kind). mu held. */
int
launch_pending (Ctx *c)
{
  if (c->pending.empty ())
    return 0;
  NvtxRange nvtx ("ttmlblend.launch_batch");
  /* finished batches are retired in bulk: an event query per launch would be a fifth of the cost
   * of a one-frame launch (wait / sync / stats retire them too) */
  if (c->batches.size () >= 32)
    reap_batches (c);
  Batch b = {};
  b.last_ticket = c->pending.back ().ticket;
  if (!c->keep_pool.empty ()) {
    b.keep = std::move (c->keep_pool.back ());
    c->keep_pool.pop_back ();
  }
  std::vector<PlaneJob> *by_kind = c->by_kind;        /* PlaneKind x {byte-granular, fast}; scratch */
  for (int k = 0; k < kPlaneKinds * 2; k++)
    by_kind[k].clear ();
  std::vector<Group> &groups = c->groups;
  groups.clear ();
  size_t n_multis = 0;
  std::vector<int> &frame_group = c->frame_group;
  frame_group.assign (c->pending.size (), -1);
  size_t fi = 0;
  for (PendingFrame &f : c->pending) {
    emit_table_jobs (f.layout->jobs, f, by_kind);
    if (f.layout->grouped) {
      int gi = -1;
      for (size_t k = 0; k < groups.size (); k++)
        if (group_accepts (groups[k], f)) {
          gi = (int) k;
          break;
        }
      if (gi < 0) {
        groups.emplace_back ();
        gi = (int) groups.size () - 1;
        group_start (groups[gi], f);
      }
      group_add (groups[gi], f);
      frame_group[fi] = gi;
    }
    fi++;
    if (f.overlay && (b.keep.empty () || b.keep.back () != f.overlay))
      b.keep.push_back (f.overlay);     /* consecutive frames of one cue: one reference */
    if (f.prep && !f.prep->blend_waited) {
      /* once per prepared overlay: later launches follow in stream order */
      f.prep->blend_waited = true;
      CU (c, cudaStreamWaitEvent (c->blend_stream, f.prep->ready, 0));
    }
    c->stats.frames_blended++;
    c->stats.algorithmic_bytes += f.layout->algo_bytes;
  }
  /* Many streams with different cue layouts give many groups of a frame or two, and a launch
   * of a few hundred chunks leaves most of the chip idle (tools/many_cues_probe.py: 61 group
   * launches for 256 1080p frames ran at half the rate of 4). Small groups are dissolved --
   * when that saves a launch, i.e. there are at least two of them or a table launch of
   * whole-vector windows is going out anyway. */
  {
    size_t n_small = 0, n_table = 0;
    for (int k = 1; k < kPlaneKinds * 2; k += 2)
      n_table += by_kind[k].size ();
    for (Group &g : groups) {
      g.dissolved = (uint64_t) g.P.h.n_frames * g.P.h.chunks_per_frame < (uint64_t) kMinGroupChunks;
      n_small += g.dissolved ? 1 : 0;
    }
    if (n_small < 2 && n_table == 0)
      for (Group &g : groups)
        g.dissolved = false;
    /* their frames are packed into multi-layout launches (band lists in the parameters,
     * up to 64 frames and kMaxMultiBands bands each); what does not fit goes to the table */
    fi = 0;
    for (PendingFrame &f : c->pending) {
      const int gi = frame_group[fi++];
      if (gi < 0 || !groups[gi].dissolved)
        continue;
      bool placed = false;
      if (c->use_multi) {
        for (size_t k = 0; k < n_multis && !placed; k++)
          placed = multi_add (*c->multis[k], f);
        if (!placed) {
          if (n_multis == c->multis.size ())
            c->multis.emplace_back (new MultiGroup ());
          multi_start (*c->multis[n_multis], f);
          placed = multi_add (*c->multis[n_multis], f);
          n_multis += placed ? 1 : 0;
        }
      }
      if (!placed)
        emit_table_jobs (f.layout->gjobs, f, by_kind);
    }
  }
  const uint64_t first_ticket = c->pending.front ().ticket;
  /* Programmatic dependent launch: a launch may start while earlier ones still run. That is
   * only right for frames that do not touch what those read or write: the ranges of every
   * batch launched since the last dependent one are remembered, and a batch that writes into
   * them, or reads what they write, has its first launch wait for everything before it
   * (JF_DEP); after such a launch only its own ranges can still be in flight. */
  int sync = 0;
  if (c->use_pdl) {
    const bool timed = c->profiling && (c->profile_seq % c->profile_every) == 0;
    bool dep = c->inflight_dst.v.size () + c->inflight_src.v.size () > 4096 || timed || c->isolate_next;
    c->isolate_next = timed;    /* an event pair measures one launch alone: neither it nor its successor overlaps */
    if (!dep)
      dep = sets_overlap (c->pending_dst, c->inflight_dst) || sets_overlap (c->pending_dst, c->inflight_src) ||
          sets_overlap (c->pending_src, c->inflight_dst);
    if (dep) {
      c->inflight_dst.clear ();
      c->inflight_src.clear ();
      c->stats.dependent_launches++;
    }
    c->inflight_dst.merge (c->pending_dst);
    c->inflight_src.merge (c->pending_src);
    sync = JF_PDL | (dep ? JF_DEP : 0);
  }
  c->pending.clear ();
  c->pending_dst.clear ();
  c->pending_src.clear ();
  size_t n_launches = n_multis;
  for (Group &g : groups)
    n_launches += g.dissolved ? 0 : 1;
  for (int k = 0; k < kPlaneKinds * 2; k++)
    n_launches += !by_kind[k].empty ();
  /* From here on the frames are no longer queued: whatever happens, a batch with a `done` event
   * is pushed, so that wait() on these tickets synchronises with what did get launched, and a
   * failure half way (out of memory for a table slot, say -- the context stays usable) is
   * remembered for them instead of being reported as finished work. */
  auto issue = [&]() -> int {
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
      if (g.dissolved)
        continue;
      CU (c, launch_group (g.P, g.kind, sync, c->blend_stream));
      sync &= ~JF_DEP;          /* the launches of one batch are independent of each other */
      c->stats.launches++;
      c->stats.group_launches++;
      if (g.P.h.flags & JF_LAZY)
        c->stats.lazy_launches++;
      if (g.P.h.flags & JF_OPAQUE)
        c->stats.opaque_skip_launches++;
    }
    for (size_t k = 0; k < n_multis; k++) {
      MultiGroup &m = *c->multis[k];
      CU (c, launch_multi (m.P, m.kind, sync, c->blend_stream));
      sync &= ~JF_DEP;
      c->stats.launches++;
      c->stats.multi_launches++;
      if (m.P.flags & JF_LAZY)
        c->stats.lazy_launches++;
      if (m.P.flags & JF_OPAQUE)
        c->stats.opaque_skip_launches++;
    }
    for (int k = 0; k < kPlaneKinds * 2; k++) {
      if (by_kind[k].empty ())
        continue;
      TableSlot &s = c->slots[c->next_slot];
      c->next_slot = (c->next_slot + 1) % kTableSlots;
      int rc = launch_jobs (c, s, by_kind[k].data (), by_kind[k].size (), k / 2, (k & 1) != 0, sync,
          c->blend_stream);
      if (rc)
        return rc;
      sync &= ~JF_DEP;
    }
    if (b.t1)
      CU (c, cudaEventRecord (b.t1, c->blend_stream));
    return 0;
  };
  const int rc = issue ();
  if (rc) {
    if (c->failed_ranges.size () >= 64)
      c->failed_ranges.pop_front ();
    c->failed_ranges.push_back ({ first_ticket, b.last_ticket, rc });
    if (b.t0) { c->timing_pool.push_back (b.t0); b.t0 = nullptr; }
    if (b.t1) { c->timing_pool.push_back (b.t1); b.t1 = nullptr; }
  }
  b.done = event_get (c);
  if (b.done && cudaEventRecord (b.done, c->blend_stream) == cudaSuccess) {
    c->batches.push_back (std::move (b));
  } else {
    /* not even an event: only a broken context gets here */
    cudaGetLastError ();
    if (b.done)
      c->event_pool.push_back (b.done);
    if (!c->sticky) {
      c->sticky = FLUC_TTMLBLEND_ERROR_CUDA;
      c->cuda_error = "launch_pending: cannot record the batch event";
    }
  }
  c->launched_cv.notify_all ();
  return rc ? rc : c->sticky;
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
        c->pending_src.clear ();
      }
      c->launched_cv.notify_all ();   /* also when the launch failed: waiters must not sleep on */
    } else {
      c->cv.wait_until (lk, deadline);
    }
  }
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

}  // namespace tbh
