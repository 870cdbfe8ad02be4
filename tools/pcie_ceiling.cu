// tools/pcie_ceiling.cu -- what can the box's PCIe / host memory system carry when SEVERAL GPUs
// move host frames at the same time? (scratch measurement, not product code)
//
// The host-frame path (fluc_ttmlblend_blend_host*) is bound by PCIe: one GPU moves ~37 GB/s each
// way. With one process per GPU the aggregate stopped growing after two GPUs in round 1; this
// program measures the ceiling the path can at best reach, with nothing of the blend in it:
// G workers (one per GPU) run the same traffic pattern at the same time between two barriers,
// as G processes (fork before CUDA is touched) or as G threads of one process.
//
//   patterns  dma_h2d      cudaMemcpyAsync pinned host -> device, 2 streams
//             dma_d2h      device -> pinned host
//             dma_both     both at once on different streams
//             zc_read      a kernel reads pinned host memory (128-bit loads), writes HBM
//             zc_write     a kernel reads HBM, writes pinned host memory
//             zc_inplace   a kernel reads and rewrites the same host bytes: the blend's own shape
//             zc_two       two kernels at once, one reading host memory, one writing it
//             zcr_dmaw     kernel reads host || copy engine writes host
//             dmar_zcw     copy engine reads host || kernel writes host
//             dma_pieces   dma_both in the 128 row pieces of a 32-frame 4K NV12 batch per direction
//             dma_2d       the same pieces as 4 two-dimensional copies per direction (frames at a constant spacing)
//             dma_2d_zcw   2-D copies in || kernel writes host
//
//   build/pcie_ceiling --gpus 0,1,2,3 [--threads] [--secs 0.5] [--sync spin|yield|block]
//                      [--pin] [--alloc default|portable|wc|register] [--patterns a,b,..]
//
// Output: one line per pattern with the aggregate and per-GPU GB/s per direction.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <sched.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf (stderr, "%s: %s\n", #x, cudaGetErrorString (e_)); _exit (3); } } while (0)

static double
now ()
{
  return std::chrono::duration<double> (std::chrono::steady_clock::now ().time_since_epoch ()).count ();
}

template <int U>
__global__ void __launch_bounds__ (256)
move16 (const uint4 *__restrict__ s, uint4 *__restrict__ d, size_t n)
{
  const size_t base = (size_t) blockIdx.x * (256 * U) + threadIdx.x;
  uint4 v[U];
#pragma unroll
  for (int k = 0; k < U; k++) {
    const size_t i = base + (size_t) k * 256;
    if (i < n)
      v[k] = s[i];
  }
#pragma unroll
  for (int k = 0; k < U; k++) {
    const size_t i = base + (size_t) k * 256;
    if (i < n) {
      v[k].x ^= 1u;             /* "blend": the bytes written differ from the bytes read */
      d[i] = v[k];
    }
  }
}

constexpr int kMaxGpus = 16, kMaxPatterns = 12;

struct Shared {
  std::atomic<int> arrived;
  std::atomic<int> generation;
  std::atomic<int> failed;
  double gbs_in[kMaxPatterns][kMaxGpus];     /* host -> device direction */
  double gbs_out[kMaxPatterns][kMaxGpus];    /* device -> host direction */
};

static void
barrier (Shared *sh, int n)
{
  const int gen = sh->generation.load ();
  if (sh->arrived.fetch_add (1) + 1 == n) {
    sh->arrived.store (0);
    sh->generation.fetch_add (1);
  } else {
    while (sh->generation.load () == gen && !sh->failed.load ())
      sched_yield ();
  }
}

struct Options {
  std::vector<int> gpus;
  bool threads = false;
  double secs = 0.5;
  std::string sync = "auto", alloc = "default";
  bool pin = false;
  std::vector<std::string> patterns = { "dma_h2d", "dma_d2h", "dma_both", "zc_read", "zc_write", "zc_inplace" };
};

static void
worker (const Options &o, int w, Shared *sh)
{
  const int G = (int) o.gpus.size ();
  if (o.pin) {
    /* worker w gets the w-th slice of the CPUs this process may use: its pinned pages are first
     * touched from there */
    cpu_set_t all, mine;
    CPU_ZERO (&mine);
    if (sched_getaffinity (0, sizeof all, &all) == 0) {
      std::vector<int> cpus;
      for (int i = 0; i < CPU_SETSIZE; i++)
        if (CPU_ISSET (i, &all))
          cpus.push_back (i);
      const size_t per = std::max<size_t> (1, cpus.size () / G);
      for (size_t i = w * per; i < std::min (cpus.size (), (w + 1) * per); i++)
        CPU_SET (cpus[i], &mine);
      if (CPU_COUNT (&mine))
        sched_setaffinity (0, sizeof mine, &mine);
    }
  }
  CK (cudaSetDevice (o.gpus[w]));
  if (o.sync == "block")
    CK (cudaSetDeviceFlags (cudaDeviceScheduleBlockingSync));
  else if (o.sync == "spin")
    CK (cudaSetDeviceFlags (cudaDeviceScheduleSpin));
  else if (o.sync == "yield")
    CK (cudaSetDeviceFlags (cudaDeviceScheduleYield));
  CK (cudaFree (0));

  /* the 4K NV12 config's pieces: rows under the two cue regions, luma + chroma, 32 frames */
  const size_t piece[4] = { 1382400, 552960, 691200, 276480 };
  size_t per_frame = 0;
  for (size_t p : piece)
    per_frame += p;
  const size_t bytes = per_frame * 32;    /* 92.9 MB, what one e2e step moves each way */
  uint8_t *h1 = nullptr, *h2 = nullptr, *d1 = nullptr, *d2 = nullptr;
  auto host_alloc = [&](uint8_t **p) {
    if (o.alloc == "portable")
      CK (cudaHostAlloc ((void **) p, bytes, cudaHostAllocPortable));
    else if (o.alloc == "wc")
      CK (cudaHostAlloc ((void **) p, bytes, cudaHostAllocWriteCombined));
    else if (o.alloc == "register") {
      void *m = mmap (nullptr, (bytes + (2u << 20) - 1) & ~((size_t) (2u << 20) - 1), PROT_READ | PROT_WRITE,
          MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
      if (m == MAP_FAILED) { perror ("mmap"); _exit (3); }
      madvise (m, bytes, MADV_HUGEPAGE);
      memset (m, 1, bytes);
      CK (cudaHostRegister (m, bytes, cudaHostRegisterDefault));
      *p = (uint8_t *) m;
    } else
      CK (cudaHostAlloc ((void **) p, bytes, cudaHostAllocDefault));
    memset (*p, 1, bytes);
  };
  host_alloc (&h1);
  host_alloc (&h2);
  CK (cudaMalloc ((void **) &d1, bytes));
  CK (cudaMalloc ((void **) &d2, bytes));
  CK (cudaMemset (d1, 2, bytes));
  CK (cudaMemset (d2, 3, bytes));
  cudaStream_t s1, s2;
  /* PCIE_DUMMY_STREAMS=n: n other streams are created first (does the order of creation decide
   * which copy engine a stream's copies use?) */
  if (const char *ds = getenv ("PCIE_DUMMY_STREAMS"))
    for (int i = 0; i < atoi (ds); i++) {
      cudaStream_t dummy;
      CK (cudaStreamCreateWithFlags (&dummy, cudaStreamNonBlocking));
      if (getenv ("PCIE_DUMMY_USED")) {         /* ... and used: a copy each way and a kernel */
        CK (cudaMemcpyAsync (d1, h1, 1 << 20, cudaMemcpyHostToDevice, dummy));
        CK (cudaMemcpyAsync (h2, d2, 1 << 20, cudaMemcpyDeviceToHost, dummy));
        move16<4><<<64, 256, 0, dummy>>> ((const uint4 *) d1, (uint4 *) d2, 65536);
        CK (cudaStreamSynchronize (dummy));
      }
    }
  CK (cudaStreamCreateWithFlags (&s1, cudaStreamNonBlocking));
  CK (cudaStreamCreateWithFlags (&s2, cudaStreamNonBlocking));
  const size_t n16 = bytes / 16;
  const unsigned grid = (unsigned) ((n16 + 1023) / 1024);

  for (size_t pi = 0; pi < o.patterns.size (); pi++) {
    const std::string &p = o.patterns[pi];
    double in_bytes = 0, out_bytes = 0;   /* per iteration */
    auto issue = [&]() {
      if (p == "dma_h2d") {
        CK (cudaMemcpyAsync (d1, h1, bytes, cudaMemcpyHostToDevice, s1));
        in_bytes = (double) bytes;
      } else if (p == "dma_d2h") {
        CK (cudaMemcpyAsync (h2, d2, bytes, cudaMemcpyDeviceToHost, s2));
        out_bytes = (double) bytes;
      } else if (p == "dma_both") {
        CK (cudaMemcpyAsync (d1, h1, bytes, cudaMemcpyHostToDevice, s1));
        CK (cudaMemcpyAsync (h2, d2, bytes, cudaMemcpyDeviceToHost, s2));
        in_bytes = out_bytes = (double) bytes;
      } else if (p == "zc_read") {
        move16<4><<<grid, 256, 0, s1>>> ((const uint4 *) h1, (uint4 *) d1, n16);
        in_bytes = (double) bytes;
      } else if (p == "zc_write") {
        move16<4><<<grid, 256, 0, s1>>> ((const uint4 *) d2, (uint4 *) h2, n16);
        out_bytes = (double) bytes;
      } else if (p == "zcr_dmaw") {           /* a kernel reads host memory while the copy engine writes it */
        move16<4><<<grid, 256, 0, s1>>> ((const uint4 *) h1, (uint4 *) d1, n16);
        CK (cudaMemcpyAsync (h2, d2, bytes, cudaMemcpyDeviceToHost, s2));
        in_bytes = out_bytes = (double) bytes;
      } else if (p == "dmar_zcw") {           /* the copy engine reads host memory while a kernel writes it */
        CK (cudaMemcpyAsync (d1, h1, bytes, cudaMemcpyHostToDevice, s1));
        move16<4><<<grid, 256, 0, s2>>> ((const uint4 *) d2, (uint4 *) h2, n16);
        in_bytes = out_bytes = (double) bytes;
      } else if (p == "zc_two") {             /* two kernels at once: one reads host, one writes host */
        move16<4><<<grid, 256, 0, s1>>> ((const uint4 *) h1, (uint4 *) d1, n16);
        move16<4><<<grid, 256, 0, s2>>> ((const uint4 *) d2, (uint4 *) h2, n16);
        in_bytes = out_bytes = (double) bytes;
      } else if (p == "dma_pieces") {         /* both directions, in the 128 row pieces of a 32-frame batch */
        size_t off = 0;
        for (int f = 0; f < 32; f++)
          for (size_t pc : piece) {
            CK (cudaMemcpyAsync (d1 + off, h1 + off, pc, cudaMemcpyHostToDevice, s1));
            CK (cudaMemcpyAsync (h2 + off, d2 + off, pc, cudaMemcpyDeviceToHost, s2));
            off += pc;
          }
        in_bytes = out_bytes = (double) bytes;
      } else if (p == "dma_2d") {             /* the same pieces, but one 2-D copy per piece shape moves it for all 32
                                               * frames (frames at a constant spacing): 4 copies per direction */
        size_t off = 0;
        for (size_t pc : piece) {
          CK (cudaMemcpy2DAsync (d1 + off, per_frame, h1 + off, per_frame, pc, 32, cudaMemcpyHostToDevice, s1));
          CK (cudaMemcpy2DAsync (h2 + off, per_frame, d2 + off, per_frame, pc, 32, cudaMemcpyDeviceToHost, s2));
          off += pc;
        }
        in_bytes = out_bytes = (double) bytes;
      } else if (p == "dma_2d_zcw") {         /* 2-D copies in, a kernel writes the result back to the host */
        size_t off = 0;
        for (size_t pc : piece) {
          CK (cudaMemcpy2DAsync (d1 + off, per_frame, h1 + off, per_frame, pc, 32, cudaMemcpyHostToDevice, s1));
          off += pc;
        }
        move16<4><<<grid, 256, 0, s2>>> ((const uint4 *) d2, (uint4 *) h2, n16);
        in_bytes = out_bytes = (double) bytes;
      } else {                  /* zc_inplace */
        move16<4><<<grid, 256, 0, s1>>> ((const uint4 *) h1, (uint4 *) h1, n16);
        in_bytes = out_bytes = (double) bytes;
      }
    };
    if (p == "dma_2d_wide") {
      /* as dma_2d, but the 32 frames are whole 4K NV12 frames (12.4 MB apart), as in the product */
      const size_t fb = 12441600, yoff[4] = { (size_t) 1728 * 3840, (size_t) 72 * 3840, (size_t) 3840 * 2160 + (size_t) 864 * 3840,
        (size_t) 3840 * 2160 + (size_t) 36 * 3840 };
      uint8_t *hw1, *hw2, *dw1, *dw2;
      CK (cudaHostAlloc ((void **) &hw1, fb * 32, cudaHostAllocDefault));
      CK (cudaHostAlloc ((void **) &hw2, fb * 32, cudaHostAllocDefault));
      CK (cudaMalloc ((void **) &dw1, fb * 32));
      CK (cudaMalloc ((void **) &dw2, fb * 32));
      memset (hw1, 1, fb * 32);
      memset (hw2, 1, fb * 32);
      auto go = [&]() {
        for (int k = 0; k < 4; k++) {
          CK (cudaMemcpy2DAsync (dw1 + yoff[k], fb, hw1 + yoff[k], fb, piece[k], 32, cudaMemcpyHostToDevice, s1));
          CK (cudaMemcpy2DAsync (hw2 + yoff[k], fb, dw2 + yoff[k], fb, piece[k], 32, cudaMemcpyDeviceToHost, s2));
        }
      };
      go ();
      CK (cudaStreamSynchronize (s1));
      CK (cudaStreamSynchronize (s2));
      barrier (sh, G);
      const double t0 = now ();
      int n = 0;
      double t = 0;
      do {
        go ();
        go ();
        CK (cudaStreamSynchronize (s1));
        CK (cudaStreamSynchronize (s2));
        n += 2;
        t = now () - t0;
      } while (t < o.secs);
      sh->gbs_in[pi][w] = sh->gbs_out[pi][w] = (double) bytes * n / t / 1e9;
      barrier (sh, G);
      cudaFreeHost (hw1); cudaFreeHost (hw2); cudaFree (dw1); cudaFree (dw2);
      continue;
    }
    if (p == "dma_pipe") {
      /* the product's DMA batch pipeline: copy-in stream, blend stream, copy-out stream, three
       * device sets; the host hands over batch i+1 once batch i-1 has come back */
      cudaStream_t s3;
      CK (cudaStreamCreateWithFlags (&s3, cudaStreamNonBlocking));
      cudaEvent_t ein, ek, done[4], setdone[3];
      CK (cudaEventCreateWithFlags (&ein, cudaEventDisableTiming));
      CK (cudaEventCreateWithFlags (&ek, cudaEventDisableTiming));
      for (auto &e : done) CK (cudaEventCreateWithFlags (&e, cudaEventDisableTiming));
      for (auto &e : setdone) CK (cudaEventCreateWithFlags (&e, cudaEventDisableTiming));
      uint8_t *dset[3];
      for (auto &d : dset) CK (cudaMalloc ((void **) &d, bytes));
      auto batch = [&](int i) {
        uint8_t *h = (i & 1) ? h2 : h1, *d = dset[i % 3];
        if (i >= 3) CK (cudaStreamWaitEvent (s1, setdone[i % 3], 0));
        size_t off = 0;
        for (size_t pc : piece) { CK (cudaMemcpy2DAsync (d + off, per_frame, h + off, per_frame, pc, 32, cudaMemcpyHostToDevice, s1)); off += pc; }
        CK (cudaEventRecord (ein, s1));
        CK (cudaStreamWaitEvent (s3, ein, 0));
        move16<4><<<grid, 256, 0, s3>>> ((const uint4 *) d, (uint4 *) d, n16);
        CK (cudaEventRecord (ek, s3));
        CK (cudaStreamWaitEvent (s2, ek, 0));
        off = 0;
        for (size_t pc : piece) { CK (cudaMemcpy2DAsync (h + off, per_frame, d + off, per_frame, pc, 32, cudaMemcpyDeviceToHost, s2)); off += pc; }
        CK (cudaEventRecord (setdone[i % 3], s2));
        CK (cudaEventRecord (done[i & 3], s2));
      };
      for (int i = 0; i < 4; i++) { batch (i); if (i) CK (cudaEventSynchronize (done[(i - 1) & 3])); }
      CK (cudaStreamSynchronize (s2));
      barrier (sh, G);
      const double t0 = now ();
      int i = 4, n = 0;
      double t = 0;
      /* PCIE_PIPE_DEPTH=2: the host runs two batches ahead, so the copy-in of batch i+1 is queued
       * while the copy-out of batch i still waits for its event */
      const int depth = getenv ("PCIE_PIPE_DEPTH") ? atoi (getenv ("PCIE_PIPE_DEPTH")) : 1;
      do {
        batch (i);
        CK (cudaEventSynchronize (done[(i - depth) & 3]));
        i++; n++;
        t = now () - t0;
      } while (t < o.secs);
      CK (cudaStreamSynchronize (s2));
      t = now () - t0;
      sh->gbs_in[pi][w] = sh->gbs_out[pi][w] = (double) bytes * n / t / 1e9;
      barrier (sh, G);
      continue;
    }
    for (int i = 0; i < 2; i++)
      issue ();
    CK (cudaStreamSynchronize (s1));
    CK (cudaStreamSynchronize (s2));
    barrier (sh, G);
    const double t0 = now ();
    int iters = 0;
    double t = 0;
    do {
      issue ();
      issue ();
      CK (cudaStreamSynchronize (s1));
      CK (cudaStreamSynchronize (s2));
      iters += 2;
      t = now () - t0;
    } while (t < o.secs);
    CK (cudaGetLastError ());
    sh->gbs_in[pi][w] = in_bytes * iters / t / 1e9;
    sh->gbs_out[pi][w] = out_bytes * iters / t / 1e9;
    barrier (sh, G);
  }
}

int
main (int argc, char **argv)
{
  Options o;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    auto next = [&]() -> std::string { return i + 1 < argc ? argv[++i] : ""; };
    auto split = [](const std::string &s) {
      std::vector<std::string> out;
      size_t b = 0;
      while (b <= s.size ()) {
        const size_t e = s.find (',', b);
        out.push_back (s.substr (b, e == std::string::npos ? std::string::npos : e - b));
        if (e == std::string::npos)
          break;
        b = e + 1;
      }
      return out;
    };
    if (a == "--gpus") {
      for (auto &s : split (next ()))
        o.gpus.push_back (atoi (s.c_str ()));
    } else if (a == "--threads")
      o.threads = true;
    else if (a == "--secs")
      o.secs = atof (next ().c_str ());
    else if (a == "--sync")
      o.sync = next ();
    else if (a == "--alloc")
      o.alloc = next ();
    else if (a == "--pin")
      o.pin = true;
    else if (a == "--patterns")
      o.patterns = split (next ());
    else {
      fprintf (stderr, "unknown option %s\n", a.c_str ());
      return 2;
    }
  }
  if (o.gpus.empty ())
    o.gpus.push_back (0);
  if (o.gpus.size () > (size_t) kMaxGpus || o.patterns.size () > (size_t) kMaxPatterns)
    return 2;
  const int G = (int) o.gpus.size ();
  Shared *sh = (Shared *) mmap (nullptr, sizeof (Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (sh == MAP_FAILED)
    return 1;
  memset ((void *) sh, 0, sizeof *sh);

  if (o.threads) {
    std::vector<std::thread> th;
    for (int w = 0; w < G; w++)
      th.emplace_back (worker, std::cref (o), w, sh);
    for (auto &t : th)
      t.join ();
  } else {
    std::vector<pid_t> kids;
    for (int w = 0; w < G; w++) {
      const pid_t pid = fork ();          /* before CUDA is touched */
      if (pid == 0) {
        worker (o, w, sh);
        _exit (0);
      }
      kids.push_back (pid);
    }
    int bad = 0;
    for (pid_t k : kids) {
      int st = 0;
      waitpid (k, &st, 0);
      if (!WIFEXITED (st) || WEXITSTATUS (st) != 0) {
        bad++;
        sh->failed.store (1);
      }
    }
    if (bad) {
      fprintf (stderr, "%d worker(s) failed\n", bad);
      return 1;
    }
  }

  printf ("# %d GPU(s) [", G);
  for (int w = 0; w < G; w++)
    printf ("%s%d", w ? "," : "", o.gpus[w]);
  printf ("] as %s, sync=%s, alloc=%s, pin=%d, %.2f s per pattern, 92.9 MB per iteration and direction\n",
      o.threads ? "threads of one process" : "one process per GPU", o.sync.c_str (), o.alloc.c_str (), o.pin ? 1 : 0,
      o.secs);
  for (size_t pi = 0; pi < o.patterns.size (); pi++) {
    double in = 0, out = 0;
    for (int w = 0; w < G; w++) {
      in += sh->gbs_in[pi][w];
      out += sh->gbs_out[pi][w];
    }
    printf ("%-11s aggregate h2d %7.1f GB/s  d2h %7.1f GB/s   per GPU:", o.patterns[pi].c_str (), in, out);
    for (int w = 0; w < G; w++)
      printf (" %.1f/%.1f", sh->gbs_in[pi][w], sh->gbs_out[pi][w]);
    printf ("\n");
  }
  return 0;
}
