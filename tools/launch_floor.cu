// tools/launch_floor.cu -- what does ONE small launch cost on this box, on the host and on the
// device? (scratch) The single-frame configs (720p I420, 1080p NV12: 1-3 MB per step) are bound
// by launch cost, not HBM; this gives the floor the runtime's own per-launch work sits on.
//   empty kernel / 1 KB / 4 KB / 20 KB of __grid_constant__ parameters, with and without an
//   event record per launch, back to back on one stream: host us per launch (issue loop) and
//   device us per launch (events around the loop).
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf ("%s: %s\n", #x, cudaGetErrorString (e_)); return 1; } } while (0)

template <int N> struct Params { uint32_t w[N / 4]; };

template <int N>
__global__ void __launch_bounds__ (256)
k_params (const __grid_constant__ Params<N> p, uint32_t *out)
{
  if (p.w[threadIdx.x % (N / 4)] == 0xdeadbeefu)
    out[0] = 1;
}

__global__ void __launch_bounds__ (256)
k_copy (const uint4 *__restrict__ s, uint4 *__restrict__ d, size_t n)
{
  const size_t base = (size_t) blockIdx.x * 1024 + threadIdx.x;
  uint4 v[4];
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (base + k * 256 < n)
      v[k] = s[base + k * 256];
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (base + k * 256 < n)
      d[base + k * 256] = v[k];
}

/* the same copy, but every CTA lets the next grid of the stream start at once (programmatic
 * dependent launch): launched with programmaticStreamSerializationAllowed, independent work */
__global__ void __launch_bounds__ (256)
k_copy_pdl (const uint4 *__restrict__ s, uint4 *__restrict__ d, size_t n)
{
  asm volatile ("griddepcontrol.launch_dependents;" ::: "memory");
  const size_t base = (size_t) blockIdx.x * 1024 + threadIdx.x;
  uint4 v[4];
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (base + k * 256 < n)
      v[k] = s[base + k * 256];
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (base + k * 256 < n)
      d[base + k * 256] = v[k];
}

static cudaError_t
launch_pdl (const uint4 *s, uint4 *d, size_t n, cudaStream_t st)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3 ((unsigned) ((n + 1023) / 1024));
  cfg.blockDim = dim3 (256);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx (&cfg, k_copy_pdl, s, d, n);
}

static double
now ()
{
  return std::chrono::duration<double> (std::chrono::steady_clock::now ().time_since_epoch ()).count ();
}

template <typename F>
static int
run (const char *name, F launch, cudaStream_t st, bool with_event, int n = 20000)
{
  cudaEvent_t a, b, evs[64];
  CK (cudaEventCreate (&a));
  CK (cudaEventCreate (&b));
  for (auto &e : evs)
    CK (cudaEventCreateWithFlags (&e, cudaEventDisableTiming));
  for (int i = 0; i < 200; i++)
    launch ();
  CK (cudaStreamSynchronize (st));
  CK (cudaEventRecord (a, st));
  const double t0 = now ();
  for (int i = 0; i < n; i++) {
    launch ();
    if (with_event)
      cudaEventRecord (evs[i & 63], st);
  }
  const double t_issue = now () - t0;
  CK (cudaEventRecord (b, st));
  CK (cudaEventSynchronize (b));
  float ms = 0;
  CK (cudaEventElapsedTime (&ms, a, b));
  CK (cudaGetLastError ());
  printf ("%-46s host %6.2f us/launch   device %6.2f us/launch\n", name, t_issue / n * 1e6, ms / n * 1e3);
  return 0;
}

int
main ()
{
  cudaStream_t st;
  CK (cudaStreamCreateWithFlags (&st, cudaStreamNonBlocking));
  uint32_t *out;
  CK (cudaMalloc (&out, 256));
  static Params<1024> p1 = {};
  static Params<4096> p4 = {};
  static Params<20480> p20 = {};
  static Params<32000> p32 = {};
  const size_t bytes = 3110400;           /* one 1080p NV12 frame */
  uint4 *s, *d;
  CK (cudaMalloc (&s, bytes));
  CK (cudaMalloc (&d, bytes));
  const size_t n16 = bytes / 16;
  const size_t big = (size_t) 32 * 12441600;
  uint4 *sb, *db;
  CK (cudaMalloc (&sb, big));
  CK (cudaMalloc (&db, big));
  const size_t nbig = big / 16;
  for (int ev = 0; ev < 2; ev++) {
    const char *sfx = ev ? " + event record" : "";
    char name[96];
    snprintf (name, sizeof name, "1 CTA, 1 KB parameters%s", sfx);
    run (name, [&] { k_params<1024><<<1, 256, 0, st>>> (p1, out); }, st, ev);
    snprintf (name, sizeof name, "1 CTA, 4 KB parameters%s", sfx);
    run (name, [&] { k_params<4096><<<1, 256, 0, st>>> (p4, out); }, st, ev);
    snprintf (name, sizeof name, "1 CTA, 20 KB parameters%s", sfx);
    run (name, [&] { k_params<20480><<<1, 256, 0, st>>> (p20, out); }, st, ev);
    snprintf (name, sizeof name, "1 CTA, 32 KB parameters%s", sfx);
    run (name, [&] { k_params<32000><<<1, 256, 0, st>>> (p32, out); }, st, ev);
    snprintf (name, sizeof name, "190 CTAs, 20 KB parameters%s", sfx);
    run (name, [&] { k_params<20480><<<190, 256, 0, st>>> (p20, out); }, st, ev);
    snprintf (name, sizeof name, "1080p NV12 frame copy (3.1 MB -> 3.1 MB)%s", sfx);
    run (name, [&] { k_copy<<<(unsigned) ((n16 + 1023) / 1024), 256, 0, st>>> (s, d, n16); }, st, ev);
    snprintf (name, sizeof name, "same, programmatic dependent launch%s", sfx);
    run (name, [&] { launch_pdl (s, d, n16, st); }, st, ev);
    snprintf (name, sizeof name, "398 MB copy (32 x 4K NV12)%s", sfx);
    run (name, [&] { k_copy<<<(unsigned) ((nbig + 1023) / 1024), 256, 0, st>>> (sb, db, nbig); }, st, ev, 400);
    snprintf (name, sizeof name, "same, programmatic dependent launch%s", sfx);
    run (name, [&] { launch_pdl (sb, db, nbig, st); }, st, ev, 400);
  }
  return 0;
}
