// tools/pciebench.cu -- PCIe ceiling for the host-frame path (scratch, not product).
// H2D only, D2H only and both at once, for frame-row sized pieces from pinned memory.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv) {
  size_t piece = argc > 1 ? atol(argv[1]) : 1382400;   // 360 rows x 3840
  int n = 64, iters = 20;
  uint8_t *h, *d;
  CK(cudaHostAlloc(&h, piece * n * 2, cudaHostAllocDefault));
  CK(cudaMalloc(&d, piece * n * 2));
  cudaStream_t s[8];
  for (auto& x : s) CK(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
  auto run = [&](const char* name, bool up, bool down, int nstreams) {
    for (int w = 0; w < 2; w++) {
      double t0 = now();
      for (int it = 0; it < iters; it++)
        for (int i = 0; i < n; i++) {
          cudaStream_t st = s[i % nstreams];
          if (up) CK(cudaMemcpyAsync(d + piece * i, h + piece * i, piece, cudaMemcpyHostToDevice, st));
          if (down) CK(cudaMemcpyAsync(h + piece * (n + i), d + piece * (n + i), piece, cudaMemcpyDeviceToHost, st));
        }
      double t_issue = now() - t0;
      CK(cudaDeviceSynchronize());
      double t = now() - t0;
      if (w) printf("%-28s piece %7zu B x%d streams: %6.1f GB/s per direction, issue %.1f us/copy, total %.1f us/piece\n", name, piece, nstreams,
                    piece * (double)n * iters / t / 1e9, t_issue / (n * iters * (up + down)) * 1e6, t / (n * iters) * 1e6);
    }
  };
  run("H2D only", true, false, 4);
  run("D2H only", false, true, 4);
  run("H2D + D2H same streams", true, true, 4);
  run("H2D + D2H 8 streams", true, true, 8);
  run("H2D + D2H 1 stream", true, true, 1);
  return 0;
}
