// tools/zcbench.cu -- how fast can SMs move pinned HOST memory over PCIe? (scratch)
// LDG/STG vs TMA bulk, read-only, write-only and both directions at once.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// mode 0: host -> device, 1: device -> host, 2: host -> host (in place style: read A write A)
template <int U>
__global__ void __launch_bounds__(256) zc_ldg(const uint4* __restrict__ s, uint4* __restrict__ d, size_t n) {
  size_t base = (size_t)blockIdx.x * (256 * U) + threadIdx.x;
  uint4 v[U];
#pragma unroll
  for (int k = 0; k < U; k++) { size_t i = base + (size_t)k * 256; if (i < n) v[k] = s[i]; }
#pragma unroll
  for (int k = 0; k < U; k++) { size_t i = base + (size_t)k * 256; if (i < n) d[i] = v[k]; }
}

// one CTA moves PIECE bytes: TMA bulk load global(host) -> smem, then bulk store smem -> global(host or device)
template <int PIECE>
__global__ void __launch_bounds__(32) zc_tma(const uint8_t* __restrict__ s, uint8_t* __restrict__ d, size_t bytes) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x != 0) return;
  size_t off = (size_t)blockIdx.x * PIECE;
  if (off >= bytes) return;
  unsigned a = (unsigned)__cvta_generic_to_shared(&bar), m = (unsigned)__cvta_generic_to_shared(sm);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(a));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(a), "r"(PIECE) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(m), "l"(s + off), "r"(PIECE), "r"(a) : "memory");
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" :: "r"(a) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(d + off), "r"(m), "r"(PIECE) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F> float timeit(F f, int iters = 10) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 2; i++) f();
  cudaEventRecord(a); for (int i = 0; i < iters; i++) f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); CK(cudaGetLastError()); return ms / iters;
}

int main() {
  size_t bytes = 96ull << 20;
  uint8_t *h1, *h2, *d1, *d2;
  CK(cudaHostAlloc(&h1, bytes, cudaHostAllocDefault)); CK(cudaHostAlloc(&h2, bytes, cudaHostAllocDefault));
  CK(cudaMalloc(&d1, bytes)); CK(cudaMalloc(&d2, bytes));
  size_t n = bytes / 16;
  auto rep = [&](const char* nm, float ms, double dirs) { printf("%-44s %8.3f ms  %6.1f GB/s per direction%s\n", nm, ms, bytes / ms / 1e6, dirs > 1 ? " (both at once)" : ""); };
  rep("LDG U4 host->device", timeit([&] { zc_ldg<4><<<(n + 1023) / 1024, 256>>>((uint4*)h1, (uint4*)d1, n); }), 1);
  rep("LDG U8 host->device", timeit([&] { zc_ldg<8><<<(n + 2047) / 2048, 256>>>((uint4*)h1, (uint4*)d1, n); }), 1);
  rep("STG U4 device->host", timeit([&] { zc_ldg<4><<<(n + 1023) / 1024, 256>>>((uint4*)d1, (uint4*)h2, n); }), 1);
  rep("LDG/STG U4 host->host (same buffer)", timeit([&] { zc_ldg<4><<<(n + 1023) / 1024, 256>>>((uint4*)h1, (uint4*)h1, n); }), 2);
  rep("LDG/STG U4 host->host (other buffer)", timeit([&] { zc_ldg<4><<<(n + 1023) / 1024, 256>>>((uint4*)h1, (uint4*)h2, n); }), 2);
  rep("LDG/STG U1 host->host", timeit([&] { zc_ldg<1><<<(n + 255) / 256, 256>>>((uint4*)h1, (uint4*)h2, n); }), 2);
  {
    constexpr int P = 16384;
    rep("TMA 16K host->device", timeit([&] { zc_tma<P><<<bytes / P, 32, P>>>(h1, d1, bytes); }), 1);
    rep("TMA 16K device->host", timeit([&] { zc_tma<P><<<bytes / P, 32, P>>>(d1, h2, bytes); }), 1);
    rep("TMA 16K host->host", timeit([&] { zc_tma<P><<<bytes / P, 32, P>>>(h1, h2, bytes); }), 2);
    constexpr int P2 = 65536;
    CK(cudaFuncSetAttribute(zc_tma<P2>, cudaFuncAttributeMaxDynamicSharedMemorySize, P2));
    rep("TMA 64K host->host", timeit([&] { zc_tma<P2><<<bytes / P2, 32, P2>>>(h1, h2, bytes); }), 2);
    constexpr int P3 = 4096;
    rep("TMA 4K host->host", timeit([&] { zc_tma<P3><<<bytes / P3, 32, P3>>>(h1, h2, bytes); }), 2);
  }
  cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
  {
    cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    auto both = [&](int U_blocks_div) {
      // two kernels at once: reader host->device on s1, writer device->host on s2
      zc_ldg<4><<<(n + 1023) / 1024, 256, 0, s1>>>((uint4*)h1, (uint4*)d1, n);
      zc_ldg<4><<<(n + 1023) / 1024, 256, 0, s2>>>((uint4*)d2, (uint4*)h2, n);
    };
    for (int w = 0; w < 2; w++) both(1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0, s1); cudaStreamWaitEvent(s2, e0, 0);
    for (int i = 0; i < 10; i++) both(1);
    cudaEventRecord(e1, s1); cudaEventRecord(e2, s2); cudaStreamWaitEvent(s1, e2, 0); cudaEventRecord(e1, s1);
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
    rep("two kernels: LDG h->d || STG d->h", ms / 10, 2);
    // kernel reads host, DMA writes host
    auto mix = [&]() {
      zc_ldg<4><<<(n + 1023) / 1024, 256, 0, s1>>>((uint4*)h1, (uint4*)d1, n);
      cudaMemcpyAsync(h2, d2, bytes, cudaMemcpyDeviceToHost, s2);
    };
    for (int w = 0; w < 2; w++) mix();
    cudaDeviceSynchronize();
    cudaEventRecord(e0, s1); cudaStreamWaitEvent(s2, e0, 0);
    for (int i = 0; i < 10; i++) mix();
    cudaEventRecord(e2, s2); cudaStreamWaitEvent(s1, e2, 0); cudaEventRecord(e1, s1);
    cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    rep("kernel LDG h->d || DMA d->h", ms / 10, 2);
    auto mix2 = [&]() {
      cudaMemcpyAsync(d1, h1, bytes, cudaMemcpyHostToDevice, s1);
      zc_ldg<4><<<(n + 1023) / 1024, 256, 0, s2>>>((uint4*)d2, (uint4*)h2, n);
    };
    for (int w = 0; w < 2; w++) mix2();
    cudaDeviceSynchronize();
    cudaEventRecord(e0, s1); cudaStreamWaitEvent(s2, e0, 0);
    for (int i = 0; i < 10; i++) mix2();
    cudaEventRecord(e2, s2); cudaStreamWaitEvent(s1, e2, 0); cudaEventRecord(e1, s1);
    cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    rep("DMA h->d || kernel STG d->h", ms / 10, 2);
  }
  {
    // the e2e pattern: per frame 4 pieces (rows under two cue regions, luma + chroma of 4K NV12),
    // DMA host->device of batch i+1 while a kernel stores batch i device->host
    const size_t piece[4] = {1382400, 552960, 691200, 276480};
    const int frames = 32; size_t per_frame = 0; for (size_t p : piece) per_frame += p;
    cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    auto go = [&]() {
      size_t off = 0;
      for (int f = 0; f < frames; f++) for (size_t p : piece) { cudaMemcpyAsync(d1 + off, h1 + off, p, cudaMemcpyHostToDevice, s1); off += p; }
      size_t nn = per_frame * frames / 16;
      zc_ldg<4><<<(nn + 1023) / 1024, 256, 0, s2>>>((uint4*)d2, (uint4*)h2, nn);
    };
    for (int w = 0; w < 2; w++) go();
    cudaDeviceSynchronize();
    cudaEventRecord(e0, s1); cudaStreamWaitEvent(s2, e0, 0);
    for (int i = 0; i < 10; i++) go();
    cudaEventRecord(e2, s2); cudaStreamWaitEvent(s1, e2, 0); cudaEventRecord(e1, s1);
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %8.3f ms per 32 frames -> %.0f frames/s, %.1f GB/s per direction\n", "frame pieces: DMA h->d || kernel STG d->h", ms / 10, frames / (ms / 10) * 1e3, per_frame * frames / (ms / 10) / 1e6);
    auto go2 = [&]() {   // zero-copy shape: one kernel reads and writes host
      size_t nn = per_frame * frames / 16;
      zc_ldg<4><<<(nn + 1023) / 1024, 256, 0, s2>>>((uint4*)h1, (uint4*)h1, nn);
    };
    for (int w = 0; w < 2; w++) go2();
    cudaDeviceSynchronize();
    cudaEventRecord(e0, s2);
    for (int i = 0; i < 10; i++) go2();
    cudaEventRecord(e1, s2); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %8.3f ms per 32 frames -> %.0f frames/s, %.1f GB/s per direction\n", "frame pieces: one kernel LDG+STG host", ms / 10, frames / (ms / 10) * 1e3, per_frame * frames / (ms / 10) / 1e6);
  }
  rep("DMA H2D + D2H concurrently", timeit([&] { cudaMemcpyAsync(d1, h1, bytes, cudaMemcpyHostToDevice, s1); cudaMemcpyAsync(h2, d2, bytes, cudaMemcpyDeviceToHost, s2); cudaStreamSynchronize(s1); cudaStreamSynchronize(s2); }, 5), 2);
  rep("DMA H2D only", timeit([&] { cudaMemcpyAsync(d1, h1, bytes, cudaMemcpyHostToDevice, 0); }, 5), 1);
  return 0;
}
