// tools/copybench.cu -- what can a frame-sized streaming copy reach on this B200?
// Scratch micro-benchmark used to pick the structure of the blend kernel's copy path
// (DESIGN.md "Kernel structure"). Not part of the product library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/copybench tools/copybench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint4 ld_na(const uint4* p) {
  uint4 r; asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x),"=r"(r.y),"=r"(r.z),"=r"(r.w) : "l"(p)); return r; }
__device__ __forceinline__ uint4 ld_nc(const uint4* p) {
  uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x),"=r"(r.y),"=r"(r.z),"=r"(r.w) : "l"(p)); return r; }
__device__ __forceinline__ void st_na(uint4* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p),"r"(v.x),"r"(v.y),"r"(v.z),"r"(v.w) : "memory"); }
__device__ __forceinline__ void st_cs(uint4* p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p),"r"(v.x),"r"(v.y),"r"(v.z),"r"(v.w) : "memory"); }

template <int U, int MODE>
__global__ void __launch_bounds__(256) copy_chunk(const uint4* __restrict__ s, uint4* __restrict__ d, size_t n) {
  size_t base = (size_t)blockIdx.x * (256 * U) + threadIdx.x;
  uint4 v[U];
#pragma unroll
  for (int k = 0; k < U; k++) { size_t i = base + (size_t)k * 256; if (i < n) v[k] = MODE == 1 ? ld_nc(s + i) : (MODE == 2 ? s[i] : ld_na(s + i)); }
#pragma unroll
  for (int k = 0; k < U; k++) { size_t i = base + (size_t)k * 256; if (i < n) { if (MODE == 3) st_cs(d + i, v[k]); else if (MODE == 2) d[i] = v[k]; else st_na(d + i, v[k]); } }
}

// occupancy-limited copy: dynamic smem only there to cap CTAs/SM
template <int U>
__global__ void __launch_bounds__(256) copy_occ(const uint4* __restrict__ s, uint4* __restrict__ d, size_t n) {
  extern __shared__ uint8_t dummy[];
  size_t base = (size_t)blockIdx.x * (256 * U) + threadIdx.x;
  uint4 v[U];
#pragma unroll
  for (int k = 0; k < U; k++) { size_t i = base + (size_t)k * 256; if (i < n) v[k] = ld_na(s + i); }
#pragma unroll
  for (int k = 0; k < U; k++) { size_t i = base + (size_t)k * 256; if (i < n) st_na(d + i, v[k]); }
  if (n == 1) dummy[threadIdx.x] = 0;
}

// copy with the blend kernel's prologue: table search (barrier) + dependent job-field loads
struct FakeJob { const uint4* s; uint4* d; unsigned begin; unsigned pad; };
template <int U>
__global__ void __launch_bounds__(256) copy_prologue(const FakeJob* __restrict__ jobs, const unsigned* __restrict__ begins, int n_jobs, size_t n) {
  extern __shared__ uint8_t dummy[];
  unsigned chunk = blockIdx.x;
  int cnt = 0;
  for (int b = 0; b < n_jobs; b += 256) { int j = b + threadIdx.x; cnt += __syncthreads_count(j < n_jobs && __ldg(begins + j) <= chunk); }
  const FakeJob* job = jobs + (cnt - 1);
  const uint4* s = (const uint4*)__ldg((const unsigned long long*)&job->s);
  uint4* d = (uint4*)__ldg((const unsigned long long*)&job->d);
  unsigned local = chunk - __ldg(&job->begin);
  size_t base = (size_t)local * (256 * U) + threadIdx.x;
  uint4 v[U];
#pragma unroll
  for (int k = 0; k < U; k++) { size_t i = base + (size_t)k * 256; v[k] = ld_na(s + i); }
#pragma unroll
  for (int k = 0; k < U; k++) { size_t i = base + (size_t)k * 256; st_na(d + i, v[k]); }
  if (n == 1) dummy[threadIdx.x] = 0;
}

template <int U>
__global__ void __launch_bounds__(256) copy_persist(const uint4* __restrict__ s, uint4* __restrict__ d, size_t n) {
  const size_t chunks = (n + 256 * U - 1) / (256 * U);
  for (size_t c = blockIdx.x; c < chunks; c += gridDim.x) {
    size_t base = c * (256 * U) + threadIdx.x;
    uint4 v[U];
#pragma unroll
    for (int k = 0; k < U; k++) { size_t i = base + (size_t)k * 256; if (i < n) v[k] = ld_na(s + i); }
#pragma unroll
    for (int k = 0; k < U; k++) { size_t i = base + (size_t)k * 256; if (i < n) st_na(d + i, v[k]); }
  }
}

// software-pipelined persistent: loads of chunk c+G are issued before the stores of chunk c
template <int U>
__global__ void __launch_bounds__(256) copy_persist_pipe(const uint4* __restrict__ s, uint4* __restrict__ d, size_t n) {
  const size_t chunks = (n + 256 * U - 1) / (256 * U);
  size_t c = blockIdx.x;
  if (c >= chunks) return;
  uint4 cur[U], nxt[U];
  {
    size_t base = c * (256 * U) + threadIdx.x;
#pragma unroll
    for (int k = 0; k < U; k++) { size_t i = base + (size_t)k * 256; if (i < n) cur[k] = ld_na(s + i); }
  }
  for (; c < chunks; c += gridDim.x) {
    size_t cn = c + gridDim.x;
    if (cn < chunks) {
      size_t base = cn * (256 * U) + threadIdx.x;
#pragma unroll
      for (int k = 0; k < U; k++) { size_t i = base + (size_t)k * 256; if (i < n) nxt[k] = ld_na(s + i); }
    }
    size_t base = c * (256 * U) + threadIdx.x;
#pragma unroll
    for (int k = 0; k < U; k++) { size_t i = base + (size_t)k * 256; if (i < n) st_na(d + i, cur[k]); }
#pragma unroll
    for (int k = 0; k < U; k++) cur[k] = nxt[k];
  }
}

// TMA bulk copy ring: one thread per CTA moves STAGE_BYTES pieces global -> smem -> global
template <int STAGES, int STAGE_BYTES>
__global__ void __launch_bounds__(32) copy_bulk(const uint8_t* __restrict__ s, uint8_t* __restrict__ d, size_t bytes) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bar[STAGES];
  const size_t pieces = bytes / STAGE_BYTES;
  if (threadIdx.x != 0) return;
  for (int i = 0; i < STAGES; i++) {
    unsigned a = (unsigned)__cvta_generic_to_shared(&bar[i]);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(a));
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  size_t p = blockIdx.x;
  unsigned phase_bits = 0;
  // prologue: fill the ring
  size_t issue = p;
  int n_issued = 0;
  for (int st = 0; st < STAGES && issue < pieces; st++, issue += gridDim.x, n_issued++) {
    unsigned a = (unsigned)__cvta_generic_to_shared(&bar[st]);
    unsigned sm = (unsigned)__cvta_generic_to_shared(smem + (size_t)st * STAGE_BYTES);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(a), "r"(STAGE_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(sm), "l"(s + issue * STAGE_BYTES), "r"(STAGE_BYTES), "r"(a) : "memory");
  }
  int st = 0;
  for (; p < pieces; p += gridDim.x) {
    unsigned a = (unsigned)__cvta_generic_to_shared(&bar[st]);
    unsigned sm = (unsigned)__cvta_generic_to_shared(smem + (size_t)st * STAGE_BYTES);
    unsigned ph = (phase_bits >> st) & 1u;
    asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" :: "r"(a), "r"(ph) : "memory");
    phase_bits ^= 1u << st;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(d + p * STAGE_BYTES), "r"(sm), "r"(STAGE_BYTES) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    // refill this stage once its store has read the smem
    if (issue < pieces) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(a), "r"(STAGE_BYTES) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   :: "r"(sm), "l"(s + issue * STAGE_BYTES), "r"(STAGE_BYTES), "r"(a) : "memory");
      issue += gridDim.x;
    }
    st = (st + 1) % STAGES;
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
float timeit(F f, int iters = 20) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; i++) f();
  cudaEventRecord(a);
  for (int i = 0; i < iters; i++) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  CK(cudaGetLastError());
  return ms / iters;
}

int main(int argc, char** argv) {
  size_t bytes = (argc > 1 ? atol(argv[1]) : 398) * 1000000ull / 65536 * 65536;
  uint8_t *s, *d;
  CK(cudaMalloc(&s, bytes)); CK(cudaMalloc(&d, bytes));
  CK(cudaMemset(s, 1, bytes)); CK(cudaMemset(d, 2, bytes));
  size_t n = bytes / 16;
  auto rep = [&](const char* name, float ms) { printf("%-34s %8.4f ms  %8.1f GB/s (read+write)\n", name, ms, 2.0 * bytes / ms / 1e6); };
  rep("cudaMemcpyAsync D2D", timeit([&] { cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, 0); }));
  rep("chunk U4 ld.na/st.na", timeit([&] { copy_chunk<4, 0><<<(n + 1023) / 1024, 256>>>((uint4*)s, (uint4*)d, n); }));
  rep("chunk U4 ld.nc/st.na", timeit([&] { copy_chunk<4, 1><<<(n + 1023) / 1024, 256>>>((uint4*)s, (uint4*)d, n); }));
  rep("chunk U4 plain", timeit([&] { copy_chunk<4, 2><<<(n + 1023) / 1024, 256>>>((uint4*)s, (uint4*)d, n); }));
  rep("chunk U4 ld.na/st.cs", timeit([&] { copy_chunk<4, 3><<<(n + 1023) / 1024, 256>>>((uint4*)s, (uint4*)d, n); }));
  rep("chunk U8 ld.na/st.na", timeit([&] { copy_chunk<8, 0><<<(n + 2047) / 2048, 256>>>((uint4*)s, (uint4*)d, n); }));
  rep("chunk U2 ld.na/st.na", timeit([&] { copy_chunk<2, 0><<<(n + 511) / 512, 256>>>((uint4*)s, (uint4*)d, n); }));
  rep("chunk U1 ld.na/st.na", timeit([&] { copy_chunk<1, 0><<<(n + 255) / 256, 256>>>((uint4*)s, (uint4*)d, n); }));
  for (int occ : {2, 3, 4, 5, 6, 8}) {
    int smem = 227 * 1024 / occ - 1024; smem = smem / 1024 * 1024;
    CK(cudaFuncSetAttribute(copy_occ<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(copy_occ<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    char nm[64]; snprintf(nm, sizeof nm, "chunk U4 capped %d CTA/SM", occ);
    rep(nm, timeit([&] { copy_occ<4><<<(n + 1023) / 1024, 256, smem>>>((uint4*)s, (uint4*)d, n); }));
    snprintf(nm, sizeof nm, "chunk U2 capped %d CTA/SM", occ);
    rep(nm, timeit([&] { copy_occ<2><<<(n + 511) / 512, 256, smem>>>((uint4*)s, (uint4*)d, n); }));
  }
  {
    // 320 jobs of equal size, like 32 frames x 10 bands
    int n_jobs = 320; size_t chunks = n / 1024, per = chunks / n_jobs;
    std::vector<FakeJob> hj(n_jobs); std::vector<unsigned> hb(n_jobs);
    for (int j = 0; j < n_jobs; j++) { hj[j].s = (const uint4*)s + (size_t)j * per * 1024; hj[j].d = (uint4*)d + (size_t)j * per * 1024; hj[j].begin = hb[j] = (unsigned)(j * per); }
    FakeJob* dj; unsigned* db; CK(cudaMalloc(&dj, n_jobs * sizeof(FakeJob))); CK(cudaMalloc(&db, n_jobs * 4));
    CK(cudaMemcpy(dj, hj.data(), n_jobs * sizeof(FakeJob), cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, hb.data(), n_jobs * 4, cudaMemcpyHostToDevice));
    for (int occ : {4, 8}) {
      int smem = 227 * 1024 / occ - 1024; smem = smem / 1024 * 1024;
      CK(cudaFuncSetAttribute(copy_prologue<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      char nm[64]; snprintf(nm, sizeof nm, "chunk U4 +prologue %d CTA/SM", occ);
      float ms = timeit([&] { copy_prologue<4><<<(unsigned)(per * n_jobs), 256, smem>>>(dj, db, n_jobs, n); });
      printf("%-34s %8.4f ms  %8.1f GB/s (read+write)\n", nm, ms, 2.0 * per * n_jobs * 16384 / ms / 1e6);
    }
  }
  for (int per : {2, 4}) {
    char nm[64]; snprintf(nm, sizeof nm, "persist U4 %d CTA/SM", per);
    rep(nm, timeit([&] { copy_persist<4><<<148 * per, 256>>>((uint4*)s, (uint4*)d, n); }));
    snprintf(nm, sizeof nm, "persist U8 %d CTA/SM", per);
    rep(nm, timeit([&] { copy_persist<8><<<148 * per, 256>>>((uint4*)s, (uint4*)d, n); }));
    snprintf(nm, sizeof nm, "persist-pipe U4 %d CTA/SM", per);
    rep(nm, timeit([&] { copy_persist_pipe<4><<<148 * per, 256>>>((uint4*)s, (uint4*)d, n); }));
  }
  {
    constexpr int SB = 16384;
    CK(cudaFuncSetAttribute(copy_bulk<4, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * SB));
    CK(cudaFuncSetAttribute(copy_bulk<8, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * SB));
    CK(cudaFuncSetAttribute(copy_bulk<2, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * SB));
    for (int per : {1, 2, 3}) {
      char nm[64];
      snprintf(nm, sizeof nm, "TMA bulk 4x16K %d CTA/SM", per);
      rep(nm, timeit([&] { copy_bulk<4, SB><<<148 * per, 32, 4 * SB>>>(s, d, bytes); }));
      snprintf(nm, sizeof nm, "TMA bulk 2x16K %d CTA/SM", per);
      rep(nm, timeit([&] { copy_bulk<2, SB><<<148 * per, 32, 2 * SB>>>(s, d, bytes); }));
    }
    rep("TMA bulk 8x16K 1 CTA/SM", timeit([&] { copy_bulk<8, SB><<<148, 32, 8 * SB>>>(s, d, bytes); }));
    constexpr int SB2 = 32768;
    CK(cudaFuncSetAttribute(copy_bulk<4, SB2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * SB2));
    rep("TMA bulk 4x32K 1 CTA/SM", timeit([&] { copy_bulk<4, SB2><<<148, 32, 4 * SB2>>>(s, d, bytes); }));
    constexpr int SB3 = 8192;
    CK(cudaFuncSetAttribute(copy_bulk<8, SB3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * SB3));
    rep("TMA bulk 8x8K 2 CTA/SM", timeit([&] { copy_bulk<8, SB3><<<148 * 2, 32, 8 * SB3>>>(s, d, bytes); }));
  }
  // verify last copy
  std::vector<uint8_t> h(1 << 20);
  CK(cudaMemcpy(h.data(), d + bytes - h.size(), h.size(), cudaMemcpyDeviceToHost));
  for (auto v : h) if (v != 1) { printf("VERIFY FAILED\n"); return 1; }
  printf("verify ok\n");
  return 0;
}
