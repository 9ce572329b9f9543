// Development microbenchmark: does remote (NVLink) traffic issued from the LSU (cp.async / st.global) block the LOCAL
// loads of the same SM?  Each CTA streams `nl` local bytes and `nr` remote bytes through shared memory.
//   lsu : both streams with cp.async 16 B (LDGSTS)            tma : remote stream with cp.async.bulk (1536 B pieces),
//                                                                    local stream with cp.async 16 B
// Reported: time of local only, remote only, both - "both ~ max" means overlap, "both ~ sum" means serialisation.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
constexpr int TH = 128, PIECE = 1536, NPIECE = 6;      // a "tile" = 6 local + 6 remote pieces of 1536 B (like one z-step tile)
constexpr int TILE_B = PIECE * NPIECE;

__device__ __forceinline__ void cpa16(void* s, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(s)), "l"(g) : "memory");
}
// mode bit0: local stream on, bit1: remote stream on; use_tma: remote via bulk copies
__global__ void mix(const char* __restrict__ loc, const char* __restrict__ rem, double* sink, size_t tiles, int mode, int use_tma) {
  extern __shared__ __align__(128) char sm[];            // [2 stages][2 streams][TILE_B] + mbarriers
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 4 * TILE_B);
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(bar + b)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned ph[2] = {0, 0};
  double acc = 0.0;
  auto issue = [&](size_t t, int st) {
    char* sl = sm + (st * 2 + 0) * TILE_B; char* sr = sm + (st * 2 + 1) * TILE_B;
    if (mode & 1) for (int i = threadIdx.x; i < TILE_B / 16; i += TH) cpa16(sl + i * 16, loc + t * TILE_B + i * 16);
    if (mode & 2) {
      if (!use_tma) { for (int i = threadIdx.x; i < TILE_B / 16; i += TH) cpa16(sr + i * 16, rem + t * TILE_B + i * 16); }
      else if (threadIdx.x < NPIECE) {
        const unsigned ba = (unsigned)__cvta_generic_to_shared(bar + st), sa = (unsigned)__cvta_generic_to_shared(sr + threadIdx.x * PIECE);
        if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"(TILE_B) : "memory");
        __syncwarp(0x3f);
        asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa), "l"(rem + t * TILE_B + threadIdx.x * PIECE), "r"(PIECE), "r"(ba) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  size_t t = blockIdx.x;
  int st = 0;
  if (t < tiles) issue(t, 0);
  for (; t < tiles; t += gridDim.x) {
    const size_t tn = t + gridDim.x;
    if (tn < tiles) issue(tn, st ^ 1); else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    if ((mode & 2) && use_tma) {
      const unsigned ba = (unsigned)__cvta_generic_to_shared(bar + st);
      unsigned ok = 0;
      while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(ba), "r"(ph[st]) : "memory");
      ph[st] ^= 1;
    }
    __syncthreads();
    const double* d = reinterpret_cast<const double*>(sm + st * 2 * TILE_B);
    for (int i = threadIdx.x; i < 2 * TILE_B / 8; i += TH * 8) acc += d[i];   // touch a little, like a consumer would
    __syncthreads();
    st ^= 1;
  }
  if (acc == 1.2345) *sink = acc;
}
int main() {
  int nd = 0; CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
  const size_t tiles = 8192, bytes = tiles * TILE_B;     // 75 MB per stream
  char *loc, *rem; double* sink;
  CK(cudaSetDevice(1)); CK(cudaMalloc(&rem, bytes)); CK(cudaMemset(rem, 1, bytes));
  CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0)); CK(cudaMalloc(&loc, bytes)); CK(cudaMemset(loc, 2, bytes)); CK(cudaMalloc(&sink, 8));
  const int smem = 4 * TILE_B + 64;
  CK(cudaFuncSetAttribute(mix, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int cps = 2; cps <= 4; cps += 2)
    for (int use_tma = 0; use_tma < 2; ++use_tma)
      for (int mode = 1; mode <= 3; ++mode) {
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
          cudaEventRecord(e0);
          mix<<<148 * cps, TH, smem>>>(loc, rem, sink, tiles, mode, use_tma);
          cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
          float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        CK(cudaGetLastError());
        printf("%d CTA/SM %s %-11s %7.1f us\n", cps, use_tma ? "tma" : "lsu", mode == 1 ? "local" : mode == 2 ? "remote" : "both", best * 1e3);
      }
  return 0;
}
