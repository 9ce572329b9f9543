// Development microbenchmark: how fast can ONE B200 move data to / from a peer's HBM over NVLink from inside a kernel?
//   push16  : st.global.v2.f64 per thread, coalesced            pull16  : ld.global 16 B per thread (unroll 8)
//   pushrun : 16 B stores in 256 B runs scattered 4 KB apart    pullcpa : cp.async 16 B into shared memory, 4 stages
//   pushtma : cp.async.bulk shared -> peer global, 2 KB pieces  pulltma : cp.async.bulk peer global -> shared + mbarrier
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o p2p_bw p2p_bw.cu ; needs 2 GPUs with peer access.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <thread>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void push16(double2* __restrict__ dst, const double2* __restrict__ src, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void pushrun(double2* __restrict__ dst, const double2* __restrict__ src, size_t n) {
  // element i -> run r = i / 16 (256 B), runs permuted: dst run = (r * 257) % nruns
  const size_t nruns = n / 16;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / 16, e = i % 16;
    dst[((r * 257) % nruns) * 16 + e] = src[i];
  }
}
__global__ void pull16(double2* __restrict__ dst, const double2* __restrict__ src, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; i + 7 * stride < n; i += 8 * stride) {
    double2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = src[i + u * stride];
#pragma unroll
    for (int u = 0; u < 8; ++u) dst[i + u * stride] = v[u];
  }
  for (; i < n; i += stride) dst[i] = src[i];
}
__global__ void pullcpa(double2* __restrict__ dst, const double2* __restrict__ src, size_t n) {
  extern __shared__ double2 sm[];   // [4][blockDim]
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
  int issued = 0, done = 0;
  size_t ii = i;
  for (int s = 0; s < 3; ++s) {
    if (ii < n) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + (unsigned)((s * blockDim.x + threadIdx.x) * 16)), "l"(src + ii));
    asm volatile("cp.async.commit_group;");
    ii += stride; ++issued;
  }
  for (; i < n; i += stride) {
    if (ii < n) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + (unsigned)(((issued & 3) * blockDim.x + threadIdx.x) * 16)), "l"(src + ii));
    asm volatile("cp.async.commit_group;");
    ii += stride; ++issued;
    asm volatile("cp.async.wait_group 3;" ::: "memory");
    dst[i] = sm[(done & 3) * blockDim.x + threadIdx.x];
    ++done;
  }
}
// TMA bulk: each CTA loops over chunks of CH bytes
template <int CH> __global__ void pushtma(char* __restrict__ dst, const char* __restrict__ src, size_t bytes) {
  extern __shared__ __align__(128) char smc[];   // [2][CH]
  const size_t nch = bytes / CH;
  int buf = 0;
  for (size_t c = blockIdx.x; c < nch; c += gridDim.x) {
    // wait until the bulk store that last used this buffer has finished reading it
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncthreads();
    const double2* s = reinterpret_cast<const double2*>(src + c * CH);
    double2* d = reinterpret_cast<double2*>(smc + buf * CH);
    for (int i = threadIdx.x; i < CH / 16; i += blockDim.x) d[i] = s[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned sa = (unsigned)__cvta_generic_to_shared(smc + buf * CH);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * CH), "r"(sa), "r"(CH) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    buf ^= 1;
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
template <int CH> __global__ void pulltma(char* __restrict__ dst, const char* __restrict__ src, size_t bytes) {
  extern __shared__ __align__(128) char smc[];   // [2][CH] + 2 mbarriers
  uint64_t* bar = reinterpret_cast<uint64_t*>(smc + 2 * CH);
  const size_t nch = bytes / CH;
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(bar + b)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](size_t c, int b) {
    const unsigned ba = (unsigned)__cvta_generic_to_shared(bar + b), sa = (unsigned)__cvta_generic_to_shared(smc + b * CH);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"(CH) : "memory");
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa), "l"(src + c * CH), "r"(CH), "r"(ba) : "memory");
  };
  size_t c = blockIdx.x;
  if (threadIdx.x == 0 && c < nch) issue(c, 0);
  int b = 0; unsigned ph[2] = {0, 0};
  for (; c < nch; c += gridDim.x) {
    const size_t cn = c + gridDim.x;
    if (threadIdx.x == 0 && cn < nch) issue(cn, b ^ 1);
    const unsigned ba = (unsigned)__cvta_generic_to_shared(bar + b);
    unsigned ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(ba), "r"(ph[b]) : "memory");
    ph[b] ^= 1;
    const double2* s = reinterpret_cast<const double2*>(smc + b * CH);
    double2* d = reinterpret_cast<double2*>(dst + c * CH);
    for (int i = threadIdx.x; i < CH / 16; i += blockDim.x) d[i] = s[i];
    __syncthreads();
    b ^= 1;
  }
}

int main() {
  int nd = 0; CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
  const size_t bytes = 64ull << 20, n = bytes / 16;
  double2 *loc, *rem, *loc2;
  CK(cudaSetDevice(1)); CK(cudaMalloc(&rem, bytes)); CK(cudaMemset(rem, 1, bytes));
  CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0));
  CK(cudaMalloc(&loc, bytes)); CK(cudaMalloc(&loc2, bytes)); CK(cudaMemset(loc, 2, bytes));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  constexpr int CH = 4096;
  CK(cudaFuncSetAttribute(pulltma<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * CH + 64));
  const int grids[] = {148, 148 * 2, 148 * 4, 148 * 8};
  for (int gi = 0; gi < 4; ++gi) {
    const int g = grids[gi];
    for (int k = 0; k < 8; ++k) {
      float best = 1e9f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        switch (k) {
          case 0: push16<<<g, 256>>>(rem, loc, n); break;
          case 1: pushrun<<<g, 256>>>(rem, loc, n); break;
          case 2: pull16<<<g, 256>>>(loc2, rem, n); break;
          case 3: pullcpa<<<g, 256, 4 * 256 * 16>>>(loc2, rem, n); break;
          case 4: pushtma<CH><<<g, 128, 2 * CH>>>((char*)rem, (const char*)loc, bytes); break;
          case 5: pulltma<CH><<<g, 128, 2 * CH + 64>>>((char*)loc2, (const char*)rem, bytes); break;
          case 6: push16<<<g, 256>>>(loc2, loc, n); break;      // local reference
          case 7: pushtma<CH><<<g, 128, 2 * CH>>>((char*)loc2, (const char*)loc, bytes); break;
        }
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      CK(cudaGetLastError());
      const char* nm[] = {"push16", "pushrun", "pull16", "pullcpa", "pushtma", "pulltma", "local16", "localtma"};
      printf("grid %4d %-8s %8.1f GB/s (%.1f us for 64 MiB)\n", g, nm[k], bytes / (best * 1e-3) / 1e9, best * 1e3);
    }
  }
  // ---- bidirectional: both GPUs move data at the same time (the all-to-all situation) ----
  double2 *loc1, *rem0, *loc1b;   // buffers of the mirrored direction
  CK(cudaSetDevice(1)); CK(cudaDeviceEnablePeerAccess(0, 0)); CK(cudaMalloc(&loc1, bytes)); CK(cudaMalloc(&loc1b, bytes)); CK(cudaMemset(loc1, 3, bytes));
  CK(cudaSetDevice(0)); CK(cudaMalloc(&rem0, bytes)); CK(cudaMemset(rem0, 4, bytes));
  for (int k = 0; k < 4; ++k) {
    float best[2] = {1e9f, 1e9f};
    for (int rep = 0; rep < 4; ++rep) {
      float ms[2];
      auto run = [&](int dev) {
        cudaSetDevice(dev);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        double2* L = dev == 0 ? loc : loc1; double2* L2 = dev == 0 ? loc2 : loc1b; double2* R = dev == 0 ? rem : rem0;
        cudaDeviceSynchronize();
        cudaEventRecord(a);
        switch (k) {
          case 0: push16<<<592, 256>>>(R, L, n); break;
          case 1: pushrun<<<592, 256>>>(R, L, n); break;
          case 2: pull16<<<592, 256>>>(L2, R, n); break;
          case 3: pullcpa<<<592, 256, 4 * 256 * 16>>>(L2, R, n); break;
        }
        cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms[dev], a, b);
      };
      std::thread t(run, 1); run(0); t.join();
      for (int d = 0; d < 2; ++d) if (ms[d] < best[d]) best[d] = ms[d];
    }
    const char* nm[] = {"push16", "pushrun", "pull16", "pullcpa"};
    printf("bidirectional %-8s GPU0 %7.1f GB/s  GPU1 %7.1f GB/s\n", nm[k], bytes / (best[0] * 1e-3) / 1e9, bytes / (best[1] * 1e-3) / 1e9);
  }
  return 0;
}
