// Development microbenchmark: cross-GPU flag ping-pong latency over NVLink and the price of system-scope fences.
// GPU0 writes a flag word in GPU1's memory and spins on its own; GPU1 answers.  One CTA, one thread each.
//   mode 0: volatile store / volatile spin, no fences
//   mode 1: __threadfence_system() before the store and after the spin
//   mode 2: st.release.sys / ld.acquire.sys
//   mode 3: __threadfence() (gpu scope) before the store, none after
// plus: cost of __threadfence_system() alone after 64 outstanding local 16-byte stores per thread of a 256-thread CTA.
#include <cstdio>
#include <cuda_runtime.h>
#include <thread>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
typedef unsigned long long ull;

__device__ __forceinline__ void put(ull* p, ull v, int mode) {
  if (mode == 1) __threadfence_system();
  if (mode == 3) __threadfence();
  if (mode == 2) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
  else *((volatile ull*)p) = v;
}
__device__ __forceinline__ void wait_for(const ull* p, ull v, int mode) {
  if (mode == 2) {
    ull x;
    do { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(x) : "l"(p) : "memory"); } while (x < v);
  } else {
    while (*((volatile const ull*)p) < v) {}
    if (mode == 1) __threadfence_system();
  }
}
__global__ void pingpong(ull* mine, ull* theirs, int first, int n, int mode, long long* cycles) {
  const long long t0 = clock64();
  for (int i = 1; i <= n; ++i) {
    if (first) { put(theirs, (ull)i, mode); wait_for(mine, (ull)i, mode); }
    else { wait_for(mine, (ull)i, mode); put(theirs, (ull)i, mode); }
  }
  *cycles = clock64() - t0;
}
__global__ void fence_cost(double2* buf, long long* cycles, int kind) {
  double2 v = make_double2(threadIdx.x, blockIdx.x);
  for (int i = 0; i < 64; ++i) buf[((size_t)blockIdx.x * 64 + i) * blockDim.x + threadIdx.x] = v;
  __syncthreads();
  const long long t0 = clock64();
  if (threadIdx.x == 0) { if (kind == 0) __threadfence_system(); else if (kind == 1) __threadfence(); }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
int main() {
  int nd = 0; CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
  ull *f0, *f1; long long *c0, *c1; double2* buf;
  CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0)); CK(cudaMalloc(&f0, 64)); CK(cudaMalloc(&c0, 8)); CK(cudaMalloc(&buf, 148ull * 4 * 64 * 256 * 16));
  CK(cudaSetDevice(1)); CK(cudaDeviceEnablePeerAccess(0, 0)); CK(cudaMalloc(&f1, 64)); CK(cudaMalloc(&c1, 8));
  const int n = 2000;
  for (int mode = 0; mode < 4; ++mode) {
    CK(cudaSetDevice(0)); CK(cudaMemset(f0, 0, 64)); CK(cudaSetDevice(1)); CK(cudaMemset(f1, 0, 64));
    CK(cudaDeviceSynchronize()); CK(cudaSetDevice(0)); CK(cudaDeviceSynchronize());
    std::thread t([&] { cudaSetDevice(1); pingpong<<<1, 1>>>(f1, f0, 0, n, mode, c1); cudaDeviceSynchronize(); });
    cudaSetDevice(0); pingpong<<<1, 1>>>(f0, f1, 1, n, mode, c0); CK(cudaDeviceSynchronize());
    t.join();
    long long cyc; CK(cudaMemcpy(&cyc, c0, 8, cudaMemcpyDeviceToHost));
    printf("mode %d: round trip %.2f us (%.0f cycles), one-way ~%.2f us\n", mode, cyc / 1.965e3 / n, (double)cyc / n, cyc / 1.965e3 / n / 2);
  }
  CK(cudaSetDevice(0));
  for (int kind = 0; kind < 3; ++kind) {
    fence_cost<<<148 * 4, 256>>>(buf, c0, kind); CK(cudaDeviceSynchronize());
    fence_cost<<<148 * 4, 256>>>(buf, c0, kind); CK(cudaDeviceSynchronize());
    long long cyc; CK(cudaMemcpy(&cyc, c0, 8, cudaMemcpyDeviceToHost));
    printf("fence after 1 MiB of stores per CTA-wave, kind %d (%s): %.2f us\n", kind, kind == 0 ? "system" : kind == 1 ? "gpu" : "none", cyc / 1.965e3);
  }
  return 0;
}
