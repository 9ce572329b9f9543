// C ABI of libsmo_b200.so (see include/smo_b200.h).  Host-side plans and launch sequences; all arithmetic
// lives in the kernels of fft_pass.cuh, xpass.cuh, kd_epilogue.cuh, sh23.cuh and reduce.cuh.
//
// Built two ways from this one source:
//   nvcc -gencode arch=compute_100a,code=sm_100a ...            -> libsmo_b200.so (the product)
//   g++ -x c++ -DSMO_EMUL ...                                   -> tests/emul/_build/libsmo_emul.so, a host
//        emulation of the same kernel bodies used ONLY by the CPU test-suite to check index logic without a GPU.
#include "../../include/smo_b200.h"
#include "fft_pass.cuh"
#include "xpass.cuh"
#include "xpass_half.cuh"
#include "kd_epilogue.cuh"
#include "zstep.cuh"
#include "sh23.cuh"
#include "reduce.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstdarg>
#include <string>
#include <vector>
#include <atomic>

#if !defined(SMO_EMUL) && defined(SMO_WITH_NCCL)
#include <nccl.h>
#include <dlfcn.h>
// NCCL is bound at run time (dlopen) so that the library shares whatever libnccl.so.2 the host program (e.g.
// PyTorch) has already loaded instead of pulling in a second copy at link time.
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
  bool ok;
};
static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    api.ok = false;
    void* so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!so) so = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (so) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(so, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(so, "ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(so, "ncclCommDestroy");
      api.GroupStart = (decltype(api.GroupStart))dlsym(so, "ncclGroupStart");
      api.GroupEnd = (decltype(api.GroupEnd))dlsym(so, "ncclGroupEnd");
      api.Send = (decltype(api.Send))dlsym(so, "ncclSend");
      api.Recv = (decltype(api.Recv))dlsym(so, "ncclRecv");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(so, "ncclGetErrorString");
      api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Send &&
               api.Recv && api.GetErrorString;
    }
  }
  return api.ok ? &api : nullptr;
}
#endif

using namespace smo;

// ------------------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};
// set by the time loops of a single-rank handle (PdlScope): launches carry the programmatic-stream-serialisation attribute
static thread_local int t_pdl = 0;
struct PdlScope {
  int old;
  explicit PdlScope(int on) : old(t_pdl) { t_pdl = on; }
  ~PdlScope() { t_pdl = old; }
};
#ifndef SMO_BULK_PUSH_DEFAULT
#define SMO_BULK_PUSH_DEFAULT 0
#endif
#ifndef SMO_PDL_DEFAULT
#define SMO_PDL_DEFAULT -1
#endif

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define SMO_E_ARG -1
#define SMO_E_CUDA -2
#define SMO_E_UNSUPPORTED -3
#define SMO_E_STATE -4
#define SMO_E_COMM -5

// ------------------------------------------------------------------------------------------------------------
// runtime abstraction (CUDA or host emulation)
// ------------------------------------------------------------------------------------------------------------
#if defined(SMO_EMUL)
typedef void* rt_stream;
static int rt_malloc(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : fail(SMO_E_CUDA, "calloc(%zu) failed", n); }
static void rt_free(void* p) { free(p); }
static int rt_h2d(void* d, const void* s, size_t n, rt_stream) { memcpy(d, s, n); return 0; }
static int rt_d2h(void* d, const void* s, size_t n, rt_stream) { memcpy(d, s, n); return 0; }
static int rt_d2d(void* d, const void* s, size_t n, rt_stream) { memmove(d, s, n); return 0; }
static int rt_copy2d(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, int, rt_stream) {
  for (size_t r = 0; r < h; ++r) memcpy((char*)d + r * dp, (const char*)s + r * sp, w);
  return 0;
}
static int rt_memset(void* d, int v, size_t n, rt_stream) { memset(d, v, n); return 0; }
static int rt_sync(rt_stream) { return 0; }
static int rt_check(const char*) { return 0; }
template <class K> static int grid_for(int nwork) { return nwork < 3 ? nwork : 3; }
template <class K> static int launch(const typename K::Params& p, rt_stream, int /*waves*/ = 1) {
  if (p.nwork <= 0) return 0;
  const int grid = p.nwork < 3 ? p.nwork : 3;   // > 1 work item per CTA exercises the persistent loop
  emul_kernel<K>(grid, K::SMEM, p);
  g_launches++;
  return 0;
}
#else
typedef cudaStream_t rt_stream;
#define CUDA_TRY(x)                                                                              \
  do {                                                                                           \
    cudaError_t e_ = (x);                                                                        \
    if (e_ != cudaSuccess) return fail(SMO_E_CUDA, "%s failed: %s", #x, cudaGetErrorString(e_)); \
  } while (0)
static int rt_malloc(void** p, size_t n) {
  CUDA_TRY(cudaMalloc(p, n ? n : 1));
  CUDA_TRY(cudaMemset(*p, 0, n ? n : 1));
  return 0;
}
static void rt_free(void* p) { if (p) cudaFree(p); }
static int rt_h2d(void* d, const void* s, size_t n, rt_stream st) { CUDA_TRY(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, st)); return 0; }
static int rt_d2h(void* d, const void* s, size_t n, rt_stream st) { CUDA_TRY(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, st)); return 0; }
static int rt_d2d(void* d, const void* s, size_t n, rt_stream st) { CUDA_TRY(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, st)); return 0; }
// kind: 0 = host->device, 1 = device->host
static int rt_copy2d(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, int kind, rt_stream st) {
  CUDA_TRY(cudaMemcpy2DAsync(d, dp, s, sp, w, h, kind == 0 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, st));
  return 0;
}
static int rt_memset(void* d, int v, size_t n, rt_stream st) { CUDA_TRY(cudaMemsetAsync(d, v, n, st)); return 0; }
static int rt_sync(rt_stream st) { CUDA_TRY(cudaStreamSynchronize(st)); return 0; }
static int rt_check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(SMO_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}
// Launch configuration is cached PER DEVICE: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device attribute, so a
// second Domain on another GPU of the same process must set it again.
constexpr int SMO_MAX_DEVICES = 64;
static int cur_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < SMO_MAX_DEVICES) ? dev : 0;
}
static int num_sms() {
  static int v[SMO_MAX_DEVICES] = {0};
  const int dev = cur_device();
  if (v[dev] == 0) {
    cudaDeviceGetAttribute(&v[dev], cudaDevAttrMultiProcessorCount, dev);
    if (v[dev] <= 0) v[dev] = 148;
  }
  return v[dev];
}
template <class K> struct LaunchCfg {
  static int blocks_per_sm() {
    static int v[SMO_MAX_DEVICES] = {0};
    const int dev = cur_device();
    if (v[dev] == 0) {
      if (K::SMEM > 48 * 1024)
        cudaFuncSetAttribute(smo_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM);
      // ask for the largest shared-memory carve-out: the occupancy query below assumes it, but without the hint the driver may
      // pick a smaller one for kernels with modest per-CTA needs (SH23: 26 KB x 8 CTAs) - then fewer CTAs are resident than the
      // grid was sized for and the persistent loop runs a second, partial wave (seen as launches that take 2x, r2l / r2m)
      cudaFuncSetAttribute(smo_kernel<K>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
      int nb = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, smo_kernel<K>, K::THREADS, K::SMEM);
      v[dev] = nb > 0 ? nb : 1;
    }
    return v[dev];
  }
};
// persistent-style grid: at most one resident wave (a multiple of the SM count), CTAs loop over work items
template <class K> static int grid_for(int nwork) {
  const long long cap = (long long)num_sms() * LaunchCfg<K>::blocks_per_sm();
  return (int)(nwork < cap ? nwork : cap);
}
// waves > 1: the grid is shrunk so that every CTA handles about `waves` work items one after the other (kernels that push
// their results to peer GPUs: the remote stores of a CTA's first item drain while it transforms the next one)
template <class K> static int launch(const typename K::Params& p, rt_stream st, int waves = 1) {
  if (p.nwork <= 0) return 0;
  int grid = grid_for<K>(p.nwork);
  if (waves > 1) {
    const int g2 = (p.nwork + waves - 1) / waves;
    if (g2 < grid) grid = g2 > 0 ? g2 : 1;
  }
  if (t_pdl) {
    // programmatic dependent launch: this grid may become resident while its predecessor in the stream drains; smo_kernel blocks
    // in griddepcontrol.wait before it touches anything the predecessor wrote (SMO_OPT_PDL)
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)K::THREADS); cfg.dynamicSmemBytes = K::SMEM; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, smo_kernel<K>, p);
    if (e != cudaSuccess) return fail(SMO_E_CUDA, "cudaLaunchKernelEx failed: %s", cudaGetErrorString(e));
  } else {
    smo_kernel<K><<<grid, K::THREADS, K::SMEM, st>>>(p);
  }
  g_launches++;
  return rt_check("kernel launch");
}
#endif

#define TRY(x)             \
  do {                     \
    int rc_ = (x);         \
    if (rc_ != 0) return rc_; \
  } while (0)

static cplx* make_twiddles(int M, int* rc) {
  std::vector<cplx> t(M);
  for (int m = 0; m < M; ++m) {
    const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)M;
    t[m].x = (double)cosl(a);
    t[m].y = (double)(-sinl(a));
  }
  void* d = nullptr;
  *rc = rt_malloc(&d, sizeof(cplx) * M);
  if (*rc) return nullptr;
  *rc = rt_h2d(d, t.data(), sizeof(cplx) * M, 0);
  if (*rc == 0) *rc = rt_sync(0);
  return (cplx*)d;
}

extern "C" int smo_version(void) { return 100; }
extern "C" const char* smo_last_error(void) { return g_err.c_str(); }
extern "C" long long smo_launch_count(void) { return g_launches.load(); }

// ============================================================================================================
// vector ops
// ============================================================================================================
static int vec_nwork(long long n) { return (int)((n + VCHUNK - 1) / VCHUNK); }

extern "C" size_t smo_vec_work_bytes(long long n) { return sizeof(double) * (2 * (size_t)vec_nwork(n) + 8); }

template <int OP>
static int vec_launch(const double* x, const double* y, double* out, long long n, double a, double b, double c,
                      double* work, rt_stream st) {
  VecParams p;
  memset(&p, 0, sizeof p);
  p.x = x; p.y = y; p.out = out; p.n = n; p.a = a; p.b = b; p.c = c;
  p.nwork = vec_nwork(n); p.nsteps = 1;
  p.partials = work ? work + 8 : nullptr;
  p.scalars = work;
  p.vec2 = ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)out) & 15) == 0) ? 1 : 0;
  return launch<VecKernel<OP>>(p, st);
}
static int final_sum(double* work, long long n, int nq, double a, rt_stream st) {
  SumParams s;
  s.partials = work + 8; s.out = work; s.nwork = nq; s.nsteps = 1; s.npart = vec_nwork(n); s.nq = nq; s.a = a;
  return launch<FinalSum>(s, st);
}

extern "C" int smo_vec_dot(const double* x, const double* y, long long n, double scale, double* out_host, void* work,
                           void* stream) {
  if (!x || !y || !out_host || !work || n <= 0) return fail(SMO_E_ARG, "smo_vec_dot: bad argument");
  rt_stream st = (rt_stream)stream;
  double* w = (double*)work;
  TRY((vec_launch<V_DOT>(x, y, nullptr, n, 0, 0, 0, w, st)));
  TRY(final_sum(w, n, 1, scale, st));
  TRY(rt_d2h(out_host, w, sizeof(double), st));
  return rt_sync(st);
}
// the same without the D2H copy and the synchronisation: the scaled sum stays in work_dev[0] (double)
extern "C" int smo_vec_dot_dev(const double* x, const double* y, long long n, double scale, void* work, void* stream) {
  if (!x || !y || !work || n <= 0) return fail(SMO_E_ARG, "smo_vec_dot_dev: bad argument");
  rt_stream st = (rt_stream)stream;
  double* w = (double*)work;
  TRY((vec_launch<V_DOT>(x, y, nullptr, n, 0, 0, 0, w, st)));
  return final_sum(w, n, 1, scale, st);
}
extern "C" int smo_vec_dot_rows(const double* x, const double* y, int rows, long long len, double scale, double* out, void* stream) {
  if (!x || !y || !out || rows <= 0 || len <= 0) return fail(SMO_E_ARG, "smo_vec_dot_rows: bad argument");
  VecParams p;
  memset(&p, 0, sizeof p);
  p.x = x; p.y = y; p.out = out; p.n = len; p.a = scale; p.nwork = rows; p.nsteps = 1;
  return launch<VecRowDot>(p, (rt_stream)stream);
}
extern "C" int smo_vec_checksum(const double* x, long long n, unsigned long long* out_host, void* work, void* stream) {
  if (!x || !out_host || !work || n <= 0) return fail(SMO_E_ARG, "smo_vec_checksum: bad argument");
  rt_stream st = (rt_stream)stream;
  double* w = (double*)work;
  VecParams p;
  memset(&p, 0, sizeof p);
  p.x = x; p.n = n; p.nwork = vec_nwork(n); p.nsteps = 1; p.partials = w + 8;
  TRY(launch<VecHash>(p, st));
  SumParams sp;
  sp.partials = w + 8; sp.out = w; sp.nwork = 1; sp.nsteps = 1; sp.npart = vec_nwork(n); sp.nq = 1; sp.a = 1.0;
  TRY(launch<HashSum>(sp, st));
  TRY(rt_d2h(out_host, w, sizeof(unsigned long long), st));
  return rt_sync(st);
}
extern "C" int smo_vec_axpby(double a, const double* x, double b, const double* y, double* out, long long n,
                             void* stream) {
  if (!x || !out || n <= 0 || (!y && b != 0.0)) return fail(SMO_E_ARG, "smo_vec_axpby: bad argument");
  rt_stream st = (rt_stream)stream;
  if (!y) return vec_launch<V_SCALE>(x, nullptr, out, n, a, 0, 0, nullptr, st);
  return vec_launch<V_AXPBY>(x, y, out, n, a, b, 0, nullptr, st);
}
extern "C" int smo_vec_project(const double* x, const double* v, double* out, long long n, void* work, void* stream) {
  if (!x || !v || !out || !work || n <= 0) return fail(SMO_E_ARG, "smo_vec_project: bad argument");
  rt_stream st = (rt_stream)stream;
  double* w = (double*)work;
  TRY((vec_launch<V_DOT2>(x, v, nullptr, n, 0, 0, 0, w, st)));   // s0 = <x,v>, s1 = <x,x>
  TRY(final_sum(w, n, 2, 1.0, st));
  return vec_launch<V_PROJ>(x, v, out, n, 0, 0, 0, w, st);
}
extern "C" int smo_vec_retract(const double* x, double alpha, const double* d, double M0, double scale, double* out,
                               long long n, void* work, void* stream) {
  if (!x || !d || !out || !work || n <= 0) return fail(SMO_E_ARG, "smo_vec_retract: bad argument");
  rt_stream st = (rt_stream)stream;
  double* w = (double*)work;
  TRY((vec_launch<V_AXPY_NRM>(x, d, out, n, alpha, 0, 0, w, st)));
  TRY(final_sum(w, n, 1, 1.0, st));
  return vec_launch<V_RESCALE>(nullptr, nullptr, out, n, 0, M0, scale, w, st);
}

// ---- measured fp64 FMA peak (denominator of the SH23 ensemble's fp64 roofline) -----------------------------
#if !defined(SMO_EMUL)
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
  double v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = (double)(threadIdx.x + i) * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fma(v[i], a, b);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true for the operands used: keeps the chains alive
}
#endif
// runs 2*64*iters flop per thread on ctas_per_sm * #SMs CTAs of 256 threads; *flops_out = total flop of the launch
extern "C" int smo_microbench_dfma(double* out_dev, int iters, int ctas_per_sm, double* flops_out, void* stream) {
#if !defined(SMO_EMUL)
  if (!out_dev || iters <= 0 || ctas_per_sm <= 0) return fail(SMO_E_ARG, "smo_microbench_dfma: bad argument");
  const int grid = num_sms() * ctas_per_sm;
  dfma_peak_kernel<<<grid, 256, 0, (rt_stream)stream>>>(out_dev, iters, 0.999999, 1e-9);
  g_launches++;
  if (flops_out) *flops_out = 2.0 * 64.0 * (double)iters * 256.0 * (double)grid;
  return rt_check("dfma microbenchmark launch");
#else
  (void)out_dev; (void)iters; (void)ctas_per_sm; (void)flops_out; (void)stream;
  return fail(SMO_E_UNSUPPORTED, "no microbenchmarks in the host emulation");
#endif
}

// ============================================================================================================
// SH23
// ============================================================================================================
struct smo_sh23 {
  int N, M, H, Nh;
  double L, a;
  cplx *twH, *twM;
  // handle-owned buffers of the *_host entry points
  double* xin; double* jout; double* gout; cplx* snaps;
  size_t cap_batch, cap_snap;
};

constexpr int SH_NI = 4;

template <int H> static int sh23_run(smo_sh23* h, bool adj, Sh23Params& p, rt_stream st) {
  typedef typename FacOf<H>::type F;
  // few instances (a single optimisation: BASELINE config 1): latency bound, occupancy is irrelevant - the variant compiled
  // without the 128-register bound; ensembles: 8 CTAs per SM
  const bool few = p.nwork <= 2 * 148;
  if (adj) return few ? launch<Sh23Adj<F, SH_NI, 2>>(p, st) : launch<Sh23Adj<F, SH_NI>>(p, st);
  return few ? launch<Sh23Fwd<F, SH_NI, 2>>(p, st) : launch<Sh23Fwd<F, SH_NI>>(p, st);
}
static int sh23_dispatch(smo_sh23* h, bool adj, Sh23Params& p, rt_stream st) {
  p.nwork = (p.batch + SH_NI - 1) / SH_NI;
  if (p.xstride == 0) p.xstride = h->M;
  if (p.n_iters >= 0) p.sstride = (long long)(smo_sh23_snapshot_bytes(h, p.n_iters) / sizeof(double));
  p.Nh = h->Nh; p.a = h->a; p.kfac = 2.0 * 3.14159265358979323846 / h->L; p.inv_dt = 1.0 / p.dt;
  p.twH = h->twH; p.twM = h->twM;
  switch (h->H) {
    case 64: return sh23_run<64>(h, adj, p, st);
    case 128: return sh23_run<128>(h, adj, p, st);
    case 256: return sh23_run<256>(h, adj, p, st);
    default: return fail(SMO_E_UNSUPPORTED, "SH23: Npts=%d not supported (64, 128, 256)", h->N);
  }
}

extern "C" int smo_sh23_create(smo_sh23_t** out, int Npts, double L, double a) {
  if (!out) return fail(SMO_E_ARG, "smo_sh23_create: null handle pointer");
  if (Npts != 64 && Npts != 128 && Npts != 256) return fail(SMO_E_UNSUPPORTED, "SH23: Npts=%d not supported (64, 128, 256)", Npts);
  if (!(L > 0)) return fail(SMO_E_ARG, "SH23: L must be positive");
  smo_sh23* h = new smo_sh23();
  memset(h, 0, sizeof *h);
  h->N = Npts; h->M = 2 * Npts; h->H = Npts; h->Nh = Npts / 2; h->L = L; h->a = a;
  int rc = 0;
  h->twH = make_twiddles(h->H, &rc);
  if (rc == 0) h->twM = make_twiddles(h->M, &rc);
  if (rc) { smo_sh23_destroy(h); return rc; }
  *out = h;
  return 0;
}
extern "C" int smo_sh23_destroy(smo_sh23_t* h) {
  if (!h) return 0;
  rt_free(h->twH); rt_free(h->twM); rt_free(h->xin); rt_free(h->jout); rt_free(h->gout); rt_free(h->snaps);
  delete h;
  return 0;
}
// per instance: (n_iters+1) states on the grid (M doubles each) + the coefficients of the final state (Nh complex)
extern "C" size_t smo_sh23_snapshot_bytes(const smo_sh23_t* h, int n_iters) {
  return h ? sizeof(double) * (size_t)(n_iters + 1) * h->M + sizeof(cplx) * (size_t)h->Nh : 0;
}
static int sh23_args(smo_sh23* h, int batch, double dt, int n_iters, const char* who) {
  if (!h) return fail(SMO_E_ARG, "%s: null handle", who);
  if (batch <= 0 || n_iters < 0 || !(dt > 0)) return fail(SMO_E_ARG, "%s: bad batch/dt/n_iters", who);
  return 0;
}
extern "C" int smo_sh23_forward(smo_sh23_t* h, const double* X, int batch, double dt, int n_iters, void* snaps,
                                double* J, void* stream) {
  TRY(sh23_args(h, batch, dt, n_iters, "smo_sh23_forward"));
  if (!X || !snaps || !J) return fail(SMO_E_ARG, "smo_sh23_forward: null buffer");
  Sh23Params p;
  memset(&p, 0, sizeof p);
  p.X = X; p.snaps = (double*)snaps; p.J = J; p.batch = batch; p.n_iters = n_iters; p.dt = dt; p.flags = 0;
  p.nsteps = n_iters + 3;
  return sh23_dispatch(h, false, p, (rt_stream)stream);
}
extern "C" int smo_sh23_prep(smo_sh23_t* h, const double* X, int batch, double dt, int n_iters, double* out,
                             void* stream) {
  TRY(sh23_args(h, batch, dt, n_iters, "smo_sh23_prep"));
  if (!X || !out) return fail(SMO_E_ARG, "smo_sh23_prep: null buffer");
  Sh23Params p;
  memset(&p, 0, sizeof p);
  p.X = X; p.grad = out; p.batch = batch; p.n_iters = n_iters; p.dt = dt; p.flags = 2;
  p.nsteps = n_iters + 3;
  return sh23_dispatch(h, false, p, (rt_stream)stream);
}
static int sh23_reserve(smo_sh23* h, int batch, int n_iters);
static int sh23_coef_of(smo_sh23* h, const double* X, long long xstride, int batch, void* coef, rt_stream st) {
  Sh23Params p;
  memset(&p, 0, sizeof p);
  p.X = X; p.xstride = xstride; p.cout = (cplx*)coef; p.batch = batch; p.n_iters = 0; p.dt = 1.0; p.flags = 8;
  p.nsteps = 1;    // step 0 only: r2c of the input, coefficients -> cout
  return sh23_dispatch(h, false, p, st);
}
extern "C" int smo_sh23_to_coef(smo_sh23_t* h, const double* X, int batch, void* coef, void* stream) {
  TRY(sh23_args(h, batch, 1.0, 0, "smo_sh23_to_coef"));
  if (!X || !coef) return fail(SMO_E_ARG, "smo_sh23_to_coef: null buffer");
  return sh23_coef_of(h, X, h->M, batch, coef, (rt_stream)stream);
}
// coefficients [batch][Nh] of stored state n of every instance of a filled snapshot store (inspection: the reference's A_fwd[:, n])
extern "C" int smo_sh23_snapshot_coef(smo_sh23_t* h, const void* snaps, int batch, int n_iters, int n, void* coef, void* stream) {
  TRY(sh23_args(h, batch, 1.0, n_iters, "smo_sh23_snapshot_coef"));
  if (!snaps || !coef || n < 0 || n > n_iters) return fail(SMO_E_ARG, "smo_sh23_snapshot_coef: bad argument");
  const long long sstride = (long long)(smo_sh23_snapshot_bytes(h, n_iters) / sizeof(double));
  return sh23_coef_of(h, (const double*)snaps + (long long)n * h->M, sstride, batch, coef, (rt_stream)stream);
}
extern "C" int smo_sh23_to_grid(smo_sh23_t* h, const void* coef, int batch, double* out, void* stream) {
  TRY(sh23_args(h, batch, 1.0, 0, "smo_sh23_to_grid"));
  if (!out || !coef) return fail(SMO_E_ARG, "smo_sh23_to_grid: null buffer");
  Sh23Params p;
  memset(&p, 0, sizeof p);
  p.cin = (const cplx*)coef; p.grad = out; p.batch = batch; p.n_iters = -1; p.dt = 1.0; p.flags = 2 | 4;
  p.nsteps = 2;   // step 0 loads the coefficients, step 1 = final inverse transform of prep mode
  return sh23_dispatch(h, false, p, (rt_stream)stream);
}
extern "C" int smo_sh23_adjoint(smo_sh23_t* h, int batch, double dt, int n_iters, const void* snaps, double* grad,
                                int flags, void* stream) {
  TRY(sh23_args(h, batch, dt, n_iters, "smo_sh23_adjoint"));
  if (!snaps || !grad) return fail(SMO_E_ARG, "smo_sh23_adjoint: null buffer");
  Sh23Params p;
  memset(&p, 0, sizeof p);
  p.snaps = (double*)const_cast<void*>(snaps); p.grad = grad; p.batch = batch; p.n_iters = n_iters; p.dt = dt;
  p.flags = (flags & SMO_ADJOINT_CONTINUOUS) ? 1 : 0;
  p.nsteps = n_iters + 2;
  return sh23_dispatch(h, true, p, (rt_stream)stream);
}
static int sh23_reserve(smo_sh23* h, int batch, int n_iters) {
  if ((size_t)batch > h->cap_batch) {
    rt_free(h->xin); rt_free(h->jout); rt_free(h->gout);
    h->xin = h->jout = h->gout = nullptr; h->cap_batch = 0;
    TRY(rt_malloc((void**)&h->xin, sizeof(double) * batch * h->M));
    TRY(rt_malloc((void**)&h->gout, sizeof(double) * batch * h->M));
    TRY(rt_malloc((void**)&h->jout, sizeof(double) * batch));
    h->cap_batch = batch;
  }
  const size_t need = n_iters < 0 ? 0 : smo_sh23_snapshot_bytes(h, n_iters) * batch;   // n_iters < 0: caller's store
  if (need > h->cap_snap) {
    rt_free(h->snaps); h->snaps = nullptr; h->cap_snap = 0;
    TRY(rt_malloc((void**)&h->snaps, need));
    h->cap_snap = need;
  }
  return 0;
}
extern "C" int smo_sh23_forward_host(smo_sh23_t* h, const double* X, int batch, double dt, int n_iters, void* snaps,
                                     double* J, void* stream) {
  TRY(sh23_args(h, batch, dt, n_iters, "smo_sh23_forward_host"));
  if (!X || !J) return fail(SMO_E_ARG, "smo_sh23_forward_host: null buffer");
  rt_stream st = (rt_stream)stream;
  TRY(sh23_reserve(h, batch, snaps ? -1 : n_iters));
  TRY(rt_h2d(h->xin, X, sizeof(double) * batch * h->M, st));
  TRY(smo_sh23_forward(h, h->xin, batch, dt, n_iters, snaps ? snaps : h->snaps, h->jout, stream));
  TRY(rt_d2h(J, h->jout, sizeof(double) * batch, st));
  return rt_sync(st);
}
extern "C" int smo_sh23_adjoint_host(smo_sh23_t* h, int batch, double dt, int n_iters, const void* snaps, double* grad,
                                     int flags, void* stream) {
  TRY(sh23_args(h, batch, dt, n_iters, "smo_sh23_adjoint_host"));
  if (!grad) return fail(SMO_E_ARG, "smo_sh23_adjoint_host: null buffer");
  if (!snaps && (!h->snaps || smo_sh23_snapshot_bytes(h, n_iters) * batch > h->cap_snap))
    return fail(SMO_E_STATE, "smo_sh23_adjoint_host: no matching forward solve on this handle");
  rt_stream st = (rt_stream)stream;
  TRY(sh23_reserve(h, batch, -1));
  TRY(smo_sh23_adjoint(h, batch, dt, n_iters, snaps ? snaps : h->snaps, h->gout, flags, stream));
  TRY(rt_d2h(grad, h->gout, sizeof(double) * batch * h->M, st));
  return rt_sync(st);
}
extern "C" int smo_sh23_prep_host(smo_sh23_t* h, const double* X, int batch, double dt, int n_iters, double* out,
                                  void* stream) {
  TRY(sh23_args(h, batch, dt, n_iters, "smo_sh23_prep_host"));
  if (!X || !out) return fail(SMO_E_ARG, "smo_sh23_prep_host: null buffer");
  rt_stream st = (rt_stream)stream;
  TRY(sh23_reserve(h, batch, -1));
  TRY(rt_h2d(h->xin, X, sizeof(double) * batch * h->M, st));
  TRY(smo_sh23_prep(h, h->xin, batch, dt, n_iters, h->gout, stream));
  TRY(rt_d2h(out, h->gout, sizeof(double) * batch * h->M, st));
  return rt_sync(st);
}

// ============================================================================================================
// kinematic dynamo
// ============================================================================================================
struct smo_kdyn {
  int N, M, Nh, Nc, kmax, Pc;
  double L, kfac;
  int rank, nranks, nkx, kx0, nz, z0;
  size_t csize, p1size, p2size, gsize;
  cplx* tw;
  cplx* p1[MAXF];    // [s][nkx][Nc][nz] (kx-slab side of the transpose)
  cplx* p1t[MAXF];   // [Nh][Nc][nz]     (z-slab side; aliases p1 on one rank)
  cplx* p2[MAXF];    // [Nh][M][nz]
  cplx* cw[MAXF];    // coefficient work
  cplx* G[3]; cplx* NU[3]; cplx* W[3];
  double* jparts; size_t jparts_cap; size_t jparts_used; size_t jparts_need;   // cost "Integrated": per-CTA partial sums of |B^n|^2 of every x pass
  cplx* acc[3];      // [Nh][M][nz] running sum over the adjoint steps of the x-spectra of (curl G) x B_f (allocated on first use)
  double* Ug[3];     // projected velocity on the grid [M][M][nz]
  double* Ut;        // the same, tile-major [M*nz/4][3][M][4] (read by the fused x passes)
  double* gwork;     // 3*gsize doubles
  double* vwork;     // reduction workspace
  bool have_U;
  // handle-owned buffers of the *_host entry points
  double* hB; double* hU; double* hGB; double* hGU; cplx* snaps; size_t cap_snap;
  void* comm;
  // profiling
  int prof_which; double prof_ms; long long prof_n;
#if !defined(SMO_EMUL)
  std::vector<cudaEvent_t>* ev;
  size_t ev_used;
#endif
  int use_graph;
  // CUDA-graph replay of the time loops (launch-gap bound small grids / many GPUs): cache of instantiated graphs
  struct GraphCache* graphs;
  int capturing; unsigned long long cap_a0, cap_b0; unsigned long long* epoch_dev;
  // peer-memory transposes (CUDA IPC): pointers to every rank's p1 / p1t buffers and flag words
  int peer_on;
  cplx* peer_p1[MAXF][MAXP]; cplx* peer_p1t[MAXF][MAXP];
  unsigned long long* flags; unsigned long long* peer_flags[MAXP]; unsigned long long epoch;
  int chunks_fwd, chunks_adj;   // z-chunked y/x/y sequence (L2-resident P2 arrays); <= 1: whole slab at once
  // in-kernel hand-shakes of the peer-memory transposes (XSync): flag words [0..MAXP) barrier kernel, [MAXP..2MAXP) "p1
  // filled" (A, signalled by the forward y pass), [2MAXP..3MAXP) "p1t filled" (B, signalled by the z kernels)
  int inkernel_sync; unsigned long long epochA, epochB; unsigned int* counters;
  int l2_hints;                 // 1: L2 residency hints on the pencil data of the time loops (single rank)
  int peer_pull;                // 0 (default, measured faster on 2 B200): producers push; 1: consumers pull from the peers' buffers
  int bulk_u;                   // 1: the x passes fetch their velocity tile with one TMA bulk copy per tile
  int tma_sin;                  // 1: ... and their spectral tiles with TMA tensor copies (tensor maps cached per source array)
  struct TmCache* tmc;          // registered x-spectral arrays (work block, snapshot store, checkpoint segment) and their tensor maps
  cplx* p2block;                // the six y-padded work arrays are one allocation (one tensor map, slice = field)
  int bulk_push;                // peer-memory pushes through TMA bulk stores of staged blocks: bit 0 = fused z step, bit 1 = forward y pass
  int pdl;                      // 1: programmatic dependent launches inside the single-rank time loops (SMO_OPT_PDL)
  int grid_acc; double* accg;   // 1: the adjoint x pass sums (curl G) x B_f on the real grid (tile-major, 3*gsize doubles) instead of on the x-spectra
  int push_waves;               // pushing kernels run ~push_waves work items per CTA so that remote stores drain under compute
  int two_streams;              // 1: the z chunks of the y -> x -> y section alternate between two streams (transfers of one
                                // chunk overlap the x pass of the next)
  unsigned int* err_host; unsigned int* err_dev;   // host-mapped error word of the bounded hand-shake waits
#if !defined(SMO_EMUL)
  cudaStream_t aux_stream; cudaEvent_t ev_fork, ev_join, ev_pipe[3];
#endif
};
// flag words of one rank: [which][chunk][source rank]; which = 0 barrier kernel, XS_A, XS_B
constexpr int MAXCH = 4;
constexpr int NFLAGW = 3 * MAXCH * MAXP;

// bounded hand-shake waits (smo_common.cuh: xs_spin) raise a host-mapped word instead of hanging the box
static int kd_check_err(smo_kdyn* h, const char* who) {
#if !defined(SMO_EMUL)
  if (h->err_host && *(volatile unsigned int*)h->err_host)
    return fail(SMO_E_COMM, "%s: a cross-GPU hand-shake wait timed out (a peer rank died or stalled); results of this handle are invalid", who);
#endif
  (void)h; (void)who;
  return 0;
}

enum { PK_Z = 1, PK_Y = 2, PK_X = 3, PK_EPI = 4, PK_A2A = 5, PK_XA = 6, PK_ZS = 7 };

#if !defined(SMO_EMUL)
static void prof_begin(smo_kdyn* h, int kind, rt_stream st) {
  if (h->prof_which != kind) return;
  if (h->ev_used + 2 > h->ev->size()) {
    for (int i = 0; i < 2; ++i) { cudaEvent_t e; cudaEventCreate(&e); h->ev->push_back(e); }
  }
  // inside a stream capture the record must become a graph node of its own (external event), re-recorded by every replay
  if (h->capturing) cudaEventRecordWithFlags((*h->ev)[h->ev_used], st, cudaEventRecordExternal);
  else cudaEventRecord((*h->ev)[h->ev_used], st);
}
static void prof_end(smo_kdyn* h, int kind, rt_stream st) {
  if (h->prof_which != kind) return;
  if (h->capturing) cudaEventRecordWithFlags((*h->ev)[h->ev_used + 1], st, cudaEventRecordExternal);
  else cudaEventRecord((*h->ev)[h->ev_used + 1], st);
  h->ev_used += 2;
}
static void prof_collect(smo_kdyn* h, rt_stream st) {
  if (!h->prof_which || h->ev_used == 0) return;
  cudaStreamSynchronize(st);
  for (size_t i = 0; i + 1 < h->ev_used; i += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, (*h->ev)[i], (*h->ev)[i + 1]) == cudaSuccess) { h->prof_ms += ms; h->prof_n++; }
    else (void)cudaGetLastError();   // do not leave a stale error for the next launch check
  }
  h->ev_used = 0;
}
#else
static void prof_begin(smo_kdyn*, int, rt_stream) {}
static void prof_end(smo_kdyn*, int, rt_stream) {}
static void prof_collect(smo_kdyn*, rt_stream) {}
#endif

// ---- all-to-all transposes between the kx-slab layout p1 and the z-slab layout p1t -------------------------
// p1  = [s][nkx][Nc][nz] : block s goes to rank s;  p1t = [s][nkx][Nc][nz] read as [Nh][Nc][nz] : block s came
// from rank s.  Both directions exchange equal contiguous blocks of nkx*Nc*nz complex numbers per peer.
#if defined(SMO_EMUL)
typedef void (*smo_emul_a2a_fn)(const void* send, void* recv, long long bytes_per_peer, void* user);
struct EmulComm { smo_emul_a2a_fn fn; void* user; };
#endif
#if !defined(SMO_EMUL)
// Cross-GPU barrier of the peer-memory transposes: thread s publishes this rank's epoch in rank s's flag word
// (after a system-wide fence, so that the remote stores of the preceding pass kernel are visible first) and then
// waits until rank s has published the same epoch here.  One tiny launch per transpose; every GPU runs its own copy.
struct PeerFlags { unsigned long long* peer[MAXP]; };
__global__ void peer_barrier_kernel(PeerFlags pf, volatile unsigned long long* mine, int rank, int nranks, unsigned long long epoch) {
  const int s = (int)threadIdx.x;
  if (s < nranks) {
    __threadfence_system();
    *((volatile unsigned long long*)(pf.peer[s] + rank)) = epoch;
    __threadfence_system();
    while (mine[s] < epoch) { /* spin */ }
    __threadfence_system();
  }
}
#endif
static int a2a(smo_kdyn* h, cplx* const* src, cplx* const* dst, int nf, rt_stream st) {
  if (h->nranks == 1) return 0;
  const size_t blk = (size_t)h->nkx * h->Nc * h->nz;
#if !defined(SMO_EMUL)
  if (h->peer_on) {   // the data already sits in the peers' buffers (stores of the preceding pass): barrier only
    (void)src; (void)dst; (void)nf; (void)blk;
    PeerFlags pf;
    for (int s = 0; s < MAXP; ++s) pf.peer[s] = h->peer_flags[s];
    h->epoch++;
    prof_begin(h, PK_A2A, st);
    peer_barrier_kernel<<<1, 32, 0, st>>>(pf, h->flags, h->rank, h->nranks, h->epoch);
    g_launches++;
    prof_end(h, PK_A2A, st);
    return rt_check("peer barrier launch");
  }
#endif
#if defined(SMO_EMUL)
  EmulComm* c = (EmulComm*)h->comm;
  for (int f = 0; f < nf; ++f) c->fn(src[f], dst[f], (long long)(blk * sizeof(cplx)), c->user);
  return 0;
#elif defined(SMO_WITH_NCCL)
  ncclComm_t comm = (ncclComm_t)h->comm;
  NcclApi* nc = nccl_api();
  if (!nc) return fail(SMO_E_COMM, "libnccl.so.2 could not be loaded");
  prof_begin(h, PK_A2A, st);
  if (nc->GroupStart() != ncclSuccess) return fail(SMO_E_COMM, "ncclGroupStart failed");
  for (int f = 0; f < nf; ++f)
    for (int s = 0; s < h->nranks; ++s) {
      ncclResult_t r1 = nc->Send(src[f] + blk * s, blk * 2, ncclDouble, s, comm, st);
      ncclResult_t r2 = nc->Recv(dst[f] + blk * s, blk * 2, ncclDouble, s, comm, st);
      if (r1 != ncclSuccess || r2 != ncclSuccess) { nc->GroupEnd(); return fail(SMO_E_COMM, "ncclSend/Recv failed: %s", nc->GetErrorString(r1 != ncclSuccess ? r1 : r2)); }
    }
  ncclResult_t r = nc->GroupEnd();
  if (r != ncclSuccess) return fail(SMO_E_COMM, "ncclGroupEnd failed: %s", nc->GetErrorString(r));
  prof_end(h, PK_A2A, st);
  return 0;
#else
  (void)src; (void)dst; (void)nf; (void)st; (void)blk;
  return fail(SMO_E_UNSUPPORTED, "library built without NCCL: nranks must be 1");
#endif
}

#ifndef SMO_TZ
#define SMO_TZ 8
#endif
#ifndef SMO_TY
#define SMO_TY 8
#endif
#ifndef SMO_TX
#define SMO_TX 4
#endif
#ifndef SMO_TXA
#define SMO_TXA 4
#endif
#ifndef SMO_TZS
#define SMO_TZS 2
#endif
// ---- CUDA graphs ----------------------------------------------------------------------------------------------
// hsh = hash over EVERY buffer pointer the steps of the loop touch (checkpoint slots depend on `every`, which the first /
// last pointers alone do not pin down: two CheckpointStores of equal size but different spacing must not share a graph)
struct GraphKey {
  int kind, n; long long opts; const void* p0; const void* p1; const void* p2; double Rm, dt; unsigned long long hsh;
  bool operator==(const GraphKey& o) const {
    return kind == o.kind && n == o.n && opts == o.opts && p0 == o.p0 && p1 == o.p1 && p2 == o.p2 && Rm == o.Rm && dt == o.dt && hsh == o.hsh;
  }
};
static unsigned long long hash_ptr(unsigned long long hsh, const void* q) {
  hsh ^= (unsigned long long)(uintptr_t)q + 0x9e3779b97f4a7c15ull + (hsh << 6) + (hsh >> 2);
  return hsh;
}
#if !defined(SMO_EMUL)
struct GraphEntry { GraphKey key; cudaGraphExec_t exec; unsigned long long nA, nB; long long nlaunch; size_t ev0, ev1; };
struct GraphCache { std::vector<GraphEntry> e; cudaStream_t stream; cudaEvent_t ev0, ev1; };
__global__ void set_epoch_base_kernel(unsigned long long* dev, unsigned long long a, unsigned long long b) { dev[1] = a; dev[2] = b; }
#else
struct GraphCache { int unused; };
#endif

// ---- in-kernel hand-shakes ----------------------------------------------------------------------------------
enum { XS_NONE = 0, XS_A = 1, XS_B = 2 };
static bool kernel_sync(const smo_kdyn* h) { return h->peer_on && h->inkernel_sync; }
// the launch about to be issued publishes "buffer `which` (z chunk `chunk`) of every peer is filled by this rank" when it has
// finished.  Epochs: one per transpose (all chunks of one transpose carry the same epoch; xs_next_epoch starts a new one).
static std::atomic<int> g_trace_n{0};     // launch numbering of the development time-stamp trace (-DSMO_XS_TRACE)
static unsigned long long xs_next_epoch(smo_kdyn* h, int which) { return (which == XS_A) ? ++h->epochA : ++h->epochB; }
static void xs_signal(smo_kdyn* h, XSync& xs, int which, int chunk = 0, unsigned long long epoch = 0) {
  if (which == XS_NONE || !kernel_sync(h)) return;
  xs.sig_n = h->nranks; xs.sig_rank = h->rank; xs.sig_sys = h->peer_pull ? 0 : 1;
  xs.sig_epoch = epoch ? epoch : xs_next_epoch(h, which);
  if (h->capturing) { xs.sig_epoch -= (which == XS_A) ? h->cap_a0 : h->cap_b0; xs.sig_base = h->epoch_dev + which; }
  for (int s = 0; s < h->nranks; ++s) xs.sig_flags[s] = h->peer_flags[s] + (which * MAXCH + chunk) * MAXP;
  xs.counter = h->counters + which * MAXCH + chunk;
  xs.err = h->err_dev;
  if (!xs.trace_id) xs.trace_id = ++g_trace_n;
}
// the launch about to be issued first waits for the latest signal `which` (all `nch` chunks of it) of every rank
static void xs_wait(smo_kdyn* h, XSync& xs, int which, int nch = 1) {
  if (which == XS_NONE || !kernel_sync(h)) return;
  xs.wait_flags = h->flags + which * MAXCH * MAXP; xs.wait_n = h->nranks * nch; xs.wait_per = h->nranks;
  xs.wait_epoch = (which == XS_A) ? h->epochA : h->epochB;
  if (h->capturing) { xs.wait_epoch -= (which == XS_A) ? h->cap_a0 : h->cap_b0; xs.wait_base = h->epoch_dev + which; }
  xs.err = h->err_dev;
  if (!xs.trace_id) xs.trace_id = ++g_trace_n;
}
#if defined(SMO_XS_TRACE) && !defined(SMO_EMUL)
// development only (not declared in the header): copy out / clear the hand-shake time-stamp trace
extern "C" int smo_debug_xs_trace(unsigned long long* out, int reset) {
  if (out) CUDA_TRY(cudaMemcpyFromSymbol(out, g_xs_trace, sizeof(unsigned long long) * XS_TRACE_SLOTS * 8));
  if (reset) { void* d = nullptr; CUDA_TRY(cudaGetSymbolAddress(&d, g_xs_trace)); CUDA_TRY(cudaMemset(d, 0, sizeof(unsigned long long) * XS_TRACE_SLOTS * 8)); }
  return 0;
}
#endif

// ---- TMA tensor maps of the x-spectral arrays ------------------------------------------------------------------------
// An x-spectral array is [slices][Nh][ncols] complex (slices = fields x stored states, contiguous).  The fused x passes fetch a
// 4-column (or 2-column) x Nh-row box of one slice per tensor copy; the map (cuTensorMapEncodeTiled: FLOAT64 elements,
// dims {2*ncols, Nh, slices}, hardware swizzle matching the kernels' si()) depends only on the array, so it is cached.
struct TmEntry { const cplx* base; long long nslices; int box_cols, box_rows; SmoTensorMap map; bool have; };
struct TmCache { std::vector<TmEntry> e; };
static void tm_destroy(smo_kdyn* h) { delete h->tmc; h->tmc = nullptr; }
// (re)register an array the x passes may read spectral tiles from
static void tm_register(smo_kdyn* h, const void* base, long long nslices) {
  if (!h->tmc) h->tmc = new TmCache();
  for (TmEntry& t : h->tmc->e)
    if (t.base == (const cplx*)base) { if (t.nslices != nslices) { t.nslices = nslices; t.have = false; } return; }
  if (h->tmc->e.size() >= 16) h->tmc->e.erase(h->tmc->e.begin() + 1);     // (entry 0 = the handle's own work block)
  TmEntry t; t.base = (const cplx*)base; t.nslices = nslices; t.box_cols = t.box_rows = 0; t.have = false;
  h->tmc->e.push_back(t);
}
#if !defined(SMO_EMUL)
typedef CUresult (*tm_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tm_encode_fn tm_encoder() {
  static tm_encode_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess) fn = (tm_encode_fn)q;
    else (void)cudaGetLastError();
  }
  return fn;
}
#endif
// tensor map + slice index of the x-spectral field at `ptr`; false: not inside a registered array / no TMA -> the launch uses cp.async
static bool tm_resolve(smo_kdyn* h, const cplx* ptr, int box_cols, int box_rows, SmoTensorMap* out, int* z) {
#if defined(SMO_EMUL)
  (void)h; (void)ptr; (void)box_cols; (void)box_rows; (void)out; (void)z;
  return false;
#else
  if (!h->tmc) return false;
  for (TmEntry& t : h->tmc->e) {
    if (ptr < t.base || ptr >= t.base + (size_t)t.nslices * h->p2size) continue;
    if ((size_t)(ptr - t.base) % h->p2size) return false;
    if (!t.have || t.box_cols != box_cols || t.box_rows != box_rows) {
      tm_encode_fn enc = tm_encoder();
      if (!enc) return false;
      const cuuint64_t ncols = (cuuint64_t)h->M * h->nz;
      const cuuint64_t dims[3] = {2 * ncols, (cuuint64_t)h->Nh, (cuuint64_t)t.nslices};
      const cuuint64_t strides[2] = {ncols * sizeof(cplx), (cuuint64_t)h->p2size * sizeof(cplx)};
      const cuuint32_t box[3] = {(cuuint32_t)(2 * box_cols), (cuuint32_t)box_rows, 1};
      const cuuint32_t es[3] = {1, 1, 1};
      const CUtensorMapSwizzle sw = (box_cols * sizeof(cplx) == 64) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
      if (enc(&t.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)t.base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
      t.have = true; t.box_cols = box_cols; t.box_rows = box_rows;
    }
    *out = t.map;
    *z = (int)((size_t)(ptr - t.base) / h->p2size);
    return true;
  }
  return false;
#endif
}

// ---- pass launchers ---------------------------------------------------------------------------------------
// which fused x pass serves grid length M: the pair-packed XFused (two real columns per complex FFT of length M), or - where M
// has no 16-thread factorisation (M = 384) - the half-length XFusedH (one real column per complex FFT of length M/2)
#if defined(SMO_TEST_HALFX)   // test-only: exercises the half-length code paths at a small size (M = 96: H = 48 = 8 x 6)
template <int M> struct UseHalfX { static constexpr bool value = (M == 384) || (M == 96); };
#else
template <int M> struct UseHalfX { static constexpr bool value = (M == 384); };
#endif
template <int M, int MODE, bool INTEG, bool GACC, bool HALF> struct XKernelOf { typedef XFused<typename FacOf<M>::type, MODE, INTEG, GACC> type; };
#ifndef SMO_XH_T
#define SMO_XH_T 2      // columns per CTA of the half-length x pass (2: 192-thread CTAs, two per SM; 4: one 384-thread CTA per SM);
                        // r2m at 256^3: adjoint x pass 1.95 ms (T = 4) -> 1.81 ms (T = 2), forward 1.02 -> 1.04 ms
#endif
template <int M, int MODE, bool INTEG, bool GACC> struct XKernelOf<M, MODE, INTEG, GACC, true> { typedef XFusedH<typename FacOf<M / 2>::type, MODE, INTEG, GACC, SMO_XH_T> type; };

template <int M> struct KdOps {
  typedef typename FacOf<M>::type F;
  static constexpr bool HALFX = UseHalfX<M>::value;
  template <int MODE, bool INTEG, bool GACC = false> using XK = typename XKernelOf<M, MODE, INTEG, GACC, HALFX>::type;
  static constexpr int TZ = SMO_TZ;    // lines per CTA, contiguous (z) passes
  static constexpr int TY = SMO_TY;    // lines per CTA, strided (y) passes
  static constexpr int TX = SMO_TX;    // columns per CTA, x passes with <= 3 fields
  static constexpr int TXA = SMO_TXA;  // columns per CTA, fused adjoint x pass (6 fields)
#ifndef SMO_TZS_WIDE
#define SMO_TZS_WIDE SMO_TZS    // ... for grids whose z FFT needs 24 stage threads (M = 384): CTA-wide barriers, so smaller CTAs may pay
#endif
  static constexpr int TZS = (FacOf<M>::type::RT > 16) ? SMO_TZS_WIDE : SMO_TZS;  // lines per CTA (x 3 components), fused z step

  static void fill(PassParams& p, smo_kdyn* h, int nf) {
    memset(&p, 0, sizeof p);
    p.nfields = nf; p.nsteps = 1; p.kmax = h->kmax; p.tw = h->tw; p.scale = 1.0;
    p.in_sN = p.out_sN = 1; p.seglen = 0; p.blk = 0; p.b0 = 0;
  }
  // coefficient [nkx][Nc][Pc] -> p1 [s][nkx][Nc][nz]   (zero-pad + inverse FFT along z)
  static int inv_z(smo_kdyn* h, const cplx* const* in, cplx* const* out, int nf, rt_stream st, int sig = XS_NONE) {
    PassParams p; fill(p, h, nf);
    xs_signal(h, p.xs, sig);
    for (int f = 0; f < nf; ++f) { p.in[f] = in[f]; p.out[f] = out[f]; }
    p.nA = 1; p.nB = h->nkx * h->Nc; p.tilesB = (p.nB + TZ - 1) / TZ;
    p.in_sA = 0; p.in_sB = h->Pc;
    p.out_sA = 0; p.out_sB = h->nz;
    if (h->nranks > 1) { p.seglen = h->nz; p.blk = (long long)h->nkx * h->Nc * h->nz; }
    if (h->peer_on && !h->peer_pull && out == h->p1) {   // fused transpose (push): segment s is stored straight into rank s's p1t
      p.peer_mode = 1; p.peer_off = (long long)h->rank * p.blk;
      for (int f = 0; f < nf; ++f) for (int s2 = 0; s2 < h->nranks; ++s2) p.peer_out[f][s2] = h->peer_p1t[f][s2];
    }
    p.nwork = nf * p.nA * p.tilesB;
    prof_begin(h, PK_Z, st);
    int rc = launch<FftPass<F, +1, false, TZ>>(p, st, p.peer_mode ? h->push_waves : 1);
    prof_end(h, PK_Z, st);
    return rc;
  }
  // p1 [s][nkx][Nc][nz] -> coefficient   (forward FFT along z + truncation), scaled by 1/M
  static int fwd_z(smo_kdyn* h, const cplx* const* in, cplx* const* out, int nf, rt_stream st) {
    PassParams p; fill(p, h, nf);
    for (int f = 0; f < nf; ++f) { p.in[f] = in[f]; p.out[f] = out[f]; }
    p.nA = 1; p.nB = h->nkx * h->Nc; p.tilesB = (p.nB + TZ - 1) / TZ;
    p.in_sA = 0; p.in_sB = h->nz;
    if (h->nranks > 1) { p.seglen = h->nz; p.blk = (long long)h->nkx * h->Nc * h->nz; }
    p.out_sA = 0; p.out_sB = h->Pc;
    p.scale = 1.0 / M;
    if (h->peer_on && h->peer_pull && in == h->p1) {   // fused transpose (pull): z segment s is read out of rank s's p1t
      p.pull_mode = 1; p.pull_off = (long long)h->kx0 * h->Nc * h->nz;
      for (int f = 0; f < nf; ++f) for (int s2 = 0; s2 < h->nranks; ++s2) p.peer_in[f][s2] = h->peer_p1t[f][s2];
    }
    p.nwork = nf * p.nA * p.tilesB;
    prof_begin(h, PK_Z, st);
    int rc = launch<FftPass<F, -1, false, TZ>>(p, st);
    prof_end(h, PK_Z, st);
    return rc;
  }
  // p1t [Nh][Nc][nz] -> p2 [Nh][M][nz]   (zero-pad + inverse FFT along y) for the z range [z0, z0+nzc)
  static int inv_y(smo_kdyn* h, const cplx* const* in, cplx* const* out, int nf, rt_stream st, int z0 = 0, int nzc = -1,
                   int wait = XS_NONE, bool loop = false) {
    PassParams p; fill(p, h, nf);
    xs_wait(h, p.xs, wait);     // (the z kernels signal with one launch: chunk word 0)
    if (loop && h->l2_hints && h->nranks == 1) p.hint_in = 1;     // pencils: last use
    for (int f = 0; f < nf; ++f) { p.in[f] = in[f]; p.out[f] = out[f]; }
    p.nA = h->Nh; p.b0 = z0; p.nB = nzc < 0 ? h->nz : nzc; p.tilesB = (p.nB + TY - 1) / TY;
    p.in_sA = (long long)h->Nc * h->nz; p.in_sB = 1; p.in_sN = h->nz;
    p.out_sA = (long long)M * h->nz; p.out_sB = 1; p.out_sN = h->nz;
    if (h->peer_on && h->peer_pull && in == h->p1t) {   // fused transpose (pull): row kx is read out of its owner's p1
      p.pull_mode = 2; p.peer_rows = h->nkx; p.perm_rows = h->nkx; p.pull_off = (long long)h->rank * h->nkx * h->Nc * h->nz;
      for (int f = 0; f < nf; ++f) for (int s2 = 0; s2 < h->nranks; ++s2) p.peer_in[f][s2] = h->peer_p1[f][s2];
    }
    p.nwork = nf * p.nA * p.tilesB;
    prof_begin(h, PK_Y, st);
    int rc = launch<FftPass<F, +1, true, TY>>(p, st);
    prof_end(h, PK_Y, st);
    return rc;
  }
  static int fwd_y(smo_kdyn* h, const cplx* const* in, cplx* const* out, int nf, rt_stream st, int z0 = 0, int nzc = -1,
                   int sig = XS_NONE, bool loop = false, int chunk = 0, unsigned long long epoch = 0) {
    PassParams p; fill(p, h, nf);
    xs_signal(h, p.xs, sig, chunk, epoch);
    if (loop && h->l2_hints && h->nranks == 1) { p.hint_in = 1; p.hint_out = 2; }   // x-spectra: last use; pencils: keep for the z step
    for (int f = 0; f < nf; ++f) { p.in[f] = in[f]; p.out[f] = out[f]; }
    p.nA = h->Nh; p.b0 = z0; p.nB = nzc < 0 ? h->nz : nzc; p.tilesB = (p.nB + TY - 1) / TY;
    p.in_sA = (long long)M * h->nz; p.in_sB = 1; p.in_sN = h->nz;
    p.out_sA = (long long)h->Nc * h->nz; p.out_sB = 1; p.out_sN = h->nz;
    p.scale = 1.0 / M;
    if (h->peer_on && !h->peer_pull && out == h->p1t) {   // fused transpose (push): row kx is stored straight into its owner's p1
      p.peer_mode = 2; p.peer_rows = h->nkx; p.perm_rows = h->nkx; p.peer_off = (long long)h->rank * h->nkx * h->Nc * h->nz;
      for (int f = 0; f < nf; ++f) for (int s2 = 0; s2 < h->nranks; ++s2) p.peer_out[f][s2] = h->peer_p1[f][s2];
    }
    p.nwork = nf * p.nA * p.tilesB;
    prof_begin(h, PK_Y, st);
    int rc;
    if (p.peer_mode == 2 && (h->bulk_push & 2) && p.nB % TY == 0) rc = launch<FftPass<F, -1, true, TY, true>>(p, st, h->push_waves);   // staged rows, TMA bulk stores
    else rc = launch<FftPass<F, -1, true, TY>>(p, st, p.peer_mode ? h->push_waves : 1);
    prof_end(h, PK_Y, st);
    return rc;
  }
  static void xfill(XParams& p, smo_kdyn* h, int T) {
    memset(&p, 0, sizeof p);
    p.nsteps = 1; p.ncols = (long long)M * h->nz; p.Nh = h->Nh; p.tw = h->tw; p.scale = 1.0 / M;
    p.nwork = (int)(p.ncols / T);
  }
  static int x_c2r(smo_kdyn* h, const cplx* const* in, double* const* out, rt_stream st) {
    XParams p; xfill(p, h, TX);
    for (int f = 0; f < 3; ++f) { p.sin[f] = in[f]; p.gout[f] = out[f]; }
    return launch<XPass<F, TX, X_C2R, 3, 3>>(p, st);
  }
  static int x_r2c(smo_kdyn* h, const double* const* in, cplx* const* out, rt_stream st) {
    XParams p; xfill(p, h, TX);
    for (int f = 0; f < 3; ++f) { p.gin[f] = in[f]; p.sout[f] = out[f]; }
    return launch<XPass<F, TX, X_R2C, 3, 3>>(p, st);
  }
  // number of z chunks for the y -> x -> y sequence: the P2 arrays of one chunk (nf fields) should stay L2 resident
  // (~126 MB L2: aim at <= 32 MB per chunk); every chunk must hold whole tiles and still fill the GPU
  static int pick_chunks(smo_kdyn* h, int requested, int nf, int tile) {
    int nch = requested;
    if (nch < 0) {
      const double bytes = (double)nf * h->p2size * sizeof(cplx);
      nch = (int)(bytes / (32.0 * 1024 * 1024) + 0.999);
    }
    if (nch < 1) nch = 1;
    while (nch > 1 && (h->nz % nch != 0 || (h->nz / nch) % tile != 0 || (h->nz / nch) % TY != 0)) --nch;
    return nch;
  }
  static void xffill(XFParams& p, smo_kdyn* h, int T, int z0, int nzc) {
    memset(&p, 0, sizeof p);
    p.nsteps = 1; p.ncols = (long long)M * h->nz; p.Nh = h->Nh; p.tw = h->tw; p.scale = 1.0 / M; p.ut = h->Ut; p.bulk_u = h->bulk_u;
    if (nzc < 0) {   // all columns: tiles of T consecutive (y,z) columns, may straddle rows
      p.tiles_per_row = (int)(p.ncols / T); p.row_tiles = 0; p.tile0 = 0; p.nwork = p.tiles_per_row;
    } else {         // z range [z0, z0+nzc) of every row (needs nz, z0, nzc multiples of T)
      p.tiles_per_row = nzc / T; p.row_tiles = h->nz / T; p.tile0 = z0 / T; p.nwork = M * p.tiles_per_row;
    }
  }
  // cost "Integrated": partial-sum slots one forward step needs = sum over its z chunks of the x-pass grids (exact)
  static size_t jparts_per_step(smo_kdyn* h) {
    const int nch = yxy_chunks(h, 0);
    XFParams p; xffill(p, h, XK<X_FWD, true>::T, 0, nch > 1 ? h->nz / nch : -1);
    return (size_t)nch * (size_t)grid_for<XK<X_FWD, true>>(p.nwork);
  }
  // TMA path of the spectral tiles: tensor map + slice of every input field (fields 0..2 share one array, fields 3..5 another)
  template <class K> static void tma_sources(smo_kdyn* h, XFParams& p, int nf) {
    p.tma_sin = 0;
    if (!h->tma_sin || !K::TMA_OK) return;
    for (int g = 0; g < nf / 3; ++g) {
      for (int c = 0; c < 3; ++c) {
        SmoTensorMap m; int z = 0;
        if (!tm_resolve(h, p.sin[3 * g + c], K::TMA_BOX_COLS, K::TMA_BOX_ROWS, &m, &z)) return;
        if (c == 0) p.tm[g] = m;
        else if (memcmp(&m, &p.tm[g], sizeof m) != 0) return;      // the three components must live in one array
        p.tz[3 * g + c] = z;
      }
    }
    p.tma_sin = 1;
  }
  // forward: x-spectra of B (the snapshot slot or the work arrays) in, x-spectra of U x B out (work arrays)
  static int x_fwd(smo_kdyn* h, cplx* const* bp2, rt_stream st, int z0 = 0, int nzc = -1, bool integ = false) {
    XFParams p; xffill(p, h, XK<X_FWD, false>::T, z0, nzc);
    for (int f = 0; f < 3; ++f) { p.sin[f] = bp2[f]; p.sout[f] = h->p2[f]; }
    tma_sources<XK<X_FWD, false>>(h, p, 3);
    int rc;
    if (integ) {   // cost "Integrated": this launch's per-CTA sums of |B|^2 go to the next free slots of jparts
      const int grid = grid_for<XK<X_FWD, true>>(p.nwork);
      if (h->jparts_used + (size_t)grid > h->jparts_cap) return fail(SMO_E_STATE, "x_fwd: partial-sum buffer too small");
      p.jpart = h->jparts + h->jparts_used; h->jparts_used += (size_t)grid;
      prof_begin(h, PK_X, st);
      rc = launch<XK<X_FWD, true>>(p, st);
    } else {
      prof_begin(h, PK_X, st);
      rc = launch<XK<X_FWD, false>>(p, st);
    }
    prof_end(h, PK_X, st);
    return rc;
  }
  // adjoint: x-spectra of curl G (work arrays) and of the forward state B_f (read straight from its snapshot slot) in
  static int x_adj(smo_kdyn* h, cplx* const* bfp2, rt_stream st, int z0 = 0, int nzc = -1, bool integ = false) {
    XFParams p; xffill(p, h, XK<X_ADJ, false>::T, z0, nzc);
    for (int f = 0; f < 3; ++f) { p.sin[f] = h->p2[f]; p.sout[f] = h->p2[f]; p.sin[3 + f] = bfp2[f]; }
    tma_sources<XK<X_ADJ, false, true>>(h, p, 6);
    p.accumulate = 1;
    prof_begin(h, PK_XA, st);
    int rc;
    if (h->grid_acc) {   // (curl G) x B_f: summed over the sweep on the real grid, transformed once afterwards (nu_finish)
      p.gacc = h->accg;
      rc = integ ? launch<XK<X_ADJ, true, true>>(p, st) : launch<XK<X_ADJ, false, true>>(p, st);
    } else {             // ... summed on the x-spectra
      for (int c = 0; c < 3; ++c) p.sout[3 + c] = h->acc[c];
      rc = integ ? launch<XK<X_ADJ, true>>(p, st) : launch<XK<X_ADJ, false>>(p, st);
    }
    prof_end(h, PK_XA, st);
    return rc;
  }
  static int u_tile(smo_kdyn* h, rt_stream st) {
    UTileParams p; memset(&p, 0, sizeof p);
    for (int c = 0; c < 3; ++c) p.in[c] = h->Ug[c];
    p.out = h->Ut; p.ncols = (long long)M * h->nz; p.M = M; p.nsteps = 1; p.half = HALFX ? XK<X_FWD, false>::T : 0;
    p.nwork = (int)(M * ((p.ncols + UTile::THREADS - 1) / UTile::THREADS));
    if (HALFX) return launch<UTileH>(p, st);
    return launch<UTile>(p, st);
  }

  // grid [3][gsize] -> coefficients [3][csize]
  static int to_coef(smo_kdyn* h, const double* grid, cplx* const* coef, rt_stream st) {
    const double* g[3] = {grid, grid + h->gsize, grid + 2 * h->gsize};
    TRY(x_r2c(h, g, h->p2, st));
    if (h->peer_on) TRY(a2a(h, h->p1t, h->p1, 3, st));   // no peer still uses the buffers the y pass is about to fill
    TRY(fwd_y(h, h->p2, h->p1t, 3, st));
    TRY(a2a(h, h->p1t, h->p1, 3, st));
    return fwd_z(h, h->p1, coef, 3, st);
  }
  static int to_grid(smo_kdyn* h, const cplx* const* coef, double* grid, rt_stream st) { return to_grid_via(h, coef, grid, h->p2, st); }
  // the same, leaving the x-spectra (y- and z-padded) of the field in xs
  static int to_grid_via(smo_kdyn* h, const cplx* const* coef, double* grid, cplx* const* xs, rt_stream st) {
    double* g[3] = {grid, grid + h->gsize, grid + 2 * h->gsize};
    if (h->peer_on) TRY(a2a(h, h->p1, h->p1t, 3, st));   // no peer still uses the buffers the z pass is about to fill
    TRY(inv_z(h, coef, h->p1, 3, st));
    TRY(a2a(h, h->p1, h->p1t, 3, st));
    TRY(inv_y(h, h->p1t, xs, 3, st));
    return x_c2r(h, xs, g, st);
  }
  // fused z step (zstep.cuh): p1 -> [forward z FFT, implicit update of the state, inverse z FFT] -> p1
  //   mode 0: state Bn -> Bnp1, next operand Bnp1;  mode 1: state G in place, next operand curl G'
  static int zstep(smo_kdyn* h, int mode, const cplx* const* Bn, cplx* const* Bnp1, bool do_inv, double Rm, double dt, rt_stream st,
                   int nch_wait = 1) {
    ZParams p; memset(&p, 0, sizeof p);
    xs_wait(h, p.xs, XS_A, nch_wait);
    if (do_inv) xs_signal(h, p.xs, XS_B);
    const int nf = 3;
    for (int f = 0; f < nf; ++f) { p.in[f] = h->p1[f]; p.out[f] = h->p1[f]; }
    if (mode == 0) {
      for (int c = 0; c < 3; ++c) { p.b[c] = Bn[c]; p.o[c] = Bnp1[c]; }
    } else {
      for (int c = 0; c < 3; ++c) { p.b[c] = h->G[c]; p.o[c] = h->G[c]; }
    }
    p.nsteps = 1; p.mode = mode; p.ntrip = 1;
    p.nlines = h->nkx * h->Nc; p.tiles = (p.nlines + TZS - 1) / TZS; p.nwork = p.tiles * p.ntrip;
    p.do_inv = do_inv ? 1 : 0;
    p.l2_hints = (h->l2_hints && h->nranks == 1) ? 1 : 0;
    p.Nc = h->Nc; p.Pc = h->Pc; p.kmax = h->kmax; p.kx0 = h->kx0;
    p.line_stride = h->nz; p.kfac = h->kfac; p.Rm = Rm; p.dt = dt; p.scale = 1.0 / M; p.tw = h->tw;
    if (h->nranks > 1) { p.seglen = h->nz; p.blk = (long long)h->nkx * h->Nc * h->nz; }
    if (h->peer_on && !h->peer_pull) {
      p.peer_mode = 1; p.peer_off = (long long)h->rank * p.blk;
      p.bulk_push = (h->bulk_push & 1) ? 1 : 0;
      for (int f = 0; f < nf; ++f) for (int s2 = 0; s2 < h->nranks; ++s2) p.peer_out[f][s2] = h->peer_p1t[f][s2];
    } else if (h->peer_on) {
      p.pull_mode = 1; p.pull_off = (long long)h->kx0 * h->Nc * h->nz;
      for (int f = 0; f < nf; ++f) for (int s2 = 0; s2 < h->nranks; ++s2) p.peer_in[f][s2] = h->peer_p1t[f][s2];
    }
    prof_begin(h, PK_ZS, st);
    int rc = launch<ZStep<F, TZS>>(p, st, (p.peer_mode && do_inv) ? h->push_waves : 1);
    prof_end(h, PK_ZS, st);
    return rc;
  }
  // y -> fused x -> y part of a step: p1t (z-slab side) -> p1t.  xs = x-spectra of the forward state: WRITTEN by the inverse
  // y pass of a forward step (the snapshot slot of state n, or the work arrays), READ by the x pass of an adjoint step.
  // (with in-kernel hand-shakes the first y pass waits for "p1t filled", the last one signals "p1 filled")
  static int yxy_chunks(smo_kdyn* h, int mode) {
    const int nch = pick_chunks(h, mode == 0 ? h->chunks_fwd : h->chunks_adj, mode == 0 ? 3 : 6, TY > 4 ? TY : 4);
    return (kernel_sync(h) && nch > MAXCH) ? 1 : nch;   // (one hand-shake flag word per chunk)
  }
  static int yxy(smo_kdyn* h, int mode, cplx* const* xs, rt_stream st, bool integ = false) {
    const int nch = yxy_chunks(h, mode);
    const int nzc = h->nz / nch;
    // every chunk's last y pass signals "p1 filled" on its own flag word with the same epoch; the z step waits for all of them
    const unsigned long long eA = kernel_sync(h) ? xs_next_epoch(h, XS_A) : 0ull;
#if !defined(SMO_EMUL)
    // two streams: odd chunks run on an auxiliary stream, so that the remote stores (and the hand-shake) of one chunk's last
    // y pass drain while the other chunk is still in its x pass.  Works eagerly and inside a stream capture (fork / join events).
    const bool two = h->two_streams && nch > 1;
    // two_streams = 2: software pipeline - a kernel of chunk c+1 starts only after the SAME kernel of chunk c (events), so the
    // chunks are staggered by one stage: chunk c's pushing y pass (NVLink transfer + hand-shake) runs beside chunk c+1's x pass
    // instead of both chunks marching in lock step (r2c: unstaggered chunks gained nothing).
    const bool stagger = two && h->two_streams >= 2;
    if (two && !h->aux_stream) {
      CUDA_TRY(cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
      CUDA_TRY(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
      for (int k = 0; k < 3; ++k) CUDA_TRY(cudaEventCreateWithFlags(&h->ev_pipe[k], cudaEventDisableTiming));
    }
    if (two) { CUDA_TRY(cudaEventRecord(h->ev_fork, st)); CUDA_TRY(cudaStreamWaitEvent(h->aux_stream, h->ev_fork, 0)); }
#else
    const bool two = false, stagger = false;
#endif
    for (int ch = 0; ch < nch; ++ch) {
      const int z0 = ch * nzc, zc = nch > 1 ? nzc : -1;
#if !defined(SMO_EMUL)
      rt_stream q = (two && (ch & 1)) ? h->aux_stream : st;
#else
      rt_stream q = st;
#endif
#if !defined(SMO_EMUL)
#define SMO_STAGE_GATE(k)                                                                              \
      if (stagger) {                                                                                   \
        if (ch > 0) CUDA_TRY(cudaStreamWaitEvent(q, h->ev_pipe[k], 0));                                 \
      }
#define SMO_STAGE_DONE(k)                                                                              \
      if (stagger && ch + 1 < nch) CUDA_TRY(cudaEventRecord(h->ev_pipe[k], q));
#else
#define SMO_STAGE_GATE(k)
#define SMO_STAGE_DONE(k)
#endif
      SMO_STAGE_GATE(0)
      TRY(inv_y(h, h->p1t, mode == 0 ? xs : h->p2, 3, q, z0, zc, (ch == 0 || (two && !stagger)) ? XS_B : XS_NONE, true));
      SMO_STAGE_DONE(0)
      SMO_STAGE_GATE(1)
      if (mode == 0) TRY(x_fwd(h, xs, q, z0, zc, integ));
      else TRY(x_adj(h, xs, q, z0, zc, integ));
      SMO_STAGE_DONE(1)
      SMO_STAGE_GATE(2)
      TRY(fwd_y(h, h->p2, h->p1t, 3, q, z0, zc, kernel_sync(h) ? XS_A : XS_NONE, true, ch, eA));
      SMO_STAGE_DONE(2)
#undef SMO_STAGE_GATE
#undef SMO_STAGE_DONE
    }
#if !defined(SMO_EMUL)
    if (two) { CUDA_TRY(cudaEventRecord(h->ev_join, h->aux_stream)); CUDA_TRY(cudaStreamWaitEvent(st, h->ev_join, 0)); }
#endif
    return 0;
  }
  // first half of a forward step only: x-spectra of the state whose z-padded form sits in p1t
  static int y_only(smo_kdyn* h, cplx* const* xs, rt_stream st) { return inv_y(h, h->p1t, xs, 3, st, 0, -1, XS_B, true); }
  static void efill(EpiParams& p, smo_kdyn* h, double Rm, double dt, int flag) {
    memset(&p, 0, sizeof p);
    p.nsteps = 1; p.n = (long long)h->csize; p.Nc = h->Nc; p.Pc = h->Pc; p.kmax = h->kmax; p.kx0 = h->kx0;
    p.kfac = h->kfac; p.Rm = Rm; p.dt = dt; p.flag = flag;
    p.nwork = (int)((p.n + 255) / 256);
  }
  static int compat(smo_kdyn* h, const cplx* const* BN, double Rm, double dt, int flag, rt_stream st) {
    EpiParams e; efill(e, h, Rm, dt, flag);
    for (int c = 0; c < 3; ++c) { e.b[c] = BN[c]; e.o[c] = h->G[c]; e.o2[c] = h->W[c]; }
    return launch<EpiKernel<EPI_COMPAT>>(e, st);
  }
  static int curl_G(smo_kdyn* h, double Rm, double dt, rt_stream st) {
    EpiParams e; efill(e, h, Rm, dt, 0);
    for (int c = 0; c < 3; ++c) { e.b[c] = h->G[c]; e.o2[c] = h->W[c]; }
    return launch<EpiKernel<EPI_CURL>>(e, st);
  }
  // start of an adjoint sweep: the running sum of the (curl G) x B_f spectra is cleared
  static int acc_begin(smo_kdyn* h, rt_stream st) {
    if (h->grid_acc) {
      if (!h->accg) TRY(rt_malloc((void**)&h->accg, sizeof(double) * 3 * h->gsize));
      return rt_memset(h->accg, 0, sizeof(double) * 3 * h->gsize, st);
    }
    for (int c = 0; c < 3; ++c) {
      if (!h->acc[c]) TRY(rt_malloc((void**)&h->acc[c], sizeof(cplx) * h->p2size));
      TRY(rt_memset(h->acc[c], 0, sizeof(cplx) * h->p2size, st));
    }
    return 0;
  }
  // end of the sweep: nu^N = -dt P_k[ y/z transforms of the running sum ]   (once instead of every step)
  static int nu_finish(smo_kdyn* h, double Rm, double dt, rt_stream st) {
    if (h->grid_acc) {   // tile-major running sum -> grid -> x-spectra (the one r2c transform the sweep skipped)
      UTileParams u; memset(&u, 0, sizeof u);
      for (int c = 0; c < 3; ++c) u.in[c] = h->gwork + (size_t)c * h->gsize;
      u.out = h->accg; u.ncols = (long long)M * h->nz; u.M = M; u.nsteps = 1; u.half = HALFX ? XK<X_FWD, false>::T : 0;
      u.nwork = (int)(M * ((u.ncols + UnTile::THREADS - 1) / UnTile::THREADS));
      TRY(launch<UnTile>(u, st));
      const double* g[3] = {h->gwork, h->gwork + h->gsize, h->gwork + 2 * h->gsize};
      TRY(x_r2c(h, g, h->p2, st));
    }
    cplx* const* accx = h->grid_acc ? h->p2 : h->acc;
    if (h->peer_on) TRY(a2a(h, h->p1t, h->p1, 3, st));
    TRY(fwd_y(h, accx, h->p1t, 3, st));
    TRY(a2a(h, h->p1t, h->p1, 3, st));
    TRY(fwd_z(h, h->p1, h->cw, 3, st));
    EpiParams e; efill(e, h, Rm, dt, 0);
    for (int c = 0; c < 3; ++c) { e.b[c] = h->cw[c]; e.o[c] = h->NU[c]; }
    return launch<EpiKernel<EPI_NUFIN>>(e, st);
  }
  static int final_scale(smo_kdyn* h, double Rm, double dt, int flag, rt_stream st) {
    EpiParams e; efill(e, h, Rm, dt, flag);
    for (int c = 0; c < 3; ++c) { e.b[c] = h->G[c]; e.o[c] = h->cw[c]; }
    return launch<EpiKernel<EPI_FINAL>>(e, st);
  }
};

// Runs body(stream) - a fixed sequence of kernel launches - eagerly, or (smo_kdyn_use_graph) replays it from a cached CUDA
// graph: the first call with a given key runs eagerly, the second one captures and instantiates, later ones only launch.
template <class Body> static int run_graphed(smo_kdyn* h, const GraphKey& key, rt_stream st, Body body) {
#if defined(SMO_EMUL)
  (void)h; (void)key;
  return body(st);
#else
  // (per-kernel profiling events are captured as event-record nodes: every replay re-records the same events)
  const bool can = h->use_graph && (h->nranks == 1 || kernel_sync(h));
  if (!can) return body(st);
  if (!h->graphs) {
    h->graphs = new GraphCache();
    CUDA_TRY(cudaStreamCreateWithFlags(&h->graphs->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&h->graphs->ev0, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->graphs->ev1, cudaEventDisableTiming));
    TRY(rt_malloc((void**)&h->epoch_dev, sizeof(unsigned long long) * 4));
  }
  GraphCache* gc = h->graphs;
  GraphEntry* en = nullptr;
  for (GraphEntry& e : gc->e) if (e.key == key) { en = &e; break; }
  if (!en) {   // first sight: eager (also instantiates every kernel's launch configuration outside a capture)
    if (gc->e.size() >= 64) { if (gc->e.front().exec) cudaGraphExecDestroy(gc->e.front().exec); gc->e.erase(gc->e.begin()); }
    GraphEntry e; e.key = key; e.exec = nullptr; e.nA = e.nB = 0; e.nlaunch = 0; e.ev0 = e.ev1 = 0;
    gc->e.push_back(e);
    return body(st);
  }
  if (!en->exec) {
    const unsigned long long a0 = h->epochA, b0 = h->epochB;
    const long long l0 = g_launches.load();
    en->ev0 = h->ev_used;
    cudaGraph_t graph = nullptr;
    CUDA_TRY(cudaStreamBeginCapture(gc->stream, cudaStreamCaptureModeThreadLocal));
    h->capturing = 1; h->cap_a0 = a0; h->cap_b0 = b0;
    const int rc = body(gc->stream);
    const cudaError_t ce = cudaStreamEndCapture(gc->stream, &graph);
    h->capturing = 0;
    en->nA = h->epochA - a0; en->nB = h->epochB - b0; en->nlaunch = g_launches.load() - l0;
    h->epochA = a0; h->epochB = b0; g_launches -= en->nlaunch;
    en->ev1 = h->ev_used; h->ev_used = en->ev0;
    if (rc != 0 || ce != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      return rc ? rc : fail(SMO_E_CUDA, "stream capture failed: %s", cudaGetErrorString(ce));
    }
    const cudaError_t ie = cudaGraphInstantiate(&en->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { en->exec = nullptr; return fail(SMO_E_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ie)); }
  }
  if (h->ev_used != en->ev0) return body(st);   // the profiling events baked into this graph are in use: run eagerly
  CUDA_TRY(cudaEventRecord(gc->ev0, st));
  CUDA_TRY(cudaStreamWaitEvent(gc->stream, gc->ev0, 0));
  if (h->nranks > 1) set_epoch_base_kernel<<<1, 1, 0, gc->stream>>>(h->epoch_dev, h->epochA, h->epochB);
  CUDA_TRY(cudaGraphLaunch(en->exec, gc->stream));
  CUDA_TRY(cudaEventRecord(gc->ev1, gc->stream));
  CUDA_TRY(cudaStreamWaitEvent(st, gc->ev1, 0));
  h->epochA += en->nA; h->epochB += en->nB; g_launches += en->nlaunch; h->ev_used = en->ev1;
  return 0;
#endif
}
// programmatic dependent launches: only where every launch of the loop is a smo_kernel on ONE stream of ONE rank.  -1 (default) =
// automatic: on for the small grids, whose steps are launch-latency bound (B200, r2z: 16^3 -16 %, 24^3 -10.5 %, 32^3 -7 %, 64^3 -1.5 %
// per Grad_f pair), off for the large ones, where CTAs that move in early make the kernels 5 % slower (128^3, 256^3)
static int pdl_on(const smo_kdyn* h) {
  if (h->nranks != 1 || h->two_streams != 0 || h->prof_which != 0) return 0;
  return h->pdl < 0 ? (h->N <= 64 ? 1 : 0) : (h->pdl ? 1 : 0);
}
static long long graph_opts(const smo_kdyn* h) {
  return (pdl_on(h) ? (1ll << 31) : 0ll) | ((long long)(h->bulk_push & 3) << 32) | ((h->prof_which & 0xf) << 24) | (h->l2_hints ? 1 : 0) | (h->peer_pull ? 2 : 0) | (h->inkernel_sync ? 4 : 0) | (h->peer_on ? 8 : 0) |
         ((h->two_streams & 1) ? 16 : 0) | ((h->two_streams & 2) ? (1 << 29) : 0) | ((h->push_waves & 3) << 5) | (h->grid_acc ? 128 : 0) | (h->bulk_u ? (1 << 28) : 0) | (h->tma_sin ? (1 << 30) : 0) | ((h->chunks_fwd & 0xff) << 8) | ((h->chunks_adj & 0xff) << 16);
}

// ---- snapshot store ------------------------------------------------------------------------------------------------
// Forward states are kept in the form the adjoint x pass consumes: their x-spectra on this rank's z-slab, [Nh][M][nz] per
// component (y- and z-padded: 2.3 x the bytes of the coefficients).  The inverse y pass of forward step n writes them
// straight into slot n - no extra traffic in the forward solve - and the adjoint sweep needs no z / y transform and no
// transpose of the forward state at all.  Layout: n_iters + 1 slots of 3 * p2size, then the coefficients of the final
// state (3 * csize; terminal condition of the adjoint).  Checkpoints (kd_*_ckpt) stay in coefficient form.
static void snap_ptrs(smo_kdyn* h, void* snaps, int n, cplx** out) {     // coefficient-form slot n (checkpoint stores)
  cplx* s = (cplx*)snaps;
  for (int c = 0; c < 3; ++c) out[c] = s + ((size_t)n * 3 + c) * h->csize;
}
static void snap_xs(smo_kdyn* h, void* snaps, int n, cplx** out) {       // x-spectral slot n
  cplx* s = (cplx*)snaps;
  for (int c = 0; c < 3; ++c) out[c] = s + ((size_t)n * 3 + c) * h->p2size;
}
static void snap_final(smo_kdyn* h, void* snaps, int n_iters, cplx** out) {
  cplx* s = (cplx*)snaps + (size_t)(n_iters + 1) * 3 * h->p2size;
  for (int c = 0; c < 3; ++c) out[c] = s + (size_t)c * h->csize;
}

// loop bodies, templated on M ---------------------------------------------------------------------------------
template <int M> static int kd_set_U(smo_kdyn* h, const double* U, rt_stream st) {
  // parameter fields are projected on the retained modes at first use [D2-8]: to_coef then to_grid
  TRY(KdOps<M>::to_coef(h, U, h->cw, st));
  double* g = h->gwork;
  TRY(KdOps<M>::to_grid(h, h->cw, g, st));
  for (int c = 0; c < 3; ++c) TRY(rt_d2d(h->Ug[c], g + (size_t)c * h->gsize, sizeof(double) * h->gsize, st));
  TRY(KdOps<M>::u_tile(h, st));
  h->have_U = true;
  return 0;
}
// buffers of forward step n: coefficients of state n in, of state n+1 out, and where the x-spectra of state n go
struct FwdStep { cplx* cin[3]; cplx* cout[3]; cplx* xs[3]; };
// Time loop of the forward problem: 4 launches per step.  The state travels  coefficients -> p1 -> [y inverse] -> x-spectra
// (snapshot slot) -> [fused x] -> [y forward] -> p1 -> [fused z step: coefficients of step n+1 written, and already on their
// way back to p1].  tail_xs_only: the last "step" only produces the x-spectra of its state (snapshot of the final state).
template <int M, class StepFn>
static int kd_forward_loop(smo_kdyn* h, int n_steps, bool tail_xs_only, double Rm, double dt, StepFn step, rt_stream st,
                           bool integ = false) {
  if (n_steps <= 0) return 0;
  // multi-GPU: the transposes are remote stores / loads of the kernels themselves; their hand-shakes are either fused into
  // the kernels (ks: producer signals at its end, consumer waits at its start) or separate barrier launches (a2a)
  const bool ks = kernel_sync(h);
  if (ks) TRY(a2a(h, h->p1, h->p1t, 3, st));     // every rank has left whatever used the pencil buffers before
  const FwdStep first = step(0), last = step(n_steps - 1);
  GraphKey key; key.kind = (tail_xs_only ? 3 : 1) + (integ ? 10 : 0); key.n = n_steps; key.opts = graph_opts(h); key.p0 = first.xs[0]; key.p1 = last.cout[0];
  key.p2 = first.cin[0]; key.Rm = Rm; key.dt = dt; key.hsh = 0;
  for (int n = 0; n < n_steps; ++n) { const FwdStep b = step(n); key.hsh = hash_ptr(hash_ptr(hash_ptr(key.hsh, b.cin[0]), b.cout[0]), b.xs[0]); }
  return run_graphed(h, key, st, [&](rt_stream s) -> int {
    PdlScope pdl(pdl_on(h));
    TRY(KdOps<M>::inv_z(h, first.cin, h->p1, 3, s, XS_B));
    if (!ks) TRY(a2a(h, h->p1, h->p1t, 3, s));
    for (int n = 0; n < n_steps; ++n) {
      const FwdStep b = step(n);
      if (tail_xs_only && n == n_steps - 1) return KdOps<M>::y_only(h, b.xs, s);
      TRY(KdOps<M>::yxy(h, 0, b.xs, s, integ));
      if (!ks) TRY(a2a(h, h->p1t, h->p1, 3, s));
      const bool more = n + 1 < n_steps;
      TRY(KdOps<M>::zstep(h, 0, b.cin, b.cout, more, Rm, dt, s, KdOps<M>::yxy_chunks(h, 0)));
      if (more && !ks) TRY(a2a(h, h->p1, h->p1t, 3, s));
    }
    return 0;
  });
}
// coefficient state n of a plain time loop ping-pongs between two scratch triplets
static void pingpong(cplx* const* a, cplx* const* b, int n, FwdStep& f) {
  for (int c = 0; c < 3; ++c) { f.cin[c] = (n & 1) ? b[c] : a[c]; f.cout[c] = (n & 1) ? a[c] : b[c]; }
}
// cost "Integrated" (KD:655-669): J = dt * sum_{n=0}^{N} <B^n,B^n>.  The states 0..N-1 pass through the forward x pass, which
// sums |B^n|^2 over its grid points into per-CTA partials (deterministic: fixed tile assignment, fixed-order sums); the final
// state is added from its grid values.
template <int M> static int jparts_begin(smo_kdyn* h, int n_steps, rt_stream st) {
  const size_t need = (size_t)n_steps * KdOps<M>::jparts_per_step(h);   // exactly the slots the x passes of this solve fill
  if (need > h->jparts_cap) {
    rt_free(h->jparts); h->jparts = nullptr; h->jparts_cap = 0;
#if !defined(SMO_EMUL)
    if (h->graphs) { for (GraphEntry& e : h->graphs->e) if (e.exec) cudaGraphExecDestroy(e.exec); h->graphs->e.clear(); }
#endif
    TRY(rt_malloc((void**)&h->jparts, sizeof(double) * need));
    h->jparts_cap = need;
  }
  h->jparts_used = 0; h->jparts_need = need;
  return rt_memset(h->jparts, 0, sizeof(double) * need, st);
}
// *J_host = J_final_term_host * dt + dt * scale * sum(partials)
static int jparts_finish(smo_kdyn* h, size_t used, double dt, double scale, double* J_host, rt_stream st) {
  SumParams sp;
  sp.partials = h->jparts; sp.out = h->vwork; sp.nwork = 1; sp.nsteps = 1; sp.npart = (int)used; sp.nq = 1; sp.a = dt * scale;
  TRY(launch<FinalSum>(sp, st));
  double part = 0.0;
  TRY(rt_d2h(&part, h->vwork, sizeof(double), st));
  TRY(rt_sync(st));
  *J_host = dt * (*J_host) + part;
  return 0;
}
template <int M> static int kd_forward(smo_kdyn* h, const double* B0, const double* U, double Rm, double dt, int n_iters,
                                       void* snaps, double* J_host, int flags, rt_stream st) {
  const bool integ = (flags & SMO_COST_INTEGRATED) != 0;
  tm_register(h, snaps, (long long)(n_iters + 1) * 3);
  TRY(kd_set_U<M>(h, U, st));
  if (integ) TRY(jparts_begin<M>(h, n_iters, st));
  TRY(KdOps<M>::to_coef(h, B0, h->G, st));
  auto step = [&](int n) { FwdStep f; pingpong(h->G, h->NU, n, f); snap_xs(h, snaps, n, f.xs); return f; };
  // n_iters steps + the x-spectra of the final state (slot n_iters: continuous adjoint, and the cost below)
  TRY((kd_forward_loop<M>(h, n_iters + 1, true, Rm, dt, step, st, integ)));
  const FwdStep fin = step(n_iters);
  cplx* sf[3];
  snap_final(h, snaps, n_iters, sf);
  for (int c = 0; c < 3; ++c) TRY(rt_d2d(sf[c], fin.cin[c], sizeof(cplx) * h->csize, st));
  // Cost "Final": J = mean over the dealiased grid of |B^N|^2 (FWD_Solve_KDyn.py:622, 671-673)
  double* g[3] = {h->gwork, h->gwork + h->gsize, h->gwork + 2 * h->gsize};
  TRY(KdOps<M>::x_c2r(h, fin.xs, g, st));
  const double scale = 1.0 / ((double)M * M * M);
  prof_collect(h, st);
  TRY(smo_vec_dot(h->gwork, h->gwork, (long long)(3 * h->gsize), scale, J_host, h->vwork, (void*)st));
  TRY(kd_check_err(h, "forward solve"));
  return integ ? jparts_finish(h, h->jparts_need, dt, scale, J_host, st) : 0;
}
template <int M> static int kd_prep(smo_kdyn* h, const double* B0, const double* U, double Rm, double dt, int n_iters,
                                    double* out, rt_stream st) {
  TRY(kd_set_U<M>(h, U, st));
  // ping-pong between G and NU as coefficient state (no snapshots kept)
  TRY(KdOps<M>::to_coef(h, B0, h->G, st));
  auto step = [&](int n) { FwdStep f; pingpong(h->G, h->NU, n, f); for (int c = 0; c < 3; ++c) f.xs[c] = h->p2[c]; return f; };
  TRY((kd_forward_loop<M>(h, n_iters + 1, false, Rm, dt, step, st)));
  return KdOps<M>::to_grid(h, step(n_iters + 1).cin, out, st);
}
// Adjoint time loop over `count` forward states, state(i) = x-spectra of the i-th state the sweep linearises about
// (descending in time).  G, the running sum of the gradient integrand (and W = curl G on entry) live in the handle and
// persist between calls, so a checkpointed sweep can call this once per recomputed segment.  3 transforms in, 3 out.
struct XsPtr { cplx* p[3]; };
template <int M, class StateFn>
static int kd_adjoint_loop(smo_kdyn* h, int count, double Rm, double dt, StateFn state, rt_stream st, bool integ = false) {
  if (count <= 0) return 0;
  const bool ks = kernel_sync(h);
  if (ks) TRY(a2a(h, h->p1, h->p1t, 3, st));
  GraphKey key; key.kind = integ ? 12 : 2; key.n = count; key.opts = graph_opts(h); key.p0 = state(0).p[0]; key.p1 = state(count - 1).p[0]; key.p2 = nullptr; key.Rm = Rm; key.dt = dt; key.hsh = 0;
  for (int i = 0; i < count; ++i) key.hsh = hash_ptr(key.hsh, state(i).p[0]);
  return run_graphed(h, key, st, [&](rt_stream q) -> int {
    PdlScope pdl(pdl_on(h));
    TRY(KdOps<M>::inv_z(h, h->W, h->p1, 3, q, XS_B));
    if (!ks) TRY(a2a(h, h->p1, h->p1t, 3, q));
    for (int i = 0; i < count; ++i) {
      const XsPtr bf = state(i);
      TRY(KdOps<M>::yxy(h, 1, bf.p, q, integ));
      if (!ks) TRY(a2a(h, h->p1t, h->p1, 3, q));
      const bool more = i + 1 < count;
      TRY(KdOps<M>::zstep(h, 1, nullptr, nullptr, more, Rm, dt, q, KdOps<M>::yxy_chunks(h, 1)));
      if (more && !ks) TRY(a2a(h, h->p1, h->p1t, 3, q));
    }
    return 0;
  });
}
template <int M> static int kd_adjoint_finish(smo_kdyn* h, double Rm, double dt, int cont, double* gB, double* gU, rt_stream st) {
  TRY(KdOps<M>::nu_finish(h, Rm, dt, st));
  TRY(KdOps<M>::final_scale(h, Rm, dt, cont, st));
  TRY(KdOps<M>::to_grid(h, h->cw, gB, st));
  TRY(KdOps<M>::to_grid(h, h->NU, gU, st));
  prof_collect(h, st);
  return 0;
}
template <int M> static int kd_adjoint(smo_kdyn* h, double Rm, double dt, int n_iters, const void* snapsc, double* gB,
                                       double* gU, int flags, rt_stream st) {
  const int cont = (flags & SMO_ADJOINT_CONTINUOUS) ? 2 : 0;
  const bool integ = (flags & SMO_COST_INTEGRATED) != 0;
  void* snaps = const_cast<void*>(snapsc);
  tm_register(h, snaps, (long long)(n_iters + 1) * 3);
  cplx* s[3];
  snap_final(h, snaps, n_iters, s);
  TRY(KdOps<M>::compat(h, s, Rm, dt, cont | (integ ? 1 : 0), st));
  TRY(KdOps<M>::acc_begin(h, st));
  // adjoint step m linearises about snapshot idx(m): snapshot_index -1-m (continuous) / -2-m (discrete)
  auto state = [&](int m) { XsPtr x; snap_xs(h, snaps, cont ? (n_iters - m) : (n_iters - 1 - m), x.p); return x; };
  TRY((kd_adjoint_loop<M>(h, n_iters, Rm, dt, state, st, integ)));
  return kd_adjoint_finish<M>(h, Rm, dt, cont, gB, gU, st);
}

// ---- checkpointed sweeps (two-level, revolve style) -----------------------------------------------------------------
// The forward solve keeps only the coefficients of the states 0, every, 2*every, ... and of the final state N (slot
// ceil(N/every)); the adjoint walks the segments backwards, recomputing the x-spectra of the states of one segment into a
// buffer of every+1 slots before it sweeps them.  Memory: ceil(N/every) + 1 coefficient states + every + 1 x-spectral
// states instead of N + 1; extra work: at most N forward steps (recompute factor rho <= 1 of one forward solve).
static int ckpt_slots(int n_iters, int every) { return (n_iters + every - 1) / every + 1; }
static int ckpt_slot_of(int n, int n_iters, int every) { return n == n_iters ? ckpt_slots(n_iters, every) - 1 : n / every; }
// coefficient state n of the checkpointing forward solve: its checkpoint slot, or one of two scratch states
static void ckpt_coef(smo_kdyn* h, void* ck, int n, int n_iters, int every, cplx** out) {
  if (n % every == 0 || n == n_iters) { snap_ptrs(h, ck, ckpt_slot_of(n, n_iters, every), out); return; }
  for (int c = 0; c < 3; ++c) out[c] = (n & 1) ? h->NU[c] : h->G[c];
}
template <int M> static int kd_forward_ckpt(smo_kdyn* h, const double* B0, const double* U, double Rm, double dt, int n_iters,
                                            int every, void* ck, double* J_host, int flags, rt_stream st) {
  const bool integ = (flags & SMO_COST_INTEGRATED) != 0;
  TRY(kd_set_U<M>(h, U, st));
  if (integ) TRY(jparts_begin<M>(h, n_iters, st));
  cplx* s0[3];
  snap_ptrs(h, ck, 0, s0);
  TRY(KdOps<M>::to_coef(h, B0, s0, st));
  auto step = [&](int n) {
    FwdStep f; ckpt_coef(h, ck, n, n_iters, every, f.cin); ckpt_coef(h, ck, n + 1, n_iters, every, f.cout);
    for (int c = 0; c < 3; ++c) f.xs[c] = h->p2[c];
    return f;
  };
  TRY((kd_forward_loop<M>(h, n_iters, false, Rm, dt, step, st, integ)));
  snap_ptrs(h, ck, ckpt_slot_of(n_iters, n_iters, every), s0);
  TRY(KdOps<M>::to_grid(h, s0, h->gwork, st));
  const double scale = 1.0 / ((double)M * M * M);
  prof_collect(h, st);
  TRY(smo_vec_dot(h->gwork, h->gwork, (long long)(3 * h->gsize), scale, J_host, h->vwork, (void*)st));
  TRY(kd_check_err(h, "forward solve"));
  return integ ? jparts_finish(h, h->jparts_need, dt, scale, J_host, st) : 0;
}
template <int M> static int kd_adjoint_ckpt(smo_kdyn* h, double Rm, double dt, int n_iters, int every, const void* ckc, void* seg,
                                            double* gB, double* gU, int flags, rt_stream st) {
  const int cont = (flags & SMO_ADJOINT_CONTINUOUS) ? 2 : 0;
  const bool integ = (flags & SMO_COST_INTEGRATED) != 0;
  void* ck = const_cast<void*>(ckc);
  tm_register(h, seg, (long long)(every + 1) * 3);
  cplx* s[3];
  snap_ptrs(h, ck, ckpt_slot_of(n_iters, n_iters, every), s);
  TRY(KdOps<M>::compat(h, s, Rm, dt, cont | (integ ? 1 : 0), st));
  TRY(KdOps<M>::acc_begin(h, st));
  const int nseg = (n_iters + every - 1) / every;
  for (int k = nseg - 1; k >= 0; --k) {
    const int n0 = k * every, n1 = (n0 + every < n_iters) ? n0 + every : n_iters;   // the segment holds the states n0 .. n1
    // states the sweep needs from this segment: n1-1 .. n0 (discrete) or n1 .. n0+1 (continuous).  Recompute from checkpoint
    // n0: full steps up to the last state needed, of which only the x-spectra are produced (tail); scratch coefficients in cw
    const int last = cont ? (n1 - n0) : (n1 - n0 - 1);
    cplx* c0[3];
    snap_ptrs(h, ck, k, c0);
    auto step = [&](int j) {
      FwdStep f;
      for (int c = 0; c < 3; ++c) {
        f.cin[c] = (j == 0) ? c0[c] : h->cw[(j & 1) ? 3 + c : c];
        f.cout[c] = h->cw[(j & 1) ? c : 3 + c];
      }
      snap_xs(h, seg, j, f.xs);
      return f;
    };
    TRY((kd_forward_loop<M>(h, last + 1, true, Rm, dt, step, st)));
    if (k != nseg - 1) TRY(KdOps<M>::curl_G(h, Rm, dt, st));   // W = curl G (the fused step keeps it on chip only)
    auto sweep = [&](int i) { XsPtr x; snap_xs(h, seg, last - i, x.p); return x; };
    TRY((kd_adjoint_loop<M>(h, n1 - n0, Rm, dt, sweep, st, integ)));
  }
  return kd_adjoint_finish<M>(h, Rm, dt, cont, gB, gU, st);
}
// inspection: coefficients of stored state n (tests, SnapshotStore['A_fwd'])
template <int M> static int kd_snapshot_coef(smo_kdyn* h, void* snaps, int n, cplx* const* coef, rt_stream st) {
  cplx* xs[3];
  snap_xs(h, snaps, n, xs);
  if (h->peer_on) TRY(a2a(h, h->p1t, h->p1, 3, st));
  TRY(KdOps<M>::fwd_y(h, xs, h->p1t, 3, st));
  TRY(a2a(h, h->p1t, h->p1, 3, st));
  return KdOps<M>::fwd_z(h, h->p1, coef, 3, st);
}

#define KD_DISPATCH(h, CALL, ...)                                                                  \
  switch ((h)->M) {                                                                                \
    case 24: return CALL<24>(__VA_ARGS__);                                                         \
    case 36: return CALL<36>(__VA_ARGS__);                                                         \
    case 48: return CALL<48>(__VA_ARGS__);                                                         \
    case 96: return CALL<96>(__VA_ARGS__);                                                         \
    case 192: return CALL<192>(__VA_ARGS__);                                                       \
    case 384: return CALL<384>(__VA_ARGS__);                                                       \
    default: return fail(SMO_E_UNSUPPORTED, "kinematic dynamo: Npts=%d not supported", (h)->N);    \
  }
template <int M> static int kd_to_coef(smo_kdyn* h, const double* grid, cplx* const* c, rt_stream st) { return KdOps<M>::to_coef(h, grid, c, st); }
template <int M> static int kd_to_grid(smo_kdyn* h, const cplx* const* c, double* grid, rt_stream st) { return KdOps<M>::to_grid(h, c, grid, st); }

static bool kd_supported(int Npts) { return Npts == 16 || Npts == 24 || Npts == 32 || Npts == 64 || Npts == 128 || Npts == 256; }

extern "C" int smo_kdyn_create(smo_kdyn_t** out, int Npts, double L, int rank, int nranks, void* comm) {
  if (!out) return fail(SMO_E_ARG, "smo_kdyn_create: null handle pointer");
  if (!kd_supported(Npts)) return fail(SMO_E_UNSUPPORTED, "kinematic dynamo: Npts=%d not supported (16, 24, 32, 64, 128, 256)", Npts);
  if (!(L > 0) || nranks < 1 || rank < 0 || rank >= nranks) return fail(SMO_E_ARG, "smo_kdyn_create: bad L/rank/nranks");
  const int M = 3 * Npts / 2, Nh = Npts / 2;
  if (Nh % nranks || M % nranks) return fail(SMO_E_ARG, "nranks=%d must divide Npts/2=%d and 3*Npts/2=%d", nranks, Nh, M);
  if (nranks > 1 && !comm) return fail(SMO_E_ARG, "smo_kdyn_create: nranks > 1 needs a communicator");
  const int nz = M / nranks;
  if (((long long)M * nz) % SMO_TX || ((long long)M * nz) % SMO_TXA)
    return fail(SMO_E_ARG, "local grid columns M*nz = %lld must be a multiple of the x-pass tiles (%d, %d)", (long long)M * nz, SMO_TX, SMO_TXA);
  smo_kdyn* h = new smo_kdyn();
  h->N = Npts; h->M = M; h->Nh = Nh; h->kmax = (Npts - 1) / 2; h->Nc = 2 * h->kmax + 1; h->Pc = h->Nc + 1;
  h->L = L; h->kfac = 2.0 * 3.14159265358979323846 / L;
  h->rank = rank; h->nranks = nranks; h->nkx = Nh / nranks; h->kx0 = rank * h->nkx; h->nz = nz; h->z0 = rank * nz;
  h->csize = (size_t)h->nkx * h->Nc * h->Pc;
  h->p1size = (size_t)h->nkx * h->Nc * M;
  h->p2size = (size_t)Nh * M * nz;
  h->gsize = (size_t)M * M * nz;
  h->comm = comm;
  h->have_U = false;
  h->prof_which = 0; h->prof_ms = 0; h->prof_n = 0; h->use_graph = 0;
  h->jparts = nullptr; h->jparts_cap = 0; h->jparts_used = 0; h->jparts_need = 0;
  h->graphs = nullptr; h->capturing = 0; h->cap_a0 = h->cap_b0 = 0; h->epoch_dev = nullptr;
  h->peer_on = 0; h->flags = nullptr; h->epoch = 0;
  h->inkernel_sync = 1; h->epochA = h->epochB = 0; h->counters = nullptr; h->peer_pull = 0; h->l2_hints = 1;
  h->tma_sin = 1; h->tmc = nullptr; h->p2block = nullptr;     // (r2k: x-adj 196 -> 192 us at 128^3, 2.11 -> 1.94 ms at 256^3)
  h->grid_acc = 1; h->accg = nullptr; h->bulk_u = 1;     // (r2e: adjoint x pass 221 -> 196 us at 128^3, 2.46 -> 2.04 ms at 256^3)
  h->push_waves = 1; h->two_streams = 0; h->err_host = nullptr; h->err_dev = nullptr;
  h->pdl = SMO_PDL_DEFAULT; h->bulk_push = SMO_BULK_PUSH_DEFAULT;
#if !defined(SMO_EMUL)
  h->aux_stream = nullptr; h->ev_fork = nullptr; h->ev_join = nullptr;
#endif
  for (int f = 0; f < MAXF; ++f) for (int s2 = 0; s2 < MAXP; ++s2) { h->peer_p1[f][s2] = h->peer_p1t[f][s2] = nullptr; }
  for (int s2 = 0; s2 < MAXP; ++s2) h->peer_flags[s2] = nullptr;
  h->chunks_fwd = h->chunks_adj = 1;    // off by default (measured slower at 128^3: the passes are not HBM-bound enough to gain)
  h->hB = h->hU = h->hGB = h->hGU = nullptr; h->snaps = nullptr; h->cap_snap = 0;
  h->tw = nullptr; h->gwork = nullptr; h->vwork = nullptr;
  for (int f = 0; f < MAXF; ++f) h->p1[f] = h->p1t[f] = h->p2[f] = h->cw[f] = nullptr;
  for (int c = 0; c < 3; ++c) { h->G[c] = h->NU[c] = h->W[c] = nullptr; h->Ug[c] = nullptr; h->acc[c] = nullptr; }
  h->Ut = nullptr;
#if !defined(SMO_EMUL)
  h->ev = new std::vector<cudaEvent_t>(); h->ev_used = 0;
#endif
  int rc = 0;
  h->tw = make_twiddles(M, &rc);
  for (int f = 0; f < MAXF && rc == 0; ++f) {
    rc = rt_malloc((void**)&h->p1[f], sizeof(cplx) * h->p1size);
    if (rc == 0) {
      if (nranks > 1) rc = rt_malloc((void**)&h->p1t[f], sizeof(cplx) * h->p1size);
      else h->p1t[f] = h->p1[f];
    }
    if (rc == 0 && f == 0) rc = rt_malloc((void**)&h->p2block, sizeof(cplx) * h->p2size * MAXF);
    if (rc == 0) h->p2[f] = h->p2block + (size_t)f * h->p2size;
    if (rc == 0) rc = rt_malloc((void**)&h->cw[f], sizeof(cplx) * h->csize);
  }
  for (int c = 0; c < 3 && rc == 0; ++c) {
    rc = rt_malloc((void**)&h->G[c], sizeof(cplx) * h->csize);
    if (rc == 0) rc = rt_malloc((void**)&h->NU[c], sizeof(cplx) * h->csize);
    if (rc == 0) rc = rt_malloc((void**)&h->W[c], sizeof(cplx) * h->csize);
    if (rc == 0) rc = rt_malloc((void**)&h->Ug[c], sizeof(double) * h->gsize);
  }
  if (rc == 0) tm_register(h, h->p2block, MAXF);
  if (rc == 0) rc = rt_malloc((void**)&h->Ut, sizeof(double) * 3 * h->gsize);
  if (rc == 0) rc = rt_malloc((void**)&h->gwork, sizeof(double) * 3 * h->gsize);
  if (rc == 0) rc = rt_malloc((void**)&h->vwork, smo_vec_work_bytes((long long)(3 * h->gsize)));
  if (rc) { smo_kdyn_destroy(h); return rc; }
  *out = h;
  return 0;
}
extern "C" int smo_kdyn_destroy(smo_kdyn_t* h) {
  if (!h) return 0;
  rt_free(h->tw);
  for (int f = 0; f < MAXF; ++f) {
    if (h->nranks > 1) rt_free(h->p1t[f]);
    rt_free(h->p1[f]); rt_free(h->cw[f]);
  }
  for (int c = 0; c < 3; ++c) { rt_free(h->G[c]); rt_free(h->NU[c]); rt_free(h->W[c]); rt_free(h->Ug[c]); rt_free(h->acc[c]); }
  rt_free(h->gwork); rt_free(h->vwork); rt_free(h->Ut); rt_free(h->jparts); rt_free(h->accg); rt_free(h->p2block);
  tm_destroy(h);
#if !defined(SMO_EMUL)
  if (h->peer_on) {
    for (int s = 0; s < h->nranks; ++s) {
      if (s == h->rank) continue;
      for (int f = 0; f < MAXF; ++f) { cudaIpcCloseMemHandle(h->peer_p1[f][s]); cudaIpcCloseMemHandle(h->peer_p1t[f][s]); }
      cudaIpcCloseMemHandle(h->peer_flags[s]);
    }
  }
  rt_free(h->flags); rt_free(h->counters); rt_free(h->epoch_dev);
  if (h->err_host) cudaFreeHost(h->err_host);
  if (h->aux_stream) { cudaStreamDestroy(h->aux_stream); cudaEventDestroy(h->ev_fork); cudaEventDestroy(h->ev_join); for (int k = 0; k < 3; ++k) cudaEventDestroy(h->ev_pipe[k]); }
  if (h->graphs) {
    for (GraphEntry& e : h->graphs->e) if (e.exec) cudaGraphExecDestroy(e.exec);
    cudaStreamDestroy(h->graphs->stream); cudaEventDestroy(h->graphs->ev0); cudaEventDestroy(h->graphs->ev1);
    delete h->graphs;
  }
#endif
  rt_free(h->hB); rt_free(h->hU); rt_free(h->hGB); rt_free(h->hGU); rt_free(h->snaps);
#if !defined(SMO_EMUL)
  if (h->ev) { for (cudaEvent_t e : *h->ev) cudaEventDestroy(e); delete h->ev; }
#endif
  delete h;
  return 0;
}
extern "C" size_t smo_kdyn_grid_elems(const smo_kdyn_t* h) { return h ? h->gsize : 0; }
extern "C" size_t smo_kdyn_coef_elems(const smo_kdyn_t* h) { return h ? h->csize : 0; }
extern "C" size_t smo_kdyn_snapshot_bytes(const smo_kdyn_t* h, int n_iters) {
  return h ? sizeof(cplx) * 3 * (h->p2size * (size_t)(n_iters + 1) + h->csize) : 0;
}
extern "C" size_t smo_kdyn_segment_bytes(const smo_kdyn_t* h, int every) {
  return h ? sizeof(cplx) * 3 * h->p2size * (size_t)(every + 1) : 0;
}
static int kd_args(smo_kdyn* h, double Rm, double dt, int n_iters, int flags, const char* who) {
  if (!h) return fail(SMO_E_ARG, "%s: null handle", who);
  TRY(kd_check_err(h, who));
  if (!(Rm > 0) || !(dt > 0) || n_iters < 0) return fail(SMO_E_ARG, "%s: bad Rm/dt/n_iters", who);
  (void)flags;
  return 0;
}
extern "C" int smo_kdyn_forward(smo_kdyn_t* h, const double* B0, const double* U, double Rm, double dt, int n_iters,
                                void* snaps, double* J_host, int flags, void* stream) {
  TRY(kd_args(h, Rm, dt, n_iters, flags, "smo_kdyn_forward"));
  if (!B0 || !U || !snaps || !J_host) return fail(SMO_E_ARG, "smo_kdyn_forward: null buffer");
  rt_stream st = (rt_stream)stream;
  KD_DISPATCH(h, kd_forward, h, B0, U, Rm, dt, n_iters, snaps, J_host, flags, st)
}
extern "C" int smo_kdyn_prep(smo_kdyn_t* h, const double* B0, const double* U, double Rm, double dt, int n_iters,
                             double* out, void* stream) {
  TRY(kd_args(h, Rm, dt, n_iters, 0, "smo_kdyn_prep"));
  if (!B0 || !U || !out) return fail(SMO_E_ARG, "smo_kdyn_prep: null buffer");
  rt_stream st = (rt_stream)stream;
  KD_DISPATCH(h, kd_prep, h, B0, U, Rm, dt, n_iters, out, st)
}
extern "C" int smo_kdyn_adjoint(smo_kdyn_t* h, double Rm, double dt, int n_iters, const void* snaps, double* gB,
                                double* gU, int flags, void* stream) {
  TRY(kd_args(h, Rm, dt, n_iters, flags, "smo_kdyn_adjoint"));
  if (!snaps || !gB || !gU) return fail(SMO_E_ARG, "smo_kdyn_adjoint: null buffer");
  if (!h->have_U) return fail(SMO_E_STATE, "smo_kdyn_adjoint: no preceding forward solve on this handle");
  rt_stream st = (rt_stream)stream;
  KD_DISPATCH(h, kd_adjoint, h, Rm, dt, n_iters, snaps, gB, gU, flags, st)
}
extern "C" size_t smo_kdyn_checkpoint_bytes(const smo_kdyn_t* h, int n_iters, int every) {
  return (h && every > 0 && n_iters >= 0) ? sizeof(cplx) * 3 * h->csize * (size_t)ckpt_slots(n_iters, every) : 0;
}
extern "C" int smo_kdyn_forward_ckpt(smo_kdyn_t* h, const double* B0, const double* U, double Rm, double dt, int n_iters,
                                     int every, void* ckpt, double* J_host, int flags, void* stream) {
  TRY(kd_args(h, Rm, dt, n_iters, flags, "smo_kdyn_forward_ckpt"));
  if (!B0 || !U || !ckpt || !J_host || every < 1) return fail(SMO_E_ARG, "smo_kdyn_forward_ckpt: bad argument");
  rt_stream st = (rt_stream)stream;
  KD_DISPATCH(h, kd_forward_ckpt, h, B0, U, Rm, dt, n_iters, every, ckpt, J_host, flags, st)
}
extern "C" int smo_kdyn_adjoint_ckpt(smo_kdyn_t* h, double Rm, double dt, int n_iters, int every, const void* ckpt, void* seg,
                                     double* gB, double* gU, int flags, void* stream) {
  TRY(kd_args(h, Rm, dt, n_iters, flags, "smo_kdyn_adjoint_ckpt"));
  if (!ckpt || !seg || !gB || !gU || every < 1) return fail(SMO_E_ARG, "smo_kdyn_adjoint_ckpt: bad argument");
  if (!h->have_U) return fail(SMO_E_STATE, "smo_kdyn_adjoint_ckpt: no preceding forward solve on this handle");
  rt_stream st = (rt_stream)stream;
  KD_DISPATCH(h, kd_adjoint_ckpt, h, Rm, dt, n_iters, every, ckpt, seg, gB, gU, flags, st)
}
extern "C" int smo_kdyn_snapshot_coef(smo_kdyn_t* h, const void* snaps, int n_iters, int n, void* coef, void* stream) {
  if (!h || !snaps || !coef || n < 0 || n > n_iters) return fail(SMO_E_ARG, "smo_kdyn_snapshot_coef: bad argument");
  rt_stream st = (rt_stream)stream;
  cplx* c[3] = {(cplx*)coef, (cplx*)coef + h->csize, (cplx*)coef + 2 * h->csize};
  KD_DISPATCH(h, kd_snapshot_coef, h, const_cast<void*>(snaps), n, c, st)
}
extern "C" int smo_kdyn_to_coef(smo_kdyn_t* h, const double* grid, void* coef, void* stream) {
  if (!h || !grid || !coef) return fail(SMO_E_ARG, "smo_kdyn_to_coef: bad argument");
  rt_stream st = (rt_stream)stream;
  cplx* c[3] = {(cplx*)coef, (cplx*)coef + h->csize, (cplx*)coef + 2 * h->csize};
  KD_DISPATCH(h, kd_to_coef, h, grid, c, st)
}
extern "C" int smo_kdyn_to_grid(smo_kdyn_t* h, const void* coef, double* grid, void* stream) {
  if (!h || !grid || !coef) return fail(SMO_E_ARG, "smo_kdyn_to_grid: bad argument");
  rt_stream st = (rt_stream)stream;
  const cplx* c[3] = {(const cplx*)coef, (const cplx*)coef + h->csize, (const cplx*)coef + 2 * h->csize};
  KD_DISPATCH(h, kd_to_grid, h, c, grid, st)
}
extern "C" int smo_kdyn_profile_set(smo_kdyn_t* h, int which) {
  if (!h) return fail(SMO_E_ARG, "smo_kdyn_profile_set: null handle");
  h->prof_which = which; h->prof_ms = 0; h->prof_n = 0;
#if !defined(SMO_EMUL)
  h->ev_used = 0;
#endif
  return 0;
}
extern "C" int smo_kdyn_profile_read(smo_kdyn_t* h, double* total_ms, long long* launches) {
  if (!h) return fail(SMO_E_ARG, "smo_kdyn_profile_read: null handle");
  if (total_ms) *total_ms = h->prof_ms;
  if (launches) *launches = h->prof_n;
  return 0;
}
// ---- peer-memory attachment (CUDA IPC) --------------------------------------------------------------------------
#if !defined(SMO_EMUL)
extern "C" int smo_kdyn_peer_handle_bytes(void) { return (int)((2 * MAXF + 1) * sizeof(cudaIpcMemHandle_t)); }
extern "C" int smo_kdyn_peer_export(smo_kdyn_t* h, void* out) {
  if (!h || !out) return fail(SMO_E_ARG, "smo_kdyn_peer_export: bad argument");
  if (h->nranks < 2) return fail(SMO_E_ARG, "smo_kdyn_peer_export: single-rank handle");
  if (h->nranks > MAXP) return fail(SMO_E_UNSUPPORTED, "peer transposes support at most %d ranks", MAXP);
  if (!h->flags) {
    TRY(rt_malloc((void**)&h->flags, sizeof(unsigned long long) * NFLAGW));
    TRY(rt_malloc((void**)&h->counters, sizeof(unsigned int) * 3 * MAXCH));
    CUDA_TRY(cudaHostAlloc((void**)&h->err_host, sizeof(unsigned int), cudaHostAllocMapped));
    *h->err_host = 0u;
    CUDA_TRY(cudaHostGetDevicePointer((void**)&h->err_dev, h->err_host, 0));
    CUDA_TRY(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t* hd = (cudaIpcMemHandle_t*)out;
  for (int f = 0; f < MAXF; ++f) {
    CUDA_TRY(cudaIpcGetMemHandle(&hd[f], h->p1[f]));
    CUDA_TRY(cudaIpcGetMemHandle(&hd[MAXF + f], h->p1t[f]));
  }
  CUDA_TRY(cudaIpcGetMemHandle(&hd[2 * MAXF], h->flags));
  return 0;
}
extern "C" int smo_kdyn_peer_attach(smo_kdyn_t* h, const void* all) {
  if (!h || !all) return fail(SMO_E_ARG, "smo_kdyn_peer_attach: bad argument");
  if (h->nranks < 2 || h->nranks > MAXP || !h->flags) return fail(SMO_E_STATE, "smo_kdyn_peer_attach: export first (2..%d ranks)", MAXP);
  const cudaIpcMemHandle_t* hd = (const cudaIpcMemHandle_t*)all;
  const int per = 2 * MAXF + 1;
  for (int s = 0; s < h->nranks; ++s) {
    if (s == h->rank) {
      for (int f = 0; f < MAXF; ++f) { h->peer_p1[f][s] = h->p1[f]; h->peer_p1t[f][s] = h->p1t[f]; }
      h->peer_flags[s] = h->flags;
      continue;
    }
    for (int f = 0; f < MAXF; ++f) {
      void* q = nullptr;
      CUDA_TRY(cudaIpcOpenMemHandle(&q, hd[s * per + f], cudaIpcMemLazyEnablePeerAccess));
      h->peer_p1[f][s] = (cplx*)q;
      CUDA_TRY(cudaIpcOpenMemHandle(&q, hd[s * per + MAXF + f], cudaIpcMemLazyEnablePeerAccess));
      h->peer_p1t[f][s] = (cplx*)q;
    }
    void* q = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&q, hd[s * per + 2 * MAXF], cudaIpcMemLazyEnablePeerAccess));
    h->peer_flags[s] = (unsigned long long*)q;
  }
  h->peer_on = 1;
  return 0;
}
#else
extern "C" int smo_kdyn_peer_handle_bytes(void) { return 64 * (2 * MAXF + 1); }
extern "C" int smo_kdyn_peer_export(smo_kdyn_t*, void*) { return fail(SMO_E_UNSUPPORTED, "no peer memory in the host emulation"); }
extern "C" int smo_kdyn_peer_attach(smo_kdyn_t*, const void*) { return fail(SMO_E_UNSUPPORTED, "no peer memory in the host emulation"); }
#endif

extern "C" int smo_kdyn_set_chunks(smo_kdyn_t* h, int chunks_fwd, int chunks_adj) {
  if (!h) return fail(SMO_E_ARG, "smo_kdyn_set_chunks: null handle");
  h->chunks_fwd = chunks_fwd; h->chunks_adj = chunks_adj;
  return 0;
}
extern "C" int smo_kdyn_set_option(smo_kdyn_t* h, int key, int value) {
  if (!h) return fail(SMO_E_ARG, "smo_kdyn_set_option: null handle");
  switch (key) {
    case SMO_OPT_KERNEL_SYNC: h->inkernel_sync = value ? 1 : 0; return 0;
    case SMO_OPT_PEER_PULL: h->peer_pull = value ? 1 : 0; return 0;
    case SMO_OPT_L2_HINTS: h->l2_hints = value ? 1 : 0; return 0;
    case SMO_OPT_PUSH_WAVES: h->push_waves = value < 1 ? 1 : value; return 0;
    case SMO_OPT_TWO_STREAMS: h->two_streams = value < 0 ? 0 : (value > 2 ? 2 : value); return 0;
    case SMO_OPT_GRID_ACC: h->grid_acc = value ? 1 : 0; return 0;
    case SMO_OPT_BULK_U: h->bulk_u = value ? 1 : 0; return 0;
    case SMO_OPT_TMA_SIN: h->tma_sin = value ? 1 : 0; return 0;
    case SMO_OPT_BULK_PUSH: h->bulk_push = value & 3; return 0;
    case SMO_OPT_PDL: h->pdl = value < 0 ? -1 : (value ? 1 : 0); return 0;
    case 99:   // development only (WRONG RESULTS): point every peer buffer at the local one to time the kernels without NVLink traffic
      for (int f = 0; f < MAXF; ++f) for (int s2 = 0; s2 < h->nranks; ++s2) { h->peer_p1[f][s2] = h->p1[f]; h->peer_p1t[f][s2] = h->p1t[f]; }
      return 0;
    default: return fail(SMO_E_ARG, "smo_kdyn_set_option: unknown key %d", key);
  }
}
extern "C" int smo_kdyn_use_graph(smo_kdyn_t* h, int on) {
  if (!h) return fail(SMO_E_ARG, "smo_kdyn_use_graph: null handle");
  h->use_graph = on;
  return 0;
}

// host-buffer forms --------------------------------------------------------------------------------------------
static int kd_reserve_host(smo_kdyn* h, int n_iters) {
  const size_t vb = sizeof(double) * 3 * h->gsize;
  if (!h->hB) {
    TRY(rt_malloc((void**)&h->hB, vb));
    TRY(rt_malloc((void**)&h->hU, vb));
    TRY(rt_malloc((void**)&h->hGB, vb));
    TRY(rt_malloc((void**)&h->hGU, vb));
  }
  const size_t need = n_iters < 0 ? 0 : smo_kdyn_snapshot_bytes(h, n_iters);   // n_iters < 0: caller's store
  if (need > h->cap_snap) {
    rt_free(h->snaps); h->snaps = nullptr; h->cap_snap = 0;
    TRY(rt_malloc((void**)&h->snaps, need));
    h->cap_snap = need;
  }
  return 0;
}
// full host vector [3][M][M][M]  <->  local slab [3][M][M][nz]
static int kd_slab_copy(smo_kdyn* h, double* dev, const double* host_in, double* host_out, rt_stream st) {
  const size_t M = h->M, rows = 3 * M * M;
  if (h->nranks == 1) {
    if (host_in) return rt_h2d(dev, host_in, sizeof(double) * 3 * h->gsize, st);
    return rt_d2h(host_out, dev, sizeof(double) * 3 * h->gsize, st);
  }
  if (host_in) return rt_copy2d(dev, h->nz * sizeof(double), host_in + h->z0, M * sizeof(double), h->nz * sizeof(double), rows, 0, st);
  return rt_copy2d(host_out + h->z0, M * sizeof(double), dev, h->nz * sizeof(double), h->nz * sizeof(double), rows, 1, st);
}
// z-slab of this rank <-> full reference vector on the host (strided 2-D copy; the host-side numpy slicing it replaces cost
// more than the copy itself at 8 ranks).  Exactly one of host_in / host_out is given.  Asynchronous on `stream`.
extern "C" int smo_kdyn_slab_copy(smo_kdyn_t* h, double* slab_dev, const double* host_in, double* host_out, void* stream) {
  if (!h || !slab_dev || (!host_in) == (!host_out)) return fail(SMO_E_ARG, "smo_kdyn_slab_copy: bad argument");
  return kd_slab_copy(h, slab_dev, host_in, host_out, (rt_stream)stream);
}
extern "C" int smo_kdyn_forward_host(smo_kdyn_t* h, const double* B0, const double* U, double Rm, double dt,
                                     int n_iters, void* snaps, double* J_host, int flags, void* stream) {
  TRY(kd_args(h, Rm, dt, n_iters, flags, "smo_kdyn_forward_host"));
  if (!B0 || !U || !J_host) return fail(SMO_E_ARG, "smo_kdyn_forward_host: null buffer");
  rt_stream st = (rt_stream)stream;
  TRY(kd_reserve_host(h, snaps ? -1 : n_iters));
  TRY(kd_slab_copy(h, h->hB, B0, nullptr, st));
  TRY(kd_slab_copy(h, h->hU, U, nullptr, st));
  return smo_kdyn_forward(h, h->hB, h->hU, Rm, dt, n_iters, snaps ? snaps : h->snaps, J_host, flags, stream);
}
extern "C" int smo_kdyn_adjoint_host(smo_kdyn_t* h, double Rm, double dt, int n_iters, const void* snaps, double* gB,
                                     double* gU, int flags, void* stream) {
  TRY(kd_args(h, Rm, dt, n_iters, flags, "smo_kdyn_adjoint_host"));
  if (!gB || !gU) return fail(SMO_E_ARG, "smo_kdyn_adjoint_host: null buffer");
  if (!snaps && (!h->snaps || smo_kdyn_snapshot_bytes(h, n_iters) > h->cap_snap))
    return fail(SMO_E_STATE, "smo_kdyn_adjoint_host: no matching forward solve on this handle");
  rt_stream st = (rt_stream)stream;
  TRY(kd_reserve_host(h, -1));
  TRY(smo_kdyn_adjoint(h, Rm, dt, n_iters, snaps ? snaps : h->snaps, h->hGB, h->hGU, flags, stream));
  TRY(kd_slab_copy(h, h->hGB, nullptr, gB, st));
  TRY(kd_slab_copy(h, h->hGU, nullptr, gU, st));
  return rt_sync(st);
}
extern "C" int smo_kdyn_prep_host(smo_kdyn_t* h, const double* B0, const double* U, double Rm, double dt, int n_iters,
                                  double* out, void* stream) {
  TRY(kd_args(h, Rm, dt, n_iters, 0, "smo_kdyn_prep_host"));
  if (!B0 || !U || !out) return fail(SMO_E_ARG, "smo_kdyn_prep_host: null buffer");
  rt_stream st = (rt_stream)stream;
  TRY(kd_reserve_host(h, -1));
  TRY(kd_slab_copy(h, h->hB, B0, nullptr, st));
  TRY(kd_slab_copy(h, h->hU, U, nullptr, st));
  TRY(smo_kdyn_prep(h, h->hB, h->hU, Rm, dt, n_iters, h->hGB, stream));
  TRY(kd_slab_copy(h, h->hGB, nullptr, out, st));
  return rt_sync(st);
}

// communicator ------------------------------------------------------------------------------------------------
#if !defined(SMO_EMUL) && defined(SMO_WITH_NCCL)
extern "C" int smo_comm_unique_id_bytes(void) { return (int)sizeof(ncclUniqueId); }
extern "C" int smo_comm_get_unique_id(void* id_out) {
  if (!id_out) return fail(SMO_E_ARG, "smo_comm_get_unique_id: null buffer");
  NcclApi* nc = nccl_api();
  if (!nc) return fail(SMO_E_COMM, "libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  ncclResult_t r = nc->GetUniqueId(&id);
  if (r != ncclSuccess) return fail(SMO_E_COMM, "ncclGetUniqueId failed: %s", nc->GetErrorString(r));
  memcpy(id_out, &id, sizeof id);
  return 0;
}
extern "C" int smo_comm_create(void** comm, const void* id_in, int nranks, int rank) {
  if (!comm || !id_in || nranks < 1 || rank < 0 || rank >= nranks) return fail(SMO_E_ARG, "smo_comm_create: bad argument");
  NcclApi* nc = nccl_api();
  if (!nc) return fail(SMO_E_COMM, "libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  memcpy(&id, id_in, sizeof id);
  ncclComm_t c;
  ncclResult_t r = nc->CommInitRank(&c, nranks, id, rank);
  if (r != ncclSuccess) return fail(SMO_E_COMM, "ncclCommInitRank failed: %s", nc->GetErrorString(r));
  *comm = (void*)c;
  return 0;
}
extern "C" int smo_comm_destroy(void* comm) {
  NcclApi* nc = nccl_api();
  if (comm && nc) nc->CommDestroy((ncclComm_t)comm);
  return 0;
}
#else
extern "C" int smo_comm_unique_id_bytes(void) { return 128; }
extern "C" int smo_comm_get_unique_id(void* id_out) {
  if (!id_out) return fail(SMO_E_ARG, "smo_comm_get_unique_id: null buffer");
  memset(id_out, 0, 128);
  return 0;
}
extern "C" int smo_comm_create(void**, const void*, int, int) {
  return fail(SMO_E_UNSUPPORTED, "library built without NCCL");
}
extern "C" int smo_comm_destroy(void*) { return 0; }
#endif

#if defined(SMO_EMUL)
// test-only: communicator made of a host callback (tests/emul drives it with torch.distributed gloo)
extern "C" void* smo_emul_make_comm(smo_emul_a2a_fn fn, void* user) {
  EmulComm* c = new EmulComm();
  c->fn = fn; c->user = user;
  return c;
}
#endif
