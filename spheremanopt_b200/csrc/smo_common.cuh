// Common definitions for the spheremanopt_b200 CUDA library (sm_100a, fp64).
//
// Kernel bodies are written as a fixed sequence of *phases* separated by CTA-wide barriers:
//
//     struct K { struct Params; struct State; static constexpr int THREADS, NPHASES, MIN_BLOCKS;
//                template <int PH> SMO_HD static void phase(const Params&, int work, int tid,
//                                                           unsigned char* smem, State&); };
//
// On the device, smo_kernel<K> runs the phases of one work item (tile) per loop trip with
// __syncthreads() between them; State lives in registers.  With -DSMO_EMUL the very same phase
// bodies are compiled by g++ and run thread-by-thread on the host (tests/emul): that build is test
// infrastructure for checking index logic without a GPU and is never loaded by the product package.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <cmath>

#if defined(SMO_EMUL)
#define SMO_HD inline
#define SMO_DEV inline
#include <vector>
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
struct alignas(64) SmoTensorMap { unsigned long long opaque[16]; };
#else
#include <cuda_runtime.h>
#include <cuda.h>
#define SMO_HD __host__ __device__ __forceinline__
#define SMO_DEV __device__ __forceinline__
typedef CUtensorMap SmoTensorMap;     // TMA descriptor (cuTensorMapEncodeTiled), passed to kernels by value inside their Params
#endif

namespace smo {

typedef double2 cplx;   // interleaved (re, im)

constexpr int MAXF = 6;  // max fields handled by one launch
constexpr int MAXP = 8;  // max ranks (GPUs of one box) of the slab decomposition

template <class T> SMO_HD T ldg(const T* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}
SMO_HD cplx ldg_c(const cplx* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

// 16-byte asynchronous global -> shared copy (LDGSTS, L2-only caching): the producer side of the software pipelines.
// Host emulation: an immediate copy.
SMO_HD void cp_async16(void* sdst, const void* gsrc) {
#if defined(__CUDA_ARCH__)
  const unsigned s = (unsigned)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
#else
  memcpy(sdst, gsrc, 16);
#endif
}
// fire-and-forget prefetch of the 128-byte line holding p into L2 (hides the HBM latency of a later plain load)
SMO_HD void prefetch_l2(const void* p) {
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}
// L2 residency hints.  Between the y pass that writes the z-padded pencils, the fused z step that rewrites them and the y
// pass that reads them back, 75 MB (128^3) of pencil data is produced and consumed within ~100 us: stored with evict_last and
// read with evict_first it can stay in the 126 MB L2 instead of making two round trips through HBM, provided the streams
// that are only touched once (x-spectra, coefficient states) are marked evict_first and do not push it out.
// kind: 1 = evict_first, 2 = evict_last.
SMO_HD unsigned long long l2_policy(int kind) {
#if defined(__CUDA_ARCH__)
  unsigned long long pol;
  if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
#else
  (void)kind;
  return 0ull;
#endif
}
SMO_HD void cp_async16_hint(void* sdst, const void* gsrc, unsigned long long pol) {
#if defined(__CUDA_ARCH__)
  const unsigned s = (unsigned)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(s), "l"(gsrc), "l"(pol) : "memory");
#else
  (void)pol;
  memcpy(sdst, gsrc, 16);
#endif
}
SMO_HD void st_cplx_hint(cplx* p, double x, double y, unsigned long long pol) {
#if defined(__CUDA_ARCH__)
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(x), "d"(y), "l"(pol) : "memory");
#else
  (void)pol;
  p->x = x; p->y = y;
#endif
}
// store of one complex value into a PEER GPU's memory (fused transposes).  Measured on 2 B200 (r2zb): st.relaxed.sys and st.wt
// flavours run exactly like the plain store; a system fence after every tile costs +35 %.
SMO_HD void st_peer(cplx* p, double x, double y) { p->x = x; p->y = y; }
SMO_HD void cp_async_commit() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N> SMO_HD void cp_async_wait() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

// TMA bulk copy (cp.async.bulk, 1-D) global -> shared with mbarrier completion: ONE thread moves a whole contiguous tile, the
// bytes travel through the async proxy instead of the LSU / L1 data pipe that the 16-byte cp.async (LDGSTS) copies share with
// every shared-memory load and store of the FFTs.  Host emulation: an immediate memcpy, waits are no-ops.
SMO_HD void mbar_init(unsigned long long* bar, int count) {
#if defined(__CUDA_ARCH__)
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#else
  (void)bar; (void)count;
#endif
}
// called by one thread: arms the barrier with the byte count and starts the copy (bytes: multiple of 16, both sides 16-byte aligned)
SMO_HD void bulk_load(void* sdst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
#if defined(__CUDA_ARCH__)
  const unsigned ba = (unsigned)__cvta_generic_to_shared(bar), sa = (unsigned)__cvta_generic_to_shared(sdst);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa), "l"(gsrc), "r"(bytes), "r"(ba) : "memory");
#else
  (void)bar;
  memcpy(sdst, gsrc, bytes);
#endif
}
// arrive on the barrier and add `bytes` to the transaction count it waits for (one call per issuing thread)
SMO_HD void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
#if defined(__CUDA_ARCH__)
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
#else
  (void)bar; (void)bytes;
#endif
}
// TMA tensor copy (cp.async.bulk.tensor, 3-D tiled map) global -> shared: one thread moves a whole box {c0.., c1.., c2} described
// by the tensor map (row pitch, swizzle and bounds live in the descriptor); completion is signalled on the mbarrier.
SMO_HD void tma_load_3d(void* sdst, const SmoTensorMap* tm, int c0, int c1, int c2, unsigned long long* bar) {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(reinterpret_cast<unsigned long long>(tm)),
                 "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
#else
  (void)sdst; (void)tm; (void)c0; (void)c1; (void)c2; (void)bar;
#endif
}
// TMA bulk STORE shared -> global (bulk async-group completion): one thread ships a contiguous staged block - also into a PEER
// GPU's memory over NVLink (any generic global address) - through the async proxy.  The transposes pushed by the time-loop
// kernels use it so that the remote stores do not sit in the SM's LSU queue in front of the local loads of the next tile.
// Protocol: writers st.shared -> bulk_fence_smem() -> barrier -> issuer bulk_store()... bulk_commit(); the staging area may be
// rewritten after bulk_wait_read(); the data is complete in (peer) memory after bulk_wait_all() + bulk_fence_global().
SMO_HD void bulk_store(void* gdst, const void* ssrc, unsigned bytes) {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
#else
  memcpy(gdst, ssrc, bytes);
#endif
}
SMO_HD void warp_sync() {
#if defined(__CUDA_ARCH__)
  __syncwarp();
#endif
}
SMO_HD void bulk_commit() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#endif
}
SMO_HD void bulk_wait_read() {      // the committed bulk stores of this thread have finished READING shared memory
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
}
SMO_HD void bulk_wait_all() {       // ... have completed (their writes are performed)
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#endif
}
SMO_HD void bulk_fence_smem() {     // generic-proxy writes to shared memory -> visible to the async proxy (TMA)
#if defined(__CUDA_ARCH__)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}
SMO_HD void bulk_fence_global() {   // async-proxy writes to global memory -> ordered before this thread's later generic-proxy operations
#if defined(__CUDA_ARCH__)
  asm volatile("fence.proxy.async.global;" ::: "memory");
#endif
}
SMO_HD void mbar_wait(unsigned long long* bar, unsigned parity) {
#if defined(__CUDA_ARCH__)
  const unsigned ba = (unsigned)__cvta_generic_to_shared(bar);
  unsigned ok = 0;
  while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(ba), "r"(parity) : "memory");
#else
  (void)bar; (void)parity;
#endif
}

// Cross-GPU hand-shake fused into a kernel (peer-memory transposes of the slab-decomposed dynamo).  A kernel whose
// Params carry an `XSync xs` member
//   * first waits until every source rank has published `wait_epoch` in this GPU's flag words (its inputs were stored
//     into this GPU's memory by the peers' preceding kernel), and
//   * after its last CTA has finished, publishes `sig_epoch` in every peer's flag word of this rank (its stores are
//     complete and visible: per-CTA device-scope fence, completion counter, ONE system-scope fence in the last CTA -
//     cumulative over all it acquired through the counter -, flag stores).  sig_sys = 1: the data was stored into the
//     peers' memory (push); 0: the data is local and the peers will read it through this GPU's L2 (pull).
// Measured on 2 B200 (tools/microbench/flag_latency.cu): flag word one way 1.0 us, every system fence +1 us (idle) to
// +3.4 us (stores in flight); so the waiting side uses no fence at all - its loads are issued after the spin and a CTA
// barrier, bypass L1 (cp.async.cg / peer addresses) and find the data already in the owning GPU's L2.
// One process per GPU; the waiting kernel only ever waits for kernels running on OTHER GPUs.
// Epochs are `*base + offset` when `base` is set (launches replayed from a CUDA graph: the offsets are baked into the
// graph, the base is bumped before every replay), else plain values.
struct XSync {
  const unsigned long long* wait_flags;   // local flag words (wait_n of them: source ranks x chunks); nullptr: no wait
  const unsigned long long* wait_base;
  unsigned long long wait_epoch;
  int wait_n, wait_per;                   // thread t < wait_n polls word (t / wait_per) * MAXP + t % wait_per
  int sig_n, sig_rank, sig_sys;           // sig_n = 0: no signal
  const unsigned long long* sig_base;
  unsigned long long sig_epoch;
  unsigned long long* sig_flags[MAXP];    // the peers' flag words of this (buffer, chunk): word [sig_rank] is this rank's
  unsigned int* counter;                  // local: CTAs of this launch that have finished
  unsigned int* err;                      // host-mapped word: set to 1 when a wait gave up (a peer died): results are invalid
  int trace_id;                           // development builds (-DSMO_XS_TRACE): slot of this launch in the time-stamp trace, 0 = none
};
#if defined(SMO_XS_TRACE) && !defined(SMO_EMUL)
// development only: per-launch time stamps (globaltimer, ns) of the hand-shake stages, read back with smo_debug_xs_trace()
constexpr int XS_TRACE_SLOTS = 8192;
__device__ unsigned long long g_xs_trace[XS_TRACE_SLOTS * 8];
__device__ __forceinline__ unsigned long long xs_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define XS_TRACE(idx, val) do { if (p.xs.trace_id) g_xs_trace[(p.xs.trace_id % XS_TRACE_SLOTS) * 8 + (idx)] = (val); } while (0)
#else
#define XS_TRACE(idx, val) do { } while (0)
#endif
// Flag loads / stores of the hand-shake.  The waiter's load is an acquire at system scope (everything the peer stored
// before its release-ordered flag store is visible to the loads that follow); the signaller's last CTA orders the whole
// launch's stores with a system-wide fence before it publishes.  SMO_XSYNC_RELAXED restores the round-1 fast path (plain
// volatile accesses, device-scope fence in the last CTA) for A/B timing.
SMO_HD unsigned long long xs_load_flag(const unsigned long long* f) {
#if defined(__CUDA_ARCH__) && !defined(SMO_XSYNC_RELAXED)
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
  return v;
#else
  return *(const volatile unsigned long long*)f;
#endif
}
// (the release ordering comes from ONE system-wide fence before the first flag store: a st.release per peer compiles to one
//  MEMBAR.ALL.SYS per store - 8 serial fences in the last CTA at 8 GPUs, seen in the r2 SASS)
SMO_HD void xs_store_flag(unsigned long long* f, unsigned long long v) {
#if defined(__CUDA_ARCH__) && !defined(SMO_XSYNC_RELAXED)
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(f), "l"(v) : "memory");
#else
  *(volatile unsigned long long*)f = v;
#endif
}
#ifndef SMO_XSYNC_TIMEOUT_CYCLES
#define SMO_XSYNC_TIMEOUT_CYCLES 20000000000ll   // ~10 s at 1.9 GHz: far beyond any legitimate wait inside a time loop
#endif
// bounded spin: returns false (and raises the error word) if the flag did not reach `want` in time
SMO_HD bool xs_spin(const unsigned long long* f, unsigned long long want, unsigned int* err) {
#if defined(__CUDA_ARCH__)
  if (xs_load_flag(f) >= want) return true;
  const long long t0 = clock64();
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 256; ++i)
      if (xs_load_flag(f) >= want) return true;
    if (clock64() - t0 > SMO_XSYNC_TIMEOUT_CYCLES) {
      if (err) *(volatile unsigned int*)err = 1u;
      return false;
    }
  }
#else
  (void)f; (void)want; (void)err;
  return true;
#endif
}
template <class K, class = void> struct has_xsync { static constexpr bool value = false; };
template <class K> struct has_xsync<K, decltype((void)((typename K::Params*)nullptr)->xs)> { static constexpr bool value = true; };

// what a CTA knows about itself (kernels with K::V2 == true get this instead of a bare tid / smem pair)
struct Ctx {
  int cta, ncta, tid;
  unsigned char* smem;
};

SMO_HD int imax(int a, int b) { return a > b ? a : b; }
SMO_HD int imin(int a, int b) { return a < b ? a : b; }

// ---------------------------------------------------------------------------------------------
// phase runner
// ---------------------------------------------------------------------------------------------
// A kernel class K provides
//   struct Params { int nwork; int nsteps; ... };   struct State { ... };
//   static constexpr int THREADS, NPHASES, MIN_BLOCKS;
//   template <int PH> SMO_HD static void phase(const Params&, int work, int step, int tid, unsigned char* smem, State&);
// For every work item the runner executes  for step in [0,nsteps): phase<0>, sync, phase<1>, sync, ...
// kernels that define `static constexpr bool V2 = true` use the Ctx interface (init + phase2)
template <class K, class = void> struct is_v2 { static constexpr bool value = false; };
template <class K> struct is_v2<K, decltype((void)K::V2)> { static constexpr bool value = K::V2; };

template <class K, class = void> struct has_finish { static constexpr bool value = false; };
template <class K> struct has_finish<K, decltype((void)K::HAS_FINISH)> { static constexpr bool value = K::HAS_FINISH; };

template <class K, class = void> struct has_sync_kinds { static constexpr bool value = false; };
template <class K> struct has_sync_kinds<K, decltype((void)K::sync_after(0))> { static constexpr bool value = true; };

#if !defined(SMO_EMUL)
template <class K, int PH, bool END> struct PhaseStep;
template <class K, int PH> struct PhaseStep<K, PH, false> {
  static __device__ __forceinline__ void run(const typename K::Params& p, int work, int step, int tid,
                                             unsigned char* smem, typename K::State& st) {
    if constexpr (is_v2<K>::value) {
      Ctx c; c.cta = (int)blockIdx.x; c.ncta = (int)gridDim.x; c.tid = tid; c.smem = smem;
      K::template phase2<PH>(p, work, step, c, st);
    } else {
      K::template phase<PH>(p, work, step, tid, smem, st);
    }
    // barrier after the phase: CTA-wide by default; kernels may declare `sync_after(PH)` = 0 none, 1 warp, 2 CTA
    if constexpr (has_sync_kinds<K>::value) {
      if constexpr (K::sync_after(PH) == 2) __syncthreads();
      else if constexpr (K::sync_after(PH) == 1) __syncwarp();
    } else {
      __syncthreads();
    }
    PhaseStep<K, PH + 1, (PH + 1 >= K::NPHASES)>::run(p, work, step, tid, smem, st);
  }
};
template <class K, int PH> struct PhaseStep<K, PH, true> {
  static __device__ __forceinline__ void run(const typename K::Params&, int, int, int, unsigned char*,
                                             typename K::State&) {}
};

// One CTA loops over work items blockIdx.x, blockIdx.x + gridDim.x, ... (persistent-style grid).
template <class K>
__global__ void __launch_bounds__(K::THREADS, K::MIN_BLOCKS) smo_kernel(const __grid_constant__ typename K::Params p) {
  extern __shared__ __align__(1024) unsigned char smo_smem[];     // (TMA swizzle patterns are functions of the shared-memory address)
  typename K::State st;
#if defined(SMO_XS_TRACE)
  if constexpr (has_xsync<K>::value) { if (blockIdx.x == 0 && threadIdx.x == 0) { XS_TRACE(0, xs_now()); XS_TRACE(6, (unsigned long long)gridDim.x); XS_TRACE(7, (unsigned long long)(K::THREADS * 1000 + K::NPHASES)); } }
#endif
  if constexpr (is_v2<K>::value) {
    Ctx c; c.cta = (int)blockIdx.x; c.ncta = (int)gridDim.x; c.tid = (int)threadIdx.x; c.smem = smo_smem;
    K::init(p, c, st);
    __syncthreads();
  }
  // Programmatic dependent launch (SMO_OPT_PDL): when the launch carries the programmatic-stream-serialisation attribute this CTA
  // may have started while the previous kernel of the stream was still draining.  Everything above touched only kernel parameters,
  // constant tables and shared memory; from here on the kernel reads and writes what its predecessor produced, so wait until that
  // grid has completed and its stores are visible - then let the NEXT kernel's CTAs move in as ours exit.  Both instructions are
  // no-ops for a launch without the attribute.  (Measured, r2z: triggering only after the work loop gains nothing at any size; an
  // explicit start stagger of the co-resident CTAs changes nothing either.)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if constexpr (has_xsync<K>::value) {
    if (p.xs.wait_flags != nullptr) {
      if ((int)threadIdx.x < p.xs.wait_n) {
        const unsigned long long want = p.xs.wait_epoch + (p.xs.wait_base ? *p.xs.wait_base : 0ull);
        xs_spin(p.xs.wait_flags + ((int)threadIdx.x / p.xs.wait_per) * MAXP + (int)threadIdx.x % p.xs.wait_per, want, p.xs.err);   // the peer's kernel runs on another GPU
      }
      __syncthreads();
    }
  }
#if defined(SMO_XS_TRACE)
  if constexpr (has_xsync<K>::value) { if (blockIdx.x == 0 && threadIdx.x == 0) XS_TRACE(1, xs_now()); }
#endif
  for (int work = blockIdx.x; work < p.nwork; work += gridDim.x) {
    for (int step = 0; step < p.nsteps; ++step)
      PhaseStep<K, 0, false>::run(p, work, step, (int)threadIdx.x, smo_smem, st);
  }
#if defined(SMO_XS_TRACE)
  if constexpr (has_xsync<K>::value) { if (blockIdx.x == 0 && threadIdx.x == 0) XS_TRACE(2, xs_now()); }
#endif
  if constexpr (has_finish<K>::value) {   // per-CTA epilogue after the last work item (e.g. deterministic partial sums)
    __syncthreads();
    Ctx c; c.cta = (int)blockIdx.x; c.ncta = (int)gridDim.x; c.tid = (int)threadIdx.x; c.smem = smo_smem;
    K::template finish<0>(p, c, st);
    __syncthreads();
    K::template finish<1>(p, c, st);
  }
  if constexpr (has_xsync<K>::value) {
    bulk_wait_all();          // TMA bulk stores issued by this thread (staged peer pushes) are complete (no-op without any) ...
    bulk_fence_global();      // ... and ordered before the fence / flag stores below
    if (p.xs.sig_n > 0) {
      __syncthreads();
      if (threadIdx.x == 0) {
        // the peers read (pull) or received (push) this launch's results: both need the stores ordered system-wide
#if defined(SMO_XSYNC_RELAXED)
        if (p.xs.sig_sys) __threadfence_system(); else __threadfence();
#elif defined(SMO_XSYNC_CTA_SYS)
        __threadfence_system();      // (A/B: one system-scope fence per CTA, the protocol up to session r2zd)
#else
        // Device-scope release of this CTA's stores: the CTA barrier above, this fence and the count below form a release pattern
        // towards the last CTA (same GPU), whose ONE system-scope fence before the flag stores is cumulative over everything it
        // acquired through the count (PTX memory model: causality order composes across scopes; the same reasoning by which a
        // single thread's fence.sys publishes the stores its whole CTA made before a bar.sync).  A system-scope fence in EVERY CTA
        // cost 10-15 us at the end of each pushing kernel (r2zc trace: it waits for the remote acknowledgements of the whole SM);
        // the device-scope fence shortens the forward step on 2 GPUs from 159 to 148 us (r2zd).
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
        const unsigned int old = atomicAdd(p.xs.counter, 1u);
#if defined(SMO_XS_TRACE)
        if (blockIdx.x == 0) XS_TRACE(3, xs_now());
        if (old == gridDim.x - 1) XS_TRACE(4, xs_now());
#endif
        if (old == gridDim.x - 1) {
          *p.xs.counter = 0u;   // ready for the next launch
#if defined(SMO_XSYNC_RELAXED)
          __threadfence();
#else
          __threadfence_system();   // acquire side of the CTA count + release of the whole launch's stores, once for all flag stores
#endif
          const unsigned long long val = p.xs.sig_epoch + (p.xs.sig_base ? *p.xs.sig_base : 0ull);
          for (int s = 0; s < p.xs.sig_n; ++s) xs_store_flag(p.xs.sig_flags[s] + p.xs.sig_rank, val);
#if defined(SMO_XS_TRACE)
          XS_TRACE(5, xs_now());
#endif
        }
      }
    }
  }
}
#else
static thread_local int g_emul_cta = 0, g_emul_ncta = 1;
template <class K, int PH, bool END> struct EmulStep;
template <class K, int PH> struct EmulStep<K, PH, false> {
  static void run(const typename K::Params& p, int work, int step, unsigned char* smem,
                  std::vector<typename K::State>& st) {
    for (int tid = 0; tid < K::THREADS; ++tid) {
      if constexpr (is_v2<K>::value) {
        Ctx c; c.cta = g_emul_cta; c.ncta = g_emul_ncta; c.tid = tid; c.smem = smem;
        K::template phase2<PH>(p, work, step, c, st[tid]);
      } else {
        K::template phase<PH>(p, work, step, tid, smem, st[tid]);
      }
    }
    EmulStep<K, PH + 1, (PH + 1 >= K::NPHASES)>::run(p, work, step, smem, st);
  }
};
template <class K, int PH> struct EmulStep<K, PH, true> {
  static void run(const typename K::Params&, int, int, unsigned char*, std::vector<typename K::State>&) {}
};
// Host emulation of smo_kernel<K>: CTAs run one after another, threads of a CTA phase by phase.
template <class K> void emul_kernel(int grid, size_t smem_bytes, const typename K::Params& p) {
  std::vector<unsigned char> smem(smem_bytes + 256);
  unsigned char* sm = smem.data();
  sm += (128 - ((uintptr_t)sm & 127)) & 127;
  std::vector<typename K::State> st(K::THREADS);
  for (int cta = 0; cta < grid; ++cta) {
    g_emul_cta = cta; g_emul_ncta = grid;
    if constexpr (is_v2<K>::value) {
      for (int tid = 0; tid < K::THREADS; ++tid) {
        Ctx c; c.cta = cta; c.ncta = grid; c.tid = tid; c.smem = sm;
        K::init(p, c, st[tid]);
      }
    }
    for (int work = cta; work < p.nwork; work += grid)
      for (int step = 0; step < p.nsteps; ++step) EmulStep<K, 0, false>::run(p, work, step, sm, st);
    if constexpr (has_finish<K>::value) {
      for (int tid = 0; tid < K::THREADS; ++tid) { Ctx c; c.cta = cta; c.ncta = grid; c.tid = tid; c.smem = sm; K::template finish<0>(p, c, st[tid]); }
      for (int tid = 0; tid < K::THREADS; ++tid) { Ctx c; c.cta = cta; c.ncta = grid; c.tid = tid; c.smem = sm; K::template finish<1>(p, c, st[tid]); }
    }
  }
}
#endif

}  // namespace smo
