// Inner product and sphere-geometry vector kernels (fp64, deterministic).
//
// Replaces Inner_Prod / Inner_Prod_3 (FWD_Solve_SH23.py:158-172, FWD_Solve_KDyn.py:173-181: the mean over the
// dealiased grid of x_j*y_j) and the numpy axpy/scale algebra of tangent_vector, transport_vector and
// Update_vector (Sphere_Grad_Descent.py:625-690).
//
// Reductions are two-stage: every work item (a fixed CHUNK of elements, independent of the grid size) writes
// one partial sum per accumulated quantity; a single-CTA kernel then adds the partials in a fixed order, so
// results are bit-reproducible from run to run and independent of the launch geometry.
#pragma once
#include "smo_common.cuh"

namespace smo {

struct VecParams {
  const double* x;
  const double* y;
  const double* z;
  double* out;
  double* partials;          // [nq][nwork]
  const double* scalars;     // device scalars produced by FinalSum
  int nwork, nsteps;
  long long n;
  double a, b, c;
  int op;
};

constexpr int VTHREADS = 256;
constexpr int VCHUNK = 256 * 32;   // elements per work item

SMO_DEV double warp_sum(double v) {
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
#endif
  return v;
}

// block-wide sum of NQ per-thread values -> partials[q*nwork + work]
template <int NQ> SMO_HD void block_reduce_store(const VecParams& p, int work, int tid, unsigned char* smem, int ph,
                                                 double* acc) {
  double* S = reinterpret_cast<double*>(smem);
#if defined(__CUDA_ARCH__)
  if (ph == 0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const double v = warp_sum(acc[q]);
      if ((tid & 31) == 0) S[q * (VTHREADS / 32) + (tid >> 5)] = v;
    }
  } else if (tid < NQ) {
    double s = 0.0;
    for (int w = 0; w < VTHREADS / 32; ++w) s += S[tid * (VTHREADS / 32) + w];
    p.partials[(long long)tid * p.nwork + work] = s;
  }
#else
  if (ph == 0) {
    for (int q = 0; q < NQ; ++q) S[q * VTHREADS + tid] = acc[q];
  } else if (tid < NQ) {
    double s = 0.0;
    for (int t = 0; t < VTHREADS; ++t) s += S[tid * VTHREADS + t];
    p.partials[(long long)tid * p.nwork + work] = s;
  }
#endif
}

enum {
  V_DOT = 0,      // partial of sum x*y
  V_DOT2 = 1,     // partials of sum x*y and sum x*x                       (tangent / transport, pass 1)
  V_AXPBY = 2,    // out = a*x + b*y
  V_SCALE = 3,    // out = a*x
  V_PROJ = 4,     // out = y - (s0/s1)*x           with device scalars       (tangent / transport, pass 2)
  V_AXPY_NRM = 5, // out = x + a*y, partial of sum out*out                  (retraction, pass 1)
  V_RESCALE = 6   // out = out*sqrt(b/(c*s0))      with device scalar        (retraction, pass 2)
};

template <int OP> struct VecKernel {
  typedef VecParams Params;
  static constexpr int THREADS = VTHREADS;
  static constexpr int NPHASES = 2;
  static constexpr int MIN_BLOCKS = 1;
  static constexpr size_t SMEM = 2 * VTHREADS * sizeof(double);
  struct State {
    double acc[2];
  };
  template <int PH>
  SMO_HD static void phase(const Params& p, int work, int, int tid, unsigned char* smem, State& st) {
    if (PH == 0) {
      const long long base = (long long)work * VCHUNK;
      double a0 = 0.0, a1 = 0.0;
      double s = 0.0;
      if (OP == V_PROJ) s = p.scalars[0] / p.scalars[1];
      if (OP == V_RESCALE) s = sqrt(p.b / (p.c * p.scalars[0]));
#pragma unroll 4
      for (int e = 0; e < VCHUNK / VTHREADS; ++e) {
        const long long i = base + (long long)e * VTHREADS + tid;
        if (i < p.n) {
          if (OP == V_DOT) a0 += p.x[i] * p.y[i];
          if (OP == V_DOT2) { const double xv = p.x[i]; a0 += xv * p.y[i]; a1 += xv * xv; }
          if (OP == V_AXPBY) p.out[i] = p.a * p.x[i] + p.b * p.y[i];
          if (OP == V_SCALE) p.out[i] = p.a * p.x[i];
          if (OP == V_PROJ) p.out[i] = p.y[i] - s * p.x[i];
          if (OP == V_AXPY_NRM) { const double f = p.x[i] + p.a * p.y[i]; p.out[i] = f; a0 += f * f; }
          if (OP == V_RESCALE) p.out[i] = p.out[i] * s;
        }
      }
      st.acc[0] = a0; st.acc[1] = a1;
      if (OP == V_DOT || OP == V_AXPY_NRM) block_reduce_store<1>(p, work, tid, smem, 0, st.acc);
      if (OP == V_DOT2) block_reduce_store<2>(p, work, tid, smem, 0, st.acc);
    } else {
      if (OP == V_DOT || OP == V_AXPY_NRM) block_reduce_store<1>(p, work, tid, smem, 1, st.acc);
      if (OP == V_DOT2) block_reduce_store<2>(p, work, tid, smem, 1, st.acc);
    }
  }
};

// out[q] = a * sum_w partials[q*npart + w], q < nq (fixed summation order).  One CTA.
struct SumParams {
  const double* partials;
  double* out;
  int nwork, nsteps;
  int npart, nq;
  double a;
};
struct FinalSum {
  typedef SumParams Params;
  static constexpr int THREADS = VTHREADS;
  static constexpr int NPHASES = 2;
  static constexpr int MIN_BLOCKS = 1;
  static constexpr size_t SMEM = VTHREADS * sizeof(double);
  struct State {};
  template <int PH> SMO_HD static void phase(const Params& p, int work, int, int tid, unsigned char* smem, State&) {
    double* S = reinterpret_cast<double*>(smem);
    const int q = work;
    if (PH == 0) {
      double s = 0.0;
      for (int w = tid; w < p.npart; w += VTHREADS) s += p.partials[(long long)q * p.npart + w];
      S[tid] = s;
    } else if (tid == 0) {
      double s = 0.0;
      for (int t = 0; t < VTHREADS; ++t) s += S[t];
      p.out[q] = p.a * s;
    }
  }
};

}  // namespace smo
