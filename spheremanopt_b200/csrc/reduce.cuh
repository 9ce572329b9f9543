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
  int vec2;                  // 1: x, y, out are 16-byte aligned -> 128-bit accesses
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

// one element of operation OP: reads xv / yv / the old output ov as the operation needs, returns the new output in ov
template <int OP> SMO_HD void vec_elem(const VecParams& p, double s, double xv, double yv, double& ov, double& a0, double& a1) {
  if (OP == V_DOT) a0 += xv * yv;
  if (OP == V_DOT2) { a0 += xv * yv; a1 += xv * xv; }
  if (OP == V_AXPBY) ov = p.a * xv + p.b * yv;
  if (OP == V_SCALE) ov = p.a * xv;
  if (OP == V_PROJ) ov = yv - s * xv;
  if (OP == V_AXPY_NRM) { const double f = xv + p.a * yv; ov = f; a0 += f * f; }
  if (OP == V_RESCALE) ov = ov * s;
}
template <int OP> struct VecTraits {
  static constexpr bool RX = (OP != V_RESCALE);
  static constexpr bool RY = (OP == V_DOT || OP == V_DOT2 || OP == V_AXPBY || OP == V_PROJ || OP == V_AXPY_NRM);
  static constexpr bool RO = (OP == V_RESCALE);
  static constexpr bool WO = (OP == V_AXPBY || OP == V_SCALE || OP == V_PROJ || OP == V_AXPY_NRM || OP == V_RESCALE);
};
SMO_HD double2 ld_pair(const double* p, long long i2) {
#if defined(__CUDA_ARCH__)
  return __ldg(reinterpret_cast<const double2*>(p) + i2);
#else
  return make_double2(p[2 * i2], p[2 * i2 + 1]);
#endif
}

// Streaming kernels: every thread moves 16 bytes per access (ld/st.global.v2.f64), 4 independent accesses per operand in
// flight (p.vec2 = 1: all operands 16-byte aligned; else the scalar path).  HBM-bound: 2-3 vector passes per operation.
template <int OP> struct VecKernel {
  typedef VecParams Params;
  typedef VecTraits<OP> Tr;
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
      if (p.vec2) {
        const long long n2 = p.n >> 1, base2 = base >> 1;
        constexpr int NE = VCHUNK / 2 / VTHREADS;
#pragma unroll 4
        for (int e = 0; e < NE; ++e) {
          const long long i2 = base2 + (long long)e * VTHREADS + tid;
          if (i2 < n2) {
            double2 xv = make_double2(0.0, 0.0), yv = xv, ov = xv;
            if (Tr::RX) xv = ld_pair(p.x, i2);
            if (Tr::RY) yv = ld_pair(p.y, i2);
            if (Tr::RO) ov = reinterpret_cast<const double2*>(p.out)[i2];
            vec_elem<OP>(p, s, xv.x, yv.x, ov.x, a0, a1);
            vec_elem<OP>(p, s, xv.y, yv.y, ov.y, a0, a1);
            if (Tr::WO) reinterpret_cast<double2*>(p.out)[i2] = ov;
          }
        }
        const long long last = p.n - 1;     // odd length: the unpaired last element
        if ((p.n & 1) && tid == 0 && last >= base && last < base + VCHUNK) {
          double ov = Tr::RO ? p.out[last] : 0.0;
          vec_elem<OP>(p, s, Tr::RX ? p.x[last] : 0.0, Tr::RY ? p.y[last] : 0.0, ov, a0, a1);
          if (Tr::WO) p.out[last] = ov;
        }
      } else {
#pragma unroll 4
        for (int e = 0; e < VCHUNK / VTHREADS; ++e) {
          const long long i = base + (long long)e * VTHREADS + tid;
          if (i < p.n) {
            double ov = Tr::RO ? p.out[i] : 0.0;
            vec_elem<OP>(p, s, Tr::RX ? p.x[i] : 0.0, Tr::RY ? p.y[i] : 0.0, ov, a0, a1);
            if (Tr::WO) p.out[i] = ov;
          }
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

// ---------------------------------------------------------------------------------------------------------------
// Row-wise inner products of two [rows][len] matrices: out[r] = a * sum_j x[r][j]*y[r][j].  One CTA per row, fixed summation
// order.  The batched Inner_Product of the ensemble driver (many small SH23 vectors per launch instead of one launch each).
struct VecRowDot {
  typedef VecParams Params;       // n = row length, nwork = rows, out[rows]
  static constexpr int THREADS = VTHREADS;
  static constexpr int NPHASES = 2;
  static constexpr int MIN_BLOCKS = 1;
  static constexpr size_t SMEM = VTHREADS * sizeof(double);
  struct State {};
  template <int PH>
  SMO_HD static void phase(const Params& p, int work, int, int tid, unsigned char* smem, State&) {
    double* S = reinterpret_cast<double*>(smem);
    if (PH == 0) {
      const double* x = p.x + (long long)work * p.n;
      const double* y = p.y + (long long)work * p.n;
      double a = 0.0;
      for (long long j = tid; j < p.n; j += VTHREADS) a += x[j] * y[j];
      S[tid] = a;
    } else if (tid == 0) {
      double s = 0.0;
      for (int t = 0; t < VTHREADS; ++t) s += S[t];
      p.out[work] = p.a * s;
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// 64-bit position-sensitive checksum of a vector's bit patterns: sum_i bits(x_i) * (2 i + 1) mod 2^64.  Used by the host
// layer to notice that the snapshot store Grad_f is about to replay was written for a different X (the reference couples
// f and Grad_f through that store); one streaming pass, deterministic, any single changed entry changes it.
struct VecHash {
  typedef VecParams Params;
  static constexpr int THREADS = VTHREADS;
  static constexpr int NPHASES = 2;
  static constexpr int MIN_BLOCKS = 1;
  static constexpr size_t SMEM = VTHREADS * sizeof(unsigned long long);
  struct State {};
  template <int PH>
  SMO_HD static void phase(const Params& p, int work, int, int tid, unsigned char* smem, State&) {
    unsigned long long* S = reinterpret_cast<unsigned long long*>(smem);
    if (PH == 0) {
      const long long base = (long long)work * VCHUNK;
      unsigned long long a = 0ull;
#pragma unroll 4
      for (int e = 0; e < VCHUNK / VTHREADS; ++e) {
        const long long i = base + (long long)e * VTHREADS + tid;
        if (i < p.n) {
          unsigned long long b;
          const double v = p.x[i];
          memcpy(&b, &v, sizeof b);
          a += b * (2ull * (unsigned long long)i + 1ull);
        }
      }
      S[tid] = a;
    } else if (tid == 0) {
      unsigned long long s = 0ull;
      for (int t = 0; t < VTHREADS; ++t) s += S[t];
      reinterpret_cast<unsigned long long*>(p.partials)[work] = s;
    }
  }
};
struct HashSum {
  typedef SumParams Params;
  static constexpr int THREADS = VTHREADS;
  static constexpr int NPHASES = 2;
  static constexpr int MIN_BLOCKS = 1;
  static constexpr size_t SMEM = VTHREADS * sizeof(unsigned long long);
  struct State {};
  template <int PH> SMO_HD static void phase(const Params& p, int, int, int tid, unsigned char* smem, State&) {
    unsigned long long* S = reinterpret_cast<unsigned long long*>(smem);
    const unsigned long long* P = reinterpret_cast<const unsigned long long*>(p.partials);
    if (PH == 0) {
      unsigned long long s = 0ull;
      for (int w = tid; w < p.npart; w += VTHREADS) s += P[w];
      S[tid] = s;
    } else if (tid == 0) {
      unsigned long long s = 0ull;
      for (int t = 0; t < VTHREADS; ++t) s += S[t];
      reinterpret_cast<unsigned long long*>(p.out)[0] = s;
    }
  }
};

}  // namespace smo
