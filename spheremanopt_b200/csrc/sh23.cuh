// Swift-Hohenberg SH23 forward and discrete-adjoint time loops, one kernel launch per solve.
//
// Replaces FWD_Solve_IVP_Lin / Compatib_Cond / ADJ_Solve_IVP_Lin (and FWD_Solve_IVP_PREP) of
// FWD_Solve_SH23.py (409-545, 552-596, 598-729, 334-407): Dedalus IVP
//     dt(u) + (1+dxx)^2 u - a u = 1.8 u^2 - u^3     with SBDF1 (FWD_Solve_SH23.py:322-325),
//     (1/dt + L_k) u^{n+1}_k = u^n_k/dt + N_k(u^n),   L_k = (1-k^2)^2 - a,
// on Npts Fourier modes (k = 0..Npts/2-1 retained) and a dealias-2 grid of M = 2*Npts points, plus its exact
// discrete adjoint.  The whole time loop runs inside the kernel: RT threads own one problem instance whose
// spectral state (Nh complex numbers) lives in shared memory; each step does a c2r and an r2c transform of
// length M as complex FFTs of length H = M/2 (two register stages + one exchange, fft_core.cuh) with the usual
// even/odd pre/post-processing, the pointwise nonlinearity in registers and the diagonal implicit solve, and
// streams the snapshot u^n_k to HBM ([batch][n_iters+1][Nh] complex) for the adjoint sweep, which replays the
// snapshots backwards.  A CTA carries NI instances, so ensembles (BASELINE config 5) shard over CTAs and GPUs
// with no communication.
//
// Forward schedule : step 0 = r2c of the input vector; step s = 1..n_iters+1 handles state n = s-1 (snapshot,
//                    J += dt*mean(u_n^2), SBDF1 update if n < n_iters); step n_iters+2 reduces J.
//                    (prep mode: n_iters+1 updates, no J/snapshots, final state written on the grid.)
// Adjoint schedule : step 0 = terminal condition q0 (Compatib_Cond) and prefetch of the first snapshot;
//                    steps 1..n_iters = adjoint SBDF1 steps; step n_iters+1 = dt*(1/dt+L) q on the grid.
#pragma once
#include "fft_core.cuh"

namespace smo {

struct Sh23Params {
  const double* X;       // [batch][M] input grid vectors (forward)
  cplx* snaps;           // [batch][n_iters+1][Nh]
  double* J;             // [batch]  (forward: dt * sum_n mean(u_n^2))
  double* grad;          // [batch][M] (adjoint output / prep output)
  int nwork, nsteps;
  int batch, n_iters, Nh;
  double dt, a, kfac;    // kfac = 2 pi / L
  double inv_dt;         // 1/dt (set by the host: the time loop itself contains no fp64 division)
  int flags;             // bit0: continuous adjoint; bit1: prep mode; bit2: initial state given as coefficients
  const cplx* cin;       // [batch][Nh] initial coefficients (bit2)
  const cplx* twH;       // exp(-2 pi i m / H)
  const cplx* twM;       // exp(-2 pi i m / M)
};

SMO_HD double sh_A(const Sh23Params& p, int k) {
  const double kk = p.kfac * (double)k;
  const double t = 1.0 - kk * kk;
  return p.inv_dt + t * t - p.a;
}

template <class F> struct Sh23Core {
  typedef typename F::Swapped FS;
  static constexpr int R1 = F::R1, R2 = F::R2, H = F::M, M = 2 * F::M, RT = F::RT;
  static constexpr int XLEN = (F::XP > FS::XP) ? ((F::XP > H) ? F::XP : H) : ((FS::XP > H) ? FS::XP : H);
  static constexpr int NHMAX = H / 2;

  // Z[k] of the half-length inverse transform from a half spectrum c (entries >= Nh are zero):
  //   Z[k] = (c[k] + conj(c[H-k])) + i e^{+2 pi i k/M} (c[k] - conj(c[H-k])) ;  Im c[0] dropped (c2r convention).
  // scaleA multiplies c[k] by dt*(1/dt + L_k) first (final un-inversion of the discrete adjoint).
  SMO_HD static void preprocess(const Sh23Params& p, const cplx* c, bool scaleA, int k, double& zr, double& zi) {
    zr = 0.0; zi = 0.0;
    const int Nh = p.Nh;
    if (k < Nh) {
      cplx v = c[k];
      if (scaleA) { const double s0 = p.dt * sh_A(p, k); v.x *= s0; v.y *= s0; }
      if (k == 0) { zr = v.x; zi = v.x; return; }
      const cplx w = ldg_c(p.twM + k);
      const double c_ = w.x, s_ = -w.y;          // e^{+2 pi i k/M} = c_ + i s_
      zr = v.x * (1.0 - s_) - v.y * c_;
      zi = v.x * c_ + v.y * (1.0 - s_);
    } else if (k > H - Nh) {
      cplx v = c[H - k];
      if (scaleA) { const double s0 = p.dt * sh_A(p, H - k); v.x *= s0; v.y *= s0; }
      v.y = -v.y;                                  // conj
      const cplx w = ldg_c(p.twM + k);
      const double c_ = w.x, s_ = -w.y;
      zr = v.x * (1.0 + s_) + v.y * c_;
      zi = -v.x * c_ + v.y * (1.0 + s_);
    }
  }
  SMO_HD static void inv_stage1(const Sh23Params& p, const cplx* c, bool scaleA, int jj, cplx* XA, double* re,
                                double* im) {
    if (jj < R2) {
#pragma unroll
      for (int i = 0; i < R1; ++i) preprocess(p, c, scaleA, jj + R2 * i, re[i], im[i]);
      stage1<F, +1>(re, im, jj, p.twH);
#pragma unroll
      for (int k1 = 0; k1 < R1; ++k1) XA[jj * F::SK + k1] = make_double2(re[k1], im[k1]);
    }
  }
  // thread k1 < R1 ends with (u[2n], u[2n+1]) in (re[k2], im[k2]) for n = k1 + R1*k2
  SMO_HD static void inv_stage2(const cplx* XA, int jj, double* re, double* im) {
    if (jj < R1) {
#pragma unroll
      for (int j = 0; j < R2; ++j) { const cplx v = XA[j * F::SK + jj]; re[j] = v.x; im[j] = v.y; }
      stage2<F, +1>(re, im);
    }
  }
  // forward stage 1 (swapped factorisation) on the values a stage-2 thread holds, -> exchange XB
  SMO_HD static void fwd_stage1(const Sh23Params& p, int jj, cplx* XB, double* re, double* im) {
    if (jj < R1) {
      stage1<FS, -1>(re, im, jj, p.twH);
#pragma unroll
      for (int k1 = 0; k1 < R2; ++k1) XB[jj * FS::SK + k1] = make_double2(re[k1], im[k1]);
    }
  }
  // forward stage 2, spectrum written in natural order to XZ
  SMO_HD static void fwd_stage2(const cplx* XB, cplx* XZ, int jj, double* re, double* im) {
    if (jj < R2) {
#pragma unroll
      for (int j = 0; j < R1; ++j) { const cplx v = XB[j * FS::SK + jj]; re[j] = v.x; im[j] = v.y; }
      stage2<FS, -1>(re, im);
#pragma unroll
      for (int k2 = 0; k2 < R1; ++k2) XZ[jj + R2 * k2] = make_double2(re[k2], im[k2]);
    }
  }
  // coefficient k (< Nh) of the length-M r2c transform, divided by M, from the half-length spectrum XZ
  SMO_HD static cplx postprocess(const Sh23Params& p, const cplx* XZ, int k) {
    const cplx zk = XZ[k], zm = XZ[(H - k) % H];
    const double evr = 0.5 * (zk.x + zm.x), evi = 0.5 * (zk.y - zm.y);
    const double odr = 0.5 * (zk.y + zm.y), odi = 0.5 * (zm.x - zk.x);
    const cplx w = ldg_c(p.twM + k);
    const double inv = 1.0 / (double)M;
    return make_double2((evr + odr * w.x - odi * w.y) * inv, (evi + odr * w.y + odi * w.x) * inv);
  }
};

// ------------------------------------------------------------------------------------------------ forward
template <class F, int NI_> struct Sh23Fwd {
  typedef Sh23Params Params;
  typedef Sh23Core<F> Cr;
  static constexpr int NI = NI_, RT = F::RT, R1 = F::R1, R2 = F::R2, M = Cr::M;
  static constexpr int THREADS = NI_ * RT;
  static constexpr int NPHASES = 4;
  static constexpr int MIN_BLOCKS = 1;
  static constexpr int PER_INST = Cr::NHMAX + 2 * Cr::XLEN + Cr::NHMAX / 2;   // C | XA | XB | 1/A_k (doubles)
  static constexpr size_t SMEM = (size_t)NI_ * PER_INST * sizeof(cplx);
  // an instance is private to its RT threads: with RT dividing 32 every barrier of the time loop is a __syncwarp
  SMO_HD static constexpr int sync_after(int) { return (32 % RT == 0) ? 1 : 2; }
  struct State {
    double re[RT], im[RT];
    double jacc;
  };
  template <int PH>
  SMO_HD static void phase(const Params& p, int work, int step, int tid, unsigned char* smem, State& st) {
    const int jj = tid % RT, li = tid / RT;
    const int inst = work * NI + li;
    const bool live = inst < p.batch;
    cplx* C = reinterpret_cast<cplx*>(smem) + (size_t)li * PER_INST;
    cplx* XA = C + Cr::NHMAX;
    cplx* XB = XA + Cr::XLEN;
    double* IA = reinterpret_cast<double*>(XB + Cr::XLEN);
    const int Nh = p.Nh, NIT = p.n_iters;
    const bool prep = (p.flags & 2) != 0;
    const int n = step - 1;
    const bool fin = (step == NIT + 2);
    const bool do_inv = (step >= 1) && !fin;
    const bool do_upd = do_inv && (n < NIT || prep);
    if (PH == 0) {
      if (step == 0) {
        st.jacc = 0.0;
        for (int k = jj; k < Nh; k += RT) IA[k] = 1.0 / sh_A(p, k);
      }
      if (do_inv) {
        if (live && !prep) {
          cplx* dst = p.snaps + ((long long)inst * (NIT + 1) + n) * Nh;
          for (int k = jj; k < Nh; k += RT) dst[k] = C[k];
        }
        Cr::inv_stage1(p, C, false, jj, XA, st.re, st.im);
      }
      if (fin && prep) Cr::inv_stage1(p, C, false, jj, XA, st.re, st.im);
    } else if (PH == 1) {
      if (step == 0) {
        if (jj < R1) {
#pragma unroll
          for (int i = 0; i < R2; ++i) {
            const int nn = jj + R1 * i;
            double a0 = 0.0, a1 = 0.0;
            if (live && !(p.flags & 4)) { a0 = p.X[(long long)inst * M + 2 * nn]; a1 = p.X[(long long)inst * M + 2 * nn + 1]; }
            st.re[i] = a0; st.im[i] = a1;
          }
        }
        Cr::fwd_stage1(p, jj, XB, st.re, st.im);
      } else if (do_inv) {
        Cr::inv_stage2(XA, jj, st.re, st.im);
        if (jj < R1) {
          double acc = 0.0;
#pragma unroll
          for (int i = 0; i < R2; ++i) {
            const double u0 = st.re[i], u1 = st.im[i];
            acc += u0 * u0 + u1 * u1;
            st.re[i] = 1.8 * u0 * u0 - u0 * u0 * u0;
            st.im[i] = 1.8 * u1 * u1 - u1 * u1 * u1;
          }
          if (n <= NIT) st.jacc += acc;
        }
        if (do_upd) Cr::fwd_stage1(p, jj, XB, st.re, st.im);
      } else if (fin && prep) {
        Cr::inv_stage2(XA, jj, st.re, st.im);
        if (jj < R1 && live) {
#pragma unroll
          for (int k2 = 0; k2 < R2; ++k2) {
            const int nn = jj + R1 * k2;
            p.grad[(long long)inst * M + 2 * nn] = st.re[k2];
            p.grad[(long long)inst * M + 2 * nn + 1] = st.im[k2];
          }
        }
      } else if (fin) {
        XB[jj] = make_double2(st.jacc, 0.0);
      }
    } else if (PH == 2) {
      if (step == 0 || do_upd) Cr::fwd_stage2(XB, XA, jj, st.re, st.im);
      if (fin && !prep && jj == 0 && live) {
        double s = 0.0;
        for (int t = 0; t < RT; ++t) s += XB[t].x;
        p.J[inst] = p.dt * s / (double)M;
      }
    } else {
      if (step == 0) {
        if (p.flags & 4) {
          for (int k = jj; k < Nh; k += RT) C[k] = live ? p.cin[(long long)inst * Nh + k] : make_double2(0.0, 0.0);
        } else {
          for (int k = jj; k < Nh; k += RT) C[k] = Cr::postprocess(p, XA, k);
        }
      } else if (do_upd) {
        for (int k = jj; k < Nh; k += RT) {
          const cplx nh = Cr::postprocess(p, XA, k);
          const cplx c = C[k];
          const double iA = IA[k];
          C[k] = make_double2((c.x * p.inv_dt + nh.x) * iA, (c.y * p.inv_dt + nh.y) * iA);
        }
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------ adjoint
template <class F, int NI_> struct Sh23Adj {
  typedef Sh23Params Params;
  typedef Sh23Core<F> Cr;
  static constexpr int NI = NI_, RT = F::RT, R1 = F::R1, R2 = F::R2, M = Cr::M;
  static constexpr int THREADS = NI_ * RT;
  static constexpr int NPHASES = 5;
  static constexpr int MIN_BLOCKS = 1;
  // C (= q) | S (snapshot) | XA | XB | 1/A_k ; the forward exchange re-uses XB one (warp) barrier after the inverse read it
  static constexpr int PER_INST = 2 * Cr::NHMAX + 2 * Cr::XLEN + Cr::NHMAX / 2;
  static constexpr size_t SMEM = (size_t)NI_ * PER_INST * sizeof(cplx);
  SMO_HD static constexpr int sync_after(int) { return (32 % RT == 0) ? 1 : 2; }
  struct State {
    double re[RT], im[RT];
    double ur[RT], ui[RT];
  };
  template <int PH>
  SMO_HD static void phase(const Params& p, int work, int step, int tid, unsigned char* smem, State& st) {
    const int jj = tid % RT, li = tid / RT;
    const int inst = work * NI + li;
    const bool live = inst < p.batch;
    cplx* C = reinterpret_cast<cplx*>(smem) + (size_t)li * PER_INST;
    cplx* S = C + Cr::NHMAX;
    cplx* XA = S + Cr::NHMAX;
    cplx* XB = XA + Cr::XLEN;
    cplx* XC = XB;
    double* IA = reinterpret_cast<double*>(XB + Cr::XLEN);
    const int Nh = p.Nh, NIT = p.n_iters;
    const bool cont = (p.flags & 1) != 0;
    const bool fin = (step == NIT + 1);
    const bool do_step = (step >= 1) && !fin;
    const cplx* snaps = p.snaps + (long long)inst * (NIT + 1) * Nh;
    if (PH == 0) {
      if (do_step) {
        Cr::inv_stage1(p, S, false, jj, XA, st.re, st.im);      // u_f
        Cr::inv_stage1(p, C, false, jj, XB, st.re, st.im);      // q
      }
      if (fin) Cr::inv_stage1(p, C, !cont, jj, XA, st.re, st.im);
    } else if (PH == 1) {
      if (do_step) {
        Cr::inv_stage2(XA, jj, st.ur, st.ui);
        Cr::inv_stage2(XB, jj, st.re, st.im);
        if (jj < R1) {
#pragma unroll
          for (int i = 0; i < R2; ++i) {
            const double u0 = st.ur[i], u1 = st.ui[i];
            st.re[i] = (3.6 * u0 - 3.0 * (u0 * u0)) * st.re[i] - 2.0 * u0;    // FWD_Solve_SH23.py:640
            st.im[i] = (3.6 * u1 - 3.0 * (u1 * u1)) * st.im[i] - 2.0 * u1;
          }
        }
      }
      if (fin) {
        Cr::inv_stage2(XA, jj, st.re, st.im);
        if (jj < R1 && live) {
#pragma unroll
          for (int k2 = 0; k2 < R2; ++k2) {
            const int nn = jj + R1 * k2;
            p.grad[(long long)inst * M + 2 * nn] = st.re[k2];
            p.grad[(long long)inst * M + 2 * nn + 1] = st.im[k2];
          }
        }
      }
    } else if (PH == 2) {
      if (do_step) Cr::fwd_stage1(p, jj, XC, st.re, st.im);
    } else if (PH == 3) {
      if (do_step) Cr::fwd_stage2(XC, XA, jj, st.re, st.im);
    } else {
      if (step == 0) {
        // terminal condition: discrete q0 = -2 u^N/(1/dt + L) (FWD_Solve_SH23.py:584), continuous q0 = 0
        for (int k = jj; k < Nh; k += RT) {
          cplx q = make_double2(0.0, 0.0);
          const double iA = 1.0 / sh_A(p, k);
          IA[k] = iA;
          if (!cont && live) { const cplx f = snaps[(long long)NIT * Nh + k]; q = make_double2(-2.0 * f.x * iA, -2.0 * f.y * iA); }
          C[k] = q;
        }
      } else if (do_step) {
        for (int k = jj; k < Nh; k += RT) {
          const cplx rh = Cr::postprocess(p, XA, k);
          const cplx c = C[k];
          const double iA = IA[k];
          C[k] = make_double2((c.x * p.inv_dt + rh.x) * iA, (c.y * p.inv_dt + rh.y) * iA);
        }
      }
      // prefetch the forward snapshot of the NEXT adjoint step m = step: index -2-m (discrete), -1-m (continuous)
      if (step < NIT) {
        const int sidx = cont ? (NIT - step) : (NIT - 1 - step);
        for (int k = jj; k < Nh; k += RT) S[k] = live ? snaps[(long long)sidx * Nh + k] : make_double2(0.0, 0.0);
      }
    }
  }
};

}  // namespace smo
