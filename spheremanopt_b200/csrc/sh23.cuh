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
// even/odd pre/post-processing, the pointwise nonlinearity in registers and the diagonal implicit solve.
//
// Snapshot store (opaque; per instance (n_iters+1)*M doubles + Nh complex): the forward solve streams the GRID values u^n
// - which it holds in registers right after its c2r transform - to HBM, plus the coefficients of the final state (terminal
// condition of the adjoint).  The adjoint sweep reads u_f = u^{N-1-m} straight back into the pointwise product, so an adjoint
// step is ONE c2r (of q) + one r2c instead of two c2r + one r2c: a third of its transforms disappears for 2x the snapshot
// bytes (16.8 GB per 4096-instance ensemble pair = 2.6 ms at the HBM roofline; the kernels are fp64 / latency bound).
// Resources are budgeted for 8 CTAs of NI = 4 instances per SM (<= 128 registers, 26.6 KB of shared memory: one exchange
// buffer per instance, one 1/A_k table per CTA): 32 resident instances per SM, so the 4096-instance ensemble of BASELINE
// config 5 (27.7 per SM) is ONE wave with 4 warps per scheduler instead of two waves with 2.
// A CTA carries NI instances, so ensembles shard over CTAs and GPUs with no communication.
//
// Forward schedule : step 0 = r2c of the input vector; step s = 1..n_iters+1 handles state n = s-1 (snapshot,
//                    J += dt*mean(u_n^2), SBDF1 update if n < n_iters); step n_iters+2 reduces J.
//                    (prep mode: n_iters+1 updates, no J/snapshots, final state written on the grid.)
// Adjoint schedule : step 0 = terminal condition q0 (Compatib_Cond); steps 1..n_iters = adjoint SBDF1 steps;
//                    step n_iters+1 = dt*(1/dt+L) q on the grid.
#pragma once
#include "fft_core.cuh"

namespace smo {

struct Sh23Params {
  const double* X;       // [batch] input grid vectors (forward), instance stride xstride
  double* snaps;         // snapshot store: instance b at snaps + b*sstride doubles: [n_iters+1][M] grid values, then Nh complex
  double* J;             // [batch]  (forward: dt * sum_n mean(u_n^2))
  double* grad;          // [batch][M] (adjoint output / prep output)
  int nwork, nsteps;
  int batch, n_iters, Nh;
  long long xstride, sstride;
  double dt, a, kfac;    // kfac = 2 pi / L
  double inv_dt;         // 1/dt (set by the host: the time loop itself contains no fp64 division)
  int flags;             // bit0: continuous adjoint; bit1: prep mode; bit2: initial state given as coefficients;
                         // bit3: only transform the input: its coefficients -> cout
  const cplx* cin;       // [batch][Nh] initial coefficients (bit2)
  cplx* cout;            // [batch][Nh] (bit3)
  const cplx* twH;       // exp(-2 pi i m / H)
  const cplx* twM;       // exp(-2 pi i m / M)
};

SMO_HD double sh_A(const Sh23Params& p, int k) {
  const double kk = p.kfac * (double)k;
  const double t = 1.0 - kk * kk;
  return p.inv_dt + t * t - p.a;
}

#ifndef SMO_SH_MB
#define SMO_SH_MB 8     // resident CTAs per SM the register allocation of the SH23 kernels is bounded for
#endif

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
  // forward stage 2: exchange buffer -> registers (thread jj < R2 ends with Z[jj + R2*k2] in (re[k2], im[k2]))
  SMO_HD static void fwd_stage2_regs(const cplx* XB, int jj, double* re, double* im) {
    if (jj < R2) {
#pragma unroll
      for (int j = 0; j < R1; ++j) { const cplx v = XB[j * FS::SK + jj]; re[j] = v.x; im[j] = v.y; }
      stage2<FS, -1>(re, im);
    }
  }
  // ... and the spectrum in natural order back into the (now fully consumed) buffer
  SMO_HD static void spectrum_store(cplx* XZ, int jj, const double* re, const double* im) {
    if (jj < R2) {
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
  // shared memory of a CTA of NI instances: per instance C (coefficient state) | X (exchange / spectrum); then 1/A_k, shared
  static constexpr int PER_INST = NHMAX + XLEN;
  template <int NI> static constexpr size_t smem_bytes() { return (size_t)(NI * PER_INST + NHMAX / 2) * sizeof(cplx); }
  SMO_HD static cplx* cbuf(unsigned char* smem, int li) { return reinterpret_cast<cplx*>(smem) + (size_t)li * PER_INST; }
  template <int NI> SMO_HD static double* iabuf(unsigned char* smem) { return reinterpret_cast<double*>(reinterpret_cast<cplx*>(smem) + (size_t)NI * PER_INST); }
};

// ------------------------------------------------------------------------------------------------ forward
// MB_ = resident CTAs per SM the register allocation is bounded for (8: ensembles; 2: few instances, latency bound - the compiler
// may keep more of a step in flight)
template <class F, int NI_, int MB_ = SMO_SH_MB> struct Sh23Fwd {
  typedef Sh23Params Params;
  typedef Sh23Core<F> Cr;
  static constexpr int NI = NI_, RT = F::RT, R1 = F::R1, R2 = F::R2, M = Cr::M;
  static constexpr int THREADS = NI_ * RT;
  static constexpr int NPHASES = 6;
  static constexpr int MIN_BLOCKS = MB_;
  static constexpr size_t SMEM = Cr::template smem_bytes<NI_>();
  // an instance is private to its RT threads: with RT dividing 32 every barrier of the time loop is a __syncwarp; the shared
  // 1/A_k table is written in step 0 and first read in step 1, with the CTA barrier after phase 5 of step 0 in between
  SMO_HD static constexpr int sync_after(int ph) { return (32 % RT == 0 && ph != 5) ? 1 : 2; }
  struct State {
    double re[RT], im[RT];
    double jacc;
  };
  template <int PH>
  SMO_HD static void phase(const Params& p, int work, int step, int tid, unsigned char* smem, State& st) {
    const int jj = tid % RT, li = tid / RT;
    const int inst = work * NI + li;
    const bool live = inst < p.batch;
    cplx* C = Cr::cbuf(smem, li);
    cplx* X = C + Cr::NHMAX;
    double* IA = Cr::template iabuf<NI>(smem);
    const int Nh = p.Nh, NIT = p.n_iters;
    const bool prep = (p.flags & 2) != 0;
    const bool coef_only = (p.flags & 8) != 0;
    const int n = step - 1;
    const bool fin = (step == NIT + 2);
    const bool do_inv = (step >= 1) && !fin;
    const bool do_upd = do_inv && (n < NIT || prep);
    double* sg = p.snaps + (long long)inst * p.sstride;     // this instance's store
    if (PH == 0) {
      if (step == 0) {
        st.jacc = 0.0;
        for (int k = tid; k < Nh; k += THREADS) IA[k] = 1.0 / sh_A(p, k);
      }
      if (do_inv || (fin && prep)) Cr::inv_stage1(p, C, false, jj, X, st.re, st.im);
    } else if (PH == 1) {
      if (step == 0) {
        if (jj < R1) {
#pragma unroll
          for (int i = 0; i < R2; ++i) {
            const int nn = jj + R1 * i;
            cplx v = make_double2(0.0, 0.0);
            if (live && !(p.flags & 4)) v = *reinterpret_cast<const cplx*>(p.X + (long long)inst * p.xstride + 2 * nn);
            st.re[i] = v.x; st.im[i] = v.y;
          }
        }
      } else if (do_inv) {
        Cr::inv_stage2(X, jj, st.re, st.im);
        if (jj < R1) {
          if (live && !prep) {   // snapshot n = the grid values just produced: 256 contiguous bytes per k2 and instance
            cplx* dst = reinterpret_cast<cplx*>(sg + (long long)n * M) + jj;
#pragma unroll
            for (int k2 = 0; k2 < R2; ++k2) dst[R1 * k2] = make_double2(st.re[k2], st.im[k2]);
          }
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
      } else if (fin && prep) {
        Cr::inv_stage2(X, jj, st.re, st.im);
        if (jj < R1 && live) {
          cplx* dst = reinterpret_cast<cplx*>(p.grad + (long long)inst * M) + jj;
#pragma unroll
          for (int k2 = 0; k2 < R2; ++k2) dst[R1 * k2] = make_double2(st.re[k2], st.im[k2]);
        }
      } else if (fin) {
        X[jj] = make_double2(st.jacc, 0.0);
      }
    } else if (PH == 2) {
      // (every thread of the instance has read the exchange buffer: it can take the forward transform's stage 1)
      if (step == 0 || do_upd) Cr::fwd_stage1(p, jj, X, st.re, st.im);
      if (fin && !prep && jj == 0 && live) {
        double s = 0.0;
        for (int t = 0; t < RT; ++t) s += X[t].x;
        p.J[inst] = p.dt * s / (double)M;
      }
    } else if (PH == 3) {
      if (step == 0 || do_upd) Cr::fwd_stage2_regs(X, jj, st.re, st.im);
    } else if (PH == 4) {
      if (step == 0 || do_upd) Cr::spectrum_store(X, jj, st.re, st.im);
    } else {
      if (step == 0) {
        if (p.flags & 4) {
          for (int k = jj; k < Nh; k += RT) C[k] = live ? p.cin[(long long)inst * Nh + k] : make_double2(0.0, 0.0);
        } else {
          for (int k = jj; k < Nh; k += RT) {
            const cplx c = Cr::postprocess(p, X, k);
            C[k] = c;
            if (coef_only && live) p.cout[(long long)inst * Nh + k] = c;
          }
        }
      } else if (do_upd) {
        for (int k = jj; k < Nh; k += RT) {
          const cplx nh = Cr::postprocess(p, X, k);
          const cplx c = C[k];
          const double iA = IA[k];
          C[k] = make_double2((c.x * p.inv_dt + nh.x) * iA, (c.y * p.inv_dt + nh.y) * iA);
        }
      } else if (do_inv && !prep && live) {
        // n == n_iters: the final state's coefficients close the store (terminal condition of the adjoint, SH:573-584)
        cplx* fc = reinterpret_cast<cplx*>(sg + (long long)(NIT + 1) * M);
        for (int k = jj; k < Nh; k += RT) fc[k] = C[k];
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------ adjoint
template <class F, int NI_, int MB_ = SMO_SH_MB> struct Sh23Adj {
  typedef Sh23Params Params;
  typedef Sh23Core<F> Cr;
  static constexpr int NI = NI_, RT = F::RT, R1 = F::R1, R2 = F::R2, M = Cr::M;
  static constexpr int THREADS = NI_ * RT;
  static constexpr int NPHASES = 6;
  static constexpr int MIN_BLOCKS = MB_;
  static constexpr size_t SMEM = Cr::template smem_bytes<NI_>();
  SMO_HD static constexpr int sync_after(int ph) { return (32 % RT == 0 && ph != 5) ? 1 : 2; }
  struct State {
    double re[RT], im[RT];
  };
  // forward state the adjoint step `step` (1-based) linearises about: snapshot_index -2-m (discrete) / -1-m (continuous), m = step-1
  SMO_HD static int sidx_of(int step, int NIT, bool cont) { return cont ? (NIT + 1 - step) : (NIT - step); }
  template <int PH>
  SMO_HD static void phase(const Params& p, int work, int step, int tid, unsigned char* smem, State& st) {
    const int jj = tid % RT, li = tid / RT;
    const int inst = work * NI + li;
    const bool live = inst < p.batch;
    cplx* C = Cr::cbuf(smem, li);        // q
    cplx* X = C + Cr::NHMAX;
    double* IA = Cr::template iabuf<NI>(smem);
    const int Nh = p.Nh, NIT = p.n_iters;
    const bool cont = (p.flags & 1) != 0;
    const bool fin = (step == NIT + 1);
    const bool do_step = (step >= 1) && !fin;
    const double* sg = p.snaps + (long long)inst * p.sstride;
    if (PH == 0) {
      if (do_step) Cr::inv_stage1(p, C, false, jj, X, st.re, st.im);      // q
      if (fin) Cr::inv_stage1(p, C, !cont, jj, X, st.re, st.im);
    } else if (PH == 1) {
      if (do_step) {
        Cr::inv_stage2(X, jj, st.re, st.im);
        if (jj < R1) {
          // u_f straight from the store (grid values; the rows were sent towards L2 one step ahead)
          const cplx* uf = reinterpret_cast<const cplx*>(sg + (long long)sidx_of(step, NIT, cont) * M) + jj;
#pragma unroll
          for (int i = 0; i < R2; ++i) {
            cplx u = make_double2(0.0, 0.0);
            if (live) u = ldg_c(uf + R1 * i);
            st.re[i] = (3.6 * u.x - 3.0 * (u.x * u.x)) * st.re[i] - 2.0 * u.x;    // FWD_Solve_SH23.py:640
            st.im[i] = (3.6 * u.y - 3.0 * (u.y * u.y)) * st.im[i] - 2.0 * u.y;
          }
        }
      }
      if (fin) {
        Cr::inv_stage2(X, jj, st.re, st.im);
        if (jj < R1 && live) {
          cplx* dst = reinterpret_cast<cplx*>(p.grad + (long long)inst * M) + jj;
#pragma unroll
          for (int k2 = 0; k2 < R2; ++k2) dst[R1 * k2] = make_double2(st.re[k2], st.im[k2]);
        }
      }
    } else if (PH == 2) {
      if (do_step) Cr::fwd_stage1(p, jj, X, st.re, st.im);
    } else if (PH == 3) {
      if (do_step) Cr::fwd_stage2_regs(X, jj, st.re, st.im);
    } else if (PH == 4) {
      if (do_step) Cr::spectrum_store(X, jj, st.re, st.im);
    } else {
      if (step == 0) {
        // terminal condition: discrete q0 = -2 u^N/(1/dt + L) (FWD_Solve_SH23.py:584), continuous q0 = 0
        const cplx* fc = reinterpret_cast<const cplx*>(sg + (long long)(NIT + 1) * M);
        for (int k = tid; k < Nh; k += THREADS) IA[k] = 1.0 / sh_A(p, k);
        for (int k = jj; k < Nh; k += RT) {
          cplx q = make_double2(0.0, 0.0);
          if (!cont && live) { const double iA = 1.0 / sh_A(p, k); const cplx f = fc[k]; q = make_double2(-2.0 * f.x * iA, -2.0 * f.y * iA); }
          C[k] = q;
        }
      } else if (do_step) {
        for (int k = jj; k < Nh; k += RT) {
          const cplx rh = Cr::postprocess(p, X, k);
          const cplx c = C[k];
          const double iA = IA[k];
          C[k] = make_double2((c.x * p.inv_dt + rh.x) * iA, (c.y * p.inv_dt + rh.y) * iA);
        }
      }
      // the forward state of the NEXT adjoint step on its way into L2 (M doubles = M/16 lines of 128 bytes)
      if (step < NIT && live) {
        const double* nx = sg + (long long)sidx_of(step + 1, NIT, cont) * M;
        for (int l = jj; l < M / 16; l += RT) prefetch_l2(nx + 16 * l);
      }
    }
  }
};

}  // namespace smo
