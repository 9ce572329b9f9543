// Real-axis (x) pass of the kinematic-dynamo pseudo-spectral step, fused with the grid-space products.
//
// Replaces what the reference evaluates through Dedalus for the RHS of the induction equation and of its
// adjoint: the c2r transform of every operand, the pointwise products U x B (FWD_Solve_KDyn.py:417-419),
// (curl G) x U and (curl G) x B_f (FWD_Solve_KDyn.py:846-848, 857-859, 875-877) on the 3/2-dealiased grid,
// and the r2c transform + truncation of the result - without the M^3 real grid ever touching HBM.
//
// Layout: "x-spectral" arrays are complex [Nh][ncols] (kx rows, ncols = M*M or M*Mz_local grid columns (y,z));
// real grid arrays are double [M][ncols].  A CTA handles T adjacent columns.  Two adjacent columns of the same
// field are transformed by ONE complex length-M FFT (z = a + i b; a, b real): for the inverse the Hermitian
// extension of the two half spectra is assembled on the fly (Im of the kx = 0 entry dropped, as FFTW's c2r
// does); for the forward the two spectra are separated from Z[k] and conj(Z[M-k]).
//
// Phases (fused modes):  0 assemble + inverse stage 1 -> exchange | 1 inverse stage 2 -> registers |
//   2 registers -> grid values in shared memory | 3 products + forward stage 1 | 4 -> exchange |
//   5 forward stage 2 | 6 spectrum -> shared | 7 split, scale, truncate, store.
// The forward transform uses the swapped factorisation (fft_core.cuh) so that all exchanges are conflict free.
#pragma once
#include "fft_core.cuh"

#ifndef SMO_X_MB
#define SMO_X_MB 2      // resident CTAs per SM the register allocation of the x passes is bounded for
#endif

namespace smo {

struct XParams {
  const cplx* sin[MAXF];     // spectral inputs
  const double* gin[MAXF];   // real-grid inputs (U for the fused modes, fields for R2C)
  cplx* sout[MAXF];
  double* gout[MAXF];
  int nwork, nsteps;         // nwork = ncols / T
  long long ncols;
  int Nh;
  double scale;
  const cplx* tw;
};

enum { X_C2R = 0, X_R2C = 1, X_FWD = 2, X_ADJ = 3 };

template <class F, int T_, int MODE, int NFI, int NFO> struct XPass {
  typedef XParams Params;
  typedef typename F::Swapped FS;
  static constexpr int T = T_, HP = T_ / 2;
  static constexpr int NJI = NFI * HP, NJO = NFO * HP, NJ = (NJI > NJO) ? NJI : NJO;
  static constexpr int R1 = F::R1, R2 = F::R2, M = F::M, RT = F::RT;
  static constexpr int THREADS = NJ * RT;
  static constexpr int NPHASES = (MODE == X_C2R) ? 2 : 8;
  static constexpr int MIN_BLOCKS = SMO_X_MB;
  static constexpr int XLEN = (F::XP > FS::XP) ? ((F::XP > M) ? F::XP : M) : ((FS::XP > M) ? FS::XP : M);
  static constexpr size_t SMEM = (size_t)NJ * XLEN * sizeof(cplx);
  static_assert(T_ % 2 == 0, "columns are processed in pairs");
  struct State {
    double re[RT], im[RT];
  };

  SMO_HD static void decode(int tid, int& f, int& pp, int& jj) {
    pp = tid % HP;
    jj = (tid / HP) % RT;
    f = tid / (HP * RT);
  }

  // value pair (columns col, col+1) of the product needed by output field fo at grid row n
  SMO_HD static void product(const Params& p, const cplx* Gs, int fo, int pp, long long col, int n, double& e0,
                             double& e1) {
    if (MODE == X_R2C) {
      const cplx v = *reinterpret_cast<const cplx*>(p.gin[fo] + (long long)n * p.ncols + col);
      e0 = v.x; e1 = v.y;
      return;
    }
    const int c = (fo >= 3) ? fo - 3 : fo;
    const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
    cplx a1, a2, b1, b2;   // (col, col+1) pairs of the two operands' components c1, c2
    if (MODE == X_FWD) {
      // E = U x B : U from HBM (real grid), B from shared memory
      a1 = *reinterpret_cast<const cplx*>(p.gin[c1] + (long long)n * p.ncols + col);
      a2 = *reinterpret_cast<const cplx*>(p.gin[c2] + (long long)n * p.ncols + col);
      b1 = Gs[(c1 * M + n) * HP + pp];
      b2 = Gs[(c2 * M + n) * HP + pp];
    } else {
      // fo < 3 : W x U ; fo >= 3 : W x B_f   (W = fields 0..2, B_f = fields 3..5 in shared memory)
      a1 = Gs[(c1 * M + n) * HP + pp];
      a2 = Gs[(c2 * M + n) * HP + pp];
      if (fo < 3) {
        b1 = *reinterpret_cast<const cplx*>(p.gin[c1] + (long long)n * p.ncols + col);
        b2 = *reinterpret_cast<const cplx*>(p.gin[c2] + (long long)n * p.ncols + col);
      } else {
        b1 = Gs[((3 + c1) * M + n) * HP + pp];
        b2 = Gs[((3 + c2) * M + n) * HP + pp];
      }
    }
    e0 = a1.x * b2.x - a2.x * b1.x;
    e1 = a1.y * b2.y - a2.y * b1.y;
  }

  template <int PH>
  SMO_HD static void phase(const Params& p, int work, int /*step*/, int tid, unsigned char* smem, State& st) {
    cplx* X = reinterpret_cast<cplx*>(smem);
    int f, pp, jj;
    decode(tid, f, pp, jj);
    const int q = f * HP + pp;
    const long long col = (long long)work * T + 2 * pp;

    // ---------------- inverse half: spectral -> grid values --------------------------------
    if (PH == 0 && MODE != X_R2C) {
      if (f < NFI && jj < R2) {
        const int j = jj;
        const cplx* A = p.sin[f] + col;
#pragma unroll
        for (int i = 0; i < R1; ++i) {
          const int n = j + R2 * i;
          double zr = 0.0, zi = 0.0;
          if (n < p.Nh) {
            const cplx v = A[(long long)n * p.ncols], w = A[(long long)n * p.ncols + 1];
            if (n == 0) { zr = v.x; zi = w.x; } else { zr = v.x - w.y; zi = v.y + w.x; }
          } else if (n > M - p.Nh) {
            const int m = M - n;
            const cplx v = A[(long long)m * p.ncols], w = A[(long long)m * p.ncols + 1];
            zr = v.x + w.y; zi = w.x - v.y;
          }
          st.re[i] = zr; st.im[i] = zi;
        }
        stage1<F, +1>(st.re, st.im, j, p.tw);
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) X[(j * F::SK + k1) * NJ + q] = make_double2(st.re[k1], st.im[k1]);
      }
    }
    if (PH == 1 && MODE != X_R2C) {
      if (f < NFI && jj < R1) {
        const int k1 = jj;
#pragma unroll
        for (int j = 0; j < R2; ++j) {
          const cplx v = X[(j * F::SK + k1) * NJ + q];
          st.re[j] = v.x; st.im[j] = v.y;
        }
        stage2<F, +1>(st.re, st.im);
        if (MODE == X_C2R) {
#pragma unroll
          for (int k2 = 0; k2 < R2; ++k2) {
            const int n = k1 + R1 * k2;
            *reinterpret_cast<cplx*>(p.gout[f] + (long long)n * p.ncols + col) = make_double2(st.re[k2], st.im[k2]);
          }
        }
      }
    }
    if (MODE == X_C2R) return;
    if (PH == 2 && MODE != X_R2C) {
      if (f < NFI && jj < R1) {
        const int k1 = jj;
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) X[(f * M + k1 + R1 * k2) * HP + pp] = make_double2(st.re[k2], st.im[k2]);
      }
    }
    // ---------------- forward half: grid products -> truncated spectra ---------------------
    if (PH == 3) {
      if (f < NFO && jj < R1) {
        const int j = jj;   // stage-1 thread of the swapped factorisation owns rows j + R1*i
#pragma unroll
        for (int i = 0; i < R2; ++i) product(p, X, f, pp, col, j + R1 * i, st.re[i], st.im[i]);
        stage1<FS, -1>(st.re, st.im, j, p.tw);
      }
    }
    if (PH == 4) {
      if (f < NFO && jj < R1) {
#pragma unroll
        for (int k1 = 0; k1 < R2; ++k1) X[(jj * FS::SK + k1) * NJ + q] = make_double2(st.re[k1], st.im[k1]);
      }
    }
    if (PH == 5) {
      if (f < NFO && jj < R2) {
        const int k1 = jj;
#pragma unroll
        for (int j = 0; j < R1; ++j) {
          const cplx v = X[(j * FS::SK + k1) * NJ + q];
          st.re[j] = v.x; st.im[j] = v.y;
        }
        stage2<FS, -1>(st.re, st.im);
      }
    }
    if (PH == 6) {
      if (f < NFO && jj < R2) {
#pragma unroll
        for (int k2 = 0; k2 < R1; ++k2) X[(jj + R2 * k2) * NJ + q] = make_double2(st.re[k2], st.im[k2]);
      }
    }
    if (PH == 7) {
      if (f < NFO) {
        cplx* O = p.sout[f] + col;
        const double h = 0.5 * p.scale;
        for (int k = jj; k < p.Nh; k += RT) {
          const cplx zk = X[k * NJ + q];
          const cplx zm = X[((M - k) % M) * NJ + q];
          O[(long long)k * p.ncols] = make_double2(h * (zk.x + zm.x), h * (zk.y - zm.y));
          O[(long long)k * p.ncols + 1] = make_double2(h * (zk.y + zm.y), h * (zm.x - zk.x));
        }
      }
    }
  }
};

}  // namespace smo
