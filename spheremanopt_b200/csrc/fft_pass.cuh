// Batched complex-to-complex FFT pass along one axis of a 3-D coefficient/pencil array, with the
// Dedalus-style zero-padding (coefficient -> dealiased grid) or truncation (grid -> coefficient) fused in.
//
// Replaces, for the y and z axes, what the reference triggers implicitly through every field['g'] / ['c']
// access and solver.step (FWD_Solve_KDyn.py:635-641, 955-961): FFTW c2c along the axis plus the pad/truncate
// copy between scales=1 and scales=3/2 ([D2-2], [D2-3] of SURVEY.md section 8(c)).
//
// A CTA handles a tile of T lines; each line is transformed by RT = max(R1,R2) threads in two register
// stages with one shared-memory exchange (fft_core.cuh).  Global loads/stores go straight from/to registers:
//   TFAST = true  : adjacent lines are adjacent in memory (strided axis); lanes run over the T lines so each
//                   warp access covers T*16 B contiguous per FFT row.
//   TFAST = false : the FFT axis itself is contiguous; lanes run over the stage threads of one line.
// Zero input rows (padding) are never loaded and dropped output rows (truncation) never stored, so HBM traffic
// is the pruned C / P1 / P2 figure of SURVEY.md section 8(d).
#pragma once
#include "fft_core.cuh"

// tuning knobs (overridable with -D for experiments; defaults are the measured best on B200)
#ifndef SMO_PASS_MB
#define SMO_PASS_MB 4   // resident CTAs per SM the register allocation of the c2c passes is bounded for
#endif

namespace smo {

struct PassParams {
  const cplx* in[MAXF];
  cplx* out[MAXF];
  int nwork, nsteps;        // nwork = nfields * nA * tilesB, nsteps = 1
  int nfields, nA, nB, tilesB;
  long long in_sA, in_sB, out_sA, out_sB;   // line (a,b) base offsets, in complex elements
  // full-length side (M entries):    off(n) = (n / split) * blk + (n % split) * sN
  // compact side (2*kmax+1 entries): off(c) = c * sN
  long long in_sN, out_sN, in_blk, out_blk;
  int in_split, out_split;
  int pad;                  // 1: input compact, output full (inverse direction); 0: input full, output compact
  int kmax;
  double scale;             // applied to outputs
  const cplx* tw;           // exp(-2 pi i m / M), m < M
};

template <class F, int DIR, bool TFAST, int T_> struct FftPass {
  typedef PassParams Params;
  static constexpr int T = T_;
  static constexpr int THREADS = T_ * F::RT;
  static constexpr int NPHASES = 2;
  static constexpr int MIN_BLOCKS = SMO_PASS_MB;
  static constexpr size_t SMEM = (size_t)T_ * F::XP * sizeof(cplx);
  struct State {
    double re[F::RT], im[F::RT];
  };

  SMO_HD static void decode(const Params& p, int work, int tid, int& f, int& a, int& b, int& t, int& jj) {
    const int per_field = p.nA * p.tilesB;
    f = work / per_field;
    const int r = work - f * per_field;
    a = r / p.tilesB;
    const int bt = r - a * p.tilesB;
    if (TFAST) { t = tid % T; jj = tid / T; } else { jj = tid % F::RT; t = tid / F::RT; }
    b = bt * T + t;
  }
  SMO_HD static int xidx(int t, int e) { return TFAST ? e * T + t : t * F::XP + e; }

  template <int PH>
  SMO_HD static void phase(const Params& p, int work, int /*step*/, int tid, unsigned char* smem, State& st) {
    cplx* X = reinterpret_cast<cplx*>(smem);
    int f, a, b, t, jj;
    decode(p, work, tid, f, a, b, t, jj);
    const bool live = b < p.nB;
    constexpr int M = F::M;
    if (PH == 0) {
      if (jj < F::R2) {
        const int j = jj;
        const cplx* src = p.in[f] + (long long)a * p.in_sA + (long long)b * p.in_sB;
#pragma unroll
        for (int i = 0; i < F::R1; ++i) {
          const int n = j + F::R2 * i;
          double vr = 0.0, vi = 0.0;
          if (live) {
            if (p.pad) {
              const int c = compact_index(n, M, p.kmax);
              if (c >= 0) { const cplx v = src[(long long)c * p.in_sN]; vr = v.x; vi = v.y; }
            } else {
              const cplx v = src[(long long)(n / p.in_split) * p.in_blk + (long long)(n % p.in_split) * p.in_sN];
              vr = v.x; vi = v.y;
            }
          }
          st.re[i] = vr; st.im[i] = vi;
        }
        stage1<F, DIR>(st.re, st.im, j, p.tw);
#pragma unroll
        for (int k1 = 0; k1 < F::R1; ++k1) X[xidx(t, j * F::SK + k1)] = make_double2(st.re[k1], st.im[k1]);
      }
    } else {
      if (jj < F::R1) {
        const int k1 = jj;
#pragma unroll
        for (int j = 0; j < F::R2; ++j) {
          const cplx v = X[xidx(t, j * F::SK + k1)];
          st.re[j] = v.x; st.im[j] = v.y;
        }
        stage2<F, DIR>(st.re, st.im);
        if (live) {
          cplx* dst = p.out[f] + (long long)a * p.out_sA + (long long)b * p.out_sB;
#pragma unroll
          for (int k2 = 0; k2 < F::R2; ++k2) {
            const int k = k1 + F::R1 * k2;
            const cplx v = make_double2(st.re[k2] * p.scale, st.im[k2] * p.scale);
            if (p.pad) {
              dst[(long long)(k / p.out_split) * p.out_blk + (long long)(k % p.out_split) * p.out_sN] = v;
            } else {
              const int c = compact_index(k, M, p.kmax);
              if (c >= 0) dst[(long long)c * p.out_sN] = v;
            }
          }
        }
      }
    }
  }
};

}  // namespace smo
