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


// ================================================================================================================
// Fused x-pass of the time loops, software pipelined (X_FWD: U x B; X_ADJ: (curl G) x U and (curl G) x B_f).
//
// Same mathematics as XPass above, restructured for HBM throughput:
//   * the spectral input tile (NFI fields x Nh rows x T columns) and the velocity tile (3 x M x T doubles) of the
//     NEXT column tile are streamed into shared memory with cp.async while the current tile is transformed (the
//     spectral buffer is free after the assembly phase, the velocity buffer after the product phase);
//   * the velocity field is read from a tile-major copy Ut[ncols/4][3][M][4] written once per forward call, so a
//     tile's velocity data is one contiguous 18 KB block per 4 columns (perfectly coalesced, no 32-byte rows);
//   * column tiles can be restricted to a z range (chunked launches keep the P2 arrays L2-resident between the
//     y pass that produces them, this kernel and the y pass that consumes its output).
// Phases: 0 wait spectral tile | 1 assemble + inverse stage 1 | 2 prefetch next spectral tile, inverse stage 2 |
//   3 grid values -> shared, wait velocity tile | 4 products + forward stage 1 | 5 prefetch next velocity tile,
//   exchange | 6 forward stage 2 | 7 spectrum -> shared | 8 split, scale, truncate, store.
// ================================================================================================================
struct XFParams {
  const cplx* sin[MAXF];     // spectral inputs  [Nh][ncols]
  cplx* sout[MAXF];          // spectral outputs [Nh][ncols]
  const double* ut;          // velocity, tile-major [ncols/4][3][M][4]
  int nwork, nsteps;         // nwork = number of column tiles of this launch
  long long ncols;
  int Nh;
  int tiles_per_row, row_tiles, tile0;   // column tile of work w: (w / tiles_per_row) * row_tiles + tile0 + w % tiles_per_row
  double scale;
  const cplx* tw;
};

template <class F, int T_, int MODE, int NFI, int NFO, bool JFAST> struct XFused {
  typedef XFParams Params;
  typedef typename F::Swapped FS;
  static constexpr bool V2 = true;
  static constexpr int T = T_, HP = T_ / 2, TB = T_ / 4;
  static constexpr int NJI = NFI * HP, NJO = NFO * HP, NJ = (NJI > NJO) ? NJI : NJO;
  static constexpr int R1 = F::R1, R2 = F::R2, M = F::M, RT = F::RT;
  static constexpr int NH = M / 3;                   // retained kx modes (dealias 3/2: Npts/2 = M/3)
  static constexpr int THREADS = NJ * RT;
  static constexpr int NPHASES = 9;
  static constexpr int MIN_BLOCKS = (THREADS <= 96) ? 2 * SMO_X_MB : SMO_X_MB;
  static constexpr int XLEN = (F::XP > FS::XP) ? ((F::XP > M) ? F::XP : M) : ((FS::XP > M) ? FS::XP : M);
  // With JFAST and exactly one warp per field (FT = 32: both column pairs of a field, 16 stage threads each) every
  // FFT exchange, the spectral tile and the spectrum hand-over stay inside one warp: only the two barriers around the
  // product phase (which reads the other fields' grid values and the shared velocity tile) are CTA-wide, so the
  // warps of a CTA drift apart and overlap each other's shared-memory and fp64 phases.
  static constexpr int FT = HP * RT;                 // threads per field
  static constexpr bool WSYNC = JFAST && (FT == 32);
  SMO_HD static constexpr int sync_after(int ph) { return (!WSYNC || ph == 3 || ph == 4) ? 2 : 1; }
  // Thread order and shared-memory layouts are chosen so that every quarter-warp access is bank-conflict free:
  //   JFAST = false (T = 8): lanes run over the HP = 4 column pairs, then over the stage threads;  X is [e][NJ]
  //   JFAST = true  (T = 4): lanes run over the RT stage threads, then over the HP = 2 pairs;      X is [q][XLP]
  static constexpr int XLP = XLEN + 4;               // q pitch of the JFAST exchange layout (= 4 mod 8)
  // The cp.async targets (spectral tile, velocity tile) are dense 128-byte lines with an XOR swizzle inside each
  // line instead of padded pitches: LDGSTS writes a line in one wavefront only when 8 lanes cover one ALIGNED line
  // (measured: padded / scattered targets cost 8-32 wavefronts per instruction), and the swizzle keeps the reads
  // of the transform phases conflict free.
  static constexpr int SIN_ELEMS = NFI * NH * T_;    // cplx (16-byte units)
  static constexpr int UBU = 3 * M * 2;              // 16-byte units per 4-column velocity block
  static constexpr int SU_UNITS = TB * UBU;
  static constexpr int X_ELEMS = JFAST ? NJ * XLP : NJ * XLEN;
  static constexpr size_t SMEM = (size_t)(SIN_ELEMS + SU_UNITS + X_ELEMS) * sizeof(cplx);
  static_assert(SIN_ELEMS % 8 == 0 && SU_UNITS % 8 == 0, "cp.async regions must be whole 128-byte lines");
  static_assert(T_ % 4 == 0, "column tiles are multiples of the 4-column velocity blocks");
  static_assert(MODE == X_FWD || MODE == X_ADJ, "fused modes only");
  static_assert(JFAST == (T_ == 4), "T = 4 uses the stage-thread-fastest order, T = 8 the pair-fastest order");
  static_assert(T_ == 4 || T_ == 8, "supported column tiles");
  static_assert(NH % 2 == 0, "two spectral rows per 128-byte line need an even number of retained modes");
  struct State {
    double re[RT], im[RT];
    double wr, wi;     // w_M^jj = exp(-2 pi i jj / M): base of this thread's inter-stage twiddles
    int it;
  };

  SMO_HD static cplx* sin_buf(unsigned char* s) { return reinterpret_cast<cplx*>(s); }
  SMO_HD static cplx* su_buf(unsigned char* s) { return sin_buf(s) + SIN_ELEMS; }
  SMO_HD static cplx* x_buf(unsigned char* s) { return su_buf(s) + SU_UNITS; }
  // unit index of spectral entry (field f, row, tile column col)
  SMO_HD static int si(int f, int row, int col) {
    if (T == 8) return (f * NH + row) * 8 + (col ^ (row & 7));
    return (((f * NH + row) >> 1) << 3) + ((((row & 1) << 2) + col) ^ ((row >> 1) & 3));   // T == 4: two rows per line
  }
  // unit index of the velocity pair (component cidx, grid row n, column pair pp)
  SMO_HD static int ui(int cidx, int n, int pp) {
    if (JFAST) return ((cidx * M + n) * 2 + pp) ^ ((n >> 2) & 1);                            // [c][n][pair], one block
    return ((pp >> 1) * UBU + (cidx * M + n) * 2 + (pp & 1)) ^ (((pp >> 1) & 1) << 2);      // [block][c][n][half]
  }

  SMO_HD static void decode(int tid, int& f, int& pp, int& jj) {
    if (JFAST) { jj = tid % RT; pp = (tid / RT) % HP; } else { pp = tid % HP; jj = (tid / HP) % RT; }
    f = tid / (HP * RT);
  }
  SMO_HD static int xe(int e, int q) { return JFAST ? q * XLP + e : e * NJ + q; }                 // exchange / spectrum
  SMO_HD static int gi(int f, int n, int pp) { return JFAST ? (f * HP + pp) * XLP + n : (f * M + n) * HP + pp; }   // grid values
  SMO_HD static long long tile_of(const Params& p, int work) {
    return (long long)(work / p.tiles_per_row) * p.row_tiles + p.tile0 + (work % p.tiles_per_row);
  }
  // asynchronous loads of the spectral tile / the velocity tile of `work`
  SMO_HD static void load_sin(const Params& p, int work, const Ctx& c) {
    cplx* S = sin_buf(c.smem);
    const long long col0 = tile_of(p, work) * T;
    if constexpr (WSYNC) {                                    // every warp streams in the field it transforms
      const int f = c.tid / FT;
      if (f < NFI)
        for (int q = c.tid % FT; q < NH * T; q += FT) {
          const int tc = q % T, row = q / T;
          cp_async16(&S[si(f, row, tc)], p.sin[f] + (long long)row * p.ncols + col0 + tc);
        }
    } else {
      for (int q = c.tid; q < NFI * NH * T; q += THREADS) {   // chunks (f, row, column), column fastest
        const int tc = q % T, r = q / T;
        const int row = r % NH, f = r / NH;
        cp_async16(&S[si(f, row, tc)], p.sin[f] + (long long)row * p.ncols + col0 + tc);
      }
    }
  }
  SMO_HD static void load_su(const Params& p, int work, const Ctx& c) {
    cplx* U = su_buf(c.smem);
    const double* src = p.ut + tile_of(p, work) * TB * (3LL * M * 4);
    for (int q = c.tid; q < SU_UNITS; q += THREADS) {          // gmem chunk q = (block, component, row, half)
      const int b = q / UBU, r = q % UBU;
      cp_async16(&U[ui(r / (2 * M), (r / 2) % M, b * 2 + (r & 1))], src + 2 * q);
    }
  }
  // velocity pair (columns 2pp, 2pp+1 of the tile) of component cidx at grid row n
  SMO_HD static cplx su_pair(const cplx* U, int cidx, int n, int pp) { return U[ui(cidx, n, pp)]; }

  SMO_HD static void init(const Params& p, const Ctx& c, State& st) {
    int f, pp, jj;
    decode(c.tid, f, pp, jj);
    const cplx w = ldg_c(p.tw + jj);
    st.wr = w.x; st.wi = w.y;
    st.it = 0;
  }

  // product needed by output field fo at grid row n (pair of columns)
  SMO_HD static void product(const cplx* Gs, const cplx* U, int fo, int pp, int n, double& e0, double& e1) {
    const int cc = (fo >= 3) ? fo - 3 : fo;
    const int c1 = (cc + 1) % 3, c2 = (cc + 2) % 3;
    cplx a1, a2, b1, b2;
    if (MODE == X_FWD) {
      a1 = su_pair(U, c1, n, pp); a2 = su_pair(U, c2, n, pp);          // E = U x B
      b1 = Gs[gi(c1, n, pp)]; b2 = Gs[gi(c2, n, pp)];
    } else {
      a1 = Gs[gi(c1, n, pp)]; a2 = Gs[gi(c2, n, pp)];                  // W x U (fo < 3), W x B_f (fo >= 3)
      if (fo < 3) { b1 = su_pair(U, c1, n, pp); b2 = su_pair(U, c2, n, pp); }
      else { b1 = Gs[gi(3 + c1, n, pp)]; b2 = Gs[gi(3 + c2, n, pp)]; }
    }
    e0 = a1.x * b2.x - a2.x * b1.x;
    e1 = a1.y * b2.y - a2.y * b1.y;
  }

  template <int PH>
  SMO_HD static void phase2(const Params& p, int work, int /*step*/, const Ctx& c, State& st) {
    cplx* S = sin_buf(c.smem);
    cplx* X = x_buf(c.smem);
    const cplx* U = su_buf(c.smem);
    int f, pp, jj;
    decode(c.tid, f, pp, jj);
    const int q = f * HP + pp;
    const bool more = work + c.ncta < p.nwork;
    if (PH == 0) {
      if (st.it == 0) {
        load_sin(p, work, c); cp_async_commit();
        load_su(p, work, c); cp_async_commit();
      }
      cp_async_wait<1>();                       // the spectral tile of this work item has landed
    }
    if (PH == 1) {
      if (f < NFI && jj < R2) {
        const int j = jj;
#pragma unroll
        for (int i = 0; i < R1; ++i) {
          const int n = j + R2 * i;
          double zr = 0.0, zi = 0.0;
          if (n < NH) {
            const cplx v = S[si(f, n, 2 * pp)], w = S[si(f, n, 2 * pp + 1)];
            if (n == 0) { zr = v.x; zi = w.x; } else { zr = v.x - w.y; zi = v.y + w.x; }
          } else if (n > M - NH) {
            const int m = M - n;
            const cplx v = S[si(f, m, 2 * pp)], w = S[si(f, m, 2 * pp + 1)];
            zr = v.x + w.y; zi = w.x - v.y;
          }
          st.re[i] = zr; st.im[i] = zi;
        }
        RegFFT<R1, +1>::run(as_arr<R1>(st.re), as_arr<R1>(st.im));
        twiddle_powers<R1>(st.re, st.im, st.wr, -st.wi);      // conj: inverse direction
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) X[xe(j * F::SK + k1, q)] = make_double2(st.re[k1], st.im[k1]);
      }
    }
    if (PH == 2) {
      if (more) load_sin(p, work + c.ncta, c);   // the spectral buffer was consumed in phase 1
      cp_async_commit();
      if (f < NFI && jj < R1) {
        const int k1 = jj;
#pragma unroll
        for (int j = 0; j < R2; ++j) {
          const cplx v = X[xe(j * F::SK + k1, q)];
          st.re[j] = v.x; st.im[j] = v.y;
        }
        stage2<F, +1>(st.re, st.im);
      }
    }
    if (PH == 3) {
      if (f < NFI && jj < R1) {
        const int k1 = jj;
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) X[gi(f, k1 + R1 * k2, pp)] = make_double2(st.re[k2], st.im[k2]);
      }
      cp_async_wait<1>();                       // the velocity tile of this work item has landed
    }
    if (PH == 4) {
      if (f < NFO && jj < R1) {
        const int j = jj;   // stage-1 thread of the swapped factorisation owns rows j + R1*i
#pragma unroll
        for (int i = 0; i < R2; ++i) product(X, U, f, pp, j + R1 * i, st.re[i], st.im[i]);
        RegFFT<R2, -1>::run(as_arr<R2>(st.re), as_arr<R2>(st.im));
        twiddle_powers<R2>(st.re, st.im, st.wr, st.wi);
      }
    }
    if (PH == 5) {
      if (more) load_su(p, work + c.ncta, c);    // the velocity buffer was consumed in phase 4
      cp_async_commit();
      if (f < NFO && jj < R1) {
#pragma unroll
        for (int k1 = 0; k1 < R2; ++k1) X[xe(jj * FS::SK + k1, q)] = make_double2(st.re[k1], st.im[k1]);
      }
    }
    if (PH == 6) {
      if (f < NFO && jj < R2) {
        const int k1 = jj;
#pragma unroll
        for (int j = 0; j < R1; ++j) {
          const cplx v = X[xe(j * FS::SK + k1, q)];
          st.re[j] = v.x; st.im[j] = v.y;
        }
        stage2<FS, -1>(st.re, st.im);
      }
    }
    if (PH == 7) {
      if (f < NFO && jj < R2) {
#pragma unroll
        for (int k2 = 0; k2 < R1; ++k2) {
          const int k = jj + R2 * k2;   // only the retained modes and their mirror images are needed by phase 8
          if (k < NH || k > M - NH) X[xe(k, q)] = make_double2(st.re[k2], st.im[k2]);
        }
      }
    }
    if (PH == 8) {
      // own thread order (column pairs fastest) so that a row's T columns are stored by adjacent lanes
      const int pp8 = c.tid % HP, kk = (c.tid / HP) % RT, f8 = c.tid / (HP * RT), q8 = f8 * HP + pp8;
      if (f8 < NFO) {
        cplx* O = p.sout[f8] + tile_of(p, work) * T + 2 * pp8;
        const double h = 0.5 * p.scale;
        for (int k = kk; k < NH; k += RT) {
          const cplx zk = X[xe(k, q8)];
          const cplx zm = X[xe((M - k) % M, q8)];
          O[(long long)k * p.ncols] = make_double2(h * (zk.x + zm.x), h * (zk.y - zm.y));
          O[(long long)k * p.ncols + 1] = make_double2(h * (zk.y + zm.y), h * (zm.x - zk.x));
        }
      }
      st.it++;
    }
  }
};

// one-off re-layout of the velocity field: grid [3][M][ncols] -> tile-major [ncols/4][3][M][4]
struct UTileParams {
  const double* in[3];
  double* out;
  int nwork, nsteps;
  long long ncols;
  int M;
};
struct UTile {
  typedef UTileParams Params;
  static constexpr int THREADS = 256;
  static constexpr int NPHASES = 1;
  static constexpr int MIN_BLOCKS = 1;
  static constexpr size_t SMEM = 0;
  struct State {};
  template <int PH> SMO_HD static void phase(const Params& p, int work, int, int tid, unsigned char*, State&) {
    // work item = (row n, block of 256 columns)
    const long long per_row = (p.ncols + THREADS - 1) / THREADS;
    const int n = (int)(work / per_row);
    const long long col = (work % per_row) * THREADS + tid;
    if (col >= p.ncols) return;
#pragma unroll
    for (int cc = 0; cc < 3; ++cc)
      p.out[(((col / 4) * 3 + cc) * p.M + n) * 4 + (col % 4)] = p.in[cc][(long long)n * p.ncols + col];
  }
};

}  // namespace smo
