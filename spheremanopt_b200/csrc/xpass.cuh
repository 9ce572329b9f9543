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
  static constexpr int MIN_BLOCKS = (F::RT > 16) ? 1 : SMO_X_MB;
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
// Same mathematics as XPass above, restructured around the resource that binds it on B200 - the shared-memory data
// pipe (ncu: l1tex data-pipe wavefronts ~75 % busy, HBM ~30 %, fp64 pipe ~35 %):
//   * a CTA handles T = 4 adjacent (y,z) columns = 2 column pairs of every field; one warp per field (FT = 32 threads:
//     2 pairs x 16 stage threads), so every FFT exchange, the spectral tile and the spectrum hand-over stay inside a
//     warp and only the two barriers around the product phase are CTA-wide (sync_after);
//   * the spectral input tile (NF fields x Nh rows x 4 columns) and the velocity tile (3 x M x 4 doubles) of the NEXT
//     column tile are streamed into shared memory with cp.async while the current tile is transformed; the velocity
//     field is read from a tile-major copy Ut[ncols/4][3][M][4] written once per forward call;
//   * the stages that use R2 (< R1) threads per FFT run on the first HP*R2 lanes of the warp ("narrow" mapping:
//     lane = pair*R2 + j) instead of leaving 4 idle lanes in every half warp: a 128-bit access of 24 dense lanes costs 3
//     wavefronts, of 2 x 12 lanes in two half warps 4;
//   * every thread keeps the grid values of its own field in registers and computes the cross-product component whose
//     formula contains them, so a product costs 3 shared-memory reads instead of 4.
// Phases: 0 wait spectral tile | 1 assemble + inverse stage 1 (narrow) | 2 prefetch next spectral tile, inverse stage 2
//   (wide) | 3 grid values -> shared, wait velocity tile | 4 products + forward stage 1 (wide) | 5 prefetch next velocity
//   tile, exchange | 6 forward stage 2 (narrow) | 7 spectrum -> shared | 8 split, scale, truncate, store.
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
  int accumulate;            // X_ADJ: outputs 3..5 (x-spectra of (curl G) x B_f) are ADDED to sout[3..5] instead of stored
  int bulk_u;                // 1: the velocity tile arrives by ONE TMA bulk copy (cp.async.bulk + mbarrier) instead of 16-byte cp.async
  int tma_sin;               // 1: the spectral tiles arrive by TMA tensor copies (one per field, box = 4 columns x NH rows, hardware
                             // swizzle = si()) described by tm[0] (fields 0..2) / tm[1] (fields 3..5), slice tz[f] of the map
  int tz[MAXF];
  SmoTensorMap tm[2];
  double* gacc;              // X_ADJ, GACC kernels: running sum of (curl G) x B_f ON THE GRID, tile-major like `ut`
  double* jpart;             // INTEG forward: [gridDim] per-CTA sums of |B|^2 over the grid points this launch visited
};

// INTEG = Cost_function "Integrated" (KD:655-669, 861-864): the forward pass also sums |B^n|^2 over its grid points
// (deterministic per-CTA partials), the adjoint pass adds the source -2 B_f to the (curl G) x U products.
// GACC (adjoint only): the products (curl G) x B_f are not transformed at all inside the sweep: nu = -dt sum_m P_k[F(...)] is
// linear, so the threads that hold them add them to a running sum on the real grid (tile-major, the layout of `ut`: a warp's
// read-modify-write of one row group is 512 contiguous bytes, straight from registers) and ONE r2c transform after the sweep
// replaces 3 of the 6 forward FFTs of every adjoint step; the warps of B_f skip the forward half of the pass.
template <class F, int MODE, bool INTEG = false, bool GACC = false> struct XFused {
  typedef XFParams Params;
  typedef typename F::Swapped FS;
  static constexpr bool V2 = true;
  static constexpr int T = 4, HP = 2;
  static constexpr int NF = (MODE == X_FWD) ? 3 : 6;  // fields in = fields out
  static constexpr int NJ = NF * HP;
  static constexpr int R1 = F::R1, R2 = F::R2, M = F::M, RT = F::RT;
  static constexpr int NH = M / 3;                   // retained kx modes (dealias 3/2: Npts/2 = M/3)
  // lanes reserved per FFT: RT, or a whole warp when RT does not divide 32 (384 = 24 x 16: 24 stage threads) - an FFT is
  // then private to one warp ("PAIRWARP": every warp owns one column pair of one field) and the warp-level barriers apply
  static constexpr int LP = (RT > 16 && RT < 32) ? 32 : RT;
  static constexpr bool PAIRWARP = (LP != RT);
  static constexpr int FT = HP * LP;                 // threads per field
  static constexpr int NARROW = HP * R2;             // active threads per field in the R2-thread stages (dense mapping)
  static constexpr int THREADS = NF * FT;
  static constexpr int NPHASES = 9;
  static constexpr int MIN_BLOCKS = (RT > 16) ? 1 : ((THREADS <= 96) ? 2 * SMO_X_MB : SMO_X_MB);
  static constexpr int XLEN = (F::XP > FS::XP) ? ((F::XP > M) ? F::XP : M) : ((FS::XP > M) ? FS::XP : M);
  static constexpr int XLP = XLEN + ((12 - XLEN % 8) % 8);   // pitch of one FFT's exchange region, = 4 (mod 8) 16-byte units
  static constexpr bool WSYNC = (FT == 32) || PAIRWARP;
  SMO_HD static constexpr int sync_after(int ph) { return (!WSYNC || ph == 3 || ph == 4) ? 2 : 1; }
  // The cp.async targets (spectral tile, velocity tile) are dense 128-byte lines with an XOR swizzle inside each
  // line instead of padded pitches: LDGSTS writes a line in one wavefront only when 8 lanes cover one ALIGNED line,
  // and the swizzle keeps the reads of the transform phases conflict free.
  static constexpr int SIN_ELEMS = NF * NH * T;      // cplx (16-byte units)
  static constexpr int SU_UNITS = 3 * M * 2;         // 16-byte units of the 4-column velocity block
  static constexpr int X_ELEMS = NJ * XLP;
  // Adjoint: the gradient integrand  nu = -dt * sum_m P_k[ F((curl G^m) x B_f^m) ]  is linear in the products, so the sum
  // over the time steps is taken on the x-spectra (this kernel adds its (curl G) x B_f outputs to a running array) and
  // the y / z transforms, the transpose and the projection are applied ONCE after the sweep instead of every step.
  // The running array's tile is streamed in with cp.async while the tile is transformed.
  static constexpr int ACC_ELEMS = (MODE == X_ADJ && !GACC) ? 3 * NH * T : 0;
  static constexpr size_t SMEM = (size_t)(SIN_ELEMS + SU_UNITS + X_ELEMS + ACC_ELEMS + 1) * sizeof(cplx);   // + the mbarrier of the velocity tile
  static_assert(R1 >= R2 && RT == R1, "the radix-R1 stage is the wide one");
  static_assert(SIN_ELEMS % 8 == 0 && SU_UNITS % 8 == 0, "cp.async regions must be whole 128-byte lines");
  static_assert(MODE == X_FWD || MODE == X_ADJ, "fused modes only");
  static_assert(!GACC || MODE == X_ADJ, "grid accumulation belongs to the adjoint pass");
  // does field f take part in the forward (r2c) half of the pass?
  SMO_HD static bool fwd_half(int f) { return !(GACC && f >= 3); }
  static_assert(NH % 2 == 0, "two spectral rows per 128-byte line need an even number of retained modes");
  struct State {
    double re[RT], im[RT];
    double wr, wi;     // w_M^jj = exp(-2 pi i jj / M): base of this thread's inter-stage twiddles (wide mapping)
    double jacc;       // INTEG forward: running sum of this thread's |B|^2
    int it;
  };
  static constexpr bool HAS_FINISH = INTEG && MODE == X_FWD;
  // fixed-order sum of the threads' partial sums -> jpart[cta]
  template <int STEP> SMO_HD static void finish(const Params& p, const Ctx& c, State& st) {
    double* R = reinterpret_cast<double*>(x_buf(c.smem));
    if (STEP == 0) { R[c.tid] = st.jacc; return; }
    if (c.tid == 0) {
      double s = 0.0;
      for (int t = 0; t < THREADS; ++t) s += R[t];
      p.jpart[c.cta] = s;
    }
  }

  SMO_HD static cplx* sin_buf(unsigned char* s) { return reinterpret_cast<cplx*>(s); }
  SMO_HD static cplx* su_buf(unsigned char* s) { return sin_buf(s) + SIN_ELEMS; }
  SMO_HD static cplx* x_buf(unsigned char* s) { return su_buf(s) + SU_UNITS; }
  SMO_HD static cplx* acc_buf(unsigned char* s) { return x_buf(s) + X_ELEMS; }
  SMO_HD static unsigned long long* ubar(unsigned char* s) { return reinterpret_cast<unsigned long long*>(acc_buf(s) + ACC_ELEMS); }
  SMO_HD static unsigned long long* sbar(unsigned char* s) { return ubar(s) + 1; }     // mbarrier of the spectral tiles (TMA path)
  static constexpr bool TMA_OK = !PAIRWARP && (NH * T * (int)sizeof(cplx)) % 1024 == 0;   // every field's tile starts on a swizzle period
  static constexpr int TMA_BOX_COLS = T, TMA_BOX_ROWS = NH;                              // box of one tensor copy
  // unit index of spectral entry (field f, row, tile column col): two rows per 128-byte line
  SMO_HD static int si(int f, int row, int col) {
    return (((f * NH + row) >> 1) << 3) + ((((row & 1) << 2) + col) ^ ((row >> 1) & 3));
  }
  // unit index of the velocity pair (component cidx, grid row n, column pair pp)
  SMO_HD static int ui(int cidx, int n, int pp) { return ((cidx * M + n) * 2 + pp) ^ ((n >> 2) & 1); }
  // field whose spectrum the threads of input field f produce (the cross-product component containing their own values)
  SMO_HD static int out_field(int f) {
    if (MODE == X_FWD) return (f + 1) % 3;
    return f < 3 ? (f + 2) % 3 : 3 + (f - 2) % 3;
  }
  SMO_HD static long long tile_of(const Params& p, int work) {
    return (long long)(work / p.tiles_per_row) * p.row_tiles + p.tile0 + (work % p.tiles_per_row);
  }
  // asynchronous loads of the spectral tile (every field by its own threads) / the velocity tile of `work`
  SMO_HD static void load_sin(const Params& p, int work, const Ctx& c) {
    cplx* S = sin_buf(c.smem);
    const long long col0 = tile_of(p, work) * T;
    const int f = c.tid / FT;
    if (TMA_OK && p.tma_sin) {
      // one TMA tensor copy per field, issued by the first lane of the field's warp (the warp has just finished reading its tile:
      // no other warp touches it); SWIZZLE_64B of the descriptor == si()
      if (c.tid % FT == 0) {
        mbar_expect_tx(sbar(c.smem), (unsigned)(NH * T * sizeof(cplx)));
        tma_load_3d(&S[f * NH * T], &p.tm[f >= 3 ? 1 : 0], (int)(col0 * 2), 0, p.tz[f], sbar(c.smem));
      }
      return;
    }
    if constexpr (PAIRWARP) {   // the warp of pair pp streams in the two columns it assembles
      const int pp = (c.tid % FT) / LP;
      for (int q = c.tid % LP; q < NH * 2; q += LP) {
        const int tc = 2 * pp + (q & 1), row = q >> 1;
        cp_async16(&S[si(f, row, tc)], p.sin[f] + (long long)row * p.ncols + col0 + tc);
      }
    } else {
      for (int q = c.tid % FT; q < NH * T; q += FT) {
        const int tc = q % T, row = q / T;
        cp_async16(&S[si(f, row, tc)], p.sin[f] + (long long)row * p.ncols + col0 + tc);
      }
    }
  }
  // running-sum tile of the field this warp produces (fields 3..5 only), consumed by phase 8 of the same work item
  SMO_HD static void load_acc(const Params& p, int work, const Ctx& c) {
    const int f = c.tid / FT;
    if (MODE != X_ADJ || GACC || f < 3) return;
    cplx* A = acc_buf(c.smem);
    const long long col0 = tile_of(p, work) * T;
    const cplx* src = p.sout[out_field(f)];
    if constexpr (PAIRWARP) {
      const int pp = (c.tid % FT) / LP;
      for (int q = c.tid % LP; q < NH * 2; q += LP) {
        const int tc = 2 * pp + (q & 1), row = q >> 1;
        cp_async16(&A[si(f - 3, row, tc)], src + (long long)row * p.ncols + col0 + tc);
      }
    } else {
      for (int q = c.tid % FT; q < NH * T; q += FT) {
        const int tc = q % T, row = q / T;
        cp_async16(&A[si(f - 3, row, tc)], src + (long long)row * p.ncols + col0 + tc);
      }
    }
  }
  // the velocity tile is stored in HBM in the (swizzled) order it has in shared memory (UTile below), so it is one contiguous
  // copy: a single TMA bulk copy issued by one thread (bulk_u), or 16-byte cp.async by everybody
  SMO_HD static void load_su(const Params& p, int work, const Ctx& c) {
    cplx* U = su_buf(c.smem);
    const double* src = p.ut + tile_of(p, work) * (3LL * M * 4);
    if (p.bulk_u) {
      if (c.tid == 0) bulk_load(U, src, (unsigned)(SU_UNITS * sizeof(cplx)), ubar(c.smem));
    } else {
      for (int q = c.tid; q < SU_UNITS; q += THREADS) cp_async16(&U[q], src + 2 * q);
    }
  }

  SMO_HD static void init(const Params& p, const Ctx& c, State& st) {
    if (c.tid == 0) { mbar_init(ubar(c.smem), 1); mbar_init(sbar(c.smem), NF); }
    const cplx w = ldg_c(p.tw + (((c.tid % FT) % LP) < RT ? ((c.tid % FT) % LP) : 0));
    st.wr = w.x; st.wi = w.y;
    st.jacc = 0.0;
    st.it = 0;
  }

  template <int PH>
  SMO_HD static void phase2(const Params& p, int work, int /*step*/, const Ctx& c, State& st) {
    cplx* S = sin_buf(c.smem);
    const cplx* U = su_buf(c.smem);
    const int f = c.tid / FT, tif = c.tid % FT;
    // wide mapping (R1 = RT threads per FFT) and narrow mapping (R2 threads per FFT, dense on the first lanes)
    const int ppw = tif / LP, jw = tif % LP;
    const bool wact = !PAIRWARP || jw < RT;
    const int ppn = PAIRWARP ? ppw : tif / R2, jn = PAIRWARP ? jw : tif % R2;
    const bool nact = PAIRWARP ? (jw < R2) : (tif < NARROW);
    cplx* Xw = x_buf(c.smem) + (f * HP + ppw) * XLP;
    cplx* Xn = x_buf(c.smem) + (f * HP + (nact ? ppn : 0)) * XLP;
    const bool more = work + c.ncta < p.nwork;
    if (PH == 0) {
      if (st.it == 0) {
        load_sin(p, work, c); cp_async_commit();
        load_su(p, work, c); cp_async_commit();
      }
      cp_async_wait<1>();                       // the spectral tile of this work item has landed
      if (TMA_OK && p.tma_sin) mbar_wait(sbar(c.smem), (unsigned)(st.it & 1));
    }
    if (PH == 1) {
      if (nact) {
        const int j = jn;
#pragma unroll
        for (int i = 0; i < R1; ++i) {
          const int n = j + R2 * i;
          double zr = 0.0, zi = 0.0;
          if (n < NH) {
            const cplx v = S[si(f, n, 2 * ppn)], w = S[si(f, n, 2 * ppn + 1)];
            if (n == 0) { zr = v.x; zi = w.x; } else { zr = v.x - w.y; zi = v.y + w.x; }
          } else if (n > M - NH) {
            const int m = M - n;
            const cplx v = S[si(f, m, 2 * ppn)], w = S[si(f, m, 2 * ppn + 1)];
            zr = v.x + w.y; zi = w.x - v.y;
          }
          st.re[i] = zr; st.im[i] = zi;
        }
        RegFFT<R1, +1>::run(as_arr<R1>(st.re), as_arr<R1>(st.im));
        const cplx wj = ldg_c(p.tw + j);
        twiddle_powers<R1>(st.re, st.im, wj.x, -wj.y);        // conj: inverse direction
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) Xn[j * F::SK + k1] = make_double2(st.re[k1], st.im[k1]);
      }
    }
    if (PH == 2) {
      if (more) load_sin(p, work + c.ncta, c);   // the spectral buffer was consumed in phase 1
      if (MODE == X_ADJ && !GACC && p.accumulate) load_acc(p, work, c);
      cp_async_commit();
      if (GACC && f >= 3 && wact) {              // the running-sum rows this thread updates in phase 4: on their way into L2
        const cplx* ga = reinterpret_cast<const cplx*>(p.gacc + tile_of(p, work) * (3LL * M * 4)) + ((f - 2) % 3 * M + jw) * 2 + ppw;
#pragma unroll
        for (int i = 0; i < R2; ++i) prefetch_l2(ga + R1 * i * 2);
      }
      if (wact) {
#pragma unroll
        for (int j = 0; j < R2; ++j) {
          const cplx v = Xw[j * F::SK + jw];
          st.re[j] = v.x; st.im[j] = v.y;
        }
        stage2<F, +1>(st.re, st.im);             // st[k2] = grid value of field f at row jw + R1*k2 (column pair ppw)
      }
    }
    if (PH == 3) {
      if (wact) {
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) Xw[jw + R1 * k2] = make_double2(st.re[k2], st.im[k2]);
      }
      cp_async_wait<1>();                       // the velocity tile of this work item has landed
      if (p.bulk_u) mbar_wait(ubar(c.smem), (unsigned)(st.it & 1));
    }
    if (PH == 4 && wact) {
      // component (out_field) of the cross product that contains this thread's own field; own values stay in registers
      const cplx* Xa = x_buf(c.smem) + ppw * XLP + jw;        // + field * HP * XLP + row
#pragma unroll
      for (int i = 0; i < R2; ++i) {
        const int n = jw + R1 * i;
        const double ox = st.re[i], oy = st.im[i];
        double e0, e1;
        if (MODE == X_FWD) {
          // E_c = U_c1 B_c2 - U_c2 B_c1 with c2 = f (own), c = f+1, c1 = f+2
          const int c1 = (f + 2) % 3;
          const cplx u1 = U[ui(c1, n, ppw)], u2 = U[ui(f, n, ppw)];
          const cplx b1 = Xa[c1 * HP * XLP + R1 * i];
          e0 = u1.x * ox - u2.x * b1.x;
          e1 = u1.y * oy - u2.y * b1.y;
          if (INTEG) st.jacc += ox * ox + oy * oy;
        } else if (f < 3) {
          // (W x U)_c = W_c1 U_c2 - W_c2 U_c1 with c1 = f (own), c = f+2, c2 = f+1
          const int c2 = (f + 1) % 3;
          const cplx u2 = U[ui(c2, n, ppw)], u1 = U[ui(f, n, ppw)];
          const cplx w2 = Xa[c2 * HP * XLP + R1 * i];
          e0 = ox * u2.x - w2.x * u1.x;
          e1 = oy * u2.y - w2.y * u1.y;
          if (INTEG) {   // source -2 B_f of the G equation (KD:862-864), component c = f+2
            const cplx bc = Xa[(3 + (f + 2) % 3) * HP * XLP + R1 * i];
            e0 -= 2.0 * bc.x; e1 -= 2.0 * bc.y;
          }
        } else {
          // (W x B)_c = W_c1 B_c2 - W_c2 B_c1 with c2 = f-3 (own), c = c2+1, c1 = c2+2
          const int g = f - 3, c1 = (g + 2) % 3;
          const cplx w1 = Xa[c1 * HP * XLP + R1 * i], w2 = Xa[g * HP * XLP + R1 * i];
          const cplx b1 = Xa[(3 + c1) * HP * XLP + R1 * i];
          e0 = w1.x * ox - w2.x * b1.x;
          e1 = w1.y * oy - w2.y * b1.y;
        }
        st.re[i] = e0; st.im[i] = e1;
      }
      if (fwd_half(f)) {
        RegFFT<R2, -1>::run(as_arr<R2>(st.re), as_arr<R2>(st.im));
        twiddle_powers<R2>(st.re, st.im, st.wr, st.wi);
      }
    }
    if (PH == 5) {
      if (more) load_su(p, work + c.ncta, c);    // the velocity buffer was consumed in phase 4
      cp_async_commit();
      if (wact && fwd_half(f)) {
#pragma unroll
        for (int k1 = 0; k1 < R2; ++k1) Xw[jw * FS::SK + k1] = make_double2(st.re[k1], st.im[k1]);
      }
      if (GACC && wact && !fwd_half(f)) {
        // the (curl G) x B_f products of phase 4 (still in registers) join the running sum on the grid: component c = f-2 (mod 3),
        // rows jw + R1*i, column pair ppw - 512 contiguous bytes per warp and i.  All loads are issued before the first add, and
        // nothing waits for this warp until the next tile's product phase: the round trip hides under the other warps' r2c half.
        cplx* ga = reinterpret_cast<cplx*>(p.gacc + tile_of(p, work) * (3LL * M * 4)) + (((f - 2) % 3) * M + jw) * 2 + ppw;
        cplx a[R2];
#pragma unroll
        for (int i = 0; i < R2; ++i) a[i] = ga[R1 * i * 2];
#pragma unroll
        for (int i = 0; i < R2; ++i) ga[R1 * i * 2] = make_double2(a[i].x + st.re[i], a[i].y + st.im[i]);
      }
    }
    if (PH == 6) {
      if (nact && fwd_half(f)) {
#pragma unroll
        for (int j = 0; j < R1; ++j) {
          const cplx v = Xn[j * FS::SK + jn];
          st.re[j] = v.x; st.im[j] = v.y;
        }
        stage2<FS, -1>(st.re, st.im);
      }
    }
    if (PH == 7) {
      if (nact && fwd_half(f)) {
#pragma unroll
        for (int k2 = 0; k2 < R1; ++k2) {
          const int k = jn + R2 * k2;   // only the retained modes and their mirror images are needed by phase 8
          if (k < NH || k > M - NH) Xn[k] = make_double2(st.re[k2], st.im[k2]);
        }
      }
      if (MODE == X_ADJ && !GACC) cp_async_wait<1>();     // the running-sum tile (committed in phase 2) has landed
    }
    if (PH == 8 && fwd_half(f)) {
      // own thread order (column pairs fastest) so that a row's T columns are stored by adjacent lanes
      const int pp8 = PAIRWARP ? tif / LP : tif % HP, kk = PAIRWARP ? tif % LP : tif / HP;
      const cplx* X8 = x_buf(c.smem) + (f * HP + pp8) * XLP;
      cplx* O = p.sout[out_field(f)] + tile_of(p, work) * T + 2 * pp8;
      const double h = 0.5 * p.scale;
      const bool addto = (MODE == X_ADJ) && !GACC && p.accumulate && f >= 3;
      const cplx* A = acc_buf(c.smem);
      for (int k = kk; k < NH; k += LP) {
        const cplx zk = X8[k];
        const cplx zm = X8[(M - k) % M];
        cplx o0 = make_double2(h * (zk.x + zm.x), h * (zk.y - zm.y));
        cplx o1 = make_double2(h * (zk.y + zm.y), h * (zm.x - zk.x));
        if (addto) {
          const cplx a0 = A[si(f - 3, k, 2 * pp8)], a1 = A[si(f - 3, k, 2 * pp8 + 1)];
          o0.x += a0.x; o0.y += a0.y; o1.x += a1.x; o1.y += a1.y;
        }
        O[(long long)k * p.ncols] = o0;
        O[(long long)k * p.ncols + 1] = o1;
      }
    }
    if (PH == 8) st.it++;
  }
};

// one-off re-layout of the velocity field: grid [3][M][ncols] -> tile-major [ncols/4][3][M][4]
struct UTileParams {
  const double* in[3];
  double* out;
  int nwork, nsteps;
  long long ncols;
  int M;
  int half;      // 0: pair layout of XFused; T (2 or 4): the half-length layout of xpass_half.cuh ([ncols/T][3][M/2][T][2])
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
    // unit (16 bytes) = (component, row, column pair), stored at its swizzled shared-memory position XFused::ui()
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
      const long long unit = (((long long)cc * p.M + n) * 2 + (col % 4) / 2) ^ ((n >> 2) & 1);
      p.out[(col / 4) * (3LL * p.M * 4) + unit * 2 + (col & 1)] = p.in[cc][(long long)n * p.ncols + col];
    }
  }
};

// inverse re-layout (once per adjoint sweep): tile-major running sum [ncols/4][3][M][4] -> grid [3][M][ncols]
struct UnTile {
  typedef UTileParams Params;    // in[c] = destination grids (written), out = tile-major source (read)
  static constexpr int THREADS = 256;
  static constexpr int NPHASES = 1;
  static constexpr int MIN_BLOCKS = 1;
  static constexpr size_t SMEM = 0;
  struct State {};
  template <int PH> SMO_HD static void phase(const Params& p, int work, int, int tid, unsigned char*, State&) {
    const long long per_row = (p.ncols + THREADS - 1) / THREADS;
    const int n = (int)(work / per_row);
    const long long col = (work % per_row) * THREADS + tid;
    if (col >= p.ncols) return;
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
      const long long src = p.half ? ((((col / p.half) * 3 + cc) * (p.M / 2) + n / 2) * p.half + (col % p.half)) * 2 + (n & 1)
                                   : (((col / 4) * 3 + cc) * p.M + n) * 4 + (col % 4);
      const_cast<double*>(p.in[cc])[(long long)n * p.ncols + col] = p.out[src];
    }
  }
};

}  // namespace smo
