// Fused x pass of the dynamo time loops for grids whose dealiased length M has no 16-thread two-stage factorisation
// (256^3: M = 384 = 24 x 16) - "half-length" variant.
//
// Same mathematics and the same phase structure as XFused (xpass.cuh): c2r of every operand, U x B | (curl G) x U and
// (curl G) x B_f on the dealiased grid (FWD_Solve_KDyn.py:417-419, 846-848, 857-859, 875-877), r2c + truncation, the real
// grid never touching HBM.  The difference is how a real column becomes a complex FFT.  XFused packs TWO real columns into
// one complex transform of length M; for M = 384 that transform needs 24 stage threads per line (24 x 16), which neither
// tiles a warp nor leaves room for two CTAs per SM (48 doubles of state per thread).  Here ONE real column of length M is
// ONE complex transform of length H = M/2 = 192 = 16 x 12 - the factorisation the 128^3 kernels are tuned for - using the
// even/odd identity (the one sh23.cuh uses for the SH23 time loop):
//     c2r:  Z[k] = (c[k] + conj(c[H-k])) + i e^{+2 pi i k/M} (c[k] - conj(c[H-k])),  z = IDFT_H(Z),  u[2n] + i u[2n+1] = z[n]
//     r2c:  Z = DFT_H(u[2n] + i u[2n+1]),  c[k] = ( (Z[k] + conj(Z[H-k])) - i e^{-2 pi i k/M} (Z[k] - conj(Z[H-k])) ) / 2
// (c[k] = 0 for k >= NH = M/3 retained modes; Im c[0] dropped as FFTW's c2r does).  A thread therefore holds the grid values of
// two consecutive x rows of one column as (re, im); the pointwise products act on both, exactly as on a column pair.
//
// Layout differences to XFused: a CTA handles T = 4 columns = 4 complex FFTs per field (FT = 4 x 16 = 64 threads = 2 warps per
// field; every FFT, its spectral tile columns and its spectrum hand-over stay inside one warp); the velocity tile is
// [ncols/4][3][H][4][2] (the two x rows of a column adjacent: one 16-byte unit per (row pair, column)).
#pragma once
#include "xpass.cuh"

namespace smo {

template <class F, int MODE, bool INTEG = false, bool GACC = false, int T_ = 4> struct XFusedH {
  typedef XFParams Params;
  typedef typename F::Swapped FS;
  static constexpr bool V2 = true;
  static constexpr int T = T_;                        // columns (= complex FFTs per field) per CTA
  static constexpr int NF = (MODE == X_FWD) ? 3 : 6;  // fields in = fields out
  static constexpr int R1 = F::R1, R2 = F::R2, H = F::M, RT = F::RT;
  static constexpr int M = 2 * H;                     // real grid length
  static constexpr int NH = M / 3;                    // retained kx modes (dealias 3/2)
  static constexpr int NJ = NF * T;
  static constexpr int FT = T * RT;                   // threads per field
  static constexpr int WT = (FT < 32) ? FT : 32;      // threads of a field inside one warp
  static constexpr int CPW = WT / RT;                 // columns per warp
  static constexpr int NARROW = CPW * R2;             // active lanes per warp in the R2-thread stages (dense mapping)
  static constexpr int THREADS = NF * FT;
  static constexpr int NPHASES = 9;
  static constexpr int MIN_BLOCKS = (THREADS <= 96) ? 4 : ((THREADS <= 192) ? 2 : 1);
  static constexpr int XLEN = (F::XP > FS::XP) ? ((F::XP > H) ? F::XP : H) : ((FS::XP > H) ? FS::XP : H);
  static constexpr int XLP = XLEN + ((12 - XLEN % 8) % 8);   // pitch of one FFT's exchange region, = 4 (mod 8) 16-byte units
  SMO_HD static constexpr int sync_after(int ph) { return (ph == 3 || ph == 4) ? 2 : 1; }
  static constexpr int SIN_ELEMS = NF * NH * T;      // cplx (16-byte units)
  static constexpr int SU_UNITS = 3 * H * T;         // 16-byte units of the T-column velocity block
  static constexpr int X_ELEMS = NJ * XLP;
  static constexpr int ACC_ELEMS = (MODE == X_ADJ && !GACC) ? 3 * NH * T : 0;   // GACC: running sum on the grid (see XFused)
  static constexpr size_t SMEM = (size_t)(SIN_ELEMS + SU_UNITS + X_ELEMS + ACC_ELEMS + 1) * sizeof(cplx);   // + the mbarrier of the velocity tile
  static_assert(R1 >= R2 && RT == R1, "the radix-R1 stage is the wide one");
  static_assert(32 % RT == 0 && FT % WT == 0 && T_ % CPW == 0, "whole FFTs per warp");
  static_assert(T_ == 2 || T_ == 4, "the shared-memory swizzles assume 2 or 4 columns (32 / 64 bytes) per row");
  static_assert(NH % 8 == 0 && H % 8 == 0 && (CPW == 2 || CPW == 4), "swizzled lines hold whole rows; blocks must start on a swizzle period");
  static_assert(NH <= H && H - NH < NH, "mode bookkeeping of the even/odd assembly");
  static_assert(MODE == X_FWD || MODE == X_ADJ, "fused modes only");
  static_assert(!GACC || MODE == X_ADJ, "grid accumulation belongs to the adjoint pass");
  SMO_HD static bool fwd_half(int f) { return !(GACC && f >= 3); }
  struct State {
    double re[RT], im[RT];
    double wr, wi;     // exp(-2 pi i jw / H): base of this thread's inter-stage twiddles of the forward transform (wide mapping)
    double jacc;       // INTEG forward: running sum of this thread's |B|^2
    int it;
  };
  static constexpr bool HAS_FINISH = INTEG && MODE == X_FWD;
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
  static constexpr bool TMA_OK = (NH * CPW * (int)sizeof(cplx)) % 1024 == 0;           // every warp's block starts on a swizzle period
  static constexpr int TMA_BOX_COLS = CPW, TMA_BOX_ROWS = NH;                          // box of one tensor copy
  // Spectral / running-sum tiles: every warp owns the CPW columns of its FFTs, stored as a block [rows][CPW columns] of its own,
  // so that the warp's cp.async of 8 consecutive lanes (LROWS rows x CPW columns) fills one ALIGNED 128-byte line in a single
  // wavefront (r2d ncu: with a row-major [row][4 columns] tile a warp touched half of every line - 11 wavefronts per LDGSTS
  // instead of 4, a quarter of all shared-memory wavefronts of the kernel).  Inside a line the position is XOR-swizzled by the
  // line number so that 8 lanes reading one column of 8 consecutive rows hit 8 different 16-byte bank groups.
  static constexpr int LROWS = 8 / CPW;              // rows per 128-byte line of a warp's block
  SMO_HD static int si(int f, int row, int col) {
    const int g = col / CPW, cc = col % CPW;
    const int r = (f * (T / CPW) + g) * NH + row;    // row inside the stack of blocks (NH is a multiple of LROWS)
    return ((r / LROWS) << 3) + (((row % LROWS) * CPW + cc) ^ ((row / LROWS) % CPW));
  }
  // velocity tile [3][H row pairs][T columns]: 8 / T rows per 128-byte line, same XOR swizzle idea
  static constexpr int LROWS_U = 8 / T;
  SMO_HD static int ui(int cidx, int n, int col) {
    const int r = cidx * H + n;                      // (H is a multiple of LROWS_U)
    return ((r / LROWS_U) << 3) + (((n % LROWS_U) * T + col) ^ ((n / LROWS_U) % T));
  }
  SMO_HD static int out_field(int f) {
    if (MODE == X_FWD) return (f + 1) % 3;
    return f < 3 ? (f + 2) % 3 : 3 + (f - 2) % 3;
  }
  SMO_HD static long long tile_of(const Params& p, int work) {
    return (long long)(work / p.tiles_per_row) * p.row_tiles + p.tile0 + (work % p.tiles_per_row);
  }
  // every warp streams in the CPW columns (CPW * 16 contiguous bytes per row) its own FFTs assemble
  SMO_HD static void load_sin(const Params& p, int work, const Ctx& c) {
    cplx* S = sin_buf(c.smem);
    const long long col0 = tile_of(p, work) * T;
    const int f = c.tid / FT, tif = c.tid % FT, wv = tif / WT, lane = tif % WT;
    if (TMA_OK && p.tma_sin) {
      // one TMA tensor copy per warp: box = its CPW columns x NH rows; SWIZZLE_32B (CPW = 2) / SWIZZLE_64B (CPW = 4) == si()
      if (lane == 0) {
        mbar_expect_tx(sbar(c.smem), (unsigned)(NH * CPW * sizeof(cplx)));
        tma_load_3d(&S[(f * (T / CPW) + wv) * NH * CPW], &p.tm[f >= 3 ? 1 : 0], (int)((col0 + wv * CPW) * 2), 0, p.tz[f], sbar(c.smem));
      }
      return;
    }
    const cplx* src = p.sin[f] + col0 + wv * CPW;
    for (int q = lane; q < NH * CPW; q += WT) {
      const int cc = q % CPW, row = q / CPW;
      cp_async16(&S[si(f, row, wv * CPW + cc)], src + (long long)row * p.ncols + cc);
    }
  }
  SMO_HD static void load_acc(const Params& p, int work, const Ctx& c) {
    const int f = c.tid / FT;
    if (MODE != X_ADJ || GACC || f < 3) return;
    cplx* A = acc_buf(c.smem);
    const long long col0 = tile_of(p, work) * T;
    const int tif = c.tid % FT, wv = tif / WT, lane = tif % WT;
    const cplx* src = p.sout[out_field(f)] + col0 + wv * CPW;
    for (int q = lane; q < NH * CPW; q += WT) {
      const int cc = q % CPW, row = q / CPW;
      cp_async16(&A[si(f - 3, row, wv * CPW + cc)], src + (long long)row * p.ncols + cc);
    }
  }
  // the velocity tile is stored in HBM in its (swizzled) shared-memory order (UTileH): one contiguous copy
  SMO_HD static void load_su(const Params& p, int work, const Ctx& c) {
    cplx* U = su_buf(c.smem);
    const double* src = p.ut + tile_of(p, work) * (3LL * M * T);
    if (p.bulk_u) {
      if (c.tid == 0) bulk_load(U, src, (unsigned)(SU_UNITS * sizeof(cplx)), ubar(c.smem));
    } else {
      for (int q = c.tid; q < SU_UNITS; q += THREADS) cp_async16(&U[q], src + 2 * q);
    }
  }

  SMO_HD static void init(const Params& p, const Ctx& c, State& st) {
    if (c.tid == 0) { mbar_init(ubar(c.smem), 1); mbar_init(sbar(c.smem), NF * (T / CPW)); }
    const int jw = (c.tid % FT) % RT;
    const cplx w = ldg_c(p.tw + 2 * jw);            // p.tw[m] = exp(-2 pi i m / M);  exp(-2 pi i jw / H) = tw[2 jw]
    st.wr = w.x; st.wi = w.y;
    st.jacc = 0.0;
    st.it = 0;
  }

  template <int PH>
  SMO_HD static void phase2(const Params& p, int work, int /*step*/, const Ctx& c, State& st) {
    cplx* S = sin_buf(c.smem);
    const cplx* U = su_buf(c.smem);
    const int f = c.tid / FT, tif = c.tid % FT, wv = tif / WT, lane = tif % WT;
    // wide mapping (R1 = RT threads per FFT) and narrow mapping (R2 threads per FFT, dense on the first lanes of the warp)
    const int cw = wv * CPW + lane / RT, jw = lane % RT;
    const bool nact = lane < NARROW;
    const int cn = wv * CPW + (nact ? lane / R2 : 0), jn = lane % R2;
    cplx* Xw = x_buf(c.smem) + (f * T + cw) * XLP;
    cplx* Xn = x_buf(c.smem) + (f * T + cn) * XLP;
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
        // twiddles e^{+2 pi i n/M} for n = j + R2*i by recurrence: cur = w^j, step = w^R2
        const cplx w0 = ldg_c(p.tw + j), ws = ldg_c(p.tw + R2);
        double cr = w0.x, ci = -w0.y;
        const double sr = ws.x, si_ = -ws.y;
#pragma unroll
        for (int i = 0; i < R1; ++i) {
          const int n = j + R2 * i;
          double ax = 0.0, ay = 0.0, bx = 0.0, by = 0.0;    // a = c[n], b = conj(c[H-n])
          if (n < NH) { const cplx v = S[si(f, n, cn)]; ax = v.x; ay = (n == 0) ? 0.0 : v.y; }
          if (n > H - NH) { const cplx v = S[si(f, H - n, cn)]; bx = v.x; by = -v.y; }
          const double sx = ax + bx, sy = ay + by, dx = ax - bx, dy = ay - by;
          // Z = s + i w^n d
          st.re[i] = sx - (ci * dx + cr * dy);
          st.im[i] = sy + (cr * dx - ci * dy);
          const double t = cr * sr - ci * si_; ci = cr * si_ + ci * sr; cr = t;
        }
        RegFFT<R1, +1>::run(as_arr<R1>(st.re), as_arr<R1>(st.im));
        const cplx wj = ldg_c(p.tw + 2 * j);
        twiddle_powers<R1>(st.re, st.im, wj.x, -wj.y);        // conj: inverse direction, exp(+2 pi i j k1 / H)
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) Xn[j * F::SK + k1] = make_double2(st.re[k1], st.im[k1]);
      }
    }
    if (PH == 2) {
      if (more) load_sin(p, work + c.ncta, c);   // the spectral buffer was consumed in phase 1
      if (MODE == X_ADJ && !GACC && p.accumulate) load_acc(p, work, c);
      cp_async_commit();
      if (GACC && f >= 3) {                      // the running-sum rows this thread updates in phase 4: on their way into L2
        const cplx* ga = reinterpret_cast<const cplx*>(p.gacc + tile_of(p, work) * (3LL * M * T)) + (((f - 2) % 3) * H + jw) * T + cw;
#pragma unroll
        for (int i = 0; i < R2; ++i) prefetch_l2(ga + R1 * i * T);
      }
#pragma unroll
      for (int j = 0; j < R2; ++j) {
        const cplx v = Xw[j * F::SK + jw];
        st.re[j] = v.x; st.im[j] = v.y;
      }
      stage2<F, +1>(st.re, st.im);               // st[k2] = (u[2n], u[2n+1]) of field f, column cw, n = jw + R1*k2
    }
    if (PH == 3) {
#pragma unroll
      for (int k2 = 0; k2 < R2; ++k2) Xw[jw + R1 * k2] = make_double2(st.re[k2], st.im[k2]);
      cp_async_wait<1>();                       // the velocity tile of this work item has landed
      if (p.bulk_u) mbar_wait(ubar(c.smem), (unsigned)(st.it & 1));
    }
    if (PH == 4) {
      const cplx* Xa = x_buf(c.smem) + cw * XLP + jw;        // + field * T * XLP + row pair
#pragma unroll
      for (int i = 0; i < R2; ++i) {
        const int n = jw + R1 * i;
        const double ox = st.re[i], oy = st.im[i];
        double e0, e1;
        if (MODE == X_FWD) {
          const int c1 = (f + 2) % 3;
          const cplx u1 = U[ui(c1, n, cw)], u2 = U[ui(f, n, cw)];
          const cplx b1 = Xa[c1 * T * XLP + R1 * i];
          e0 = u1.x * ox - u2.x * b1.x;
          e1 = u1.y * oy - u2.y * b1.y;
          if (INTEG) st.jacc += ox * ox + oy * oy;
        } else if (f < 3) {
          const int c2 = (f + 1) % 3;
          const cplx u2 = U[ui(c2, n, cw)], u1 = U[ui(f, n, cw)];
          const cplx w2 = Xa[c2 * T * XLP + R1 * i];
          e0 = ox * u2.x - w2.x * u1.x;
          e1 = oy * u2.y - w2.y * u1.y;
          if (INTEG) {   // source -2 B_f of the G equation (KD:862-864), component c = f+2
            const cplx bc = Xa[(3 + (f + 2) % 3) * T * XLP + R1 * i];
            e0 -= 2.0 * bc.x; e1 -= 2.0 * bc.y;
          }
        } else {
          const int g = f - 3, c1 = (g + 2) % 3;
          const cplx w1 = Xa[c1 * T * XLP + R1 * i], w2 = Xa[g * T * XLP + R1 * i];
          const cplx b1 = Xa[(3 + c1) * T * XLP + R1 * i];
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
      if (fwd_half(f)) {
#pragma unroll
        for (int k1 = 0; k1 < R2; ++k1) Xw[jw * FS::SK + k1] = make_double2(st.re[k1], st.im[k1]);
      } else if (GACC) {
        // the (curl G) x B_f products of phase 4 (still in registers) join the running sum on the grid (see XFused)
        cplx* ga = reinterpret_cast<cplx*>(p.gacc + tile_of(p, work) * (3LL * M * T)) + (((f - 2) % 3) * H + jw) * T + cw;
        cplx a[R2];
#pragma unroll
        for (int i = 0; i < R2; ++i) a[i] = ga[R1 * i * T];
#pragma unroll
        for (int i = 0; i < R2; ++i) ga[R1 * i * T] = make_double2(a[i].x + st.re[i], a[i].y + st.im[i]);
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
        for (int k2 = 0; k2 < R1; ++k2) Xn[jn + R2 * k2] = make_double2(st.re[k2], st.im[k2]);   // Z[k], all k (the split needs Z[H-k])
      }
      if (MODE == X_ADJ && !GACC) cp_async_wait<1>();     // the running-sum tile (committed in phase 2) has landed
    }
    if (PH == 8 && fwd_half(f)) {
      // own thread order (columns fastest) so that the warp's CPW columns of a row are stored by adjacent lanes
      const int c8 = wv * CPW + lane % CPW, kk = lane / CPW;
      const cplx* X8 = x_buf(c.smem) + (f * T + c8) * XLP;
      cplx* O = p.sout[out_field(f)] + tile_of(p, work) * T + c8;
      const double h = 0.5 * p.scale;
      const bool addto = (MODE == X_ADJ) && !GACC && p.accumulate && f >= 3;
      const cplx* A = acc_buf(c.smem);
      // e^{-2 pi i k/M} for k = kk + RT*t by recurrence
      const cplx w0 = ldg_c(p.tw + kk), ws = ldg_c(p.tw + RT);
      double cr = w0.x, ci = w0.y;
#pragma unroll 1
      for (int k = kk; k < NH; k += RT) {
        const cplx zk = X8[k];
        const cplx zm = X8[(H - k) % H];
        // s = Zk + conj(Zm), d = Zk - conj(Zm);  c = h * (s - i w d)
        const double sx = zk.x + zm.x, sy = zk.y - zm.y, dx = zk.x - zm.x, dy = zk.y + zm.y;
        cplx o = make_double2(h * (sx + (ci * dx + cr * dy)), h * (sy - (cr * dx - ci * dy)));
        if (addto) { const cplx a = A[si(f - 3, k, c8)]; o.x += a.x; o.y += a.y; }
        O[(long long)k * p.ncols] = o;
        const double t = cr * ws.x - ci * ws.y; ci = cr * ws.y + ci * ws.x; cr = t;
      }
    }
    if (PH == 8) st.it++;
  }
};

// one-off re-layout of the velocity field for XFusedH: grid [3][M][ncols] -> [ncols/T][3][M/2][T][2], T = p.half columns per tile
struct UTileH {
  typedef UTileParams Params;
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
    // unit (16 bytes) = (component, row pair, column), stored at its swizzled shared-memory position XFusedH::ui()
    const int T = p.half, LR = 8 / T;
    const int H = p.M / 2, n2 = n / 2, cq = (int)(col % T);
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
      const long long r = (long long)cc * H + n2;
      const long long unit = ((r / LR) << 3) + (((n2 % LR) * T + cq) ^ ((n2 / LR) % T));
      p.out[(col / T) * (3LL * p.M * T) + unit * 2 + (n & 1)] = p.in[cc][(long long)n * p.ncols + col];
    }
  }
};

}  // namespace smo
