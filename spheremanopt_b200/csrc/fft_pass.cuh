// Batched complex-to-complex FFT pass along one axis of a 3-D coefficient/pencil array, with the
// Dedalus-style zero-padding (coefficient -> dealiased grid) or truncation (grid -> coefficient) fused in.
//
// Replaces, for the y and z axes, what the reference triggers implicitly through every field['g'] / ['c']
// access and solver.step (FWD_Solve_KDyn.py:635-641, 955-961): FFTW c2c along the axis plus the pad/truncate
// copy between scales=1 and scales=3/2 ([D2-2], [D2-3] of SURVEY.md section 8(c)).
//
// Structure (software-pipelined, HBM-bound):
//   * a CTA owns tiles of T lines (persistent loop over tiles); the input tile of tile i+1 is streamed into
//     shared memory with 16-byte asynchronous copies (cp.async / LDGSTS) while tile i is being transformed;
//   * each line is transformed by RT = max(R1,R2) threads in two register stages (generated radix-R codelets)
//     with one shared-memory exchange that re-uses the (already consumed) input buffer; twiddles come from a
//     shared-memory table laid out so that every access is a compile-time offset from a per-thread base;
//   * results go straight from registers to HBM; all per-element index arithmetic (padding map, truncation map,
//     multi-rank segment map) is precomputed once per CTA, so the steady state has no integer divisions.
//   TFAST = true  : adjacent lines are adjacent in memory (the y pass: FFT axis strided, lines along z);
//   TFAST = false : the FFT axis itself is contiguous (the z pass).
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
  int b0;                   // first line index along B handled by this launch (chunked launches), lines b0 .. b0+nB-1
  long long in_sA, in_sB, out_sA, out_sB;   // line (a,b) base offsets, in complex elements
  long long in_sN, out_sN;  // element stride along the FFT axis (1 when the axis is contiguous)
  // full-length side of the z pass in the multi-rank P1 layout: off(n) = (n / seglen) * blk + (n % seglen)
  int seglen;               // 0: no segmentation
  long long blk;
  int kmax;
  double scale;             // applied to outputs
  const cplx* tw;           // exp(-2 pi i m / M), m < M
  // Transposes fused into the stores (multi-GPU slab decomposition): results are written straight into the peers'
  // buffers over NVLink instead of a local buffer followed by an all-to-all.
  //   peer_mode 1 (inverse z pass): full-length index k goes to peer k / seglen, at peer_off + line*seglen + k % seglen
  //   peer_mode 2 (forward y pass): row a (= kx) goes to peer a / peer_rows, at peer_off + (a % peer_rows)*out_sA + ...
  int perm_rows;            // > 0: row order of the work items is interleaved over the perm_rows-row blocks of the peers, so
                            // that local and remote rows alternate in time instead of coming in bursts
  int peer_mode, peer_rows;
  long long peer_off;       // this rank's block inside every peer buffer (rank * blk)
  cplx* peer_out[MAXF][MAXP];
  // Transposes fused into the LOADS (pull): the input tile is streamed with cp.async straight out of the peers' buffers
  // over NVLink; all stores stay local.  Data arrives CTA by CTA, so the transfer of later tiles overlaps the transforms
  // of earlier ones even when the launch is a single wave.
  //   pull_mode 1 (forward z pass): full-length index n is read from peer n / seglen, at pull_off + line*seglen + n % seglen
  //   pull_mode 2 (inverse y pass): row a (= kx) is read from peer a / peer_rows, at pull_off + (a % peer_rows)*in_sA + ...
  int pull_mode;
  long long pull_off;
  const cplx* peer_in[MAXF][MAXP];
  XSync xs;                 // cross-GPU wait / signal fused into the launch (smo_common.cuh)
  int hint_in, hint_out;    // L2 residency hints of the loads / stores (0 none, 1 evict_first, 2 evict_last; y passes only)
};
// STAGE (forward y pass with peer_mode 2 only): the truncated result tile [NC rows][T columns] is staged in shared memory (in
// place of the consumed exchange buffer) and shipped to the row's owner with TMA bulk stores (one 16*T-byte row each, issued by
// the lanes of warp 0) instead of 16-byte stores from every thread: the remote traffic leaves the LSU path of the SM.

template <class F, int DIR, bool TFAST, int T_, bool STAGE = false> struct FftPass {
  typedef PassParams Params;
  static constexpr bool V2 = true;
  static constexpr bool PAD = DIR > 0;   // inverse direction: compact input, full-length output
  static constexpr int T = T_, M = F::M, R1 = F::R1, R2 = F::R2, RT = F::RT;
  static constexpr int KMAX = M / 3 - 1, NC = 2 * KMAX + 1;      // dealias 3/2: Npts = 2M/3, kmax = Npts/2 - 1
  static constexpr int THREADS = T_ * F::RT;
  static constexpr int NPHASES = STAGE ? 6 : 4;
  static_assert(!STAGE || (TFAST && DIR < 0), "staged push: forward y pass only");
  static constexpr int MIN_BLOCKS = (F::RT > 16) ? 2 : ((T_ * F::RT <= 64) ? 2 * SMO_PASS_MB : SMO_PASS_MB);   // RT > 16: 48+ doubles per thread
  // input tile: y pass [rows][T]; z pass [T][LENP]
  static constexpr int NIN = PAD ? NC : M;
  static constexpr int LENP = PAD ? NC + 1 : M;
  static constexpr int IN_ELEMS = TFAST ? NIN * T_ : T_ * LENP;
  static constexpr int X_ELEMS = T_ * F::XP;
  static constexpr int BUF = (IN_ELEMS > X_ELEMS) ? IN_ELEMS : X_ELEMS;
  static constexpr size_t SMEM = (size_t)(2 * BUF + M) * sizeof(cplx) + (size_t)M * sizeof(int) + 16;
  struct State {
    double re[F::RT], im[F::RT];
    int ooff[F::R2];   // output offset of the k2-th result of this thread (complex elements), -1 = dropped
    int it;            // tiles done by this CTA (buffer parity)
  };

  SMO_HD static cplx* buf(unsigned char* smem, int which) { return reinterpret_cast<cplx*>(smem) + (size_t)which * BUF; }
  SMO_HD static cplx* twid(unsigned char* smem) { return reinterpret_cast<cplx*>(smem) + 2 * (size_t)BUF; }
  SMO_HD static int* segtab(unsigned char* smem) { return reinterpret_cast<int*>(twid(smem) + M); }
  SMO_HD static unsigned long long* pols(unsigned char* smem) { return reinterpret_cast<unsigned long long*>(segtab(smem) + M); }

  SMO_HD static void split_tid(int tid, int& t, int& jj) {
    if (TFAST) { t = tid % T; jj = tid / T; } else { jj = tid % RT; t = tid / RT; }
  }
  SMO_HD static void decode(const Params& p, int work, int& f, int& a, int& bt) {
    const int per_field = p.nA * p.tilesB;
    f = work / per_field;
    const int r = work - f * per_field;
    a = r / p.tilesB;
    bt = r - a * p.tilesB;
    if (p.perm_rows > 0) { const int np = p.nA / p.perm_rows; a = (a % np) * p.perm_rows + a / np; }
  }
  SMO_HD static int xidx(int t, int e) { return TFAST ? e * T + t : t * F::XP + e; }
  // offset of full-length index n on a (possibly segmented) side
  SMO_HD static long long full_off(const Params& p, int n, long long sN) {
    if (!TFAST && p.seglen > 0) return (long long)(n / p.seglen) * p.blk + (long long)(n % p.seglen);
    return (long long)n * sN;
  }

  // stream the input tile of `work` into buffer `which` (asynchronous)
  SMO_HD static void load_tile(const Params& p, int work, int which, const Ctx& c) {
    int f, a, bt, t, jj;
    decode(p, work, f, a, bt);
    split_tid(c.tid, t, jj);
    cplx* B = buf(c.smem, which);
    const int b = bt * T + t;   // line within this launch
    if (b < p.nB) {
      const cplx* src = p.in[f] + (long long)a * p.in_sA + (long long)(p.b0 + b) * p.in_sB;
      if (TFAST && PAD && p.pull_mode == 2)
        src = p.peer_in[f][a / p.peer_rows] + p.pull_off + (long long)(a % p.peer_rows) * p.in_sA + (long long)(p.b0 + b) * p.in_sB;
      if (TFAST && p.hint_in) {
        const unsigned long long pol = pols(c.smem)[0];
        for (int row = jj; row < NIN; row += RT) cp_async16_hint(&B[row * T + t], src + (long long)row * p.in_sN, pol);
      } else if (TFAST) {
        // rows of T adjacent lines: lane t walks along z (16 B each, T*16 B contiguous per row)
        for (int row = jj; row < NIN; row += RT) cp_async16(&B[row * T + t], src + (long long)row * p.in_sN);
      } else if (PAD || p.seglen <= 0) {
        for (int e = jj; e < LENP; e += RT) cp_async16(&B[t * LENP + e], src + e);
      } else if (p.pull_mode == 1) {
        const int* so = segtab(c.smem);     // packed (peer << 24 | offset inside the segment)
        const long long line = p.pull_off + (long long)(p.b0 + b) * p.in_sB;
        for (int e = jj; e < LENP; e += RT) cp_async16(&B[t * LENP + e], p.peer_in[f][so[e] >> 24] + line + (so[e] & 0xffffff));
      } else {
        const int* so = segtab(c.smem);
        for (int e = jj; e < LENP; e += RT) cp_async16(&B[t * LENP + e], src + so[e]);
      }
    }
  }

  SMO_HD static void init(const Params& p, const Ctx& c, State& st) {
    int t, jj;
    split_tid(c.tid, t, jj);
    cplx* W = twid(c.smem);
    // twiddle table [k1][j]: w_M^(j*k1); a stage-1 thread reads W[k1*R2 + j] (compile-time offsets from W + j)
    for (int m = c.tid; m < M; m += THREADS) W[m] = ldg_c(p.tw + ((m % R2) * (m / R2)) % M);
    if (!TFAST) {
      int* so = segtab(c.smem);
      for (int m = c.tid; m < M; m += THREADS)
        so[m] = (!PAD && p.pull_mode == 1) ? (((m / p.seglen) << 24) | (m % p.seglen)) : (int)full_off(p, m, 1);
    }
    // per-thread output map of stage 2 (thread k1 = jj holds X[k1 + R1*k2])
#pragma unroll
    for (int k2 = 0; k2 < R2; ++k2) {
      const int k = jj + R1 * k2;
      int off = -1;
      if (jj < R1) {
        if (PAD && p.peer_mode == 1) off = ((k / p.seglen) << 24) | (k % p.seglen);   // (peer, offset inside the segment)
        else if (PAD) off = (int)full_off(p, k, p.out_sN);
        else { const int cidx = compact_index(k, M, KMAX); if (cidx >= 0) off = (int)((long long)cidx * p.out_sN); }
      }
      st.ooff[k2] = off;
    }
    st.it = 0;
    if (c.tid == 0) { pols(c.smem)[0] = l2_policy(p.hint_in ? p.hint_in : 1); pols(c.smem)[1] = l2_policy(p.hint_out ? p.hint_out : 1); }
  }

  template <int PH>
  SMO_HD static void phase2(const Params& p, int work, int /*step*/, const Ctx& c, State& st) {
    int t, jj;
    split_tid(c.tid, t, jj);
    const int cur = st.it & 1;
    cplx* B = buf(c.smem, cur);
    if (PH == 0) {
      // (first tile of this CTA: start its own stream;) prefetch the next tile of this CTA into the other buffer,
      // then wait for the current one
      if (STAGE) {
        // the other buffer was the staging area of the previous tile: its bulk stores must have read it before the prefetch
        // (issued in phase 1, after the barrier) may overwrite it
        if (st.it == 0) { load_tile(p, work, cur, c); cp_async_commit(); }
        if (c.tid < 32) bulk_wait_read();
        cp_async_wait<0>();
      } else {
        if (st.it == 0) { load_tile(p, work, cur, c); cp_async_commit(); }
        if (work + c.ncta < p.nwork) load_tile(p, work + c.ncta, cur ^ 1, c);
        cp_async_commit();
        cp_async_wait<1>();
      }
    } else if (PH == 1) {
      if (STAGE) {
        if (work + c.ncta < p.nwork) load_tile(p, work + c.ncta, cur ^ 1, c);
        cp_async_commit();
      }
      if (jj < R2) {
        const int j = jj;
#pragma unroll
        for (int i = 0; i < R1; ++i) {
          const int n = j + R2 * i;
          int row = n;
          bool nz = true;
          if (PAD) {
            if (n > KMAX) { row = n - (M - NC); nz = (n >= M - KMAX); }
          }
          cplx v = make_double2(0.0, 0.0);
          if (nz) v = TFAST ? B[row * T + t] : B[t * LENP + row];
          st.re[i] = v.x; st.im[i] = v.y;
        }
        const cplx* W = twid(c.smem) + j;
        RegFFT<R1, DIR>::run(as_arr<R1>(st.re), as_arr<R1>(st.im));
#pragma unroll
        for (int k1 = 1; k1 < R1; ++k1) {
          const cplx w = W[k1 * R2];
          const double cc = w.x, ss = (DIR > 0) ? -w.y : w.y;
          const double a = st.re[k1], b = st.im[k1];
          st.re[k1] = a * cc - b * ss;
          st.im[k1] = a * ss + b * cc;
        }
      }
    } else if (PH == 2) {
      if (jj < R2) {
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) B[xidx(t, jj * F::SK + k1)] = make_double2(st.re[k1], st.im[k1]);
      }
    } else if (PH == 3) {
      if (jj < R1) {
        const int k1 = jj;
#pragma unroll
        for (int j = 0; j < R2; ++j) {
          const cplx v = B[xidx(t, j * F::SK + k1)];
          st.re[j] = v.x; st.im[j] = v.y;
        }
        stage2<F, DIR>(st.re, st.im);
        int f, a, bt;
        decode(p, work, f, a, bt);
        const int b = bt * T + t;
        if (STAGE) {
          // results stay in registers until every thread has read the exchange buffer (barrier), phases 4 / 5 ship them
        } else if (b < p.nB) {
          if (PAD && !TFAST && p.peer_mode == 1) {
            const long long line = p.peer_off + (long long)(p.b0 + b) * p.out_sB;
#pragma unroll
            for (int k2 = 0; k2 < R2; ++k2) {
              const int off = st.ooff[k2];
              p.peer_out[f][off >> 24][line + (off & 0xffffff)] = make_double2(st.re[k2] * p.scale, st.im[k2] * p.scale);
            }
          } else {
            cplx* dst;
            if (!PAD && TFAST && p.peer_mode == 2)
              dst = p.peer_out[f][a / p.peer_rows] + p.peer_off + (long long)(a % p.peer_rows) * p.out_sA + (long long)(p.b0 + b) * p.out_sB;
            else
              dst = p.out[f] + (long long)a * p.out_sA + (long long)(p.b0 + b) * p.out_sB;
            if (!PAD && TFAST && p.peer_mode == 2) {
#pragma unroll
              for (int k2 = 0; k2 < R2; ++k2) {
                const int off = st.ooff[k2];
                if (off >= 0) st_peer(dst + off, st.re[k2] * p.scale, st.im[k2] * p.scale);
              }
            } else if (p.hint_out) {
              const unsigned long long pol = pols(c.smem)[1];
#pragma unroll
              for (int k2 = 0; k2 < R2; ++k2) {
                const int off = st.ooff[k2];
                if (PAD || off >= 0) st_cplx_hint(dst + off, st.re[k2] * p.scale, st.im[k2] * p.scale, pol);
              }
            } else {
#pragma unroll
              for (int k2 = 0; k2 < R2; ++k2) {
                const int off = st.ooff[k2];
                if (PAD || off >= 0) dst[off] = make_double2(st.re[k2] * p.scale, st.im[k2] * p.scale);
              }
            }
          }
        }
      }
      if (!STAGE) st.it++;
    }
    if (STAGE && PH == 4) {
      // truncated, scaled results -> staging tile [compact row][T] in the consumed exchange buffer
      if (jj < R1) {
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) {
          const int cidx = compact_index(jj + R1 * k2, M, KMAX);
          if (cidx >= 0) B[cidx * T + t] = make_double2(st.re[k2] * p.scale, st.im[k2] * p.scale);
        }
      }
      bulk_fence_smem();
    }
    if (STAGE && PH == 5) {
      if (c.tid < 32) {
        int f, a, bt;
        decode(p, work, f, a, bt);
        cplx* dst = p.peer_out[f][a / p.peer_rows] + p.peer_off + (long long)(a % p.peer_rows) * p.out_sA + (long long)(p.b0 + bt * T) * p.out_sB;
        for (int row = c.tid; row < NC; row += 32) bulk_store(dst + (long long)row * p.out_sN, &B[row * T], (unsigned)(T * sizeof(cplx)));
        bulk_commit();
      }
      st.it++;
    }
  }
};

}  // namespace smo
