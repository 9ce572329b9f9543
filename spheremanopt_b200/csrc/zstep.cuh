// Fused z pass of the kinematic-dynamo time loops: forward FFT along z + truncation, the diagonal-in-Fourier
// implicit update of the time stepper, and the zero-padded inverse FFT along z of the NEXT step's operands - one
// kernel, one HBM round trip of the pencil data per time step.
//
// Replaces, per time step of FWD_Solve_KDyn.py:635-641 (forward) and :955-961 (adjoint), the sequence
//   FftPass<-1> (p1 -> coefficients)  +  a pointwise update kernel  +  FftPass<+1> (coefficients -> p1)
// of an unfused path (the first working version of this library): the right-hand side coefficients never travel to HBM, the state (B^n | G, nu) is read once
// and written once, and what the next step needs on the grid side (B^{n+1} | curl G', B_f of the next snapshot)
// goes straight from shared memory into the inverse transform.
//
// A work item is a *triplet*: the three vector components of T adjacent z lines (a z line = all kz of one (kx, ky)).
//   forward step : 1 triplet per tile  - in = to_p1(U x B^n),        state B^n -> B^{n+1},   next operand B^{n+1}
//   adjoint step : 1 triplet per tile  - in = to_p1((curl G) x U),   state G -> G',          next operand curl G'
//   (the gradient integrand nu is accumulated on the x-spectra by the x pass and transformed once after the sweep; the
//    forward state B_f is read by the x pass in x-spectral form straight from its snapshot slot)
// Thread (f, t, jj): component f, line t, stage thread jj < RT.  With RT dividing 32 all exchanges of a line stay
// inside one warp, so only the two barriers around the pointwise update (which mixes the three components) are
// CTA-wide; everything else is a __syncwarp (smo_common.cuh: sync_after).
// Shared memory: landing buffer [3][T][M] (cp.async target of the NEXT tile, prefetched as soon as stage 1 has
// consumed the current one), work buffer [3][T][XP] (exchange -> compact coefficients -> exchange), twiddles, maps.
#pragma once
#include "fft_core.cuh"
#include "kd_epilogue.cuh"

#ifndef SMO_ZS_MB
#define SMO_ZS_MB 4     // resident CTAs per SM the register allocation of the fused z step is bounded for (55 KB of smem each)
#endif

namespace smo {

struct ZParams {
  const cplx* in[MAXF];     // p1 inputs (kx-slab side), one per field
  cplx* out[MAXF];          // p1 outputs (local buffer; peer buffers below when peer_mode != 0)
  const cplx* b[MAXF];      // coefficient state in  : forward B^n[3]    | adjoint G[3]
  cplx* o[MAXF];            // coefficient state out : forward B^{n+1}[3] | adjoint G'[3]
  int nwork, nsteps;        // nwork = tiles * ntrip
  int ntrip, mode;          // mode 0: forward CNAB1 step; mode 1: adjoint step; ntrip = triplets per tile (1)
  int nlines, tiles;        // z lines of this rank = nkx*Nc
  int do_inv;               // 0: last step of a loop, nothing follows on the grid side
  int Nc, Pc, kmax, kx0;
  int line_stride;          // p1 elements between consecutive lines inside a block (= nz)
  int seglen;               // multi-rank p1 layout [s][nkx][Nc][nz]: z index n lives in block n / seglen (0: one rank)
  long long blk;
  double kfac, Rm, dt, scale;
  const cplx* tw;
  int peer_mode;            // 1: segment s of every output line is stored straight into rank s's buffer
  long long peer_off;
  cplx* peer_out[MAXF][MAXP];
  int pull_mode;            // 1: segment s of every input line is read straight out of rank s's buffer (stores stay local)
  long long pull_off;
  const cplx* peer_in[MAXF][MAXP];
  XSync xs;                 // cross-GPU wait / signal fused into the launch (smo_common.cuh)
  int l2_hints;             // 1: pencil lines are read evict_first and written evict_last, the coefficient state evict_first
  int bulk_push;            // peer_mode 1: the output lines are staged in shared memory (per warp: [peer][line][seglen], in place of the
                            // consumed exchange lines) and shipped with ONE TMA bulk store per (warp, peer) instead of 16-byte stores
};

template <class F, int T_> struct ZStep {
  typedef ZParams Params;
  static constexpr bool V2 = true;
  static constexpr int T = T_, M = F::M, R1 = F::R1, R2 = F::R2, RT = F::RT, XP = F::XP;
  static constexpr int KMAX = M / 3 - 1, NC = 2 * KMAX + 1, PC = NC + 1;
  // lanes reserved per line: RT, or a whole warp when RT does not divide 32 (384 = 24 x 16: 24 stage threads) - a line is then
  // private to one warp (8 idle lanes in the FFT stages, all 32 in the copies) and the warp-level barriers apply
  static constexpr int LP = (F::RT > 16 && F::RT < 32) ? 32 : F::RT;
  static constexpr int THREADS = 3 * T_ * LP;
  static constexpr int NPHASES = 9;
  static constexpr int MIN_BLOCKS = (F::RT > 16) ? 2 : SMO_ZS_MB;
  static constexpr bool WARP_OK = (32 % LP == 0);          // the threads of a line never straddle a warp
  static constexpr int LW = WARP_OK ? ((32 / LP < T_) ? 32 / LP : T_) : 1;   // lines of one component a warp handles
  static constexpr bool BULK_OK = WARP_OK && (T_ % LW == 0) && (LP * LW == 32);   // staged push: a warp owns LW whole lines of ONE component
  static constexpr int LAND = 3 * T_ * M, WORK = 3 * T_ * XP, STATE = 3 * T_ * PC;
  static constexpr size_t SMEM = (size_t)(LAND + WORK + STATE + M) * sizeof(cplx) + 2 * (size_t)M * sizeof(int) + 16;
  static_assert(XP >= PC, "compact coefficient line must fit into the exchange line");
  struct State {
    double re[F::RT], im[F::RT];
    int it;
  };
  // barrier after phase PH: 1 = warp, 2 = CTA
  SMO_HD static constexpr int sync_after(int ph) { return (!WARP_OK || ph == 4 || ph == 5) ? 2 : 1; }

  SMO_HD static cplx* land(unsigned char* s) { return reinterpret_cast<cplx*>(s); }
  SMO_HD static cplx* wrk(unsigned char* s) { return land(s) + LAND; }
  SMO_HD static cplx* stl(unsigned char* s) { return wrk(s) + WORK; }      // coefficient state of the tile (cp.async target)
  SMO_HD static cplx* twid(unsigned char* s) { return stl(s) + STATE; }
  SMO_HD static int* segidx(unsigned char* s) { return reinterpret_cast<int*>(twid(s) + M); }   // n / seglen
  SMO_HD static int* segrem(unsigned char* s) { return segidx(s) + M; }                         // n % seglen
  SMO_HD static unsigned long long* pols(unsigned char* s) { return reinterpret_cast<unsigned long long*>(segrem(s) + M); }

  SMO_HD static void split_tid(int tid, int& f, int& t, int& jj) {
    jj = tid % LP;
    t = (tid / LP) % T;
    f = tid / (LP * T);
  }
  SMO_HD static void decode(const Params& p, int work, int& tile, int& trip) {
    tile = work / p.ntrip;
    trip = work - tile * p.ntrip;
  }
  // the thread's own line of the landing buffer <- p1 line (asynchronous; only the RT threads of a line touch it)
  SMO_HD static void load_tile(const Params& p, int work, const Ctx& c) {
    int f, t, jj, tile, trip;
    split_tid(c.tid, f, t, jj);
    decode(p, work, tile, trip);
    const int b = tile * T + t;
    if (b >= p.nlines) return;
    cplx* Ld = land(c.smem) + (f * T + t) * M;
    const cplx* src = p.in[3 * trip + f] + (long long)b * p.line_stride;
    if (p.seglen <= 0 && p.l2_hints) {
      const unsigned long long pol = pols(c.smem)[0];
      for (int e = jj; e < M; e += LP) cp_async16_hint(&Ld[e], src + e, pol);
    } else if (p.seglen <= 0) {
      for (int e = jj; e < M; e += LP) cp_async16(&Ld[e], src + e);
    } else if (p.pull_mode == 1) {
      const int* si = segidx(c.smem);
      const int* sr = segrem(c.smem);
      const long long line = p.pull_off + (long long)b * p.line_stride;
      for (int e = jj; e < M; e += LP) cp_async16(&Ld[e], p.peer_in[3 * trip + f][si[e]] + line + sr[e]);
    } else {
      const int* si = segidx(c.smem);
      const int* sr = segrem(c.smem);
      for (int e = jj; e < M; e += LP) cp_async16(&Ld[e], src + (long long)si[e] * p.blk + sr[e]);
    }
  }

  // the thread's own line of the coefficient state (B^n | G | nu) -> shared memory, asynchronously; issued once the
  // pointwise update of the previous tile has consumed the buffer
  SMO_HD static void load_state(const Params& p, int work, const Ctx& c) {
    int f, t, jj, tile, trip;
    split_tid(c.tid, f, t, jj);
    decode(p, work, tile, trip);
    const int b = tile * T + t;
    if (b >= p.nlines) return;
    cplx* Sd = stl(c.smem) + (f * T + t) * PC;
    const cplx* src = p.b[3 * trip + f] + (long long)b * p.Pc;
    if (p.l2_hints) {
      const unsigned long long pol = pols(c.smem)[0];
      for (int e = jj; e < PC; e += LP) cp_async16_hint(&Sd[e], src + e, pol);
    } else {
      for (int e = jj; e < PC; e += LP) cp_async16(&Sd[e], src + e);
    }
  }
  SMO_HD static void init(const Params& p, const Ctx& c, State& st) {
    cplx* W = twid(c.smem);
    for (int m = c.tid; m < M; m += THREADS) W[m] = ldg_c(p.tw + ((m % R2) * (m / R2)) % M);
    int* si = segidx(c.smem);
    int* sr = segrem(c.smem);
    for (int m = c.tid; m < M; m += THREADS) {
      si[m] = p.seglen > 0 ? m / p.seglen : 0;
      sr[m] = p.seglen > 0 ? m % p.seglen : m;
    }
    if (c.tid == 0) { pols(c.smem)[0] = l2_policy(1); pols(c.smem)[1] = l2_policy(2); }
    st.it = 0;
  }

  template <int PH>
  SMO_HD static void phase2(const Params& p, int work, int /*step*/, const Ctx& c, State& st) {
    int f, t, jj, tile, trip;
    split_tid(c.tid, f, t, jj);
    decode(p, work, tile, trip);
    const int b = tile * T + t;
    const bool live = b < p.nlines;
    const bool fwd = live;
    cplx* Ld = land(c.smem) + (f * T + t) * M;
    cplx* Wk = wrk(c.smem) + (f * T + t) * XP;
    if (PH == 0) {
      if (st.it == 0) { load_tile(p, work, c); load_state(p, work, c); }
      cp_async_commit();
      cp_async_wait<0>();
    } else if (PH == 1) {
      // forward stage 1: thread j < R2 owns z samples j + R2*i
      if (jj < R2 && fwd) {
#pragma unroll
        for (int i = 0; i < R1; ++i) {
          const cplx v = Ld[jj + R2 * i];
          st.re[i] = v.x; st.im[i] = v.y;
        }
        RegFFT<R1, -1>::run(as_arr<R1>(st.re), as_arr<R1>(st.im));
        const cplx* W = twid(c.smem) + jj;
#pragma unroll
        for (int k1 = 1; k1 < R1; ++k1) {
          const cplx w = W[k1 * R2];
          const double a = st.re[k1], bb = st.im[k1];
          st.re[k1] = a * w.x - bb * w.y;
          st.im[k1] = a * w.y + bb * w.x;
        }
      }
    } else if (PH == 2) {
      // the landing line is consumed: stream in the same line of this CTA's next work item
      if (work + c.ncta < p.nwork) load_tile(p, work + c.ncta, c);
      cp_async_commit();
      if (BULK_OK && p.bulk_push) {      // the previous tile's bulk stores (issued by this warp's first lane) have read the lines
        if ((c.tid & 31) == 0) bulk_wait_read();
        warp_sync();
      }
      if (jj < R2 && fwd) {
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) Wk[jj * F::SK + k1] = make_double2(st.re[k1], st.im[k1]);
      }
    } else if (PH == 3) {
      if (jj < R1 && fwd) {
#pragma unroll
        for (int j = 0; j < R2; ++j) {
          const cplx v = Wk[j * F::SK + jj];
          st.re[j] = v.x; st.im[j] = v.y;
        }
        stage2<F, -1>(st.re, st.im);
      }
    } else if (PH == 4) {
      // truncate: retained modes, scaled, in the compact order [0..kmax, -kmax..-1] (overwrites the exchange line)
      if (jj < R1 && fwd) {
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) {
          const int cidx = compact_index(jj + R1 * k2, M, KMAX);
          if (cidx >= 0) Wk[cidx] = make_double2(st.re[k2] * p.scale, st.im[k2] * p.scale);
        }
      }
    } else if (PH == 5) {
      // pointwise implicit update of the three components at (line, kz); consecutive threads -> consecutive kz
      const int kind = p.mode;                            // 0 forward step, 1 adjoint G
      cplx* Wt = wrk(c.smem);
      for (int e = c.tid; e < T * PC; e += THREADS) {
        const int tt = e / PC, iz = e - tt * PC;
        const int bl = tile * T + tt;
        if (bl >= p.nlines || iz >= NC) continue;
        cplx* w0 = Wt + (0 * T + tt) * XP + iz;
        cplx* w1 = Wt + (1 * T + tt) * XP + iz;
        cplx* w2 = Wt + (2 * T + tt) * XP + iz;
        const cplx* Ss = stl(c.smem) + tt * PC + iz;
        C3 S; S.x = Ss[0]; S.y = Ss[T * PC]; S.z = Ss[2 * T * PC];
        const int ix = bl / p.Nc, iy = bl - ix * p.Nc;
        Wave w;
        w.kx = p.kfac * (double)(p.kx0 + ix);
        w.ky = p.kfac * (double)(iy <= p.kmax ? iy : iy - p.Nc);
        w.kz = p.kfac * (double)(iz <= p.kmax ? iz : iz - p.Nc);
        w.k2 = w.kx * w.kx + w.ky * w.ky + w.kz * w.kz;
        w.valid = true;
        const long long idx = (long long)bl * p.Pc + iz;
        C3 A; A.x = *w0; A.y = *w1; A.z = *w2;
        C3 nw = zero3(), so = zero3();    // next operand (-> inverse transform), new state (-> HBM)
        if (w.k2 != 0.0) {
          const double alpha = 1.0 / p.dt + w.k2 / (2.0 * p.Rm), beta = 1.0 / p.dt - w.k2 / (2.0 * p.Rm);
          if (kind == 0) {
            so = proj_scale_minus(w, axpy3(beta, S, curl3(w, A)), 1.0 / alpha, kdot_over_k2(w, S));
            nw = so;
          } else {
            so = proj_scale_minus(w, axpy3(beta, S, A), 1.0 / alpha, kdot_over_k2(w, S));
            nw = curl3(w, so);
          }
        }
        store3(p.o, 0, idx, so);
        *w0 = nw.x; *w1 = nw.y; *w2 = nw.z;
      }
    } else if (PH == 6) {
      // the state buffer is consumed (CTA barrier after phase 5): stream in the state of this CTA's next work item
      if (work + c.ncta < p.nwork) load_state(p, work + c.ncta, c);
      cp_async_commit();
      // inverse stage 1 on the zero-padded compact line
      if (p.do_inv && jj < R2 && live) {
#pragma unroll
        for (int i = 0; i < R1; ++i) {
          const int n = jj + R2 * i;
          cplx v = make_double2(0.0, 0.0);
          if (n <= KMAX) v = Wk[n];
          else if (n >= M - KMAX) v = Wk[n - (M - NC)];
          st.re[i] = v.x; st.im[i] = v.y;
        }
        RegFFT<R1, +1>::run(as_arr<R1>(st.re), as_arr<R1>(st.im));
        const cplx* W = twid(c.smem) + jj;
#pragma unroll
        for (int k1 = 1; k1 < R1; ++k1) {
          const cplx w = W[k1 * R2];
          const double a = st.re[k1], bb = st.im[k1];
          st.re[k1] = a * w.x + bb * w.y;
          st.im[k1] = bb * w.x - a * w.y;
        }
      }
    } else if (PH == 7) {
      if (p.do_inv && jj < R2 && live) {
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) Wk[jj * F::SK + k1] = make_double2(st.re[k1], st.im[k1]);
      }
    } else if (BULK_OK && p.bulk_push && p.peer_mode == 1 && p.do_inv) {
      // staged push: the LW lines of this warp go, segment by segment, to their owners with one bulk store per peer
      const bool act = jj < R1 && live;
      if (act) {
#pragma unroll
        for (int j = 0; j < R2; ++j) {
          const cplx v = Wk[j * F::SK + jj];
          st.re[j] = v.x; st.im[j] = v.y;
        }
      }
      warp_sync();                                   // every input of the warp's lines is in registers: the lines may be overwritten
      const int t0 = t - t % LW;                       // first line of this warp
      cplx* Sg = wrk(c.smem) + (f * T + t0) * XP;     // staging area = the warp's LW exchange lines, as [peer][line][seglen]
      if (act) {
        stage2<F, +1>(st.re, st.im);
        const int* si = segidx(c.smem);
        const int* sr = segrem(c.smem);
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) {
          const int k = jj + R1 * k2;
          Sg[(si[k] * LW + (t - t0)) * p.seglen + sr[k]] = make_double2(st.re[k2], st.im[k2]);
        }
      }
      bulk_fence_smem();
      warp_sync();
      if ((c.tid & 31) == 0) {
        const int b0 = tile * T + t0;
        const int nlive = imin(LW, p.nlines - b0);
        if (nlive > 0) {
          const int fo = 3 * trip + f;
          const int nseg = M / p.seglen;
          for (int s2 = 0; s2 < nseg; ++s2)
            bulk_store(p.peer_out[fo][s2] + p.peer_off + (long long)b0 * p.line_stride, Sg + (size_t)s2 * LW * p.seglen,
                       (unsigned)(nlive * p.seglen * sizeof(cplx)));
          bulk_commit();
        }
      }
      st.it++;
    } else {
      if (p.do_inv && jj < R1 && live) {
#pragma unroll
        for (int j = 0; j < R2; ++j) {
          const cplx v = Wk[j * F::SK + jj];
          st.re[j] = v.x; st.im[j] = v.y;
        }
        stage2<F, +1>(st.re, st.im);
        const int fo = 3 * trip + f;
        const int* si = segidx(c.smem);
        const int* sr = segrem(c.smem);
        const long long line = (long long)b * p.line_stride;
        if (p.peer_mode == 1) {
#pragma unroll
          for (int k2 = 0; k2 < R2; ++k2) {
            const int k = jj + R1 * k2;
            st_peer(&p.peer_out[fo][si[k]][p.peer_off + line + sr[k]], st.re[k2], st.im[k2]);
          }
        } else if (p.l2_hints) {
          cplx* dst = p.out[fo] + line;
          const unsigned long long pol = pols(c.smem)[1];
#pragma unroll
          for (int k2 = 0; k2 < R2; ++k2) {
            const int k = jj + R1 * k2;
            st_cplx_hint(dst + ((long long)si[k] * p.blk + sr[k]), st.re[k2], st.im[k2], pol);
          }
        } else {
          cplx* dst = p.out[fo] + line;
#pragma unroll
          for (int k2 = 0; k2 < R2; ++k2) {
            const int k = jj + R1 * k2;
            dst[(long long)si[k] * p.blk + sr[k]] = make_double2(st.re[k2], st.im[k2]);
          }
        }
      }
      st.it++;
    }
  }
};

}  // namespace smo
