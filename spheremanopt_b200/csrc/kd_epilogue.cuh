// Coefficient-space (diagonal-in-Fourier) kernels of the kinematic-dynamo path.
//
// Replaces Dedalus' per-pencil sparse solves of (M/dt + L/2) X^{n+1} = F^n + (M/dt - L/2) X^n for the
// 4-variable (Pi,A,B,C) CNAB1 system of FWD_Solve_KDyn.py:431-443 and the 8-variable adjoint system of
// FWD_Solve_KDyn.py:855-886, the LBVP of Compatib_Cond (FWD_Solve_KDyn.py:733-748) and the final
// "undo the last implicit solve" of FWD_Solve_KDyn.py:979-989, by their closed forms.
//
// Eliminating Pi from the 4x4 pencil (alpha = 1/dt + k^2/2Rm, beta = 1/dt - k^2/2Rm, P_k = I - k k^T/k^2):
//     X' = P_k[F + beta X]/alpha - k (k.X)/k^2 ,   k = 0 -> 0,
// which is exactly the CNAB1 update for ANY input (Pi drops out; for solenoidal X the last term vanishes).
// Coefficient arrays are complex [nkx][Nc][Pc] (kx slab, ky, kz with pitch Pc >= Nc), Dedalus mode order.
#pragma once
#include "smo_common.cuh"

namespace smo {

struct EpiParams {
  const cplx* a[MAXF];   // inputs  (meaning depends on the kernel)
  const cplx* b[MAXF];
  cplx* o[MAXF];         // outputs
  cplx* o2[MAXF];
  int nwork, nsteps;
  long long n;           // nkx*Nc*Pc
  int Nc, Pc, kmax, kx0;
  double kfac;           // 2 pi / L
  double Rm, dt;
  int flag;
};

struct Wave {
  double kx, ky, kz, k2;
  bool valid;
};
SMO_HD Wave wave_of(const EpiParams& p, long long idx) {
  Wave w;
  const int iz = (int)(idx % p.Pc);
  const long long r = idx / p.Pc;
  const int iy = (int)(r % p.Nc);
  const int ix = (int)(r / p.Nc);
  w.valid = iz < p.Nc;
  w.kx = p.kfac * (double)(p.kx0 + ix);
  w.ky = p.kfac * (double)(iy <= p.kmax ? iy : iy - p.Nc);
  w.kz = p.kfac * (double)(iz <= p.kmax ? iz : iz - p.Nc);
  w.k2 = w.kx * w.kx + w.ky * w.ky + w.kz * w.kz;
  return w;
}

struct C3 { cplx x, y, z; };
SMO_HD C3 load3(const cplx* const* a, int off, long long idx) {
  C3 v; v.x = a[off][idx]; v.y = a[off + 1][idx]; v.z = a[off + 2][idx]; return v;
}
SMO_HD void store3(cplx* const* o, int off, long long idx, const C3& v) {
  o[off][idx] = v.x; o[off + 1][idx] = v.y; o[off + 2][idx] = v.z;
}
SMO_HD cplx cmk(double r, double i) { return make_double2(r, i); }
// i k x v
SMO_HD C3 curl3(const Wave& w, const C3& v) {
  C3 c;
  c.x = cmk(-(w.ky * v.z.y - w.kz * v.y.y), w.ky * v.z.x - w.kz * v.y.x);
  c.y = cmk(-(w.kz * v.x.y - w.kx * v.z.y), w.kz * v.x.x - w.kx * v.z.x);
  c.z = cmk(-(w.kx * v.y.y - w.ky * v.x.y), w.kx * v.y.x - w.ky * v.x.x);
  return c;
}
// (k.v)/k2
SMO_HD cplx kdot_over_k2(const Wave& w, const C3& v) {
  const double inv = 1.0 / w.k2;
  return cmk((w.kx * v.x.x + w.ky * v.y.x + w.kz * v.z.x) * inv, (w.kx * v.x.y + w.ky * v.y.y + w.kz * v.z.y) * inv);
}
// P_k v * s  -  k * d
SMO_HD C3 proj_scale_minus(const Wave& w, const C3& v, double s, cplx d) {
  const cplx kv = kdot_over_k2(w, v);
  C3 o;
  o.x = cmk((v.x.x - w.kx * kv.x) * s - w.kx * d.x, (v.x.y - w.kx * kv.y) * s - w.kx * d.y);
  o.y = cmk((v.y.x - w.ky * kv.x) * s - w.ky * d.x, (v.y.y - w.ky * kv.y) * s - w.ky * d.y);
  o.z = cmk((v.z.x - w.kz * kv.x) * s - w.kz * d.x, (v.z.y - w.kz * kv.y) * s - w.kz * d.y);
  return o;
}
SMO_HD C3 zero3() { C3 z; z.x = z.y = z.z = cmk(0.0, 0.0); return z; }
SMO_HD C3 axpy3(double s, const C3& x, const C3& y) {   // s*x + y
  C3 o;
  o.x = cmk(s * x.x.x + y.x.x, s * x.x.y + y.x.y);
  o.y = cmk(s * x.y.x + y.y.x, s * x.y.y + y.y.y);
  o.z = cmk(s * x.z.x + y.z.x, s * x.z.y + y.z.y);
  return o;
}

enum { EPI_COMPAT = 2, EPI_FINAL = 3, EPI_CURL = 4, EPI_NUFIN = 5 };

// (the per-step updates - CNAB1 step of B, adjoint step of G - live in the fused z step, zstep.cuh, and use the helpers above)
// EPI_COMPAT: b[0..2] = B^N -> o[0..2] = G^0, o2[0..2] = i k x G^0   (flag&1 Integrated, flag&2 Continuous)
// EPI_FINAL : b[0..2] = G^N -> o[0..2] = dt*alpha*G^N  (flag&2 Continuous: plain copy)
// EPI_CURL  : b[0..2] = G   -> o2[0..2] = i k x G        (segment boundaries of a checkpointed adjoint sweep)
// EPI_NUFIN : b[0..2] = to_coef(sum_m (curl G^m) x B_f^m) -> o[0..2] = nu^N = -dt P_k[b]   (KD:874-877 summed over the sweep:
//             nu' = nu + dt P_k[-H] with nu^0 = 0 is linear in H, so the sum is taken before the transforms, see xpass.cuh)
template <int KIND> struct EpiKernel {
  typedef EpiParams Params;
  static constexpr int THREADS = 256;
  static constexpr int NPHASES = 1;
  static constexpr int MIN_BLOCKS = 1;
  static constexpr size_t SMEM = 0;
  struct State {};

  template <int PH>
  SMO_HD static void phase(const Params& p, int work, int, int tid, unsigned char*, State&) {
    const long long idx = (long long)work * THREADS + tid;
    if (idx >= p.n) return;
    const Wave w = wave_of(p, idx);
    if (!w.valid) return;
    const bool k0 = (w.k2 == 0.0);
    const double alpha = 1.0 / p.dt + w.k2 / (2.0 * p.Rm);
    if (KIND == EPI_COMPAT) {
      C3 G = zero3(), Wn = zero3();
      const C3 f = load3(p.b, 0, idx);
      if (p.flag & 2) {
        G = axpy3(-2.0, f, zero3());
        Wn = curl3(w, G);
      } else if (!k0) {
        const double den = (p.flag & 1) ? alpha : (1.0 + p.dt * w.k2 / (2.0 * p.Rm));
        G = proj_scale_minus(w, f, -2.0 / den, cmk(0.0, 0.0));
        Wn = curl3(w, G);
      }
      store3(p.o, 0, idx, G);
      store3(p.o2, 0, idx, Wn);
    } else if (KIND == EPI_NUFIN) {
      store3(p.o, 0, idx, k0 ? zero3() : proj_scale_minus(w, load3(p.b, 0, idx), -p.dt, cmk(0.0, 0.0)));
    } else if (KIND == EPI_CURL) {
      store3(p.o2, 0, idx, k0 ? zero3() : curl3(w, load3(p.b, 0, idx)));
    } else {
      const C3 G = load3(p.b, 0, idx);
      const double s = (p.flag & 2) ? 1.0 : p.dt * alpha;
      store3(p.o, 0, idx, axpy3(s, G, zero3()));
    }
  }
};

}  // namespace smo
