// Two-stage shared-memory FFT building blocks (fp64).
//
// A length-M = R1*R2 complex DFT is split Cooley-Tukey style into
//   stage 1: R2 threads, thread j owns x[j + R2*i], i < R1, does a radix-R1 register FFT (codelets.cuh)
//            and multiplies output k1 by the twiddle w_M^(j*k1);
//   exchange through shared memory;
//   stage 2: R1 threads, thread k1 gathers the R2 values of that k1, does a radix-R2 register FFT and
//            holds X[k1 + R1*k2], k2 < R2.
// The transform with the roles of R1 and R2 swapped consumes x[j' + R1*i'] in stage 1, i.e. exactly what a
// stage-2 thread of the un-swapped transform holds; inverse-then-forward chains (fused x-pass, SH23 time loop)
// use that to hand data from one transform to the next in registers or with one coalesced exchange.
#pragma once
#include "smo_common.cuh"
#include "codelets.cuh"

namespace smo {

template <int R1_, int R2_> struct Fac {
  static constexpr int R1 = R1_, R2 = R2_, M = R1_ * R2_;
  static constexpr int RT = (R1_ > R2_) ? R1_ : R2_;   // threads per line
  static constexpr int SK = (R1_ % 2 == 0) ? R1_ + 1 : R1_;   // odd pitch of one stage-1 thread's outputs
  static constexpr int XP = R2_ * SK;                  // exchange-buffer length of one line
  typedef Fac<R2_, R1_> Swapped;
};

// supported dealiased lengths M -> factorisation (R1 >= R2 keeps the radix-R1 stage the wide one)
template <int M> struct FacOf;
template <> struct FacOf<24>  { typedef Fac<6, 4> type; };
template <> struct FacOf<36>  { typedef Fac<6, 6> type; };
#if defined(SMO_TEST_R24)   // test-only factorisation: exercises the 24-thread (one FFT per warp) code paths of M = 384 at a small size
template <> struct FacOf<48>  { typedef Fac<24, 2> type; };
#else
template <> struct FacOf<48>  { typedef Fac<8, 6> type; };
#endif
template <> struct FacOf<64>  { typedef Fac<8, 8> type; };
template <> struct FacOf<72>  { typedef Fac<9, 8> type; };
template <> struct FacOf<96>  { typedef Fac<12, 8> type; };
template <> struct FacOf<128> { typedef Fac<16, 8> type; };
template <> struct FacOf<144> { typedef Fac<12, 12> type; };
template <> struct FacOf<192> { typedef Fac<16, 12> type; };
template <> struct FacOf<256> { typedef Fac<16, 16> type; };
template <> struct FacOf<288> { typedef Fac<18, 16> type; };
template <> struct FacOf<384> { typedef Fac<24, 16> type; };

template <int R> SMO_HD double (&as_arr(double* p))[R] { return *reinterpret_cast<double(*)[R]>(p); }

// radix-R1 register FFT followed by the inter-stage twiddles w_M^(DIR*j*k1).  tw[m] = exp(-2 pi i m / M).
template <class F, int DIR> SMO_HD void stage1(double* xr, double* xi, int j, const cplx* tw) {
  RegFFT<F::R1, DIR>::run(as_arr<F::R1>(xr), as_arr<F::R1>(xi));
#pragma unroll
  for (int k1 = 1; k1 < F::R1; ++k1) {
    const cplx w = ldg_c(tw + j * k1);
    const double c = w.x, s = (DIR > 0) ? -w.y : w.y;
    const double a = xr[k1], b = xi[k1];
    xr[k1] = a * c - b * s;
    xi[k1] = a * s + b * c;
  }
}
template <class F, int DIR> SMO_HD void stage2(double* xr, double* xi) {
  RegFFT<F::R2, DIR>::run(as_arr<F::R2>(xr), as_arr<F::R2>(xi));
}

// x[k] *= (wr + i*wi)^k for k = 1..R-1, the powers generated on the fly by two interleaved recurrences (odd / even
// exponents) instead of being loaded: the fused x passes are bound by shared-memory bandwidth, the fp64 pipe has room.
template <int R> SMO_HD void twiddle_powers(double* xr, double* xi, double wr, double wi) {
  const double w2r = wr * wr - wi * wi, w2i = 2.0 * wr * wi;
  double or_ = wr, oi = wi;      // odd exponents: w, w^3, ...
  double er = w2r, ei = w2i;     // even exponents: w^2, w^4, ...
#pragma unroll
  for (int k = 1; k < R; ++k) {
    const double a = xr[k], b = xi[k];
    if (k & 1) {
      xr[k] = a * or_ - b * oi; xi[k] = a * oi + b * or_;
      const double t = or_ * w2r - oi * w2i; oi = or_ * w2i + oi * w2r; or_ = t;
    } else {
      xr[k] = a * er - b * ei; xi[k] = a * ei + b * er;
      const double t = er * w2r - ei * w2i; ei = er * w2i + ei * w2r; er = t;
    }
  }
}

// position of FFT index n inside a compact array of the 2*kmax+1 retained modes [0..kmax, -kmax..-1]; -1 if dropped
SMO_HD int compact_index(int n, int M, int kmax) {
  if (n <= kmax) return n;
  if (n >= M - kmax) return n - (M - (2 * kmax + 1));
  return -1;
}

}  // namespace smo
