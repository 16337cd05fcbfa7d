// epgx_common.cuh -- shared definitions of the epgx engine (sm_100a only).
//
// Arithmetic forms of the tape records.  Every form acts on ONE configuration order: the triple
// (F+, F-, Z) held as six reals.  Reference formulas: epgpy/transition.py:114-196 (T and its
// derivatives), evolution.py:220-256 (E/P/R), opscalar.py:213-232, opmatrix.py:199-221,
// diffusion.py:60-79, exchange.py:89-120; see include/epgx.h for the coefficient layouts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "epgx.h"

// internal record of the merged stream: ends a segment (unit shift / reset of the segment that just ran:
// off[0] = shift, off[1] = n_old, off[2] = n_new, aux1 = segment flags) and opens the next one (aux = nact)
#define EPGX_OP_SEG 64
// internal: one whole TR = FUSED + plain ADC + the SEG that follows, merged by the stream builder when a
// segment consists of exactly [FUSED, CONT, ADC(F0, no scale)].  First record as FUSED; second record:
// off[0..1] / pat[0..1] post-E blocks, aux = ADC row, flags = (shift + 1) | segment flags << 2,
// off[2] = n_old << 16 | n_new, aux1 = nact of the next segment
#define EPGX_OP_TR 65
// bit of the first word (flags bit 15) of the FIRST record of a tape window: the window consists of
// TAPE_CHUNK / 2 EPGX_OP_TR record pairs with shift = +1 and no segment flags
#define EPGX_CHUNK_PURE_TR 0x80000000
// complex counterpart: [D?] [FUSED (any pulse kind)] [ADC(F0, optional scale)] + the segment's close, in three
// records: FUSED as TRC; CONT' as for TR; CONT2: flags bit 0 = ADC scale in off[0] / pat[0], bit 1 = D table in
// off[1] / pat[1].  A window of TAPE_CHUNK records holds (TAPE_CHUNK - 1) / 3 of them (flags bit 14).
#define EPGX_OP_TRC 66
// derivative counterpart (real-valued tapes, <= 3 variables): [INJ(DIAG)*] [E] [INJ(T_RE)*] T_RE [INJ(DIAG)*] [E] ADC
// + the segment's close, in five records: TRJ (T block, E_pre blocks; flags PRE / POST / PARTIALS),
// CONT' (E_post blocks, ADC signal row, close), then three CONT records with the blocks of the pre-E, pulse and
// post-E injections of variables 0..2 (off[v] / pat[v], presence in flags bits 0..2); the Jacobian row rides in rsv1
// of CONT'.  Groups start at multiples of five records inside a window (<= 12 per window); flags bit 13 of the
// first record of a window says that it holds at least one of them.
#define EPGX_OP_TRJ 67
#define EPGX_CHUNK_ANY_TRJ 0x20000000
// flags bit 12: the window is made of 12 plain groups (unit shift +1, no segment flags) followed by NOP records
#define EPGX_CHUNK_PURE_TRJ 0x10000000
#define EPGX_CHUNK_PURE_TRC 0x40000000

namespace epgx {

template <typename real> __device__ __forceinline__ real ldc(const real *p) { return __ldg(p); }

template <typename real> struct vec2;
template <> struct vec2<float> { typedef float2 type; };
template <> struct vec2<double> { typedef double2 type; };

// one configuration order of one state set
template <typename real> struct Tri {
  real pr, pi, mr, mi, zr, zi;
};

template <typename real> __device__ __forceinline__ void tri_zero(Tri<real> &t) {
  t.pr = t.pi = t.mr = t.mi = t.zr = t.zi = real(0);
}

template <typename real> __device__ __forceinline__ void tri_add(Tri<real> &t, const Tri<real> &s) {
  t.pr += s.pr; t.pi += s.pi; t.mr += s.mr; t.mi += s.mi; t.zr += s.zr; t.zi += s.zi;
}

// F+' = a F+ + B F- + U Z ; F-' = conj(B) F+ + a F- + conj(U) Z ; Z' = -1/2 (conj(U) F+ + U F-) + w Z
template <typename real>
__device__ __forceinline__ Tri<real> form_t_gen(const Tri<real> &s, real a, real w, real Br, real Bi, real Ur, real Ui) {
  Tri<real> o;
  const real h = real(-0.5);
  o.pr = a * s.pr + Br * s.mr - Bi * s.mi + Ur * s.zr - Ui * s.zi;
  o.pi = a * s.pi + Br * s.mi + Bi * s.mr + Ur * s.zi + Ui * s.zr;
  o.mr = a * s.mr + Br * s.pr + Bi * s.pi + Ur * s.zr + Ui * s.zi;
  o.mi = a * s.mi + Br * s.pi - Bi * s.pr + Ur * s.zi - Ui * s.zr;
  o.zr = w * s.zr + h * (Ur * s.pr + Ui * s.pi + Ur * s.mr - Ui * s.mi);
  o.zi = w * s.zi + h * (Ur * s.pi - Ui * s.pr + Ur * s.mi + Ui * s.mr);
  return o;
}

// B = b, U = u real (phi = +-90 deg): real and imaginary parts decouple
template <typename real>
__device__ __forceinline__ Tri<real> form_t_re(const Tri<real> &s, real a, real w, real b, real u) {
  Tri<real> o;
  const real hu = real(-0.5) * u;
  o.pr = a * s.pr + b * s.mr + u * s.zr;
  o.pi = a * s.pi + b * s.mi + u * s.zi;
  o.mr = a * s.mr + b * s.pr + u * s.zr;
  o.mi = a * s.mi + b * s.pi + u * s.zi;
  o.zr = w * s.zr + hu * (s.pr + s.mr);
  o.zi = w * s.zi + hu * (s.pi + s.mi);
  return o;
}

// B = b real, U = -i u (phi = 0 / 180 deg)
template <typename real>
__device__ __forceinline__ Tri<real> form_t_im(const Tri<real> &s, real a, real w, real b, real u) {
  Tri<real> o;
  const real hu = real(0.5) * u;
  o.pr = a * s.pr + b * s.mr + u * s.zi;
  o.pi = a * s.pi + b * s.mi - u * s.zr;
  o.mr = a * s.mr + b * s.pr - u * s.zi;
  o.mi = a * s.mi + b * s.pi + u * s.zr;
  o.zr = w * s.zr + hu * (s.pi - s.mi);
  o.zi = w * s.zi - hu * (s.pr - s.mr);
  return o;
}

// fused E.T.E with real couplings (T_RE kind): five independent coefficients
template <typename real>
__device__ __forceinline__ Tri<real> form_t5_re(const Tri<real> &s, real a, real w, real b, real u, real h) {
  Tri<real> o;
  o.pr = a * s.pr + b * s.mr + u * s.zr;
  o.pi = a * s.pi + b * s.mi + u * s.zi;
  o.mr = a * s.mr + b * s.pr + u * s.zr;
  o.mi = a * s.mi + b * s.pi + u * s.zi;
  o.zr = w * s.zr + h * (s.pr + s.mr);
  o.zi = w * s.zi + h * (s.pi + s.mi);
  return o;
}

// fused E.T.E, T_IM kind (U = -i u): h = +1/2 e1' e2 u
template <typename real>
__device__ __forceinline__ Tri<real> form_t5_im(const Tri<real> &s, real a, real w, real b, real u, real h) {
  Tri<real> o;
  o.pr = a * s.pr + b * s.mr + u * s.zi;
  o.pi = a * s.pi + b * s.mi - u * s.zr;
  o.mr = a * s.mr + b * s.pr - u * s.zi;
  o.mi = a * s.mi + b * s.pi + u * s.zr;
  o.zr = w * s.zr + h * (s.pi - s.mi);
  o.zi = w * s.zi - h * (s.pr - s.mr);
  return o;
}

// coefficient assembly of a FUSED record (see include/epgx.h)
template <typename real> struct Fused5 {
  real a, w, b, u, h; // per-order coefficients
  real fz, zz;        // affine terms at k = 0: F+-(0) += fz (RE) / -+ i fz (IM), Z(0) += zz
};

template <typename real>
__device__ __forceinline__ Fused5<real> fuse5(real ta, real tw, real tb, real tu, bool pre, real e1a, real r0a, real e2a,
                                              bool post, real e1b, real r0b, real e2b, bool im, real m0) {
  if (!pre) { e1a = real(1); e2a = real(1); r0a = real(0); }
  if (!post) { e1b = real(1); e2b = real(1); r0b = real(0); }
  Fused5<real> f;
  const real ff = e2b * e2a;
  f.a = ff * ta;
  f.b = ff * tb;
  f.u = e2b * e1a * tu;
  f.h = (im ? real(0.5) : real(-0.5)) * e1b * e2a * tu;
  f.w = e1b * e1a * tw;
  const real c1 = r0a * m0; // Z(0) offset of E_pre, pushed through T and E_post
  f.fz = e2b * tu * c1;
  f.zz = e1b * tw * c1 + r0b * m0;
  return f;
}

// fused E.T.E for ANY pulse kind: a, w real; B, U, H complex; H = -1/2 e1' e2 U
//   F+' = a F+ + B F- + U Z ; F-' = conj(B) F+ + a F- + conj(U) Z ; Z' = w Z + conj(H) F+ + H F-
template <typename real> struct Fused8 {
  real a, w, Br, Bi, Ur, Ui, Hr, Hi;
  real fzr, fzi, zz; // affine terms at k = 0: F+(0) += fz, F-(0) += conj(fz), Z(0) += zz
};

template <typename real>
__device__ __forceinline__ Tri<real> form_t8(const Tri<real> &s, const Fused8<real> &f) {
  Tri<real> o;
  o.pr = f.a * s.pr + f.Br * s.mr - f.Bi * s.mi + f.Ur * s.zr - f.Ui * s.zi;
  o.pi = f.a * s.pi + f.Br * s.mi + f.Bi * s.mr + f.Ur * s.zi + f.Ui * s.zr;
  o.mr = f.a * s.mr + f.Br * s.pr + f.Bi * s.pi + f.Ur * s.zr + f.Ui * s.zi;
  o.mi = f.a * s.mi + f.Br * s.pi - f.Bi * s.pr + f.Ur * s.zi - f.Ui * s.zr;
  o.zr = f.w * s.zr + f.Hr * s.pr + f.Hi * s.pi + f.Hr * s.mr - f.Hi * s.mi;
  o.zi = f.w * s.zi + f.Hr * s.pi - f.Hi * s.pr + f.Hr * s.mi + f.Hi * s.mr;
  return o;
}

// T block (a, w, B, U) with real diagonal E_pre = (e2a, e2a, e1a; r0a) and E_post = (e2b, e2b, e1b; r0b)
template <typename real>
__device__ __forceinline__ Fused8<real> fuse8(real ta, real tw, real tBr, real tBi, real tUr, real tUi, bool pre, real e1a,
                                              real r0a, real e2a, bool post, real e1b, real r0b, real e2b, real m0) {
  if (!pre) { e1a = real(1); e2a = real(1); r0a = real(0); }
  if (!post) { e1b = real(1); e2b = real(1); r0b = real(0); }
  Fused8<real> f;
  const real ff = e2b * e2a, fu = e2b * e1a, fh = real(-0.5) * e1b * e2a;
  f.a = ff * ta;
  f.Br = ff * tBr; f.Bi = ff * tBi;
  f.Ur = fu * tUr; f.Ui = fu * tUi;
  f.Hr = fh * tUr; f.Hi = fh * tUi;
  f.w = e1b * e1a * tw;
  const real c1 = r0a * m0; // Z(0) offset of E_pre, pushed through T and E_post
  f.fzr = e2b * tUr * c1; f.fzi = e2b * tUi * c1;
  f.zz = e1b * tw * c1 + r0b * m0;
  return f;
}

// decode the T block of a FUSED record of any kind into (a, w, B, U)
template <typename real>
__device__ __forceinline__ void fused_pulse(const real *ct, int flags, real &a, real &w, real &Br, real &Bi, real &Ur, real &Ui) {
  a = ldc(ct); w = ldc(ct + 1);
  if (flags & EPGX_FLAG_GEN) { Br = ldc(ct + 2); Bi = ldc(ct + 3); Ur = ldc(ct + 4); Ui = ldc(ct + 5); }
  else {
    Br = ldc(ct + 2); Bi = real(0);
    const real u = ldc(ct + 3);
    if (flags & EPGX_FLAG_IM) { Ur = real(0); Ui = -u; } else { Ur = u; Ui = real(0); }
  }
}

// F+ *= (er + i ei), F- *= (er - i ei), Z *= e1     (er + i ei = e2 cis(2 pi g tau))
template <typename real>
__device__ __forceinline__ Tri<real> form_e_g(const Tri<real> &s, real e1, real er, real ei) {
  Tri<real> o;
  o.pr = er * s.pr - ei * s.pi;
  o.pi = er * s.pi + ei * s.pr;
  o.mr = er * s.mr + ei * s.mi;
  o.mi = er * s.mi - ei * s.mr;
  o.zr = e1 * s.zr;
  o.zi = e1 * s.zi;
  return o;
}

template <typename real>
__device__ __forceinline__ Tri<real> form_e(const Tri<real> &s, real e1, real e2) {
  Tri<real> o;
  o.pr = e2 * s.pr; o.pi = e2 * s.pi; o.mr = e2 * s.mr; o.mi = e2 * s.mi;
  o.zr = e1 * s.zr; o.zi = e1 * s.zi;
  return o;
}

// generic diagonal: c = (aP.re, aP.im, aM.re, aM.im, aZ.re, aZ.im)
template <typename real>
__device__ __forceinline__ Tri<real> form_diag(const Tri<real> &s, const real *c) {
  Tri<real> o;
  o.pr = c[0] * s.pr - c[1] * s.pi;
  o.pi = c[0] * s.pi + c[1] * s.pr;
  o.mr = c[2] * s.mr - c[3] * s.mi;
  o.mi = c[2] * s.mi + c[3] * s.mr;
  o.zr = c[4] * s.zr - c[5] * s.zi;
  o.zi = c[4] * s.zi + c[5] * s.zr;
  return o;
}

// generic 3x3 complex, row-major m[18]
template <typename real>
__device__ __forceinline__ Tri<real> form_matrix(const Tri<real> &s, const real *m) {
  Tri<real> o;
  o.pr = m[0] * s.pr - m[1] * s.pi + m[2] * s.mr - m[3] * s.mi + m[4] * s.zr - m[5] * s.zi;
  o.pi = m[0] * s.pi + m[1] * s.pr + m[2] * s.mi + m[3] * s.mr + m[4] * s.zi + m[5] * s.zr;
  o.mr = m[6] * s.pr - m[7] * s.pi + m[8] * s.mr - m[9] * s.mi + m[10] * s.zr - m[11] * s.zi;
  o.mi = m[6] * s.pi + m[7] * s.pr + m[8] * s.mi + m[9] * s.mr + m[10] * s.zi + m[11] * s.zr;
  o.zr = m[12] * s.pr - m[13] * s.pi + m[14] * s.mr - m[15] * s.mi + m[16] * s.zr - m[17] * s.zi;
  o.zi = m[12] * s.pi + m[13] * s.pr + m[14] * s.mi + m[15] * s.mr + m[16] * s.zi + m[17] * s.zr;
  return o;
}

// kernel parameters (by value; < 4 KB)
struct KParams {
  const epgx_op *ops;
  const epgx_segment *segs;
  const void *coef;
  const void *stream; // reg kernel: segments and records merged in one record stream (EPGX_OP_SEG markers)
  int nstream;
  const int *pats; // [npattern][EPGX_MAX_DIMS + 1]: axis strides then pool stride (reals)
  void *signal;
  void *jac;
  void *state; // ring kernel: final base state complex[atom_count][npool][C][3], or null
  long long atom_begin, atom_count;
  long long sig_stride, jac_stride; // atoms per output row (>= atom_count)
  int shape[EPGX_MAX_DIMS];
  int ndim, npattern, nseg;
  int G, A, C, nvar;
  int nvar1;        // order-1 variables among the nvar partial state sets (the rest: order-2 pairs)
  const int *tiles; // ring kernel: [gridDim.y][3] variables resident per tile (-1: empty), or null: consecutive
  const int *maps;  // ring kernel: gather maps of the lattice shifts (EPGX_SEG_LATTICE)
  int lattice;      // the tape has lattice segments: one more ring per atom as the gather's temporary
  int out_real;     // real kernel: `signal` holds rows of reals (epgx_simulate_real)
  int bounded; // the tape has segments that truncate at max_nstate (EPGX_SEG_MASK_TOP)
  unsigned init_off, m0_off;
  int init_pat, m0_pat, init_n;
};

} // namespace epgx
