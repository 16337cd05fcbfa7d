// epgx_realjac.cuh -- register kernel for real-valued phase graphs WITH order-1 partial states.
//
// Same eligibility as epgx_real.cuh (+-90 degree pulses, no precession, real initial state) extended to
// derivative tapes whose injection records are real too (dT/dalpha at phi = +-90, dE/dT1, dE/dT2, dE/dtau:
// epgpy/transition.py:172-186, evolution.py:360-389).  An atom keeps 1 + NV state sets (base + NV partial
// states, epgpy/diff.py:264-288) of three reals per order in registers, spread over G = 32 W lanes
// (W warps; order k in slot k / G of lane k % G) so that four sets of 500 orders fit: the 1000-TR FISP
// dictionary with its (B1, T1, T2) Jacobian runs with W = 4, NS = 4.  Variables beyond NV are tiled over
// blockIdx.y (the base state is recomputed per tile).  Warps of one atom exchange the two boundary values
// of every slot through shared memory at each unit shift (double-buffered, one named barrier).
#pragma once
#include <cuda_pipeline.h>

#include "epgx_common.cuh"
#include "epgx_reg.cuh"

namespace epgx {

// ---- slot-count-specialised bodies (K active slots, compile time): straight-line code, no per-slot predicate

// F+' = a F+ + b F- + u Z ; F-' = b F+ + a F- + u Z ; Z' = w Z + h (F+ + F-)
template <typename real, int NS, int SETS, int K>
__device__ __forceinline__ void rj_lin5(real (&P)[SETS][NS], real (&M)[SETS][NS], real (&Z)[SETS][NS], real a, real w, real b,
                                        real u, real h, bool on_base, bool on_part, bool inject, int iset) {
  if constexpr (K <= NS) {
    if (inject) {
#pragma unroll
      for (int q = 1; q < SETS; ++q)
        if (q == iset) { // uniform branch: one target set
#pragma unroll
          for (int s = 0; s < K; ++s) {
            const real p_ = P[0][s], m_ = M[0][s], z_ = Z[0][s];
            P[q][s] += a * p_ + b * m_ + u * z_;
            M[q][s] += a * m_ + b * p_ + u * z_;
            Z[q][s] += w * z_ + h * (p_ + m_);
          }
        }
    } else {
#pragma unroll
      for (int q = 0; q < SETS; ++q)
        if (q == 0 ? on_base : on_part) {
#pragma unroll
          for (int s = 0; s < K; ++s) {
            const real p_ = P[q][s], m_ = M[q][s], z_ = Z[q][s];
            P[q][s] = a * p_ + b * m_ + u * z_;
            M[q][s] = a * m_ + b * p_ + u * z_;
            Z[q][s] = w * z_ + h * (p_ + m_);
          }
        }
    }
  }
}

// F+- *= dp / dm, Z *= dz (the affine term is added by the caller)
template <typename real, int NS, int SETS, int K>
__device__ __forceinline__ void rj_diag3(real (&P)[SETS][NS], real (&M)[SETS][NS], real (&Z)[SETS][NS], real dp, real dm,
                                         real dz, bool on_base, bool on_part, bool inject, int iset) {
  if constexpr (K <= NS) {
    if (inject) {
#pragma unroll
      for (int q = 1; q < SETS; ++q)
        if (q == iset) {
#pragma unroll
          for (int s = 0; s < K; ++s) { P[q][s] += dp * P[0][s]; M[q][s] += dm * M[0][s]; Z[q][s] += dz * Z[0][s]; }
        }
    } else {
#pragma unroll
      for (int q = 0; q < SETS; ++q)
        if (q == 0 ? on_base : on_part) {
#pragma unroll
          for (int s = 0; s < K; ++s) { P[q][s] *= dp; M[q][s] *= dm; Z[q][s] *= dz; }
        }
    }
  }
}

#define RJ_DISPATCH(n, FN, ...)                                        \
  switch (n) {                                                         \
  case 1: FN<real, NS, SETS, 1>(__VA_ARGS__); break;                   \
  case 2: FN<real, NS, SETS, 2>(__VA_ARGS__); break;                   \
  case 3: FN<real, NS, SETS, 3>(__VA_ARGS__); break;                   \
  case 4: FN<real, NS, SETS, 4>(__VA_ARGS__); break;                   \
  case 5: FN<real, NS, SETS, 5>(__VA_ARGS__); break;                   \
  case 6: FN<real, NS, SETS, 6>(__VA_ARGS__); break;                   \
  case 7: FN<real, NS, SETS, 7>(__VA_ARGS__); break;                   \
  case 8: FN<real, NS, SETS, 8>(__VA_ARGS__); break;                   \
  default: break;                                                      \
  }

// unit shift of all state sets, K slots, one warp or less per atom (G <= 32 lanes): U moves up, D moves down, the new
// order 0 of U is the old order 1 of D
template <typename real, int NS, int SETS, int K>
__device__ __forceinline__ void rj_shift(real (&U)[SETS][NS], real (&D)[SETS][NS], int G, int gbase, int srcUp, int srcDn,
                                         bool is_first, bool is_last, bool has1) {
  const unsigned FULL = 0xffffffffu;
  if constexpr (K <= NS) {
#pragma unroll
    for (int q = 0; q < SETS; ++q) {
      real c1;
      if (G == 1) c1 = K > 1 ? D[q][K > 1 ? 1 : 0] : real(0);
      else c1 = __shfl_sync(FULL, D[q][0], gbase | 1); // used by lane 0 of the atom only
      if (!has1) c1 = real(0);
#pragma unroll
      for (int s = K - 1; s >= 0; --s) { // descending = in place
        real v = U[q][s];
        if (is_last) v = s > 0 ? U[q][s > 0 ? s - 1 : 0] : c1;
        U[q][s] = __shfl_sync(FULL, v, srcUp);
      }
      real keep = real(0);
#pragma unroll
      for (int s = K - 1; s >= 0; --s) {
        const real cur = D[q][s];
        D[q][s] = __shfl_sync(FULL, is_first ? keep : cur, srcDn);
        keep = cur;
      }
    }
  }
}

// predicated shared-memory access by 32-bit shared address (no branch, no address re-materialisation)
__device__ __forceinline__ void lds_if(double &v, unsigned addr, bool pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p ld.shared.f64 %0, [%1];\n\t}" : "+d"(v) : "r"(addr), "r"((unsigned)pred));
}
__device__ __forceinline__ void lds_if(float &v, unsigned addr, bool pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p ld.shared.f32 %0, [%1];\n\t}" : "+f"(v) : "r"(addr), "r"((unsigned)pred));
}
__device__ __forceinline__ void sts_if(unsigned addr, double v, bool pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p st.shared.f64 [%0], %1;\n\t}" ::"r"(addr), "d"(v), "r"((unsigned)pred) : "memory");
}
__device__ __forceinline__ void sts_if(unsigned addr, float v, bool pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p st.shared.f32 [%0], %1;\n\t}" ::"r"(addr), "f"(v), "r"((unsigned)pred) : "memory");
}
__device__ __forceinline__ unsigned smem_addr(const void *ptr) {
  unsigned a = (unsigned)__cvta_generic_to_shared(ptr);
  asm volatile("mov.u32 %0, %0;" : "+r"(a)); // opaque: keep it in a register
  return a;
}

// unit shift with several warps per atom (G = 32 W lanes): rotation by one lane inside each warp (shuffles); the value
// that crosses a warp boundary goes through shared memory rows [warp][set][NS + 2], written by the SENDING warp into
// the row of the receiving one: "up" rows hold the new order 0 in entry 1 and slot s of the sender's lane 31 in
// entry s + 2 (warp 0 reads from entry 1: its lane 0 takes slot s - 1 of the last warp), "down" rows hold slot s of
// the sender's lane 0 in entry s followed by a zero (the last warp reads from entry 1).  One named barrier per shift;
// the rows are double-buffered by the caller.  Arguments are shared-memory byte addresses.
template <typename real, int NS, int SETS, int K>
__device__ __forceinline__ void rj_shift_mw(real (&U)[SETS][NS], real (&D)[SETS][NS], int G, int lq, int srcUp, int srcDn, bool c1lane,
                                            bool has1, unsigned up_w, unsigned up_r, unsigned up_c1, unsigned dn_w, unsigned dn_r,
                                            int barrier_id) {
  const unsigned FULL = 0xffffffffu;
  constexpr int NSX = NS + 2, RS = (int)sizeof(real);
  if constexpr (K <= NS) {
    const bool first = lq == 0, last = lq == 31;
    if (last) { // one divergent region per sender (ptxas turns predicated stores into a branch each)
#pragma unroll
      for (int q = 0; q < SETS; ++q)
#pragma unroll
        for (int s = 0; s < K; ++s) sts_if(up_w + (q * NSX + s) * RS, U[q][s], true);
    }
    if (first) {
#pragma unroll
      for (int q = 0; q < SETS; ++q) {
#pragma unroll
        for (int s = 0; s < K; ++s) sts_if(dn_w + (q * NSX + s) * RS, D[q][s], true);
        sts_if(dn_w + (q * NSX + K) * RS, real(0), true);
      }
    }
    if (c1lane) {
#pragma unroll
      for (int q = 0; q < SETS; ++q) sts_if(up_c1 + q * NSX * RS, has1 ? D[q][0] : real(0), true);
    }
    asm volatile("bar.sync %0, %1;" ::"r"(barrier_id), "r"(G) : "memory");
#pragma unroll
    for (int q = 0; q < SETS; ++q)
#pragma unroll
      for (int s = 0; s < K; ++s) {
        real v = __shfl_sync(FULL, U[q][s], srcUp);
        lds_if(v, up_r + (q * NSX + s) * RS, first);
        U[q][s] = v;
      }
#pragma unroll
    for (int q = 0; q < SETS; ++q)
#pragma unroll
      for (int s = 0; s < K; ++s) {
        real v = __shfl_sync(FULL, D[q][s], srcDn);
        lds_if(v, dn_r + (q * NSX + s) * RS, last);
        D[q][s] = v;
      }
  }
}

// ---- whole-TR groups of derivative tapes (EPGX_OP_TRJ, see epgx_common.cuh)
//
// Every operator of such a TR lies in the five-coefficient class L(a, w, b, u, h):
//   F+' = a F+ + b F- + u Z ; F-' = b F+ + a F- + u Z ; Z' = w Z + h (F+ + F-)
// (closed under products; E and the DIAG injections are its diagonal members), so the whole TR collapses to
//   x' = F x + c_0 ,  y_v' = F y_v + J_v x + c_v          (x base state, y_v partial states, c at k = 0 only)
// with F = E_post T E_pre and J_v = F P_v + E_post T G_v E_pre + E_post Q_v T E_pre, where P_v / G_v / Q_v are the
// pre-E, pulse and post-E injection generators of variable v (epgpy/diff.py:264-288 in pre-injected form).
// The 4 x 7 coefficients of a group (a, w, b, u, h, f, z per state set; f, z per unit M0) are assembled once per
// tape window by one lane per (group, set) and kept in shared memory; the per-TR path is then straight-line FMAs.
constexpr int TRJ_PER_WINDOW = 12;
constexpr int TRJ_REALS = 32;
static_assert(TRJ_PER_WINDOW == kTrjPerWindow && TRJ_REALS == kTrjReals && TRJ_PER_WINDOW * 5 <= TAPE_CHUNK, "TRJ window layout");

template <typename real>
__device__ __forceinline__ void trj_assemble(const int4 *g, int part, const real *__restrict__ coef, const int *patoff, real *out) {
  const int4 a0 = g[0], a1 = g[1], b0 = g[2], b1 = g[3];
  const int flags = (a0.x >> 16) & 0xffff;
  const real *ct = coef + (unsigned)a0.z + patoff[a1.y & 0xff];
  const real ta = ldc(ct), tw = ldc(ct + 1), tb = ldc(ct + 2), tu = ldc(ct + 3), th = real(-0.5) * tu;
  real al1 = real(1), al2 = real(1), ra = real(0), be1 = real(1), be2 = real(1), rb = real(0);
  if (flags & EPGX_FLAG_PRE) {
    const real *c = coef + (unsigned)a0.w + patoff[(a1.y >> 8) & 0xff];
    al1 = ldc(c); ra = ldc(c + 1);
    al2 = ldc(coef + (unsigned)a1.x + patoff[(a1.y >> 16) & 0xff]);
  }
  if (flags & EPGX_FLAG_POST) {
    const real *c = coef + (unsigned)b0.z + patoff[b1.y & 0xff];
    be1 = ldc(c); rb = ldc(c + 1);
    be2 = ldc(coef + (unsigned)b0.w + patoff[(b1.y >> 8) & 0xff]);
  }
  const real Fa = be2 * al2 * ta, Fb = be2 * al2 * tb, Fu = be2 * al1 * tu, Fh = be1 * al2 * th, Fw = be1 * al1 * tw;
  if (part == 0) {
    out[0] = Fa; out[1] = Fw; out[2] = Fb; out[3] = Fu; out[4] = Fh;
    out[5] = be2 * tu * ra;
    out[6] = be1 * tw * ra + rb;
    out[7] = real(0);
    return;
  }
  const int v = part - 1;
  // block v of an injection record (two int4): off[v] / pat[v], presence in flags bit v
  auto blk = [&](const int4 lo, const int4 hi) -> const real * {
    const unsigned off = v == 0 ? (unsigned)lo.z : v == 1 ? (unsigned)lo.w : (unsigned)hi.x;
    return coef + off + patoff[(hi.y >> (8 * v)) & 0xff];
  };
  real pp = real(0), pz = real(0), p0 = real(0), ga = real(0), gw = real(0), gb = real(0), gu = real(0), qq = real(0),
       qz = real(0), q0 = real(0);
  if ((g[4].x >> (16 + v)) & 1) {
    const real *c = blk(g[4], g[5]);
    pp = ldc(c); pz = ldc(c + 4); p0 = ldc(c + 6);
  }
  if ((g[6].x >> (16 + v)) & 1) {
    const real *c = blk(g[6], g[7]);
    ga = ldc(c); gw = ldc(c + 1); gb = ldc(c + 2); gu = ldc(c + 3);
  }
  if ((g[8].x >> (16 + v)) & 1) {
    const real *c = blk(g[8], g[9]);
    qq = ldc(c); qz = ldc(c + 4); q0 = ldc(c + 6);
  }
  const real gh = real(-0.5) * gu;
  // T G_v
  const real A = ta * ga + tb * gb + tu * gh, B = ta * gb + tb * ga + tu * gh, U = (ta + tb) * gu + tu * gw;
  const real H = th * (ga + gb) + tw * gh, W = real(2) * th * gu + tw * gw;
  out[0] = Fa * (pp + qq) + be2 * al2 * A;
  out[1] = Fw * (pz + qz) + be1 * al1 * W;
  out[2] = Fb * (pp + qq) + be2 * al2 * B;
  out[3] = Fu * (pz + qq) + be2 * al1 * U;
  out[4] = Fh * (pp + qz) + be1 * al2 * H;
  out[5] = Fu * p0 + be2 * (U + qq * tu) * ra;
  out[6] = Fw * p0 + be1 * ((W + qz * tw) * ra + q0);
  out[7] = real(0);
}

// the per-TR arithmetic, K slots: 7 + 14 (SETS - 1) floating-point instructions per order
template <typename real, int NS, int SETS, int K>
__device__ __forceinline__ void rj_trj(real (&P)[SETS][NS], real (&M)[SETS][NS], real (&Z)[SETS][NS], const real *cf) {
  typedef typename vec2<real>::type real2;
  if constexpr (K <= NS) {
    const real2 f0 = ((const real2 *)cf)[0], f1 = ((const real2 *)cf)[1], f2 = ((const real2 *)cf)[2];
    const real a = f0.x, w = f0.y, b = f1.x, u = f1.y, h = f2.x, c = a - b;
#pragma unroll
    for (int q = 1; q < SETS; ++q) {
      const real2 j0 = ((const real2 *)cf)[4 * q], j1 = ((const real2 *)cf)[4 * q + 1], j2 = ((const real2 *)cf)[4 * q + 2];
      const real ja = j0.x, jw = j0.y, jb = j1.x, ju = j1.y, jh = j2.x;
#pragma unroll
      for (int s = 0; s < K; ++s) {
        const real xp = P[0][s], xm = M[0][s], xz = Z[0][s];
        const real p_ = P[q][s], m_ = M[q][s], z_ = Z[q][s];
        // shared by the rows of F+ and F- (epgx_real.cuh): q = b s + u Z + jb xs + ju x_Z -- 14 instead of 18 per order
        const real s_ = p_ + m_, xs = xp + xm;
        const real q_ = fma(b, s_, fma(u, z_, fma(jb, xs, ju * xz)));
        P[q][s] = fma(c, p_, fma(ja - jb, xp, q_));
        M[q][s] = fma(c, m_, fma(ja - jb, xm, q_));
        Z[q][s] = fma(w, z_, fma(h, s_, fma(jw, xz, jh * xs)));
      }
    }
#pragma unroll
    for (int s = 0; s < K; ++s) {
      const real p_ = P[0][s], m_ = M[0][s], z_ = Z[0][s];
      const real s_ = p_ + m_, q_ = fma(b, s_, u * z_);
      P[0][s] = fma(c, p_, q_);
      M[0][s] = fma(c, m_, q_);
      Z[0][s] = fma(w, z_, h * s_);
    }
  }
}

// per-thread constants of the whole-TR fast loop
template <typename real> struct RjCtx {
  int G, lgG, lane, lq, gbase, srcUp, srcDn, barrier_id, nvar;
  bool is_first, is_last, valid;
  unsigned up_w, up_r, up_c1, dn_w, dn_r, par_bytes; // shared-memory byte addresses (rj_shift_mw)
  typename vec2<real>::type *sig, *jac;               // output rows of this atom
  long long sig_stride, jac_stride;
};

// consecutive plain whole-TR groups (unit shift +1, no segment flags) of the current tape window, executed with K
// slots: the slot count is a compile-time constant of the LOOP, not of each TR, so the state never leaves its
// registers between TRs.  K covers orders 0..min(n_new, nact + 1): orders above nact + 1 are unobservable
// (lowering.py) and need not move.  On entry r is the first record of a group that qualifies; on exit it is the
// last record of the last group done.
template <typename real, int NS, int SETS, bool MW, int K>
__device__ __forceinline__ void rj_trj_run(real (&P)[SETS][NS], real (&M)[SETS][NS], real (&Z)[SETS][NS], const RjCtx<real> &c,
                                           const int4 *tb, const real *trjc, real m0, int cnt, int &r, int &nact, int &parity) {
  typedef typename vec2<real>::type real2;
  if constexpr (K <= NS) {
    for (;;) {
      const int4 h0 = tb[2 * r], q0 = tb[2 * r + 2], q1 = tb[2 * r + 3];
      const real *cf = trjc + (r / 5) * (SETS * 8);
      rj_trj<real, NS, SETS, K>(P, M, Z, cf);
      if (c.lane == 0) {
#pragma unroll
        for (int q = 0; q < SETS; ++q) {
          const real f = cf[8 * q + 5] * m0;
          P[q][0] += f; M[q][0] += f; Z[q][0] += cf[8 * q + 6] * m0;
        }
        if (c.valid) {
          c.sig[(long long)q0.y * c.sig_stride] = real2{P[0][0], real(0)};
          if (h0.x & (EPGX_FLAG_PARTIALS << 16)) {
#pragma unroll
            for (int q = 1; q < SETS; ++q)
              if (q - 1 < c.nvar) c.jac[((long long)q1.w * c.nvar + q - 1) * c.jac_stride] = real2{P[q][0], real(0)};
          }
        }
      }
      const int n_old = (int)((unsigned)q1.x >> 16);
      if constexpr (MW) {
        const unsigned po = parity * c.par_bytes;
        rj_shift_mw<real, NS, SETS, K>(P, M, c.G, c.lq, c.srcUp, c.srcDn, c.lane == 1, n_old >= 1, c.up_w + po, c.up_r + po,
                                       c.up_c1 + po, c.dn_w + po, c.dn_r + po, c.barrier_id);
        parity ^= 1;
      } else {
        rj_shift<real, NS, SETS, K>(P, M, c.G, c.gbase, c.srcUp, c.srcDn, c.is_first, c.is_last, n_old >= 1);
      }
      nact = q1.z;
      r += 5;
      if (r + 5 > cnt) break;
      const int4 g0 = tb[2 * r], g2 = tb[2 * r + 2], g3 = tb[2 * r + 3];
      if ((g0.x & 0xffff) != EPGX_OP_TRJ || ((g2.x >> 16) & 0xffff) != 2 || nact < 0) break;
      if (((min((int)((unsigned)g3.x & 0xffff), nact + 1) >> c.lgG) + 1) != K) break;
    }
    r -= 1;
  }
}

// MAXT = 128: instance for CTAs of at most 128 threads, capped at 168 registers (three CTAs per SM: +30 % on the FISP
// Jacobian in FP64 over the 192-register build)
template <typename real, int NS, int NV, bool MW, int MAXT = 256>
__global__ void __launch_bounds__(MAXT, MAXT <= 128 ? 3 : 1) realjac_kernel(const KParams p) {
  typedef typename vec2<real>::type real2;
  constexpr int SETS = 1 + NV;
  extern __shared__ __align__(16) unsigned char smem_raw[];

  const int G = p.G;
  const int tid = threadIdx.x;
  const int al = tid / G;
  const int lane = tid - al * G;
  const int lw = tid & 31;
  const int W = MW ? G >> 5 : 1; // MW: several warps per atom (G = 32 W), else one warp or less
  const int GW = MW ? 32 : G;
  const int wq = MW ? lane >> 5 : 0;
  const int gbase = lw & ~(GW - 1);
  const int lq = lw - gbase;
  const int srcUp = gbase | ((lq - 1) & (GW - 1));
  const int srcDn = gbase | ((lq + 1) & (GW - 1));
  const bool is_first = lq == 0, is_last = lq == GW - 1;
  int lgG = 0;
  while ((1 << lgG) < G) ++lgG;

  const long long a_rel = (long long)blockIdx.x * p.A + al;
  const bool valid = a_rel < p.atom_count;
  const long long atom = p.atom_begin + (valid ? a_rel : p.atom_count - 1);
  const int v0 = blockIdx.y * NV;
  const real *__restrict__ coef = (const real *)p.coef;

  // shared memory: tape window, pattern offsets [A][npattern], boundary exchange [2][A][up | down][W][SETS][NS + 2],
  // TRJ coefficients [A][TRJ_PER_WINDOW][SETS][8]
  int4 *tbuf = (int4 *)smem_raw;
  int *patoff = (int *)(tbuf + 2 * TAPE_CHUNK * 2) + al * p.npattern;
  real *xbuf = (real *)((int *)(tbuf + 2 * TAPE_CHUNK * 2) + ((p.A * p.npattern + 3) & ~3));
  constexpr int NSX = NS + 2, ROW = SETS * NSX; // boundary exchange rows (rj_shift_mw)
  const int par_stride = p.A * W * 2 * ROW;
  real *trjc = xbuf + (MW ? (size_t)2 * par_stride : 0) + (size_t)al * TRJ_PER_WINDOW * TRJ_REALS;
  real *const xup = xbuf + (size_t)al * W * 2 * ROW, *const xdn = xup + W * ROW;
  const unsigned up_w = smem_addr(xup + (wq + 1 == W ? 0 : wq + 1) * ROW + 2), up_c1 = smem_addr(xup + 1);
  const unsigned up_r = smem_addr(xup + wq * ROW + (wq == 0 ? 1 : 2));
  const unsigned dn_w = smem_addr(xdn + (wq == 0 ? W - 1 : wq - 1) * ROW);
  const unsigned dn_r = smem_addr(xdn + wq * ROW + (wq == W - 1 ? 1 : 0));
  RjCtx<real> ctx;
  ctx.G = G; ctx.lgG = lgG; ctx.lane = lane; ctx.lq = lq; ctx.gbase = gbase; ctx.srcUp = srcUp; ctx.srcDn = srcDn;
  ctx.barrier_id = 1 + al; ctx.nvar = p.nvar;
  ctx.is_first = is_first; ctx.is_last = is_last; ctx.valid = valid;
  ctx.up_w = up_w; ctx.up_r = up_r; ctx.up_c1 = up_c1; ctx.dn_w = dn_w; ctx.dn_r = dn_r;
  ctx.par_bytes = (unsigned)par_stride * (unsigned)sizeof(real);
  ctx.sig = (typename vec2<real>::type *)p.signal + a_rel; ctx.jac = (typename vec2<real>::type *)p.jac + a_rel;
  ctx.sig_stride = p.sig_stride; ctx.jac_stride = p.jac_stride;
  {
    int idx[EPGX_MAX_DIMS];
    long long r = atom;
    for (int d = p.ndim - 1; d >= 0; --d) {
      idx[d] = (int)(r % p.shape[d]);
      r /= p.shape[d];
    }
    for (int q = lane; q < p.npattern; q += G) {
      const int *st = p.pats + q * (EPGX_MAX_DIMS + 1);
      int o = 0;
      for (int d = 0; d < p.ndim; ++d) o += idx[d] * st[d];
      patoff[q] = o;
    }
  }
  __syncthreads();

  real P[SETS][NS], M[SETS][NS], Z[SETS][NS];
  real m0 = ldc(coef + p.m0_off + patoff[p.m0_pat]);
  {
    const real *ib = coef + p.init_off + patoff[p.init_pat];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const int k = s * G + lane;
      const bool in = k <= p.init_n;
      P[0][s] = in ? ldc(ib + 6 * k) : real(0);
      M[0][s] = in ? ldc(ib + 6 * k + 2) : real(0);
      Z[0][s] = in ? ldc(ib + 6 * k + 4) : real(0);
#pragma unroll
      for (int q = 1; q < SETS; ++q) P[q][s] = M[q][s] = Z[q][s] = real(0);
    }
  }
  real2 *sig = (real2 *)p.signal;
  real2 *jac = (real2 *)p.jac;
  int parity = 0;

  const int4 *stream = (const int4 *)p.stream;
  const int nthreads = blockDim.x;
  for (int i = tid; i < 2 * TAPE_CHUNK && i < 2 * p.nstream; i += nthreads) __pipeline_memcpy_async(tbuf + i, stream + i, 16);
  __pipeline_commit();
  int nact = -1, nslot = 0;
  bool alive = false; // has a derivative been injected into this tile's partial states yet?
  for (int base = 0, chunk = 0; base < p.nstream; base += TAPE_CHUNK, ++chunk) {
    __pipeline_wait_prior(0);
    __syncthreads();
    {
      const int nb = base + TAPE_CHUNK;
      int4 *dst = tbuf + ((chunk + 1) & 1) * 2 * TAPE_CHUNK;
      for (int i = tid; i < 2 * TAPE_CHUNK && nb * 2 + i < 2 * p.nstream; i += nthreads)
        __pipeline_memcpy_async(dst + i, stream + (size_t)nb * 2 + i, 16);
      __pipeline_commit();
    }
    const int4 *tb = tbuf + (chunk & 1) * 2 * TAPE_CHUNK;
    const int cnt = min(TAPE_CHUNK, p.nstream - base);
    if (tb[0].x & EPGX_CHUNK_ANY_TRJ) { // block-uniform: coefficient assembly of the window's TRJ groups
      for (int item = lane; item < TRJ_PER_WINDOW * SETS; item += G) {
        const int j = item / SETS, part = item - j * SETS;
        if (5 * j + 5 <= cnt && (tb[10 * j].x & 0xffff) == EPGX_OP_TRJ)
          trj_assemble<real>(tb + 10 * j, part, coef, patoff, trjc + (j * SETS + part) * 8);
      }
      __syncthreads();
    }
    for (int r = 0; r < cnt; ++r) {
      const int4 r0 = tb[2 * r], r1 = tb[2 * r + 1];
      const int code = r0.x & 0xffff, flags = (r0.x >> 16) & 0xffff, aux = r0.y;
      const unsigned off0 = (unsigned)r0.z, off1 = (unsigned)r0.w, off2 = (unsigned)r1.x;
      const int pat0 = r1.y & 0xff, pat1 = (r1.y >> 8) & 0xff;
      const bool inject = flags & EPGX_FLAG_INJECT;
      const int iset = aux - v0 + 1; // target set of an injection
      // the partial states of this variable tile are exactly zero until its first injection (per-pulse variables: most
      // of the sequence for the late tiles): linear operators leave them zero, so they are skipped
      if (inject && iset >= 1 && iset < SETS) alive = true;
      const bool on_base = flags & EPGX_FLAG_BASE, on_part = (flags & EPGX_FLAG_PARTIALS) && alive;
      const bool aff0 = (flags & EPGX_FLAG_AFFINE) && lane == 0 && nslot > 0;
      (void)off2;

#define LIN5(a, w, b, u, h) RJ_DISPATCH(nslot, rj_lin5, P, M, Z, a, w, b, u, h, on_base, on_part, inject, iset)
#define DIAG3(dp, dm, dz, z0)                                                                  \
  {                                                                                            \
    RJ_DISPATCH(nslot, rj_diag3, P, M, Z, dp, dm, dz, on_base, on_part, inject, iset)          \
    if (aff0) {                                                                                \
      if (inject) {                                                                            \
        _Pragma("unroll") for (int q = 1; q < SETS; ++q) if (q == iset) Z[q][0] += z0;         \
      } else if (on_base) Z[0][0] += z0;                                                       \
    }                                                                                          \
  }

#define DO_SEG(SHIFT, NOLD, NNEW, SFLAGS, NEXT_NACT)                                                                  \
  {                                                                                                                    \
    const int shift = (SHIFT), n_old = (NOLD), n_new = (NNEW), sflags = (SFLAGS);                                      \
    const int n_move = max(min(n_new, nact + 1), 0); /* orders above nact + 1 are unobservable: they stay */           \
    nact = (NEXT_NACT);                                                                                                \
    nslot = nact < 0 ? 0 : (nact >> lgG) + 1;                                                                          \
    if (sflags & EPGX_SEG_RESET) {                                                                                     \
      _Pragma("unroll") for (int q = 0; q < SETS; ++q)                                                                 \
        _Pragma("unroll") for (int s = 0; s < NS; ++s) P[q][s] = M[q][s] = Z[q][s] = real(0);                          \
      if (lane == 0) Z[0][0] = m0;                                                                                     \
    } else if (shift != 0) {                                                                                           \
      const int nsl = (n_move >> lgG) + 1;                                                                             \
      if constexpr (MW) {                                                                                              \
        const unsigned po = parity * par_stride * (int)sizeof(real);                                                                           \
        if (shift > 0) { RJ_DISPATCH(nsl, rj_shift_mw, P, M, G, lq, srcUp, srcDn, lane == 1, n_old >= 1, up_w + po, up_r + po, up_c1 + po, dn_w + po, dn_r + po, 1 + al) } \
        else { RJ_DISPATCH(nsl, rj_shift_mw, M, P, G, lq, srcUp, srcDn, lane == 1, n_old >= 1, up_w + po, up_r + po, up_c1 + po, dn_w + po, dn_r + po, 1 + al) } \
        parity ^= 1;                                                                                                   \
      } else {                                                                                                         \
        if (shift > 0) { RJ_DISPATCH(nsl, rj_shift, P, M, G, gbase, srcUp, srcDn, is_first, is_last, n_old >= 1) } \
        else { RJ_DISPATCH(nsl, rj_shift, M, P, G, gbase, srcUp, srcDn, is_first, is_last, n_old >= 1) } \
      }                                                                                                                \
      if (sflags & EPGX_SEG_MASK_TOP) {                                                                                \
        _Pragma("unroll") for (int q = 0; q < SETS; ++q)                                                               \
          _Pragma("unroll") for (int s = 0; s < NS; ++s)                                                               \
            if (s * G + lane > n_new) {                                                                                \
              if (shift > 0) P[q][s] = real(0); else M[q][s] = real(0);                                                \
            }                                                                                                          \
      }                                                                                                                \
    }                                                                                                                  \
  }

      switch (code) {
      case EPGX_OP_T_RE: {
        const real *c = coef + off0 + patoff[pat0];
        const real a = ldc(c), w = ldc(c + 1), b = ldc(c + 2), u = ldc(c + 3), h = real(-0.5) * u;
        LIN5(a, w, b, u, h)
      } break;
      case EPGX_OP_E: {
        const real *c0 = coef + off0 + patoff[pat0];
        const real e1 = ldc(c0), z0 = ldc(c0 + 1) * m0;
        const real e2 = ldc(coef + off1 + patoff[pat1]);
        DIAG3(e2, e2, e1, z0)
      } break;
      case EPGX_OP_DIAG: { // real entries only (checked by the host): (aP, 0, aM, 0, aZ, 0, a0, 0)
        const real *c = coef + off0 + patoff[pat0];
        const real dp = ldc(c), dm = ldc(c + 2), dz = ldc(c + 4), z0 = ldc(c + 6) * m0;
        DIAG3(dp, dm, dz, z0)
      } break;
      case EPGX_OP_D: {
        const real *c = coef + off0 + patoff[pat0];
#pragma unroll
        for (int s = 0; s < NS; ++s)
          if (s < nslot) {
            const int k = min(s * G + lane, p.C - 1);
            const real dp = ldc(c + 3 * k), dm = ldc(c + 3 * k + 1), dl = ldc(c + 3 * k + 2);
#pragma unroll
            for (int q = 0; q < SETS; ++q)
              if (q == 0 ? on_base : on_part) { P[q][s] *= dp; M[q][s] *= dm; Z[q][s] *= dl; }
          }
      } break;
      case EPGX_OP_SPOIL:
#pragma unroll
        for (int q = 0; q < SETS; ++q)
          if (q == 0 ? on_base : on_part) {
#pragma unroll
            for (int s = 0; s < NS; ++s)
              if (s < nslot) { P[q][s] = real(0); M[q][s] = real(0); }
          }
        break;
      case EPGX_OP_PD:
        m0 = ldc(coef + off0 + patoff[pat0]);
        break;
      case EPGX_OP_ADC:
        if (lane == 0 && valid) {
          real fr = real(1), fi = real(0);
          if (flags & EPGX_FLAG_SCALE) {
            const real *c = coef + off0 + patoff[pat0];
            fr = ldc(c); fi = ldc(c + 1);
          }
          const bool z0 = flags & EPGX_FLAG_Z0;
          if (on_base && blockIdx.y == 0) {
            const real x = z0 ? Z[0][0] : P[0][0];
            sig[(long long)aux * p.sig_stride + a_rel] = real2{x * fr, x * fi};
          }
          if (flags & EPGX_FLAG_PARTIALS) {
#pragma unroll
            for (int q = 1; q < SETS; ++q) {
              const int v = v0 + q - 1;
              if (v < p.nvar) {
                const real x = z0 ? Z[q][0] : P[q][0];
                jac[((long long)r1.z * p.nvar + v) * p.jac_stride + a_rel] = real2{x * fr, x * fi};
              }
            }
          }
        }
        break;
      case EPGX_OP_SEG:
        DO_SEG((int)off0, (int)off1, (int)off2, r1.z, aux)
        break;
      case EPGX_OP_TRJ: { // one whole TR with its injections, ADC and the segment's close (coefficients: trjc)
        if (((tb[2 * r + 2].x >> 16) & 0xffff) == 2 && nact >= 0 && v0 == 0) { // plain TRs: per-slot-count loops
          const int ks = (min((int)((unsigned)tb[2 * r + 3].x & 0xffff), nact + 1) >> lgG) + 1;
#define RUN(K) case K: rj_trj_run<real, NS, SETS, MW, K>(P, M, Z, ctx, tb, trjc, m0, cnt, r, nact, parity); break;
          switch (ks) { RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) default: break; }
#undef RUN
          if (ks <= NS) {
            nslot = nact < 0 ? 0 : (nact >> lgG) + 1;
            break;
          }
        }
        const real *cf = trjc + (r / 5) * (SETS * 8);
        RJ_DISPATCH(nslot, rj_trj, P, M, Z, cf)
        if (lane == 0 && nslot > 0) {
#pragma unroll
          for (int q = 0; q < SETS; ++q) {
            const real f = cf[8 * q + 5] * m0;
            P[q][0] += f; M[q][0] += f; Z[q][0] += cf[8 * q + 6] * m0;
          }
        }
        const int4 q0 = tb[2 * r + 2], q1 = tb[2 * r + 3];
        if (lane == 0 && valid) {
          if (blockIdx.y == 0) sig[(long long)q0.y * p.sig_stride + a_rel] = real2{P[0][0], real(0)};
          if (flags & EPGX_FLAG_PARTIALS) {
#pragma unroll
            for (int q = 1; q < SETS; ++q) {
              const int v = v0 + q - 1;
              if (v < p.nvar) jac[((long long)q1.w * p.nvar + v) * p.jac_stride + a_rel] = real2{P[q][0], real(0)};
            }
          }
        }
        const int segw = (q0.x >> 16) & 0xffff; // (shift + 1) | segment flags << 2
        DO_SEG((segw & 3) - 1, (int)((unsigned)q1.x >> 16), (int)((unsigned)q1.x & 0xffff), segw >> 2, q1.z)
        r += 4;
      } break;
      default:
        break;
      }
#undef LIN5
#undef DIAG3
#undef DO_SEG
    }
  }
}

} // namespace epgx
