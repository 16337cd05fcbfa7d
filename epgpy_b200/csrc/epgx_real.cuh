// epgx_real.cuh -- register kernel for REAL-VALUED phase graphs.
//
// When every RF pulse of a sequence has a phase of +-90 degrees (B and U real: EPGX_OP_T_RE / the RE
// kind of EPGX_OP_FUSED), no operator precesses (E without g, D, SPOILER, PD) and the initial state is
// real, the imaginary part of every F+/F-/Z coefficient stays exactly zero for the whole sequence (the
// reference computes them anyway and returns ~1e-17 round-off, epgpy/transition.py:114-151).  The MRF
// FISP / phase-alternated bSSFP families are of this kind.  This kernel keeps three reals per order
// instead of six: half the FMAs, half the shuffles of the unit shift, half the registers (so twice the
// resident warps) of epgx_reg.cuh.  One warp (or a sub-warp group of G lanes) per atom.
//
// Layout: orders are dealt to the lanes in BLOCKS OF TWO -- order k = 2 b + i lives in lane b % G, register
// slot 2 (b / G) + i.  A unit shift moves F+(k - 1) to F+(k): inside a block that is a change of register
// (free), only the value that leaves a block crosses to the next lane.  In the whole-TR fast path the register
// roles alternate with the parity of the TR (tr_window), so a shift costs ONE shuffled register per block
// and component instead of two; the generic record path keeps the canonical roles and copies.
#pragma once
#include <cuda_pipeline.h>

#include "epgx_common.cuh"
#include "epgx_reg.cuh"

// Compile-time switches of measured experiments.  The rejected variants stay in the source on
// purpose: ptxas' register assignment in the whole-TR loops is sensitive to the surrounding code -- removing the unused
// per-lane affine pointer (EPGX_REAL_AFF) changed the loop bodies and cost 2 % in FP32 (120.7 against 118.2 ms).
#ifndef EPGX_REAL_UNROLL
#define EPGX_REAL_UNROLL 1 // iterations (of two TRs) unrolled in the whole-TR loop; 2: measured 10 % slower
#endif
#ifndef EPGX_REAL_AFF
#define EPGX_REAL_AFF 0 // affine terms of order 0 through a per-lane pointer (1: measured 2 % slower) or selects (0)
#endif
#ifndef EPGX_REAL_SCALED
#define EPGX_REAL_SCALED 1 // 0: every whole-TR window runs the unscaled seven-instruction form (experiments)
#endif

namespace epgx {

// 1 / x to full precision without the division subroutine (|x| normal: the caller checks the range)
__device__ __forceinline__ double rcp_newton(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(fma(-x, r, 1.0), r, r);
  return fma(fma(-x, r, 1.0), r, r);
}
__device__ __forceinline__ float rcp_newton(float x) { return __frcp_rn(x); }

// One tape window of TAPE_CHUNK / 2 whole-TR records (fused E.T.E, plain ADC, unit shift +1) for one atom
// per warp, with a COMPILE-TIME number KP of active register pairs: the per-TR body is a single basic block (no
// slot dispatch), so the scheduler overlaps the coefficient loads, the 18 KP FMAs and the 2 KP shuffles /
// selects of the shift.  KP is the largest pair count of the window; registers above the populated orders
// hold zeros (or unobservable values), so over-covering changes nothing.  The fused coefficients of the
// window's TRs wait in the warp's shared-memory rows cw[TR][8] = (c, w | b, u | h, mask | fz, zz) (written by lane TR);
// lane 0 leaves the echo of TR j in sb[j]; mask != 0 flags a shift that truncates at max_nstate.  The affine terms
// (fz, zz) belong to order 0 only: lane 0 reads them from the row, every other lane from a row of zeros (af / afs:
// per-lane base and stride -- one load instead of four selects per TR).
//
// Two TRs per iteration.  Even TR ("phase 0"): the even order of a block is in register 2 sp, the odd one in
// 2 sp + 1 (canonical).  Its shift rotates the ODD F+ registers up one lane (they become the even orders of
// the next block) and the EVEN F- registers down one lane (they become the odd orders of the previous block);
// the other registers stay where they are and change role.  Odd TR ("phase 1"): roles swapped, its shift
// rotates the even F+ and the odd F- registers and restores the canonical roles.  Z never moves.
//
// Arithmetic of one TR on one order (the fused E.T.E is the real matrix [[a b u] [b a u] [h h w]]):
//   s = F+ + F-,  q = b s + u Z,  F+' = c F+ + q,  F-' = c F- + q,  Z' = w Z + h s      with c = a - b
// -- SEVEN floating-point instructions (the row-by-row form takes eight: q is shared by F+ and F-, s by q and Z').
// SC (scaled window): the lanes keep Zs = u_j Z instead of Z while the window runs (u_j: the coefficient of the TR
// about to be applied), which removes the product u Z as well -- SIX instructions:
//   q = b s + Zs,  Zs' = w~ Zs + h~ s,   w~ = u_(j+1) w / u_j,  h~ = u_(j+1) h,  zz~ = u_(j+1) zz   (staged by the prologue)
// The last TR of a window takes u_(j+1) := 1 and leaves plain Z; a window with a vanishing u (a pulse of 0 or 180
// degrees) runs unscaled.  Only Z is rescaled: rescaling F as well would remove h, but the two scales then grow
// like prod 1 / (h u) -- not bounded over a window in FP32.
template <typename real, int NS, int KP, bool MASK, bool SC>
__device__ __forceinline__ void tr_window(real (&P)[NS], real (&M)[NS], real (&Z)[NS], const real *cw, real *sb, bool lane0,
                                          bool is_first, bool is_last, int srcUp, int srcDn, unsigned mtop, int jbeg, int jend,
                                          const real *af, int afs) {
  typedef typename vec2<real>::type real2;
  const unsigned FULL = 0xffffffffu;
  if constexpr (2 * KP <= NS) {
    if (SC && jbeg == 0) { // entering a scaled window: Zs = u_0 Z (registers above the pair count hold zeros or unobservable orders)
      const real u0 = cw[3];
#pragma unroll
      for (int r = 0; r < 2 * KP; ++r) Z[r] *= u0;
    }
    constexpr int kUnroll = EPGX_REAL_UNROLL;
#pragma unroll(kUnroll)
    for (int j = jbeg; j < jend; j += 2) {
#pragma unroll
      for (int ph = 0; ph < 2; ++ph) {
        // coefficients of the TR: broadcast loads from the warp's staging rows
        const real2 c0 = ((const real2 *)cw)[4 * (j + ph)], c1v = ((const real2 *)cw)[4 * (j + ph) + 1],
                    c2 = ((const real2 *)cw)[4 * (j + ph) + 2];
        const real c = c0.x, w = c0.y, b = c1v.x, u = c1v.y, h = c2.x;
        if constexpr (sizeof(real) == 4) {
          // FP32: the two orders of a block sit in adjacent registers and take the same coefficients -- Blackwell's
          // packed FFMA2 / FMUL2 / FADD2 (sm_100: fma.rn.f32x2) update both at once, 7 (6 scaled) instructions per register
          // pair; the scalar coefficients are broadcast operands and the role swap of the odd TRs is the
          // LO_HI operand swizzle on the Z pair (both free).  The kernel is issue-bound in FP32.
#pragma unroll
          for (int sp = 0; sp < KP; ++sp) {
            const float2 p2 = make_float2(P[2 * sp], P[2 * sp + 1]), m2 = make_float2(M[2 * sp], M[2 * sp + 1]);
            const float2 z2 = ph ? make_float2(Z[2 * sp + 1], Z[2 * sp]) : make_float2(Z[2 * sp], Z[2 * sp + 1]);
            const float2 cc2 = make_float2(c, c), b2 = make_float2(b, b), u2 = make_float2(u, u), w2 = make_float2(w, w),
                         h2 = make_float2(h, h);
            const float2 s2 = __fadd2_rn(p2, m2);
            const float2 q2 = SC ? __ffma2_rn(b2, s2, z2) : __ffma2_rn(b2, s2, __fmul2_rn(u2, z2));
            const float2 np = __ffma2_rn(cc2, p2, q2);
            const float2 nm = __ffma2_rn(cc2, m2, q2);
            const float2 nz = __ffma2_rn(w2, z2, __fmul2_rn(h2, s2));
            P[2 * sp] = np.x; P[2 * sp + 1] = np.y; M[2 * sp] = nm.x; M[2 * sp + 1] = nm.y;
            if (ph) { Z[2 * sp + 1] = nz.x; Z[2 * sp] = nz.y; } else { Z[2 * sp] = nz.x; Z[2 * sp + 1] = nz.y; }
          }
        } else {
#if EPGX_REAL_AFF
          const real2 aff = *(const real2 *)(af + afs * (j + ph));
          const real fz0 = aff.x, zz0 = aff.y;
#else
          const real fz0 = lane0 ? cw[8 * (j + ph) + 6] : real(0), zz0 = lane0 ? cw[8 * (j + ph) + 7] : real(0);
#endif
#pragma unroll
          for (int sp = 0; sp < KP; ++sp)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int r = 2 * sp + (i ^ ph); // register of F+- (k = 2 b + i) in this phase; Z(k) is in 2 sp + i
              const real p_ = P[r], m_ = M[r], z_ = Z[2 * sp + i];
              // seven (scaled window: six) FP64 instructions per order, written out with fma().  Order 0 (pair 0,
              // i = 0, lane 0) takes the affine terms of the two E operators as the addends of its first products,
              // which costs two selects instead of three FP64 additions for the whole warp
              const real s_ = p_ + m_;
              if (sp == 0 && i == 0) {
                const real q_ = fma(b, s_, SC ? z_ + fz0 : fma(u, z_, fz0));
                P[r] = fma(c, p_, q_);
                M[r] = fma(c, m_, q_);
                Z[0] = fma(w, z_, fma(h, s_, zz0));
              } else {
                const real q_ = SC ? fma(b, s_, z_) : fma(b, s_, u * z_);
                P[r] = fma(c, p_, q_);
                M[r] = fma(c, m_, q_);
                Z[2 * sp + i] = fma(w, z_, h * s_);
              }
            }
          if (lane0) sb[j + ph] = P[ph]; // the echo of the TR; written to HBM by lane j + ph after the window
        }
        if constexpr (sizeof(real) == 4) {
          if (lane0) {
            const real fz = cw[8 * (j + ph) + 6], zz = cw[8 * (j + ph) + 7];
            P[ph] += fz; M[ph] += fz; Z[0] += zz;
            sb[j + ph] = P[ph];
          }
        }
        // unit shift +1.  F+: registers of the odd orders (2 sp + 1 - ph) rotate up one lane, the last lane sending
        // the value of its previous pair; order 0 <- F-(1), which lane 0 holds itself.  F-: registers of the even
        // orders (2 sp + ph) rotate down one lane, the first lane sending the value of its next pair.
        const real f1 = M[1 - ph];
#pragma unroll
        for (int sp = KP - 1; sp >= 0; --sp) {
          const int r = 2 * sp + 1 - ph;
          const real v = (is_last && sp > 0) ? P[sp > 0 ? r - 2 : r] : P[r];
          P[r] = __shfl_sync(FULL, v, srcUp);
        }
        if (is_first) P[1 - ph] = f1;
        real keep = real(0);
#pragma unroll
        for (int sp = KP - 1; sp >= 0; --sp) {
          const int r = 2 * sp + ph;
          const real cur = M[r];
          M[r] = __shfl_sync(FULL, is_first ? keep : cur, srcDn);
          keep = cur;
        }
        // truncation at max_nstate (EPGX_SEG_MASK_TOP, flag staged with the coefficients): F+ of the order that moved
        // above the cap reads as zero.  mtop: bit of its canonical register in this lane (0: another lane / not held);
        // in the roles after this shift the register is (canonical ^ (1 - ph))
        if constexpr (MASK) {
          if (c2.y != real(0)) { // (bit tests: an index comparison would turn P[] into a local array)
#pragma unroll
            for (int r = 0; r < 2 * KP; ++r) P[r] = ((mtop >> (r ^ (1 - ph))) & 1u) ? real(0) : P[r];
          }
        }
      }
    }
  }
}

// register budget: CTAs of at most 128 threads (MAXT = 128, the default launch shape: four warps) get 168 registers
// when the state takes 96 of them (FP64 NS = 16: three CTAs per SM; measured 246 ms against 282 ms at 128 registers
// and 258 ms uncapped for the 1 M-atom FISP dictionary) and 128 registers for the smaller states
constexpr int real_min_blocks(int state_bytes, int maxt) {
  return maxt <= 128 ? (state_bytes <= 64 ? 4 : state_bytes <= 128 ? 3 : 2) : (state_bytes <= 64 ? 3 : state_bytes <= 128 ? 2 : 1);
}

// BOUNDED: the tape truncates at max_nstate (EPGX_SEG_MASK_TOP segments) -- only these instances carry the masking
// code of the whole-TR windows (it costs the unbounded FISP dictionary 10 % through register pressure otherwise)
template <typename real, int NS, int MAXT = 256, bool BOUNDED = false>
__global__ void __launch_bounds__(MAXT, real_min_blocks(sizeof(real) * NS, MAXT)) real_kernel(const KParams p) {
  typedef typename vec2<real>::type real2;
  extern __shared__ __align__(16) unsigned char smem_raw[];

  const int G = p.G; // 1..32, power of two
  const int tid = threadIdx.x;
  const int al = tid / G;
  const int lane = tid - al * G;
  const int lw = tid & 31;
  const int gbase = lw & ~(G - 1);
  const int srcUp = gbase | ((lane - 1) & (G - 1));
  const int srcDn = gbase | ((lane + 1) & (G - 1));
  const unsigned FULL = 0xffffffffu;
  const bool is_last = lane == G - 1, is_first = lane == 0;
  int lgG = 0;
  while ((1 << lgG) < G) ++lgG;

  const long long a_rel = (long long)blockIdx.x * p.A + al;
  const bool valid = a_rel < p.atom_count;
  const long long atom = p.atom_begin + (valid ? a_rel : p.atom_count - 1);
  const real *__restrict__ coef = (const real *)p.coef;

  // The kernel steps through the tape KW host windows at a time (kRealWindows = 2: 128 records, double-buffered): two
  // consecutive whole-TR windows run as ONE window of 64 TRs -- one prologue (decode, gathers, fusing, staging: its
  // latencies took 18 % of the kernel with 32-TR windows) for twice the TRs.
  // (one warp per atom only: with several atoms per warp the staging rows of a 64-TR window cost CTAs -- measured
  // FP32 122.3 -> 118.8 ms, FP64 unchanged, but max_nstate = 32 at 8 lanes per atom 61.4 -> 65.6 ms when joined)
  constexpr int KW = kRealWindows, RCH = KW * TAPE_CHUNK;
  const int WTR = (G == 32 ? KW : 1) * (TAPE_CHUNK / 2); // rows of the staging buffers: windows are joined for G = 32 only
  int4 *tbuf = (int4 *)smem_raw;
  int *patoff = (int *)(tbuf + 2 * RCH * 2) + al * p.npattern;
  // whole-TR windows (G >= 8): per atom, the echoes of the window [KW * 32] then the coefficient rows [WTR][8] (the echo
  // buffer first: at a compile-time distance from the rows)
  constexpr int SBN = KW * (TAPE_CHUNK / 2);
  real *cw0 = (real *)((int *)(tbuf + 2 * RCH * 2) + ((p.A * p.npattern + 3) & ~3));
  real *sb = cw0 + (size_t)al * (SBN + WTR * 8);
  real *cw = sb + SBN;
  real *zrow = cw0 + (G >= 8 ? (size_t)p.A * (SBN + WTR * 8) : 0); // two zeros: the affine terms of the lanes that do not hold order 0
  if (tid < 4) zrow[tid] = real(0);
  const real *af = lane == 0 ? cw + 6 : zrow;
  const int afs = lane == 0 ? 8 : 0;
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

  static_assert(NS % 2 == 0, "register slots come in pairs (blocks of two orders)");
  // order held by register slot s of this lane (canonical roles): k = 2 ((s / 2) G + lane) + s % 2
#define ORDER_OF(s) (((((s) >> 1) << lgG) + lane) * 2 + ((s) & 1))
// register slots (whole pairs) that hold orders 0..n
#define SLOTS_FOR(n) ((((n) >> (lgG + 1)) + 1) * 2)
  // bit of the canonical register of order C = max_order + 1 (the one a truncating shift must clear) if this lane holds it
  const unsigned mtop = (((p.C >> 1) & (G - 1)) == lane && ((p.C >> 1) >> lgG) * 2 + 1 < NS) ? 1u << (((p.C >> 1) >> lgG) * 2 + (p.C & 1)) : 0u;
  real P[NS], M[NS], Z[NS];
  real m0 = ldc(coef + p.m0_off + patoff[p.m0_pat]);
  {
    const real *ib = coef + p.init_off + patoff[p.init_pat];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const int k = ORDER_OF(s);
      const bool in = k <= p.init_n;
      P[s] = in ? ldc(ib + 6 * k) : real(0);
      M[s] = in ? ldc(ib + 6 * k + 2) : real(0);
      Z[s] = in ? ldc(ib + 6 * k + 4) : real(0);
    }
  }
  real2 *sig = (real2 *)p.signal;
  real *sigr = (real *)p.signal; // p.out_real: rows of REALS (the imaginary parts of this kernel's signal are exact zeros)

#define SLOT_CASE(K, ...)                      \
  case (K) + 1:                                \
    if (NS > (K)) {                            \
      constexpr int s = (K) < NS ? (K) : 0;    \
      __VA_ARGS__                              \
    }
#define DUFF(n, ...)                                                                                        \
  switch (n) {                                                                                              \
    SLOT_CASE(31, __VA_ARGS__) SLOT_CASE(30, __VA_ARGS__) SLOT_CASE(29, __VA_ARGS__) SLOT_CASE(28, __VA_ARGS__) \
    SLOT_CASE(27, __VA_ARGS__) SLOT_CASE(26, __VA_ARGS__) SLOT_CASE(25, __VA_ARGS__) SLOT_CASE(24, __VA_ARGS__) \
    SLOT_CASE(23, __VA_ARGS__) SLOT_CASE(22, __VA_ARGS__) SLOT_CASE(21, __VA_ARGS__) SLOT_CASE(20, __VA_ARGS__) \
    SLOT_CASE(19, __VA_ARGS__) SLOT_CASE(18, __VA_ARGS__) SLOT_CASE(17, __VA_ARGS__) SLOT_CASE(16, __VA_ARGS__) \
    SLOT_CASE(15, __VA_ARGS__) SLOT_CASE(14, __VA_ARGS__) SLOT_CASE(13, __VA_ARGS__) SLOT_CASE(12, __VA_ARGS__) \
    SLOT_CASE(11, __VA_ARGS__) SLOT_CASE(10, __VA_ARGS__) SLOT_CASE(9, __VA_ARGS__) SLOT_CASE(8, __VA_ARGS__)   \
    SLOT_CASE(7, __VA_ARGS__) SLOT_CASE(6, __VA_ARGS__) SLOT_CASE(5, __VA_ARGS__) SLOT_CASE(4, __VA_ARGS__)     \
    SLOT_CASE(3, __VA_ARGS__) SLOT_CASE(2, __VA_ARGS__) SLOT_CASE(1, __VA_ARGS__) SLOT_CASE(0, __VA_ARGS__)     \
  default:                                                                                                  \
    break;                                                                                                  \
  }
// F+' = a F+ + b F- + u Z ; F-' = b F+ + a F- + u Z ; Z' = w Z + h (F+ + F-)
#define APPLY5(a, w, b, u, h)                                              \
  DUFF(nslot, {                                                            \
    const real p_ = P[s], m_ = M[s], z_ = Z[s], s_ = p_ + m_, q_ = fma(b, s_, u * z_); \
    P[s] = fma(a - b, p_, q_);                                             \
    M[s] = fma(a - b, m_, q_);                                             \
    Z[s] = fma(w, z_, h * s_);                                             \
  })

// close a segment (reset / unit shift) and open the next one (order count of the next pass)
#define DO_SEG(SHIFT_, NOLD_, NNEW_, SFLAGS_, NEXT_)                                                       \
  {                                                                                                        \
    const int shift = (SHIFT_), n_old = (NOLD_), n_new = (NNEW_), sflags = (SFLAGS_);                      \
    const int n_move = max(min(n_new, nact + 1), 0); /* orders above nact + 1 are unobservable: they stay */ \
    nact = (NEXT_);                                                                                        \
    nslot = nact < 0 ? 0 : SLOTS_FOR(nact);                                                                \
    if (sflags & EPGX_SEG_RESET) {                                                                         \
      _Pragma("unroll") for (int s = 0; s < NS; ++s) P[s] = M[s] = Z[s] = real(0);                         \
      if (lane == 0) Z[0] = m0;                                                                            \
    } else if (shift != 0) {                                                                               \
      const int npair = (n_move >> (lgG + 1)) + 1;                                                         \
      if (shift > 0) SHIFT_REAL(P, M) else SHIFT_REAL(M, P)                                                \
      if (sflags & EPGX_SEG_MASK_TOP) {                                                                    \
        _Pragma("unroll") for (int s = 0; s < NS; ++s)                                                     \
          if (ORDER_OF(s) > n_new) {                                                                       \
            if (shift > 0) P[s] = real(0); else M[s] = real(0);                                            \
          }                                                                                                \
      }                                                                                                    \
    }                                                                                                      \
  }
// U: the component whose orders move up (F+ for shift > 0), D: the other one; the new order 0 of U is the old
// order 1 of D (real state: no conjugation), which lane 0 holds itself.  Canonical roles before and after:
// up: the odd register of every pair rotates one lane up (the LAST lane sending the value of its previous pair)
// and becomes the even one, the old even one becomes the odd one; dn: mirrored.
#define PAIR_CASE(K, ...)                           \
  case (K) + 1:                                     \
    if (NS > 2 * (K)) {                             \
      constexpr int sp = 2 * (K) < NS ? (K) : 0;    \
      __VA_ARGS__                                   \
    }
#define DUFFP(n, ...)                                                                                       \
  switch (n) {                                                                                              \
    PAIR_CASE(15, __VA_ARGS__) PAIR_CASE(14, __VA_ARGS__) PAIR_CASE(13, __VA_ARGS__) PAIR_CASE(12, __VA_ARGS__) \
    PAIR_CASE(11, __VA_ARGS__) PAIR_CASE(10, __VA_ARGS__) PAIR_CASE(9, __VA_ARGS__) PAIR_CASE(8, __VA_ARGS__)   \
    PAIR_CASE(7, __VA_ARGS__) PAIR_CASE(6, __VA_ARGS__) PAIR_CASE(5, __VA_ARGS__) PAIR_CASE(4, __VA_ARGS__)     \
    PAIR_CASE(3, __VA_ARGS__) PAIR_CASE(2, __VA_ARGS__) PAIR_CASE(1, __VA_ARGS__) PAIR_CASE(0, __VA_ARGS__)     \
  default:                                                                                                  \
    break;                                                                                                  \
  }
#define SHIFT_REAL(U, D)                                                                                   \
  {                                                                                                        \
    const real f1 = n_old < 1 ? real(0) : D[1];                                                            \
    DUFFP(npair, {                                                                                         \
      const real v = (is_last && sp > 0) ? U[sp > 0 ? 2 * sp - 1 : 1] : U[2 * sp + 1];                     \
      U[2 * sp + 1] = U[2 * sp];                                                                           \
      U[2 * sp] = __shfl_sync(FULL, v, srcUp);                                                             \
    })                                                                                                     \
    if (is_first) U[0] = f1;                                                                               \
    real keep = real(0); /* old even value of the pair above (zero above the populated orders) */          \
    DUFFP(npair, {                                                                                         \
      const real cur = D[2 * sp];                                                                          \
      D[2 * sp] = D[2 * sp + 1];                                                                           \
      D[2 * sp + 1] = __shfl_sync(FULL, is_first ? keep : cur, srcDn);                                     \
      keep = cur;                                                                                          \
    })                                                                                                     \
  }

  const int4 *stream = (const int4 *)p.stream;
  const int nthreads = blockDim.x;
  for (int i = tid; i < 2 * RCH && i < 2 * p.nstream; i += nthreads) __pipeline_memcpy_async(tbuf + i, stream + i, 16);
  __pipeline_commit();
  int nact = -1, nslot = 0;
  for (int base0 = 0, chunk = 0; base0 < p.nstream; base0 += RCH, ++chunk) {
    __pipeline_wait_prior(0);
    __syncthreads();
    {
      const int nb = base0 + RCH;
      int4 *dst = tbuf + ((chunk + 1) & 1) * 2 * RCH;
      for (int i = tid; i < 2 * RCH && nb * 2 + i < 2 * p.nstream; i += nthreads)
        __pipeline_memcpy_async(dst + i, stream + (size_t)nb * 2 + i, 16);
      __pipeline_commit();
    }
    const int4 *tb0 = tbuf + (chunk & 1) * 2 * RCH;
    // host windows of this kernel window that are whole-TR windows, joined while the previous one is full
    int joined = 0; // host windows consumed by the joined fast path that starts at sub-window `h`
    for (int h = 0; h < KW; h += joined > 0 ? joined : 1) {
    const int base = base0 + h * TAPE_CHUNK;
    if (base >= p.nstream) break;
    const int4 *tb = tb0 + h * 2 * TAPE_CHUNK;
    const int cnt = min(TAPE_CHUNK, p.nstream - base);
    joined = 0;
    if (G >= 8 && (tb[0].x & EPGX_CHUNK_PURE_TR)) {
      // ---- fast path: the window holds whole-TR records (shift +1, no flags).  Phase 1: the G lanes
      // of an atom decode the TRs (lane, lane + G, ...), gather their coefficients, fuse them and stage them in the
      // atom's shared-memory rows -- one vectorised pass for the window.  Phase 2 (tr_window) runs the TRs in order: no
      // global load and no decode on the per-TR path.  need: register pairs (2 G orders each) the window needs -- a TR
      // applies to orders 0..nact and shifts orders 0..min(n_new, nact + 1); what lies above nact + 1 is unobservable
      // (lowering.py) and need not move; nact of TR j is the "next nact" of TR j - 1
      // (the window runs in parts of 16 TRs, each with its own pair count: 0.13 pair less per TR on average)
      constexpr int HALF = TAPE_CHUNK / 4;
      static_assert(kRealWindows == 1 || kRealWindows == 2, "one or two host windows per kernel window");
      // whole-TR records of this window: TAPE_CHUNK / 2, or an even number below it (rest: NOP padding); the next host
      // window joins when this one is full and the run goes on (the TR records are contiguous then)
      const int ntr0 = tb[1].w;
      const bool join = KW == 2 && G == 32 && h == 0 && ntr0 == TAPE_CHUNK / 2 && base + TAPE_CHUNK < p.nstream &&
                        (tb[2 * TAPE_CHUNK].x & EPGX_CHUNK_PURE_TR);
      const int ntr = join ? ntr0 + tb[2 * TAPE_CHUNK + 1].w : ntr0;
      joined = join ? 2 : 1;
      int need0 = 0, need1 = 0, need2 = 0, need3 = 0;
      bool bad = false;
      for (int j = lane; j < ntr; j += G) {
        const int4 a0 = tb[4 * j], a1 = tb[4 * j + 1], b0 = tb[4 * j + 2], b1 = tb[4 * j + 3];
        const int fl = (a0.x >> 16) & 0xffff;
        real g[10]; // T (a, w, b, u), E_pre (e1, r0, e2), E_post (e1, r0, e2)
        {
          const real *ct = coef + (unsigned)a0.z + patoff[a1.y & 0xff];
          const real *ca = coef + (unsigned)a0.w + patoff[(a1.y >> 8) & 0xff];
          const real *cb = coef + (unsigned)b0.z + patoff[b1.y & 0xff];
          g[0] = ldc(ct); g[1] = ldc(ct + 1); g[2] = ldc(ct + 2); g[3] = ldc(ct + 3);
          g[4] = ldc(ca); g[5] = ldc(ca + 1); g[6] = ldc(coef + (unsigned)a1.x + patoff[(a1.y >> 16) & 0xff]);
          g[7] = ldc(cb); g[8] = ldc(cb + 1); g[9] = ldc(coef + (unsigned)b0.w + patoff[(b1.y >> 8) & 0xff]);
        }
        const Fused5<real> fv = fuse5<real>(g[0], g[1], g[2], g[3], fl & EPGX_FLAG_PRE, g[4], g[5], g[6], fl & EPGX_FLAG_POST, g[7],
                                            g[8], g[9], false, m0);
        real2 *c = (real2 *)(cw + 8 * j);
        c[0] = real2{fv.a - fv.b, fv.w}; c[1] = real2{fv.b, fv.u};
        c[2] = real2{fv.h, ((b0.x >> 18) & EPGX_SEG_MASK_TOP) ? real(1) : real(0)}; c[3] = real2{fv.fz, fv.zz};
        const real au = fabs(fv.u);
        bad = bad || !(au >= (sizeof(real) == 4 ? real(1e-15) : real(1e-100)) && au <= real(4));
        const int cur = j == 0 ? nact : tb[4 * j - 1].z;
        const int nd = (max(max(min((int)((unsigned)b1.x & 0xffff), cur + 1), cur), 0) >> (lgG + 1)) + 1;
        if (j < HALF) need0 = max(need0, nd); else if (j < 2 * HALF) need1 = max(need1, nd);
        else if (j < 3 * HALF) need2 = max(need2, nd); else need3 = max(need3, nd);
      }
      need0 = __reduce_max_sync(FULL, need0);
      need1 = __reduce_max_sync(FULL, need1);
      need2 = __reduce_max_sync(FULL, need2);
      need3 = __reduce_max_sync(FULL, need3);
      const int needsum = need0 + need1 + need2 + need3;
      // scaled window (see tr_window): every u of the window usable as a scale for every atom of the warp
      // (worth it when the orders in flight outweigh the reciprocal per TR of the second staging pass)
      const bool sc = EPGX_REAL_SCALED && needsum * G * 2 >= 96 * ((ntr + HALF - 1) / HALF) && !__any_sync(FULL, bad);
      __syncwarp();
      if (sc) {
        for (int j = lane; j < ntr; j += G) {
          real *c = cw + 8 * j;
          const real un = j + 1 < ntr ? c[8 + 3] : real(1);
          c[1] = c[1] * (un * rcp_newton(c[3]));
          c[4] *= un;
          c[7] *= un;
        }
        __syncwarp();
      }
#define TRW(K_)                                                                                                              \
  case K_:                                                                                                                   \
    if (sc) tr_window<real, NS, K_, BOUNDED, true>(P, M, Z, cw, sb, lane == 0, is_first, is_last, srcUp, srcDn, mtop, jb, min(jb + HALF, ntr), af, afs); \
    else tr_window<real, NS, K_, BOUNDED, false>(P, M, Z, cw, sb, lane == 0, is_first, is_last, srcUp, srcDn, mtop, jb, min(jb + HALF, ntr), af, afs); \
    break;
#pragma unroll 1
      for (int jb = 0; jb < ntr; jb += HALF) {
        switch (jb < HALF ? need0 : jb < 2 * HALF ? need1 : jb < 3 * HALF ? need2 : need3) {
          TRW(1) TRW(2) TRW(3) TRW(4) TRW(5) TRW(6) TRW(7) TRW(8) TRW(9) TRW(10) TRW(11) TRW(12) TRW(13) TRW(14) TRW(15) TRW(16)
        default: break;
        }
      }
#undef TRW
      __syncwarp();
      if (valid) // lane j (+ G, ...) stores the echoes of its TRs
        for (int j = lane; j < ntr; j += G) {
          const long long o = (long long)tb[4 * j + 2].y * p.sig_stride + a_rel;
          if (p.out_real) sigr[o] = sb[j]; else sig[o] = real2{sb[j], real(0)};
        }
      nact = tb[4 * (ntr - 1) + 3].z;
      nslot = nact < 0 ? 0 : SLOTS_FOR(nact);
      continue;
    }
    for (int r = 0; r < cnt; ++r) {
      const int4 r0 = tb[2 * r], r1 = tb[2 * r + 1];
      const int code = r0.x & 0xffff, flags = (r0.x >> 16) & 0xffff, aux = r0.y;
      const unsigned off0 = (unsigned)r0.z, off1 = (unsigned)r0.w, off2 = (unsigned)r1.x;
      const int pat0 = r1.y & 0xff, pat1 = (r1.y >> 8) & 0xff, pat2 = (r1.y >> 16) & 0xff;

      switch (code) {
      case EPGX_OP_NOP:
        if (aux) r = cnt; // padding up to the next window (a run of whole-TR records starts there)
        break;
      case EPGX_OP_FUSED: {
        const int4 q0 = tb[2 * r + 2], q1 = tb[2 * r + 3]; // the CONT record (never split from FUSED)
        const real *ct = coef + off0 + patoff[pat0];
        const real *ca = coef + off1 + patoff[pat1];
        const real *cb = coef + (unsigned)q0.z + patoff[q1.y & 0xff];
        const Fused5<real> f = fuse5<real>(ldc(ct), ldc(ct + 1), ldc(ct + 2), ldc(ct + 3), flags & EPGX_FLAG_PRE, ldc(ca),
                                           ldc(ca + 1), ldc(coef + off2 + patoff[pat2]), flags & EPGX_FLAG_POST, ldc(cb),
                                           ldc(cb + 1), ldc(coef + (unsigned)q0.w + patoff[(q1.y >> 8) & 0xff]), false, m0);
        APPLY5(f.a, f.w, f.b, f.u, f.h)
        if (lane == 0 && nslot > 0) { P[0] += f.fz; M[0] += f.fz; Z[0] += f.zz; }
        ++r;
      } break;
      case EPGX_OP_T_RE: {
        const real *c = coef + off0 + patoff[pat0];
        const real a = ldc(c), w = ldc(c + 1), b = ldc(c + 2), u = ldc(c + 3), h = real(-0.5) * u;
        APPLY5(a, w, b, u, h)
      } break;
      case EPGX_OP_E: {
        const real *c0 = coef + off0 + patoff[pat0];
        const real e1 = ldc(c0), r0v = ldc(c0 + 1);
        const real e2 = ldc(coef + off1 + patoff[pat1]);
        DUFF(nslot, { P[s] *= e2; M[s] *= e2; Z[s] *= e1; })
        if ((flags & EPGX_FLAG_AFFINE) && lane == 0 && nslot > 0) Z[0] += r0v * m0;
      } break;
      case EPGX_OP_D: {
        const real *c = coef + off0 + patoff[pat0];
        DUFF(nslot, {
          const int k = min(ORDER_OF(s), p.C - 1);
          P[s] *= ldc(c + 3 * k); M[s] *= ldc(c + 3 * k + 1); Z[s] *= ldc(c + 3 * k + 2);
        })
      } break;
      case EPGX_OP_SPOIL:
        DUFF(nslot, { P[s] = real(0); M[s] = real(0); })
        break;
      case EPGX_OP_PD:
        m0 = ldc(coef + off0 + patoff[pat0]);
        break;
      case EPGX_OP_ADC:
        if (lane == 0 && valid && (flags & EPGX_FLAG_BASE)) {
          real fr = real(1), fi = real(0);
          if (flags & EPGX_FLAG_SCALE) {
            const real *c = coef + off0 + patoff[pat0];
            fr = ldc(c); fi = ldc(c + 1);
          }
          const real x = (flags & EPGX_FLAG_Z0) ? Z[0] : P[0];
          const long long o = (long long)aux * p.sig_stride + a_rel;
          if (p.out_real) sigr[o] = x * fr; else sig[o] = real2{x * fr, x * fi};
        }
        break;
      case EPGX_OP_SEG: {
        DO_SEG((int)off0, (int)off1, (int)off2, r1.z, aux)
      } break;
      case EPGX_OP_TR: {
        // one whole TR: FUSED (E.T.E) + plain ADC + the segment's unit shift, one decode (see epgx.cu)
        const int4 q0 = tb[2 * r + 2], q1 = tb[2 * r + 3];
        const real *ct = coef + off0 + patoff[pat0];
        const real *ca = coef + off1 + patoff[pat1];
        const real *cb = coef + (unsigned)q0.z + patoff[q1.y & 0xff];
        const Fused5<real> f = fuse5<real>(ldc(ct), ldc(ct + 1), ldc(ct + 2), ldc(ct + 3), flags & EPGX_FLAG_PRE, ldc(ca),
                                           ldc(ca + 1), ldc(coef + off2 + patoff[pat2]), flags & EPGX_FLAG_POST, ldc(cb),
                                           ldc(cb + 1), ldc(coef + (unsigned)q0.w + patoff[(q1.y >> 8) & 0xff]), false, m0);
        APPLY5(f.a, f.w, f.b, f.u, f.h)
        if (lane == 0 && nslot > 0) { P[0] += f.fz; M[0] += f.fz; Z[0] += f.zz; }
        if (lane == 0 && valid) {
          const long long o = (long long)q0.y * p.sig_stride + a_rel;
          if (p.out_real) sigr[o] = P[0]; else sig[o] = real2{P[0], real(0)};
        }
        const int segw = (q0.x >> 16) & 0xffff; // (shift + 1) | segment flags << 2
        DO_SEG((segw & 3) - 1, (int)((unsigned)q1.x >> 16), (int)((unsigned)q1.x & 0xffff), segw >> 2, q1.z)
        ++r;
      } break;
      default:
        break;
      }
    }
    } // host windows of this kernel window
  }
#undef APPLY5
#undef DO_SEG
#undef SHIFT_REAL
#undef DUFF
#undef SLOT_CASE
#undef DUFFP
#undef PAIR_CASE
#undef ORDER_OF
#undef SLOTS_FOR
}

} // namespace epgx
