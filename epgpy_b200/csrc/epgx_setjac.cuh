// epgx_setjac.cuh -- derivative kernel for real-valued phase graphs with ONE WARP PER STATE SET.
//
// Same eligibility as epgx_realjac.cuh (real-valued tape, real-valued order-1 injections: epgpy/diff.py:264-288,
// transition.py:172-186, evolution.py:360-389) for at most three variables and at most 32 NS orders.  Where
// realjac_kernel spreads the ORDERS of an atom over several warps (every warp holds base + partial states of
// its orders, and every unit shift exchanges boundary values between the warps), this kernel gives every STATE
// SET its own warp: warp 0 owns the base state x, warp v + 1 the partial state y_v, each with all the orders in
// the blocked layout of epgx_real.cuh (order k = 2 b + i in lane b % 32, register 2 (b / 32) + i), so a shift
// is the cheap single-warp rotation.  What crosses warps is the base state itself: before a whole-TR group
// (EPGX_OP_TRJ: x' = F x + c_0, y_v' = F y_v + J_v x + c_v, see epgx_realjac.cuh) warp 0 publishes x in shared
// memory (three conflict-free STS.128 per register pair, double-buffered), one named barrier, and the partial
// warps read it back (three LDS.128 per pair) for their J_v x term.  Per TR and order that is 9 (base) or 18
// (partial) FP64 instructions against ~1.5 memory instructions and one shuffled register per block and
// component -- 76 % arithmetic in the partial warps, against 42 % for the orders-over-warps kernel.
//
// Records that are not whole-TR groups run through a generic path with the same roles (an injection record
// publishes the base state first); sequences made mostly of such records are better served by realjac_kernel
// (the host chooses: epgx.cu, choose_variant).
#pragma once
#include <cuda_pipeline.h>

#include "epgx_common.cuh"
#include "epgx_realjac.cuh"
#include "epgx_reg.cuh"

namespace epgx {

constexpr int SJ_ROW = 16; // staged coefficients of one TR: F (a, w, b, u, h, f, z, -) then J_q (same order)

template <typename real> struct SjCtx {
  int q, lane, srcUp, srcDn, nthreads;
  bool is_first, is_last;
  real *xpub;   // published base state: [2 buffers][F+ | F- | Z][32 NS orders]
  const real *cw; // this warp's coefficient rows [TRJ_PER_WINDOW][SJ_ROW]
  real *sb;     // this warp's echoes of the window
};

// publish the (pre-operator) base state, KP register pairs, in the roles of phase ph; order k = 64 sp + 2 lane + i
template <typename real, int NS, int KP>
__device__ __forceinline__ void sj_publish(const real (&P)[NS], const real (&M)[NS], const real (&Z)[NS], real *xb, int lane, int ph) {
  typedef typename vec2<real>::type real2;
  real2 *xp = (real2 *)xb, *xm = xp + NS * 16, *xz = xm + NS * 16;
#pragma unroll
  for (int sp = 0; sp < KP; ++sp) {
    xp[32 * sp + lane] = ph ? real2{P[2 * sp + 1], P[2 * sp]} : real2{P[2 * sp], P[2 * sp + 1]};
    xm[32 * sp + lane] = ph ? real2{M[2 * sp + 1], M[2 * sp]} : real2{M[2 * sp], M[2 * sp + 1]};
    xz[32 * sp + lane] = real2{Z[2 * sp], Z[2 * sp + 1]};
  }
}

// one whole TR (without its shift) on KP register pairs in the roles of phase ph: base warp x' = F x, partial
// warp y' = F y + J x with x read from the published buffer xb.  F = [[a b u] [b a u] [h h w]] and J of the same shape:
// with s = F+ + F- and c = a - b the rows of F+ and F- share q = b s + u Z (epgx_real.cuh) -- 7 instead of 8
// floating-point instructions per order for the base warp, 14 instead of 18 for a partial warp
template <typename real, int NS, int KP>
__device__ __forceinline__ void sj_apply(real (&P)[NS], real (&M)[NS], real (&Z)[NS], const real *cf, const real *xb, int q, int lane,
                                         int ph) {
  typedef typename vec2<real>::type real2;
  const real2 c0 = ((const real2 *)cf)[0], c1v = ((const real2 *)cf)[1], c2 = ((const real2 *)cf)[2];
  const real a = c0.x, w = c0.y, b = c1v.x, u = c1v.y, h = c2.x;
  if constexpr (sizeof(real) == 4) {
    // FP32, canonical roles (ph == 0 for every caller of the float instances): packed f32x2 arithmetic on the two
    // orders of a block (epgx_real.cuh), 8 (base) / 18 (partial) instructions per register pair instead of 18 / 36
    const float2 c2v = make_float2(a - b, a - b), b2 = make_float2(b, b), u2 = make_float2(u, u), w2 = make_float2(w, w), h2 = make_float2(h, h);
    if (q == 0) {
#pragma unroll
      for (int sp = 0; sp < KP; ++sp) {
        const float2 p2 = make_float2(P[2 * sp], P[2 * sp + 1]), m2 = make_float2(M[2 * sp], M[2 * sp + 1]);
        const float2 z2 = make_float2(Z[2 * sp], Z[2 * sp + 1]);
        const float2 s2 = __fadd2_rn(p2, m2);
        const float2 q2 = __ffma2_rn(b2, s2, __fmul2_rn(u2, z2));
        const float2 np = __ffma2_rn(c2v, p2, q2), nm = __ffma2_rn(c2v, m2, q2);
        const float2 nz = __ffma2_rn(w2, z2, __fmul2_rn(h2, s2));
        P[2 * sp] = np.x; P[2 * sp + 1] = np.y; M[2 * sp] = nm.x; M[2 * sp + 1] = nm.y; Z[2 * sp] = nz.x; Z[2 * sp + 1] = nz.y;
      }
    } else {
      const float2 j0 = ((const float2 *)cf)[4], j1 = ((const float2 *)cf)[5], j2 = ((const float2 *)cf)[6];
      const float2 jc2 = make_float2(j0.x - j1.x, j0.x - j1.x), jw2 = make_float2(j0.y, j0.y), jb2 = make_float2(j1.x, j1.x),
                   ju2 = make_float2(j1.y, j1.y), jh2 = make_float2(j2.x, j2.x);
      const float2 *xp = (const float2 *)xb, *xm = xp + NS * 16, *xz = xm + NS * 16;
#pragma unroll
      for (int sp = 0; sp < KP; ++sp) {
        const float2 vp = xp[32 * sp + lane], vm = xm[32 * sp + lane], vz = xz[32 * sp + lane];
        const float2 p2 = make_float2(P[2 * sp], P[2 * sp + 1]), m2 = make_float2(M[2 * sp], M[2 * sp + 1]);
        const float2 z2 = make_float2(Z[2 * sp], Z[2 * sp + 1]);
        const float2 s2 = __fadd2_rn(p2, m2), xs2 = __fadd2_rn(vp, vm);
        // q = b s + u Z + jb xs + ju x_Z, shared by F+ and F-
        const float2 q2 = __ffma2_rn(b2, s2, __ffma2_rn(u2, z2, __ffma2_rn(jb2, xs2, __fmul2_rn(ju2, vz))));
        const float2 np = __ffma2_rn(c2v, p2, __ffma2_rn(jc2, vp, q2));
        const float2 nm = __ffma2_rn(c2v, m2, __ffma2_rn(jc2, vm, q2));
        const float2 nz = __ffma2_rn(w2, z2, __ffma2_rn(h2, s2, __ffma2_rn(jw2, vz, __fmul2_rn(jh2, xs2))));
        P[2 * sp] = np.x; P[2 * sp + 1] = np.y; M[2 * sp] = nm.x; M[2 * sp + 1] = nm.y; Z[2 * sp] = nz.x; Z[2 * sp + 1] = nz.y;
      }
    }
  } else if (q == 0) {
    const real c = a - b;
#pragma unroll
    for (int sp = 0; sp < KP; ++sp)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = ph ? 2 * sp + 1 - i : 2 * sp + i;
        const real p_ = P[r], m_ = M[r], z_ = Z[2 * sp + i];
        const real s_ = p_ + m_, q_ = fma(b, s_, u * z_);
        P[r] = fma(c, p_, q_);
        M[r] = fma(c, m_, q_);
        Z[2 * sp + i] = fma(w, z_, h * s_);
      }
  } else {
    const real2 j0 = ((const real2 *)cf)[4], j1 = ((const real2 *)cf)[5], j2 = ((const real2 *)cf)[6];
    const real jc = j0.x - j1.x, jw = j0.y, jb = j1.x, ju = j1.y, jh = j2.x, c = a - b;
    const real2 *xp = (const real2 *)xb, *xm = xp + NS * 16, *xz = xm + NS * 16;
#pragma unroll
    for (int sp = 0; sp < KP; ++sp) {
      const real2 vp = xp[32 * sp + lane], vm = xm[32 * sp + lane], vz = xz[32 * sp + lane];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = ph ? 2 * sp + 1 - i : 2 * sp + i;
        const real x_p = i ? vp.y : vp.x, x_m = i ? vm.y : vm.x, x_z = i ? vz.y : vz.x;
        const real p_ = P[r], m_ = M[r], z_ = Z[2 * sp + i];
        const real s_ = p_ + m_, xs = x_p + x_m;
        const real q_ = fma(b, s_, fma(u, z_, fma(jb, xs, ju * x_z)));
        P[r] = fma(c, p_, fma(jc, x_p, q_));
        M[r] = fma(c, m_, fma(jc, x_m, q_));
        Z[2 * sp + i] = fma(w, z_, fma(h, s_, fma(jw, x_z, jh * xs)));
      }
    }
  }
}

// One tape window of TRJ_PER_WINDOW plain whole-TR groups (unit shift +1, no segment flags), KP register pairs.
// One TR per iteration with the canonical register roles: unlike tr_window of epgx_real.cuh the body is NOT unrolled
// over the two role phases -- a partial warp runs 36 FP64 instructions per register pair, and a two-phase body for
// eight pairs (13 KB of code next to the base warp's 6 KB) overflowed the instruction caches (ncu: 3.4 cycles of
// "no instruction" per issue); the two register copies per pair and component of the canonical shift are cheaper.
template <typename real, int NS, int KP>
__device__ __forceinline__ void sj_window(real (&P)[NS], real (&M)[NS], real (&Z)[NS], const SjCtx<real> &c, real m0, int &par) {
  const unsigned FULL = 0xffffffffu;
  if constexpr (2 * KP <= NS) {
#pragma unroll 1
    for (int j = 0; j < TRJ_PER_WINDOW; ++j) {
      const real *cf = c.cw + j * SJ_ROW;
      real *xb = c.xpub + par * (3 * NS * 32);
      if (c.q == 0) sj_publish<real, NS, KP>(P, M, Z, xb, c.lane, 0);
      asm volatile("bar.sync 1, %0;" ::"r"(c.nthreads) : "memory");
      par ^= 1;
      sj_apply<real, NS, KP>(P, M, Z, cf, xb, c.q, c.lane, 0);
      if (c.lane == 0) {
        const real *ca = cf + (c.q == 0 ? 0 : 8);
        const real f = ca[5] * m0;
        P[0] += f; M[0] += f; Z[0] += ca[6] * m0;
        c.sb[j] = P[0]; // the echo (or its derivative) of the TR; stored by lane j after the window
      }
      // unit shift +1, canonical roles (epgx_real.cuh: SHIFT_REAL)
      const real f1 = M[1];
#pragma unroll
      for (int sp = KP - 1; sp >= 0; --sp) {
        const real v = (c.is_last && sp > 0) ? P[sp > 0 ? 2 * sp - 1 : 1] : P[2 * sp + 1];
        P[2 * sp + 1] = P[2 * sp];
        P[2 * sp] = __shfl_sync(FULL, v, c.srcUp);
      }
      if (c.is_first) P[0] = f1;
      real keep = real(0);
#pragma unroll
      for (int sp = KP - 1; sp >= 0; --sp) {
        const real cur = M[2 * sp];
        M[2 * sp] = M[2 * sp + 1];
        M[2 * sp + 1] = __shfl_sync(FULL, c.is_first ? keep : cur, c.srcDn);
        keep = cur;
      }
    }
  }
}

template <typename real, int NS>
__global__ void __launch_bounds__(128, (sizeof(real) * NS <= 64 ? 4 : 3)) setjac_kernel(const KParams p) {
  typedef typename vec2<real>::type real2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  static_assert(NS % 2 == 0, "register slots come in pairs (blocks of two orders)");
  constexpr int lgG = 5;

  const int tid = threadIdx.x;
  const int q = tid >> 5; // state set of this warp: 0 = base, v + 1 = partial state of variable v
  const int lane = tid & 31;
  const int nthreads = blockDim.x;
  const int srcUp = (lane - 1) & 31, srcDn = (lane + 1) & 31;
  const unsigned FULL = 0xffffffffu;
  const bool is_last = lane == 31, is_first = lane == 0;

  const long long a_rel = blockIdx.x;
  const long long atom = p.atom_begin + a_rel;
  const real *__restrict__ coef = (const real *)p.coef;

  // shared memory: tape window [2][TAPE_CHUNK] records, pattern offsets [npattern], published base state
  // [2][3][32 NS], then per warp: coefficient rows [TRJ_PER_WINDOW][SJ_ROW] and echoes [16]
  int4 *tbuf = (int4 *)smem_raw;
  int *patoff = (int *)(tbuf + 2 * TAPE_CHUNK * 2);
  real *xpub = (real *)(patoff + ((p.npattern + 3) & ~3));
  real *cw = xpub + 2 * 3 * NS * 32 + (size_t)q * (TRJ_PER_WINDOW * SJ_ROW + 16);
  real *sb = cw + TRJ_PER_WINDOW * SJ_ROW;
  {
    int idx[EPGX_MAX_DIMS];
    long long r = atom;
    for (int d = p.ndim - 1; d >= 0; --d) {
      idx[d] = (int)(r % p.shape[d]);
      r /= p.shape[d];
    }
    for (int g = tid; g < p.npattern; g += nthreads) {
      const int *st = p.pats + g * (EPGX_MAX_DIMS + 1);
      int o = 0;
      for (int d = 0; d < p.ndim; ++d) o += idx[d] * st[d];
      patoff[g] = o;
    }
  }
  __syncthreads();

#define ORDER_OF(s) (((((s) >> 1) << lgG) + lane) * 2 + ((s) & 1))
#define SLOTS_FOR(n) ((((n) >> (lgG + 1)) + 1) * 2)
  real P[NS], M[NS], Z[NS];
  real m0 = ldc(coef + p.m0_off + patoff[p.m0_pat]);
  {
    const real *ib = coef + p.init_off + patoff[p.init_pat];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const int k = ORDER_OF(s);
      const bool in = q == 0 && k <= p.init_n;
      P[s] = in ? ldc(ib + 6 * k) : real(0);
      M[s] = in ? ldc(ib + 6 * k + 2) : real(0);
      Z[s] = in ? ldc(ib + 6 * k + 4) : real(0);
    }
  }
  real2 *sig = (real2 *)p.signal + a_rel;
  real2 *jac = (real2 *)p.jac + a_rel;
  int par = 0; // publication buffer

  SjCtx<real> ctx;
  ctx.q = q; ctx.lane = lane; ctx.srcUp = srcUp; ctx.srcDn = srcDn; ctx.nthreads = nthreads;
  ctx.is_first = is_first; ctx.is_last = is_last; ctx.xpub = xpub; ctx.cw = cw; ctx.sb = sb;

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
// canonical unit shift of the blocked layout (epgx_real.cuh: SHIFT_REAL)
#define SHIFT_REAL(U, D)                                                                                   \
  {                                                                                                        \
    const real f1 = n_old < 1 ? real(0) : D[1];                                                            \
    DUFFP(npair, {                                                                                         \
      const real v = (is_last && sp > 0) ? U[sp > 0 ? 2 * sp - 1 : 1] : U[2 * sp + 1];                     \
      U[2 * sp + 1] = U[2 * sp];                                                                           \
      U[2 * sp] = __shfl_sync(FULL, v, srcUp);                                                             \
    })                                                                                                     \
    if (is_first) U[0] = f1;                                                                               \
    real keep = real(0);                                                                                   \
    DUFFP(npair, {                                                                                         \
      const real cur = D[2 * sp];                                                                          \
      D[2 * sp] = D[2 * sp + 1];                                                                           \
      D[2 * sp + 1] = __shfl_sync(FULL, is_first ? keep : cur, srcDn);                                     \
      keep = cur;                                                                                          \
    })                                                                                                     \
  }
#define DO_SEG(SHIFT_, NOLD_, NNEW_, SFLAGS_, NEXT_)                                                       \
  {                                                                                                        \
    const int shift = (SHIFT_), n_old = (NOLD_), n_new = (NNEW_), sflags = (SFLAGS_);                      \
    const int n_move = max(min(n_new, nact + 1), 0); /* orders above nact + 1 are unobservable: they stay */ \
    nact = (NEXT_);                                                                                        \
    nslot = nact < 0 ? 0 : SLOTS_FOR(nact);                                                                \
    if (sflags & EPGX_SEG_RESET) {                                                                         \
      _Pragma("unroll") for (int s = 0; s < NS; ++s) P[s] = M[s] = Z[s] = real(0);                         \
      if (lane == 0 && q == 0) Z[0] = m0;                                                                  \
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
// make the whole base state visible to the partial warps (canonical roles); every warp takes part
#define PUBLISH_ALL()                                                                                      \
  real *xb = xpub + par * (3 * NS * 32);                                                                   \
  if (q == 0) sj_publish<real, NS, NS / 2>(P, M, Z, xb, lane, 0);                                          \
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");                                              \
  par ^= 1;                                                                                                \
  const real *xP = xb, *xM = xb + NS * 32, *xZ = xb + 2 * NS * 32;

  const int4 *stream = (const int4 *)p.stream;
  for (int i = tid; i < 2 * TAPE_CHUNK && i < 2 * p.nstream; i += nthreads) __pipeline_memcpy_async(tbuf + i, stream + i, 16);
  __pipeline_commit();
  int nact = -1, nslot = 0;
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
    const int wflags = tb[0].x;
    int rowv = 0, jrowv = 0;
    bool jpart = false;
    if (wflags & EPGX_CHUNK_ANY_TRJ) {
      // coefficient assembly: lane j of every warp fuses the coefficients of group j for its own state set
      int nnewv = 0, nextv = -1;
      if (lane < TRJ_PER_WINDOW && 5 * lane + 5 <= cnt && (tb[10 * lane].x & 0xffff) == EPGX_OP_TRJ) {
        const int4 *g = tb + 10 * lane;
        real *row = cw + lane * SJ_ROW;
        trj_assemble<real>(g, 0, coef, patoff, row);
        if (q > 0) trj_assemble<real>(g, q, coef, patoff, row + 8);
        rowv = g[2].y; jrowv = g[3].w; jpart = (g[0].x >> 16) & EPGX_FLAG_PARTIALS;
        nnewv = (int)((unsigned)g[3].x & 0xffff); nextv = g[3].z;
      }
      __syncwarp();
      if (wflags & EPGX_CHUNK_PURE_TRJ) {
        // ---- fast path: TRJ_PER_WINDOW plain groups.  Register pairs the window needs: a TR applies to orders
        // 0..nact and shifts orders 0..min(n_new, nact + 1); nact of TR j is the "next nact" of TR j - 1
        int curv = __shfl_up_sync(FULL, nextv, 1);
        if (lane == 0) curv = nact;
        int need = lane < TRJ_PER_WINDOW ? (max(min(nnewv, curv + 1), 0) >> 6) + 1 : 0;
        need = max(__reduce_max_sync(FULL, need), nslot >> 1);
#define SJW(K_) case K_: sj_window<real, NS, K_>(P, M, Z, ctx, m0, par); break;
        switch (need) {
          SJW(1) SJW(2) SJW(3) SJW(4) SJW(5) SJW(6) SJW(7) SJW(8) SJW(9) SJW(10) SJW(11) SJW(12) SJW(13) SJW(14) SJW(15) SJW(16)
        default: break;
        }
#undef SJW
        __syncwarp();
        if (lane < TRJ_PER_WINDOW) {
          if (q == 0) sig[(long long)rowv * p.sig_stride] = real2{sb[lane], real(0)};
          else if (jpart) jac[((long long)jrowv * p.nvar + q - 1) * p.jac_stride] = real2{sb[lane], real(0)};
        }
        nact = __shfl_sync(FULL, nextv, TRJ_PER_WINDOW - 1);
        nslot = nact < 0 ? 0 : SLOTS_FOR(nact);
        continue;
      }
    }
    for (int r = 0; r < cnt; ++r) {
      const int4 r0 = tb[2 * r], r1 = tb[2 * r + 1];
      const int code = r0.x & 0xffff, flags = (r0.x >> 16) & 0xffff, aux = r0.y;
      const unsigned off0 = (unsigned)r0.z, off1 = (unsigned)r0.w, off2 = (unsigned)r1.x;
      const int pat0 = r1.y & 0xff, pat1 = (r1.y >> 8) & 0xff;
      const bool inject = flags & EPGX_FLAG_INJECT;
      // does this warp's own state take the (non-injection) record?
      const bool mine = q == 0 ? (flags & EPGX_FLAG_BASE) : (flags & EPGX_FLAG_PARTIALS);
      const bool target = inject && q == aux + 1;
      const bool aff0 = (flags & EPGX_FLAG_AFFINE) && lane == 0 && nslot > 0;

      switch (code) {
      case EPGX_OP_T_RE: {
        const real *c = coef + off0 + patoff[pat0];
        const real a = ldc(c), w = ldc(c + 1), b = ldc(c + 2), u = ldc(c + 3), h = real(-0.5) * u;
        if (inject) {
          PUBLISH_ALL()
          if (target) {
            DUFF(nslot, {
              const int k = ORDER_OF(s);
              const real p_ = xP[k], m_ = xM[k], z_ = xZ[k];
              P[s] += a * p_ + b * m_ + u * z_;
              M[s] += a * m_ + b * p_ + u * z_;
              Z[s] += w * z_ + h * (p_ + m_);
            })
          }
        } else if (mine) {
          DUFF(nslot, {
            const real p_ = P[s], m_ = M[s], z_ = Z[s];
            P[s] = a * p_ + b * m_ + u * z_;
            M[s] = a * m_ + b * p_ + u * z_;
            Z[s] = w * z_ + h * (p_ + m_);
          })
        }
      } break;
      case EPGX_OP_E: {
        const real *c0 = coef + off0 + patoff[pat0];
        const real e1 = ldc(c0), z0 = ldc(c0 + 1) * m0;
        const real e2 = ldc(coef + off1 + patoff[pat1]);
        if (mine) {
          DUFF(nslot, { P[s] *= e2; M[s] *= e2; Z[s] *= e1; })
          if (aff0 && q == 0) Z[0] += z0;
        }
      } break;
      case EPGX_OP_DIAG: { // real entries only (checked by the host): (aP, 0, aM, 0, aZ, 0, a0, 0)
        const real *c = coef + off0 + patoff[pat0];
        const real dp = ldc(c), dm = ldc(c + 2), dz = ldc(c + 4), z0 = ldc(c + 6) * m0;
        if (inject) {
          PUBLISH_ALL()
          if (target) {
            DUFF(nslot, {
              const int k = ORDER_OF(s);
              P[s] += dp * xP[k]; M[s] += dm * xM[k]; Z[s] += dz * xZ[k];
            })
            if (aff0) Z[0] += z0;
          }
        } else if (mine) {
          DUFF(nslot, { P[s] *= dp; M[s] *= dm; Z[s] *= dz; })
          if (aff0 && q == 0) Z[0] += z0;
        }
      } break;
      case EPGX_OP_D: {
        const real *c = coef + off0 + patoff[pat0];
        if (mine) {
          DUFF(nslot, {
            const int k = min(ORDER_OF(s), p.C - 1);
            P[s] *= ldc(c + 3 * k); M[s] *= ldc(c + 3 * k + 1); Z[s] *= ldc(c + 3 * k + 2);
          })
        }
      } break;
      case EPGX_OP_SPOIL:
        if (mine) { DUFF(nslot, { P[s] = real(0); M[s] = real(0); }) }
        break;
      case EPGX_OP_PD:
        m0 = ldc(coef + off0 + patoff[pat0]);
        break;
      case EPGX_OP_ADC:
        if (lane == 0 && mine) {
          real fr = real(1), fi = real(0);
          if (flags & EPGX_FLAG_SCALE) {
            const real *c = coef + off0 + patoff[pat0];
            fr = ldc(c); fi = ldc(c + 1);
          }
          const real x = (flags & EPGX_FLAG_Z0) ? Z[0] : P[0];
          if (q == 0) sig[(long long)aux * p.sig_stride] = real2{x * fr, x * fi};
          else if (q - 1 < p.nvar) jac[((long long)r1.z * p.nvar + q - 1) * p.jac_stride] = real2{x * fr, x * fi};
        }
        break;
      case EPGX_OP_SEG:
        DO_SEG((int)off0, (int)off1, (int)off2, r1.z, aux)
        break;
      case EPGX_OP_TRJ: { // one whole-TR group outside a pure window: canonical roles, generic close
        const real *cf = cw + (r / 5) * SJ_ROW;
        const int4 q0 = tb[2 * r + 2], q1 = tb[2 * r + 3];
        PUBLISH_ALL()
        (void)xP; (void)xM; (void)xZ;
        if (nslot > 0) {
#define SJA(K_) case K_: sj_apply<real, NS, K_>(P, M, Z, cf, xb, q, lane, 0); break;
          switch (nslot >> 1) {
            SJA(1) SJA(2) SJA(3) SJA(4) SJA(5) SJA(6) SJA(7) SJA(8) SJA(9) SJA(10) SJA(11) SJA(12) SJA(13) SJA(14) SJA(15) SJA(16)
          default: break;
          }
#undef SJA
        }
        if (lane == 0) {
          if (nslot > 0) {
            const real *ca = cf + (q == 0 ? 0 : 8);
            const real f = ca[5] * m0;
            P[0] += f; M[0] += f; Z[0] += ca[6] * m0;
          }
          if (q == 0) sig[(long long)q0.y * p.sig_stride] = real2{P[0], real(0)};
          else if ((flags & EPGX_FLAG_PARTIALS) && q - 1 < p.nvar) jac[((long long)q1.w * p.nvar + q - 1) * p.jac_stride] = real2{P[0], real(0)};
        }
        const int segw = (q0.x >> 16) & 0xffff; // (shift + 1) | segment flags << 2
        DO_SEG((segw & 3) - 1, (int)((unsigned)q1.x >> 16), (int)((unsigned)q1.x & 0xffff), segw >> 2, q1.z)
        r += 4;
      } break;
      default:
        break;
      }
    }
  }
#undef PUBLISH_ALL
#undef DO_SEG
#undef SHIFT_REAL
#undef DUFFP
#undef PAIR_CASE
#undef DUFF
#undef SLOT_CASE
#undef ORDER_OF
#undef SLOTS_FOR
}

} // namespace epgx
