// epgx_reg.cuh -- the fast forward kernel: the whole F+/F-/Z state of an atom lives in REGISTERS
// for the entire sequence.
//
//  * G lanes (G = 1..32 inside a warp, or G = 32 W with W warps) own one atom; order k sits in slot
//    k / G of lane k % G (cyclic layout), so the growing triangle of populated orders keeps all lanes
//    busy and a pass touches only the slots s with s*G <= nact (warp-uniform predicate);
//  * every record of a segment is decoded ONCE per lane and applied to up to NS orders with fully
//    unrolled FMA code: coefficient loads and tape decode are amortised over the slots;
//  * the unit shift S(+-1) (epgpy/shift.py:283-292) is a rotate-by-one-lane of the F+ registers
//    (__shfl_sync) with the wrap-around lane moving to the next slot, the mirrored rotation of F-,
//    and F+(0) <- conj(F-(1)) (symmetry of epgpy/statematrix.py:418-421); warps of a multi-warp atom
//    exchange one boundary value per slot through shared memory;
//  * shared memory holds nothing but the per-atom pattern offsets and those boundary values; HBM sees
//    only the coefficient tables (L2-resident) and the ADC samples.
// Forward simulation, one pool.  Partial derivatives and exchange run in the ring kernel.
#pragma once
#include <cuda_pipeline.h>

#include "epgx_common.cuh"

namespace epgx {

constexpr int TAPE_CHUNK = 64; // records per shared-memory tape window

constexpr int TRC_PER_WINDOW = (TAPE_CHUNK - 1) / 3; // whole-TR triples per tape window
constexpr int TRC_REALS = 14;                        // staged coefficients per TR (13 used)

// One tape window of whole-TR triples ([D] . fused E.T.E of any pulse kind . ADC with optional phase . unit shift
// +1) for one atom per warp, K active slots at compile time: the complex counterpart of tr_window in
// epgx_real.cuh.  The per-TR coefficients were decoded, gathered and fused by the lanes in parallel and staged in
// shared memory (cw / ci); each TR reads them back with broadcast loads.
template <typename real, int NS, int K>
__device__ __forceinline__ void trc_window(real (&Pr)[NS], real (&Pi)[NS], real (&Mr)[NS], real (&Mi)[NS], real (&Zr)[NS],
                                           real (&Zi)[NS], const real *cw, const int *ci, int ntr, const real *coef, int C,
                                           int lane, bool valid, bool is_first, bool is_last, int srcUp, int srcDn,
                                           typename vec2<real>::type *sig, long long sig_stride, long long a_rel) {
  typedef typename vec2<real>::type real2;
  const unsigned FULL = 0xffffffffu;
  if constexpr (K <= NS) {
#pragma unroll 1
    for (int j = 0; j < ntr; ++j) {
      const real *c = cw + j * TRC_REALS;
      Fused8<real> f;
      f.a = c[0]; f.w = c[1]; f.Br = c[2]; f.Bi = c[3]; f.Ur = c[4]; f.Ui = c[5]; f.Hr = c[6]; f.Hi = c[7];
      f.fzr = c[8]; f.fzi = c[9]; f.zz = c[10];
      const real fr = c[11], fi = c[12];
      const int row = ci[4 * j], dbase = ci[4 * j + 1];
      if (dbase >= 0) { // diffusion attenuation of the order held by each slot
#pragma unroll
        for (int s = 0; s < K; ++s) {
          const real *d = coef + dbase + 3 * min(s * 32 + lane, C - 1);
          const real dp = ldc(d), dm = ldc(d + 1), dl = ldc(d + 2);
          Pr[s] *= dp; Pi[s] *= dp; Mr[s] *= dm; Mi[s] *= dm; Zr[s] *= dl; Zi[s] *= dl;
        }
      }
#pragma unroll
      for (int s = 0; s < K; ++s) {
        const Tri<real> t_ = {Pr[s], Pi[s], Mr[s], Mi[s], Zr[s], Zi[s]};
        const Tri<real> o_ = form_t8(t_, f);
        Pr[s] = o_.pr; Pi[s] = o_.pi; Mr[s] = o_.mr; Mi[s] = o_.mi; Zr[s] = o_.zr; Zi[s] = o_.zi;
      }
      if (lane == 0) {
        Pr[0] += f.fzr; Pi[0] += f.fzi; Mr[0] += f.fzr; Mi[0] -= f.fzi; Zr[0] += f.zz;
        if (valid) sig[(long long)row * sig_stride + a_rel] = real2{Pr[0] * fr - Pi[0] * fi, Pr[0] * fi + Pi[0] * fr};
      }
      // unit shift +1: F+(k) <- F+(k-1), F+(0) <- conj F-(1), F-(k) <- F-(k+1)
      const real c1r = __shfl_sync(FULL, Mr[0], 1), c1i = -__shfl_sync(FULL, Mi[0], 1);
#pragma unroll
      for (int s = K - 1; s >= 0; --s) {
        const real vr = is_last ? (s > 0 ? Pr[s > 0 ? s - 1 : 0] : c1r) : Pr[s];
        const real vi = is_last ? (s > 0 ? Pi[s > 0 ? s - 1 : 0] : c1i) : Pi[s];
        Pr[s] = __shfl_sync(FULL, vr, srcUp);
        Pi[s] = __shfl_sync(FULL, vi, srcUp);
      }
      real kr = real(0), ki = real(0);
#pragma unroll
      for (int s = K - 1; s >= 0; --s) {
        const real cr_ = Mr[s], ci_ = Mi[s];
        Mr[s] = __shfl_sync(FULL, is_first ? kr : cr_, srcDn);
        Mi[s] = __shfl_sync(FULL, is_first ? ki : ci_, srcDn);
        kr = cr_; ki = ci_;
      }
    }
  }
}

// MAXT = 128: instances for CTAs of at most 128 threads with a register budget chosen by measurement (as in
// epgx_real.cuh): REG_MINB_128 CTAs per SM when the state takes 96 registers (FP64 NS = 8, FP32 NS = 16)
#ifndef REG_MINB_128
#define REG_MINB_128 3
#endif
constexpr int reg_min_blocks(int state_bytes, int maxt) {
  return maxt <= 128 ? (state_bytes <= 192 ? 4 : state_bytes <= 384 ? REG_MINB_128 : 1) : (state_bytes <= 192 ? 2 : 1);
}
template <typename real, int NS, int MAXT = 256>
__global__ void __launch_bounds__(MAXT, reg_min_blocks(6 * NS * sizeof(real), MAXT)) reg_kernel(const KParams p) {
  typedef typename vec2<real>::type real2;
  extern __shared__ __align__(16) unsigned char smem_raw[];

  const int G = p.G;
  const int tid = threadIdx.x;
  const int al = tid / G;
  const int lane = tid - al * G;
  const int lw = tid & 31;
  const int W = G > 32 ? G >> 5 : 1;       // warps per atom
  const int GW = G > 32 ? 32 : G;          // lanes of the atom inside this warp
  const int wq = G > 32 ? lane >> 5 : 0;   // warp index inside the atom
  const int gbase = lw & ~(GW - 1);
  const int lq = lw - gbase;               // lane inside the warp-local group
  const int srcUp = gbase | ((lq - 1) & (GW - 1));
  const int srcDn = gbase | ((lq + 1) & (GW - 1));
  const unsigned FULL = 0xffffffffu;
  const bool is_first = lq == 0, is_last = lq == GW - 1;
  int lgG = 0;
  while ((1 << lgG) < G) ++lgG;

  const long long a_rel = (long long)blockIdx.x * p.A + al;
  const bool valid = a_rel < p.atom_count;
  const long long atom = p.atom_begin + (valid ? a_rel : p.atom_count - 1);
  const real *__restrict__ coef = (const real *)p.coef;

  // shared memory: tape window [2][TAPE_CHUNK] records, pattern offsets [A][npattern] int, then
  // (multi-warp atoms) boundary exchange [2 parity][A][W][2 (up, dn)][NS]
  int4 *tbuf = (int4 *)smem_raw;
  int *patoff = (int *)(tbuf + 2 * TAPE_CHUNK * 2) + al * p.npattern;
  real2 *xbuf = (real2 *)((int *)(tbuf + 2 * TAPE_CHUNK * 2) + ((p.A * p.npattern + 3) & ~3));
  // staged whole-TR coefficients of this warp (one atom per warp only): [TRC_PER_WINDOW][TRC_REALS] reals + [..][4] ints
  real *cwbuf = (real *)(xbuf + (G > 32 ? (size_t)2 * p.A * W * 2 * NS : 0)) + (size_t)(tid >> 5) * TRC_PER_WINDOW * TRC_REALS;
  int *cibuf = (int *)((real *)(xbuf + (G > 32 ? (size_t)2 * p.A * W * 2 * NS : 0)) + (size_t)(blockDim.x >> 5) * TRC_PER_WINDOW * TRC_REALS) +
               (size_t)(tid >> 5) * TRC_PER_WINDOW * 4;
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

  real Pr[NS], Pi[NS], Mr[NS], Mi[NS], Zr[NS], Zi[NS];
  real m0 = ldc(coef + p.m0_off + patoff[p.m0_pat]);
  {
    const real *ib = coef + p.init_off + patoff[p.init_pat];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const int k = s * G + lane;
      const bool in = k <= p.init_n;
      Pr[s] = in ? ldc(ib + 6 * k) : real(0);
      Pi[s] = in ? ldc(ib + 6 * k + 1) : real(0);
      Mr[s] = in ? ldc(ib + 6 * k + 2) : real(0);
      Mi[s] = in ? ldc(ib + 6 * k + 3) : real(0);
      Zr[s] = in ? ldc(ib + 6 * k + 4) : real(0);
      Zi[s] = in ? ldc(ib + 6 * k + 5) : real(0);
    }
  }

  real2 *sig = (real2 *)p.signal;
  int parity = 0;

#define LOAD_TRI(s) Tri<real> t_ = {Pr[s], Pi[s], Mr[s], Mi[s], Zr[s], Zi[s]}
#define STORE_TRI(s, o) { Pr[s] = o.pr; Pi[s] = o.pi; Mr[s] = o.mr; Mi[s] = o.mi; Zr[s] = o.zr; Zi[s] = o.zi; }
// Duff-style dispatch: run BODY for slots n-1 .. 0 with ONE branch (the slots are independent, or are
// written so that descending order is the in-place order); `s` is a compile-time constant in BODY
#define SLOT_CASE(K, ...)                      \
  case (K) + 1:                                \
    if (NS > (K)) {                            \
      constexpr int s = (K) < NS ? (K) : 0;    \
      __VA_ARGS__                              \
    }
#define DUFF(n, ...)                                                                                        \
  switch (n) {                                                                                              \
    SLOT_CASE(15, __VA_ARGS__) SLOT_CASE(14, __VA_ARGS__) SLOT_CASE(13, __VA_ARGS__) SLOT_CASE(12, __VA_ARGS__) \
    SLOT_CASE(11, __VA_ARGS__) SLOT_CASE(10, __VA_ARGS__) SLOT_CASE(9, __VA_ARGS__) SLOT_CASE(8, __VA_ARGS__)   \
    SLOT_CASE(7, __VA_ARGS__) SLOT_CASE(6, __VA_ARGS__) SLOT_CASE(5, __VA_ARGS__) SLOT_CASE(4, __VA_ARGS__)     \
    SLOT_CASE(3, __VA_ARGS__) SLOT_CASE(2, __VA_ARGS__) SLOT_CASE(1, __VA_ARGS__) SLOT_CASE(0, __VA_ARGS__)     \
  default:                                                                                                  \
    break;                                                                                                  \
  }
#define FOR_SLOTS(EXPR) DUFF(nslot, { LOAD_TRI(s); const Tri<real> o_ = EXPR; STORE_TRI(s, o_); })

// segment close / open, shared by EPGX_OP_SEG and EPGX_OP_TR
#define SHIFT_HEAD(DR, DI)                                                                                  \
    real cr, ci;                                                                                             \
    if (G == 1) { cr = NS > 1 ? DR[NS > 1 ? 1 : 0] : real(0); ci = NS > 1 ? DI[NS > 1 ? 1 : 0] : real(0); }  \
    else { cr = __shfl_sync(FULL, DR[0], gbase | 1); ci = __shfl_sync(FULL, DI[0], gbase | 1); }             \
    if (n_old < 1) { cr = real(0); ci = real(0); }                                                           \
    ci = -ci;
#define SHIFT_W1(UR, UI, DR, DI)                                                                             \
  {                                                                                                          \
    SHIFT_HEAD(DR, DI)                                                                                       \
    /* up: the LAST lane first takes over the value of its previous slot (order 0's new value for slot 0), \
       then one rotate-by-one-lane delivers every order to its new owner; descending = in place */         \
    DUFF(nsl, {                                                                                              \
      const real vr = is_last ? (s > 0 ? UR[s > 0 ? s - 1 : 0] : cr) : UR[s];                                \
      const real vi = is_last ? (s > 0 ? UI[s > 0 ? s - 1 : 0] : ci) : UI[s];                                \
      UR[s] = __shfl_sync(FULL, vr, srcUp);                                                                  \
      UI[s] = __shfl_sync(FULL, vi, srcUp);                                                                  \
    })                                                                                                       \
    /* dn: rotate the other way; the last lane receives the first lane's value of the NEXT slot (zero     \
       above the populated orders) */                                                                       \
    real nr = real(0), ni = real(0);                                                                         \
    DUFF(nsl, {                                                                                              \
      const real xr = __shfl_sync(FULL, DR[s], srcDn), xi = __shfl_sync(FULL, DI[s], srcDn);                 \
      DR[s] = is_last ? nr : xr;                                                                             \
      DI[s] = is_last ? ni : xi;                                                                             \
      nr = xr; ni = xi;                                                                                      \
    })                                                                                                       \
  }
#define SHIFT_WN(UR, UI, DR, DI)                                                                             \
  {                                                                                                          \
    SHIFT_HEAD(DR, DI)                                                                                       \
    real2 *xb = xbuf + (size_t)(parity * p.A + al) * W * 2 * NS; /* [warp][up | dn][slot] */                 \
    if (lq == 31) { _Pragma("unroll") for (int s = 0; s < NS; ++s) xb[(wq * 2 + 0) * NS + s] = real2{UR[s], UI[s]}; } \
    if (lq == 0) { _Pragma("unroll") for (int s = 0; s < NS; ++s) xb[(wq * 2 + 1) * NS + s] = real2{DR[s], DI[s]}; }  \
    asm volatile("bar.sync %0, %1;" ::"r"(1 + al), "r"(G) : "memory");                                       \
    parity ^= 1;                                                                                             \
    _Pragma("unroll") for (int s = 0; s < NS; ++s) {                                                         \
      if (s >= nsl) break;                                                                                   \
      real vr = __shfl_sync(FULL, UR[s], srcUp), vi = __shfl_sync(FULL, UI[s], srcUp);                       \
      if (lq == 0) {                                                                                         \
        if (wq == 0) { vr = cr; vi = ci; const real2 x = xb[((W - 1) * 2 + 0) * NS + s]; cr = x.x; ci = x.y; } \
        else { const real2 x = xb[((wq - 1) * 2 + 0) * NS + s]; vr = x.x; vi = x.y; }                        \
      }                                                                                                      \
      UR[s] = vr; UI[s] = vi;                                                                                \
    }                                                                                                        \
    _Pragma("unroll") for (int s = 0; s < NS; ++s) {                                                         \
      if (s >= nsl) break;                                                                                   \
      real vr = __shfl_sync(FULL, DR[s], srcDn), vi = __shfl_sync(FULL, DI[s], srcDn);                       \
      if (lq == 31) {                                                                                        \
        if (wq < W - 1) { const real2 x = xb[((wq + 1) * 2 + 1) * NS + s]; vr = x.x; vi = x.y; }             \
        else if (s + 1 < NS) { const real2 x = xb[(0 * 2 + 1) * NS + (s + 1 < NS ? s + 1 : s)]; vr = x.x; vi = x.y; } \
        else { vr = real(0); vi = real(0); }                                                                 \
      }                                                                                                      \
      DR[s] = vr; DI[s] = vi;                                                                                \
    }                                                                                                        \
  }
#define DO_SEG(SHIFT_, NOLD_, NNEW_, SFLAGS_, NEXT_) \
  { \
    const int shift = (SHIFT_), n_old = (NOLD_), n_new = (NNEW_), sflags = (SFLAGS_); \
    const int n_move = max(min(n_new, nact + 1), 0); /* orders above nact + 1 are unobservable: they stay */ \
    nact = (NEXT_); \
    nslot = nact < 0 ? 0 : (nact >> lgG) + 1; \
        if (sflags & EPGX_SEG_RESET) { \
_Pragma("unroll") \
          for (int s = 0; s < NS; ++s) Pr[s] = Pi[s] = Mr[s] = Mi[s] = Zr[s] = Zi[s] = real(0); \
          if (lane == 0) Zr[0] = m0; \
        } else if (shift != 0) { \
      const int nsl = (n_move >> lgG) + 1; \
      if (W == 1) { \
        if (shift > 0) SHIFT_W1(Pr, Pi, Mr, Mi) else SHIFT_W1(Mr, Mi, Pr, Pi) \
      } else { \
        if (shift > 0) SHIFT_WN(Pr, Pi, Mr, Mi) else SHIFT_WN(Mr, Mi, Pr, Pi) \
      } \
          if (sflags & EPGX_SEG_MASK_TOP) { \
_Pragma("unroll") \
            for (int s = 0; s < NS; ++s) \
              if (s * G + lane > n_new) { \
                if (shift > 0) { Pr[s] = real(0); Pi[s] = real(0); } else { Mr[s] = real(0); Mi[s] = real(0); } \
              } \
          } \
        } \
  }

  // ---- the tape is streamed through shared memory in chunks of TAPE_CHUNK records (cp.async, double
  // buffered): a record fetch is two broadcast LDS instead of two dependent global loads
  const int4 *stream = (const int4 *)p.stream;
  const int nthreads = blockDim.x;
  for (int i = tid; i < 2 * TAPE_CHUNK && i < 2 * p.nstream; i += nthreads) __pipeline_memcpy_async(tbuf + i, stream + i, 16);
  __pipeline_commit();
  int nact = -1, nslot = 0;
  for (int base = 0, chunk = 0; base < p.nstream; base += TAPE_CHUNK, ++chunk) {
    __pipeline_wait_prior(0);
    __syncthreads(); // chunk `chunk` has landed; every warp is done with the buffer about to be refilled
    {
      const int nb = base + TAPE_CHUNK;
      int4 *dst = tbuf + ((chunk + 1) & 1) * 2 * TAPE_CHUNK;
      for (int i = tid; i < 2 * TAPE_CHUNK && nb * 2 + i < 2 * p.nstream; i += nthreads)
        __pipeline_memcpy_async(dst + i, stream + (size_t)nb * 2 + i, 16);
      __pipeline_commit();
    }
    const int4 *tb = tbuf + (chunk & 1) * 2 * TAPE_CHUNK;
    const int cnt = min(TAPE_CHUNK, p.nstream - base);
    if (G == 32 && (tb[0].x & EPGX_CHUNK_PURE_TRC)) {
      // ---- fast path: the window holds TRC_PER_WINDOW whole-TR triples.  Phase 1: lane j decodes TR j, gathers and
      // fuses its coefficients and stages them in shared memory; phase 2 (trc_window) runs the TRs in order.
      int need = 0, nnew = 0, nxt = -1;
      if (lw < TRC_PER_WINDOW) {
        const int4 a0 = tb[6 * lw], a1 = tb[6 * lw + 1], b0 = tb[6 * lw + 2], b1 = tb[6 * lw + 3];
        const int4 c0 = tb[6 * lw + 4], c1 = tb[6 * lw + 5];
        const int fl = (a0.x >> 16) & 0xffff, f2 = (c0.x >> 16) & 0xffff;
        const real *ca = coef + (unsigned)a0.w + patoff[(a1.y >> 8) & 0xff];
        const real *cb = coef + (unsigned)b0.z + patoff[b1.y & 0xff];
        real ta, tw, tBr, tBi, tUr, tUi;
        fused_pulse<real>(coef + (unsigned)a0.z + patoff[a1.y & 0xff], fl, ta, tw, tBr, tBi, tUr, tUi);
        const Fused8<real> f = fuse8<real>(ta, tw, tBr, tBi, tUr, tUi, fl & EPGX_FLAG_PRE, ldc(ca), ldc(ca + 1),
                                           ldc(coef + (unsigned)a1.x + patoff[(a1.y >> 16) & 0xff]), fl & EPGX_FLAG_POST,
                                           ldc(cb), ldc(cb + 1), ldc(coef + (unsigned)b0.w + patoff[(b1.y >> 8) & 0xff]), m0);
        real fr = real(1), fi = real(0);
        if (f2 & 1) { const real *cs = coef + (unsigned)c0.z + patoff[c1.y & 0xff]; fr = ldc(cs); fi = ldc(cs + 1); }
        real *c = cwbuf + lw * TRC_REALS;
        c[0] = f.a; c[1] = f.w; c[2] = f.Br; c[3] = f.Bi; c[4] = f.Ur; c[5] = f.Ui; c[6] = f.Hr; c[7] = f.Hi;
        c[8] = f.fzr; c[9] = f.fzi; c[10] = f.zz; c[11] = fr; c[12] = fi;
        cibuf[4 * lw] = b0.y;                                                                  // ADC row
        cibuf[4 * lw + 1] = (f2 & 2) ? (int)((unsigned)c0.w + (unsigned)patoff[(c1.y >> 8) & 0xff]) : -1; // D table
        nnew = (int)((unsigned)b1.x & 0xffff); nxt = b1.z;
        if (lw == TRC_PER_WINDOW - 1) cibuf[4 * lw + 2] = nxt;
      }
      // slots a TR needs: it applies to orders 0..nact and shifts orders 0..min(n_new, nact + 1) (what lies above
      // nact + 1 is unobservable and need not move); nact of TR j is the "next nact" of TR j - 1
      int curv = __shfl_up_sync(FULL, nxt, 1);
      if (lw == 0) curv = nact;
      if (lw < TRC_PER_WINDOW) need = (max(min(nnew, curv + 1), 0) >> 5) + 1;
      need = max(__reduce_max_sync(FULL, need), nslot);
      __syncwarp();
#define TRCW(K_) case K_: trc_window<real, NS, K_>(Pr, Pi, Mr, Mi, Zr, Zi, cwbuf, cibuf, TRC_PER_WINDOW, coef, p.C, lane, valid, is_first, is_last, srcUp, srcDn, sig, p.sig_stride, a_rel); break;
      switch (need) {
        TRCW(1) TRCW(2) TRCW(3) TRCW(4) TRCW(5) TRCW(6) TRCW(7) TRCW(8) TRCW(9) TRCW(10) TRCW(11) TRCW(12) TRCW(13) TRCW(14) TRCW(15) TRCW(16)
      default: break;
      }
#undef TRCW
      nact = cibuf[4 * (TRC_PER_WINDOW - 1) + 2];
      nslot = nact < 0 ? 0 : (nact >> lgG) + 1;
      __syncwarp();
      continue;
    }
    for (int r = 0; r < cnt; ++r) {
      const int4 r0 = tb[2 * r], r1 = tb[2 * r + 1];
      const int code = r0.x & 0xffff, flags = (r0.x >> 16) & 0xffff, aux = r0.y;
      const unsigned off0 = (unsigned)r0.z, off1 = (unsigned)r0.w, off2 = (unsigned)r1.x;
      const int pat0 = r1.y & 0xff, pat1 = (r1.y >> 8) & 0xff, pat2 = (r1.y >> 16) & 0xff;
      const bool aff = (flags & EPGX_FLAG_AFFINE) && lane == 0 && nslot > 0;

      switch (code) {
      case EPGX_OP_T_GEN: {
        const real *c = coef + off0 + patoff[pat0];
        const real a = ldc(c), w = ldc(c + 1), Br = ldc(c + 2), Bi = ldc(c + 3), Ur = ldc(c + 4), Ui = ldc(c + 5);
        FOR_SLOTS(form_t_gen(t_, a, w, Br, Bi, Ur, Ui))
      } break;
      case EPGX_OP_T_RE: {
        const real *c = coef + off0 + patoff[pat0];
        const real a = ldc(c), w = ldc(c + 1), b = ldc(c + 2), u = ldc(c + 3);
        FOR_SLOTS(form_t_re(t_, a, w, b, u))
      } break;
      case EPGX_OP_T_IM: {
        const real *c = coef + off0 + patoff[pat0];
        const real a = ldc(c), w = ldc(c + 1), b = ldc(c + 2), u = ldc(c + 3);
        FOR_SLOTS(form_t_im(t_, a, w, b, u))
      } break;
      case EPGX_OP_E: {
        const real *c0 = coef + off0 + patoff[pat0];
        const real e1 = ldc(c0), r0v = ldc(c0 + 1);
        const real e2 = ldc(coef + off1 + patoff[pat1]);
        if (flags & EPGX_FLAG_G) {
          const real *c2 = coef + off2 + patoff[pat2];
          const real er = e2 * ldc(c2), ei = e2 * ldc(c2 + 1);
          FOR_SLOTS(form_e_g(t_, e1, er, ei))
        } else {
          FOR_SLOTS(form_e(t_, e1, e2))
        }
        if (aff) Zr[0] += r0v * m0;
      } break;
      case EPGX_OP_DIAG: {
        const real *c = coef + off0 + patoff[pat0];
        real d[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = ldc(c + i);
        FOR_SLOTS(form_diag(t_, d))
        if (aff) { Zr[0] += d[6] * m0; Zi[0] += d[7] * m0; }
      } break;
      case EPGX_OP_MATRIX: {
        const real *c = coef + off0 + patoff[pat0];
        real m[18];
#pragma unroll
        for (int i = 0; i < 18; ++i) m[i] = ldc(c + i);
        FOR_SLOTS(form_matrix(t_, m))
        if (aff) {
          const real *c1 = coef + off1 + patoff[pat1];
          Pr[0] += ldc(c1) * m0; Pi[0] += ldc(c1 + 1) * m0; Mr[0] += ldc(c1 + 2) * m0;
          Mi[0] += ldc(c1 + 3) * m0; Zr[0] += ldc(c1 + 4) * m0; Zi[0] += ldc(c1 + 5) * m0;
        }
      } break;
      case EPGX_OP_D: {
        // orders above nact inside the last active slot hold zeros / unobservable values: clamp the row
        const real *c = coef + off0 + patoff[pat0];
        DUFF(nslot, {
          const int k = min(s * G + lane, p.C - 1);
          const real dp = ldc(c + 3 * k), dm = ldc(c + 3 * k + 1), dl = ldc(c + 3 * k + 2);
          Pr[s] *= dp; Pi[s] *= dp; Mr[s] *= dm; Mi[s] *= dm; Zr[s] *= dl; Zi[s] *= dl;
        })
      } break;
      case EPGX_OP_FUSED: {
        const int4 q0 = tb[2 * r + 2], q1 = tb[2 * r + 3]; // the CONT record (never split from FUSED)
        const real *ca = coef + off1 + patoff[pat1];
        const real *cb = coef + (unsigned)q0.z + patoff[q1.y & 0xff];
        real ta, tw, tBr, tBi, tUr, tUi;
        fused_pulse<real>(coef + off0 + patoff[pat0], flags, ta, tw, tBr, tBi, tUr, tUi);
        const Fused8<real> f = fuse8<real>(ta, tw, tBr, tBi, tUr, tUi, flags & EPGX_FLAG_PRE, ldc(ca), ldc(ca + 1),
                                           ldc(coef + off2 + patoff[pat2]), flags & EPGX_FLAG_POST, ldc(cb), ldc(cb + 1),
                                           ldc(coef + (unsigned)q0.w + patoff[(q1.y >> 8) & 0xff]), m0);
        FOR_SLOTS(form_t8(t_, f))
        if (lane == 0 && nslot > 0) { Pr[0] += f.fzr; Pi[0] += f.fzi; Mr[0] += f.fzr; Mi[0] -= f.fzi; Zr[0] += f.zz; }
        ++r; // the CONT record
      } break;
      case EPGX_OP_SPOIL:
        DUFF(nslot, { Pr[s] = Pi[s] = Mr[s] = Mi[s] = real(0); })
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
          const bool z0 = flags & EPGX_FLAG_Z0;
          const real xr = z0 ? Zr[0] : Pr[0], xi = z0 ? Zi[0] : Pi[0];
          sig[(long long)aux * p.sig_stride + a_rel] = real2{xr * fr - xi * fi, xr * fi + xi * fr};
        }
        break;
      case EPGX_OP_SEG: {
        // end of a segment: unit shift / reset of the previous one, then the next pass's order count
        DO_SEG((int)off0, (int)off1, (int)off2, r1.z, aux)
      } break;
      case EPGX_OP_TRC: {
        // [D] . FUSED . ADC(optional scale) . segment close, in three records (see epgx_common.cuh)
        const int4 q0 = tb[2 * r + 2], q1 = tb[2 * r + 3], w0 = tb[2 * r + 4], w1 = tb[2 * r + 5];
        const int f2 = (w0.x >> 16) & 0xffff;
        if (f2 & 2) {
          const real *c = coef + (unsigned)w0.w + patoff[(w1.y >> 8) & 0xff];
          DUFF(nslot, {
            const int k = min(s * G + lane, p.C - 1);
            const real dp = ldc(c + 3 * k), dm = ldc(c + 3 * k + 1), dl = ldc(c + 3 * k + 2);
            Pr[s] *= dp; Pi[s] *= dp; Mr[s] *= dm; Mi[s] *= dm; Zr[s] *= dl; Zi[s] *= dl;
          })
        }
        const real *ca = coef + off1 + patoff[pat1];
        const real *cb = coef + (unsigned)q0.z + patoff[q1.y & 0xff];
        real ta, tw, tBr, tBi, tUr, tUi;
        fused_pulse<real>(coef + off0 + patoff[pat0], flags, ta, tw, tBr, tBi, tUr, tUi);
        const Fused8<real> f = fuse8<real>(ta, tw, tBr, tBi, tUr, tUi, flags & EPGX_FLAG_PRE, ldc(ca), ldc(ca + 1),
                                           ldc(coef + off2 + patoff[pat2]), flags & EPGX_FLAG_POST, ldc(cb), ldc(cb + 1),
                                           ldc(coef + (unsigned)q0.w + patoff[(q1.y >> 8) & 0xff]), m0);
        FOR_SLOTS(form_t8(t_, f))
        if (lane == 0 && nslot > 0) { Pr[0] += f.fzr; Pi[0] += f.fzi; Mr[0] += f.fzr; Mi[0] -= f.fzi; Zr[0] += f.zz; }
        if (lane == 0 && valid) {
          real fr = real(1), fi = real(0);
          if (f2 & 1) { const real *cs = coef + (unsigned)w0.z + patoff[w1.y & 0xff]; fr = ldc(cs); fi = ldc(cs + 1); }
          sig[(long long)q0.y * p.sig_stride + a_rel] = real2{Pr[0] * fr - Pi[0] * fi, Pr[0] * fi + Pi[0] * fr};
        }
        const int segw = (q0.x >> 16) & 0xffff; // (shift + 1) | segment flags << 2
        DO_SEG((segw & 3) - 1, (int)((unsigned)q1.x >> 16), (int)((unsigned)q1.x & 0xffff), segw >> 2, q1.z)
        r += 2;
      } break;
      case EPGX_OP_TR: {
        // one whole TR: FUSED (E.T.E, RE kind) + plain ADC + the segment's unit shift (see epgx.cu)
        const int4 q0 = tb[2 * r + 2], q1 = tb[2 * r + 3];
        const real *ct = coef + off0 + patoff[pat0];
        const real *ca = coef + off1 + patoff[pat1];
        const real *cb = coef + (unsigned)q0.z + patoff[q1.y & 0xff];
        const Fused5<real> f = fuse5<real>(ldc(ct), ldc(ct + 1), ldc(ct + 2), ldc(ct + 3), flags & EPGX_FLAG_PRE, ldc(ca),
                                           ldc(ca + 1), ldc(coef + off2 + patoff[pat2]), flags & EPGX_FLAG_POST, ldc(cb),
                                           ldc(cb + 1), ldc(coef + (unsigned)q0.w + patoff[(q1.y >> 8) & 0xff]), false, m0);
        FOR_SLOTS(form_t5_re(t_, f.a, f.w, f.b, f.u, f.h))
        if (lane == 0 && nslot > 0) { Pr[0] += f.fz; Mr[0] += f.fz; Zr[0] += f.zz; }
        if (lane == 0 && valid) sig[(long long)q0.y * p.sig_stride + a_rel] = real2{Pr[0], Pi[0]};
        const int segw = (q0.x >> 16) & 0xffff; // (shift + 1) | segment flags << 2
        DO_SEG((segw & 3) - 1, (int)((unsigned)q1.x >> 16), (int)((unsigned)q1.x & 0xffff), segw >> 2, q1.z)
        ++r;
      } break;
      default:
        break;
      }
    }
  }
#undef FOR_SLOTS
#undef DO_SEG
#undef SHIFT_HEAD
#undef SHIFT_W1
#undef SHIFT_WN
#undef DUFF
#undef SLOT_CASE
#undef LOAD_TRI
#undef STORE_TRI
}

} // namespace epgx
