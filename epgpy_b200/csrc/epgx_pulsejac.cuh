// epgx_pulsejac.cuh -- one THREAD per state set: per-pulse derivative variables on bounded, real-valued phase graphs.
//
// The reference's sequence optimisation differentiates the signal with respect to the flip angle of EVERY pulse
// (examples/differentiation/optim_mrf.py:96-149: 400 TRs, 800 variables, max_nstate = 10; epgpy/diff.py:264-288).  Such
// a variable is injected exactly once -- at its pulse -- and from then on its partial state obeys the same linear
// recurrence as the base state, affine terms dropped.  With a few orders per state (max_nstate + 1 <= NO <= 16) a whole
// state set fits the registers of ONE thread, so
//   * thread v of an atom runs the base recurrence until the injection of variable v, turns its state into the partial
//     state there (x_v = form(x_0)) and goes on with the partial recurrence: no base state is shared, published or
//     synchronised, and nothing is recomputed per variable tile beyond that prefix;
//   * a unit shift is a renaming of the thread's own registers (no shuffle, no shared memory);
//   * all threads of a CTA belong to one atom: records and coefficients are CTA-uniform loads, every thread has
//     3 NO independent FMA chains per operator;
//   * thread nvar of an atom never injects: it is the base state and writes the signal rows.
// The variable-tiled kernels (epgx_realjac.cuh: 3 variables per CTA, 334 tiles for 1000 variables, the base state
// recomputed and every coefficient re-read per tile) ran this workload at 5 % of the FP64 FMA peak.
//
// Eligibility (epgx.cu): real-valued derivative tape (as epgx_realjac.cuh), one pool, max_order < 16, every variable
// injected by exactly one record, at least four variables.
#pragma once
#include <cuda_pipeline.h>

#include "epgx_common.cuh"

namespace epgx {

constexpr int PJ_THREADS = 128;
constexpr int PJ_WIN = 64; // records per shared-memory window (one int4 per thread and window)

// register budget: four CTAs per SM while the state takes at most 72 registers (12 orders in FP64), three above
constexpr int pj_min_blocks(int state_bytes) { return state_bytes <= 288 ? 4 : 3; }

template <typename real, int NO>
__global__ void __launch_bounds__(PJ_THREADS, pj_min_blocks(3 * NO * sizeof(real))) pulsejac_kernel(const KParams p) {
  typedef typename vec2<real>::type real2;
  __shared__ int patoff[EPGX_MAX_PATTERNS];
  __shared__ int4 srec[3][2 * PJ_WIN];             // record windows w, w + 1, w + 2
  __shared__ __align__(16) real scoef[2][PJ_WIN * 5]; // coefficient entries of the records of windows w, w + 1
  const int tid = threadIdx.x;
  const long long a_rel = blockIdx.x;
  const long long atom = p.atom_begin + a_rel;
  const int vid = blockIdx.y * PJ_THREADS + tid; // variable of this thread; nvar: the base state; beyond: idle
  const real *__restrict__ coef = (const real *)p.coef;
  {
    int idx[EPGX_MAX_DIMS];
    long long r = atom;
    for (int d = p.ndim - 1; d >= 0; --d) {
      idx[d] = (int)(r % p.shape[d]);
      r /= p.shape[d];
    }
    for (int q = tid; q < p.npattern; q += PJ_THREADS) {
      const int *st = p.pats + q * (EPGX_MAX_DIMS + 1);
      int o = 0;
      for (int d = 0; d < p.ndim; ++d) o += idx[d] * st[d];
      patoff[q] = o;
    }
  }
  __syncthreads();
  const bool active = vid <= p.nvar; // (idle threads of the last tile still serve the staging and the barriers)

  real P[NO], M[NO], Z[NO];
  real m0 = ldc(coef + p.m0_off + patoff[p.m0_pat]);
  {
    const real *ib = coef + p.init_off + patoff[p.init_pat];
#pragma unroll
    for (int k = 0; k < NO; ++k) {
      const bool in = k <= p.init_n;
      P[k] = in ? ldc(ib + 6 * k) : real(0);
      M[k] = in ? ldc(ib + 6 * k + 2) : real(0);
      Z[k] = in ? ldc(ib + 6 * k + 4) : real(0);
    }
  }
  bool partial = false; // false: this thread still carries the base state
  real2 *sig = (real2 *)p.signal;
  real2 *jac = (real2 *)p.jac;

  // The merged record stream (records + EPGX_OP_SEG markers, epgx.cu) is staged through shared memory one window
  // ahead, records AND the coefficient entries they point to: window w + 2's records (cp.async, one int4 per thread) and
  // window w + 1's coefficients (thread j gathers the entries of record j with cp.async, which holds no register) are
  // in flight while window w executes from shared memory.  With a handful of CTAs per SM (64 atoms, or ONE atom in the
  // reference's optimisation loop) nothing else hides the two dependent L2 latencies of "read the record, then read
  // what it points to": 1.2 us per record without the staging.
  const int4 *stream = (const int4 *)p.stream;
  const int n = p.nstream, nwin = (n + PJ_WIN - 1) / PJ_WIN;
  auto stage_records = [&](int w) { // window w -> srec[w % 3]
    const int i = 2 * PJ_WIN * w + tid;
    if (tid < 2 * PJ_WIN && i < 2 * n) __pipeline_memcpy_async(&srec[w % 3][tid], stream + i, 16);
  };
  auto stage_coefs = [&](int w) { // coefficient entries of window w (its records are in srec[w % 3]) -> scoef[w & 1]
    const int j = tid;
    if (j < PJ_WIN && w * PJ_WIN + j < n) {
      const int4 w0 = srec[w % 3][2 * j], w1 = srec[w % 3][2 * j + 1];
      const int code = w0.x & 0xffff, flags = (w0.x >> 16) & 0xffff;
      const real *b0 = coef + (unsigned)w0.z + patoff[w1.y & 0xff];
      real *dst = &scoef[w & 1][5 * j];
      int nb = 0, stride = 1;
      switch (code) {
      case EPGX_OP_T_RE: nb = 4; break;
      case EPGX_OP_E: nb = 2; __pipeline_memcpy_async(dst + 2, coef + (unsigned)w0.w + patoff[(w1.y >> 8) & 0xff], sizeof(real)); break;
      case EPGX_OP_DIAG: nb = 4; stride = 2; break; // real-valued: (aP, -, aM, -, aZ, -, a0Z, -)
      case EPGX_OP_ADC: nb = (flags & EPGX_FLAG_SCALE) ? 2 : 0; break;
      case EPGX_OP_PD: nb = 1; break;
      default: break;
      }
      for (int q = 0; q < nb; ++q) __pipeline_memcpy_async(dst + q, b0 + q * stride, sizeof(real));
    }
  };
  stage_records(0);
  stage_records(1);
  __pipeline_commit();
  __pipeline_wait_prior(0);
  __syncthreads();
  stage_coefs(0);
  __pipeline_commit();
  for (int w = 0; w < nwin; ++w) {
    __pipeline_wait_prior(0); // records of window w + 1, coefficients of window w
    __syncthreads();
    if (w + 2 < nwin) stage_records(w + 2);
    if (w + 1 < nwin) stage_coefs(w + 1);
    __pipeline_commit();
    const int4 *rec = srec[w % 3];
    const real *cwin = scoef[w & 1];
    const int cnt = min(PJ_WIN, n - w * PJ_WIN);
    for (int i = 0; i < cnt; ++i) {
      const int4 a0 = rec[2 * i], a1 = rec[2 * i + 1];
      const real *ca = cwin + 5 * i;
    {
      const int code = a0.x & 0xffff, flags = (a0.x >> 16) & 0xffff, aux = a0.y;
      const bool inject = flags & EPGX_FLAG_INJECT;
      // an injection turns the base state of ITS thread into the partial state; other records act on the threads whose
      // kind of state they name
      const bool mine = active && (inject ? (aux == vid && !partial) : (partial ? (flags & EPGX_FLAG_PARTIALS) : (flags & EPGX_FLAG_BASE)));
      const bool aff = (flags & EPGX_FLAG_AFFINE) && !partial;
      switch (code) {
      case EPGX_OP_T_RE: {
        const real a = ca[0], w = ca[1], b = ca[2], u = ca[3], h = real(-0.5) * u, c = a - b;
        if (mine) {
#pragma unroll
          for (int k = 0; k < NO; ++k) { // seven instructions per order: q = b s + u Z shared by F+ and F- (epgx_real.cuh)
            const real p_ = P[k], m_ = M[k], z_ = Z[k], s_ = p_ + m_, q_ = fma(b, s_, u * z_);
            P[k] = fma(c, p_, q_);
            M[k] = fma(c, m_, q_);
            Z[k] = fma(w, z_, h * s_);
          }
        }
      } break;
      case EPGX_OP_E:
        if (mine) {
          const real e1 = ca[0], e2 = ca[2];
#pragma unroll
          for (int k = 0; k < NO; ++k) { P[k] *= e2; M[k] *= e2; Z[k] *= e1; }
          if (aff) Z[0] = fma(ca[1], m0, Z[0]);
        }
        break;
      case EPGX_OP_DIAG: // real-valued: (aP, -, aM, -, aZ, -, a0Z, -)
        if (mine) {
#pragma unroll
          for (int k = 0; k < NO; ++k) { P[k] *= ca[0]; M[k] *= ca[1]; Z[k] *= ca[2]; }
          if (aff) Z[0] = fma(ca[3], m0, Z[0]);
        }
        break;
      case EPGX_OP_D:
        if (mine) {
          const real *c = coef + (unsigned)a0.z + patoff[a1.y & 0xff];
#pragma unroll
          for (int k = 0; k < NO; ++k) {
            const int kk = min(k, p.C - 1);
            P[k] *= ldc(c + 3 * kk); M[k] *= ldc(c + 3 * kk + 1); Z[k] *= ldc(c + 3 * kk + 2);
          }
        }
        break;
      case EPGX_OP_SPOIL:
        if (mine) {
#pragma unroll
          for (int k = 0; k < NO; ++k) P[k] = M[k] = real(0);
        }
        break;
      case EPGX_OP_PD:
        m0 = ca[0];
        break;
      case EPGX_OP_ADC: {
        const real fr = (flags & EPGX_FLAG_SCALE) ? ca[0] : real(1), fi = (flags & EPGX_FLAG_SCALE) ? ca[1] : real(0);
        const real x = (flags & EPGX_FLAG_Z0) ? Z[0] : P[0];
        if ((flags & EPGX_FLAG_BASE) && vid == p.nvar) sig[(long long)aux * p.sig_stride + a_rel] = real2{x * fr, x * fi};
        if ((flags & EPGX_FLAG_PARTIALS) && vid < p.nvar) {
          const real y = partial ? x : real(0); // the partial state is zero until its injection
          jac[((long long)a1.z * p.nvar + vid) * p.jac_stride + a_rel] = real2{y * fr, y * fi};
        }
      } break;
      case EPGX_OP_SEG: { // close the segment: reset / unit shift (a renaming of this thread's registers)
        const int shift = (int)a0.z, n_new = (int)a1.x, sflags = a1.z;
        if (sflags & EPGX_SEG_RESET) {
#pragma unroll
          for (int k = 0; k < NO; ++k) P[k] = M[k] = Z[k] = real(0);
          if (!partial) Z[0] = m0;
        } else if (shift > 0) {
          const real f1 = NO > 1 ? M[NO > 1 ? 1 : 0] : real(0);
#pragma unroll
          for (int k = NO - 1; k >= 1; --k) P[k] = P[k - 1];
          P[0] = f1;
#pragma unroll
          for (int k = 0; k + 1 < NO; ++k) M[k] = M[k + 1];
          M[NO - 1] = real(0);
#pragma unroll
          for (int k = 0; k < NO; ++k)
            if (k > n_new) P[k] = real(0); // truncation at max_nstate / orders beyond the schedule
        } else if (shift < 0) {
          const real f1 = NO > 1 ? P[NO > 1 ? 1 : 0] : real(0);
#pragma unroll
          for (int k = NO - 1; k >= 1; --k) M[k] = M[k - 1];
          M[0] = f1;
#pragma unroll
          for (int k = 0; k + 1 < NO; ++k) P[k] = P[k + 1];
          P[NO - 1] = real(0);
#pragma unroll
          for (int k = 0; k < NO; ++k)
            if (k > n_new) M[k] = real(0);
        }
      } break;
      default:
        break;
      }
      if (inject && mine) partial = true;
    }
    }
  }
}

} // namespace epgx
