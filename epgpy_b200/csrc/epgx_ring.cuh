// epgx_ring.cuh -- the general fused EPG kernel: state sets of an atom live in shared-memory ring
// buffers (half storage, orders k >= 0), G lanes of a CTA sweep the orders of one atom.
//
//  * a SEGMENT (all records between two shifts) is one pass: each lane pulls its order(s) into
//    registers, applies every record of the segment, writes them back -- one smem round trip per
//    segment, not per operator (the reference makes 1-4 HBM passes per operator,
//    epgpy/opscalar.py:213-232, opmatrix.py:199-221);
//  * a unit shift S(+-1) (epgpy/shift.py:283-292) is a rotation of two ring offsets plus a boundary
//    fix-up F+(0) <- conj(F-(1)) (from the symmetry of epgpy/statematrix.py:418-421): no copy;
//  * order-1 partial states (epgpy/diff.py:264-288) are NVT more state sets of the same atom,
//    updated in the same pass; variables beyond NVT are tiled over blockIdx.y;
//  * only ADC samples leave the chip (epgpy/probe.py:138-165).
#pragma once
#include "epgx_common.cuh"

namespace epgx {

template <typename real, int NP, int NVT>
__global__ void __launch_bounds__(256) ring_kernel(const KParams p) {
  typedef typename vec2<real>::type real2;
  constexpr int NSET = 1 + NVT;
  extern __shared__ __align__(16) unsigned char smem_raw[];

  const int G = p.G, C = p.C;
  const int tid = threadIdx.x;
  const int al = tid / G;
  const int lane = tid - al * G;
  const long long a_rel = (long long)blockIdx.x * p.A + al;
  const bool valid = a_rel < p.atom_count;
  const long long atom = p.atom_begin + (valid ? a_rel : p.atom_count - 1);
  // variables resident in this tile (set s holds variable tv[s - 1]; -1: empty slot)
  int tv[NVT > 0 ? NVT : 1];
#pragma unroll
  for (int s = 0; s < NVT; ++s) {
    if (p.tiles != nullptr) tv[s] = s < 3 ? __ldg(p.tiles + blockIdx.y * 3 + s) : -1;
    else tv[s] = (int)(blockIdx.y * NVT + s) < p.nvar ? (int)(blockIdx.y * NVT + s) : -1;
  }
  // set (1..NVT) that holds variable v, 0 if it is not resident
  auto set_of = [&](int v) {
    int r = 0;
#pragma unroll
    for (int s = 0; s < NVT; ++s)
      if (tv[s] == v && v >= 0) r = s + 1;
    return r;
  };
  const real *__restrict__ coef = (const real *)p.coef;

  // shared memory: rings [A][NSET][NP][3][C] complex (+ one ring per atom as the temporary of lattice gathers), then
  // pattern offsets [A][npattern] int
  real2 *rings = (real2 *)smem_raw;
  real2 *gtmp = rings + (size_t)p.A * NSET * NP * 3 * C + (size_t)al * C;
  int *patoff_all = (int *)(rings + (size_t)p.A * (NSET * NP * 3 + (p.lattice ? 1 : 0)) * C);
  real2 *ring = rings + (size_t)al * NSET * NP * 3 * C;
  int *patoff = patoff_all + al * p.npattern;
#define RING(set, pool, comp) (ring + (((set) * NP + (pool)) * 3 + (comp)) * C)

  // ---- pattern offsets of this atom (left-aligned broadcasting -> strides, common.py:273-334)
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
    const real2 z = {real(0), real(0)};
    for (int i = lane; i < NSET * NP * 3 * C; i += G) ring[i] = z;
  }
  if (G > 32) __syncthreads(); else __syncwarp();
#define POFF(pat, pool) (patoff[pat] + (pool) * p.pats[(pat) * (EPGX_MAX_DIMS + 1) + EPGX_MAX_DIMS])

  // ---- initial state (statematrix.py:12-80) and equilibrium density
  real m0[NP];
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    m0[q] = ldc(coef + p.m0_off + POFF(p.m0_pat, q));
    const real *ib = coef + p.init_off + POFF(p.init_pat, q);
    for (int k = lane; k <= p.init_n; k += G) {
      RING(0, q, 0)[k] = real2{ldc(ib + 6 * k), ldc(ib + 6 * k + 1)};
      RING(0, q, 1)[k] = real2{ldc(ib + 6 * k + 2), ldc(ib + 6 * k + 3)};
      RING(0, q, 2)[k] = real2{ldc(ib + 6 * k + 4), ldc(ib + 6 * k + 5)};
    }
  }
  if (G > 32) __syncthreads(); else __syncwarp();

  int baseP = 0, baseM = 0; // logical order k of F+ sits at ring slot (baseP + k) mod C
  bool alive = false;       // has a derivative been injected into this tile's partial states yet?
  real2 *sig = (real2 *)p.signal;
  real2 *jac = (real2 *)p.jac;

  for (int sg = 0; sg < p.nseg; ++sg) {
    const int4 s0 = __ldg((const int4 *)(p.segs + sg));
    const int4 s1 = __ldg((const int4 *)(p.segs + sg) + 1);
    const int first = s0.x, count = s0.y, nact = s0.z, shift = s0.w;
    const int n_old = s1.x, n_new = s1.y, sflags = s1.z;
    const int kz = (sflags & EPGX_SEG_LATTICE) ? (sflags >> 16) : 0; // slot of the order k = 0

    // the partial states of this variable tile are exactly zero until its first injection (per-pulse variables: most
    // of the sequence for the late tiles): linear operators leave them zero, so they are skipped.  Decided per
    // segment by every thread alike (a lane without an order in this segment must learn it too)
    if (NVT > 0 && !alive)
      for (int r = first; r < first + count; ++r) {
        const int2 h = __ldg((const int2 *)(p.ops + r));
        if (((h.x >> 16) & EPGX_FLAG_INJECT) && set_of(h.y) >= 1) alive = true;
      }
    if (count > 0) {
      for (int k = lane; k <= nact; k += G) {
        int iP = baseP + k; if (iP >= C) iP -= C;
        int iM = baseM + k; if (iM >= C) iM -= C;
        Tri<real> st[NSET][NP];
#pragma unroll
        for (int s = 0; s < NSET; ++s)
#pragma unroll
          for (int q = 0; q < NP; ++q) {
            const real2 a = RING(s, q, 0)[iP], b = RING(s, q, 1)[iM], c = RING(s, q, 2)[k];
            st[s][q].pr = a.x; st[s][q].pi = a.y; st[s][q].mr = b.x; st[s][q].mi = b.y;
            st[s][q].zr = c.x; st[s][q].zi = c.y;
          }

        for (int r = first; r < first + count; ++r) {
          const int4 r0 = __ldg((const int4 *)(p.ops + r));
          const int4 r1 = __ldg((const int4 *)(p.ops + r) + 1);
          const int code = r0.x & 0xffff, flags = (r0.x >> 16) & 0xffff, aux = r0.y;
          const unsigned off0 = (unsigned)r0.z, off1 = (unsigned)r0.w, off2 = (unsigned)r1.x;
          const int pat0 = r1.y & 0xff, pat1 = (r1.y >> 8) & 0xff, pat2 = (r1.y >> 16) & 0xff;
          const int aux1 = r1.z;
          const bool inject = flags & EPGX_FLAG_INJECT;
          const int iset = inject ? set_of(aux) : 0;                  // target set of an injection (0: not resident)
          const int isrc = inject && aux1 > 0 ? set_of(aux1 - 1) : 0; // its source set: 0 = the base state
          const bool src_ok = !inject || aux1 == 0 || isrc >= 1;
          const bool on_base = flags & EPGX_FLAG_BASE, on_part = (flags & EPGX_FLAG_PARTIALS) && alive;
          // order-1 / order-2 partial states only (EPGX_FLAG_P1 / P2)
          bool sel[NSET];
          sel[0] = on_base;
#pragma unroll
          for (int s = 1; s < NSET; ++s)
            sel[s] = on_part && tv[s - 1] >= 0 && !((flags & EPGX_FLAG_P1) && tv[s - 1] >= p.nvar1) &&
                     !((flags & EPGX_FLAG_P2) && tv[s - 1] < p.nvar1);
          const bool aff = (flags & EPGX_FLAG_AFFINE) && k == kz && isrc == 0;

          // linear forms: out = form(in) for the selected sets, or partial += form(base)
#define APPLY_FORM(EXPR, AFFINE_STMT)                                         \
  {                                                                            \
    if (inject) {                                                              \
      if (iset >= 1 && iset < NSET && src_ok) {                                \
        Tri<real> s_ = st[0][q];                                               \
        _Pragma("unroll") for (int s = 1; s < NSET; ++s) if (s == isrc) s_ = st[s][q]; \
        Tri<real> o_ = EXPR;                                                   \
        if (aff) { AFFINE_STMT; }                                              \
        _Pragma("unroll") for (int s = 1; s < NSET; ++s) if (s == iset) tri_add(st[s][q], o_); \
      }                                                                        \
    } else {                                                                   \
      _Pragma("unroll") for (int s = 0; s < NSET; ++s) {                       \
        if (sel[s]) {                                                          \
          const Tri<real> s_ = st[s][q];                                       \
          Tri<real> o_ = EXPR;                                                 \
          if (s == 0 && aff) { AFFINE_STMT; }                                  \
          st[s][q] = o_;                                                       \
        }                                                                      \
      }                                                                        \
    }                                                                          \
  }

          switch (code) {
          case EPGX_OP_T_GEN:
#pragma unroll
            for (int q = 0; q < NP; ++q) {
              const real *c = coef + off0 + POFF(pat0, q);
              const real a = ldc(c), w = ldc(c + 1), Br = ldc(c + 2), Bi = ldc(c + 3), Ur = ldc(c + 4), Ui = ldc(c + 5);
              APPLY_FORM(form_t_gen(s_, a, w, Br, Bi, Ur, Ui), )
            }
            break;
          case EPGX_OP_T_RE:
#pragma unroll
            for (int q = 0; q < NP; ++q) {
              const real *c = coef + off0 + POFF(pat0, q);
              const real a = ldc(c), w = ldc(c + 1), b = ldc(c + 2), u = ldc(c + 3);
              APPLY_FORM(form_t_re(s_, a, w, b, u), )
            }
            break;
          case EPGX_OP_T_IM:
#pragma unroll
            for (int q = 0; q < NP; ++q) {
              const real *c = coef + off0 + POFF(pat0, q);
              const real a = ldc(c), w = ldc(c + 1), b = ldc(c + 2), u = ldc(c + 3);
              APPLY_FORM(form_t_im(s_, a, w, b, u), )
            }
            break;
          case EPGX_OP_E:
#pragma unroll
            for (int q = 0; q < NP; ++q) {
              const real *c0 = coef + off0 + POFF(pat0, q);
              const real e1 = ldc(c0), r0v = ldc(c0 + 1);
              const real e2 = ldc(coef + off1 + POFF(pat1, q));
              if (flags & EPGX_FLAG_G) {
                const real *c2 = coef + off2 + POFF(pat2, q);
                const real er = e2 * ldc(c2), ei = e2 * ldc(c2 + 1);
                APPLY_FORM(form_e_g(s_, e1, er, ei), o_.zr += r0v * m0[q])
              } else {
                APPLY_FORM(form_e(s_, e1, e2), o_.zr += r0v * m0[q])
              }
            }
            break;
          case EPGX_OP_DIAG:
#pragma unroll
            for (int q = 0; q < NP; ++q) {
              const real *c = coef + off0 + POFF(pat0, q);
              real d[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) d[i] = ldc(c + i);
              APPLY_FORM(form_diag(s_, d), { o_.zr += d[6] * m0[q]; o_.zi += d[7] * m0[q]; })
            }
            break;
          case EPGX_OP_MATRIX:
#pragma unroll
            for (int q = 0; q < NP; ++q) {
              const real *c = coef + off0 + POFF(pat0, q);
              real m[18];
#pragma unroll
              for (int i = 0; i < 18; ++i) m[i] = ldc(c + i);
              real v[6] = {0, 0, 0, 0, 0, 0};
              if (aff) {
                const real *c1 = coef + off1 + POFF(pat1, q);
#pragma unroll
                for (int i = 0; i < 6; ++i) v[i] = ldc(c1 + i) * m0[q];
              }
              APPLY_FORM(form_matrix(s_, m), {
                o_.pr += v[0]; o_.pi += v[1]; o_.mr += v[2]; o_.mi += v[3]; o_.zr += v[4]; o_.zi += v[5];
              })
            }
            break;
          case EPGX_OP_D:
#pragma unroll
            for (int q = 0; q < NP; ++q) {
              const real *c = coef + off0 + POFF(pat0, q) + 3 * k;
              const real dp = ldc(c), dm = ldc(c + 1), dl = ldc(c + 2);
#pragma unroll
              for (int s = 0; s < NSET; ++s)
                if (sel[s]) {
                  st[s][q].pr *= dp; st[s][q].pi *= dp; st[s][q].mr *= dm; st[s][q].mi *= dm;
                  st[s][q].zr *= dl; st[s][q].zi *= dl;
                }
            }
            break;
          case EPGX_OP_X: {
            // s_c <- m_c (s_c - eq_c) + eq_c over the pool axis (exchange.py:102-119)
            const real *c = coef + off0 + patoff[pat0];
            real mt[NP * NP * 2], ml[NP * NP * 2];
#pragma unroll
            for (int i = 0; i < NP * NP * 2; ++i) { mt[i] = ldc(c + i); ml[i] = ldc(c + NP * NP * 2 + i); }
#pragma unroll
            for (int s = 0; s < NSET; ++s)
              if (sel[s]) {
                Tri<real> o[NP];
#pragma unroll
                for (int i = 0; i < NP; ++i) {
                  tri_zero(o[i]);
#pragma unroll
                  for (int j = 0; j < NP; ++j) {
                    const real tr = mt[(i * NP + j) * 2], ti = mt[(i * NP + j) * 2 + 1];
                    const real lr = ml[(i * NP + j) * 2], li = ml[(i * NP + j) * 2 + 1];
                    const Tri<real> &x = st[s][j];
                    const real zr = x.zr - ((s == 0 && k == kz) ? m0[j] : real(0));
                    o[i].pr += tr * x.pr - ti * x.pi;
                    o[i].pi += tr * x.pi + ti * x.pr;
                    o[i].mr += tr * x.mr + ti * x.mi;
                    o[i].mi += tr * x.mi - ti * x.mr;
                    o[i].zr += lr * zr - li * x.zi;
                    o[i].zi += lr * x.zi + li * zr;
                  }
                  if (s == 0 && k == kz) o[i].zr += m0[i];
                }
#pragma unroll
                for (int i = 0; i < NP; ++i) st[s][i] = o[i];
              }
          } break;
          case EPGX_OP_FUSED: {
            const int4 q0 = __ldg((const int4 *)(p.ops + r + 1));
            const int4 q1 = __ldg((const int4 *)(p.ops + r + 1) + 1);
#pragma unroll
            for (int q = 0; q < NP; ++q) {
              const real *ca = coef + off1 + POFF(pat1, q);
              const real *cb = coef + (unsigned)q0.z + POFF(q1.y & 0xff, q);
              real ta, tw, tBr, tBi, tUr, tUi;
              fused_pulse<real>(coef + off0 + POFF(pat0, q), flags, ta, tw, tBr, tBi, tUr, tUi);
              const Fused8<real> f =
                  fuse8<real>(ta, tw, tBr, tBi, tUr, tUi, flags & EPGX_FLAG_PRE, ldc(ca), ldc(ca + 1),
                              ldc(coef + off2 + POFF(pat2, q)), flags & EPGX_FLAG_POST, ldc(cb), ldc(cb + 1),
                              ldc(coef + (unsigned)q0.w + POFF((q1.y >> 8) & 0xff, q)), m0[q]);
              if (on_base) {
                const Tri<real> s_ = st[0][q];
                st[0][q] = form_t8(s_, f);
                if (k == kz) {
                  st[0][q].pr += f.fzr; st[0][q].pi += f.fzi; st[0][q].mr += f.fzr; st[0][q].mi -= f.fzi;
                  st[0][q].zr += f.zz;
                }
              }
            }
            ++r; // the CONT record
          } break;
          case EPGX_OP_SPOIL:
#pragma unroll
            for (int q = 0; q < NP; ++q)
#pragma unroll
              for (int s = 0; s < NSET; ++s)
                if (sel[s]) st[s][q].pr = st[s][q].pi = st[s][q].mr = st[s][q].mi = real(0);
            break;
          case EPGX_OP_PD:
#pragma unroll
            for (int q = 0; q < NP; ++q) m0[q] = ldc(coef + off0 + POFF(pat0, q));
            break;
          case EPGX_OP_ADC:
            if (k == ((flags & EPGX_FLAG_SLOT) ? aux1 : kz) && valid) { // (EPGX_FLAG_SLOT: any lattice slot, base state only)
#pragma unroll
              for (int q = 0; q < NP; ++q) {
                real fr = real(1), fi = real(0);
                if (flags & EPGX_FLAG_SCALE) {
                  const real *c = coef + off0 + POFF(pat0, q);
                  fr = ldc(c); fi = ldc(c + 1);
                }
                const bool z0 = flags & EPGX_FLAG_Z0;
                if (on_base && blockIdx.y == 0) {
                  const real xr = z0 ? st[0][q].zr : st[0][q].pr, xi = z0 ? st[0][q].zi : st[0][q].pi;
                  sig[((long long)aux * p.sig_stride + a_rel) * NP + q] = real2{xr * fr - xi * fi, xr * fi + xi * fr};
                }
                if (flags & EPGX_FLAG_PARTIALS) {
#pragma unroll
                  for (int s = 1; s < NSET; ++s) {
                    const int v = tv[s - 1];
                    if (v >= 0 && v < p.nvar) {
                      const real xr = z0 ? st[s][q].zr : st[s][q].pr, xi = z0 ? st[s][q].zi : st[s][q].pi;
                      jac[(((long long)aux1 * p.nvar + v) * p.jac_stride + a_rel) * NP + q] =
                          real2{xr * fr - xi * fi, xr * fi + xi * fr};
                    }
                  }
                }
              }
            }
            break;
          default:
            break;
          }
#undef APPLY_FORM
        }

#pragma unroll
        for (int s = 0; s < NSET; ++s)
#pragma unroll
          for (int q = 0; q < NP; ++q) {
            RING(s, q, 0)[iP] = real2{st[s][q].pr, st[s][q].pi};
            RING(s, q, 1)[iM] = real2{st[s][q].mr, st[s][q].mi};
            RING(s, q, 2)[k] = real2{st[s][q].zr, st[s][q].zi};
          }
      }
    }

    // PD changes m0 inside a pass executed only by lanes with an order to process: replay it for all
    // (cheap, uniform): every lane must hold the same m0.
    if (count > 0) {
      for (int r = first; r < first + count; ++r) {
        const int4 r0 = __ldg((const int4 *)(p.ops + r));
        if ((r0.x & 0xffff) == EPGX_OP_PD) {
          const int4 r1 = __ldg((const int4 *)(p.ops + r) + 1);
#pragma unroll
          for (int q = 0; q < NP; ++q) m0[q] = ldc(coef + (unsigned)r0.z + POFF(r1.y & 0xff, q));
        }
      }
    }

    if (shift != 0 || (sflags & EPGX_SEG_RESET)) {
      if (G > 32) __syncthreads(); else __syncwarp();
      if (sflags & EPGX_SEG_RESET) {
        // state <- equilibrium, order <- 0 (operator.py:297-304); partial sets are cleared
        const real2 z = {real(0), real(0)};
        for (int i = lane; i < NSET * NP * 3 * C; i += G) ring[i] = z;
        if (G > 32) __syncthreads(); else __syncwarp();
        baseP = baseM = 0;
        if (lane == 0) {
#pragma unroll
          for (int q = 0; q < NP; ++q) RING(0, q, 2)[0] = real2{m0[q], real(0)};
        }
      } else if (shift == 2) {
        // lattice gather (EPGX_SEG_LATTICE): new slot j <- old slot map[j], per component, through the temporary ring
        const int nn = n_new + 1;
        const int *mp = p.maps + s1.w;
        const real2 z = {real(0), real(0)};
#pragma unroll 1
        for (int sq = 0; sq < NSET * NP; ++sq)
#pragma unroll 1
          for (int c = 0; c < 3; ++c) {
            real2 *r = ring + ((size_t)sq * 3 + c) * C;
            for (int j = lane; j < nn; j += G) {
              const int src = __ldg(mp + c * nn + j);
              gtmp[j] = src >= 0 ? r[src] : z;
            }
            if (G > 32) __syncthreads(); else __syncwarp();
            for (int j = lane; j < C; j += G) r[j] = j < nn ? gtmp[j] : z;
            if (G > 32) __syncthreads(); else __syncwarp();
          }
      } else if (shift > 0) {
        // F+(k) <- F+(k-1), F+(0) <- conj(F-(1)), F-(k) <- F-(k+1)
        if (lane == 0) {
          int i1 = baseM + 1; if (i1 >= C) i1 -= C;
          const int nbP = baseP == 0 ? C - 1 : baseP - 1;
          const int nbM = i1;
#pragma unroll
          for (int s = 0; s < NSET; ++s)
#pragma unroll
            for (int q = 0; q < NP; ++q) {
              real2 v = {real(0), real(0)};
              if (n_old >= 1) { v = RING(s, q, 1)[i1]; v.y = -v.y; }
              RING(s, q, 0)[nbP] = v;
              for (int k = n_old; k <= n_new; ++k) {
                int i = nbM + k; if (i >= C) i -= C;
                RING(s, q, 1)[i] = real2{real(0), real(0)};
              }
            }
        }
        baseP = baseP == 0 ? C - 1 : baseP - 1;
        baseM = baseM + 1 == C ? 0 : baseM + 1;
      } else {
        if (lane == 0) {
          int i1 = baseP + 1; if (i1 >= C) i1 -= C;
          const int nbM = baseM == 0 ? C - 1 : baseM - 1;
          const int nbP = i1;
#pragma unroll
          for (int s = 0; s < NSET; ++s)
#pragma unroll
            for (int q = 0; q < NP; ++q) {
              real2 v = {real(0), real(0)};
              if (n_old >= 1) { v = RING(s, q, 0)[i1]; v.y = -v.y; }
              RING(s, q, 1)[nbM] = v;
              for (int k = n_old; k <= n_new; ++k) {
                int i = nbP + k; if (i >= C) i -= C;
                RING(s, q, 0)[i] = real2{real(0), real(0)};
              }
            }
        }
        baseM = baseM == 0 ? C - 1 : baseM - 1;
        baseP = baseP + 1 == C ? 0 : baseP + 1;
      }
      if (G > 32) __syncthreads(); else __syncwarp();
    }
  }
  // ---- final-state read-back (epgx_simulate_state): the base state in logical order, half storage
  if (p.state != nullptr && blockIdx.y == 0) {
    if (G > 32) __syncthreads(); else __syncwarp();
    if (valid) {
      real2 *out = (real2 *)p.state + (size_t)a_rel * NP * C * 3;
#pragma unroll
      for (int q = 0; q < NP; ++q)
        for (int k = lane; k < C; k += G) {
          int iP = baseP + k; if (iP >= C) iP -= C;
          int iM = baseM + k; if (iM >= C) iM -= C;
          real2 *o = out + ((size_t)q * C + k) * 3;
          o[0] = RING(0, q, 0)[iP]; o[1] = RING(0, q, 1)[iM]; o[2] = RING(0, q, 2)[k];
        }
    }
  }
#undef RING
#undef POFF
}

} // namespace epgx
