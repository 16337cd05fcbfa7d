// epgx.cu -- C ABI of libepgx.so (include/epgx.h): plan management, kernel selection and launch.
// Host-side here is bookkeeping only; every arithmetic step of the EPG path runs in the sm_100a
// kernels of epgx_ring.cuh / epgx_reg.cuh.  There is no CPU fallback.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "epgx_common.cuh"
#include "epgx_launch.h"

using namespace epgx;

static thread_local std::string g_err;

static int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

#define CUDA_TRY(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(EPGX_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));    \
  } while (0)

struct epgx_plan {
  epgx_tape tape; // pointers redirected to the vectors below
  std::vector<epgx_op> ops;
  std::vector<epgx_segment> segs;
  std::vector<double> coef64;
  std::vector<float> coef32;
  std::vector<int> pats; // [npattern][MAX_DIMS+1]
  std::vector<int> tiles; // [ntile][3] variables resident per tile of the shared-memory kernel (order-2 tapes)
  std::vector<int> maps;  // gather maps of the lattice shifts
  bool lattice = false;   // EPGX_SEG_LATTICE segments: the shared-memory kernel only
  bool order2 = false;    // injections from partial states / P1, P2 records: the shared-memory kernel only
  std::vector<epgx_op> stream; // segments + records merged (register kernel)
  int64_t natoms;
  double flops_cplx, flops_real, updates; // executed real flops per atom (complex / real-valued kernels)
  bool realjac_ok; // real-valued graph with real-valued derivative injections
  bool pulsejac_ok = false; // ... whose variables are injected by exactly one record each (per-pulse variables)
  int64_t ntrj = 0; // whole-TR derivative groups (EPGX_OP_TRJ) in the merged stream
  bool bounded = false; // some segment truncates at max_nstate (EPGX_SEG_MASK_TOP)
  bool real_ok; // real-valued phase graph: eligible for the three-reals-per-order kernel
  epgx_config cfg;
  // workspace layout (bytes)
  int64_t off_ops, off_segs, off_pats, off_stream, off_tiles, off_maps, off_coef, ws_bytes;
};

static const int kSmemLimit = 227 * 1024;

static int pow2ceil(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

static int nvt_choice(int nvar, int want, int npool) {
  // instantiated tile sizes of the ring kernel
  const int opts[3] = {0, 1, 3};
  if (nvar == 0) return 0;
  if (npool > 2) return 1; // three / four pools: one resident partial state
  if (want > 0) {
    for (int o : opts)
      if (o == want) return o;
  }
  return nvar == 1 ? 1 : 3;
}

static double form_flops(int code, int flags, int npool) {
  switch (code) {
  case EPGX_OP_T_GEN: return 56;
  case EPGX_OP_T_RE:
  case EPGX_OP_T_IM: return 28;
  case EPGX_OP_FUSED: return 56; // applied in the general (a, w, B, U, H) form by the complex kernels
  case EPGX_OP_E: return (flags & EPGX_FLAG_G) ? 14 : 6;
  case EPGX_OP_DIAG: return 18;
  case EPGX_OP_MATRIX: return 66;
  case EPGX_OP_D: return 6;
  case EPGX_OP_X: return 3.0 * (8.0 * npool - 2.0); // per pool: 3 components x (N cmul + (N-1) cadd)
  default: return 0;
  }
}

static const int kRegSlots[5] = {1, 2, 4, 8, 16};

// flops per order in the real-valued kernels (three reals per order).  A pulse / fused E.T.E takes seven instructions:
// s = F+ + F- (1), u Z (1), q = b s + u Z (2), F+' = c F+ + q (2), F-' = c F- + q (2), h s (1), Z' = w Z + h s (2) = 11
// flops (round 1 counted the row-by-row form: 14); scaled windows (epgx_real.cuh) drop the product u Z and execute 10
static double form_flops_real(int code) {
  switch (code) {
  case EPGX_OP_T_RE:
  case EPGX_OP_FUSED: return 11;
  case EPGX_OP_E:
  case EPGX_OP_D:
  case EPGX_OP_DIAG: return 3;
  default: return 0;
  }
}

static int choose_variant(const epgx_plan *pl, epgx_config &c, int kernel, int lanes, int vars, int atoms) {
  const epgx_tape &t = pl->tape;
  const int rsz = t.dtype == EPGX_F64 ? 8 : 4;
  const int C = t.max_order + 1;
  const bool reg_ok = t.nvar == 0 && t.npool == 1 && !pl->lattice;
  if (kernel == 2 && !reg_ok)
    return fail(EPGX_ERR_UNSUPPORTED, "the register kernel runs forward simulations of one pool only");
  if ((kernel == 4 || kernel == 5) && !pl->realjac_ok)
    return fail(EPGX_ERR_UNSUPPORTED, "the real-valued derivative kernels need a real-valued tape with order-1 variables");
  if (kernel == 5 && t.nvar > 3)
    return fail(EPGX_ERR_UNSUPPORTED, "the warp-per-state-set derivative kernel holds at most three variables");
  // ---- one thread per state set (epgx_pulsejac.cuh): per-pulse variables on bounded graphs
  if (kernel == 6 && !(pl->pulsejac_ok && C <= 16))
    return fail(EPGX_ERR_UNSUPPORTED, "the thread-per-state-set kernel needs a real-valued derivative tape of at most 16 orders "
                                      "whose variables are injected once each");
  if (pl->pulsejac_ok && C <= 16 && (kernel == 6 || (kernel == 0 && t.nvar >= 4))) {
    c.kernel = 5;
    c.lanes_per_atom = 1;
    c.slots_per_lane = C <= 4 ? 4 : C <= 8 ? 8 : C <= 12 ? 12 : 16;
    c.vars_per_pass = 128;
    c.var_tiles = (t.nvar + 1 + 127) / 128;
    c.atoms_per_cta = 1;
    c.threads_per_cta = 128;
    c.smem_bytes = 0;
    c.ring = C;
    return EPGX_OK;
  }
  // ---- one warp per state set (epgx_setjac.cuh): whole-TR derivative groups, 32 NS orders.  Automatic choice when
  // most of the segments are such groups and the orders would not fit one warp of the orders-over-warps kernel
  if (pl->realjac_ok && t.nvar <= 3 && (kernel == 5 || (kernel == 0 && C > 128 && 2 * pl->ntrj >= t.nseg))) {
    const int need = (C + 31) / 32;
    int NS = 0;
    for (int o : {2, 4, 8, 16})
      if (o >= need && !NS) NS = o;
    if (NS) {
      const int W = 1 + t.nvar;
      c.kernel = 4;
      c.lanes_per_atom = 32;
      c.slots_per_lane = NS;
      c.vars_per_pass = t.nvar;
      c.var_tiles = 1;
      c.atoms_per_cta = 1;
      c.threads_per_cta = 32 * W;
      c.smem_bytes = 2 * epgx::kTapeChunk * 32 + ((t.npattern + 3) & ~3) * 4 + 2 * 3 * NS * 32 * rsz +
                     W * (epgx::kTrjPerWindow * epgx::kSjRow + 16) * rsz + 64;
      c.ring = C;
      return EPGX_OK;
    }
    if (kernel == 5) return fail(EPGX_ERR_CAPACITY, "no warp-per-state-set instance holds " + std::to_string(C) + " orders");
  }
  if (pl->realjac_ok && (kernel == 0 || kernel == 4)) {
    // ---- real-valued register kernel with 3 resident partial states: G = 32 W lanes x NS slots >= C orders
    const int ns_max = 4; // 4 slots x (1 + 3) state sets x 3 reals: the most that leaves three (FP64) / four (FP32) CTAs per SM
    int G = lanes > 0 ? pow2ceil(lanes) : pow2ceil((C + ns_max - 1) / ns_max);
    if (G > 256) G = 256;
    const int need = (C + G - 1) / G;
    int NS = 0;
    for (int o : {1, 2, 4, 8})
      if (o >= need && !NS) NS = o;
    if (NS && (NS <= ns_max || lanes > 0)) {
      const int W = G > 32 ? G / 32 : 1;
      int A = atoms > 0 ? atoms : (128 / G > 0 ? 128 / G : 1);
      if (W > 1 && A > 15) A = 15;
      const bool trj = t.nvar <= 3; // whole-TR groups (EPGX_OP_TRJ): per-atom coefficient windows in shared memory
      if (trj && A > 16 && atoms <= 0) A = 16;
      while (A * G > 256 && A > 1) --A;
      while (trj && A > 1 && A * epgx::kTrjPerWindow * epgx::kTrjReals * rsz > 32 * 1024) --A;
      if (A * G <= 256) {
        c.kernel = 3;
        c.lanes_per_atom = G;
        c.slots_per_lane = NS;
        c.vars_per_pass = 3;
        c.var_tiles = (t.nvar + 2) / 3;
        c.atoms_per_cta = A;
        c.threads_per_cta = A * G;
        c.smem_bytes = 2 * epgx::kTapeChunk * 32 + ((A * t.npattern + 3) & ~3) * 4 + (W > 1 ? 2 * A * W * 4 * 2 * (NS + 2) * rsz : 0) +
                       (trj ? A * epgx::kTrjPerWindow * epgx::kTrjReals * rsz : 0) + 32;
        c.ring = C;
        return EPGX_OK;
      }
    }
    if (kernel == 4) return fail(EPGX_ERR_CAPACITY, "no real-derivative-kernel instance holds " + std::to_string(C) + " orders");
  }
  if (kernel == 3 && !pl->real_ok)
    return fail(EPGX_ERR_UNSUPPORTED, "the real-valued kernel needs +-90 degree pulses, no precession and a real initial state");
  if (pl->real_ok && (kernel == 0 || kernel == 3)) {
    // ---- real-valued register kernel: one warp (or less) per atom, up to 32 slots
    const int ns_max = t.dtype == EPGX_F64 ? 16 : 32;
    int G = lanes > 0 ? pow2ceil(lanes) : pow2ceil((C + 7) / 8); // about 8 orders per lane; small graphs pack several atoms per warp
    if (G > 32) G = 32;
    const int need = (C + G - 1) / G;
    int NS = 0;
    for (int o : {2, 4, 8, 16, 32}) // register slots come in pairs: blocks of two orders per lane
      if (o >= need && !NS) NS = o;
    if (NS && NS <= ns_max) {
      int A = atoms > 0 ? atoms : (G == 32 ? 2 : 128 / G); // (one warp per atom: CTAs of two warps, measured 2 % faster than four)
      while (A * G > 256 && A > 1) --A;
      while (G >= 8 && A > 1 && A * 32 * (epgx::kRealWindows + 8 * (G == 32 ? epgx::kRealWindows : 1)) * rsz > 48 * 1024) --A; // staging rows of the whole-TR windows
      c.kernel = 2;
      c.lanes_per_atom = G;
      c.slots_per_lane = NS;
      c.vars_per_pass = 0;
      c.var_tiles = 1;
      c.atoms_per_cta = A;
      c.threads_per_cta = A * G;
      // tape windows (two buffers of kRealWindows host windows), pattern offsets, staging rows + echoes of the whole-TR
      // windows [A][32 kRealWindows][9], a row of zeros (epgx_real.cuh)
      const int kw = G == 32 ? epgx::kRealWindows : 1;
      c.smem_bytes = 2 * epgx::kRealWindows * epgx::kTapeChunk * 32 + ((A * t.npattern + 3) & ~3) * 4 + (G >= 8 ? A * 32 * (epgx::kRealWindows + 8 * kw) * rsz : 0) + 4 * rsz + 32;
      c.ring = C;
      return EPGX_OK;
    }
    if (kernel == 3) return fail(EPGX_ERR_CAPACITY, "no real-kernel instance holds " + std::to_string(C) + " orders");
  }
  if (reg_ok && kernel != 1) {
    // ---- register kernel: G lanes x NS slots >= C orders
    const int ns_max = 16;
    int G;
    if (lanes > 0) {
      G = pow2ceil(lanes);
    } else {
      G = pow2ceil((C + 7) / 8);
      if (G > 32 && C <= 32 * ns_max) G = 32;
    }
    if (G > 256) G = 256;
    int need = (C + G - 1) / G, NS = 0;
    for (int o : kRegSlots)
      if (o >= need && !NS) NS = o;
    if (NS && (NS <= ns_max || lanes > 0 || kernel == 2)) {
      int A = atoms > 0 ? atoms : (128 / G > 0 ? 128 / G : 1);
      if (G > 32 && A > 15) A = 15; // named barriers 1..15
      while (A * G > 256 && A > 1) --A;
      if (A * G <= 256) {
        const int W = G > 32 ? G / 32 : 1;
        c.kernel = 1;
        c.lanes_per_atom = G;
        c.slots_per_lane = NS;
        c.vars_per_pass = 0;
        c.var_tiles = 1;
        c.atoms_per_cta = A;
        c.threads_per_cta = A * G;
        c.smem_bytes = 2 * epgx::kTapeChunk * 32 + ((A * t.npattern + 3) & ~3) * 4 + (W > 1 ? 2 * A * W * 2 * NS * 2 * rsz : 0) +
                       ((A * G + 31) / 32) * epgx::kTrcPerWindow * (epgx::kTrcReals * rsz + 16) + 64;
        c.ring = C;
        return EPGX_OK;
      }
    }
    if (kernel == 2) return fail(EPGX_ERR_CAPACITY, "no register-kernel instance holds " + std::to_string(C) + " orders");
  }
  int G = lanes > 0 ? pow2ceil(lanes) : pow2ceil((C + 3) / 4);
  if (G > 256) G = 256;
  int nvt = nvt_choice(t.nvar, vars, t.npool);
  const bool tiled = !pl->tiles.empty();
  if (tiled) {
    if (t.npool > 2) return fail(EPGX_ERR_UNSUPPORTED, "order-2 partial states with more than two exchange pools");
    nvt = 3; // a pair tile holds (a, b, ab)
  } else if (pl->order2 && (t.nvar > 3 || t.npool > 2)) {
    return fail(EPGX_ERR_INVALID, "order-2 tape without variable tiles");
  } else if (pl->order2) {
    nvt = 3;
  }
  int64_t per_atom;
  for (;;) {
    per_atom = ((int64_t)(1 + nvt) * t.npool * 3 + (pl->lattice ? 1 : 0)) * C * 2 * rsz + (int64_t)t.npattern * 4;
    if (per_atom <= kSmemLimit - 1024 || nvt <= 1 || pl->order2) break;
    nvt = nvt == 3 ? 1 : 0;
  }
  if (per_atom > kSmemLimit - 1024)
    return fail(EPGX_ERR_CAPACITY, "state of one atom (" + std::to_string(per_atom) +
                                       " bytes) exceeds the shared memory of an SM; use max_nstate or float32");
  int A = atoms > 0 ? atoms : (128 / G > 0 ? 128 / G : 1);
  int64_t budget = atoms > 0 ? kSmemLimit - 1024 : 110 * 1024;
  if (per_atom > budget) budget = kSmemLimit - 1024;
  while (A > 1 && A * per_atom > budget) --A;
  while (A * G > 256) --A;
  if (A < 1) A = 1;
  if (atoms <= 0) { // small grids: smaller CTAs, so that every SM gets a few of them (latency-bound record loop)
    const int tiles = tiled ? (int)(pl->tiles.size() / 3) : nvt ? (t.nvar + nvt - 1) / nvt : 1;
    while (A > 1 && A * G > 32 && ((pl->natoms + A - 1) / A) * tiles < 4 * 148) A = (A + 1) / 2;
  }
  c.kernel = 0;
  c.lanes_per_atom = G;
  c.slots_per_lane = 0;
  c.vars_per_pass = nvt;
  c.var_tiles = tiled ? (int)(pl->tiles.size() / 3) : nvt ? (t.nvar + nvt - 1) / nvt : 1;
  c.atoms_per_cta = A;
  c.threads_per_cta = A * G;
  c.smem_bytes = (int)(A * per_atom + 16);
  c.ring = C;
  return EPGX_OK;
}

extern "C" int epgx_version(void) { return EPGX_VERSION; }

extern "C" const char *epgx_last_error(void) { return g_err.c_str(); }

extern "C" int epgx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

static int entry_reals(int code, int blk, int flags, const epgx_tape &t) {
  switch (code) {
  case EPGX_OP_T_GEN: return blk == 0 ? 6 : 0;
  case EPGX_OP_T_RE:
  case EPGX_OP_T_IM: return blk == 0 ? 4 : 0;
  case EPGX_OP_E: return blk == 0 ? 2 : blk == 1 ? 1 : ((flags & EPGX_FLAG_G) ? 2 : 0);
  case EPGX_OP_DIAG: return blk == 0 ? 8 : 0;
  case EPGX_OP_MATRIX: return blk == 0 ? 18 : (blk == 1 && (flags & EPGX_FLAG_AFFINE)) ? 6 : 0;
  case EPGX_OP_D: return blk == 0 ? 3 * (t.max_order + 1) : 0;
  case EPGX_OP_X: return blk == 0 ? 4 * t.npool * t.npool : 0;
  case EPGX_OP_PD: return blk == 0 ? 1 : 0;
  case EPGX_OP_FUSED: return blk == 0 ? ((flags & EPGX_FLAG_GEN) ? 6 : 4) : blk == 1 ? 2 : 1;
  case EPGX_OP_CONT: return blk == 0 ? 2 : blk == 1 ? 1 : 0;
  case EPGX_OP_ADC: return (blk == 0 && (flags & EPGX_FLAG_SCALE)) ? 2 : 0;
  default: return 0;
  }
}

extern "C" int epgx_plan_create(const epgx_tape *t, epgx_plan **out) {
  if (!t || !out) return fail(EPGX_ERR_INVALID, "null argument");
  *out = nullptr;
  if (t->dtype != EPGX_F64 && t->dtype != EPGX_F32) return fail(EPGX_ERR_INVALID, "bad dtype");
  if (t->ndim < 1 || t->ndim > EPGX_MAX_DIMS) return fail(EPGX_ERR_INVALID, "ndim out of range");
  if (t->npool < 1 || t->npool > EPGX_MAX_POOLS)
    return fail(EPGX_ERR_UNSUPPORTED, "number of exchange pools must be 1.." + std::to_string(EPGX_MAX_POOLS));
  if (t->npattern < 1 || t->npattern > EPGX_MAX_PATTERNS) return fail(EPGX_ERR_INVALID, "npattern out of range");
  if (t->nop < 0 || t->nseg < 0 || t->ncoef < 1 || !t->coef || (t->nop && !t->ops) || (t->nseg && !t->segs))
    return fail(EPGX_ERR_INVALID, "empty or null tape arrays");
  if (t->ncoef >= (int64_t)1 << 31) return fail(EPGX_ERR_CAPACITY, "coefficient table too large (>= 2^31 reals)");
  if (t->max_order < 0 || t->init_n < 0 || t->init_n > t->max_order || t->nvar < 0 || t->nadc < 0 || t->nvar1 < 0 ||
      t->nvar1 > t->nvar)
    return fail(EPGX_ERR_INVALID, "bad order / variable counts");
  if (t->ntile < 0 || (t->ntile > 0 && !t->tiles)) return fail(EPGX_ERR_INVALID, "bad variable tiles");
  for (int64_t i = 0; i < 3 * (int64_t)t->ntile; ++i)
    if (t->tiles[i] < -1 || t->tiles[i] >= t->nvar) return fail(EPGX_ERR_INVALID, "variable tile refers to an unknown variable");
  int64_t natoms = 1;
  for (int d = 0; d < t->ndim; ++d) {
    if (t->shape[d] < 1) return fail(EPGX_ERR_INVALID, "bad grid shape");
    natoms *= t->shape[d];
    if (natoms >= (int64_t)1 << 31) return fail(EPGX_ERR_CAPACITY, "more than 2^31 atoms");
  }
  // largest entry offset reachable through each pattern
  std::vector<int64_t> pat_span(t->npattern, 0);
  for (int q = 0; q < t->npattern; ++q) {
    int64_t span = 0;
    for (int d = 0; d < t->ndim; ++d) {
      if (t->stride[q][d] < 0) return fail(EPGX_ERR_INVALID, "negative pattern stride");
      span += (int64_t)t->stride[q][d] * (t->shape[d] - 1);
    }
    if (t->pool_stride[q] < 0) return fail(EPGX_ERR_INVALID, "negative pool stride");
    span += (int64_t)t->pool_stride[q] * (t->npool - 1);
    pat_span[q] = span;
  }
  auto block_ok = [&](uint32_t off, int pat, int64_t reals) {
    return pat >= 0 && pat < t->npattern && (int64_t)off + pat_span[pat] + reals <= t->ncoef;
  };
  if (!block_ok(t->init_off, t->init_pat, 6 * (int64_t)(t->init_n + 1)) || !block_ok(t->m0_off, t->m0_pat, 1))
    return fail(EPGX_ERR_INVALID, "initial-state block out of range");
  for (int64_t i = 0; i < t->nop; ++i) {
    const epgx_op &o = t->ops[i];
    if (o.code >= EPGX_OP_COUNT) return fail(EPGX_ERR_INVALID, "unknown opcode in record " + std::to_string(i));
    for (int b = 0; b < 3; ++b) {
      const int n = entry_reals(o.code, b, o.flags, *t);
      if (n && !block_ok(o.off[b], o.pat[b], n))
        return fail(EPGX_ERR_INVALID, "coefficient block out of range in record " + std::to_string(i));
    }
    if (o.code == EPGX_OP_FUSED && (i + 1 >= t->nop || t->ops[i + 1].code != EPGX_OP_CONT))
      return fail(EPGX_ERR_INVALID, "FUSED record without its CONT record at " + std::to_string(i));
    if ((o.flags & EPGX_FLAG_INJECT) && (o.aux < 0 || o.aux >= t->nvar))
      return fail(EPGX_ERR_INVALID, "injection into unknown variable in record " + std::to_string(i));
    if ((o.flags & EPGX_FLAG_INJECT) && (o.aux1 < 0 || o.aux1 > t->nvar))
      return fail(EPGX_ERR_INVALID, "injection from unknown state set in record " + std::to_string(i));
    if (o.code == EPGX_OP_ADC) {
      if ((o.flags & EPGX_FLAG_BASE) && (o.aux < 0 || o.aux >= t->nadc))
        return fail(EPGX_ERR_INVALID, "ADC row out of range in record " + std::to_string(i));
      if ((o.flags & EPGX_FLAG_PARTIALS) && (o.aux1 < 0 || o.aux1 >= t->njac))
        return fail(EPGX_ERR_INVALID, "jacobian row out of range in record " + std::to_string(i));
    }
  }
  double flops = 0, flops_r = 0, updates = 0;
  for (int64_t i = 0; i < t->nseg; ++i) {
    const epgx_segment &s = t->segs[i];
    const bool lat = s.flags & EPGX_SEG_LATTICE;
    if (s.first < 0 || s.count < 0 || (int64_t)s.first + s.count > t->nop || s.nact < -1 || s.nact > t->max_order ||
        s.n_old < 0 || s.n_new < 0 || s.n_new > t->max_order || s.n_old > t->max_order)
      return fail(EPGX_ERR_INVALID, "bad segment " + std::to_string(i));
    if (!lat && (s.shift < -1 || s.shift > 1 || s.n_new < s.n_old || s.n_new > s.n_old + 1))
      return fail(EPGX_ERR_INVALID, "bad segment " + std::to_string(i));
    if (lat) {
      if ((s.flags >> 16) < 0 || (s.flags >> 16) > s.n_old || (s.shift != 0 && s.shift != 2))
        return fail(EPGX_ERR_INVALID, "bad lattice segment " + std::to_string(i));
      if (s.shift == 2) {
        const int64_t nn = s.n_new + 1;
        if (!t->maps || s.rsv < 0 || (int64_t)s.rsv + 3 * nn > t->nmap)
          return fail(EPGX_ERR_INVALID, "gather maps out of range in segment " + std::to_string(i));
        for (int64_t j = 0; j < 3 * nn; ++j)
          if (t->maps[s.rsv + j] < -1 || t->maps[s.rsv + j] > s.n_old)
            return fail(EPGX_ERR_INVALID, "gather map refers to an unknown slot in segment " + std::to_string(i));
      }
    }
    for (int r = s.first; r < s.first + s.count; ++r) {
      const epgx_op &o = t->ops[r];
      if (o.code == EPGX_OP_FUSED && r + 1 >= s.first + s.count)
        return fail(EPGX_ERR_INVALID, "FUSED record split from its CONT record in segment " + std::to_string(i));
      const double f = form_flops(o.code, o.flags, t->npool) * t->npool * (s.nact + 1.0);
      if (f == 0) continue;
      double sets = 0;
      if (o.flags & EPGX_FLAG_INJECT) sets = 1;
      else sets = ((o.flags & EPGX_FLAG_BASE) ? 1 : 0) + ((o.flags & EPGX_FLAG_PARTIALS) ? t->nvar : 0);
      flops += f * sets;
      flops_r += form_flops_real(o.code) * t->npool * (s.nact + 1.0) * sets;
      if (!(o.flags & EPGX_FLAG_INJECT) && (o.flags & EPGX_FLAG_BASE)) updates += s.nact + 1.0;
    }
    if (s.shift) updates += s.n_new + 1.0;
  }

  epgx_plan *pl = new (std::nothrow) epgx_plan();
  if (!pl) return fail(EPGX_ERR_INVALID, "out of host memory");
  pl->tape = *t;
  pl->ops.assign(t->ops, t->ops + t->nop);
  pl->segs.assign(t->segs, t->segs + t->nseg);
  for (const epgx_segment &sg : pl->segs) pl->bounded = pl->bounded || (sg.flags & EPGX_SEG_MASK_TOP);
  if (t->dtype == EPGX_F64) pl->coef64.assign(t->coef, t->coef + t->ncoef);
  else {
    pl->coef32.resize(t->ncoef);
    for (int64_t i = 0; i < t->ncoef; ++i) pl->coef32[i] = (float)t->coef[i];
  }
  if (t->ntile) pl->tiles.assign(t->tiles, t->tiles + 3 * (size_t)t->ntile);
  pl->tape.tiles = nullptr;
  if (t->nmap > 0 && t->maps) pl->maps.assign(t->maps, t->maps + t->nmap);
  pl->tape.maps = nullptr;
  for (const epgx_segment &sg : pl->segs) pl->lattice = pl->lattice || (sg.flags & EPGX_SEG_LATTICE);
  if (pl->tape.nvar1 == 0) pl->tape.nvar1 = t->nvar;
  pl->order2 = t->ntile > 0 || pl->tape.nvar1 < t->nvar;
  for (int64_t i = 0; i < t->nop; ++i)
    if (((t->ops[i].flags & EPGX_FLAG_INJECT) && t->ops[i].aux1 > 0) || (t->ops[i].flags & (EPGX_FLAG_P1 | EPGX_FLAG_P2)))
      pl->order2 = true;
  pl->pats.assign((size_t)t->npattern * (EPGX_MAX_DIMS + 1), 0);
  for (int q = 0; q < t->npattern; ++q) {
    for (int d = 0; d < t->ndim; ++d) pl->pats[q * (EPGX_MAX_DIMS + 1) + d] = t->stride[q][d];
    pl->pats[q * (EPGX_MAX_DIMS + 1) + EPGX_MAX_DIMS] = t->pool_stride[q];
  }
  {
    bool ok = t->nvar == 0 && t->npool == 1 && !pl->lattice;
    for (int64_t i = 0; ok && i < t->nop; ++i) {
      const epgx_op &o = t->ops[i];
      switch (o.code) {
      case EPGX_OP_NOP: case EPGX_OP_T_RE: case EPGX_OP_D: case EPGX_OP_SPOIL: case EPGX_OP_PD: case EPGX_OP_ADC:
      case EPGX_OP_CONT: break;
      case EPGX_OP_E: ok = !(o.flags & EPGX_FLAG_G); break;
      case EPGX_OP_FUSED: ok = !(o.flags & EPGX_FLAG_IM); break;
      default: ok = false;
      }
    }
    if (ok) { // imaginary parts of the initial state
      const int64_t end = (int64_t)t->init_off + pat_span[t->init_pat] + 6 * (int64_t)(t->init_n + 1);
      for (int64_t i = t->init_off; ok && i < end; ++i)
        if (((i - t->init_off) % 6) % 2 == 1 && t->coef[i] != 0.0) ok = false;
    }
    pl->real_ok = ok;
    // the same with order-1 partial states: injections must be real as well
    bool okj = t->nvar > 0 && t->npool == 1 && !pl->order2 && !pl->lattice;
    for (int64_t i = 0; okj && i < t->nop; ++i) {
      const epgx_op &o = t->ops[i];
      switch (o.code) {
      case EPGX_OP_NOP: case EPGX_OP_T_RE: case EPGX_OP_D: case EPGX_OP_SPOIL: case EPGX_OP_PD: case EPGX_OP_ADC: break;
      case EPGX_OP_E: okj = !(o.flags & EPGX_FLAG_G); break;
      case EPGX_OP_DIAG: {
        const int64_t end = (int64_t)o.off[0] + pat_span[o.pat[0]] + 8;
        for (int64_t j = o.off[0]; okj && j < end; ++j)
          if (((j - o.off[0]) % 8) % 2 == 1 && t->coef[j] != 0.0) okj = false;
      } break;
      default: okj = false;
      }
    }
    if (okj) {
      const int64_t end = (int64_t)t->init_off + pat_span[t->init_pat] + 6 * (int64_t)(t->init_n + 1);
      for (int64_t i = t->init_off; okj && i < end; ++i)
        if (((i - t->init_off) % 6) % 2 == 1 && t->coef[i] != 0.0) okj = false;
    }
    pl->realjac_ok = okj;
    if (okj) { // every variable injected exactly once?
      std::vector<int> ninj(t->nvar, 0);
      for (int64_t i = 0; i < t->nop; ++i)
        if (t->ops[i].flags & EPGX_FLAG_INJECT) ++ninj[t->ops[i].aux];
      bool once = true;
      for (int v = 0; v < t->nvar; ++v) once = once && ninj[v] == 1;
      pl->pulsejac_ok = once;
    }
  }
  {
    // merged stream: [SEG(open seg 0)] seg_0 seg_1 ... where every segment is emitted as
    //   its records + SEG(close it, open the next)                                   (general form)
    //   TR  = [FUSED(RE) as TR, CONT']                 when it is exactly [FUSED, ADC(plain)] in a real-valued plan
    //   TRC = [FUSED(any) as TRC, CONT', CONT2]        when it is [D?] [FUSED] [ADC(F0, optional scale)] otherwise
    // (unit shift +-1 or none, any segment flags; the closing information rides in CONT').  TR pairs start at
    // even positions and TRC triples at multiples of 3 inside a TAPE_CHUNK window (NOP padding); a FUSED record is
    // never the last of a window.  Windows made only of plain TR / TRC records are flagged for the fast paths.
    auto seg_rec = [](int shift, int n_old, int n_new, int flags, int next_nact) {
      epgx_op o;
      memset(&o, 0, sizeof(o));
      o.code = EPGX_OP_SEG;
      o.aux = next_nact;
      o.off[0] = (uint32_t)shift;
      o.off[1] = (uint32_t)n_old;
      o.off[2] = (uint32_t)n_new;
      o.aux1 = flags;
      return o;
    };
    std::vector<epgx_op> &st = pl->stream;
    epgx_op nop;
    memset(&nop, 0, sizeof(nop));
    const int CH = epgx::kTapeChunk;
    st.push_back(seg_rec(0, 0, 0, 0, t->nseg ? t->segs[0].nact : -1));
    // records of every segment (copied: derivative tapes are regrouped below)
    std::vector<std::vector<epgx_op>> segrecs(t->nseg);
    for (int64_t i = 0; i < t->nseg; ++i) segrecs[i].assign(t->ops + t->segs[i].first, t->ops + t->segs[i].first + t->segs[i].count);
    const bool trj_ok = pl->realjac_ok && t->nvar <= 3;
    const int kAll = EPGX_FLAG_BASE | EPGX_FLAG_PARTIALS;
    auto is_inj = [&](const epgx_op &o, int code) {
      if (o.code != code || !(o.flags & EPGX_FLAG_INJECT) || o.aux < 0 || o.aux >= 3) return false;
      if (code == EPGX_OP_DIAG) { // the fused form needs the same factor on F+ and F-
        const int64_t end = (int64_t)o.off[0] + pat_span[o.pat[0]] + 8;
        for (int64_t j = o.off[0]; j < end; j += 8)
          if (t->coef[j] != t->coef[j + 2]) return false;
      }
      return true;
    };
    auto is_e = [&](const epgx_op &o) { return o.code == EPGX_OP_E && (o.flags & (kAll | EPGX_FLAG_INJECT | EPGX_FLAG_G)) == kAll; };
    if (trj_ok) {
      // a trailing [INJECT(DIAG)* E] group acts identically on every order: it commutes with the unit shift and
      // becomes the head of the next segment (the derivative counterpart of the E_pre hop of fuse_records)
      for (int64_t i = 0; i + 1 < t->nseg; ++i) {
        const epgx_segment &sg = t->segs[i];
        std::vector<epgx_op> &rs = segrecs[i];
        if (sg.shift == 0 || (sg.flags & EPGX_SEG_RESET) || rs.empty() || !is_e(rs.back()) || segrecs[i + 1].empty()) continue;
        size_t b = rs.size() - 1;
        while (b > 0 && is_inj(rs[b - 1], EPGX_OP_DIAG)) --b;
        if (b == 0) continue; // keep at least one record (the ADC) in the segment
        segrecs[i + 1].insert(segrecs[i + 1].begin(), rs.begin() + b, rs.end());
        rs.erase(rs.begin() + b, rs.end());
      }
    }
    for (int64_t i = 0; i < t->nseg; ++i) {
      epgx_segment sg = t->segs[i];
      sg.count = (int)segrecs[i].size();
      const epgx_op *rs = segrecs[i].data();
      const int next_nact = i + 1 < t->nseg ? t->segs[i + 1].nact : -1;
      const bool small = sg.n_old < 65536 && sg.n_new < 65536;
      auto close_into = [&](epgx_op &c) { // closing information of the segment in a CONT' record
        c.flags = (uint16_t)((sg.shift + 1) | (sg.flags << 2));
        c.off[2] = ((uint32_t)sg.n_old << 16) | (uint32_t)sg.n_new;
        c.aux1 = next_nact;
      };
      if (pl->real_ok && small && sg.count == 3 && rs[0].code == EPGX_OP_FUSED && rs[2].code == EPGX_OP_ADC &&
          rs[2].flags == EPGX_FLAG_BASE) {
        if (st.size() & 1) st.push_back(nop);
        epgx_op f = rs[0], c = rs[1];
        f.code = EPGX_OP_TR;
        c.aux = rs[2].aux;
        close_into(c);
        st.push_back(f);
        st.push_back(c);
        continue;
      }
      const int nd = (sg.count >= 1 && rs[0].code == EPGX_OP_D && rs[0].flags == EPGX_FLAG_BASE) ? 1 : 0;
      if (!pl->real_ok && t->nvar == 0 && t->npool == 1 && small && sg.count == nd + 3 && rs[nd].code == EPGX_OP_FUSED &&
          rs[nd + 2].code == EPGX_OP_ADC && (rs[nd + 2].flags & ~EPGX_FLAG_SCALE) == EPGX_FLAG_BASE) {
        while ((st.size() % CH) % 3 != 0 || (int)(st.size() % CH) == CH - 1) st.push_back(nop);
        epgx_op f = rs[nd], c = rs[nd + 1], c2 = nop;
        f.code = EPGX_OP_TRC;
        c.aux = rs[nd + 2].aux;
        close_into(c);
        c2.code = EPGX_OP_CONT;
        if (rs[nd + 2].flags & EPGX_FLAG_SCALE) { c2.flags |= 1; c2.off[0] = rs[nd + 2].off[0]; c2.pat[0] = rs[nd + 2].pat[0]; }
        if (nd) { c2.flags |= 2; c2.off[1] = rs[0].off[0]; c2.pat[1] = rs[0].pat[0]; }
        st.push_back(f);
        st.push_back(c);
        st.push_back(c2);
        continue;
      }
      if (trj_ok && small) {
        // [INJ(DIAG)*] [E]? [INJ(T_RE)*] T_RE [INJ(DIAG)*] [E]? ADC  ->  one TRJ group of five records
        int pg[3] = {-1, -1, -1}, tg[3] = {-1, -1, -1}, qg[3] = {-1, -1, -1}, epre = -1, epost = -1, tt = -1, adc = -1, q = 0;
        bool ok = true;
        for (; q < sg.count && is_inj(rs[q], EPGX_OP_DIAG); ++q) { ok = ok && pg[rs[q].aux] < 0; pg[rs[q].aux] = q; }
        if (q < sg.count && is_e(rs[q])) epre = q++;
        else ok = ok && pg[0] < 0 && pg[1] < 0 && pg[2] < 0;
        for (; q < sg.count && is_inj(rs[q], EPGX_OP_T_RE); ++q) { ok = ok && tg[rs[q].aux] < 0; tg[rs[q].aux] = q; }
        if (q < sg.count && rs[q].code == EPGX_OP_T_RE && (rs[q].flags & (kAll | EPGX_FLAG_INJECT)) == kAll) tt = q++;
        else ok = false;
        for (; ok && q < sg.count && is_inj(rs[q], EPGX_OP_DIAG); ++q) { ok = ok && qg[rs[q].aux] < 0; qg[rs[q].aux] = q; }
        if (ok && q < sg.count && is_e(rs[q])) epost = q++;
        else ok = ok && qg[0] < 0 && qg[1] < 0 && qg[2] < 0;
        // read-out of F0: one record for the signal row (BASE) and / or one for the Jacobian row (PARTIALS)
        int jrow = -1;
        while (ok && q < sg.count && rs[q].code == EPGX_OP_ADC && (rs[q].flags & kAll) && !(rs[q].flags & ~kAll)) {
          if (rs[q].flags & EPGX_FLAG_BASE) { ok = adc < 0; adc = q; }
          if (rs[q].flags & EPGX_FLAG_PARTIALS) { ok = ok && jrow < 0; jrow = rs[q].aux1; }
          ++q;
        }
        if (ok && adc >= 0 && q == sg.count) {
          while ((st.size() % CH) % 5 != 0 || (int)(st.size() % CH) > CH - 5) st.push_back(nop);
          epgx_op g[5];
          for (auto &o : g) o = nop;
          g[0].code = EPGX_OP_TRJ;
          g[0].flags = (uint16_t)((epre >= 0 ? EPGX_FLAG_PRE : 0) | (epost >= 0 ? EPGX_FLAG_POST : 0) |
                                  (jrow >= 0 ? EPGX_FLAG_PARTIALS : 0));
          g[0].off[0] = rs[tt].off[0]; g[0].pat[0] = rs[tt].pat[0];
          if (epre >= 0) { g[0].off[1] = rs[epre].off[0]; g[0].pat[1] = rs[epre].pat[0]; g[0].off[2] = rs[epre].off[1]; g[0].pat[2] = rs[epre].pat[1]; }
          for (int k = 1; k < 5; ++k) g[k].code = EPGX_OP_CONT;
          if (epost >= 0) { g[1].off[0] = rs[epost].off[0]; g[1].pat[0] = rs[epost].pat[0]; g[1].off[1] = rs[epost].off[1]; g[1].pat[1] = rs[epost].pat[1]; }
          g[1].aux = rs[adc].aux;
          close_into(g[1]);
          g[1].rsv1 = jrow; // Jacobian row
          for (int v = 0; v < 3; ++v) {
            if (pg[v] >= 0) { g[2].flags |= 1 << v; g[2].off[v] = rs[pg[v]].off[0]; g[2].pat[v] = rs[pg[v]].pat[0]; }
            if (tg[v] >= 0) { g[3].flags |= 1 << v; g[3].off[v] = rs[tg[v]].off[0]; g[3].pat[v] = rs[tg[v]].pat[0]; }
            if (qg[v] >= 0) { g[4].flags |= 1 << v; g[4].off[v] = rs[qg[v]].off[0]; g[4].pat[v] = rs[qg[v]].pat[0]; }
          }
          for (auto &o : g) st.push_back(o);
          continue;
        }
      }
      for (int r = 0; r < sg.count; ++r) {
        if (rs[r].code == EPGX_OP_FUSED && (int)(st.size() % CH) == CH - 1) st.push_back(nop);
        st.push_back(rs[r]);
      }
      st.push_back(seg_rec(sg.shift, sg.n_old, sg.n_new, sg.flags, next_nact));
    }
    // plain whole-TR pair: unit shift +1, no segment flag but (possibly) the truncation at max_nstate
    auto plain_tr = [&](const std::vector<epgx_op> &v, size_t k) {
      return k + 1 < v.size() && v[k].code == EPGX_OP_TR && (v[k + 1].flags & ~(EPGX_SEG_MASK_TOP << 2)) == 2;
    };
    if (pl->real_ok) {
      // Runs of plain whole-TR pairs start at a window boundary: the records in front of a run are padded with NOPs
      // (the first one, aux = 1, tells the record loop to skip to the next window), the run fills whole windows and,
      // unless the stream ends there, its last window is padded as well.  Windows hold an EVEN number of pairs (the
      // fast path runs two TRs per iteration, epgx_real.cuh), so the first pair of an odd run stays in front of the
      // padding as a generic record.  Without this the first and the last TRs of a FISP train -- 4 % of the TRs, where
      // few orders are populated -- went through the record-at-a-time path and took 20 % of the kernel's time.
      std::vector<epgx_op> in;
      in.swap(st);
      const size_t n = in.size();
      auto pad_window = [&]() {
        if (st.size() % CH == 0) return;
        epgx_op skip = nop;
        skip.aux = 1;
        st.push_back(skip);
        while (st.size() % CH) st.push_back(nop);
      };
      size_t i = 0;
      while (i < n) {
        if (in[i].code == EPGX_OP_NOP) { ++i; continue; } // alignment padding of the first pass
        std::vector<size_t> run;
        size_t j = i;
        for (;;) {
          while (j < n && in[j].code == EPGX_OP_NOP) ++j;
          if (!plain_tr(in, j)) break;
          run.push_back(j);
          j += 2;
        }
        if (run.size() >= 4) {
          size_t k = 0;
          if (run.size() & 1) {
            if (st.size() & 1) st.push_back(nop);
            st.push_back(in[run[0]]);
            st.push_back(in[run[0] + 1]);
            k = 1;
          }
          pad_window();
          for (; k < run.size(); ++k) { st.push_back(in[run[k]]); st.push_back(in[run[k] + 1]); }
          i = j;
          while (i < n && in[i].code == EPGX_OP_NOP) ++i;
          if (i < n) pad_window();
          continue;
        }
        const int code = in[i].code;
        if (code == EPGX_OP_TR) { // (not plain, or a short run) a pair at an even position: never split by a window
          if (st.size() & 1) st.push_back(nop);
          st.push_back(in[i]);
          st.push_back(in[i + 1]);
          i += 2;
          continue;
        }
        if (code == EPGX_OP_FUSED && (int)(st.size() % CH) == CH - 1) st.push_back(nop);
        st.push_back(in[i++]);
      }
    }
    for (size_t b0 = 0; b0 < st.size(); b0 += CH) {
      // windows made of an even number of plain TR pairs followed by nothing but padding: the fast path of the real
      // kernel; the pair count rides in rsv1 of the window's first record
      int m = 0;
      while (m < CH / 2 && plain_tr(st, b0 + 2 * m)) ++m;
      bool pure = m >= 2 && m % 2 == 0;
      for (size_t k = b0 + 2 * m; pure && k < b0 + CH && k < st.size(); ++k) pure = st[k].code == EPGX_OP_NOP;
      if (pure) { st[b0].flags |= 0x8000; st[b0].rsv1 = m; }
    }
    for (size_t b0 = 0; b0 + CH <= st.size(); b0 += CH) {
      bool purec = st[b0 + CH - 1].code == EPGX_OP_NOP;
      for (int j = 0; purec && j + 3 <= CH - 1; j += 3) purec = st[b0 + j].code == EPGX_OP_TRC && st[b0 + j + 1].flags == 2;
      if (purec) st[b0].flags |= 0x4000;
    }
    for (size_t b0 = 0; b0 < st.size(); b0 += CH) { // windows holding TRJ groups: coefficient assembly pass
      int nj = 0, nplain = 0;
      for (size_t j = b0; j + 1 < st.size() && j < b0 + CH; j += 5)
        if (st[j].code == EPGX_OP_TRJ) { ++nj; if (st[j + 1].flags == 2) ++nplain; }
      pl->ntrj += nj;
      if (nj) st[b0].flags |= 0x2000;
      bool pure = nplain == epgx::kTrjPerWindow && b0 + CH <= st.size();
      for (int j = 5 * epgx::kTrjPerWindow; pure && j < CH; ++j) pure = st[b0 + j].code == EPGX_OP_NOP;
      if (pure) st[b0].flags |= 0x1000;
    }
  }
  pl->tape.ops = pl->ops.data();
  pl->tape.segs = pl->segs.data();
  pl->tape.coef = nullptr;
  pl->natoms = natoms;
  memset(&pl->cfg, 0, sizeof(pl->cfg));
  int rc = choose_variant(pl, pl->cfg, 0, 0, 0, 0);
  if (rc != EPGX_OK) {
    delete pl;
    return rc;
  }
  pl->flops_cplx = flops;
  pl->flops_real = flops_r;
  pl->updates = updates;
  pl->cfg.updates_per_atom = updates;
  pl->cfg.flops_per_atom = pl->cfg.kernel >= 2 ? pl->flops_real : pl->flops_cplx;
  auto align = [](int64_t x) { return (x + 255) & ~(int64_t)255; };
  const int rsz = t->dtype == EPGX_F64 ? 8 : 4;
  pl->off_ops = 0;
  pl->off_segs = align(pl->off_ops + (int64_t)sizeof(epgx_op) * (t->nop ? t->nop : 1));
  pl->off_pats = align(pl->off_segs + (int64_t)sizeof(epgx_segment) * (t->nseg ? t->nseg : 1));
  pl->off_stream = align(pl->off_pats + (int64_t)pl->pats.size() * 4);
  pl->off_tiles = align(pl->off_stream + (int64_t)sizeof(epgx_op) * pl->stream.size());
  pl->off_maps = align(pl->off_tiles + (int64_t)pl->tiles.size() * 4);
  pl->off_coef = align(pl->off_maps + (int64_t)pl->maps.size() * 4);
  pl->ws_bytes = align(pl->off_coef + t->ncoef * rsz);
  *out = pl;
  return EPGX_OK;
}

extern "C" int epgx_plan_destroy(epgx_plan *pl) {
  delete pl;
  return EPGX_OK;
}

extern "C" int epgx_plan_config(const epgx_plan *pl, epgx_config *cfg) {
  if (!pl || !cfg) return fail(EPGX_ERR_INVALID, "null argument");
  *cfg = pl->cfg;
  return EPGX_OK;
}

extern "C" int epgx_plan_set_variant(epgx_plan *pl, int kernel, int lanes, int vars, int atoms) {
  if (!pl) return fail(EPGX_ERR_INVALID, "null plan");
  int rc = choose_variant(pl, pl->cfg, kernel, lanes, vars, atoms);
  pl->cfg.flops_per_atom = pl->cfg.kernel >= 2 ? pl->flops_real : pl->flops_cplx;
  pl->cfg.updates_per_atom = pl->updates;
  return rc;
}

extern "C" int epgx_plan_stream(const epgx_plan *pl, const epgx_op **records, int64_t *count) {
  if (!pl || !records || !count) return fail(EPGX_ERR_INVALID, "null argument");
  *records = pl->stream.data();
  *count = (int64_t)pl->stream.size();
  return EPGX_OK;
}

extern "C" int epgx_plan_workspace_bytes(const epgx_plan *pl, int64_t *bytes) {
  if (!pl || !bytes) return fail(EPGX_ERR_INVALID, "null argument");
  *bytes = pl->ws_bytes;
  return EPGX_OK;
}

extern "C" int epgx_plan_upload(const epgx_plan *pl, void *ws, void *stream) {
  if (!pl || !ws) return fail(EPGX_ERR_INVALID, "null argument");
  if ((uintptr_t)ws & 255) return fail(EPGX_ERR_INVALID, "workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  char *w = (char *)ws;
  const epgx_tape &t = pl->tape;
  if (t.nop) CUDA_TRY(cudaMemcpyAsync(w + pl->off_ops, pl->ops.data(), sizeof(epgx_op) * t.nop, cudaMemcpyHostToDevice, st));
  if (t.nseg)
    CUDA_TRY(cudaMemcpyAsync(w + pl->off_segs, pl->segs.data(), sizeof(epgx_segment) * t.nseg, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(w + pl->off_pats, pl->pats.data(), pl->pats.size() * 4, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(w + pl->off_stream, pl->stream.data(), sizeof(epgx_op) * pl->stream.size(), cudaMemcpyHostToDevice, st));
  if (!pl->tiles.empty())
    CUDA_TRY(cudaMemcpyAsync(w + pl->off_tiles, pl->tiles.data(), pl->tiles.size() * 4, cudaMemcpyHostToDevice, st));
  if (!pl->maps.empty())
    CUDA_TRY(cudaMemcpyAsync(w + pl->off_maps, pl->maps.data(), pl->maps.size() * 4, cudaMemcpyHostToDevice, st));
  if (t.dtype == EPGX_F64)
    CUDA_TRY(cudaMemcpyAsync(w + pl->off_coef, pl->coef64.data(), t.ncoef * 8, cudaMemcpyHostToDevice, st));
  else
    CUDA_TRY(cudaMemcpyAsync(w + pl->off_coef, pl->coef32.data(), t.ncoef * 4, cudaMemcpyHostToDevice, st));
  return EPGX_OK;
}

static int dispatch(const epgx_plan *pl, const epgx_config &c, const KParams &kp, cudaStream_t st) {
  const bool f64 = pl->tape.dtype == EPGX_F64;
  dim3 grid((unsigned)((kp.atom_count + c.atoms_per_cta - 1) / c.atoms_per_cta), (unsigned)c.var_tiles);
  cudaError_t e;
  switch (c.kernel) {
  case 5: e = f64 ? launch_pulsejac<double>(c.slots_per_lane, kp, grid, c.threads_per_cta, c.smem_bytes, st)
                  : launch_pulsejac<float>(c.slots_per_lane, kp, grid, c.threads_per_cta, c.smem_bytes, st); break;
  case 4: e = f64 ? launch_setjac<double>(c.slots_per_lane, kp, grid, c.threads_per_cta, c.smem_bytes, st)
                  : launch_setjac<float>(c.slots_per_lane, kp, grid, c.threads_per_cta, c.smem_bytes, st); break;
  case 3: e = f64 ? launch_realjac<double>(c.slots_per_lane, kp, grid, c.threads_per_cta, c.smem_bytes, st)
                  : launch_realjac<float>(c.slots_per_lane, kp, grid, c.threads_per_cta, c.smem_bytes, st); break;
  case 2: e = f64 ? launch_real<double>(c.slots_per_lane, kp, grid, c.threads_per_cta, c.smem_bytes, st)
                  : launch_real<float>(c.slots_per_lane, kp, grid, c.threads_per_cta, c.smem_bytes, st); break;
  case 1: e = f64 ? launch_reg<double>(c.slots_per_lane, kp, grid, c.threads_per_cta, c.smem_bytes, st)
                  : launch_reg<float>(c.slots_per_lane, kp, grid, c.threads_per_cta, c.smem_bytes, st); break;
  default: e = f64 ? launch_ring<double>(pl->tape.npool, c.vars_per_pass, kp, grid, c.threads_per_cta, c.smem_bytes, st)
                   : launch_ring<float>(pl->tape.npool, c.vars_per_pass, kp, grid, c.threads_per_cta, c.smem_bytes, st);
  }
  if (e == cudaErrorInvalidValue)
    return fail(EPGX_ERR_UNSUPPORTED, "no kernel instance for kernel=" + std::to_string(c.kernel) + " slots=" +
                                          std::to_string(c.slots_per_lane) + " npool=" + std::to_string(pl->tape.npool) +
                                          " vars_per_pass=" + std::to_string(c.vars_per_pass));
  if (e != cudaSuccess) return fail(EPGX_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
  return EPGX_OK;
}

extern "C" int epgx_simulate(const epgx_plan *pl, const void *ws, int64_t atom_begin, int64_t atom_count, void *signal,
                             void *jacobian, void *stream) {
  return epgx_simulate_strided(pl, ws, atom_begin, atom_count, signal, atom_count, jacobian, atom_count, stream);
}

static bool real_signal_ok(const epgx_plan *pl) {
  if (!pl || pl->cfg.kernel != 2 || pl->tape.nvar != 0) return false;
  for (const epgx_op &o : pl->ops)
    if (o.code == EPGX_OP_ADC && (o.flags & EPGX_FLAG_SCALE)) return false; // a complex read-out factor
  return true;
}

static int run_range(const epgx_plan *pl, const void *ws, int64_t atom_begin, int64_t atom_count, void *signal,
                     int64_t signal_stride, void *jacobian, int64_t jacobian_stride, void *state, void *stream,
                     bool out_real = false) {
  if (!pl || !ws) return fail(EPGX_ERR_INVALID, "null argument");
  if (signal_stride < atom_count || jacobian_stride < atom_count) return fail(EPGX_ERR_INVALID, "row stride < atom_count");
  if (atom_begin < 0 || atom_count < 0 || atom_begin + atom_count > pl->natoms)
    return fail(EPGX_ERR_INVALID, "atom range out of the grid");
  if (atom_count == 0) return EPGX_OK;
  const epgx_tape &t = pl->tape;
  if (t.nadc && !signal) return fail(EPGX_ERR_INVALID, "null signal buffer");
  if (t.nvar && t.njac && !jacobian) return fail(EPGX_ERR_INVALID, "null jacobian buffer");
  epgx_config cfg = pl->cfg;
  if (state && cfg.kernel != 0) { // the state leaves the chip through the shared-memory kernel only
    const int rc = choose_variant(pl, cfg, 1, 0, 0, 0);
    if (rc != EPGX_OK) return rc;
  }
  const char *w = (const char *)ws;
  KParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.ops = (const epgx_op *)(w + pl->off_ops);
  kp.segs = (const epgx_segment *)(w + pl->off_segs);
  kp.pats = (const int *)(w + pl->off_pats);
  kp.stream = w + pl->off_stream;
  kp.nstream = (int)pl->stream.size();
  kp.bounded = pl->bounded ? 1 : 0;
  kp.coef = w + pl->off_coef;
  kp.signal = signal;
  kp.jac = jacobian;
  kp.state = state;
  kp.out_real = out_real ? 1 : 0;
  kp.atom_begin = atom_begin;
  kp.atom_count = atom_count;
  kp.sig_stride = signal_stride;
  kp.jac_stride = jacobian_stride;
  for (int d = 0; d < t.ndim; ++d) kp.shape[d] = (int)t.shape[d];
  kp.ndim = t.ndim;
  kp.npattern = t.npattern;
  kp.nseg = (int)t.nseg;
  kp.G = cfg.lanes_per_atom;
  kp.A = cfg.atoms_per_cta;
  kp.C = cfg.ring;
  kp.nvar = t.nvar;
  kp.nvar1 = t.nvar1;
  kp.tiles = pl->tiles.empty() ? nullptr : (const int *)(w + pl->off_tiles);
  kp.maps = pl->maps.empty() ? nullptr : (const int *)(w + pl->off_maps);
  kp.lattice = pl->lattice ? 1 : 0;
  kp.init_off = t.init_off;
  kp.m0_off = t.m0_off;
  kp.init_pat = t.init_pat;
  kp.m0_pat = t.m0_pat;
  kp.init_n = t.init_n;
  return dispatch(pl, cfg, kp, (cudaStream_t)stream);
}

extern "C" int epgx_simulate_strided(const epgx_plan *pl, const void *ws, int64_t atom_begin, int64_t atom_count,
                                     void *signal, int64_t signal_stride, void *jacobian, int64_t jacobian_stride,
                                     void *stream) {
  return run_range(pl, ws, atom_begin, atom_count, signal, signal_stride, jacobian, jacobian_stride, nullptr, stream);
}

extern "C" int epgx_plan_real_signal(const epgx_plan *pl) { return real_signal_ok(pl) ? 1 : 0; }

extern "C" int epgx_simulate_real(const epgx_plan *pl, const void *ws, int64_t atom_begin, int64_t atom_count, void *signal,
                                  int64_t signal_stride, void *stream) {
  if (!real_signal_ok(pl)) return fail(EPGX_ERR_UNSUPPORTED, "the signal of this plan is not real-valued (see epgx_plan_real_signal)");
  return run_range(pl, ws, atom_begin, atom_count, signal, signal_stride, nullptr, atom_count, nullptr, stream, true);
}

extern "C" int epgx_simulate_state(const epgx_plan *pl, const void *ws, int64_t atom_begin, int64_t atom_count, void *signal,
                                   int64_t signal_stride, void *jacobian, int64_t jacobian_stride, void *state, void *stream) {
  if (!state) return fail(EPGX_ERR_INVALID, "null state buffer");
  if (pl && pl->lattice) return fail(EPGX_ERR_UNSUPPORTED, "state read-back of a lattice tape (the half-storage format is 1-d)");
  return run_range(pl, ws, atom_begin, atom_count, signal, signal_stride, jacobian, jacobian_stride, state, stream);
}

extern "C" int epgx_simulate_host(const epgx_plan *pl, int device, int64_t atom_begin, int64_t atom_count, void *signal,
                                  void *jacobian) {
  if (!pl) return fail(EPGX_ERR_INVALID, "null plan");
  int ndev = epgx_device_count();
  if (ndev == 0) return fail(EPGX_ERR_NO_DEVICE, "no CUDA device: the epgx engine has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(EPGX_ERR_INVALID, "bad device index");
  CUDA_TRY(cudaSetDevice(device));
  const epgx_tape &t = pl->tape;
  const int64_t csz = t.dtype == EPGX_F64 ? 16 : 8;
  const int64_t sig_bytes = csz * t.nadc * atom_count * t.npool;
  const int64_t jac_bytes = csz * t.njac * t.nvar * atom_count * t.npool;
  void *ws = nullptr, *dsig = nullptr, *djac = nullptr;
  int rc = EPGX_OK;
  cudaError_t e = cudaMalloc(&ws, pl->ws_bytes);
  if (e == cudaSuccess && sig_bytes) e = cudaMalloc(&dsig, sig_bytes);
  if (e == cudaSuccess && jac_bytes) e = cudaMalloc(&djac, jac_bytes);
  if (e != cudaSuccess) rc = fail(EPGX_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  if (rc == EPGX_OK) rc = epgx_plan_upload(pl, ws, nullptr);
  if (rc == EPGX_OK) rc = epgx_simulate(pl, ws, atom_begin, atom_count, dsig, djac, nullptr);
  if (rc == EPGX_OK && sig_bytes) {
    e = cudaMemcpy(signal, dsig, sig_bytes, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(EPGX_ERR_CUDA, std::string("D2H signal: ") + cudaGetErrorString(e));
  }
  if (rc == EPGX_OK && jac_bytes) {
    e = cudaMemcpy(jacobian, djac, jac_bytes, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(EPGX_ERR_CUDA, std::string("D2H jacobian: ") + cudaGetErrorString(e));
  }
  if (rc == EPGX_OK) {
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) rc = fail(EPGX_ERR_CUDA, std::string("sync: ") + cudaGetErrorString(e));
  }
  cudaFree(ws);
  cudaFree(dsig);
  cudaFree(djac);
  return rc;
}

extern "C" int epgx_peer_alloc(int64_t bytes, void **ptr, char handle[64]) {
  if (!ptr || !handle || bytes <= 0) return fail(EPGX_ERR_INVALID, "bad peer_alloc arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  *ptr = nullptr;
  CUDA_TRY(cudaMalloc(ptr, (size_t)bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, *ptr);
  if (e != cudaSuccess) {
    cudaFree(*ptr);
    *ptr = nullptr;
    return fail(EPGX_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  }
  memcpy(handle, &h, 64);
  return EPGX_OK;
}

extern "C" int epgx_peer_open(const char handle[64], void **ptr) {
  if (!ptr || !handle) return fail(EPGX_ERR_INVALID, "null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return EPGX_OK;
}

extern "C" int epgx_peer_close(void *ptr) {
  if (ptr) CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return EPGX_OK;
}

extern "C" int epgx_peer_free(void *ptr) {
  if (ptr) CUDA_TRY(cudaFree(ptr));
  return EPGX_OK;
}

extern "C" int epgx_copy2d_device(void *dst, int64_t dst_pitch, const void *src, int64_t src_pitch, int64_t width,
                                  int64_t height, void *stream) {
  if (!dst || !src) return fail(EPGX_ERR_INVALID, "null argument");
  CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)width, (size_t)height,
                             cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return EPGX_OK;
}

extern "C" int epgx_copy2d_to_host(void *dst, int64_t dst_pitch, const void *src, int64_t src_pitch, int64_t width,
                                   int64_t height, void *stream) {
  if (!dst || !src) return fail(EPGX_ERR_INVALID, "null argument");
  CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)width, (size_t)height,
                             cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return EPGX_OK;
}

// ---- Adc(reduce=axis): out[o][i] = sum_r in[o][r][i]   (probe.py:148-153)
template <typename real2> __global__ void reduce_kernel(const real2 *in, real2 *out, long long nouter, long long nred, long long ninner) {
  const long long total = nouter * ninner;
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < total; j += (long long)gridDim.x * blockDim.x) {
    const long long o = j / ninner, i = j - o * ninner;
    const real2 *src = in + o * nred * ninner + i;
    real2 acc = src[0];
    for (long long r = 1; r < nred; ++r) {
      const real2 v = src[r * ninner];
      acc.x += v.x;
      acc.y += v.y;
    }
    out[j] = acc;
  }
}

extern "C" int epgx_reduce(int dtype, const void *in, void *out, int64_t nouter, int64_t nred, int64_t ninner, void *stream) {
  if (!in || !out || nouter < 0 || nred < 1 || ninner < 0) return fail(EPGX_ERR_INVALID, "bad reduce arguments");
  const int64_t total = nouter * ninner;
  if (total == 0) return EPGX_OK;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == EPGX_F64) reduce_kernel<double2><<<blocks, 256, 0, st>>>((const double2 *)in, (double2 *)out, nouter, nred, ninner);
  else if (dtype == EPGX_F32) reduce_kernel<float2><<<blocks, 256, 0, st>>>((const float2 *)in, (float2 *)out, nouter, nred, ninner);
  else return fail(EPGX_ERR_INVALID, "bad dtype");
  CUDA_TRY(cudaGetLastError());
  return EPGX_OK;
}

// ---- FMA roofline denominator: 8 independent dependent-FMA chains per thread
template <typename real> __global__ void fma_kernel(real *out, int iters, real a, real b) {
  real x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = real(threadIdx.x + i) * real(1e-3);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = x[i] * a + b;
  }
  real s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == real(123.456)) out[0] = s;
}

extern "C" int epgx_fma_peak(int device, int dtype, double seconds, double *tflops) {
  if (!tflops) return fail(EPGX_ERR_INVALID, "null argument");
  int ndev = epgx_device_count();
  if (ndev == 0) return fail(EPGX_ERR_NO_DEVICE, "no CUDA device");
  if (device < 0 || device >= ndev) return fail(EPGX_ERR_INVALID, "bad device index");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  void *buf = nullptr;
  CUDA_TRY(cudaMalloc(&buf, 64));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int threads = 256, blocks = prop.multiProcessorCount * 8;
  int iters = 2000;
  double best = 0, spent = 0;
  if (seconds <= 0) seconds = 0.2;
  for (int rep = 0; rep < 200 && spent < seconds; ++rep) {
    cudaEventRecord(e0);
    if (dtype == EPGX_F64) fma_kernel<double><<<blocks, threads>>>((double *)buf, iters, 1.0000001, 1e-9);
    else fma_kernel<float><<<blocks, threads>>>((float *)buf, iters, 1.0000001f, 1e-9f);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) {
      cudaFree(buf);
      return fail(EPGX_ERR_CUDA, std::string("fma kernel: ") + cudaGetErrorString(e));
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 8 * 16 * (double)iters * threads * blocks;
    const double tf = fl / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
    spent += ms * 1e-3;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  *tflops = best;
  return EPGX_OK;
}
