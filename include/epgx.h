/*
 * epgx.h -- C ABI of libepgx.so: the B200 (sm_100a) execution engine behind the epgpy
 * operator API.
 *
 * What this replaces in the reference (py-baudin/epgpy, all Python, no FFI of its own):
 *   - the array-module seam         epgpy/common.py:21-74   (numpy|cupy dispatch)
 *   - the per-operator loop         epgpy/functions.py:173-192 (simulate_simple)
 *   - the per-operator "kernels"    epgpy/opmatrix.py:199-221 (T, MatrixOp), opscalar.py:213-232
 *                                   (E, P, R, ScalarOp), shift.py:271-294 (S, shift-1d),
 *                                   diffusion.py:60-79 (D), exchange.py:89-120 (X),
 *                                   probe.py:63-66,138-165 (ADC), diff.py:264-288 (order-1 partials),
 *                                   operator.py:281-341 (SPOILER, RESET, PD)
 *   - the state container storage   epgpy/statematrix.py:388-422, 793-804
 *
 * Model.  The Python host lowers a flattened operator sequence into
 *   (1) a COEFFICIENT TABLE : reals, one block per operator parameter group, kept in the group's
 *                             own (un-broadcast) shape -- e.g. E(tau, T1[100,1,1], T2[1,100,1])
 *                             stores 100 (e1, 1-e1) pairs and 100 e2 values,
 *   (2) PATTERNS            : per-grid-axis strides in reals (0 = broadcast) that map an atom to its
 *                             entry of a block -- the left-aligned broadcasting of
 *                             epgpy/common.py:273-334 becomes integer strides,
 *   (3) an OP TAPE          : one 32-byte record per operator application, referring to up to three
 *                             coefficient blocks through (offset, pattern) pairs,
 *   (4) SEGMENTS            : maximal runs of records that act on each configuration order
 *                             independently (everything except S / RESET); a segment is ONE pass
 *                             over the on-chip state, followed by at most one unit shift.
 * One launch of the fused kernel runs the WHOLE tape for a slab of atoms.  The F+/F-/Z state of an
 * atom (half storage, orders k >= 0) stays on chip for the whole sequence; only ADC samples go to
 * HBM.  Order-1 partial states (diff.py) are further state sets of the same atom and are propagated
 * in the same pass.
 *
 * Conventions: every entry point returns 0 on success, a negative epgx_status otherwise, and
 * never throws.  epgx_last_error() returns a thread-local message.  The CALLER owns every
 * buffer; device pointers typically come from torch tensors (tensor.data_ptr()).  Launches are
 * asynchronous on the given CUDA stream (cudaStream_t passed as void*; NULL = default stream);
 * the library never synchronises the device except in epgx_simulate_host and epgx_fma_peak.
 * A plan is immutable after creation (except epgx_plan_set_variant) and may be used from several
 * host threads / devices concurrently.
 */
#ifndef EPGX_H
#define EPGX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EPGX_VERSION 109 /* 0.1.9: EPGX_FLAG_SLOT (read-outs of any lattice slot: accumulated-time F0, DFT / Imaging probes); 0.1.8: peer windows, real-valued signal rows, lattices, order 2 */
#define EPGX_MAX_DIMS 8
#define EPGX_MAX_PATTERNS 64
#define EPGX_MAX_POOLS 4

typedef enum {
  EPGX_OK = 0,
  EPGX_ERR_INVALID = -1,     /* malformed tape / arguments */
  EPGX_ERR_CAPACITY = -2,    /* the state of one atom does not fit the on-chip memory of an SM */
  EPGX_ERR_CUDA = -3,        /* CUDA runtime error (message in epgx_last_error) */
  EPGX_ERR_UNSUPPORTED = -4, /* valid request the engine does not implement */
  EPGX_ERR_NO_DEVICE = -5
} epgx_status;

typedef enum { EPGX_F64 = 0, EPGX_F32 = 1 } epgx_dtype;

/* ---- tape record codes.  "blk i: (...)" lists the reals of one entry of coefficient block i. */
typedef enum {
  EPGX_OP_NOP = 0,
  /* RF pulse and its derivatives in rotated-Rx form (transition.py:114-196).  With a, w real and
   * B, U complex:
   *   F+' = a F+ + B F- + U Z ; F-' = conj(B) F+ + a F- + conj(U) Z ; Z' = -1/2 (conj(U) F+ + U F-) + w Z
   * T = Rz(phi)Rx(alpha)Rz(-phi) has a = cos^2(alpha/2), B = sin^2(alpha/2) e^{2i phi},
   * U = -i sin(alpha) e^{i phi}, w = cos(alpha); dT/dalpha and dT/dphi have the same shape.
   *   T_GEN blk0: (a, w, B.re, B.im, U.re, U.im)
   *   T_RE  blk0: (a, w, b, u)  B = b, U = u   (phi = +-90 deg)
   *   T_IM  blk0: (a, w, b, u)  B = b, U = -i u (phi = 0 / 180 deg) */
  EPGX_OP_T_GEN = 1,
  EPGX_OP_T_RE = 2,
  EPGX_OP_T_IM = 3,
  /* relaxation (evolution.py:220-256): F+ *= e2 (c + i s), F- *= e2 (c - i s), Z *= e1,
   * Z(0) += r0 * M0 (EPGX_FLAG_AFFINE).
   *   blk0: (e1, r0)   blk1: (e2)   blk2: (c, s) = cis(2 pi g tau), only with EPGX_FLAG_G */
  EPGX_OP_E = 4,
  /* generic diagonal operator (opscalar.py:213-232):
   *   blk0: (aP.re, aP.im, aM.re, aM.im, aZ.re, aZ.im, a0Z.re, a0Z.im); Z(0) += a0Z * M0 (AFFINE) */
  EPGX_OP_DIAG = 5,
  /* generic 3x3 operator (opmatrix.py:199-221): blk0: 9 complex row-major (18 reals),
   *   blk1: mat0[:,2] (3 complex), only with EPGX_FLAG_AFFINE: state(0) += mat0[:,2] * M0 */
  EPGX_OP_MATRIX = 6,
  /* diffusion (diffusion.py:60-79): blk0: per-order rows (dT+[k], dT-[k], dL[k]), one entry =
   *   3*(max_order+1) reals: F+(k) *= dT+[k], F-(k) *= dT-[k], Z(k) *= dL[k] */
  EPGX_OP_D = 7,
  /* exchange (exchange.py:89-120): blk0: N*N complex mT[dst][src] then N*N complex mL[dst][src];
   *   s_c <- m_c (s_c - eq_c) + eq_c over the pool axis, m = (mT, conj mT, mL) */
  EPGX_OP_X = 8,
  EPGX_OP_SPOIL = 9, /* F+ = F- = 0 (operator.py:281-286) */
  EPGX_OP_PD = 10,   /* blk0: (M0): set the equilibrium density (operator.py:315-341) */
  /* read-out (probe.py:138-165): value = F+(0) or Z(0), times blk0: (re, im) if EPGX_FLAG_SCALE
   *   (ADC phasor x weights).  aux = row of `signal` (EPGX_FLAG_BASE), aux1 = row of `jacobian`
   *   (EPGX_FLAG_PARTIALS). */
  EPGX_OP_ADC = 11,
  /* fused E_pre -> T -> E_post (the host's peephole pass over [E, T_RE|T_IM, E] runs without precession;
   * an E that closes a segment commutes with the shift and becomes the E_pre of the next segment).
   * Two records: FUSED carries blk0: T (a, w, b, u) -- or the six reals of T_GEN with EPGX_FLAG_GEN --, blk1: pre
 * (e1, r0), blk2: pre (e2);
   * the CONT record that follows carries blk0: post (e1, r0), blk1: post (e2).  Per atom the kernel
   * assembles A = e2' e2 a, B = e2' e2 b, U = e2' e1 u, H = -+1/2 e1' e2 u, W = e1' e1 w and applies
   *   F+' = A F+ + B F- + U' Z ; F-' = B F+ + A F- + conj(U') Z ; Z' = W Z + H (..)   (U' = U or -iU)
   * to every order, then the affine terms of both E at k = 0. */
  EPGX_OP_FUSED = 12,
  EPGX_OP_CONT = 13,
  EPGX_OP_COUNT = 14
} epgx_opcode;

enum {
  EPGX_FLAG_BASE = 1 << 0,     /* apply to the base state                                      */
  EPGX_FLAG_PARTIALS = 1 << 1, /* apply (without affine term) to the order-1 partial states    */
  /* derivative injection (diff.py:279-286): partial[aux] += form(source), where `form` is the record's code and the
   * source is the current (pre-operator) base state when aux1 == 0 (affine term included) or the partial state of
   * variable aux1 - 1 (order-2 cross terms, diff.py:333-362: no affine term, partial states have no equilibrium) */
  EPGX_FLAG_INJECT = 1 << 2,
  EPGX_FLAG_G = 1 << 3,      /* E: precession phasor present                  */
  EPGX_FLAG_AFFINE = 1 << 4, /* E/DIAG/MATRIX: equilibrium term present       */
  EPGX_FLAG_Z0 = 1 << 5,     /* ADC: read Z(0) instead of F+(0)               */
  EPGX_FLAG_SCALE = 1 << 6,  /* ADC: multiply by the complex factor in blk0   */
  EPGX_FLAG_PRE = 1 << 7,    /* FUSED: E_pre present                          */
  EPGX_FLAG_POST = 1 << 8,   /* FUSED: E_post present (in the CONT record)    */
  EPGX_FLAG_IM = 1 << 9,     /* FUSED: T is of the T_IM kind (else T_RE)      */
  EPGX_FLAG_GEN = 1 << 10,   /* FUSED: T is of the T_GEN kind: blk0 = (a, w, B.re, B.im, U.re, U.im) */
  /* with EPGX_FLAG_PARTIALS: restrict the record to the order-1 partial states (variables < nvar1) or to the
   * order-2 ones (variables >= nvar1): an operator is applied to the order-2 states, then their injections are made
   * from the PRE-operator order-1 states, then the order-1 states follow (diff.py:119-131) */
  EPGX_FLAG_P1 = 1 << 11,
  EPGX_FLAG_P2 = 1 << 12,
  /* ADC in a lattice segment (EPGX_SEG_LATTICE), base state only: read the lattice slot aux1 instead of the slot of
   * k = 0.  Probes that are linear functionals of ALL configurations -- the reference's F0 with accumulated time
   * (statematrix.py:149-156: sum of exp(-|t|) F over the states with zero wavenumber), DFT and Imaging (probe.py:168-219) --
   * read the slots they need into rows of `signal` and the host combines the rows with weights it knows from the lattice */
  EPGX_FLAG_SLOT = 1 << 13
};

/* tape record, 32 bytes */
typedef struct {
  uint16_t code;
  uint16_t flags;
  int32_t aux;
  uint32_t off[3]; /* offset (reals) of coefficient block i in the coefficient table */
  uint8_t pat[3];  /* pattern of block i: entry offset = sum_axis idx[axis]*stride + pool*pool_stride */
  uint8_t rsv;
  int32_t aux1;
  int32_t rsv1;
} epgx_op;

enum {
  EPGX_SEG_RESET = 1 << 0, /* after the pass: state <- equilibrium, order <- 0 (operator.py:297-304) */
  /* the shift truncates at max_nstate (n_new == n_old): what moves above n_new must read as zero */
  EPGX_SEG_MASK_TOP = 1 << 1,
  /* general integer n-d shifts (shift.py:103-117, 297-364): the state is stored on a LATTICE of configurations, one
   * slot per lattice point (full storage: k and -k), every slot 0..nact takes part in the pass and the order k = 0 sits
   * at slot (flags >> 16).  shift == 2 closes such a segment with a gather: new slot j of F+ / F- / Z takes old slot
   * maps[rsv + c (n_new + 1) + j], c = 0, 1, 2 (-1: empty).  Lattice tapes run in the shared-memory kernel. */
  EPGX_SEG_LATTICE = 1 << 2
};

/* segment, 32 bytes: one pass over orders 0..nact applying records [first, first+count), then
 * one unit shift (shift = -1, 0, +1; shift.py:86-101; S(k) lowers to |k| unit shifts), the highest
 * order going from n_old to n_new = min(n_old + |shift|, cap) */
typedef struct {
  int32_t first;
  int32_t count;
  int32_t nact;  /* orders 0..nact are updated (-1: none; <= n_old, fewer when the top orders are unobservable) */
  int32_t shift;
  int32_t n_old;
  int32_t n_new;
  int32_t flags;
  int32_t rsv;   /* EPGX_SEG_LATTICE with shift == 2: offset of the three gather maps in epgx_tape.maps */
} epgx_segment;

/* the lowered sequence (host memory, copied by epgx_plan_create) */
typedef struct {
  int32_t dtype;                /* epgx_dtype */
  int32_t ndim;                 /* grid dimensions WITHOUT the pool axis, outermost first */
  int64_t shape[EPGX_MAX_DIMS]; /* atoms = prod(shape); an atom = all pools of one grid point */
  int32_t npool;                /* exchange compartments coupled by X (1 = none) */
  int32_t npattern;
  int32_t stride[EPGX_MAX_PATTERNS][EPGX_MAX_DIMS]; /* pattern strides, in reals */
  int32_t pool_stride[EPGX_MAX_PATTERNS];
  int64_t nop;
  const epgx_op *ops;
  int64_t nseg;
  const epgx_segment *segs;
  int64_t ncoef;
  const double *coef; /* coefficient table (converted to float for EPGX_F32) */
  /* initial state: block of (init_n+1) x (F+.re, F+.im, F-.re, F-.im, Z.re, Z.im) and block of (M0) */
  uint32_t init_off;
  uint32_t m0_off;
  uint8_t init_pat;
  uint8_t m0_pat;
  uint8_t rsv0[2];
  int32_t init_n;    /* highest order of the initial state */
  int32_t nadc;      /* rows of `signal`   */
  int32_t njac;      /* rows of `jacobian` */
  int32_t nvar;      /* partial state sets: order-1 variables, then order-2 pairs of variables (0 = forward only) */
  int32_t max_order; /* highest configuration order of the whole tape */
  int32_t nvar1;     /* order-1 variables among them (0: all of them) */
  int32_t rsv[2];
  /* variable tiles of the shared-memory kernel: tile i keeps the partial states tiles[3 i .. 3 i + 2] (-1: empty slot)
   * resident next to the base state; an injection's source must sit in the tile of its target -- a pair tile is
   * (a, b, ab).  ntile == 0: consecutive variables (order-1 tapes). */
  int32_t ntile;
  const int32_t *tiles;
  int64_t nmap; /* gather maps of the lattice shifts (EPGX_SEG_LATTICE), concatenated */
  const int32_t *maps;
} epgx_tape;

typedef struct epgx_plan epgx_plan;

/* kernel configuration chosen for a plan (reported for DESIGN/bench/roofline accounting) */
typedef struct {
  int32_t kernel;         /* 0 = ring (state in shared memory), 1 = reg (state in registers), 2 = real (registers,
                             real-valued phase graphs: three reals per order), 3 = realjac (the same with
                             order-1 partial states, the orders of an atom over one or several warps),
                             4 = setjac (the same, one warp per state set), 5 = pulsejac (the same for variables injected
                             once each on bounded graphs: one thread per state set) */
  int32_t lanes_per_atom; /* G */
  int32_t slots_per_lane; /* reg kernel: orders held per lane */
  int32_t vars_per_pass;  /* partial states resident per atom */
  int32_t var_tiles;      /* ceil(nvar / vars_per_pass): grid.y */
  int32_t atoms_per_cta;
  int32_t threads_per_cta;
  int32_t smem_bytes;     /* dynamic shared memory per CTA */
  int32_t ring;           /* ring-buffer length C = max_order + 1 */
  int32_t rsv[3];
  double flops_per_atom;   /* real flops one atom executes through the tape (half storage, base + partials) */
  double updates_per_atom; /* state-updates: sum over T/E/D/X/S records of orders touched */
} epgx_config;

int epgx_version(void);
int epgx_device_count(void);
const char *epgx_last_error(void);

/* validate + copy the tape, choose the kernel variant.  Host-only; no CUDA call. */
int epgx_plan_create(const epgx_tape *tape, epgx_plan **plan);
int epgx_plan_destroy(epgx_plan *plan);
int epgx_plan_config(const epgx_plan *plan, epgx_config *cfg);
/* force a kernel variant (tuning / tests): a 0 / negative argument keeps the automatic choice;
 * kernel: 1 = ring (shared-memory state), 2 = reg (register state; forward, one pool only),
 * 3 = real (register state, real-valued phase graphs only), 4 = realjac (real-valued with partials, orders of an atom
 * over several warps), 5 = setjac (real-valued with at most three partials, one warp per state set), 6 = pulsejac
 * (real-valued, at most 16 orders, every variable injected by one record: one thread per state set) */
int epgx_plan_set_variant(epgx_plan *plan, int kernel, int lanes_per_atom, int vars_per_pass,
                          int atoms_per_cta);

/* diagnostic: the merged record stream the register kernels execute (segment markers and whole-TR groups
 * included; internal codes >= 64 are described in csrc/epgx_common.cuh).  `*records` points into the plan. */
int epgx_plan_stream(const epgx_plan *plan, const epgx_op **records, int64_t *count);

/* bytes of device workspace needed by epgx_plan_upload (tape + segments + coefficient table) */
int epgx_plan_workspace_bytes(const epgx_plan *plan, int64_t *bytes);
/* H2D copy of the tape, segments, patterns and coefficient table into `workspace` on `stream`.
 * `workspace` (256-byte aligned) must stay alive while the plan is simulated from it. */
int epgx_plan_upload(const epgx_plan *plan, void *workspace, void *stream);

/* Run the whole tape for atoms [atom_begin, atom_begin+atom_count) of the enumeration.
 *   signal    device, complex<real>[nadc][atom_count][npool]
 *   jacobian  device, complex<real>[njac][nvar][atom_count][npool], or NULL when nvar == 0 */
int epgx_simulate(const epgx_plan *plan, const void *workspace, int64_t atom_begin, int64_t atom_count,
                  void *signal, void *jacobian, void *stream);

/* Same, writing rows that are `signal_stride` / `jacobian_stride` atoms apart (>= atom_count), so that
 * several launches over atom sub-ranges fill one [row][all atoms][npool] buffer:
 *   signal[(row * signal_stride + a) * npool + pool], a = atom - atom_begin; pass the pointer of the
 *   sub-range's first column. */
int epgx_simulate_strided(const epgx_plan *plan, const void *workspace, int64_t atom_begin, int64_t atom_count,
                          void *signal, int64_t signal_stride, void *jacobian, int64_t jacobian_stride,
                          void *stream);

/* Same as epgx_simulate_strided, and the FINAL state of every atom is written too: what the reference hands back from
 * Operator.__call__(sm) (epgpy/operator.py:96-104) and lets a later simulate(init=sm) resume from
 * (epgpy/functions.py:133-144).  Half storage (orders k >= 0; F-(k) = conj F+(-k), Z(k) = conj Z(-k),
 * epgpy/statematrix.py:418-421):
 *   state  device, complex<real>[atom_count][npool][max_order + 1][3]   columns (F+, F-, Z) of the BASE state;
 *          orders above the final order count of the tape read as zero.
 * Runs the shared-memory (ring) kernel whatever variant the plan prefers for plain runs; the tape must have been
 * lowered with every order kept up to date (no pruning of unobservable orders).  signal / jacobian may be NULL when
 * the tape has no read-out rows. */
int epgx_simulate_state(const epgx_plan *plan, const void *workspace, int64_t atom_begin, int64_t atom_count,
                        void *signal, int64_t signal_stride, void *jacobian, int64_t jacobian_stride, void *state,
                        void *stream);

/* REAL-VALUED signal rows.  On real-valued phase graphs (kernel 2: +-90 degree pulses, no precession, real initial state,
 * read-outs without a complex factor) the imaginary part of every sample is an exact zero; sending only the real parts
 * halves the device->host traffic of a dictionary (8 instead of 16 GB for 1 M atoms x 1000 TRs in FP64), which is what
 * bounds the end-to-end rate.  epgx_plan_real_signal tells whether a plan qualifies; epgx_simulate_real is
 * epgx_simulate_strided with `signal` = real[nadc][signal_stride] (no Jacobian); epgx_expand_real widens rows of reals
 * in HOST memory to the complex rows the reference's API returns (imaginary parts zero), with `nthreads` host threads
 * and streaming stores:  dst[r * dst_pitch + c] = (src[r * src_pitch + c], 0),  pitches in elements. */
int epgx_plan_real_signal(const epgx_plan *plan);
int epgx_simulate_real(const epgx_plan *plan, const void *workspace, int64_t atom_begin, int64_t atom_count, void *signal,
                       int64_t signal_stride, void *stream);
int epgx_expand_real(int dtype, const void *src, int64_t src_pitch, void *dst, int64_t dst_pitch, int64_t rows, int64_t cols,
                     int nthreads);

/* PEER WINDOWS: the final gather of the signal slabs over NVLink without a collective kernel.  Every rank (one process per
 * GPU) allocates its result buffer with epgx_peer_alloc, which also returns a 64-byte CUDA IPC handle; the ranks exchange
 * the handles (any transport) and map each other's buffers with epgx_peer_open.  After the kernel of a column chunk has
 * written the rank's own buffer, epgx_copy2d_device pushes the chunk into every peer's buffer with the COPY ENGINES
 * (cudaMemcpy2DAsync device -> device), at its final place: no SM is taken from the simulation of the next chunk -- an
 * NCCL all-gather's copy kernels only got their CTAs placed once the simulation had drained -- and no scatter pass
 * follows. */
int epgx_peer_alloc(int64_t bytes, void **ptr, char handle[64]);
int epgx_peer_open(const char handle[64], void **ptr);
int epgx_peer_close(void *ptr);
int epgx_peer_free(void *ptr);
int epgx_copy2d_device(void *dst, int64_t dst_pitch, const void *src, int64_t src_pitch, int64_t width, int64_t height,
                       void *stream);

/* asynchronous pitched device->host copy (cudaMemcpy2DAsync) of `height` rows of `width` bytes: brings a
 * column range of the signal slab to (pinned) host memory while the next range is computed */
int epgx_copy2d_to_host(void *dst, int64_t dst_pitch, const void *src, int64_t src_pitch, int64_t width,
                        int64_t height, void *stream);

/* Convenience end-to-end call on HOST buffers (allocates, copies H2D, runs, copies D2H, frees,
 * synchronises): same layouts as epgx_simulate but host pointers. */
int epgx_simulate_host(const epgx_plan *plan, int device, int64_t atom_begin, int64_t atom_count,
                       void *signal, void *jacobian);

/* sum over one axis of a complex array (Adc(reduce=), probe.py:148-153):
 *   out[o][i] = sum_r in[o][r][i],  in: complex<real>[nouter][nred][ninner] (device) */
int epgx_reduce(int dtype, const void *in, void *out, int64_t nouter, int64_t nred, int64_t ninner,
                void *stream);

/* dependent-FMA micro-benchmark: sustained FMA throughput of the CUDA cores in TFLOP/s
 * (2 flops per FMA) for the roofline denominator; synchronises. */
int epgx_fma_peak(int device, int dtype, double seconds, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* EPGX_H */
