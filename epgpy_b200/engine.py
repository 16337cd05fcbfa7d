"""ctypes binding of libepgx.so (include/epgx.h) + device-buffer plumbing.

torch is used for device memory, streams and (in multi-process runs) torch.distributed only; every
arithmetic step of the EPG path is a launch of the hand-written sm_100a kernels through the C ABI.
There is NO CPU fallback: without the shared library or without a CUDA device the calls raise.
"""

import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EPGX_LIB") or os.path.join(_HERE, "libepgx.so")  # EPGX_LIB: experimental builds

MAX_DIMS, MAX_PATTERNS = 8, 64


class EpgxError(RuntimeError):
    pass


class _Tape(ctypes.Structure):
    _fields_ = [
        ("dtype", ctypes.c_int32), ("ndim", ctypes.c_int32), ("shape", ctypes.c_int64 * MAX_DIMS),
        ("npool", ctypes.c_int32), ("npattern", ctypes.c_int32),
        ("stride", (ctypes.c_int32 * MAX_DIMS) * MAX_PATTERNS), ("pool_stride", ctypes.c_int32 * MAX_PATTERNS),
        ("nop", ctypes.c_int64), ("ops", ctypes.c_void_p), ("nseg", ctypes.c_int64), ("segs", ctypes.c_void_p),
        ("ncoef", ctypes.c_int64), ("coef", ctypes.c_void_p),
        ("init_off", ctypes.c_uint32), ("m0_off", ctypes.c_uint32), ("init_pat", ctypes.c_uint8),
        ("m0_pat", ctypes.c_uint8), ("rsv0", ctypes.c_uint8 * 2), ("init_n", ctypes.c_int32),
        ("nadc", ctypes.c_int32), ("njac", ctypes.c_int32), ("nvar", ctypes.c_int32), ("max_order", ctypes.c_int32),
        ("nvar1", ctypes.c_int32), ("rsv", ctypes.c_int32 * 2), ("ntile", ctypes.c_int32), ("tiles", ctypes.c_void_p),
        ("nmap", ctypes.c_int64), ("maps", ctypes.c_void_p),
    ]


class Config(ctypes.Structure):
    _fields_ = [
        ("kernel", ctypes.c_int32), ("lanes_per_atom", ctypes.c_int32), ("slots_per_lane", ctypes.c_int32),
        ("vars_per_pass", ctypes.c_int32), ("var_tiles", ctypes.c_int32), ("atoms_per_cta", ctypes.c_int32),
        ("threads_per_cta", ctypes.c_int32), ("smem_bytes", ctypes.c_int32), ("ring", ctypes.c_int32),
        ("rsv", ctypes.c_int32 * 3), ("flops_per_atom", ctypes.c_double), ("updates_per_atom", ctypes.c_double),
    ]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "rsv"}


_lib = None
_lock = threading.Lock()
LAUNCHES = 0  # kernel launches issued through this binding (bench.py reports the count of its timed regions)
H2D_BYTES = 0  # bytes this binding copied host -> device (tape + coefficient table uploads)
D2H_BYTES = 0  # bytes this binding copied device -> host (result rows)


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


def _moved(h2d=0, d2h=0):
    global H2D_BYTES, D2H_BYTES
    H2D_BYTES += h2d
    D2H_BYTES += d2h

EXPORTS = [
    "epgx_version", "epgx_device_count", "epgx_last_error", "epgx_plan_create", "epgx_plan_destroy",
    "epgx_plan_config", "epgx_plan_stream", "epgx_plan_set_variant", "epgx_plan_workspace_bytes", "epgx_plan_upload",
    "epgx_simulate", "epgx_simulate_strided", "epgx_simulate_state", "epgx_plan_real_signal", "epgx_simulate_real",
    "epgx_expand_real", "epgx_peer_alloc", "epgx_peer_open", "epgx_peer_close", "epgx_peer_free", "epgx_copy2d_device",
    "epgx_copy2d_to_host", "epgx_simulate_host", "epgx_reduce",
    "epgx_fma_peak",
]


def lib():
    """load libepgx.so (built in-tree by __graft_entry__.build / csrc/build.py); raises if missing"""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise EpgxError(
                    f"{LIB_PATH} not found: build it with `python -m epgpy_b200.build` (nvcc, sm_100a). "
                    "There is no CPU fallback."
                )
            L = ctypes.CDLL(LIB_PATH)
            vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
            L.epgx_version.restype = i32
            L.epgx_device_count.restype = i32
            L.epgx_last_error.restype = ctypes.c_char_p
            L.epgx_plan_create.argtypes = [ctypes.POINTER(_Tape), ctypes.POINTER(vp)]
            L.epgx_plan_destroy.argtypes = [vp]
            L.epgx_plan_config.argtypes = [vp, ctypes.POINTER(Config)]
            L.epgx_plan_stream.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(i64)]
            L.epgx_plan_set_variant.argtypes = [vp, i32, i32, i32, i32]
            L.epgx_plan_workspace_bytes.argtypes = [vp, ctypes.POINTER(i64)]
            L.epgx_plan_upload.argtypes = [vp, vp, vp]
            L.epgx_simulate.argtypes = [vp, vp, i64, i64, vp, vp, vp]
            L.epgx_simulate_strided.argtypes = [vp, vp, i64, i64, vp, i64, vp, i64, vp]
            L.epgx_simulate_state.argtypes = [vp, vp, i64, i64, vp, i64, vp, i64, vp, vp]
            L.epgx_plan_real_signal.argtypes = [vp]
            L.epgx_simulate_real.argtypes = [vp, vp, i64, i64, vp, i64, vp]
            L.epgx_expand_real.argtypes = [i32, vp, i64, vp, i64, i64, i64, i32]
            L.epgx_peer_alloc.argtypes = [i64, ctypes.POINTER(vp), ctypes.c_char_p]
            L.epgx_peer_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(vp)]
            L.epgx_peer_close.argtypes = [vp]
            L.epgx_peer_free.argtypes = [vp]
            L.epgx_copy2d_device.argtypes = [vp, i64, vp, i64, i64, i64, vp]
            L.epgx_copy2d_to_host.argtypes = [vp, i64, vp, i64, i64, i64, vp]
            L.epgx_simulate_host.argtypes = [vp, i32, i64, i64, vp, vp]
            L.epgx_reduce.argtypes = [i32, vp, vp, i64, i64, i64, vp]
            L.epgx_fma_peak.argtypes = [i32, i32, ctypes.c_double, ctypes.POINTER(ctypes.c_double)]
            _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        msg = lib().epgx_last_error().decode(errors="replace")
        exc = {-2: MemoryError, -4: NotImplementedError}.get(rc, EpgxError)
        raise exc(f"epgx error {rc}: {msg}")


DTYPES = {"f64": 0, "f32": 1, "float64": 0, "float32": 1, "complex128": 0, "complex64": 1}


def norm_dtype(dtype):
    key = np.dtype(dtype).name if not isinstance(dtype, str) else dtype
    if key not in DTYPES:
        raise ValueError(f"dtype must be float64 or float32, got {dtype!r}")
    return "f64" if DTYPES[key] == 0 else "f32"


class Plan:
    """a lowered sequence registered with the engine (host object; owns the C plan)"""

    def __init__(self, low):
        self.low = low
        L = lib()
        t = _Tape()
        t.dtype = DTYPES[low.dtype]
        ashape = low.atom_shape
        t.ndim = len(ashape)
        for i, d in enumerate(ashape):
            t.shape[i] = d
        t.npool = low.npool
        t.npattern = len(low.patterns)
        for q, (strides, ps) in enumerate(low.patterns):
            for i, s in enumerate(strides):
                t.stride[q][i] = s
            t.pool_stride[q] = ps
        self._ops = np.ascontiguousarray(low.ops)
        self._segs = np.ascontiguousarray(low.segs)
        self._coef = np.ascontiguousarray(low.coef, dtype=np.float64)
        t.nop, t.ops = len(self._ops), self._ops.ctypes.data
        t.nseg, t.segs = len(self._segs), self._segs.ctypes.data
        t.ncoef, t.coef = len(self._coef), self._coef.ctypes.data
        t.init_off, t.init_pat = low.init_ref
        t.m0_off, t.m0_pat = low.m0_ref
        t.init_n = low.init_n
        t.nadc, t.njac, t.nvar, t.max_order = low.nadc, low.njac, low.nvar, low.max_order
        t.nvar1 = getattr(low, "nvar1", low.nvar)
        self._tiles = np.ascontiguousarray(getattr(low, "tiles", np.zeros((0, 3))), dtype=np.int32)
        t.ntile, t.tiles = len(self._tiles), (self._tiles.ctypes.data if len(self._tiles) else None)
        self._maps = np.ascontiguousarray(getattr(low, "maps", np.zeros(0)), dtype=np.int32)
        t.nmap, t.maps = len(self._maps), (self._maps.ctypes.data if len(self._maps) else None)
        self._h = ctypes.c_void_p()
        _check(L.epgx_plan_create(ctypes.byref(t), ctypes.byref(self._h)))
        self._ws = {}  # device index -> (workspace tensor, upload event)
        self._ws_lock = threading.Lock()

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.epgx_plan_destroy(h)

    @property
    def handle(self):
        return self._h

    def config(self):
        c = Config()
        _check(lib().epgx_plan_config(self._h, ctypes.byref(c)))
        return c.asdict()

    def stream(self):
        """diagnostic: the merged record stream of the register kernels as a structured array (a copy)"""
        ptr, n = ctypes.c_void_p(), ctypes.c_int64()
        _check(lib().epgx_plan_stream(self._h, ctypes.byref(ptr), ctypes.byref(n)))
        from .lowering import OP_DTYPE
        if not n.value:
            return np.zeros(0, dtype=OP_DTYPE)
        buf = (ctypes.c_char * (n.value * OP_DTYPE.itemsize)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=OP_DTYPE).copy()

    def set_variant(self, kernel=0, lanes_per_atom=0, vars_per_pass=0, atoms_per_cta=0):
        _check(lib().epgx_plan_set_variant(self._h, kernel, lanes_per_atom, vars_per_pass, atoms_per_cta))

    def workspace_bytes(self):
        n = ctypes.c_int64()
        _check(lib().epgx_plan_workspace_bytes(self._h, ctypes.byref(n)))
        return n.value

    # ---- device side
    def upload(self, device, force=False):
        """H2D of tape + coefficient table (cached per device unless `force`); returns the workspace tensor.
        The copies are enqueued on torch's current stream of the device; an event recorded behind them is waited on by
        every later launch stream (`_wait_upload`), so a run on another stream / thread never reads a partly
        uploaded table."""
        import torch

        dev = torch.device("cuda", device)
        with self._ws_lock:
            if device not in self._ws or force:
                with torch.cuda.device(dev):
                    ws = self._ws.get(device, (None, None))[0]
                    if ws is None:
                        ws = torch.empty(self.workspace_bytes(), dtype=torch.uint8, device=dev)
                    stream = torch.cuda.current_stream(dev)
                    _check(lib().epgx_plan_upload(self._h, ws.data_ptr(), stream.cuda_stream))
                    _moved(h2d=ws.numel())
                    ev = torch.cuda.Event()
                    ev.record(stream)
                self._ws[device] = (ws, ev)
            return self._ws[device][0]

    def _wait_upload(self, device, stream):
        ev = self._ws[device][1]
        stream.wait_event(ev)

    def _check_buffer(self, t, name, rows, atoms, device=None):
        """caller-supplied buffers: dtype, contiguity, size (and device) -- a wrong buffer would be written out of bounds"""
        import torch

        cdt = torch.complex128 if self.low.dtype == "f64" else torch.complex64
        need = rows * atoms * self.low.npool
        if isinstance(t, np.ndarray):
            ok = t.dtype == (np.complex128 if self.low.dtype == "f64" else np.complex64) and t.flags.c_contiguous and t.size >= need
            if device is not None:
                raise ValueError(f"{name}: expected a CUDA tensor")
        else:
            ok = t.dtype == cdt and t.is_contiguous() and t.numel() >= need
            if device is not None and (not t.is_cuda or t.device.index != device):
                raise ValueError(f"{name}: expected a tensor on cuda:{device}, got {t.device}")
            if device is None and t.is_cuda:
                raise ValueError(f"{name}: expected a host buffer, got {t.device}")
        if not ok:
            raise ValueError(f"{name}: need a contiguous {cdt} buffer of at least {rows} x {atoms} x {self.low.npool} elements")

    def run(self, device, atom_begin=0, atom_count=None, signal=None, jacobian=None):
        """launch the fused kernel for an atom range on torch's current stream of `device`.
        Returns (signal, jacobian) device tensors: complex [nadc][atoms][npool], [njac][nvar][atoms][npool]"""
        import torch

        require_cuda()
        low = self.low
        if atom_count is None:
            atom_count = low.natoms - atom_begin
        dev = torch.device("cuda", device)
        cdt = torch.complex128 if low.dtype == "f64" else torch.complex64
        has_jac = bool(low.nvar and low.njac)
        with torch.cuda.device(dev):
            ws = self.upload(device)
            if signal is None:
                signal = torch.empty((low.nadc, atom_count, low.npool), dtype=cdt, device=dev)
            else:
                self._check_buffer(signal, "signal", low.nadc, atom_count, device)
            if jacobian is None and has_jac:
                jacobian = torch.empty((low.njac, low.nvar, atom_count, low.npool), dtype=cdt, device=dev)
            elif has_jac:
                self._check_buffer(jacobian, "jacobian", low.njac * low.nvar, atom_count, device)
            if atom_count == 0:
                return signal, jacobian
            stream = torch.cuda.current_stream(dev)
            self._wait_upload(device, stream)
            _check(lib().epgx_simulate(self._h, ws.data_ptr(), atom_begin, atom_count,
                                       signal.data_ptr() if signal is not None and signal.numel() else None,
                                       jacobian.data_ptr() if jacobian is not None and jacobian.numel() else None,
                                       stream.cuda_stream))
            _count()
        return signal, jacobian

    def run_state(self, device, atom_begin=0, atom_count=None):
        """epgx_simulate_state: run the tape and read the FINAL base state of every atom back (the shared-memory kernel;
        lower with prune_unobservable=False).  Returns device tensors (signal, jacobian, state) with
        state complex [atoms][npool][max_order + 1][3] in half storage, columns (F+, F-, Z)."""
        import torch

        require_cuda()
        low = self.low
        if atom_count is None:
            atom_count = low.natoms - atom_begin
        dev = torch.device("cuda", device)
        cdt = torch.complex128 if low.dtype == "f64" else torch.complex64
        has_jac = bool(low.nvar and low.njac)
        with torch.cuda.device(dev):
            ws = self.upload(device)
            signal = torch.empty((low.nadc, atom_count, low.npool), dtype=cdt, device=dev)
            jacobian = torch.empty((low.njac, low.nvar, atom_count, low.npool), dtype=cdt, device=dev) if has_jac else None
            state = torch.zeros((atom_count, low.npool, low.max_order + 1, 3), dtype=cdt, device=dev)
            if atom_count:
                stream = torch.cuda.current_stream(dev)
                self._wait_upload(device, stream)
                _check(lib().epgx_simulate_state(self._h, ws.data_ptr(), atom_begin, atom_count,
                                                 signal.data_ptr() if signal.numel() else None, atom_count,
                                                 jacobian.data_ptr() if has_jac and jacobian.numel() else None, atom_count,
                                                 state.data_ptr(), stream.cuda_stream))
                _count()
        return signal, jacobian, state

    def run_strided(self, device, atom_begin, atom_count, signal, signal_stride, jacobian=None, jacobian_stride=0):
        """epgx_simulate_strided on torch's current stream: rows of `signal` are `signal_stride` atoms apart, so that
        launches over atom sub-ranges fill one [row][atoms][npool] buffer (pass the view that starts at the range's
        first column)"""
        import torch

        require_cuda()
        dev = torch.device("cuda", device)
        with torch.cuda.device(dev):
            ws = self.upload(device)
            stream = torch.cuda.current_stream(dev)
            self._wait_upload(device, stream)
            _check(lib().epgx_simulate_strided(self._h, ws.data_ptr(), atom_begin, atom_count, signal.data_ptr(), signal_stride,
                                               jacobian.data_ptr() if jacobian is not None else None,
                                               jacobian_stride or signal_stride, stream.cuda_stream))
            _count()

    def run_to_host(self, device, out_signal, atom_begin=0, atom_count=None, nchunk=8, out_jacobian=None,
                    dev_signal=None, dev_jacobian=None, host_col=0, host_atoms=None):
        """pipelined end-to-end run: H2D of the tape, then the atom range is cut in `nchunk` column ranges;
        range i+1 is computed while range i is copied to the (pinned) host buffers with pitched D2H copies.
        out_signal: host torch tensor / numpy array [nadc][host_atoms][npool] complex (pinned for overlap); the range
        lands in columns [host_col, host_col + atom_count) (host_atoms defaults to atom_count: the buffer holds exactly
        this range).  out_jacobian: [njac * nvar][host_atoms][npool], required when the plan has derivative rows.
        Synchronises before returning.  Returns the device slabs (dev_signal, dev_jacobian)."""
        import torch

        require_cuda()
        low = self.low
        if atom_count is None:
            atom_count = low.natoms - atom_begin
        if host_atoms is None:
            host_atoms = atom_count
        if host_col < 0 or host_col + atom_count > host_atoms:
            raise ValueError("host column range outside the host buffer")
        dev = torch.device("cuda", device)
        cdt = torch.complex128 if low.dtype == "f64" else torch.complex64
        csz = 16 if low.dtype == "f64" else 8
        has_jac = bool(low.nvar and low.njac)
        if low.nadc:
            if out_signal is None:
                raise ValueError("out_signal is required: the plan has read-out rows")
            self._check_buffer(out_signal, "out_signal", low.nadc, host_atoms)
        if has_jac:
            if out_jacobian is None:
                raise ValueError("out_jacobian is required: the plan has derivative rows")
            self._check_buffer(out_jacobian, "out_jacobian", low.njac * low.nvar, host_atoms)
        L = lib()
        with torch.cuda.device(dev):
            ws = self.upload(device, force=True)
            if dev_signal is None:
                dev_signal = torch.empty((low.nadc, atom_count, low.npool), dtype=cdt, device=dev)
            else:
                self._check_buffer(dev_signal, "dev_signal", low.nadc, atom_count, device)
            if has_jac and dev_jacobian is None:
                dev_jacobian = torch.empty((low.njac * low.nvar, atom_count, low.npool), dtype=cdt, device=dev)
            elif has_jac:
                self._check_buffer(dev_jacobian, "dev_jacobian", low.njac * low.nvar, atom_count, device)
            if atom_count == 0:
                return dev_signal, dev_jacobian
            compute = torch.cuda.current_stream(dev)
            self._wait_upload(device, compute)
            copy = torch.cuda.Stream(dev)
            per = -(-atom_count // max(1, nchunk))
            ptr = lambda t: t.data_ptr() if hasattr(t, "data_ptr") else t.ctypes.data
            hs = ptr(out_signal) + host_col * low.npool * csz if low.nadc else None
            hj = ptr(out_jacobian) + host_col * low.npool * csz if has_jac else None
            pitch = atom_count * low.npool * csz
            hpitch = host_atoms * low.npool * csz
            for b in range(0, atom_count, per):
                c = min(per, atom_count - b)
                colb = b * low.npool * csz
                _check(L.epgx_simulate_strided(
                    self._h, ws.data_ptr(), atom_begin + b, c, dev_signal.data_ptr() + colb, atom_count,
                    (dev_jacobian.data_ptr() + colb) if has_jac else None, atom_count, compute.cuda_stream))
                _count()
                ev = torch.cuda.Event()
                ev.record(compute)
                copy.wait_event(ev)
                if low.nadc:
                    _check(L.epgx_copy2d_to_host(hs + colb, hpitch, dev_signal.data_ptr() + colb, pitch, c * low.npool * csz,
                                                 low.nadc, copy.cuda_stream))
                    _moved(d2h=c * low.npool * csz * low.nadc)
                if has_jac:
                    _check(L.epgx_copy2d_to_host(hj + colb, hpitch, dev_jacobian.data_ptr() + colb, pitch,
                                                 c * low.npool * csz, low.njac * low.nvar, copy.cuda_stream))
                    _moved(d2h=c * low.npool * csz * low.njac * low.nvar)
            copy.synchronize()
            compute.synchronize()
        return dev_signal, dev_jacobian

    def real_signal(self):
        """True when the signal of the plan is real-valued and the kernel can emit rows of reals (epgx_plan_real_signal)"""
        return bool(lib().epgx_plan_real_signal(self._h))

    def run_to_host_real(self, device, out_signal, atom_begin=0, atom_count=None, nchunk=16, host_col=0, host_atoms=None,
                         nthreads=None, nstage=4):
        """`run_to_host` for plans with a real-valued signal (`real_signal()`): only the real parts cross PCIe -- half the
        bytes of the complex result, and the device->host copy is what bounds the end-to-end rate of a dictionary.
        Pipeline per column chunk: kernel (rows of reals into one of two device buffers) -> D2H into a pinned staging
        ring -> `epgx_expand_real` (host threads, streaming stores) into the complex result `out_signal`
        [nadc][host_atoms][1].  The expansion of chunk i runs while chunk i + 1 is copied and chunk i + 2 computed.
        Synchronises before returning."""
        import queue
        import threading as _th

        import torch

        require_cuda()
        low = self.low
        if not self.real_signal() or low.npool != 1:
            raise EpgxError("this plan has no real-valued signal")
        if atom_count is None:
            atom_count = low.natoms - atom_begin
        if host_atoms is None:
            host_atoms = atom_count
        if host_col < 0 or host_col + atom_count > host_atoms:
            raise ValueError("host column range outside the host buffer")
        self._check_buffer(out_signal, "out_signal", low.nadc, host_atoms)
        if atom_count == 0 or low.nadc == 0:
            return
        if nthreads is None:
            nthreads = int(os.environ.get("EPGX_HOST_THREADS", 0)) or max(1, min(16, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", 1)))))
        dev = torch.device("cuda", device)
        rdt = torch.float64 if low.dtype == "f64" else torch.float32
        rsz = 8 if low.dtype == "f64" else 4
        L = lib()
        per = -(-atom_count // max(1, nchunk))
        nstage = max(2, nstage)
        dst0 = (out_signal.data_ptr() if hasattr(out_signal, "data_ptr") else out_signal.ctypes.data) + host_col * 2 * rsz
        with torch.cuda.device(dev):
            ws = self.upload(device, force=True)
            dbuf = [torch.empty((low.nadc, per), dtype=rdt, device=dev) for _ in range(2)]
            stage = [torch.empty((low.nadc, per), dtype=rdt, pin_memory=True) for _ in range(nstage)]
            compute = torch.cuda.current_stream(dev)
            self._wait_upload(device, compute)
            copy = torch.cuda.Stream(dev)
            jobs, errors = queue.Queue(), []
            free = [_th.Event() for _ in range(nstage)]
            for ev in free:
                ev.set()

            def expander():
                while True:
                    job = jobs.get()
                    if job is None:
                        return
                    slot, b, c, ev = job
                    try:
                        ev.synchronize()  # the D2H copy of this chunk has landed in the staging slot
                        rc = L.epgx_expand_real(DTYPES[low.dtype], stage[slot].data_ptr(), per, dst0 + b * 2 * rsz, host_atoms,
                                                low.nadc, c, nthreads)
                        if rc != 0:
                            errors.append(EpgxError(f"epgx_expand_real: {rc}"))
                    except BaseException as ex:  # surfaced by the caller
                        errors.append(ex)
                    finally:
                        free[slot].set()

            worker = _th.Thread(target=expander, daemon=True)
            worker.start()
            copied = [None, None]  # event behind the copy that last read device buffer k
            try:
                for i, b in enumerate(range(0, atom_count, per)):
                    c = min(per, atom_count - b)
                    k, slot = i % 2, i % nstage
                    if copied[k] is not None:
                        compute.wait_event(copied[k])
                    _check(L.epgx_simulate_real(self._h, ws.data_ptr(), atom_begin + b, c, dbuf[k].data_ptr(), per, compute.cuda_stream))
                    _count()
                    done = torch.cuda.Event()
                    done.record(compute)
                    free[slot].wait()   # the expansion that last used this staging slot is over
                    free[slot].clear()
                    copy.wait_event(done)
                    with torch.cuda.stream(copy):
                        stage[slot].copy_(dbuf[k], non_blocking=True)
                        _moved(d2h=dbuf[k].numel() * rsz)
                        ev = torch.cuda.Event()
                        ev.record(copy)
                    copied[k] = ev
                    jobs.put((slot, b, c, ev))
            finally:
                jobs.put(None)
                worker.join()
            copy.synchronize()
            compute.synchronize()
            if errors:
                raise errors[0]

    def run_host(self, device, atom_begin, atom_count, signal, jacobian=None):
        """epgx_simulate_host: the plain C-ABI call on HOST numpy buffers (alloc + H2D + run + D2H)"""
        require_cuda()
        _check(lib().epgx_simulate_host(self._h, device, atom_begin, atom_count,
                                        signal.ctypes.data if signal is not None and signal.size else None,
                                        jacobian.ctypes.data if jacobian is not None and jacobian.size else None))


def require_cuda():
    if lib().epgx_device_count() < 1:
        raise EpgxError("no CUDA device visible: the epgx engine has no CPU fallback")


def device_reduce(t, axis):
    """sum a complex device tensor over one axis with the engine's reduction kernel (Adc(reduce=))"""
    import torch

    t = t.contiguous()
    shape = list(t.shape)
    nouter = int(np.prod(shape[:axis])) if axis else 1
    nred = shape[axis]
    ninner = int(np.prod(shape[axis + 1:])) if axis + 1 < len(shape) else 1
    out = torch.empty(shape[:axis] + shape[axis + 1:], dtype=t.dtype, device=t.device)
    with torch.cuda.device(t.device):
        st = torch.cuda.current_stream(t.device).cuda_stream
        _check(lib().epgx_reduce(0 if t.dtype == torch.complex128 else 1, t.data_ptr(), out.data_ptr(), nouter, nred,
                                 ninner, st))
    return out


def fma_peak(device=0, dtype="f64", seconds=0.3):
    """measured CUDA-core FMA throughput (TFLOP/s) -- the roofline denominator of this path"""
    require_cuda()
    out = ctypes.c_double()
    _check(lib().epgx_fma_peak(device, DTYPES[norm_dtype(dtype)], seconds, ctypes.byref(out)))
    return out.value
