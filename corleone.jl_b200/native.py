"""ctypes binding of libcorleone_b200.so (include/corleone_b200.h) -- the only compute path.

There is deliberately no fallback: if the shared library is missing, or no CUDA device is present,
loading / handle creation raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CORLEONE_B200_LIB") or os.path.join(_HERE, "libcorleone_b200.so")
_lib = None

WANT_XEND, WANT_SENS, WANT_TRAJ, WANT_DEFECTS, WANT_OBJ, WANT_GRAD, WANT_FIM, WANT_JAC, WANT_CRIT, WANT_GRAD_COMPACT, WANT_TRAJ_SENS = 1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024
MEM_DEVICE, PS_SOA, OUT_SOA, COMM_REDUCE, COMM_GATHER = 1, 2, 4, 8, 16
TRANSPORT_NONE, TRANSPORT_NCCL, TRANSPORT_P2P = 0, 1, 2
COMM_ID_BYTES = 128

EXPORTS = (
    "cb200_version", "cb200_last_error", "cb200_create", "cb200_destroy", "cb200_get_sizes",
    "cb200_block_structure", "cb200_set_stream", "cb200_eval", "cb200_jac_structure",
    "cb200_last_kernel_ms", "cb200_kernel_ms_history", "cb200_launch_count", "cb200_probe_fp64_tflops",
    "cb200_probe_fp64_mma_tflops", "cb200_forward_init", "cb200_grad_structure", "cb200_set_interval_range",
    "cb200_resident", "cb200_abi_layout", "cb200_comm_unique_id", "cb200_comm_init", "cb200_comm_init_host",
    "cb200_comm_destroy", "cb200_comm_query", "cb200_comm_gathered", "cb200_sample_times",
)

_dp = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)


class Desc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("model", C.c_int32), ("method", C.c_int32), ("steps_per_piece", C.c_int32),
        ("nx", C.c_int32), ("np_all", C.c_int32), ("n_controls", C.c_int32), ("n_tunable_p", C.c_int32),
        ("n_intervals", C.c_int32), ("n_quadrature", C.c_int32), ("device", C.c_int32),
        ("fim_offset", C.c_int32), ("fim_n", C.c_int32), ("fim_mode", C.c_int32),
        ("u0", _dp), ("tunable_p", _i64p), ("control_indices", _i64p), ("quadrature_indices", _i64p),
        ("n_pieces", _i32p), ("tspans", _dp), ("index_grid", _i64p), ("n_u0", _i32p),
        ("n_local_controls", _i32p), ("is_shooting", _i32p), ("ic_sorting", _i64p),
        ("ic_n_constants", _i32p), ("ic_constants", _i64p), ("n_shooting_indices", _i32p),
        ("shooting_indices", _i64p), ("model_data", _dp), ("model_data_len", C.c_int64),
        ("fim_base_nx", C.c_int32), ("n_observed", C.c_int32), ("n_observation_rows", _i32p),
        ("observation_grid", _i64p), ("abstol", C.c_double), ("reltol", C.c_double),
        ("saveat", _dp), ("n_saveat", C.c_int32), ("reserved_desc", C.c_int32),
    ]


class Out(C.Structure):
    _fields_ = [
        ("xend", C.c_void_p), ("sens", C.c_void_p), ("traj", C.c_void_p), ("defects_kind", C.c_void_p),
        ("defects_stage", C.c_void_p), ("objective", C.c_void_p), ("grad", C.c_void_p), ("fim", C.c_void_p),
        ("fim_sum", C.c_void_p), ("objective_sum", C.c_void_p), ("grad_sum", C.c_void_p),
        ("fim_init", C.c_void_p), ("objective_row", C.c_int32), ("reserved", C.c_int32),
        ("jac_vals", C.c_void_p), ("criteria", C.c_void_p), ("grad_compact", C.c_void_p), ("traj_sens", C.c_void_p), ("status", C.c_void_p),
        ("defects_gathered", C.c_void_p), ("comm_B_total", C.c_int64), ("comm_b0", C.c_int64),
        ("resident", C.c_uint32), ("reserved2", C.c_uint32),
    ]


class CommInfo(C.Structure):
    _fields_ = [("nranks", C.c_int32), ("rank", C.c_int32), ("transport", C.c_int32), ("reserved", C.c_int32),
                ("epoch", C.c_int64), ("max_B_total", C.c_int64), ("error", C.c_int64)]


# int (*)(void *ctx, const void *send, int64_t bytes, void *recv)
ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)


class Sizes(C.Structure):
    _fields_ = [(n, C.c_int64) for n in
                ("nvars", "n_defects", "n_samples", "n_row", "n_sens", "n_units", "jac_nnz", "jac_nnz_var", "n_traj_sens")]


class NativeError(RuntimeError):
    pass


class IntegrationFailure(RuntimeError):
    """Some candidates could not be integrated (cb200_out.status: a non-finite end state; in the adaptive mode also
    OrdinaryDiffEq's MaxIters / DtLessThanMin / Unstable).  The reference never inspects the solver's return code and would
    hand NaN objectives, gradients and Jacobians to the optimiser; the host mirror refuses instead."""

    def __init__(self, where, bad):
        self.candidates = bad
        super().__init__(f"{where}: the integration failed for candidate(s) {bad[:8].tolist()}"
                         f"{' ...' if len(bad) > 8 else ''} (status flags set; non-finite or diverging trajectory)")


def check_status(result, where="cb200_eval"):
    st = result.get("status")
    if st is not None and np.any(st):
        raise IntegrationFailure(where, np.nonzero(np.asarray(st).ravel())[0])


def lib():
    """Load libcorleone_b200.so; raises if it has not been built (python corleone.jl_b200/build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(f"{LIB_PATH} not found: build it with `python corleone.jl_b200/build.py` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.cb200_version.restype = C.c_int
        L.cb200_last_error.restype = C.c_char_p
        L.cb200_create.argtypes = [C.POINTER(Desc), C.POINTER(C.c_void_p)]
        L.cb200_destroy.argtypes = [C.c_void_p]
        L.cb200_get_sizes.argtypes = [C.c_void_p, C.POINTER(Sizes)]
        L.cb200_block_structure.argtypes = [C.c_void_p, _i64p]
        L.cb200_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.cb200_set_interval_range.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.cb200_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_uint32, C.c_uint32,
                                 C.POINTER(Out)]
        L.cb200_forward_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_uint32, _i32p, C.c_int32]
        L.cb200_grad_structure.argtypes = [C.c_void_p, C.c_int32, _i64p]
        L.cb200_grad_structure.restype = C.c_int64
        L.cb200_jac_structure.argtypes = [C.c_void_p, _i64p, _i64p, _i64p]
        L.cb200_last_kernel_ms.argtypes = [C.c_void_p]
        L.cb200_last_kernel_ms.restype = C.c_float
        L.cb200_kernel_ms_history.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int]
        L.cb200_kernel_ms_history.restype = C.c_int
        L.cb200_launch_count.argtypes = [C.c_void_p]
        L.cb200_launch_count.restype = C.c_int64
        L.cb200_probe_fp64_tflops.argtypes = [C.c_int, C.c_int]
        L.cb200_probe_fp64_tflops.restype = C.c_double
        L.cb200_probe_fp64_mma_tflops.argtypes = [C.c_int, C.c_int]
        L.cb200_probe_fp64_mma_tflops.restype = C.c_double
        L.cb200_resident.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), _i64p]
        L.cb200_abi_layout.argtypes = [C.c_int32, _i64p, C.c_int32]
        L.cb200_comm_unique_id.argtypes = [C.c_void_p]
        L.cb200_comm_init.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64]
        L.cb200_comm_init_host.argtypes = [C.c_void_p, C.c_int32, C.c_int32, ALLGATHER_FN, C.c_void_p, C.c_int64]
        L.cb200_comm_destroy.argtypes = [C.c_void_p]
        L.cb200_comm_query.argtypes = [C.c_void_p, C.POINTER(CommInfo)]
        L.cb200_comm_gathered.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), _i64p]
        L.cb200_sample_times.argtypes = [C.c_void_p, _dp]
        _lib = L
    return _lib


def _check(rc: int, what: str):
    if rc != 0:
        raise NativeError(f"{what} failed ({rc}): {lib().cb200_last_error().decode()}")


def _arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class NativeLayer:
    """Owns one cb200_handle built from a layer's static tables (the reference's `st`)."""

    def __init__(self, *, model_id, method, steps_per_piece, nx, np_all, u0, tunable_p, control_indices,
                 quadrature_indices, intervals, device=0, fim_offset=0, fim_n=0, model_data=None, fim_mode=0,
                 fim_base_nx=0, n_observed=0, observation_rows=None, abstol=1e-6, reltol=1e-3, saveat=None):
        """intervals: list of dicts with tspans [(t0,t1)...], index_grid (n_controls x M int64 1-based),
        n_u0, n_c, is_shooting, sorting, constants, shooting_indices (all 1-based as the reference).
        observation_rows: per interval an (rows x n_observed) int array = WeightedObservation.grid (fim_mode 2, 3).
        saveat: problem.kwargs saveat (sorted times) or None."""
        L = lib()
        nact = len(control_indices)
        keep = {}
        d = Desc()
        d.struct_size = C.sizeof(Desc)
        d.model, d.method, d.steps_per_piece = model_id, method, steps_per_piece
        d.nx, d.np_all, d.n_controls, d.n_tunable_p = nx, np_all, nact, len(tunable_p)
        d.n_intervals, d.n_quadrature, d.device = len(intervals), len(quadrature_indices), device
        d.fim_offset, d.fim_n, d.fim_mode = fim_offset, fim_n, fim_mode
        d.fim_base_nx, d.n_observed = fim_base_nx, n_observed
        d.abstol, d.reltol = abstol, reltol

        def put(name, values, dtype, ptr):
            a = _arr(values if len(values) else [0], dtype)
            keep[name] = a
            setattr(d, name, a.ctypes.data_as(ptr))

        put("u0", u0, np.float64, _dp)
        put("tunable_p", tunable_p, np.int64, _i64p)
        put("control_indices", control_indices, np.int64, _i64p)
        put("quadrature_indices", quadrature_indices, np.int64, _i64p)
        put("n_pieces", [len(iv["tspans"]) for iv in intervals], np.int32, _i32p)
        put("tspans", [x for iv in intervals for ts in iv["tspans"] for x in ts], np.float64, _dp)
        ig = [np.asarray(iv["index_grid"], dtype=np.int64).reshape(nact, -1).ravel(order="F") if nact
              else np.zeros(0, dtype=np.int64) for iv in intervals]   # no controls: empty index grid
        put("index_grid", np.concatenate(ig) if ig else [], np.int64, _i64p)
        put("n_u0", [iv["n_u0"] for iv in intervals], np.int32, _i32p)
        put("n_local_controls", [iv["n_c"] for iv in intervals], np.int32, _i32p)
        put("is_shooting", [int(iv["is_shooting"]) for iv in intervals], np.int32, _i32p)
        put("ic_sorting", [x for iv in intervals for x in iv["sorting"]], np.int64, _i64p)
        put("ic_n_constants", [len(iv["constants"]) for iv in intervals], np.int32, _i32p)
        put("ic_constants", [x for iv in intervals for x in iv["constants"]], np.int64, _i64p)
        put("n_shooting_indices", [len(iv["shooting_indices"]) for iv in intervals], np.int32, _i32p)
        put("shooting_indices", [x for iv in intervals for x in iv["shooting_indices"]], np.int64, _i64p)
        if model_data is not None:
            put("model_data", model_data, np.float64, _dp)
            d.model_data_len = len(model_data)
        if saveat is not None and len(saveat):
            put("saveat", np.asarray(saveat, dtype=np.float64), np.float64, _dp)
            d.n_saveat = len(saveat)
        if observation_rows is not None:
            rows = [np.asarray(r, dtype=np.int64).reshape(-1, n_observed) for r in observation_rows]
            put("n_observation_rows", [len(r) for r in rows], np.int32, _i32p)
            put("observation_grid", np.concatenate([r.ravel() for r in rows]) if rows else [], np.int64, _i64p)
        self._h = C.c_void_p()
        _check(L.cb200_create(C.byref(d), C.byref(self._h)), "cb200_create")
        s = Sizes()
        _check(L.cb200_get_sizes(self._h, C.byref(s)), "cb200_get_sizes")
        self.sizes = s
        self.nx, self.I, self.nvars = nx, len(intervals), int(s.nvars)
        self.n_defects, self.n_samples, self.n_row = int(s.n_defects), int(s.n_samples), int(s.n_row)
        self.n_sens, self.jac_nnz, self.jac_nnz_var = int(s.n_sens), int(s.jac_nnz), int(s.jac_nnz_var)
        self.n_traj_sens = int(s.n_traj_sens)
        self.fim_n = fim_n
        self.fim_offset = fim_offset
        self.fim_mode = fim_mode
        self.device = device

    def sample_times(self) -> np.ndarray:
        """times of the merged trajectory's samples (piece starts / saveat times / piece ends)"""
        t = np.empty(self.n_samples)
        _check(lib().cb200_sample_times(self._h, t.ctypes.data_as(_dp)), "cb200_sample_times")
        return t

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().cb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- multi-GPU data plane (cb200_comm_*)
    def comm_init(self, nranks: int, rank: int, unique_id: bytes, max_B_total: int = 0):
        """collective: NCCL bootstrap from the 128-byte id rank 0 made with comm_unique_id()"""
        buf = C.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        _check(lib().cb200_comm_init(self._h, nranks, rank, buf, int(max_B_total)), "cb200_comm_init")

    def comm_init_host(self, nranks: int, rank: int, allgather, max_B_total: int = 0):
        """collective: `allgather(send: bytes) -> list[bytes]` is the host's bootstrap channel (a torch.distributed
        all_gather_object, MPI, a threading.Barrier between threads ...)"""
        def _fn(_ctx, send, nbytes, recv):
            try:
                parts = allgather(C.string_at(send, nbytes))
                assert len(parts) == nranks and all(len(p) == nbytes for p in parts)
                C.memmove(recv, b"".join(parts), nbytes * nranks)
                return 0
            except Exception:   # never let an exception cross the C frame
                import traceback
                traceback.print_exc()
                return 1
        cb = ALLGATHER_FN(_fn)
        self._comm_cb = cb   # keep the trampoline alive as long as the communicator may call it (only during init)
        _check(lib().cb200_comm_init_host(self._h, nranks, rank, cb, None, int(max_B_total)), "cb200_comm_init_host")

    def comm_destroy(self):
        _check(lib().cb200_comm_destroy(self._h), "cb200_comm_destroy")

    def comm_info(self) -> dict:
        ci = CommInfo()
        _check(lib().cb200_comm_query(self._h, C.byref(ci)), "cb200_comm_query")
        return {n: int(getattr(ci, n)) for n, _ in CommInfo._fields_ if n != "reserved"}

    def comm_gathered(self):
        """(device pointer, leading dimension) of the all-gathered defects [q * ld + b] in the handle's window"""
        p, ld = C.c_void_p(), C.c_int64()
        _check(lib().cb200_comm_gathered(self._h, C.byref(p), C.byref(ld)), "cb200_comm_gathered")
        return int(p.value), int(ld.value)

    def resident(self, want_bit: int):
        """(device pointer, leading dimension) of a result the last evaluation left on the device (Out.resident)"""
        p, ld = C.c_void_p(), C.c_int64()
        _check(lib().cb200_resident(self._h, want_bit, C.byref(p), C.byref(ld)), "cb200_resident")
        return int(p.value), int(ld.value)

    def set_interval_range(self, first: int, last: int):
        """restrict the following evaluations to intervals first:last (1-based, inclusive; 1:I = all)"""
        _check(lib().cb200_set_interval_range(self._h, int(first), int(last)), "cb200_set_interval_range")

    # ---- queries
    def block_structure(self) -> np.ndarray:
        b = np.zeros(self.I + 1, dtype=np.int64)
        _check(lib().cb200_block_structure(self._h, b.ctypes.data_as(_i64p)), "cb200_block_structure")
        return b

    def jac_structure(self):
        r, c, s = (np.zeros(max(self.jac_nnz, 1), dtype=np.int64) for _ in range(3))
        _check(lib().cb200_jac_structure(self._h, r.ctypes.data_as(_i64p), c.ctypes.data_as(_i64p),
                                         s.ctypes.data_as(_i64p)), "cb200_jac_structure")
        n = self.jac_nnz
        return r[:n], c[:n], s[:n]

    def grad_structure(self, objective_row: int):
        """(first variable (0-based), count) of the contiguous variable range the objective row depends on"""
        first = C.c_int64()
        n = lib().cb200_grad_structure(self._h, objective_row, C.byref(first))
        if n < 0:
            raise NativeError(f"cb200_grad_structure failed ({n}): {lib().cb200_last_error().decode()}")
        return int(first.value) - 1, int(n)

    def set_stream(self, cuda_stream: int):
        _check(lib().cb200_set_stream(self._h, C.c_void_p(cuda_stream)), "cb200_set_stream")

    def last_kernel_ms(self) -> float:
        return float(lib().cb200_last_kernel_ms(self._h))

    def kernel_ms_history(self, n: int) -> np.ndarray:
        buf = (C.c_float * max(n, 1))()
        m = lib().cb200_kernel_ms_history(self._h, buf, n)
        return np.array(buf[:m], dtype=np.float64)

    def launch_count(self) -> int:
        return int(lib().cb200_launch_count(self._h))

    def forward_init(self, ps, fixed_indices=()):
        """forward_initialization for a batch: ps (B, nvars) candidate-major; returns the updated copy."""
        ps = np.array(ps, dtype=np.float64, order="C", copy=True)
        if ps.ndim == 1:
            ps = ps[None, :]
        fx = _arr(list(fixed_indices) if len(fixed_indices) else [0], np.int32)
        _check(lib().cb200_forward_init(self._h, C.c_void_p(ps.ctypes.data), ps.shape[0], max(self.nvars, 1), 0,
                                        fx.ctypes.data_as(_i32p), len(fixed_indices)), "cb200_forward_init")
        return ps

    # ---- evaluation
    def out_sizes(self):
        return dict(xend=self.nx * self.I, sens=self.n_sens, traj=self.n_row * self.n_samples,
                    defects_kind=self.n_defects, defects_stage=self.n_defects, objective=1, grad=self.nvars,
                    fim=self.fim_n * self.fim_n, jac_vals=self.jac_nnz_var, criteria=6, traj_sens=self.n_traj_sens)

    def eval_raw(self, ps_ptr: int, B: int, ldb: int, want: int, flags: int, ptrs: dict, objective_row: int = 1,
                 resident: int = 0, comm_B_total: int = 0, comm_b0: int = 0):
        """Raw pointer call (host or device pointers as `flags` says). ptrs: field name -> int address.
        resident: want-bits of results that stay on the device (NativeLayer.resident)."""
        o = Out()
        for k, v in ptrs.items():
            setattr(o, k, v)
        o.objective_row = objective_row
        o.resident, o.comm_B_total, o.comm_b0 = resident, comm_B_total, comm_b0
        _check(lib().cb200_eval(self._h, C.c_void_p(ps_ptr), B, ldb, want, flags, C.byref(o)), "cb200_eval")

    def prepare_call(self, ps_ptr: int, B: int, ldb: int, want: int, flags: int, ptrs: dict, objective_row: int = 1,
                     resident: int = 0):
        """The same call with the argument block built once: returns a function that issues cb200_eval and nothing else
        (what a Julia `ccall` costs; eval_raw fills a ctypes struct per call, which is microseconds of Python)."""
        o = Out()
        for k, v in ptrs.items():
            setattr(o, k, v)
        o.objective_row = objective_row
        o.resident = resident
        fn, h, ref, psp = lib().cb200_eval, self._h, C.byref(o), C.c_void_p(ps_ptr)

        def call():
            rc = fn(h, psp, B, ldb, want, flags, ref)
            if rc:
                _check(rc, "cb200_eval")
        call._keepalive = o
        return call

    def eval_host(self, ps, want: int, objective_row: int = 1, fim_init=None, pinned_out: dict | None = None,
                  with_sums: bool = False):
        """ps: (B, nvars) C-contiguous float64 == a Julia nvars x B matrix. Returns dict of (B, n) arrays."""
        ps = np.ascontiguousarray(ps, dtype=np.float64)
        if ps.ndim == 1:
            ps = ps[None, :]
        B = ps.shape[0]
        assert ps.shape[1] == self.nvars, (ps.shape, self.nvars)
        sz = self.out_sizes()
        fields = []
        if want & WANT_XEND: fields.append("xend")
        if want & WANT_SENS: fields.append("sens")
        if want & WANT_TRAJ: fields.append("traj")
        if want & WANT_DEFECTS: fields += ["defects_kind", "defects_stage"]
        if want & WANT_OBJ: fields.append("objective")
        if want & WANT_GRAD: fields.append("grad")
        if want & WANT_FIM: fields.append("fim")
        if want & WANT_JAC: fields.append("jac_vals")
        if want & WANT_CRIT: fields.append("criteria")
        if want & WANT_TRAJ_SENS: fields.append("traj_sens")
        if want & WANT_GRAD_COMPACT:
            fields.append("grad_compact")
            sz["grad_compact"] = self.grad_structure(objective_row)[1]
        res = {}
        for f in fields:
            if pinned_out is not None and f in pinned_out:
                res[f] = pinned_out[f]
            else:
                res[f] = np.empty((B, sz[f]), dtype=np.float64)
        ptrs = {f: res[f].ctypes.data for f in fields}
        if want & WANT_FIM:
            res["fim_sum"] = np.zeros(sz["fim"], dtype=np.float64)
            ptrs["fim_sum"] = res["fim_sum"].ctypes.data
        if (want & (WANT_FIM | WANT_CRIT)) and fim_init is not None:
            fi = np.asfortranarray(fim_init, dtype=np.float64)   # kept alive until the call returns
            ptrs["fim_init"] = fi.ctypes.data
        res["status"] = np.zeros(B, dtype=np.int32)   # failure flags per candidate (bit 0: non-finite state)
        ptrs["status"] = res["status"].ctypes.data
        if with_sums:   # cross-candidate sums accumulated on the device (the per-rank partials of a sharded run)
            if want & WANT_OBJ:
                res["objective_sum"] = np.zeros(1, dtype=np.float64)
                ptrs["objective_sum"] = res["objective_sum"].ctypes.data
            if want & WANT_GRAD:
                res["grad_sum"] = np.zeros(max(self.nvars, 1), dtype=np.float64)[: self.nvars]
                ptrs["grad_sum"] = res["grad_sum"].ctypes.data
        if self.nvars == 0:  # nothing tunable: still one candidate per row
            ps = np.zeros((B, 1))
        self.eval_raw(ps.ctypes.data, B, max(self.nvars, 1), want, 0, ptrs, objective_row)
        return res


def comm_unique_id() -> bytes:
    """the 128-byte NCCL id rank 0 creates and the host hands to the other ranks"""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    _check(lib().cb200_comm_unique_id(buf), "cb200_comm_unique_id")
    return buf.raw


def abi_layout(which: int) -> list:
    v = (C.c_int64 * 16)()
    n = lib().cb200_abi_layout(which, v, 16)
    return [int(x) for x in v[:n]]


def probe_fp64_tflops(device: int = 0, iters: int = 20000) -> float:
    return float(lib().cb200_probe_fp64_tflops(device, iters))


def probe_fp64_mma_tflops(device: int = 0, iters: int = 20000) -> float:
    return float(lib().cb200_probe_fp64_mma_tflops(device, iters))
