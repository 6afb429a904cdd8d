"""ctypes binding of libdexb200.so (include/dexb200.h) plus thin object wrappers.

This is the only module that touches the native library.  PyTorch is used for what
the task calls plumbing — device memory, streams, pinned buffers, torch.distributed —
and never for arithmetic: every number the evaluation entry points return is computed by
the CUDA kernels in csrc/.  There is no CPU fallback: if the library is missing, or no
CUDA device is visible, the calls below raise.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

from .node import MAX_DEGREE, WIRE_DTYPE, Node, to_wire_population
from .operators import OperatorEnum

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DEXB200_LIB") or os.path.join(_HERE, "lib", "libdexb200.so")

OK = 0
F32, F64 = 0, 1
EVAL_EARLY_EXIT = 1
EVAL_SKIP_INCOMPLETE = 2
PACK_FUSED, PACK_BUMPER = 1, 2
GRAD_CONSTANTS, GRAD_FEATURES, GRAD_BOTH = 0, 1, 2

# every symbol include/dexb200.h declares: (name, restype, argtypes)
_P, _I32, _I64, _U8P = C.c_void_p, C.c_int32, C.c_int64, C.c_void_p
ABI = {
    "dex_abi_version": (C.c_int, []),
    "dex_strerror": (C.c_char_p, [C.c_int]),
    "dex_device_count": (C.c_int, []),
    "dex_opcode_from_name": (C.c_int, [C.c_char_p, C.c_int]),
    "dex_opcode_name": (C.c_char_p, [C.c_int]),
    "dex_opcode_degree": (C.c_int, [C.c_int]),
    "dex_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "dex_ctx_destroy": (C.c_int, [_P]),
    "dex_ctx_set_stream": (C.c_int, [_P, _P]),
    "dex_ctx_use_own_stream": (C.c_int, [_P]),
    "dex_ctx_synchronize": (C.c_int, [_P]),
    "dex_last_error": (C.c_char_p, [_P]),
    "dex_optable_create": (C.c_int, [_P, _P, C.c_int, C.POINTER(_P)]),
    "dex_optable_destroy": (C.c_int, [_P]),
    "dex_population_pack": (C.c_int, [_P, _P, _P, _P, _I64, C.c_int, C.c_int, C.POINTER(_P)]),
    "dex_population_destroy": (C.c_int, [_P]),
    "dex_population_get_info": (C.c_int, [_P, _P]),
    "dex_population_constant_counts": (C.c_int, [_P, _P]),
    "dex_population_get_constants": (C.c_int, [_P, _P, _P, _I64]),
    "dex_population_set_constants": (C.c_int, [_P, _P, _P, _I64]),
    "dex_eval": (C.c_int, [_P, _P, _P, _I32, _I64, _I64, _P, _I64, _U8P, C.c_int]),
    "dex_eval_parametric": (C.c_int, [_P, _P, _P, _I32, _I64, _I64, _P, _I32, _I32, _P, _P, _I64,
                                      _U8P, C.c_int]),
    "dex_eval_grad": (C.c_int, [_P, _P, _P, _I32, _I64, _I64, C.c_int, _P, _I64, _P, _P, _U8P]),
    "dex_eval_grad_parametric": (C.c_int, [_P, _P, _P, _I32, _I64, _I64, _P, _I32, _I32, _P, C.c_int, _P, _I64, _P,
                                           _P, _U8P]),
    "dex_grad_offsets": (C.c_int, [_P, _I32, _I64, C.c_int, _P]),
    "dex_eval_diff": (C.c_int, [_P, _P, _P, _I32, _I64, _I64, _I32, _P, _P, _I64, _U8P]),
    "dex_eval_loss": (C.c_int, [_P, _P, _P, _I32, _I64, _I64, _P, _P, _P, _U8P, C.c_int]),
    "dex_eval_loss_grad": (C.c_int, [_P, _P, _P, _I32, _I64, _I64, _P, _P, C.c_int, _P, _P, _P, _U8P]),
    "dex_eval_host": (C.c_int, [_P, _P, _P, _I32, _I64, _I64, _P, _I64, _U8P, C.c_int]),
    "dex_shard_eval_host": (C.c_int, [_P, _P, _I32, _P, _I32, _I64, _I64, _P, _I64, _U8P, C.c_int]),
    "dex_shard_eval": (C.c_int, [_P, _P, _I32, _P, _I32, _I64, _I64, _P, _I64, _U8P, _I32, C.c_int]),
    "dex_copy_to_device": (C.c_int, [_P, _P, _P, _I64]),
    "dex_copy_to_host": (C.c_int, [_P, _P, _P, _I64]),
    "dex_host_alloc": (C.c_int, [C.POINTER(_P), _I64]),
    "dex_host_free": (C.c_int, [_P]),
    "dex_device_alloc": (C.c_int, [_P, C.POINTER(_P), _I64]),
    "dex_device_free": (C.c_int, [_P, _P]),
    "dex_ipc_export": (C.c_int, [_P, _P, _P]),
    "dex_ipc_open": (C.c_int, [_P, _P, C.POINTER(_P)]),
    "dex_ipc_close": (C.c_int, [_P, _P]),
    "dex_ctx_launch_count": (_I64, [_P]),
    "dex_eval_launch_info": (C.c_int, [_P, C.c_int32, _I64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32),
                                       C.POINTER(_I64), C.POINTER(_I64), C.POINTER(C.c_int32)]),
    "dex_population_copy_tape": (_I64, [_P, _P, _I64, _P]),
    "dex_population_copy_folded": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "dex_handler_name": (C.c_char_p, [C.c_int]),
}


class DexError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libdexb200 error {code}: {msg}")
        self.code = code


class _Info(C.Structure):
    _fields_ = [("n_trees", _I64), ("n_nodes", _I64), ("n_instructions", _I64),
                ("n_constants", _I64), ("max_stack", _I32), ("max_feature", _I32),
                ("max_parameter", _I32), ("dtype", _I32), ("n_generic", _I64), ("n_checks", _I64),
                ("n_folded_instructions", _I64), ("n_scalar_instructions", _I64),
                ("n_folded_subtrees", _I64), ("folded_max_stack", _I32), ("reserved0", _I32)]


_lib = None
_lib_lock = threading.Lock()


def lib():
    """Load libdexb200.so (built by ``__graft_entry__.build()`` / ``make -C csrc``)."""
    global _lib
    if _lib is None:
        with _lib_lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                        f"g.build()'` (needs nvcc).  dexb200 has no CPU fallback.")
                l = C.CDLL(LIB_PATH)
                for name, (res, args) in ABI.items():
                    fn = getattr(l, name)
                    fn.restype = res
                    fn.argtypes = args
                if l.dex_abi_version() != 1:
                    raise RuntimeError("libdexb200 ABI version mismatch")
                _lib = l
    return _lib


def _ptr(a):
    """Raw address of a numpy array / torch tensor / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


def _dtype_code(dt):
    import torch
    if dt in (np.float32, torch.float32, np.dtype(np.float32)):
        return F32
    if dt in (np.float64, torch.float64, np.dtype(np.float64)):
        return F64
    raise TypeError(f"the device path evaluates Float32/Float64 only, got {dt}")


def _torch_dtype(code):
    import torch
    return torch.float32 if code == F32 else torch.float64


class Context:
    """One ``dex_ctx`` per (host thread, device) — like the reference, nothing is shared
    between concurrently evaluating threads."""

    _tls = threading.local()

    def __init__(self, device=0):
        self.device = int(device)
        h = _P()
        rc = lib().dex_ctx_create(self.device, C.byref(h))
        if rc != OK:
            raise DexError(rc, lib().dex_strerror(rc).decode() +
                           f" (creating a context on device {device}; dexb200 needs a CUDA GPU)")
        self.h = h
        self._optables = {}

    @classmethod
    def get(cls, device=None):
        import torch
        if device is None:
            if not torch.cuda.is_available():
                raise DexError(-3, "no CUDA device is available and dexb200 has no CPU fallback")
            device = torch.cuda.current_device()
        if isinstance(device, torch.device):
            device = device.index if device.index is not None else torch.cuda.current_device()
        cache = cls._tls.__dict__.setdefault("ctxs", {})
        if device not in cache:
            cache[device] = Context(device)
        return cache[device]

    def check(self, rc):
        if rc != OK:
            raise DexError(rc, lib().dex_last_error(self.h).decode() or lib().dex_strerror(rc).decode())

    def use_current_stream(self):
        import torch
        s = torch.cuda.current_stream(self.device).cuda_stream
        self.check(lib().dex_ctx_set_stream(self.h, _P(s)))

    def synchronize(self):
        self.check(lib().dex_ctx_synchronize(self.h))

    @property
    def launch_count(self):
        return int(lib().dex_ctx_launch_count(self.h))

    def optable(self, operators: OperatorEnum):
        key = operators.key()
        t = self._optables.get(key)
        if t is None:
            flat, offs = operators.flat_opcodes()
            h = _P()
            rc = lib().dex_optable_create(_ptr(flat), _ptr(offs), MAX_DEGREE, C.byref(h))
            if rc != OK:
                raise DexError(rc, lib().dex_strerror(rc).decode())
            t = self._optables[key] = h
        return t

    def __del__(self):
        try:
            for t in self._optables.values():
                lib().dex_optable_destroy(t)
            lib().dex_ctx_destroy(self.h)
        except Exception:
            pass


class PeerBuffer:
    """A cudaMalloc'ed device buffer that other processes can map (CUDA IPC) — the landing zone of
    the fused evaluate + gather (include/dexb200.h "peer memory").  ``PeerBuffer(ctx, nbytes)``
    allocates and owns; ``PeerBuffer.open(ctx, handle)`` maps a peer's buffer."""

    def __init__(self, ctx: "Context", nbytes: int):
        self.ctx, self.nbytes, self.owner = ctx, int(nbytes), True
        p = _P()
        ctx.check(lib().dex_device_alloc(ctx.h, C.byref(p), self.nbytes))
        self.ptr = int(p.value)

    def handle(self) -> bytes:
        h = (C.c_uint8 * 64)()
        self.ctx.check(lib().dex_ipc_export(self.ctx.h, _P(self.ptr), C.cast(h, _P)))
        return bytes(h)

    @classmethod
    def open(cls, ctx: "Context", handle: bytes):
        self = cls.__new__(cls)
        self.ctx, self.nbytes, self.owner = ctx, None, False
        h = (C.c_uint8 * 64).from_buffer_copy(handle)
        p = _P()
        ctx.check(lib().dex_ipc_open(ctx.h, C.cast(h, _P), C.byref(p)))
        self.ptr = int(p.value)
        return self

    def as_tensor(self, shape, dtype):
        """torch view of an OWNED buffer (for reading the gathered result on the root)."""
        import torch

        class _Iface:
            pass

        holder = _Iface()
        typestr = {torch.float32: "<f4", torch.float64: "<f8", torch.uint8: "|u1"}[dtype]
        holder.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": typestr,
                                           "data": (self.ptr, False), "version": 2, "strides": None}
        t = torch.as_tensor(holder, device=f"cuda:{self.ctx.device}")
        t._dex_keepalive = self
        return t

    def close(self):
        if self.ptr:
            if self.owner:
                lib().dex_device_free(self.ctx.h, _P(self.ptr))
            else:
                lib().dex_ipc_close(self.ctx.h, _P(self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def host_context():
    """A context that can pack/validate but not evaluate (no CUDA calls)."""
    c = Context.__new__(Context)
    c.device = -1
    h = _P()
    rc = lib().dex_ctx_create(-1, C.byref(h))
    if rc != OK:
        raise DexError(rc, lib().dex_strerror(rc).decode())
    c.h = h
    c._optables = {}
    return c


def as_device_matrix(X, device, dtype_code=None):
    """Return (tensor with the memory of a column-major F x N matrix, F, N, ldx).

    ``X`` has the reference's shape (F, N).  A torch CUDA tensor whose strides are already
    (1, ldx) — e.g. ``Xt.T`` of a contiguous (N, F) tensor — is used in place; anything else
    is copied (numpy -> pinned -> device, or a device transpose copy)."""
    import torch
    if isinstance(X, np.ndarray):
        if X.ndim == 1:
            X = X.reshape(-1, 1)
        if dtype_code is not None:
            X = X.astype(np.float32 if dtype_code == F32 else np.float64, copy=False)
        Xt = torch.from_numpy(np.ascontiguousarray(X.T))
        Xd = Xt.to(f"cuda:{device}", non_blocking=False)
        F, N = X.shape
        return Xd, F, N, F
    if X.dim() == 1:
        X = X.reshape(-1, 1)
    if dtype_code is not None and X.dtype != _torch_dtype(dtype_code):
        X = X.to(_torch_dtype(dtype_code))
    F, N = X.shape
    if not X.is_cuda or X.device.index != device:
        X = X.to(f"cuda:{device}")
    if N > 1 and F > 0 and X.stride(0) == 1 and X.stride(1) >= F:
        return X, F, N, X.stride(1)
    Xc = X.T.contiguous()  # (N, F) row-major == (F, N) column-major
    return Xc, F, N, F


class Population:
    """A packed population of trees on one device (``dex_population``)."""

    def __init__(self, trees, operators: OperatorEnum, dtype=np.float32, *, ctx: Context = None,
                 bumper=False, use_fused=True, wire=None, n_params=0):
        self.ctx = ctx or Context.get()
        self.operators = operators
        self.dtype_code = _dtype_code(dtype)
        if wire is not None:
            nodes, offsets = wire
        else:
            if isinstance(trees, Node):
                trees = [trees]
            nodes, offsets = to_wire_population(list(trees))
        nodes = np.ascontiguousarray(nodes, dtype=WIRE_DTYPE)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.n_trees = len(offsets) - 1
        # n_params: size(parameters, 1) of a ParametricExpression (DEX_PACK_PARAM_ROWS): every
        # parameter gets its row and gradient direction even when the trees do not use all of them
        flags = (PACK_BUMPER if bumper else 0) | (PACK_FUSED if use_fused else 0) | ((int(n_params) & 0xffff) << 8)
        h = _P()
        self.ctx.check(lib().dex_population_pack(self.ctx.h, self.ctx.optable(operators), _ptr(nodes),
                                                 _ptr(offsets), self.n_trees, self.dtype_code, flags,
                                                 C.byref(h)))
        self.h = h
        info = _Info()
        lib().dex_population_get_info(self.h, C.byref(info))
        self.info = {k: getattr(info, k) for k, _ in _Info._fields_}
        self.n_nodes = self.info["n_nodes"]

    def launch_info(self, nfeatures, nsamples, *, early_exit=True, loss=False, parametric=False):
        """Launch geometry ``dex_eval`` / ``dex_eval_loss`` / ``dex_eval_parametric`` would choose (no device
        needed): threads per CTA, shared memory per CTA, sample tiles, rows kept in shared memory
        (0 = all; otherwise the remaining feature rows are read through L1 — dex_eval.cu eval_num_tiles)."""
        th, rows = C.c_int32(), C.c_int32()
        smem, tiles = _I64(), _I64()
        rc = lib().dex_eval_launch_info(self.h, int(nfeatures), int(nsamples), EVAL_EARLY_EXIT if early_exit else 0,
                                        int(bool(loss)), int(bool(parametric)), C.byref(th), C.byref(smem),
                                        C.byref(tiles), C.byref(rows))
        if rc != OK:
            raise ValueError("dex_eval_launch_info: invalid arguments")
        return {"threads": th.value, "smem_bytes": smem.value, "n_tiles": tiles.value, "smem_rows": rows.value}

    def __del__(self):
        try:
            lib().dex_population_destroy(self.h)
        except Exception:
            pass

    # -- constants -------------------------------------------------------------------
    def constant_counts(self):
        c = np.zeros(self.n_trees, dtype=np.int32)
        lib().dex_population_constant_counts(self.h, _ptr(c))
        return c

    def get_constants(self):
        v = np.zeros(self.info["n_constants"], dtype=np.float32 if self.dtype_code == F32 else np.float64)
        self.ctx.check(lib().dex_population_get_constants(self.ctx.h, self.h, _ptr(v), len(v)))
        return v

    def set_constants(self, values):
        v = np.ascontiguousarray(values, dtype=np.float32 if self.dtype_code == F32 else np.float64)
        self.ctx.check(lib().dex_population_set_constants(self.ctx.h, self.h, _ptr(v), len(v)))

    def tape(self):
        """Host copy of the evaluation tape: (uint32[n, 4], offsets[n_trees + 1])."""
        n = self.info["n_instructions"]
        ins = np.zeros((n, 4), dtype=np.uint32)
        off = np.zeros(self.n_trees + 1, dtype=np.int64)
        lib().dex_population_copy_tape(self.h, _ptr(ins), n, _ptr(off))
        return ins, off

    def folded(self):
        """Host copy of the image dex_eval* run: dict(tape, offsets, scalar_tape, segs, seg_offsets);
        segs[k] = (scalar begin, scalar end, target instruction in `tape`)."""
        nf, ns, nk = (self.info[k] for k in ("n_folded_instructions", "n_scalar_instructions", "n_folded_subtrees"))
        ins = np.zeros((nf, 4), dtype=np.uint32)
        off = np.zeros(self.n_trees + 1, dtype=np.int64)
        sc = np.zeros((ns, 4), dtype=np.uint32)
        segs = np.zeros((nk, 3), dtype=np.int64)
        soff = np.zeros(self.n_trees + 1, dtype=np.int64)
        lib().dex_population_copy_folded(self.h, _ptr(ins), _ptr(off), _ptr(sc), _ptr(segs), _ptr(soff))
        return dict(tape=ins, offsets=off, scalar_tape=sc, segs=segs, seg_offsets=soff)

    # -- evaluation ------------------------------------------------------------------
    def _outputs(self, N, out, ok):
        import torch
        dev = f"cuda:{self.ctx.device}"
        if out is None:
            out = torch.empty((self.n_trees, N), dtype=_torch_dtype(self.dtype_code), device=dev)
        if ok is None:
            ok = torch.empty(self.n_trees, dtype=torch.uint8, device=dev)
        assert out.is_cuda and out.stride(-1) == 1 and out.shape == (self.n_trees, N)
        return out, ok

    def eval(self, X, *, early_exit=True, out=None, ok=None):
        """Batched ``eval_tree_array``: returns (out[P, N], ok[P]) as CUDA tensors."""
        Xd, F, N, ldx = as_device_matrix(X, self.ctx.device, self.dtype_code)
        out, ok = self._outputs(N, out, ok)
        self.ctx.use_current_stream()
        self.ctx.check(lib().dex_eval(self.ctx.h, self.h, _ptr(Xd), F, N, ldx, _ptr(out),
                                      out.stride(0) if self.n_trees else N, _ptr(ok),
                                      EVAL_EARLY_EXIT if early_exit else 0))
        return out, ok

    def eval_into(self, X, out_ptr, ldo, *, ok=None, early_exit=True):
        """Batched ``eval_tree_array`` whose result rows are stored at a RAW device address with
        row stride ``ldo`` (elements) — e.g. a column block inside a peer GPU's gathered
        ``(P, N_total)`` matrix mapped with :class:`PeerBuffer` (fused evaluate + gather)."""
        import torch
        Xd, F, N, ldx = as_device_matrix(X, self.ctx.device, self.dtype_code)
        if ok is None:
            ok = torch.empty(self.n_trees, dtype=torch.uint8, device=f"cuda:{self.ctx.device}")
        self.ctx.use_current_stream()
        self.ctx.check(lib().dex_eval(self.ctx.h, self.h, _ptr(Xd), F, N, ldx, _P(int(out_ptr)), int(ldo),
                                      _ptr(ok), EVAL_EARLY_EXIT if early_exit else 0))
        return ok

    def eval_parametric(self, X, parameters, classes0, *, early_exit=True, out=None, ok=None):
        """``parameters``: (P, n_params, n_classes); ``classes0``: 0-based class per sample."""
        import torch
        Xd, F, N, ldx = as_device_matrix(X, self.ctx.device, self.dtype_code)
        dev = f"cuda:{self.ctx.device}"
        tdt = _torch_dtype(self.dtype_code)
        params = torch.as_tensor(parameters, dtype=tdt)
        if params.dim() == 2:
            params = params.unsqueeze(0)
        P, n_params, n_classes = params.shape
        assert P == self.n_trees
        pd = params.to(dev).permute(0, 2, 1).contiguous()  # per tree column-major (n_params x n_classes)
        cl = torch.as_tensor(classes0).to(device=dev, dtype=torch.int32).contiguous()
        assert cl.numel() == N
        if N and (int(cl.min()) < 0 or int(cl.max()) >= n_classes):
            raise DexError(-5, "class index out of range")
        out, ok = self._outputs(N, out, ok)
        self.ctx.use_current_stream()
        self.ctx.check(lib().dex_eval_parametric(self.ctx.h, self.h, _ptr(Xd), F, N, ldx, _ptr(pd),
                                                 n_params, n_classes, _ptr(cl), _ptr(out),
                                                 out.stride(0) if self.n_trees else N, _ptr(ok),
                                                 EVAL_EARLY_EXIT if early_exit else 0))
        return out, ok

    def eval_parametric_prepared(self, X, params_cm, classes_i32, n_params, n_classes, *, early_exit=True,
                                 out=None, ok=None):
        """As :meth:`eval_parametric` for callers that keep their arguments in the library's own
        layout on the device (no per-call conversion or range check): ``params_cm`` is the
        per-tree column-major (P, n_classes, n_params) block, ``classes_i32`` the 0-based int32
        class of every sample (the caller guarantees 0 <= class < n_classes)."""
        Xd, F, N, ldx = as_device_matrix(X, self.ctx.device, self.dtype_code)
        out, ok = self._outputs(N, out, ok)
        self.ctx.use_current_stream()
        self.ctx.check(lib().dex_eval_parametric(self.ctx.h, self.h, _ptr(Xd), F, N, ldx, _ptr(params_cm),
                                                 int(n_params), int(n_classes), _ptr(classes_i32), _ptr(out),
                                                 out.stride(0) if self.n_trees else N, _ptr(ok),
                                                 EVAL_EARLY_EXIT if early_exit else 0))
        return out, ok

    def grad_offsets(self, F, N, mode):
        off = np.zeros(self.n_trees + 1, dtype=np.int64)
        rc = lib().dex_grad_offsets(self.h, F, N, mode, _ptr(off))
        if rc != OK:
            raise DexError(rc, "dex_grad_offsets")
        return off

    def eval_grad(self, X, mode=GRAD_FEATURES):
        """Batched ``eval_grad_tree_array``: (out[P, N], grad_flat, offsets, ok[P]).
        Tree t's gradient is ``grad_flat[off[t]:off[t+1]].view(N, G_t).T`` (G_t x N)."""
        import torch
        Xd, F, N, ldx = as_device_matrix(X, self.ctx.device, self.dtype_code)
        out, ok = self._outputs(N, None, None)
        off = self.grad_offsets(F, N, mode)
        grad = torch.empty(int(off[-1]), dtype=out.dtype, device=out.device)
        self.ctx.use_current_stream()
        self.ctx.check(lib().dex_eval_grad(self.ctx.h, self.h, _ptr(Xd), F, N, ldx, mode, _ptr(out),
                                           out.stride(0) if self.n_trees else N, _ptr(grad),
                                           _ptr(off), _ptr(ok)))
        return out, grad, off, ok

    def eval_grad_parametric(self, X, parameters, classes0, mode=GRAD_FEATURES):
        """Batched ``eval_grad_tree_array`` of ParametricExpressions, as the reference differentiates
        them: the per-sample parameter rows ``parameters[:, classes]`` are the first ``n_params``
        feature directions, the rows of X follow, then the constants.  Returns
        (out[P, N], grad_flat, offsets, ok[P]); tree t's block is ``(N, G_t)`` row-major = the
        reference's ``(G_t, N)`` column-major matrix."""
        import torch
        Xd, F, N, ldx = as_device_matrix(X, self.ctx.device, self.dtype_code)
        dev = f"cuda:{self.ctx.device}"
        tdt = _torch_dtype(self.dtype_code)
        params = torch.as_tensor(parameters, dtype=tdt)
        if params.dim() == 2:
            params = params.unsqueeze(0)
        P, n_params, n_classes = params.shape
        assert P == self.n_trees
        pd = params.to(dev).permute(0, 2, 1).contiguous()
        cl = torch.as_tensor(classes0).to(device=dev, dtype=torch.int32).contiguous()
        assert cl.numel() == N
        if N and (int(cl.min()) < 0 or int(cl.max()) >= n_classes):
            raise DexError(-5, "class index out of range")
        out, ok = self._outputs(N, None, None)
        off = self.grad_offsets(n_params + F, N, mode)
        grad = torch.empty(int(off[-1]), dtype=out.dtype, device=out.device)
        self.ctx.use_current_stream()
        self.ctx.check(lib().dex_eval_grad_parametric(self.ctx.h, self.h, _ptr(Xd), F, N, ldx, _ptr(pd), n_params,
                                                      n_classes, _ptr(cl), mode, _ptr(out),
                                                      out.stride(0) if self.n_trees else N, _ptr(grad), _ptr(off),
                                                      _ptr(ok)))
        return out, grad, off, ok

    def eval_diff(self, X, direction0):
        import torch
        Xd, F, N, ldx = as_device_matrix(X, self.ctx.device, self.dtype_code)
        out, ok = self._outputs(N, None, None)
        dout = torch.empty_like(out)
        self.ctx.use_current_stream()
        self.ctx.check(lib().dex_eval_diff(self.ctx.h, self.h, _ptr(Xd), F, N, ldx, int(direction0),
                                           _ptr(out), _ptr(dout), out.stride(0) if self.n_trees else N,
                                           _ptr(ok)))
        return out, dout, ok

    def eval_loss(self, X, y, *, weights=None, early_exit=True):
        """Fused (weighted) mean-squared-error per tree without materialising the results:
        (loss[P] float64, ok[P])."""
        import torch
        Xd, F, N, ldx = as_device_matrix(X, self.ctx.device, self.dtype_code)
        dev = f"cuda:{self.ctx.device}"
        yd = torch.as_tensor(y).to(device=dev, dtype=_torch_dtype(self.dtype_code)).contiguous()
        wd = None if weights is None else torch.as_tensor(weights).to(device=dev, dtype=_torch_dtype(self.dtype_code)).contiguous()
        assert yd.numel() == N and (wd is None or wd.numel() == N)
        loss = torch.empty(self.n_trees, dtype=torch.float64, device=dev)
        ok = torch.empty(self.n_trees, dtype=torch.uint8, device=dev)
        self.ctx.use_current_stream()
        self.ctx.check(lib().dex_eval_loss(self.ctx.h, self.h, _ptr(Xd), F, N, ldx, _ptr(yd), _ptr(wd),
                                           _ptr(loss), _ptr(ok), EVAL_EARLY_EXIT if early_exit else 0))
        return loss, ok

    def eval_loss_grad(self, X, y, mode=GRAD_CONSTANTS, *, weights=None):
        """Fused (weighted) mean-squared error per tree AND its gradient w.r.t. the constants
        and/or features, without materialising values or (G x N) gradients:
        (loss[P] float64, grad[sum G] float64, offsets[P + 1], ok[P])."""
        import torch
        Xd, F, N, ldx = as_device_matrix(X, self.ctx.device, self.dtype_code)
        dev = f"cuda:{self.ctx.device}"
        tdt = _torch_dtype(self.dtype_code)
        yd = torch.as_tensor(y).to(device=dev, dtype=tdt).contiguous()
        wd = None if weights is None else torch.as_tensor(weights).to(device=dev, dtype=tdt).contiguous()
        assert yd.numel() == N and (wd is None or wd.numel() == N)
        off = self.grad_offsets(F, 1, mode)
        loss = torch.empty(self.n_trees, dtype=torch.float64, device=dev)
        grad = torch.empty(max(int(off[-1]), 1), dtype=torch.float64, device=dev)
        ok = torch.empty(self.n_trees, dtype=torch.uint8, device=dev)
        self.ctx.use_current_stream()
        self.ctx.check(lib().dex_eval_loss_grad(self.ctx.h, self.h, _ptr(Xd), F, N, ldx, _ptr(yd), _ptr(wd), mode,
                                                _ptr(loss), _ptr(grad), _ptr(off), _ptr(ok)))
        return loss, grad[: int(off[-1])], off, ok

    def eval_host(self, X_host, out_host, ok_host, *, early_exit=True, skip_incomplete=False):
        """The host-buffer entry point (dex_eval_host): numpy / pinned torch CPU tensors in the
        library's layouts — ``X_host`` is (N, F) row-major, ``out_host`` (P, N), ``ok_host`` (P,).
        ``skip_incomplete``: rows of trees whose flag comes back 0 are not transferred
        (DEX_EVAL_SKIP_INCOMPLETE; they are unspecified under early exit anyway)."""
        N, F = X_host.shape
        flags = (EVAL_EARLY_EXIT if early_exit else 0) | (EVAL_SKIP_INCOMPLETE if skip_incomplete else 0)
        self.ctx.check(lib().dex_eval_host(self.ctx.h, self.h, _ptr(X_host), F, N, F, _ptr(out_host),
                                           N, _ptr(ok_host), flags))
        return out_host, ok_host


def shard_eval_host(pops, X_host, out_host, ok_host, *, early_exit=True):
    """``dex_shard_eval_host``: one process, one :class:`Population` (same trees) per device; device d
    evaluates its column block of ``X_host`` ((N, F) row-major) and the rows land in ``out_host``."""
    N, F = X_host.shape
    n = len(pops)
    ctxs = (_P * n)(*[p.ctx.h for p in pops])
    hs = (_P * n)(*[p.h for p in pops])
    pops[0].ctx.check(lib().dex_shard_eval_host(ctxs, hs, n, _ptr(X_host), F, N, F, _ptr(out_host), N, _ptr(ok_host),
                                                EVAL_EARLY_EXIT if early_exit else 0))
    return out_host, ok_host


def shard_eval(pops, X_blocks, out_root, ok_root, *, root=0, early_exit=True):
    """``dex_shard_eval``: one process, one :class:`Population` (same trees) per device.  ``X_blocks[d]`` is
    device d's column block of X as a CUDA tensor of shape (F, n_d) with column-major memory (``Xt.T`` of a
    contiguous (n_d, F) tensor); the (P, N) result and the flags are CUDA tensors on ``pops[root]``'s
    device, written by every device through peer memory.  Asynchronous in the root context's stream."""
    n = len(pops)
    F = X_blocks[0].shape[0]
    N = sum(int(x.shape[1]) for x in X_blocks)
    for d, x in enumerate(X_blocks):
        s, e = N * d // n, N * (d + 1) // n
        assert x.is_cuda and x.device.index == pops[d].ctx.device and x.shape == (F, e - s)
        assert x.shape[1] <= 1 or (x.stride(0) == 1 and x.stride(1) >= F), "column-major (F, n) blocks"
    ldx = max([int(x.stride(1)) for x in X_blocks if x.shape[1] > 1] + [F])
    assert all(x.shape[1] <= 1 or x.stride(1) == ldx for x in X_blocks)
    assert out_root.is_cuda and out_root.device.index == pops[root].ctx.device and out_root.stride(1) == 1
    ctxs = (_P * n)(*[p.ctx.h for p in pops])
    hs = (_P * n)(*[p.h for p in pops])
    xs = (_P * n)(*[_ptr(x) for x in X_blocks])
    for p in pops:
        p.ctx.use_current_stream()
    pops[root].ctx.check(lib().dex_shard_eval(ctxs, hs, n, xs, F, N, ldx, _ptr(out_root), out_root.stride(0), _ptr(ok_root),
                                              root, EVAL_EARLY_EXIT if early_exit else 0))
    return out_root, ok_root
