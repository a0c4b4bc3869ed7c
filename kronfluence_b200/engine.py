"""ctypes binding of libkfb.so — the C ABI declared in include/kfb.h.

This module is the ONLY place where Python talks to the CUDA library.  It passes raw device pointers
(`tensor.data_ptr()`), sizes and the current CUDA stream; no torch types cross the boundary.  There is
no CPU implementation behind it: if the library is missing or no sm_100 device is visible, the calls
raise.
"""

import ctypes
import os
from typing import Optional, Tuple

import torch

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libkfb.so")

KFB_OK = 0
KFB_ERR_INVALID = -1
KFB_ERR_CUDA = -2
KFB_ERR_OOM = -3
KFB_ERR_NO_DEVICE = -4
KFB_ERR_WORKSPACE = -5
KFB_ERR_NOT_CONVERGED = -6

KFB_F32, KFB_BF16, KFB_F16, KFB_F64 = 0, 1, 2, 3
PREC_FP32, PREC_BF16, PREC_STRICT = 0, 1, 2
LINEAR, CONV2D = 0, 1
PRECOND_IDENTITY, PRECOND_DIAGONAL, PRECOND_EIGEN = 0, 1, 2
EPI_STORE, EPI_ROWDOT, EPI_SQACC = 0, 1, 2

_DTYPE_CODES = {torch.float32: KFB_F32, torch.bfloat16: KFB_BF16, torch.float16: KFB_F16, torch.float64: KFB_F64}


class KfbLayer(ctypes.Structure):
    _fields_ = [(name, ctypes.c_int32) for name in (
        "kind", "d_in", "d_out", "has_bias", "c_in", "h_in", "w_in", "groups", "k_h", "k_w",
        "stride_h", "stride_w", "pad_h", "pad_w", "dil_h", "dil_w", "h_out", "w_out")]


class KfbSplit(ctypes.Structure):
    _fields_ = [("hi", ctypes.c_void_p), ("lo", ctypes.c_void_p), ("absmax", ctypes.c_void_p), ("rows", ctypes.c_int64),
                ("cols", ctypes.c_int64), ("ld", ctypes.c_int64), ("batch", ctypes.c_int64),
                ("batch_stride", ctypes.c_int64)]


class KfbEpilogue(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("out_f32", ctypes.c_void_p), ("ldo", ctypes.c_int64),
                ("out_batch_stride", ctypes.c_int64), ("out_split", KfbSplit), ("mul", ctypes.c_void_p),
                ("ldmul", ctypes.c_int64), ("transpose_out", ctypes.c_int32), ("square", ctypes.c_int32),
                ("accumulate", ctypes.c_int32), ("alpha", ctypes.c_float), ("g", ctypes.c_void_p),
                ("ldg", ctypes.c_int64), ("reduce_sq", ctypes.c_int32), ("row_group", ctypes.c_int32),
                ("g_batch_stride", ctypes.c_int64), ("symmetric", ctypes.c_int32), ("col_group", ctypes.c_int64)]


# Every symbol include/kfb.h declares, with its ctypes signature (restype, argtypes).
_vp, _i32, _i64, _sz, _f32, _f64 = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t,
                                    ctypes.c_float, ctypes.c_double)
_LP, _SP, _EP = ctypes.POINTER(KfbLayer), ctypes.POINTER(KfbSplit), ctypes.POINTER(KfbEpilogue)
SIGNATURES = {
    "kfb_version": (ctypes.c_int, []),
    "kfb_last_error": (ctypes.c_char_p, []),
    "kfb_struct_sizes": (None, [ctypes.POINTER(ctypes.c_int)] * 3),
    "kfb_device_info": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)] * 3),
    "kfb_set_gemm_backend": (ctypes.c_int, [ctypes.c_int]),
    "kfb_set_cta_pairs": (ctypes.c_int, [ctypes.c_int]),
    "kfb_set_tma_store": (ctypes.c_int, [ctypes.c_int]),
    "kfb_set_multicast": (ctypes.c_int, [ctypes.c_int]),
    "kfb_set_strict_pass_k": (ctypes.c_int, [ctypes.c_int]),
    "kfb_set_wide_regacc": (ctypes.c_int, [ctypes.c_int]),
    "kfb_set_idle_fill": (ctypes.c_int, [ctypes.c_int]),
    "kfb_set_group_sync": (ctypes.c_int, [ctypes.c_int]),
    "kfb_launch_count": (_i64, []),
    "kfb_split_gather": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.POINTER(_i64), _vp, _SP, ctypes.c_int, _vp]),
    "kfb_split_im2col": (ctypes.c_int, [_LP, _vp, ctypes.c_int, _i64, _i32, _SP, ctypes.c_int, _vp]),
    "kfb_gemm_nt": (ctypes.c_int, [_SP, _SP, _EP, ctypes.c_int, _vp]),
    "kfb_cov_workspace_bytes": (_sz, [_LP, _i64, _i64]),
    "kfb_cov_accum_activation": (ctypes.c_int, [_LP, _vp, ctypes.c_int, _i64, _i64, _vp, _vp, _vp, _sz, ctypes.c_int, _vp]),
    "kfb_cov_accum_gradient": (ctypes.c_int, [_LP, _vp, ctypes.c_int, _i64, _i64, _f32, _vp, _vp, _sz, ctypes.c_int, _vp]),
    "kfb_eigh_jacobi_max_dim": (ctypes.c_int, []),
    "kfb_eigh_workspace_bytes": (_sz, [_i32]),
    "kfb_eigh_sym": (ctypes.c_int, [_vp, _f64, _i32, _vp, _vp, _vp, _sz, _vp]),
    "kfb_eigh_last_sweeps": (ctypes.c_int, [_vp, _i32]),
    "kfb_eigh_status": (ctypes.c_int, [_vp]),
    "kfb_set_cusolver_path": (ctypes.c_int, [ctypes.c_char_p]),
    "kfb_eigen_operands": (ctypes.c_int, [_vp, _i32, _SP, _SP, ctypes.c_int, _vp]),
    "kfb_lambda_workspace_bytes": (_sz, [_LP, _i64, _i64]),
    "kfb_lambda_accum": (ctypes.c_int, [_LP, _vp, ctypes.c_int, _vp, ctypes.c_int, _i64, _i64, _i32, _SP, _SP, _f32, _vp, _vp, _sz, ctypes.c_int, _vp]),
    "kfb_lambda_invert": (ctypes.c_int, [_vp, _i64, _f64, _f64, _vp, _vp, _sz, _vp]),
    "kfb_precondition_workspace_bytes": (_sz, [_LP, _i64, _i64]),
    "kfb_precondition": (ctypes.c_int, [_LP, _vp, ctypes.c_int, _vp, ctypes.c_int, _i64, _i64, _i32, _SP, _SP, _SP, _SP, _vp, _f32, _SP, _i64, _vp, _vp, _sz, ctypes.c_int, _vp]),
    "kfb_pairwise_workspace_bytes": (_sz, [_LP, _i64, _i64]),
    "kfb_pairwise_scores": (ctypes.c_int, [_LP, _SP, _i64, _vp, ctypes.c_int, _vp, ctypes.c_int, _i64, _i64, _i32, _SP, _SP, _f32, _vp, _i64, _i64, _i32, _vp, _sz, ctypes.c_int, _vp]),
    "kfb_pairwise_operand_bytes": (_sz, [_LP, _i64, _i64, ctypes.c_int]),
    "kfb_pairwise_prepare_workspace_bytes": (_sz, [_LP, _i64, _i64]),
    "kfb_pairwise_prepare": (ctypes.c_int, [_LP, _vp, ctypes.c_int, _vp, ctypes.c_int, _i64, _i64, _i32, _SP, _SP, _vp, _sz,
                                            _vp, _sz, ctypes.c_int, _vp]),
    "kfb_pairwise_prepared_workspace_bytes": (_sz, [_LP, _i64, _i64]),
    "kfb_pairwise_scores_prepared": (ctypes.c_int, [_LP, _SP, _i64, _vp, _sz, _i64, _i64, _f32, _vp, _i64, _i64, _i32, _vp,
                                                    _sz, ctypes.c_int, _vp]),
    "kfb_self_workspace_bytes": (_sz, [_LP, _i64, _i64]),
    "kfb_self_scores": (ctypes.c_int, [_LP, _vp, ctypes.c_int, _vp, ctypes.c_int, _i64, _i64, _i32, _SP, _SP, _vp, _f32, _vp, _i64, _i32, _vp, _sz, ctypes.c_int, _vp]),
    "kfb_pairwise_lowrank_workspace_bytes": (_sz, [_LP, _i64, _i64, _i64, _i64]),
    "kfb_pairwise_scores_lowrank": (ctypes.c_int, [_LP, _SP, _SP, _i64, _vp, ctypes.c_int, _vp, ctypes.c_int, _i64, _i64,
                                                   _i32, _SP, _SP, _f32, _vp, _i64, _i64, _i32, _i32, _vp, _sz,
                                                   ctypes.c_int, _vp]),
    "kfb_aggregate_workspace_bytes": (_sz, [_LP, _i64, _i64]),
    "kfb_aggregate_gradient": (ctypes.c_int, [_LP, _vp, ctypes.c_int, _vp, ctypes.c_int, _i64, _i64, _SP, _SP, _vp, _f32, _vp,
                                              _vp, _sz, ctypes.c_int, _vp]),
    "kfb_pairwise_explicit_workspace_bytes": (_sz, [_LP, _i64]),
    "kfb_pairwise_scores_explicit": (ctypes.c_int, [_LP, _SP, _i64, _vp, _i64, _f32, _vp, _i64, _i64, _i32, _vp, _sz,
                                                    ctypes.c_int, _vp]),
    "kfb_per_sample_gradient_workspace_bytes": (_sz, [_LP, _i64, _i64]),
    "kfb_per_sample_gradient": (ctypes.c_int, [_LP, _vp, ctypes.c_int, _vp, ctypes.c_int, _i64, _i64, _f32, _vp, _vp, _sz,
                                               ctypes.c_int, _vp]),
    "kfb_transform_gradient_workspace_bytes": (_sz, [_LP, _i64]),
    "kfb_transform_gradient": (ctypes.c_int, [_LP, _vp, _i64, _SP, _SP, _vp, _f32, _vp, _SP, _i64, _vp, _sz, ctypes.c_int, _vp]),
    "kfb_sq_accum": (ctypes.c_int, [_vp, _i64, _i64, _f32, _vp, _vp]),
    "kfb_weighted_sqnorm": (ctypes.c_int, [_vp, _vp, _i64, _i64, _f32, _vp, _i32, _vp]),
    "kfb_pairwise_scores_host": (ctypes.c_int, [_LP, _SP, _i64, _vp, ctypes.c_int, _vp, ctypes.c_int, _i64, _i64, _i32, _SP, _SP, _f32, _vp, _vp, _vp, _vp, _vp, _sz, ctypes.c_int, _vp]),
}

_lib = None


class KfbError(RuntimeError):
    """A CUDA-side failure reported by libkfb."""


def library_path() -> str:
    return _LIB_PATH


def load_library() -> ctypes.CDLL:
    """Loads libkfb.so and types every exported symbol.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise KfbError(
            f"{_LIB_PATH} is missing: build the CUDA library with `python -m kronfluence_b200.build` "
            "(there is no CPU or PyTorch fallback for the EK-FAC hot path)."
        )
    lib = ctypes.CDLL(_LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the binding drift apart
        fn.restype = restype
        fn.argtypes = argtypes
    sizes = [ctypes.c_int() for _ in range(3)]
    lib.kfb_struct_sizes(*[ctypes.byref(x) for x in sizes])
    mine = (ctypes.sizeof(KfbLayer), ctypes.sizeof(KfbSplit), ctypes.sizeof(KfbEpilogue))
    if tuple(x.value for x in sizes) != mine:
        raise KfbError(f"ABI mismatch: libkfb structs are {[x.value for x in sizes]} bytes, the binding's {list(mine)}; "
                       "rebuild with `python -m kronfluence_b200.build`")
    _lib = lib
    # debugging switches of the GEMM engine (A/B measurements): KFB_MULTICAST=0, KFB_TMA_STORE=0, KFB_CTA_PAIRS=0
    if os.environ.get("KFB_IDLE_FILL", "").lstrip("-").isdigit():
        lib.kfb_set_idle_fill(int(os.environ["KFB_IDLE_FILL"]))
    for env, setter in (("KFB_MULTICAST", lib.kfb_set_multicast), ("KFB_TMA_STORE", lib.kfb_set_tma_store),
                        ("KFB_CTA_PAIRS", lib.kfb_set_cta_pairs), ("KFB_WIDE_REGACC", lib.kfb_set_wide_regacc),
                        ("KFB_GROUP_SYNC", lib.kfb_set_group_sync)):
        if os.environ.get(env, "").isdigit():
            setter(int(os.environ[env]))
    _register_cusolver(lib)
    return lib


def _register_cusolver(lib: ctypes.CDLL) -> None:
    """Tells libkfb where the cuSOLVER bundled with torch lives (used only above the Jacobi limit)."""
    try:
        import nvidia.cusolver  # type: ignore

        root = os.path.join(list(nvidia.cusolver.__path__)[0], "lib")
        for fname in sorted(os.listdir(root)):
            if fname.startswith("libcusolver.so"):
                lib.kfb_set_cusolver_path(os.path.join(root, fname).encode())
                return
    except Exception:  # pylint: disable=broad-exception-caught
        pass


def last_error() -> str:
    return load_library().kfb_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    """Maps a kfb_status to the exception kronfluence's callers expect (SURVEY.md §8b Errors)."""
    if rc == KFB_OK:
        return
    msg = last_error()
    if rc == KFB_ERR_INVALID:
        raise ValueError(f"libkfb: {msg}")
    if rc == KFB_ERR_OOM:
        # utils/dataset.py:66-101 halves the batch size on exactly this text.
        raise RuntimeError(f"CUDA out of memory. libkfb: {msg}")
    if rc == KFB_ERR_NO_DEVICE:
        raise KfbError(f"libkfb needs an sm_100 (B200) device and has no CPU path: {msg}")
    if rc == KFB_ERR_WORKSPACE:
        raise KfbError(f"libkfb workspace too small: {msg}")
    if rc == KFB_ERR_NOT_CONVERGED:
        raise KfbError(f"libkfb: {msg}")
    raise KfbError(f"libkfb failure ({rc}): {msg}")


def require_device() -> Tuple[int, int, int]:
    lib = load_library()
    sm, major, minor = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    check(lib.kfb_device_info(ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor)))
    return sm.value, major.value, minor.value


def dtype_code(dtype: torch.dtype) -> int:
    if dtype not in _DTYPE_CODES:
        raise ValueError(f"libkfb does not accept tensors of dtype {dtype}")
    return _DTYPE_CODES[dtype]


def stream_ptr(device: Optional[torch.device] = None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def strict_scale(absmax: float) -> float:
    """The power of two the strict operands are stored with (kfb_common.cuh strict_scale): absmax * scale in
    [2^13, 2^14)."""
    import math

    if not absmax > 0.0 or math.isinf(absmax):
        return 1.0
    _, exponent = math.frexp(absmax)
    return math.ldexp(1.0, 14 - exponent)


class Split:
    """Device storage for a batch of matrices in tensor-core operand layout: two bf16 planes (one for PREC_BF16),
    or, for PREC_STRICT, two FP16 planes of the scaled operand plus the device word holding its absmax."""

    def __init__(self, rows: int, cols: int, batch: int = 1, device=None, precision: int = PREC_FP32,
                 zero: bool = False):
        self.rows, self.cols, self.batch = int(rows), int(cols), int(batch)
        self.ld = round_up(max(self.cols, 1), 8)
        self.batch_stride = self.rows * self.ld
        self.precision = precision
        planes = {PREC_FP32: 2, PREC_BF16: 1, PREC_STRICT: 2}[precision]
        alloc = torch.zeros if zero else torch.empty
        self.storage = alloc((planes, self.batch, self.rows, self.ld), dtype=torch.bfloat16, device=device)
        self.absmax = torch.zeros(1, dtype=torch.float32, device=device) if precision == PREC_STRICT else None

    def struct(self, batch_offset: int = 0, batch: Optional[int] = None) -> KfbSplit:
        nb = self.batch - batch_offset if batch is None else batch
        off = batch_offset * self.batch_stride * 2
        hi = self.storage[0].data_ptr() + off
        lo = self.storage[1].data_ptr() + off if self.storage.shape[0] >= 2 else None
        absmax = self.absmax.data_ptr() if self.absmax is not None else None
        return KfbSplit(hi, lo, absmax, self.rows, self.cols, self.ld, nb, self.batch_stride)

    def to_float(self) -> torch.Tensor:
        """hi + lo as fp32 [batch, rows, cols] (tests / debugging); strict operands are unscaled in fp64 first."""
        if self.absmax is not None:
            planes = self.storage.view(torch.float16).double()
            val = (planes[0] + planes[1]) / strict_scale(float(self.absmax.item()))
            return val[:, :, : self.cols]
        val = self.storage[0].float()
        for plane in range(1, self.storage.shape[0]):
            val = val + self.storage[plane].float()
        return val[:, :, : self.cols]

    def nbytes(self) -> int:
        return self.storage.numel() * 2


class Workspace:
    """A grow-only device scratch buffer handed to libkfb (the library never allocates)."""

    def __init__(self, device):
        self.device = device
        self.buf: Optional[torch.Tensor] = None

    def get(self, nbytes: int) -> Tuple[int, int]:
        nbytes = max(int(nbytes), 256)
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = None
            self.buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self.buf.data_ptr(), self.buf.numel()

    def release(self) -> None:
        self.buf = None


def split_from_tensor(x: torch.Tensor, precision: int = PREC_FP32, ones_col: bool = False) -> Split:
    """Row-major [batch?, rows, cols] tensor -> Split (tests / generic GEMM use)."""
    lib = load_library()
    if x.dim() == 2:
        x = x.unsqueeze(0)
    x = x.contiguous()
    b, r, c = x.shape
    dst = Split(r, c + int(ones_col), b, device=x.device, precision=precision)
    desc = (ctypes.c_int64 * 9)(r * c, c, 0, 1, r, 1, c, 1 if ones_col else 0, 0)
    st = dst.struct()
    check(lib.kfb_split_gather(x.data_ptr(), dtype_code(x.dtype), desc, None, ctypes.byref(st), precision,
                               stream_ptr(x.device)))
    return dst


def gemm_nt(a: Split, b: Split, epilogue: KfbEpilogue, precision: int = PREC_FP32) -> None:
    lib = load_library()
    sa, sb = a.struct(), b.struct()
    check(lib.kfb_gemm_nt(ctypes.byref(sa), ctypes.byref(sb), ctypes.byref(epilogue), precision,
                          stream_ptr(a.storage.device)))
