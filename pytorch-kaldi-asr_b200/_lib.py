"""ctypes binding of libpka_b200.so (the C ABI declared in include/pka_b200.h).

There is deliberately NO fallback: if the shared object is missing, or a call returns a non-zero status, a
RuntimeError is raised.  PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpka_b200.so")

PKA_F32, PKA_BF16 = 0, 1
MAX_CTX = 8


class Dropout(C.Structure):
    _fields_ = [("p", C.c_float), ("site", C.c_uint32), ("seed", C.c_uint64), ("step_ptr", C.c_void_p)]


class GemmDesc(C.Structure):
    _fields_ = [("A", C.c_void_p), ("B", C.c_void_p), ("C", C.c_void_p), ("bias", C.c_void_p), ("residual", C.c_void_p),
                ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("nseg", C.c_int32), ("nbatch", C.c_int32),
                ("lda", C.c_int32), ("ldb", C.c_int32), ("ldc", C.c_int32), ("ldr", C.c_int32),
                ("transA", C.c_int32), ("transB", C.c_int32),
                ("a_seg_off", C.c_int64), ("b_seg_off", C.c_int64),
                ("a_batch_off", C.c_int64), ("b_batch_off", C.c_int64), ("c_batch_off", C.c_int64),
                ("shiftA", C.c_int32 * MAX_CTX), ("shiftB", C.c_int32 * MAX_CTX),
                ("T", C.c_int32), ("relu", C.c_int32), ("accumulate", C.c_int32), ("splitk", C.c_int32),
                ("drop", Dropout), ("splitk_ws", C.c_void_p)]


class TcDesc(C.Structure):
    _fields_ = [("A", C.c_void_p), ("B", C.c_void_p), ("C", C.c_void_p), ("Ct", C.c_void_p), ("bias", C.c_void_p),
                ("mode", C.c_int32), ("Bt", C.c_int32), ("T", C.c_int32), ("Tp", C.c_int32),
                ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("nseg", C.c_int32),
                ("lda", C.c_int32), ("ldb", C.c_int32), ("ldc", C.c_int32),
                ("a_seg_col", C.c_int32), ("b_seg_col", C.c_int32),
                ("shift", C.c_int32 * MAX_CTX),
                ("relu", C.c_int32), ("c_dtype", C.c_int32), ("splits", C.c_int32), ("reserved", C.c_int32),
                ("drop", Dropout), ("addend", C.c_void_p), ("ldadd", C.c_int32), ("reserved2", C.c_int32)]


class AttnDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("H", C.c_int32), ("Lq", C.c_int32), ("Lk", C.c_int32), ("dk", C.c_int32),
                ("dv", C.c_int32), ("ldq", C.c_int32), ("ldk", C.c_int32), ("ldv", C.c_int32), ("ldo", C.c_int32),
                ("use_band", C.c_int32), ("band_start", C.c_int32), ("band_end", C.c_int32), ("scale", C.c_float),
                ("drop", Dropout)]


class ReduceJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("n", C.c_int64), ("split_stride", C.c_int64),
                ("splits", C.c_int32), ("kind", C.c_int32), ("accumulate", C.c_int32), ("D", C.c_int32),
                ("dk", C.c_int32), ("reserved", C.c_int32)]


class RelayoutJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("wf", C.c_void_p), ("wd", C.c_void_p),
                ("N", C.c_int32), ("K", C.c_int32), ("nseg", C.c_int32), ("kind", C.c_int32),
                ("ldf", C.c_int32), ("ldd", C.c_int32), ("n0", C.c_int32), ("reserved", C.c_int32)]


REDUCE_PLAIN, REDUCE_HEADS = 0, 1
RELAYOUT_PLAIN, RELAYOUT_HEADS = 0, 1


class BeamDesc(C.Structure):
    _fields_ = [("n_utt", C.c_int32), ("beam", C.c_int32), ("V", C.c_int32), ("max_edges", C.c_int32),
                ("max_len", C.c_int32), ("eos", C.c_int32), ("force_full_length", C.c_int32),
                ("inputs_are_logprobs", C.c_int32)]


_lib = None


def lib() -> C.CDLL:
    """Load (once) and return the kernel library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libpka_b200.so is missing (%s). Build it with `python pytorch-kaldi-asr_b200/build.py` -- there is no "
            "CPU or eager fallback for the sm_100a kernels." % LIB_PATH)
    _lib = C.CDLL(LIB_PATH)
    _lib.pka_last_error.restype = C.c_char_p
    _lib.pka_launch_count.restype = C.c_uint64
    return _lib


def launch_count() -> int:
    return int(lib().pka_launch_count())


def check(status: int, what: str = ""):
    if status != 0:
        msg = lib().pka_last_error().decode("utf-8", "replace")
        raise RuntimeError("libpka_b200 %s failed (status %d): %s" % (what, status, msg))


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return PKA_F32
    if t.dtype == torch.bfloat16:
        return PKA_BF16
    raise RuntimeError("libpka_b200: unsupported activation dtype %s" % t.dtype)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("libpka_b200 ops run on CUDA (sm_100a) tensors only; got a %s tensor. "
                               "There is no CPU fallback." % t.device)


def make_dropout(p: float, site: int, seed: int, step_tensor) -> Dropout:
    d = Dropout()
    d.p = float(p)
    d.site = int(site) & 0xFFFFFFFF
    d.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    d.step_ptr = step_tensor.data_ptr() if step_tensor is not None else 0
    return d


NO_DROPOUT = make_dropout(0.0, 0, 0, None)
