"""ctypes binding of libmsx.so (the C-ABI boundary declared in include/msx.h).

Every entry point takes plain device pointers, sizes and a ``cudaStream_t``; tensors are passed as
``tensor.data_ptr()``.  A missing library or symbol is a hard error — there is no fallback path.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmsx.so")

_lib = None


class MsxError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MsxError("libmsx.so is not built (%s missing): run `python __graft_entry__.py` — "
                           "the product path has no CPU fallback" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.msx_last_error.restype = ctypes.c_char_p
    return _lib


def check(rc, what):
    if rc != 0:
        raise MsxError("%s failed (%d): %s" % (what, rc, load().msx_last_error().decode()))


def ptr(t):
    """Pointer of a torch tensor (or None) as void*."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


# kernels enqueued per C-ABI call (for the bench's gpu_launches claim)
_KERNELS_PER_CALL = {"msx_adam_step": 2, "msx_token_sort": 3}
LAUNCHES = 0
_profile = None     # optional {"match": substring, "events": [(start, end, tag)]} set by bench.py


def call(name, *args, tag=None):
    global LAUNCHES
    fn = getattr(load(), name)
    prof = _profile
    if prof is not None and prof["match"] in name:
        import torch
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        check(fn(*args), name)
        e.record()
        prof["events"].append((s, e, tag if tag is not None or not prof.get("names") else name))
    else:
        check(fn(*args), name)
    LAUNCHES += _KERNELS_PER_CALL.get(name, 1)
