"""ctypes binding of liblivecc_sm100a.so (include/livecc_b200.h).

There is no fallback: if the library is missing or the device is not sm_100, this raises.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "liblivecc_sm100a.so"
HEADER_PATH = PKG_DIR.parent / "include" / "livecc_b200.h"

ABI_VERSION = 1

# epilogue codes (LCC_EPI_*)
EPI_NONE, EPI_BIAS, EPI_BIAS_QUICKGELU, EPI_BIAS_GELU, EPI_RESIDUAL, EPI_BIAS_RESIDUAL, EPI_SWIGLU = range(7)


class LiveCCNativeError(RuntimeError):
    pass


_lib = None


def header_symbols() -> list[str]:
    """Names of every function declared in include/livecc_b200.h (used by the CPU export test)."""
    text = HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lcc_[a-z0-9_]+)\s*\(", text)))


def load_library(build_if_missing: bool = True) -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if not build_if_missing:
            raise LiveCCNativeError(f"{LIB_PATH} is missing (run `python -m livecc_b200.build`)")
        from . import build as _build

        _build.build()
    lib = C.CDLL(str(LIB_PATH))
    lib.lcc_abi_version.restype = C.c_int
    if lib.lcc_abi_version() != ABI_VERSION:
        raise LiveCCNativeError("liblivecc_sm100a.so ABI version mismatch; rebuild it")
    lib.lcc_create.restype = C.c_void_p
    lib.lcc_create.argtypes = [C.c_int]
    lib.lcc_destroy.argtypes = [C.c_void_p]
    lib.lcc_last_error.restype = C.c_char_p
    lib.lcc_last_error.argtypes = [C.c_void_p]
    lib.lcc_num_sms.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def _ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


class Context:
    """One lcc_ctx bound to one CUDA device. Thin argument marshalling only."""

    def __init__(self, device_index: int):
        self.lib = load_library()
        self.handle = self.lib.lcc_create(int(device_index))
        if not self.handle:
            raise LiveCCNativeError(
                f"lcc_create({device_index}) failed: an sm_100 (B200) device is required; no fallback path exists"
            )
        self.device_index = device_index
        self.num_sms = self.lib.lcc_num_sms(C.c_void_p(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.lcc_destroy(C.c_void_p(self.handle))
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int, what: str):
        if rc != 0:
            msg = self.lib.lcc_last_error(C.c_void_p(self.handle)).decode()
            raise LiveCCNativeError(f"{what} failed ({rc}): {msg}")

    def call(self, name: str, *args):
        fn = getattr(self.lib, name)
        rc = fn(C.c_void_p(self.handle), *args)
        self.check(rc, name)

    @staticmethod
    def stream_ptr():
        import torch

        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    # ---- ops -------------------------------------------------------------------------------
    def gemm(self, a, b, out=None, bias=None, residual=None, epilogue=EPI_NONE, block_n=0):
        """out[M,N] = epilogue(a[M,K] @ b[N,K]^T); a/b/out are 2-D bf16 CUDA tensors (last dim contiguous)."""
        import torch

        M, K = a.shape
        N = b.shape[0]
        assert b.shape[1] == K and a.stride(1) == 1 and b.stride(1) == 1
        n_out = N // 2 if epilogue == EPI_SWIGLU else N
        if out is None:
            out = torch.empty((M, n_out), dtype=torch.bfloat16, device=a.device)
        assert out.stride(1) == 1 and out.shape == (M, n_out)
        ldr = residual.stride(0) if residual is not None else 0
        self.call(
            "lcc_gemm_bf16", _ptr(a), C.c_int(a.stride(0)), _ptr(b), C.c_int(b.stride(0)), _ptr(out),
            C.c_int(out.stride(0)), C.c_int(M), C.c_int(N), C.c_int(K), _ptr(bias), _ptr(residual),
            C.c_int(ldr), C.c_int(epilogue), C.c_int(block_n), self.stream_ptr(),
        )
        return out
