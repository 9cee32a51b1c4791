"""ctypes binding of libbmkg_b200.so (include/bmkg_b200.h).

The product path has no CPU fallback: if the library is missing this module
raises at import time, and every wrapper raises on a non-zero return code.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libbmkg_b200.so")


class BmkgError(RuntimeError):
    pass


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: biomedkg_b200 has no CPU fallback. Build the sm_100a kernel "
            "library first: `python biomedkg_b200/build.py` (or `python -c 'import __graft_entry__ as g; g.build()'`)."
        )
    return ctypes.CDLL(LIB_PATH)


lib = _load()

P, I64, I, F, SZ, U64 = c_void_p, c_int64, c_int, c_float, c_size_t, c_uint64

# name -> (restype, argtypes); mirrors include/bmkg_b200.h one to one
SIGNATURES = {
    "bmkg_abi_version": (I, []),
    "bmkg_error_string": (c_char_p, [I]),
    "bmkg_last_driver_status": (I, []),
    "bmkg_bind_device": (I, [I]),
    "bmkg_edge_sort_workspace_bytes": (SZ, [I64, I64]),
    "bmkg_edge_sort": (I, [P, I64, I64, I, P, P, P, P, P, P, SZ, P]),
    "bmkg_csr_filter_workspace_bytes": (SZ, [I64, I64]),
    "bmkg_csr_filter": (I, [P, P, P, P, P, P, P, I64, I64, P, P, P, P, P, P, SZ, P]),
    "bmkg_gcn_aggregate": (I, [P, P, P, P, I64, I, P, I, F, U64, P, P, I, P]),
    "bmkg_mask_cast": (I, [P, P, P, I64, P, P, P, P]),
    "bmkg_modality_mean": (I, [P, I64, I, I, P, P, P]),
    "bmkg_colsum_workspace_bytes": (SZ, [I64, I]),
    "bmkg_relu_dropout_bwd": (I, [P, P, F, I64, I, P, P, P, SZ, P]),
    "bmkg_colsum": (I, [P, P, I64, I, P, P, SZ, P]),
    "bmkg_l2norm_scale": (I, [P, I64, I, F, P, P, P]),
    "bmkg_l2norm_scale_bwd": (I, [P, P, P, I64, I, F, P, P]),
    "bmkg_colmean_sigmoid": (I, [P, I64, I, P, P, SZ, P]),
    "bmkg_rowdot": (I, [P, P, I64, I, P, P]),
    "bmkg_rowdot_bwd": (I, [P, P, I64, I, P, P]),
    "bmkg_softplus_pair_workspace_bytes": (SZ, [I64]),
    "bmkg_softplus_pair_sum": (I, [P, P, I64, P, P, SZ, P]),
    "bmkg_softplus_pair_bwd": (I, [P, P, P, I64, P, P, P]),
    "bmkg_fusion_attn_fwd": (I, [P, I64, I, I, P, P, P]),
    "bmkg_fusion_attn_bwd": (I, [P, P, P, I64, I, I, P, P]),
    "bmkg_infonce_padded_rows": (I64, [I64]),
    "bmkg_infonce_workspace_bytes": (SZ, [I64, I]),
    "bmkg_infonce_fwd": (I, [P, I64, I, P, P, P, SZ, P]),
    "bmkg_infonce_bwd": (I, [P, P, P, I64, I, P, P]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header/library mismatch
    _fn.restype = _res
    _fn.argtypes = _args

if lib.bmkg_abi_version() != 1:
    raise ImportError("libbmkg_b200.so ABI version mismatch; rebuild with `python biomedkg_b200/build.py --force`")

#: number of kernel-launching C-ABI calls made so far (bench.py reports it as gpu_launches evidence)
call_count = 0


def check(rc: int, what: str) -> None:
    if rc != 0:
        extra = f", driver status {lib.bmkg_last_driver_status()}" if rc == -6 else ""
        raise BmkgError(f"{what} failed: {lib.bmkg_error_string(rc).decode()} (code {rc}{extra})")


_tls = threading.local()


def bind_thread(device_index: int) -> None:
    """Bind this host thread (main thread or autograd worker) to the tensors' device inside the library's runtime."""
    if getattr(_tls, "device", None) != device_index:
        check(lib.bmkg_bind_device(int(device_index)), "bmkg_bind_device")
        _tls.device = device_index


def call(name: str, *args) -> None:
    """Launch one C-ABI entry point on the calling thread's current torch device."""
    global call_count
    call_count += 1
    bind_thread(torch.cuda.current_device())
    check(getattr(lib, name)(*args), name)
