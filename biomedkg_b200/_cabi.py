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
LIB_PATH = os.environ.get("BMKG_LIB_PATH") or os.path.join(_HERE, "_lib", "libbmkg_b200.so")   # env override: tuning builds


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
    "bmkg_hub_info_len": (I64, [I64, I64]),
    "bmkg_csr_filter": (I, [P, P, P, P, P, P, P, I64, I64, P, P, P, P, P, P, P, SZ, P]),
    "bmkg_gcn_aggregate_workspace_bytes": (SZ, [I64, I]),
    "bmkg_gcn_aggregate": (I, [P, P, P, P, I64, I, P, I, F, U64, P, P, I, I64, P, P, SZ, P]),
    "bmkg_gcn_aggregate_rows": (I, [P, P, P, P, I64, I64, I64, I, P, I, F, U64, P, P, I, I64, P, P, SZ, P]),
    "bmkg_gcn_star_aggregate": (I, [P, P, P, P, P, I64, I, P, I, P, I, P]),
    "bmkg_gat_scores": (I, [P, P, P, I64, I, I, P, P, P]),
    "bmkg_gat_workspace_bytes": (SZ, [I64, I, I]),
    "bmkg_gat_aggregate": (I, [P, P, P, P, P, I64, I, I, F, P, I, F, U64, P, P, I, P, P, I64, P, P, SZ, P]),
    "bmkg_gat_aggregate_bwd": (I, [P, P, P, P, P, P, P, P, P, P, P, P, I64, I, I, F, P, P, P, P, I64, P, P, P, SZ, P]),
    "bmkg_mask_cast": (I, [P, P, P, I64, P, P, P, P]),
    "bmkg_modality_mean": (I, [P, I64, I, I, P, P, P]),
    "bmkg_colsum_workspace_bytes": (SZ, [I64, I]),
    "bmkg_relu_dropout_bwd": (I, [P, P, F, I64, I, P, P, P, SZ, P]),
    "bmkg_colsum": (I, [P, P, I64, I, P, P, SZ, P]),
    "bmkg_center_cast": (I, [P, P, I64, I, P, P]),
    "bmkg_l2norm_colsum": (I, [P, I64, I, P, P, P, SZ, P]),
    "bmkg_center_scale": (I, [P, P, P, I64, I, F, P, P, P]),
    "bmkg_l2norm_scale_bwd": (I, [P, P, P, I64, I, F, P, P]),
    "bmkg_colmean_sigmoid": (I, [P, I64, I, P, P, SZ, P]),
    "bmkg_rowdot": (I, [P, P, I64, I, P, P]),
    "bmkg_rowdot_bwd": (I, [P, P, I64, I, P, P]),
    "bmkg_softplus_pair_workspace_bytes": (SZ, [I64]),
    "bmkg_softplus_pair_sum": (I, [P, P, I64, P, P, SZ, P]),
    "bmkg_softplus_pair_bwd": (I, [P, P, P, I64, P, P, P]),
    "bmkg_fusion_attn_fwd": (I, [P, P, I64, I, I, P, P, P]),
    "bmkg_fusion_attn_bwd": (I, [P, P, P, P, I64, I, I, P, P]),
    "bmkg_mask_cast_bwd": (I, [P, P, P, P, P, I64, P, P]),
    "bmkg_sample_workspace_bytes": (SZ, [I64]),
    "bmkg_sample_count": (I, [P, P, I64, I, P, P, SZ, P]),
    "bmkg_sample_pick": (I, [P, P, P, P, I64, I, P, U64, I, I64, P, P, P, P]),
    "bmkg_sample_relabel": (I, [P, I64, I64, P, P, P, P, P, P, SZ, P]),
    "bmkg_sample_set_ids": (I, [P, I64, P, I, P]),
    "bmkg_redaf_partial_rows": (I64, [I64, I]),
    "bmkg_redaf_fwd": (I, [P, P, P, I64, I, I, F, U64, P, P, P]),
    "bmkg_redaf_bwd": (I, [P, P, P, P, I64, I, I, F, U64, P, P, P, P]),
    "bmkg_colsum_bf16": (I, [P, P, P, I64, I, I, P, P, P, SZ, P]),
    "bmkg_linear_supported": (I, [I64, I, I]),
    "bmkg_linear_nt": (I, [P, P, P, I64, I, I, I, I, P, P, P, I, P, P, P]),
    "bmkg_linear_tn_workspace_bytes": (SZ, [I64, I, I]),
    "bmkg_linear_tn": (I, [P, P, P, I64, I, I, P, P, SZ, P]),
    "bmkg_infonce_stacked_rows": (I64, [I64, I64]),
    "bmkg_infonce_padded_rows": (I64, [I64, I64]),
    "bmkg_infonce_workspace_bytes": (SZ, [I64, I]),
    "bmkg_infonce_e_store_bytes": (SZ, [I64, I64, I64, I64]),
    "bmkg_infonce_ext": (I, [P, I64, I64, P, P]),
    "bmkg_infonce_fwd": (I, [P, P, P, I64, I, P, P, P, P, SZ, P]),
    "bmkg_infonce_bwd": (I, [P, P, P, P, P, I64, I, P, P, SZ, P]),
    "bmkg_infonce_bwd_workspace_bytes": (SZ, [I64, I64, I, I64, I64]),
    "bmkg_infonce_set_phase_bytes": (I64, [I64]),
    "bmkg_infonce_workspace_bytes_rows": (SZ, [I64, I64, I, I64, I64]),
    "bmkg_infonce_fwd_rows": (I, [P, P, P, I64, I64, I, I64, I64, P, P, P, P, SZ, P]),
    "bmkg_infonce_bwd_rows": (I, [P, P, P, P, P, I64, I64, I, I64, I64, P, P, SZ, P]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header/library mismatch
    _fn.restype = _res
    _fn.argtypes = _args

if lib.bmkg_abi_version() != 4:
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


#: CUDA kernels each entry point launches per call (for bench.py's gpu_launches count; the optional split-row hub
#: pre-passes of the aggregation kernels are not counted - a lower bound)
KERNELS_PER_CALL = {
    "bmkg_edge_sort": None,            # data dependent: 3 + 5 * passes + 2 (counted by formula in bench.py)
    "bmkg_csr_filter": 6, "bmkg_gcn_aggregate": 1, "bmkg_gcn_aggregate_rows": 1, "bmkg_gcn_star_aggregate": 1, "bmkg_mask_cast": 1, "bmkg_modality_mean": 1,
    "bmkg_relu_dropout_bwd": 2, "bmkg_colsum": 2, "bmkg_linear_nt": 1, "bmkg_linear_tn": 2, "bmkg_center_cast": 1, "bmkg_l2norm_colsum": 2, "bmkg_center_scale": 1, "bmkg_l2norm_scale_bwd": 1,
    "bmkg_colmean_sigmoid": 3, "bmkg_rowdot": 1, "bmkg_rowdot_bwd": 1, "bmkg_softplus_pair_sum": 2,
    "bmkg_softplus_pair_bwd": 1, "bmkg_fusion_attn_fwd": 1, "bmkg_fusion_attn_bwd": 1, "bmkg_infonce_ext": 1, "bmkg_infonce_fwd": 3,
    "bmkg_infonce_bwd": 1, "bmkg_infonce_fwd_rows": 3, "bmkg_infonce_bwd_rows": 1, "bmkg_gat_scores": 1, "bmkg_gat_aggregate": 1, "bmkg_gat_aggregate_bwd": 2, "bmkg_mask_cast_bwd": 1, "bmkg_colsum_bf16": 3,
    "bmkg_redaf_fwd": 1, "bmkg_redaf_bwd": 1, "bmkg_sample_count": 3, "bmkg_sample_pick": 1, "bmkg_sample_relabel": 6,
    "bmkg_sample_set_ids": 1,
}
kernel_launches = 0

#: optional per-entry-point device timing (bench.py): name -> list of (start_event, end_event) on the launch stream
timed_entries: set = set()
timings: dict = {}


def call(name: str, *args) -> None:
    """Launch one C-ABI entry point on the calling thread's current torch device."""
    global call_count, kernel_launches
    call_count += 1
    kernel_launches += KERNELS_PER_CALL.get(name) or 12
    bind_thread(torch.cuda.current_device())
    if name in timed_entries:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(getattr(lib, name)(*args), name)
        e1.record()
        timings.setdefault(name, []).append((e0, e1))
        return
    check(getattr(lib, name)(*args), name)
