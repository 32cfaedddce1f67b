"""ctypes binding of liboo_b200.so (C ABI declared in include/oo_b200.h).

There is no CPU fallback: if the shared library is missing, or no sm_100 device is visible when a
context is created, the product path raises."""
from __future__ import annotations

import ctypes as C
import glob
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboo_b200.so")

SYMBOLS = [
    "oo_last_error", "oo_version", "oo_device_count", "oo_create", "oo_destroy", "oo_set_stream",
    "oo_synchronize", "oo_set_integrals", "oo_check_v4_symmetry", "oo_set_rdms", "oo_energy_grad",
    "oo_energy_grad_host", "oo_transform", "oo_orth", "oo_bb_update", "oo_optimize",
    "oo_nccl_unique_id", "oo_comm_init", "oo_allreduce", "oo_set_timing", "oo_last_timing",
    "oo_launch_count", "oo_measure_peaks", "oo_set_pair_symmetry", "oo_streamed_slabs",
    "oo_ingest_spin_g", "oo_set_rdms_spin", "oo_energy_grad_allreduce", "oo_peer_export",
    "oo_peer_attach", "oo_peer_status", "oo_set_integrals_generic",
    "oo_retraction_stats", "oo_set_callback", "oo_pair_slab_list", "oo_pack_pair_slabs",
    "oo_eval_submit", "oo_eval_wait", "oo_request_stop", "oo_set_peer_timeout_ms",
    "oo_ingest_spin_g_rows",
]

OO_G_V4_SYMMETRIC = 1
OO_G_PAIR_PACKED = 2
CALLBACK_T = C.CFUNCTYPE(None, C.c_int, C.c_double, C.c_void_p)

_lib = None


class OOError(RuntimeError):
    """An error reported by liboo_b200 (status code + oo_last_error())."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"liboo_b200 error {code}: {msg}")
        self.code = code


def _point_at_torch_nccl() -> None:
    """Prefer the NCCL that torch already loaded (same soname) over the system one."""
    if os.environ.get("OO_NCCL_LIB"):
        return
    try:
        import nvidia.nccl  # type: ignore

        for base in list(getattr(nvidia.nccl, "__path__", [])):
            hits = glob.glob(os.path.join(base, "lib", "libnccl.so*"))
            if hits:
                os.environ["OO_NCCL_LIB"] = sorted(hits)[0]
                return
    except Exception:
        pass


def load() -> C.CDLL:
    """Load the library once; raises ImportError (loudly) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            f"g.build()'` (or `make -C {os.path.join(_HERE, 'csrc')}`). There is no CPU fallback.")
    _point_at_torch_nccl()
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    dp, vp, ip, sz = C.POINTER(C.c_double), C.c_void_p, C.POINTER(C.c_int), C.c_size_t
    lib.oo_last_error.restype = C.c_char_p
    lib.oo_version.restype = C.c_char_p
    lib.oo_device_count.restype = C.c_int
    lib.oo_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    lib.oo_destroy.argtypes = [vp]
    lib.oo_set_stream.argtypes = [vp, vp]
    lib.oo_synchronize.argtypes = [vp]
    lib.oo_set_integrals.argtypes = [vp, vp, vp, C.c_uint]
    lib.oo_set_integrals_generic.argtypes = [vp, vp, vp, vp]
    lib.oo_check_v4_symmetry.argtypes = [C.c_int, vp, C.c_int, dp]
    lib.oo_set_rdms.argtypes = [vp, vp, vp]
    lib.oo_energy_grad.argtypes = [vp, vp, vp]
    lib.oo_energy_grad_host.argtypes = [vp, vp, vp, vp]
    lib.oo_energy_grad_allreduce.argtypes = [vp, vp, vp]
    lib.oo_eval_submit.argtypes = [vp, vp, C.c_int]
    lib.oo_eval_wait.argtypes = [vp, C.c_int, vp, vp]
    lib.oo_request_stop.argtypes = [vp]
    lib.oo_set_peer_timeout_ms.argtypes = [vp, C.c_double]
    lib.oo_peer_export.argtypes = [vp, vp]
    lib.oo_peer_attach.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.oo_peer_status.argtypes = [vp]
    lib.oo_transform.argtypes = [vp, vp, vp, vp]
    lib.oo_orth.argtypes = [vp, vp, vp]
    lib.oo_bb_update.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp]
    lib.oo_optimize.argtypes = [vp, vp, C.c_double, C.c_double, C.c_int, C.c_double, vp, C.c_int,
                                ip, dp, dp]
    lib.oo_nccl_unique_id.argtypes = [vp]
    lib.oo_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.oo_allreduce.argtypes = [vp, vp, sz]
    lib.oo_set_timing.argtypes = [vp, C.c_int]
    lib.oo_last_timing.argtypes = [vp, C.POINTER(C.c_float)]
    lib.oo_launch_count.argtypes = [vp]
    lib.oo_launch_count.restype = C.c_longlong
    lib.oo_measure_peaks.argtypes = [C.c_int, sz, dp]
    lib.oo_retraction_stats.argtypes = [vp, ip, ip]
    lib.oo_set_callback.argtypes = [vp, CALLBACK_T, vp]
    lib.oo_set_pair_symmetry.argtypes = [vp, C.c_int]
    lib.oo_streamed_slabs.argtypes = [vp]
    lib.oo_pair_slab_list.argtypes = [C.c_int, C.c_int, C.c_int, ip, C.c_int]
    lib.oo_pack_pair_slabs.argtypes = [vp, vp, vp]
    lib.oo_ingest_spin_g.argtypes = [C.c_int, vp, C.c_int, C.c_double, vp, C.POINTER(C.c_uint), dp]
    lib.oo_ingest_spin_g_rows.argtypes = [C.c_int, vp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                          vp, C.POINTER(C.c_uint), dp]
    lib.oo_set_rdms_spin.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), dp, C.c_int, C.c_uint]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("oo_device_count",):
            pass
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != 0:
        raise OOError(code, load().oo_last_error().decode("utf-8", "replace"))
