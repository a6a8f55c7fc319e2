"""ctypes binding of libcable_b200.so (the C ABI in include/cable_b200.h).

This is plumbing only.  The product path is the CUDA library; if it is missing we
raise -- there is deliberately no Python/NumPy fallback for the physics.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcable_b200.so")
# tuning aid: CABLE_B200_LIB points at a variant build of the same library (cable_b200/csrc/Makefile, OUT=...)
if os.environ.get("CABLE_B200_LIB"):
    LIB_PATH = os.path.abspath(os.environ["CABLE_B200_LIB"])

MS, MSN, MF, NRB, NCP, NCS = 6, 3, 2, 3, 3, 2


class CableCfg(C.Structure):
    """Mirror of `struct cable_cfg` (include/cable_b200.h)."""
    _fields_ = [
        ("struct_bytes", C.c_int),
        ("gs_switch", C.c_int), ("fwsoil_switch", C.c_int), ("ssnow_potev", C.c_int),
        ("diag_soil_resp_on", C.c_int), ("l_new_runoff_speed", C.c_int), ("l_new_reduce_soilevp", C.c_int),
        ("litter", C.c_int), ("or_evap", C.c_int), ("gw_model", C.c_int), ("l_rev_corr", C.c_int),
        ("soil_thermal_fix", C.c_int), ("l_new_roughness_soil", C.c_int), ("call_climate", C.c_int),
        ("redistrb", C.c_int), ("soil_struc_sli", C.c_int),
        ("runtime_um", C.c_int), ("icycle", C.c_int), ("mvtype", C.c_int),
        ("snmin", C.c_float), ("max_glacier_snowd", C.c_float), ("snow_ccnsw", C.c_float),
        ("max_ssdn", C.c_float), ("max_sconds", C.c_float), ("frozen_limit", C.c_float),
        ("wiltParam", C.c_float), ("satuParam", C.c_float),
        ("zse", C.c_float * MS), ("zshh", C.c_float * (MS + 1)),
        ("ratecp", C.c_float * NCP), ("ratecs", C.c_float * NCS),
        ("met_tv_is_tk", C.c_int), ("caller_duties", C.c_int), ("output_level", C.c_int),
        ("n_forcing_slots", C.c_int), ("threads_per_block", C.c_int),
    ]


class FieldInfo(C.Structure):
    _fields_ = [("name", C.c_char_p), ("dtype", C.c_int), ("n1", C.c_int), ("n2", C.c_int),
                ("role", C.c_uint), ("flags", C.c_uint)]


class Counters(C.Structure):
    _fields_ = [("steps", C.c_longlong), ("kernel_launches", C.c_longlong), ("h2d_bytes", C.c_longlong),
                ("d2h_bytes", C.c_longlong), ("kernel_ms", C.c_double), ("kernel_ms_count", C.c_longlong),
                ("n_dryleaf_warn", C.c_longlong), ("n_fastdiv_redo_blocks", C.c_longlong)]


EXPORTS = [
    "cable_b200_abi_version", "cable_b200_last_error", "cable_b200_nfields", "cable_b200_field_id",
    "cable_b200_field_info", "cable_b200_default_cfg", "cable_b200_create", "cable_b200_destroy",
    "cable_b200_bind_field", "cable_b200_upload", "cable_b200_download", "cable_b200_set_forcing_async",
    "cable_b200_step", "cable_b200_cbm", "cable_b200_sync", "cable_b200_device_ptr",
    "cable_b200_compute_stream", "cable_b200_profile", "cable_b200_get_counters",
    "cable_b200_reset_counters", "cable_b200_grid_reduce", "cable_b200_mark_dirty", "cable_b200_set_output_mask",
    "cable_b200_param_table_classes",
    # driver stages either side of cbm()
    "cable_b200_driver_init", "cable_b200_set_met_async", "cable_b200_upload_lai", "cable_b200_post_step",
    "cable_b200_output_plan", "cable_b200_driver_field_id", "cable_b200_output_accumulate",
    "cable_b200_output_fetch_async", "cable_b200_output_wait", "cable_b200_driver_download",
    # multi-GPU gather of the output block (NCCL, bound at run time)
    "cable_b200_comm_unique_id", "cable_b200_comm_init", "cable_b200_comm_destroy", "cable_b200_output_gather_async",
]

MET_ROWS = ("SWdown", "Tair", "Qair", "PSurf", "Wind", "Rainf", "Snowf", "LWdown", "CO2air", "hod", "doy")
AGG = {"point": 0, "mean": 1, "sum": 2, "min": 3, "max": 4}


class MetConvert(C.Structure):
    """Mirror of `struct cable_met_convert` (cable_input.F90:1053-1209 convert%*)."""
    _fields_ = [("tair_offset", C.c_float), ("psurf_scale", C.c_float), ("rainf_scale", C.c_float),
                ("co2_scale", C.c_float), ("snowf_from_tair", C.c_int)]

_lib = None


def load() -> C.CDLL:
    """Load the CUDA extension; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(cable_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    H = C.c_void_p
    lib.cable_b200_abi_version.restype = C.c_int
    lib.cable_b200_last_error.restype = C.c_char_p
    lib.cable_b200_nfields.restype = C.c_int
    lib.cable_b200_field_id.argtypes = [C.c_char_p]
    lib.cable_b200_field_info.argtypes = [C.c_int, C.POINTER(FieldInfo)]
    lib.cable_b200_default_cfg.argtypes = [C.POINTER(CableCfg)]
    lib.cable_b200_default_cfg.restype = None
    lib.cable_b200_create.argtypes = [C.c_int, C.POINTER(CableCfg), C.c_int, C.POINTER(H)]
    lib.cable_b200_destroy.argtypes = [H]
    lib.cable_b200_bind_field.argtypes = [H, C.c_int, C.c_void_p]
    lib.cable_b200_upload.argtypes = [H, C.c_uint]
    lib.cable_b200_download.argtypes = [H, C.c_uint, C.c_uint]
    lib.cable_b200_set_forcing_async.argtypes = [H, C.c_int]
    lib.cable_b200_step.argtypes = [H, C.c_int, C.c_float, C.c_int]
    lib.cable_b200_cbm.argtypes = [H, C.c_int, C.c_float]
    lib.cable_b200_sync.argtypes = [H]
    lib.cable_b200_device_ptr.argtypes = [H, C.c_int, C.c_int]
    lib.cable_b200_device_ptr.restype = C.c_void_p
    lib.cable_b200_compute_stream.argtypes = [H]
    lib.cable_b200_compute_stream.restype = C.c_void_p
    lib.cable_b200_profile.argtypes = [H, C.c_int]
    lib.cable_b200_get_counters.argtypes = [H, C.POINTER(Counters)]
    lib.cable_b200_reset_counters.argtypes = [H]
    lib.cable_b200_grid_reduce.argtypes = [H, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.cable_b200_param_table_classes.argtypes = [H]
    lib.cable_b200_mark_dirty.argtypes = [H, C.c_int]
    lib.cable_b200_set_output_mask.argtypes = [H, C.c_void_p, C.c_int]
    lib.cable_b200_driver_init.argtypes = [H, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.cable_b200_set_met_async.argtypes = [H, C.c_int, C.c_void_p, C.POINTER(MetConvert)]
    lib.cable_b200_upload_lai.argtypes = [H]
    lib.cable_b200_post_step.argtypes = [H, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int]
    lib.cable_b200_output_plan.argtypes = [H, C.c_int] + [C.c_void_p] * 6
    lib.cable_b200_driver_field_id.argtypes = [C.c_char_p]
    lib.cable_b200_output_accumulate.argtypes = [H]
    lib.cable_b200_output_fetch_async.argtypes = [H, C.c_void_p]
    lib.cable_b200_output_wait.argtypes = [H]
    lib.cable_b200_driver_download.argtypes = [H, C.c_char_p, C.c_void_p]
    lib.cable_b200_comm_unique_id.argtypes = [C.c_void_p]
    lib.cable_b200_comm_init.argtypes = [H, C.c_void_p, C.c_int, C.c_int]
    lib.cable_b200_comm_destroy.argtypes = [H]
    lib.cable_b200_output_gather_async.argtypes = [H, C.c_int, C.c_void_p, C.c_void_p]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is None and name != "cable_b200_default_cfg":
            fn.restype = C.c_int
    _lib = lib
    return lib


def default_cfg() -> CableCfg:
    cfg = CableCfg()
    load().cable_b200_default_cfg(C.byref(cfg))
    return cfg


class CableError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"cable_b200 error {code}: {msg}")
        self.code = code


def check(rc: int) -> None:
    if rc != 0:
        raise CableError(rc, load().cable_b200_last_error().decode())
