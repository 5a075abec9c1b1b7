"""ctypes binding of the C ABI declared in include/mepp_b200.h.

The product path has no CPU fallback: if the shared library is missing this
module raises at import, and every device entry point fails when no sm_100
CUDA device is present.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmepp_b200.so")

MAX_MOTIF_LEN = 32
MAX_MARGIN = 16
ENTRY_LISTS = 64          # MEPP_ENTRY_LISTS
BP_WORDS = 72             # MEPP_BP_WORDS
X_SCALE = 256.0

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -m mepp_b200.build` "
        "(nvcc, sm_100a). mepp_b200 has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)

_i64, _i32, _f64, _vp = C.c_int64, C.c_int, C.c_double, C.c_void_p
_pd, _pf, _pi = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int)

# name -> (restype, argtypes); must list every symbol of include/mepp_b200.h
SIGNATURES = {
    "mepp_abi_version": (_i32, []),
    "mepp_last_error": (C.c_char_p, []),
    "mepp_device_ok": (_i32, []),
    "mepp_log_odds": (_i32, [_pd, _i32, _pd, _f64, _pd]),
    "mepp_threshold_from_p": (_i32, [_pd, _i32, _pd, _f64, _f64, _pd]),
    "mepp_motif_table": (_i32, [_pd, _i32, _i32, _f64, _f64, _pd, _i32, _f64, _pf, _pf, _pi, _pd, _vp, _vp]),
    "mepp_motif_table_with_threshold": (_i32, [_pd, _i32, _i32, _f64, _pd, _i32, _f64, _pf, _pf, _pi, _vp, _vp]),
    "mepp_prepare_tables": (_i32, [_vp, _vp, _vp, _i32, _f64, _f64, _pd, _i32, _f64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mepp_thresholds_dp": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "mepp_task_thresholds": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "mepp_sym_words": (_i64, [_i32]),
    "mepp_pad32": (_i64, [_i64]),
    "mepp_x_planes_bytes": (_i64, [_i64, _i64]),
    "mepp_choose_bn": (_i32, [_i32, _pi]),
    "mepp_y_planes_bytes": (_i64, [_i32, _i64]),
    "mepp_pack_dna": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp]),
    "mepp_build_y": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "mepp_x_scales": (_i32, [_vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "mepp_scan": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32,
                         _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "mepp_plane_words": (_i64, [_i32]),
    "mepp_pack_planes": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "mepp_scan_bitpar_tables": (_i32, [_vp, _vp, _vp, _vp, _i32, _f64, _vp, _vp]),
    "mepp_scan_bitpar": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32,
                                _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mepp_scan_lean_bucket_bytes": (_i64, [_i32, _i64]),
    "mepp_scan_lean": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32,
                              _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "mepp_onehot_words": (_i64, [_i64, _i32]),
    "mepp_pack_onehot": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "mepp_scan_tc_work_bytes": (_i64, [_i64, _i32, _i32]),
    "mepp_scan_tc": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp,
                            _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i32, _vp]),
    "mepp_scan_tc_timing": (_i32, [_i32, _pd, _pd, _pd, _pi]),
    "mepp_heatmap_work_bytes": (_i64, [_i32]),
    "mepp_heatmap": (_i32, [_vp, _vp, _i64, _i32, _i64, _i32, _i32, _i32, C.c_float, _vp, _vp, _vp]),
    "mepp_y_stats": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "mepp_build_y_cols": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "mepp_y_rows_ld": (_i64, [_i32]),
    "mepp_build_y_rows": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "mepp_profile_matches_work_bytes": (_i64, [_i64, _i64, _i32]),
    "mepp_profile_matches": (_i32, [_vp, _vp, _i64, _vp, _i32, _i32, _i32, _vp, _i64, _vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "mepp_profile_sparse_work_bytes": (_i64, [_i64, _i64]),
    "mepp_profile_sparse": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _vp, _vp]),
    "mepp_finalize_out_bytes": (_i64, [_i32, _i32]),
    "mepp_finalize_positions": (_i32, [_vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32,
                                       C.c_double, C.c_double, C.c_double, C.c_double, C.c_float, _vp, _vp]),
    "mepp_write_tsv": (_i32, [C.c_char_p, C.c_char_p, _i32, _vp, _vp, _vp, _i64]),
    "mepp_density_u16": (_i32, [_vp, _i64, _vp, _vp]),
    "mepp_x_planes_reset": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp]),
    "mepp_profile_gemm": (_i32, [_vp, _vp, _vp, _i64, _i32, _i64, _vp]),
    "mepp_gemm_counters": (_i32, [C.POINTER(C.c_ulonglong), _i32]),
    "mepp_profile_gemm_check": (_i32, [_vp, _vp, _vp, _i64, _i32, _i64, _vp, _i32, _vp]),
    "mepp_correlation": (_i32, [_vp, _i64, _vp, _vp, _vp, _pd, _i64, _i64, _i32, _i32, _vp, _i64, _vp]),
    "mepp_col_constants": (_i32, [_vp, _vp, _pd, _i64, _i32, _i32, _vp, _vp]),
    "mepp_col_constants_dev": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "mepp_profile_corr_gemm": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i64, _vp, _i32, _vp]),
    "mepp_perm_stats_work_bytes": (_i64, [_i32, _i32, _i32]),
    "mepp_perm_stats": (_i32, [_vp, _i64, _i32, _i32, _i32, _i32, _i32, _i64, _i64,
                               _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


def last_error():
    return lib.mepp_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError(f"mepp_b200 {what}: {last_error()}")


def device_ok():
    return bool(lib.mepp_device_ok())


def choose_bn(C_cols):
    nt = C.c_int(0)
    bn = lib.mepp_choose_bn(int(C_cols), C.byref(nt))
    return bn, nt.value
