"""ctypes binding of include/sejonggo_b200.h.  No CPU fallback: if the CUDA
library is missing or a symbol is absent this raises, it never degrades."""
import ctypes as C
import os
import re

from . import _build

HEADER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "sejonggo_b200.h")


class SgoConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("size", C.c_int32), ("n_games", C.c_int32), ("trees_per_game", C.c_int32),
                ("max_leaves", C.c_int32), ("arena_blocks", C.c_int32), ("komi", C.c_float), ("reserved", C.c_int32 * 9)]


def header_symbols():
    """Every function the header declares (the drop-in boundary)."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sgo_[a-z0-9_]+)\s*\(", text)))


_lib = None
vp, i32, i64, u64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double

_PROTOS = {
    "sgo_create": [C.POINTER(SgoConfig), C.POINTER(vp)],
    "sgo_destroy": [vp],
    "sgo_check_errors_sync": [vp, vp, C.POINTER(i32)],
    "sgo_abi_version": [],
    "sgo_games_reset": [vp, i32, i32, vp],
    "sgo_apply_moves": [vp, i32, i32, vp, vp, vp],
    "sgo_legal_masks": [vp, i32, i32, vp, vp],
    "sgo_score": [vp, i32, i32, vp, vp],
    "sgo_import_boards": [vp, i32, i32, vp, vp],
    "sgo_export_boards": [vp, i32, i32, vp, vp],
    "sgo_export_packed": [vp, i32, i32, i32, vp, vp],
    "sgo_random_playouts": [vp, i32, i32, u64, i32, vp, vp, vp],
    "sgo_export_planes": [vp, i32, i32, i32, i32, vp, vp, vp],
    "sgo_export_planes_indexed": [vp, i32, vp, i32, vp, vp, vp],
    "sgo_policy_unsym": [vp, i32, i32, vp, vp, vp, vp],
    "sgo_tree_new": [vp, vp, vp, vp, f64, i32, vp],
    "sgo_tree_reset": [vp, vp],
    "sgo_tree_free": [vp, vp, vp],
    "sgo_games_restart": [vp, vp, vp],
    "sgo_pool_stats_sync": [vp, vp, vp],
    "sgo_tree_sizes": [vp, vp, vp],
    "sgo_tree_select_a": [vp, vp, i32, vp],
    "sgo_tree_select_b_sync": [vp, vp, i32, i32, C.POINTER(i32), vp],
    "sgo_tree_expand": [vp, vp, vp, vp, vp],
    "sgo_tree_backup_a": [vp, vp, vp],
    "sgo_tree_backup_b": [vp, vp, i32, vp],
    "sgo_tree_pick": [vp, vp, vp, vp, vp, vp, vp],
    "sgo_tree_reroot": [vp, vp, vp],
    "sgo_tree_child_stats": [vp, vp, vp, vp, vp, vp],
    "sgo_tree_download_sync": [vp, i32, vp, i32, vp, vp],
    "sgo_tree_upload_sync": [vp, i32, vp, i32, vp, vp],
    "sgo_leaf_counts": [vp, vp, vp],
    "sgo_tree_valid": [vp, vp, vp, vp],
    "sgo_leaf_compact_sync": [vp, vp, C.POINTER(i32), vp],
    "sgo_tower_load_weights": [vp, i32, vp, i32, vp],
    "sgo_tower_free": [vp, i32],
    "sgo_tower_max_positions": [vp, i32],
    "sgo_selfplay_step": [vp, i32, vp, vp, i32, i32, vp, i32, C.POINTER(i32), vp],
    "sgo_record_words": [vp],
    "sgo_records_pack": [vp, vp, vp, vp, vp, vp],
    "sgo_tower_forward": [vp, i32, i32, vp, i32, vp, i32, vp, vp, vp],
    "sgo_tower_check_sync": [vp, i32, C.POINTER(i32), vp],
    "sgo_tower_profile": [vp, i32, i32],
    "sgo_tower_profile_read_sync": [vp, i32, vp],
    "sgo_tower_debug_conv": [vp, i32, i32, i32, i32, i32, i32, vp],
    "sgo_tower_act_copy": [vp, i32, i32, i32, vp, i32, vp],
}


class SgoTowerWeights(C.Structure):
    _fields_ = [("n_blocks", C.c_int32), ("channels", C.c_int32), ("size", C.c_int32), ("reserved", C.c_int32)] + \
        [(k, C.c_void_p) for k in ("stem_w", "stem_b", "conv_w", "conv_b", "pol_conv_w", "pol_conv_b", "pol_fc_w", "pol_fc_b",
                                   "val_conv_w", "val_conv_b", "val_fc1_w", "val_fc1_b", "val_fc2_w", "val_fc2_b")]


def load(build_if_needed=True):
    global _lib
    if _lib is not None:
        return _lib
    so = _build.SO
    if os.environ.get("SGO_LIBRARY"):                 # A/B runs against another build of the same ABI (tools/, profiles/)
        so = os.environ["SGO_LIBRARY"]
    elif build_if_needed:
        so = _build.build()
    if not os.path.isfile(so):
        raise RuntimeError("sejonggo_b200: CUDA library %s is missing — run `python -m sejonggo_b200._build`; "
                           "there is no CPU fallback" % so)
    lib = C.CDLL(so)
    for name in header_symbols():
        if not hasattr(lib, name):
            raise RuntimeError("sejonggo_b200: %s does not export %s" % (so, name))
    for name, args in _PROTOS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.sgo_launch_count.argtypes = [vp]
    lib.sgo_launch_count.restype = C.c_int64
    lib.sgo_last_error.argtypes = [vp]
    lib.sgo_last_error.restype = C.c_char_p
    _lib = lib
    return lib
