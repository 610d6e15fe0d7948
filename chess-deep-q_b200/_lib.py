"""ctypes binding of libcdq.so (include/cdq.h).  Fails loudly when the library is missing or a
call returns an error; nothing here falls back to the CPU."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("CDQ_LIBCDQ") or os.path.join(_HERE, "libcdq.so")  # env override: A/B builds

c_void_p, c_int64, c_int, c_uint64, c_size_t = (ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                                 ctypes.c_uint64, ctypes.c_size_t)

# name -> (restype, argtypes); must list every symbol include/cdq.h declares
SIGNATURES = {
    "cdq_version": (c_int, []),
    "cdq_error_string": (ctypes.c_char_p, [c_int]),
    "cdq_set_option": (c_int, [ctypes.c_char_p, c_int]),
    "cdq_get_option": (c_int, [ctypes.c_char_p, ctypes.POINTER(c_int)]),
    "cdq_eval_terms": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cdq_attack_maps": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "cdq_encode_planes": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "cdq_net_create": (c_int, [ctypes.POINTER(c_void_p)]),
    "cdq_net_load": (c_int, [c_void_p, c_void_p, c_void_p]),
    "cdq_net_destroy": (None, [c_void_p]),
    "cdq_workspace_bytes": (c_size_t, [c_int64]),
    "cdq_value_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cdq_leaf_eval": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_size_t, c_void_p]),
    "cdq_leaf_eval_host": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "cdq_host_ctx_create": (c_int, [ctypes.POINTER(c_void_p)]),
    "cdq_host_ctx_destroy": (None, [c_void_p]),
    "cdq_leaf_eval_host_ctx": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "cdq_host_release": (None, []),
    "cdq_replay_append": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                  c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, c_void_p]),
    "cdq_replay_sample": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_void_p, c_int64, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    "cdq_replay_permute": (ctypes.c_uint32, [ctypes.c_uint32, ctypes.c_uint32, c_uint64, ctypes.c_uint32]),
    "cdq_train_workspace_bytes": (c_size_t, [c_int64]),
    "cdq_train_fwd_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                  ctypes.c_double, ctypes.c_float, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cdq_train_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                               ctypes.c_double, ctypes.c_float, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "cdq_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, ctypes.c_float, ctypes.c_float,
                              ctypes.c_float, ctypes.c_float, ctypes.c_float, c_void_p]),
    "cdq_dp_adam_step": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, ctypes.c_float,
                                 ctypes.c_float, ctypes.c_float, ctypes.c_float, c_void_p]),
    "cdq_dp_adam_status": (c_int, [c_void_p, c_void_p]),
    "cdq_ipc_alloc": (c_int, [c_size_t, ctypes.POINTER(c_void_p), c_void_p]),
    "cdq_ipc_open": (c_int, [c_void_p, ctypes.POINTER(c_void_p)]),
    "cdq_ipc_close": (c_int, [c_void_p]),
    "cdq_ipc_free": (c_int, [c_void_p]),
    "cdq_random_playouts": (c_int, [c_void_p, c_int64, c_uint64, c_int, c_void_p]),
    "cdq_expand": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "cdq_expand_weighted": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cdq_sample_children": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                    c_void_p]),
    "cdq_forest_create": (c_int, [c_void_p, ctypes.POINTER(c_void_p)]),
    "cdq_forest_destroy": (None, [c_void_p]),
    "cdq_forest_set_games": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "cdq_forest_set_active": (c_int, [c_void_p, c_void_p, c_void_p]),
    "cdq_forest_search": (c_int, [c_void_p, c_void_p, c_int, ctypes.c_float, ctypes.c_float, c_int, c_void_p]),
    "cdq_forest_advance": (c_int, [c_void_p, c_void_p]),
    "cdq_forest_view": (c_int, [c_void_p, c_void_p]),
    "cdq_launch_count": (c_uint64, []),
    "cdq_profile_enable": (None, [c_int]),
    "cdq_profile_collect": (c_int, [c_void_p, c_void_p]),
    "cdq_profile_collect_ex": (c_int, [c_void_p, c_void_p]),
    "cdq_debug_conv": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
}


class CdqWeights(ctypes.Structure):
    _fields_ = [(k, c_void_p) for k in ("conv1_w", "conv1_b", "conv2_w", "conv2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b")]


class CdqForestConfig(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in ("n_games", "workers", "depth", "width", "max_iterations")]


class CdqForestView(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in ("n_games", "workers", "depth", "width", "max_nodes")] + \
               [(k, c_void_p) for k in ("game_boards", "game_halfmove_clock", "game_ply", "game_active", "game_status",
                                        "game_history_len", "best_move", "best_visits", "root_children", "edge_move",
                                        "edge_visits", "edge_value", "node_children", "node_count", "leaf_boards",
                                        "leaf_values", "leaf_depth", "positions_evaluated")]


MAX_RANKS = 8


class CdqDpPeers(ctypes.Structure):
    _fields_ = [("params", c_void_p * MAX_RANKS), ("grads", c_void_p * MAX_RANKS), ("flags", c_void_p * MAX_RANKS),
                ("mc_params", c_void_p), ("mc_grads", c_void_p)]


class CdqOptimizer(ctypes.Structure):
    _fields_ = [("peers", CdqDpPeers), ("rank", ctypes.c_int32), ("world", ctypes.c_int32), ("d_m", c_void_p), ("d_v", c_void_p),
                ("n_params_padded", c_int64), ("d_step", c_void_p), ("d_grid_bar", c_void_p), ("lr", ctypes.c_float),
                ("beta1", ctypes.c_float), ("beta2", ctypes.c_float), ("eps", ctypes.c_float)]


class CdqError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libcdq.so (built by chess-deep-q_b200/build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise CdqError("libcdq.so is not built: run `python chess-deep-q_b200/build.py` "
                           "(there is no CPU fallback)")
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def set_option(name, value):
    """cdq_set_option: run-time switch (see include/cdq.h); returns the previous value."""
    old = c_int()
    check(lib().cdq_get_option(name.encode(), ctypes.byref(old)))
    check(lib().cdq_set_option(name.encode(), int(value)))
    return old.value


def check(code):
    if code != 0:
        raise CdqError("libcdq error %d: %s" % (code, lib().cdq_error_string(code).decode()))


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise CdqError("no CUDA device: chess-deep-q_b200 computes only on a B200 (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())
