"""ctypes binding of include/smg_b200.h.  Fails loudly if the CUDA library is missing: there is no CPU fallback."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libsmg_b200.so")

c_float_p = ctypes.POINTER(ctypes.c_float)
c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
c_int64_p = ctypes.POINTER(ctypes.c_int64)
c_void_pp = ctypes.POINTER(ctypes.c_void_p)
VP = ctypes.c_void_p
I = ctypes.c_int

class TrainStepArgs(ctypes.Structure):
    """smg_train_step_args of include/smg_b200.h."""
    _fields_ = [("trunk_id", ctypes.c_int32), ("head_id", ctypes.c_int32), ("rot_idx", ctypes.c_int32),
                ("num_rotations", ctypes.c_int32), ("hm_size", ctypes.c_int32), ("loss_kind", ctypes.c_int32),
                ("adam_step", ctypes.c_int32), ("label", ctypes.c_float), ("mean", ctypes.c_double),
                ("stddev", ctypes.c_double), ("class_weight", ctypes.c_float * 3), ("lr", ctypes.c_float),
                ("beta1", ctypes.c_float), ("beta2", ctypes.c_float), ("eps", ctypes.c_float), ("flags", ctypes.c_int32)]


# name -> (restype, argtypes); must list every symbol include/smg_b200.h declares
PROTOTYPES = {
    "smg_version": (I, []),
    "smg_last_error": (ctypes.c_char_p, []),
    "smg_create": (I, [I, I, I, c_void_pp]),
    "smg_destroy": (I, [VP]),
    "smg_set_precision": (I, [VP, I]),
    "smg_get_precision": (I, [VP]),
    "smg_workspace_bytes": (ctypes.c_int64, [VP]),
    "smg_set_pack_layouts": (I, [VP, I]),
    "smg_set_trunk_weights": (I, [VP, I, c_void_pp, I, VP]),
    "smg_set_head_weights": (I, [VP, I, c_void_pp, I, VP]),
    "smg_prep": (I, [VP, VP, I, I, ctypes.c_double, ctypes.c_double, VP, VP]),
    "smg_rotate": (I, [VP, VP, c_int_p, I, I, VP, VP]),
    "smg_rotate_index_map": (I, [VP, I, I, VP, VP]),
    "smg_trunk_forward": (I, [VP, I, VP, I, VP, VP, VP, VP]),
    "smg_qforward": (I, [VP, I, I, VP, VP, I, c_int_p, I, I, VP, VP, VP, VP]),
    "smg_qpartials": (I, [VP, I, I, VP, c_int_p, I, I, VP, I, I, ctypes.c_double, ctypes.c_double, VP, VP]),
    "smg_qcombine": (I, [VP, I, VP, I, VP, I, VP, VP]),
    "smg_qforward_maps": (I, [VP, I, I, VP, VP, I, I, ctypes.c_double, ctypes.c_double, c_int_p, I, I, VP, VP, VP, VP]),
    "smg_qforward_maps_batch": (I, [VP, I, I, VP, VP, I, I, I, ctypes.c_double, ctypes.c_double, c_int_p, I, I, VP, VP, VP, VP]),
    "smg_head_bn_stats": (I, [VP, VP, I, VP]),
    "smg_qforward_train": (I, [VP, I, I, VP, VP, I, I, VP, VP, VP, VP]),
    "smg_qbackward": (I, [VP, VP, c_void_pp, c_void_pp, VP]),
    "smg_train_pass_id": (ctypes.c_int64, [VP]),
    "smg_train_step": (I, [VP, ctypes.POINTER(TrainStepArgs), VP, VP, c_void_pp, c_void_pp, c_void_pp, c_void_pp, I, VP, VP,
                           VP, VP, VP]),
    "smg_adam_step": (I, [VP, c_void_pp, c_void_pp, c_void_pp, c_void_pp, c_int64_p, I, I,
                          ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, VP]),
    "smg_argmax": (I, [VP, VP, I, VP, VP, VP]),
    "smg_heightmap": (I, [VP, VP, c_double_p, c_double_p, VP, VP, c_double_p, VP]),
    "smg_heightmap_color": (I, [VP, VP, VP, VP, VP]),
    "smg_resize_masks": (I, [VP, VP, I, I, I, VP, VP]),
    "smg_geometry": (I, [VP, I, VP, I, I, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, I, I, I, c_double_p,
                         c_double_p, VP]),
    "smg_nms": (I, [VP, VP, I, ctypes.c_float, ctypes.c_float, ctypes.c_float, VP, VP, VP]),
    "smg_launch_count": (ctypes.c_int64, [VP]),
    "smg_debug_read": (I, [VP, ctypes.c_char_p, I, VP, ctypes.c_int64, VP]),
    "smg_profile_enable": (I, [VP, I]),
    "smg_profile_read": (I, [VP, c_double_p, c_int64_p, c_double_p, c_double_p]),
    "smg_debug_dgrad": (I, [VP, I, VP, I, I, I, I, I, I, VP, I, VP, VP]),
    "smg_debug_wgrad": (I, [VP, I, VP, I, I, VP, I, I, I, I, VP, I, VP, VP, VP, VP]),
    "smg_debug_bn_bwd": (I, [VP, VP, I, I, VP, I, VP, I, VP, VP, I, I, I, I, VP, VP, I, I, VP, VP, VP]),
    "smg_debug_conv": (I, [VP, I, VP, I, I, I, I, VP, VP, I, I, I, VP, I, VP, I, I, VP, VP]),
}

_lib = None


class SmgError(RuntimeError):
    pass


def load():
    """Load libsmg_b200.so (built by build.py / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SmgError(
            "libsmg_b200.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU or PyTorch fallback for this path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise SmgError("libsmg_b200: status %d: %s" % (status, load().smg_last_error().decode("utf-8", "replace")))
