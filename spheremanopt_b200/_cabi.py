"""ctypes binding of the C ABI declared in include/smo_b200.h.

``load()`` opens the in-tree CUDA library ``spheremanopt_b200/libsmo_b200.so`` (built by
``__graft_entry__.build()`` / ``python -m spheremanopt_b200._build``) and nothing else: there is no
CPU fallback in the product path - a missing library raises ``ImportError`` loudly.

``bind(cdll)`` only attaches argument/return types to an already opened library; the CPU test-suite uses it on
the host-emulation build of the same sources (tests/emul), which is test infrastructure.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# (SMO_B200_LIB: development only - an alternative build of the same sources, e.g. a tuning variant under build/)
LIB_PATH = os.environ.get("SMO_B200_LIB") or os.path.join(_HERE, "libsmo_b200.so")

vp, dp, ll, i32, f64, sz = C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_double, C.c_size_t

# name -> (restype, argtypes).  Mirrors include/smo_b200.h one to one (tests check the two against each other).
SIGNATURES = {
    "smo_version": (i32, []),
    "smo_last_error": (C.c_char_p, []),
    "smo_launch_count": (ll, []),
    # SH23
    "smo_sh23_create": (i32, [C.POINTER(vp), i32, f64, f64]),
    "smo_sh23_destroy": (i32, [vp]),
    "smo_sh23_snapshot_bytes": (sz, [vp, i32]),
    "smo_sh23_snapshot_coef": (i32, [vp, vp, i32, i32, i32, vp, vp]),
    "smo_sh23_forward": (i32, [vp, dp, i32, f64, i32, vp, dp, vp]),
    "smo_sh23_adjoint": (i32, [vp, i32, f64, i32, vp, dp, i32, vp]),
    "smo_sh23_prep": (i32, [vp, dp, i32, f64, i32, dp, vp]),
    "smo_sh23_to_coef": (i32, [vp, dp, i32, vp, vp]),
    "smo_sh23_to_grid": (i32, [vp, vp, i32, dp, vp]),
    "smo_sh23_forward_host": (i32, [vp, dp, i32, f64, i32, vp, dp, vp]),
    "smo_sh23_adjoint_host": (i32, [vp, i32, f64, i32, vp, dp, i32, vp]),
    "smo_sh23_prep_host": (i32, [vp, dp, i32, f64, i32, dp, vp]),
    # kinematic dynamo
    "smo_kdyn_create": (i32, [C.POINTER(vp), i32, f64, i32, i32, vp]),
    "smo_kdyn_destroy": (i32, [vp]),
    "smo_kdyn_grid_elems": (sz, [vp]),
    "smo_kdyn_coef_elems": (sz, [vp]),
    "smo_kdyn_snapshot_bytes": (sz, [vp, i32]),
    "smo_kdyn_segment_bytes": (sz, [vp, i32]),
    "smo_kdyn_snapshot_coef": (i32, [vp, vp, i32, i32, vp, vp]),
    "smo_kdyn_forward": (i32, [vp, dp, dp, f64, f64, i32, vp, C.POINTER(f64), i32, vp]),
    "smo_kdyn_adjoint": (i32, [vp, f64, f64, i32, vp, dp, dp, i32, vp]),
    "smo_kdyn_prep": (i32, [vp, dp, dp, f64, f64, i32, dp, vp]),
    "smo_kdyn_checkpoint_bytes": (sz, [vp, i32, i32]),
    "smo_kdyn_forward_ckpt": (i32, [vp, dp, dp, f64, f64, i32, i32, vp, C.POINTER(f64), i32, vp]),
    "smo_kdyn_adjoint_ckpt": (i32, [vp, f64, f64, i32, i32, vp, vp, dp, dp, i32, vp]),
    "smo_kdyn_forward_host": (i32, [vp, dp, dp, f64, f64, i32, vp, C.POINTER(f64), i32, vp]),
    "smo_kdyn_adjoint_host": (i32, [vp, f64, f64, i32, vp, dp, dp, i32, vp]),
    "smo_kdyn_prep_host": (i32, [vp, dp, dp, f64, f64, i32, dp, vp]),
    "smo_kdyn_slab_copy": (i32, [vp, dp, dp, dp, vp]),
    "smo_kdyn_to_coef": (i32, [vp, dp, vp, vp]),
    "smo_kdyn_to_grid": (i32, [vp, vp, dp, vp]),
    "smo_kdyn_profile_set": (i32, [vp, i32]),
    "smo_kdyn_profile_read": (i32, [vp, C.POINTER(f64), C.POINTER(ll)]),
    "smo_kdyn_peer_handle_bytes": (i32, []),
    "smo_kdyn_peer_export": (i32, [vp, vp]),
    "smo_kdyn_peer_attach": (i32, [vp, vp]),
    "smo_kdyn_set_chunks": (i32, [vp, i32, i32]),
    "smo_kdyn_use_graph": (i32, [vp, i32]),
    "smo_kdyn_set_option": (i32, [vp, i32, i32]),
    # communicator (multi-GPU slab decomposition)
    "smo_comm_unique_id_bytes": (i32, []),
    "smo_comm_get_unique_id": (i32, [vp]),
    "smo_comm_create": (i32, [C.POINTER(vp), vp, i32, i32]),
    "smo_comm_destroy": (i32, [vp]),
    # vectors
    "smo_vec_work_bytes": (sz, [ll]),
    "smo_vec_dot": (i32, [dp, dp, ll, f64, C.POINTER(f64), vp, vp]),
    "smo_vec_dot_dev": (i32, [dp, dp, ll, f64, vp, vp]),
    "smo_microbench_dfma": (i32, [dp, i32, i32, C.POINTER(f64), vp]),
    "smo_vec_dot_rows": (i32, [dp, dp, i32, ll, f64, dp, vp]),
    "smo_vec_checksum": (i32, [dp, ll, C.POINTER(C.c_ulonglong), vp, vp]),
    "smo_vec_axpby": (i32, [f64, dp, f64, dp, dp, ll, vp]),
    "smo_vec_project": (i32, [dp, dp, dp, ll, vp, vp]),
    "smo_vec_retract": (i32, [dp, f64, dp, f64, f64, dp, ll, vp, vp]),
}

SMO_ADJOINT_CONTINUOUS = 1
SMO_COST_INTEGRATED = 2
SMO_OPT_KERNEL_SYNC = 2
SMO_OPT_PEER_PULL = 3
SMO_OPT_L2_HINTS = 4
SMO_OPT_PUSH_WAVES = 5
SMO_OPT_TWO_STREAMS = 6
SMO_OPT_GRID_ACC = 7
SMO_OPT_BULK_U = 8
SMO_OPT_TMA_SIN = 9
SMO_OPT_PDL = 10
SMO_OPT_BULK_PUSH = 11


def bind(cdll):
    """attach the header's prototypes to an opened library; raises AttributeError on a missing symbol"""
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(cdll, name)
        fn.restype = res
        fn.argtypes = args
    return cdll


_lib = None


def load():
    """the product library (CUDA, sm_100a).  No fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "spheremanopt_b200: CUDA library %s is missing - build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
                "There is no CPU fallback." % LIB_PATH)
        _lib = bind(C.CDLL(LIB_PATH))
    return _lib


def check(lib, rc):
    if rc != 0:
        msg = lib.smo_last_error()
        raise RuntimeError("libsmo_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))
