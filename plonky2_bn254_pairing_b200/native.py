"""ctypes binding of libbnp.so (include/bnp.h).  The library is the product; this module only loads it.

There is deliberately no fallback: if the shared library is missing, or no CUDA device is usable,
every entry point raises.  Build it with `python -c "import __graft_entry__ as g; g.build()"`.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BNP_LIB", os.path.join(_HERE, "libbnp.so"))

_u64p = ctypes.c_void_p
_sz = ctypes.c_size_t
_int = ctypes.c_int

# name -> (restype, argtypes); every symbol include/bnp.h declares
SIGNATURES = {
    "bnp_init": (_int, [ctypes.POINTER(_int), _int]),
    "bnp_shutdown": (None, []),
    "bnp_device_count": (_int, []),
    "bnp_strerror": (ctypes.c_char_p, [_int]),
    "bnp_last_error": (ctypes.c_char_p, []),
    "bnp_miller_loop_batch": (_int, [_u64p, _u64p, _u64p, _sz]),
    "bnp_multi_miller_loop_batch": (_int, [_u64p, _u64p, _u64p, _sz, _int]),
    "bnp_final_exp_batch": (_int, [_u64p, _u64p, _sz, _int]),
    "bnp_final_exp_witness_batch": (_int, [_u64p, _u64p, _sz]),
    "bnp_pairing_batch": (_int, [_u64p, _u64p, _u64p, _sz, _int]),
    "bnp_multi_pairing_batch": (_int, [_u64p, _u64p, _u64p, _sz, _int, _int]),
    "bnp_pairing_product": (_int, [_u64p, _u64p, _u64p, _sz, _int]),
    "bnp_frobenius_batch": (_int, [_u64p, _u64p, _sz, _sz]),
    "bnp_fq12_mul_batch": (_int, [_u64p, _u64p, _u64p, _sz]),
    "bnp_validate_batch": (_int, [_u64p, _u64p, ctypes.c_void_p, _sz]),
    "bnp_validate_dev": (_int, [_int, ctypes.c_void_p, _u64p, _u64p, ctypes.c_void_p, _sz]),
    "bnp_scalar_mul_batch": (_int, [_int, _u64p, _u64p, _u64p, ctypes.c_void_p, _sz]),
    "bnp_scalar_mul_dev": (_int, [_int, ctypes.c_void_p, _int, _u64p, _u64p, _u64p, ctypes.c_void_p, _sz]),
    "bnp_pow_u64_batch": (_int, [_u64p, _u64p, _sz, _u64p, _sz]),
    "bnp_pow_u64_dev": (_int, [_int, ctypes.c_void_p, _u64p, _u64p, _sz, _u64p, _sz]),
    "bnp_miller_loop_dev": (_int, [_int, ctypes.c_void_p, _u64p, _u64p, _u64p, _sz, _int]),
    "bnp_miller_loop_fused_dev": (_int, [_int, ctypes.c_void_p, _u64p, _u64p, _u64p, _sz]),
    "bnp_final_exp_dev": (_int, [_int, ctypes.c_void_p, _u64p, _u64p, _sz, _int]),
    "bnp_final_exp_witness_dev": (_int, [_int, ctypes.c_void_p, _u64p, _u64p, _sz]),
    "bnp_pairing_dev": (_int, [_int, ctypes.c_void_p, _u64p, _u64p, _u64p, _sz, _int, _int]),
    "bnp_frobenius_dev": (_int, [_int, ctypes.c_void_p, _u64p, _u64p, _sz, _sz]),
    "bnp_fq12_mul_dev": (_int, [_int, ctypes.c_void_p, _u64p, _u64p, _u64p, _sz]),
    "bnp_fq12_product_dev": (_int, [_int, ctypes.c_void_p, _u64p, _u64p, _sz]),
    "bnp_program_macs": (ctypes.c_uint64, [ctypes.c_char_p]),
    "bnp_program_macs_executed": (ctypes.c_uint64, [ctypes.c_char_p]),
    "bnp_launch_count": (ctypes.c_uint64, []),
    "bnp_gather_transport": (ctypes.c_char_p, []),
    "bnp_imad_peak": (_int, [_int, ctypes.POINTER(ctypes.c_double)]),
    "bnp_imad32_peak": (_int, [_int, ctypes.POINTER(ctypes.c_double)]),
    "bnp_run_program_dev": (_int, [_int, ctypes.c_void_p, ctypes.c_char_p, _u64p, _u64p, _u64p, _u64p, _u64p, _sz]),
    "bnp_set_launch_config": (_int, [_int, _int]),
    "bnp_threads_per_block": (_int, []),
    "bnp_g2_prepare_batch": (_int, [_u64p, _u64p, _sz]),
    "bnp_pairing_prepared_batch": (_int, [_u64p, _u64p, _u64p, _u64p, _sz, _int, _int, _int]),
    "bnp_pairing_prepared_dev": (_int, [_int, ctypes.c_void_p, _u64p, _u64p, _u64p, _u64p, _sz, _int, _int, _int]),
    "bnp_decode_g1_batch": (_int, [_int, ctypes.c_void_p, _sz, _u64p, ctypes.c_void_p]),
    "bnp_decode_g2_batch": (_int, [_int, ctypes.c_void_p, _sz, _u64p, ctypes.c_void_p, _int]),
    "bnp_encode_fq12_batch": (_int, [_u64p, _sz, ctypes.c_void_p]),
    "bnp_decode_fq12_batch": (_int, [ctypes.c_void_p, _sz, _u64p, ctypes.c_void_p]),
    "bnp_eip197_pairing_check": (_int, [ctypes.c_void_p, _sz, ctypes.POINTER(_int)]),
}


class BnpError(RuntimeError):
    pass


_lib = None


def load(path=None):
    """dlopen libbnp.so and declare all prototypes.  Raises if the library or a symbol is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise BnpError("libbnp.so not found at %s - the CUDA extension is not built (run __graft_entry__.build())" % p)
    lib = ctypes.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def check(rc, lib=None):
    if rc != 0:
        lib = lib or load()
        raise BnpError("libbnp: %s (%d) %s" % (lib.bnp_strerror(rc).decode(), rc, lib.bnp_last_error().decode()))


_initialised = False


def init(devices=None):
    """bnp_init on the given CUDA ordinals (default: LOCAL_RANK if set, else 0)."""
    global _initialised
    lib = load()
    if devices is None:
        devices = [int(os.environ.get("LOCAL_RANK", "0"))]
    arr = (_int * len(devices))(*devices)
    check(lib.bnp_init(arr, len(devices)))
    # tuning knobs for experiments (bnp.h::bnp_set_launch_config); unset = library defaults
    tpb, phase = int(os.environ.get("BNP_THREADS", "0")), int(os.environ.get("BNP_PHASE_MODE", "0"))
    if tpb or phase:
        check(lib.bnp_set_launch_config(tpb, phase))
    _initialised = True
    return lib


def lib():
    if not _initialised:
        init()
    return load()
