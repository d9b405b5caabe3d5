"""ctypes binding of libblake3wit.so (include/blake3wit.h).  Fails loudly when the library is absent."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

B3W_OK, B3W_ERR_INVALID, B3W_ERR_CUDA, B3W_ERR_NOMEM, B3W_ERR_DOMAIN, B3W_ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
B3W_CIRCOM_ASSERT = 4
B3W_R1CS_VIOLATION = 7
B3W_NO_ROW = 0xFFFFFFFF
B3W_FLAG_FUSED_CHECK = 1
B3W_EXT_ASSERT = 127
B3W_FLAG_COMPRESSIBLE_RING = 2
B3W_FLAG_PLAIN_RING = 4
B3W_FLAG_REFERENCE_SIBLINGS = 8
B3W_FLAG_BYTE_CHECK = 16
B3W_MEM_COMPRESSIBLE = 1
B3W_MAX_SAMPLES = 1024
B3W_VERSION = 0x000200

EXPORTS = ("b3w_version", "b3w_last_error", "b3w_create", "b3w_destroy", "b3w_circuit_info", "b3w_wtns_header",
           "b3w_input_signal", "b3w_witness_one", "b3w_witness_batch", "b3w_witness_batch_device",
           "b3w_checksum_device", "b3w_calib_fill", "b3w_calib_fill_items", "b3w_calib_fill_bulk", "b3w_host_alloc", "b3w_host_alloc_near", "b3w_host_free",
           "b3w_witness_batch_device_checked", "b3w_r1cs_check_device", "b3w_r1cs_info", "b3w_debug_inject_fault",
           "b3w_nova_chain_size", "b3w_nova_chain", "b3w_debug_set_launch", "b3w_assert_trace",
           "b3w_r1cs_load", "b3w_r1cs_load_file", "b3w_inputs_from_fr", "b3w_witness_batch_fr", "b3w_packed_words", "b3w_witness_batch_packed_device", "b3w_witness_batch_packed", "b3w_unpack_device",
           "b3w_inputs_from_fr_wide", "b3w_witness_batch_wide", "b3w_witness_batch_device_wide", "b3w_assert_trace_fr",
           "b3w_witness_batch_ex", "b3w_witness_batch_fr_ex", "b3w_witness_batch_device_ex", "b3w_last_timing", "b3w_r1cs_program_info",
           "b3w_r1cs_compile_stats", "b3w_r1cs_compile_stats_ex", "b3w_debug_r1cs_program", "b3w_nova_chain_device", "b3w_unpack_host", "b3w_witness_batch_hybrid", "b3w_multi_witness_batch_ex",
           "b3w_debug_set_store_mode", "b3w_debug_side_layout",
           "b3w_device_alloc", "b3w_device_free", "b3w_multi_create", "b3w_multi_destroy", "b3w_multi_size", "b3w_shard_range", "b3w_multi_witness_batch", "b3w_multi_nova_chain")


class B3WError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class Config(C.Structure):
    _fields_ = [("circuit", C.c_uint32), ("device", C.c_int32), ("chunk", C.c_uint32), ("flags", C.c_uint32)]


class BatchExtras(C.Structure):
    """b3w_batch_extras (include/blake3wit.h)"""
    _fields_ = [("sums", C.c_void_p), ("sample_idx", C.c_void_p), ("n_samples", C.c_uint32), ("sample_out", C.c_void_p),
                ("first_bad", C.c_void_p)]


class Timing(C.Structure):
    """b3w_timing (include/blake3wit.h)"""
    _fields_ = [("total_ms", C.c_double), ("kernel_ms", C.c_double), ("host_ms", C.c_double), ("launches", C.c_uint64),
                ("instances", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]


class Info(C.Structure):
    _fields_ = [("witness_size", C.c_uint32), ("n_inputs", C.c_uint32), ("n32", C.c_uint32), ("n_public", C.c_uint32),
                ("version", C.c_uint32 * 3), ("prime", C.c_uint8 * 32)]


def lib_path():
    return os.path.join(_HERE, "libblake3wit.so")


def lib():
    """Load the CUDA library.  There is no fallback: a missing build is an error."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise B3WError(B3W_ERR_CUDA, "%s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)" % p)
    L = C.CDLL(p)
    vp, u64, u32p = C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)
    L.b3w_version.restype = C.c_int
    L.b3w_last_error.restype = C.c_char_p
    L.b3w_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.b3w_destroy.argtypes = [vp]
    L.b3w_destroy.restype = None
    L.b3w_circuit_info.argtypes = [C.c_uint32, C.POINTER(Info)]
    L.b3w_wtns_header.argtypes = [C.c_uint32, vp]
    L.b3w_input_signal.argtypes = [C.c_uint32, C.c_char_p, u32p, u32p]
    L.b3w_witness_one.argtypes = [vp, vp, vp]
    L.b3w_assert_trace.argtypes = [C.c_uint32, vp, C.c_char_p, C.c_size_t]
    L.b3w_witness_batch.argtypes = [vp, vp, u64, vp, vp, vp]
    L.b3w_witness_batch_device.argtypes = [vp, vp, u64, vp, vp, vp, vp]
    L.b3w_checksum_device.argtypes = [vp, vp, u64, vp, vp]
    L.b3w_witness_batch_device_checked.argtypes = [vp, vp, u64, vp, vp, vp, vp, vp]
    L.b3w_r1cs_check_device.argtypes = [vp, vp, u64, vp, vp, vp]
    L.b3w_r1cs_info.argtypes = [C.c_uint32, u32p, u32p]
    L.b3w_debug_inject_fault.argtypes = [vp, C.c_uint32, C.c_uint32]
    L.b3w_debug_set_launch.argtypes = [vp, C.c_int, C.c_uint32]
    L.b3w_nova_chain_size.argtypes = [u64, C.POINTER(u64), C.POINTER(u64)]
    L.b3w_nova_chain.argtypes = [vp, vp, u64, vp, vp, vp, vp, vp, vp]
    L.b3w_calib_fill.argtypes = [vp, vp, u64, vp]
    L.b3w_calib_fill_items.argtypes = [vp, vp, u64, vp]
    L.b3w_calib_fill_bulk.argtypes = [vp, vp, u64, vp]
    L.b3w_r1cs_load.argtypes = [vp, vp, C.c_size_t, u32p]
    L.b3w_r1cs_load_file.argtypes = [vp, C.c_char_p, u32p]
    L.b3w_inputs_from_fr.argtypes = [C.c_uint32, vp, u64, vp]
    L.b3w_witness_batch_fr.argtypes = [vp, vp, u64, vp, vp, vp]
    L.b3w_inputs_from_fr_wide.argtypes = [C.c_uint32, vp, u64, vp, vp, C.POINTER(u64)]
    L.b3w_witness_batch_wide.argtypes = [vp, vp, vp, u64, vp, vp, vp]
    L.b3w_witness_batch_device_wide.argtypes = [vp, vp, vp, u64, vp, vp, vp, vp, vp]
    L.b3w_assert_trace_fr.argtypes = [C.c_uint32, vp, C.c_char_p, C.c_size_t]
    L.b3w_device_alloc.argtypes = [vp, C.c_size_t, C.c_uint32, C.POINTER(vp), u32p]
    L.b3w_device_free.argtypes = [vp, vp]
    L.b3w_packed_words.argtypes = [C.c_uint32, u32p]
    L.b3w_witness_batch_packed_device.argtypes = [vp, vp, u64, vp, vp, vp, vp]
    L.b3w_witness_batch_packed.argtypes = [vp, vp, u64, vp, vp, vp]
    L.b3w_unpack_device.argtypes = [vp, vp, u64, vp, vp]
    L.b3w_multi_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_int32), C.c_uint32, C.POINTER(vp)]
    L.b3w_multi_destroy.argtypes = [vp]
    L.b3w_multi_destroy.restype = None
    L.b3w_multi_size.argtypes = [vp]
    L.b3w_multi_size.restype = C.c_uint32
    L.b3w_shard_range.argtypes = [u64, C.c_uint32, C.c_uint32, C.POINTER(u64), C.POINTER(u64)]
    L.b3w_multi_witness_batch.argtypes = [vp, vp, u64, vp, vp, vp]
    L.b3w_multi_nova_chain.argtypes = [vp, vp, u64, vp, vp, vp, vp, vp, vp]
    ex = C.POINTER(BatchExtras)
    L.b3w_witness_batch_ex.argtypes = [vp, vp, u64, vp, vp, vp, ex]
    L.b3w_witness_batch_fr_ex.argtypes = [vp, vp, u64, vp, vp, vp, ex]
    L.b3w_witness_batch_device_ex.argtypes = [vp, vp, vp, u64, vp, vp, vp, vp, vp, C.c_int, vp]
    L.b3w_multi_witness_batch_ex.argtypes = [vp, vp, u64, vp, vp, vp, ex]
    L.b3w_last_timing.argtypes = [vp, C.POINTER(Timing)]
    L.b3w_r1cs_program_info.argtypes = [vp, u32p, u32p, u32p, u32p]
    L.b3w_r1cs_compile_stats.argtypes = [C.c_uint32, u32p, u32p, u32p, u32p, u32p]
    L.b3w_r1cs_compile_stats_ex.argtypes = [C.c_uint32, u32p, C.c_uint32]
    L.b3w_debug_r1cs_program.argtypes = [vp, C.c_size_t, vp, C.c_uint32, C.c_int, C.c_uint32, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.b3w_nova_chain_device.argtypes = [vp, vp, u64, vp, vp, vp, vp, vp, vp]
    L.b3w_unpack_host.argtypes = [vp, vp, u64, vp, C.c_uint32]
    L.b3w_witness_batch_hybrid.argtypes = [vp, vp, u64, vp, vp, vp, C.c_uint32]
    L.b3w_debug_set_store_mode.argtypes = [vp, C.c_int]
    L.b3w_debug_side_layout.argtypes = [C.c_uint32, C.c_uint32, vp, C.c_uint32, vp, vp]
    L.b3w_host_alloc.argtypes = [C.c_size_t]
    L.b3w_host_alloc.restype = vp
    L.b3w_host_alloc_near.argtypes = [C.c_size_t, C.c_int]
    L.b3w_host_alloc_near.restype = vp
    L.b3w_host_free.argtypes = [vp]
    L.b3w_host_free.restype = None
    _LIB = L
    return L


def check(rc, allow=()):
    if rc != 0 and rc not in allow:
        raise B3WError(rc, "libblake3wit error %d: %s" % (rc, lib().b3w_last_error().decode(errors="replace")))
    return rc
