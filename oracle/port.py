"""port.py -- ctypes binding of Oracle B (oracle/libb3w_oracle.so, built from circuit_oracle.c).
TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
VARIANT_ID = {"compression": 0, "nova_bn_o2": 1, "nova_pasta_o2": 2, "nova_bn_o1": 3}
_L = None


def lib():
    global _L
    if _L is None:
        so = os.path.join(_HERE, "libb3w_oracle.so")
        srcs = [os.path.join(_HERE, f) for f in ("circuit_oracle.c", "circuit_oracle_nova.inc", "w2s_tables.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.run(["make", "-s", "-C", _HERE, "port"], check=True)
        L = C.CDLL(so)
        for f in ("b3o_witness_size", "b3o_n_inputs", "b3o_n_signals"):
            getattr(L, f).argtypes = [C.c_int]
            getattr(L, f).restype = C.c_uint32
        L.b3o_witness.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.b3o_signals.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.b3o_witness_batch_u32.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _L = L
    return _L


def witness_size(variant):
    return lib().b3o_witness_size(VARIANT_ID[variant])


def witness_fr(variant, values):
    """values: list of n_inputs Python ints (already reduced mod p).  -> (rc, np.uint8[ws*32] | None)"""
    v = VARIANT_ID[variant]
    L = lib()
    assert len(values) == L.b3o_n_inputs(v)
    inp = np.frombuffer(b"".join(int(x).to_bytes(32, "little") for x in values), np.uint8).copy()
    out = np.zeros(L.b3o_witness_size(v) * 32, np.uint8)
    rc = L.b3o_witness(v, inp.ctypes.data, out.ctypes.data)
    return rc, (out if rc == 0 else None)


def signals_fr(variant, values):
    v = VARIANT_ID[variant]
    L = lib()
    inp = np.frombuffer(b"".join(int(x).to_bytes(32, "little") for x in values), np.uint8).copy()
    out = np.zeros(L.b3o_n_signals(v) * 32, np.uint8)
    rc = L.b3o_signals(v, inp.ctypes.data, out.ctypes.data)
    return rc, out


def witness_batch(variant, rows, nthreads=None, want="witness"):
    """rows: (n, n_inputs) uint32.  want = "witness" -> (n, ws*32) u8;  "sums" -> u64[n] checksums
    (same definition as b3w_checksum_device);  "sums+status" -> (sums, status);  "both" -> (witness, sums, status)."""
    v = VARIANT_ID[variant]
    L = lib()
    rows = np.ascontiguousarray(rows, np.uint32)
    n = rows.shape[0]
    assert rows.shape[1] == L.b3o_n_inputs(v)
    ws = L.b3o_witness_size(v)
    out = np.zeros((n, ws * 32), np.uint8) if want in ("witness", "both") else None
    sums = np.zeros(n, np.uint64) if want in ("sums", "sums+status", "both") else None
    status = np.zeros(n, np.int32)
    L.b3o_witness_batch_u32(v, rows.ctypes.data, n, out.ctypes.data if out is not None else None,
                            sums.ctypes.data if sums is not None else None, status.ctypes.data,
                            int(nthreads or os.cpu_count() or 1))
    if want == "witness":
        assert (status == 0).all(), "oracle: assert failed for some instance"
        return out
    if want == "sums":
        return sums
    if want == "sums+status":
        return sums, status
    return out, sums, status
