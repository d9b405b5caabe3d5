"""ref_wasm.py -- ctypes binding of Oracle A (oracle/_ref/libref_*.so).  TEST INFRASTRUCTURE ONLY.

Oracle A is the reference's own circom witness program (the committed .wasm), translated to C by
oracle/wasm2c.py and driven by oracle/wasm_harness.c with the protocol of
/root/reference/blake3_nova_js/witness_calculator.js:131-272.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
VARIANTS = ("compression", "nova_bn_o2", "nova_pasta_o2", "nova_bn_o1")

# Input signals in circuit declaration order (circuits/blake3_compression.circom:172-176,
# circuits/blake3_nova.circom:173-191); sizes are confirmed against getInputSignalSize at load.
INPUT_PLAN = {
    "compression": (("h", 8), ("m", 16), ("t", 2), ("b", 1), ("d", 1)),
    "nova": (("n_blocks", 1), ("block_count", 1), ("h", 8), ("chunk_idx_low", 1), ("chunk_idx_high", 1),
             ("leaf_depth", 1), ("total_depth", 1), ("depth", 1), ("m", 16), ("b", 1)),
}

ERR_TEXT = {1: "Signal not found.\n", 2: "Too many signals set.\n", 3: "Signal already set.\n",
            4: "Assert Failed.\n", 5: "Not enough memory.\n", 6: "Input signal array access exceeds the size.\n"}


def available(variant="compression"):
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_%s.so" % variant))


def fnv1a64(name):
    h = 0xCBF29CE484222325
    for ch in name:
        h ^= ord(ch)
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


class RefWasm:
    """One instance of a reference witness program."""

    def __init__(self, variant):
        path = os.path.join(_HERE, "_ref", "libref_%s.so" % variant)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
        L = C.CDLL(path)
        self.L, self.variant = L, variant
        L.ref_new.restype = C.c_void_p
        L.ref_free.argtypes = [C.c_void_p]
        for f in ("ref_version", "ref_minor_version", "ref_patch_version", "ref_n32", "ref_witness_size", "ref_input_size"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = C.c_uint32
        L.ref_prime.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_input_signal_size.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.ref_input_signal_size.restype = C.c_int32
        L.ref_err_msg.argtypes = [C.c_void_p]
        L.ref_err_msg.restype = C.c_char_p
        L.ref_log_msg.argtypes = [C.c_void_p]
        L.ref_log_msg.restype = C.c_char_p
        L.ref_memory.argtypes = [C.c_void_p]
        L.ref_memory.restype = C.c_void_p
        L.ref_calculate.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 5
        L.ref_calculate.restype = C.c_int
        L.ref_batch.argtypes = [C.c_uint32] + [C.c_void_p] * 4 + [C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_batch.restype = C.c_double
        self.h = L.ref_new()
        self.version = L.ref_version(self.h)
        self.n32 = L.ref_n32(self.h)
        self.witness_size = L.ref_witness_size(self.h)
        self.input_size = L.ref_input_size(self.h)
        limbs = np.zeros(8, np.uint32)
        L.ref_prime(self.h, limbs.ctypes.data)
        self.prime = int.from_bytes(limbs.tobytes(), "little")
        self.plan = INPUT_PLAN["compression" if variant == "compression" else "nova"]
        for name, size in self.plan:
            assert self.input_signal_size(name) == size, (name, size)

    def __del__(self):
        try:
            self.L.ref_free(self.h)
        except Exception:
            pass

    def input_signal_size(self, name):
        h = fnv1a64(name)
        return self.L.ref_input_signal_size(self.h, h >> 32, h & 0xFFFFFFFF)

    def _plan_arrays(self, items):
        """items: list of (name, [int values]) -> hmsb, hlsb, pos, vals(n,8) arrays."""
        hm, hl, ps, vals = [], [], [], []
        for name, vs in items:
            h = fnv1a64(name)
            for i, v in enumerate(vs):
                hm.append(h >> 32)
                hl.append(h & 0xFFFFFFFF)
                ps.append(i)
                v = int(v) % self.prime
                vals.append(np.frombuffer(v.to_bytes(32, "little"), np.uint32))
        return (np.array(hm, np.uint32), np.array(hl, np.uint32), np.array(ps, np.uint32),
                np.ascontiguousarray(np.stack(vals)) if vals else np.zeros((0, 8), np.uint32))

    def calculate(self, inputs):
        """inputs: dict name -> int | (nested) list of ints, any field values (reduced mod p).
        Returns (code, witness bytes as np.uint8[witness_size*32] or None).  code follows the C harness."""
        def flat(a):
            if isinstance(a, (list, tuple, np.ndarray)):
                out = []
                for x in a:
                    out += flat(x)
                return out
            return [a]
        items = [(k, flat(v)) for k, v in inputs.items()]
        hm, hl, ps, vals = self._plan_arrays(items)
        out = np.zeros(self.witness_size * 32, np.uint8)
        rc = self.L.ref_calculate(self.h, len(hm), hm.ctypes.data, hl.ctypes.data, ps.ctypes.data,
                                  vals.ctypes.data, out.ctypes.data)
        return rc, (out if rc == 0 else None)

    def err_msg(self):
        return self.L.ref_err_msg(self.h).decode()

    def log_msg(self):
        return self.L.ref_log_msg(self.h).decode()

    def memory(self, off, n):
        base = self.L.ref_memory(self.h)
        return C.string_at(base + off, n)

    def batch_u32(self, in_u32, nthreads=1, want_out=True):
        """in_u32: (n, n_inputs) uint32 in declaration order.  Returns (witnesses (n, ws*32) u8 | None,
        status int32[n], seconds)."""
        in_u32 = np.ascontiguousarray(in_u32, np.uint32)
        n, k = in_u32.shape
        assert k == self.input_size
        hm, hl, ps = [], [], []
        for name, size in self.plan:
            h = fnv1a64(name)
            for i in range(size):
                hm.append(h >> 32), hl.append(h & 0xFFFFFFFF), ps.append(i)
        hm, hl, ps = (np.array(x, np.uint32) for x in (hm, hl, ps))
        out = np.zeros((n, self.witness_size * 32), np.uint8) if want_out else None
        status = np.zeros(n, np.int32)
        secs = self.L.ref_batch(k, hm.ctypes.data, hl.ctypes.data, ps.ctypes.data, in_u32.ctypes.data, n,
                                out.ctypes.data if want_out else None, status.ctypes.data, int(nthreads))
        return out, status, secs
