"""witness_calculator.py -- host-side mirror of the reference's witness_calculator.js, backed by
libblake3wit.so (CUDA, sm_100a) instead of the circom wasm program.

Reference interface mirrored (paths under /root/reference):
    builder(code, options)                  blake3_nova_js/witness_calculator.js:1-106
    WitnessCalculator fields                :108-125  (version, n32, prime, witnessSize, sanityCheck)
    circom_version()                        :127-129
    _doCalculateWitness(input, sanityCheck) :131-169  (same checks, same error messages)
    calculateWitness / calculateBinWitness / calculateWTNSBin   :171-272
plus the NEW batched entry point calculateWitnessBatch().  Node is not available in this image, so the
N-API addon of INTEGRATION.md cannot be built here; this module is the executable host layer and the
parity tests are written against it the way the reference's tests use witness_calculator.js.

Differences that are deliberate and loud:
  * `code` (the .wasm bytes) only selects the circuit (by sha256); an unknown wasm raises -- there is no
    WebAssembly engine here to fall back to.
  * every input the reference takes is taken (any field element: the circuit's own constraints decide between a
    witness and "Assert Failed.", as in the wasm).  u32 inputs -- all that the reference's drivers produce -- run on
    the hot kernels; field-valued ones go through b3w_witness_batch_fr (u32 row arrays stay u32: _row() raises
    B3WError(B3W_ERR_DOMAIN) for anything else).
  * methods are plain (synchronous) functions.
"""
import ctypes as C
import hashlib

import numpy as np

from . import _lib
from ._lib import B3WError

# sha256 of the reference's committed witness programs -> (circuit id, name)
CIRCUITS = {
    "6faf23ddfd697bbb7e8e922577589c2c06486258968a5a14f96fb5a16091b142": (0, "blake3_compression"),
    "020bd11f289864c54c7d02cd05723dcf8323e31fa5c77d8700c618232685978e": (1, "blake3_nova (bn128, O2)"),
    "b982f960ebbfcabe957fe13857ea47adfeee30e18fbe05474e9b982eab187f46": (2, "blake3_nova_pasta (vesta prime, O2)"),
    "8d6317b72eab34d34e12dfd7bd310dce40f4190768669772f992a9510c441fca": (3, "blake3_nova (bn128, O1, circomkit)"),
}
CIRCUIT_IDS = {"blake3_compression": 0, "blake3_nova": 1, "blake3_nova_pasta": 2, "blake3_nova_o1": 3}


def circuit_from_wasm(code):
    h = hashlib.sha256(bytes(code)).hexdigest()
    if h not in CIRCUITS:
        raise B3WError(_lib.B3W_ERR_UNSUPPORTED,
                       "unknown witness program (sha256 %s...): only the reference's BLAKE3 circuits are built in" % h[:16])
    return CIRCUITS[h][0]


def builder(code, options=None, device=-1, chunk=0, lazy=False, fused_check=False, compressible_ring=True, reference_siblings=False,
            byte_check=False):
    """builder(code, options) -> WitnessCalculator   (witness_calculator.js:1).
    `code`: bytes of one of the reference's .wasm files, or a circuit name / id.
    lazy=True defers the creation of the GPU context to the first witness call (host-logic tests).
    fused_check=True makes every batch call also run the fused on-device R1CS check (status 7 = violation).
    byte_check=True (B3W_FLAG_BYTE_CHECK) re-reads every chunk's witnesses from the HBM ring and evaluates every row of the
    constraint system on those bytes before they leave the GPU (the consumer-side check of rust_fold/src/utils.rs:78-85).
    compressible_ring=False keeps the internal HBM ring of the host-buffer calls in ordinary memory (B3W_FLAG_PLAIN_RING); the
    default -- here, in C and in the N-API addon -- is compressible memory with a silent fall-back to ordinary memory.
    reference_siblings=True makes novaChain pick parent-step siblings by the reference's rule (rust_fold/src/blake3_hash.rs:60-78)."""
    if isinstance(code, int):
        cid = code
    elif isinstance(code, str):
        cid = CIRCUIT_IDS[code]
    else:
        cid = circuit_from_wasm(code)
    return WitnessCalculator(cid, options or {}, device=device, chunk=chunk, lazy=lazy, fused_check=fused_check,
                             compressible_ring=compressible_ring, reference_siblings=reference_siblings, byte_check=byte_check)


def _flat_array(a):
    """flatArray (witness_calculator.js:303-317)"""
    res = []

    def fill(x):
        if isinstance(x, (list, tuple, np.ndarray)):
            for y in x:
                fill(y)
        else:
            res.append(x)
    fill(a)
    return res


def _to_bigint(n):
    """BigInt(n) for the value types JSON / JS callers pass."""
    if isinstance(n, (bool, np.bool_)):
        return int(n)
    if isinstance(n, (int, np.integer)):
        return int(n)
    if isinstance(n, str):
        s = n.strip()
        return int(s, 0) if s.lower().startswith(("0x", "-0x", "0b", "0o")) else int(s)
    if isinstance(n, float) and n == int(n):
        return int(n)
    raise TypeError("Cannot convert %r to a BigInt" % (n,))


def pinned_array(shape, dtype):
    """numpy array in pinned host memory (b3w_host_alloc): copies to / from the GPU run at full PCIe rate and overlap
    with kernels.  Freed when the array is garbage collected."""
    import weakref
    L = _lib.lib()
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    p = L.b3w_host_alloc(max(n, 1))
    if not p:
        raise B3WError(_lib.B3W_ERR_NOMEM, L.b3w_last_error().decode(errors="replace"))
    buf = (C.c_uint8 * max(n, 1)).from_address(p)
    arr = np.frombuffer(buf, dt, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, L.b3w_host_free, p)
    return arr


PINNED_FROM = 32 << 20     # witness buffers of at least this many bytes are allocated pinned (full PCIe rate, overlap)


def _witness_buffer(n, row_bytes):
    if n * row_bytes >= PINNED_FROM:
        try:
            return pinned_array((n, row_bytes), np.uint8)
        except B3WError:
            pass                                   # locked-memory limit: a pageable buffer still works, only slower
    return np.empty((n, row_bytes), np.uint8)


def _nova_chain(L, call, handle, witness_size, data, want_witness):
    data = bytes(data)
    nc, ns = C.c_uint64(), C.c_uint64()
    _lib.check(L.b3w_nova_chain_size(len(data), C.byref(nc), C.byref(ns)))
    nc, ns = nc.value, ns.value
    buf = np.frombuffer(data, np.uint8) if data else np.zeros(1, np.uint8)
    rows = pinned_array((ns, 32), np.uint32)
    step_off = np.zeros(nc + 1, np.uint64)
    status = pinned_array((ns,), np.uint8)
    pub = pinned_array((ns, 15), np.uint32)
    out = pinned_array((ns, witness_size * 32), np.uint8) if want_witness else None
    root = np.zeros(32, np.uint8)
    _lib.check(call(handle, buf.ctypes.data, len(data), out.ctypes.data if want_witness else None, status.ctypes.data,
                    pub.ctypes.data, rows.ctypes.data, step_off.ctypes.data, root.ctypes.data))
    return {"n_chunks": nc, "total_steps": ns, "step_off": step_off, "rows": rows, "status": status, "pub": pub,
            "witness": out, "root": root.tobytes()}


class WitnessCalculator:
    def __init__(self, circuit, sanity_check, device=-1, chunk=0, lazy=False, fused_check=False, compressible_ring=True,
                 reference_siblings=False, byte_check=False):
        L = _lib.lib()
        self._L, self._ctx = L, None
        self._cfg = _lib.Config(circuit, device, chunk, (_lib.B3W_FLAG_FUSED_CHECK if fused_check else 0) |
                                (0 if compressible_ring else _lib.B3W_FLAG_PLAIN_RING) |
                                (_lib.B3W_FLAG_REFERENCE_SIBLINGS if reference_siblings else 0) |
                                (_lib.B3W_FLAG_BYTE_CHECK if byte_check else 0))
        info = _lib.Info()
        _lib.check(L.b3w_circuit_info(circuit, C.byref(info)))
        if not lazy:
            self._h                               # instantiate now, like WebAssembly.instantiate in builder()
        self.instance = self                      # the reference exposes the wasm instance here
        self.circuit = circuit
        self.version = info.version[0]
        self.n32 = info.n32
        self.prime = int.from_bytes(bytes(info.prime), "little")
        self.witnessSize = info.witness_size
        self.nInputs = info.n_inputs
        self.nPublic = info.n_public
        self.sanityCheck = sanity_check

    @property
    def _h(self):
        """The GPU context (created on first use)."""
        if self._ctx is None:
            h = C.c_void_p()
            _lib.check(self._L.b3w_create(C.byref(self._cfg), C.byref(h)))
            self._ctx = h
        return self._ctx

    def close(self):
        if getattr(self, "_ctx", None):
            self._L.b3w_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def circom_version(self):
        return self.version

    # ---- input normalisation: _doCalculateWitness (witness_calculator.js:131-169) ----
    def _input_signal_size(self, name):
        off, size = C.c_uint32(), C.c_uint32()
        rc = self._L.b3w_input_signal(self.circuit, name.encode(), C.byref(off), C.byref(size))
        if rc != 0:
            return 0, 0          # the wasm's getInputSignalSize returns 0 for an unknown name
        return off.value, size.value

    def _values(self, inp):
        """The checks of _doCalculateWitness (:138-168) -> the nInputs values in declaration order, each BigInt(n) % prime
        made non-negative (normalize, :319-323)."""
        vals = [0] * self.nInputs
        input_counter = 0
        for k in inp.keys():
            f_arr = _flat_array(inp[k])
            off, signal_size = self._input_signal_size(k)
            if signal_size < 0:
                raise RuntimeError("Signal %s not found\n" % k)
            if len(f_arr) < signal_size:
                raise RuntimeError("Not enough values for input signal %s\n" % k)
            if len(f_arr) > signal_size:
                raise RuntimeError("Too many values for input signal %s\n" % k)
            for i, v in enumerate(f_arr):
                vals[off + i] = _to_bigint(v) % self.prime
                input_counter += 1
        if input_counter < self.nInputs:
            raise RuntimeError("Not all inputs have been set. Only %d out of %d" % (input_counter, self.nInputs))
        return vals

    def _row(self, inp):
        """-> the u32 input row; values outside [0, 2^32) raise B3WError(B3W_ERR_DOMAIN)."""
        vals = self._values(inp)
        for idx, x in enumerate(vals):
            if x >> 32:
                raise B3WError(_lib.B3W_ERR_DOMAIN, "input %s = %d is outside the u32 domain" % (self._signal_at(idx), x))
        return np.array(vals, np.uint32)

    def _signal_at(self, idx):
        for name in ("h", "m", "t", "b", "d", "n_blocks", "block_count", "chunk_idx_low", "chunk_idx_high", "leaf_depth",
                     "total_depth", "depth"):
            off, size = self._input_signal_size(name)
            if size and off <= idx < off + size:
                return "%s[%d]" % (name, idx - off)
        return "#%d" % idx

    @staticmethod
    def _fr_bytes(vals):
        return np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vals), np.uint8).copy()

    def _do_calculate(self, inp):
        vals = self._values(inp)
        u32 = all(x >> 32 == 0 for x in vals)
        if self.circuit != 0:
            # the nova circuits execute log("D_FLAGS: ", D_FLAGS) once per witness (circuits/blake3_nova.circom:166);
            # witness_calculator.js:44-61 prints it with console.log.  The batched entry point stays silent.
            print("D_FLAGS:  0")
        out = np.empty(self.witnessSize * 32, np.uint8)
        if u32:
            row = np.array(vals, np.uint32)
            rc = self._L.b3w_witness_one(self._h, row.ctypes.data, out.ctypes.data)
            trace = lambda: self.assertTrace(row)
        else:
            # any field element is an input (only the circuit's own constraints decide, as in the reference)
            fr = self._fr_bytes(vals)
            status = np.zeros(1, np.uint8)
            _lib.check(self._L.b3w_witness_batch_fr(self._h, fr.ctypes.data, 1, out.ctypes.data, status.ctypes.data, None))
            rc = int(status[0])
            trace = lambda: self.assertTraceFr(vals)
        if rc == _lib.B3W_CIRCOM_ASSERT:
            # witness_calculator.js:21-39,159-162: Error("Assert Failed.\n" + the printErrorMessage lines), re-wrapped
            raise RuntimeError("Error: Assert Failed.\n" + trace())
        _lib.check(rc)
        return out

    def assertTrace(self, row):
        """The per-template error trace of the reference for an input row that asserts ("" if it does not)."""
        row = np.ascontiguousarray(row, np.uint32)
        buf = C.create_string_buffer(1024)
        rc = self._L.b3w_assert_trace(self.circuit, row.ctypes.data, buf, len(buf))
        if rc not in (0, _lib.B3W_CIRCOM_ASSERT):
            _lib.check(rc)
        return buf.value.decode()

    def assertTraceFr(self, vals):
        """ditto for nInputs field elements (ints, already reduced or not)."""
        fr = self._fr_bytes([int(v) % self.prime for v in vals])
        buf = C.create_string_buffer(1024)
        rc = self._L.b3w_assert_trace_fr(self.circuit, fr.ctypes.data, buf, len(buf))
        if rc not in (0, _lib.B3W_CIRCOM_ASSERT):
            _lib.check(rc)
        return buf.value.decode()

    # ---- the three reference read-outs ----
    def calculateWitness(self, inp, sanityCheck=0):
        """-> list of witnessSize Python ints (the reference returns BigInt[])."""
        b = self._do_calculate(inp).tobytes()
        return [int.from_bytes(b[32 * i:32 * i + 32], "little") for i in range(self.witnessSize)]

    def calculateBinWitness(self, inp, sanityCheck=0):
        """-> np.uint8[witnessSize*32], little-endian limbs."""
        return self._do_calculate(inp)

    def calculateWTNSBin(self, inp, sanityCheck=0):
        """-> np.uint8[76 + witnessSize*32]: the .wtns file image."""
        body = self._do_calculate(inp)
        hdr = np.empty(76, np.uint8)
        _lib.check(self._L.b3w_wtns_header(self.circuit, hdr.ctypes.data))
        return np.concatenate([hdr, body])

    # ---- NEW: the consumer's side -- are these witnesses valid? ----
    def checkWitnesses(self, witness):
        """witness: (n, witnessSize*32) u8 in HOST memory (.wtns bodies, e.g. read back from files) -> (status u8[n], first_bad
        u32[n]): every row of the circuit's constraint system evaluated on the bytes by the stand-alone GPU checker
        (b3w_r1cs_check_device) -- what circom_tester's expectPass (test/blake3_hash.test.ts:36,57), `snarkjs wtns check` and
        bellpepper's enforce (rust_fold/src/utils.rs:78-85) decide.  status 0 = all rows hold, 7 = B3W_R1CS_VIOLATION with
        first_bad = the violated row (B3W_NO_ROW otherwise)."""
        import torch
        w = np.ascontiguousarray(witness, np.uint8).reshape(-1, self.witnessSize * 32)
        if not w.flags.writeable:                         # a view of a bytes object: torch wants memory it may write to
            w = w.copy()
        n = w.shape[0]
        with torch.cuda.device(self._cfg.device if self._cfg.device >= 0 else torch.cuda.current_device()):
            self._h
            d_w = torch.from_numpy(w).cuda()
            d_st = torch.full((n,), 255, dtype=torch.uint8, device="cuda")
            d_bad = torch.zeros(n, dtype=torch.int32, device="cuda")
            self.r1cs_check_device(d_w.data_ptr(), n, d_st.data_ptr(), d_bad.data_ptr(), torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            return d_st.cpu().numpy(), d_bad.cpu().numpy().view(np.uint32)

    def wtnsBody(self, buf):
        """a .wtns image (what calculateWTNSBin / the CLI wrote) -> its witness section, after checking that the container is
        THIS circuit's: same 76-byte header as b3w_wtns_header (prime, witnessSize).  Host-only."""
        from .wtns import parse_wtns, WtnsError
        w = parse_wtns(buf)
        if w["prime"] != self.prime:
            raise WtnsError("the file's prime is not this circuit's (%d bits vs %d)" % (w["prime"].bit_length(), self.prime.bit_length()))
        if w["n8"] != 32 or w["n_witness"] != self.witnessSize:
            raise WtnsError("the file holds %d witness values, this circuit has %d" % (w["n_witness"], self.witnessSize))
        return w["body"]

    def checkWTNSBin(self, buf):
        """-> (ok, first_bad): the .wtns image's witness checked against every constraint on the GPU (see checkWitnesses)."""
        status, bad = self.checkWitnesses(self.wtnsBody(buf)[None, :])
        return bool(status[0] == 0), int(bad[0])

    # ---- NEW: batched entry point ----
    def _extras(self, n, sums, samples, first_bad):
        """-> (BatchExtras or None, dict of the arrays it points into)"""
        if not (sums or first_bad or samples is not None):
            return None, {}
        ex, keep = _lib.BatchExtras(), {}
        if sums:
            keep["sums"] = np.zeros(n, np.uint64)
            ex.sums = keep["sums"].ctypes.data
        if first_bad:
            keep["first_bad"] = np.zeros(n, np.uint32)
            ex.first_bad = keep["first_bad"].ctypes.data
        if samples is not None:
            idx = np.ascontiguousarray(samples, np.uint64)
            keep["sample_idx"] = idx
            keep["samples"] = np.zeros((idx.size, self.witnessSize * 32), np.uint8)
            ex.sample_idx, ex.n_samples, ex.sample_out = idx.ctypes.data, idx.size, keep["samples"].ctypes.data
        return ex, keep

    def calculateWitnessBatch(self, inputs, want_witness=True, out=None, sums=False, samples=None, first_bad=False):
        """inputs: list of input objects (as for calculateWitness) or an (n, nInputs) uint32 array in
        circuit declaration order.  Returns dict(witness=(n, witnessSize*32) u8 | None, status=u8[n],
        pub=(n, nPublic) u32).  Extra results (b3w_batch_extras): sums=True adds "sums" (u64[n], the witness checksum the
        expansion warps compute from what they store), samples=<indices> adds "samples" (their full witnesses, also when
        want_witness=False), first_bad=True adds "first_bad" (fused check: smallest violated row)."""
        fr = None
        if isinstance(inputs, np.ndarray):
            rows = np.ascontiguousarray(inputs, np.uint32)
            if rows.ndim != 2 or rows.shape[1] != self.nInputs:
                raise ValueError("expected an (n, %d) uint32 array" % self.nInputs)
        else:
            vals = [self._values(i) for i in inputs]
            if any(x >> 32 for v in vals for x in v):
                fr = self._fr_bytes([x for v in vals for x in v])          # field-element inputs: b3w_witness_batch_fr
                rows = np.zeros((len(vals), self.nInputs), np.uint32)
            else:
                rows = np.stack([self._row(i) for i in inputs]) if len(inputs) else np.zeros((0, self.nInputs), np.uint32)
        n = rows.shape[0]
        if want_witness and out is None:
            out = _witness_buffer(n, self.witnessSize * 32)
        status = np.zeros(n, np.uint8)
        pub = np.zeros((n, self.nPublic), np.uint32)
        ex, keep = self._extras(n, sums, samples, first_bad)
        call = self._L.b3w_witness_batch_fr_ex if fr is not None else self._L.b3w_witness_batch_ex
        src = fr if fr is not None else rows
        _lib.check(call(self._h, src.ctypes.data, n, out.ctypes.data if want_witness else None, status.ctypes.data, pub.ctypes.data,
                        C.byref(ex) if ex is not None else None))
        res = {"witness": out if want_witness else None, "status": status, "pub": pub}
        res.update({k: v for k, v in keep.items() if k != "sample_idx"})
        return res

    def calculateWitnessBatchFr(self, values, want_witness=True, sums=False, samples=None, first_bad=False):
        """values: (n, nInputs) Python ints / an (n, nInputs, 32) uint8 array of little-endian field elements.
        Every input the reference accepts.  The rows are converted on the device; u32 instances run on the hot kernels,
        only the instances that hold a field-valued input take the general path (into the same outputs)."""
        if isinstance(values, np.ndarray) and values.dtype == np.uint8:
            fr = np.ascontiguousarray(values).reshape(-1)
            n = values.shape[0]
        else:
            n = len(values)
            fr = self._fr_bytes([int(x) % self.prime for v in values for x in v]) if n else np.zeros(0, np.uint8)
        out = _witness_buffer(n, self.witnessSize * 32) if want_witness else None
        status = np.zeros(n, np.uint8)
        pub = np.zeros((n, self.nPublic), np.uint32)
        ex, keep = self._extras(n, sums, samples, first_bad)
        _lib.check(self._L.b3w_witness_batch_fr_ex(self._h, fr.ctypes.data, n, out.ctypes.data if want_witness else None,
                                                   status.ctypes.data, pub.ctypes.data, C.byref(ex) if ex is not None else None))
        res = {"witness": out, "status": status, "pub": pub}
        res.update({k: v for k, v in keep.items() if k != "sample_idx"})
        return res

    def lastTiming(self):
        """b3w_last_timing: dict(total_ms, kernel_ms, host_ms, launches, instances, h2d_bytes, d2h_bytes) of the last host-buffer call"""
        t = _lib.Timing()
        _lib.check(self._L.b3w_last_timing(self._h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in _lib.Timing._fields_}

    # ---- NEW: compact witnesses (the per-instance trace; ~200x smaller than the .wtns body) ----
    @property
    def packedWords(self):
        w = C.c_uint32()
        _lib.check(self._L.b3w_packed_words(self.circuit, C.byref(w)))
        return w.value

    def calculateWitnessBatchPacked(self, inputs):
        """As calculateWitnessBatch, but returns dict(packed=(n, packedWords) u32, status, pub): the witnesses in compact
        form.  unpackWitnesses() (GPU) turns any subset of them into .wtns bodies."""
        rows = np.ascontiguousarray(inputs, np.uint32) if isinstance(inputs, np.ndarray) else \
            (np.stack([self._row(i) for i in inputs]) if len(inputs) else np.zeros((0, self.nInputs), np.uint32))
        if rows.ndim != 2 or rows.shape[1] != self.nInputs:
            raise ValueError("expected an (n, %d) uint32 array" % self.nInputs)
        n = rows.shape[0]
        packed = np.zeros((n, self.packedWords), np.uint32)
        status = np.zeros(n, np.uint8)
        pub = np.zeros((n, self.nPublic), np.uint32)
        _lib.check(self._L.b3w_witness_batch_packed(self._h, rows.ctypes.data, n, packed.ctypes.data, status.ctypes.data,
                                                    pub.ctypes.data))
        return {"packed": packed, "status": status, "pub": pub}

    def unpackWitnesses(self, packed):
        """(n, packedWords) u32 -> (n, witnessSize*32) u8: the lazy .wtns-body export, expanded on the GPU."""
        import torch
        packed = np.ascontiguousarray(packed, np.uint32)
        n = packed.shape[0]
        with torch.cuda.device(self._cfg.device if self._cfg.device >= 0 else torch.cuda.current_device()):
            self._h
            d_p = torch.from_numpy(packed.view(np.int32)).cuda()
            d_o = torch.empty((n, self.witnessSize * 32), dtype=torch.uint8, device="cuda")
            self.unpack_device(d_p.data_ptr(), n, d_o.data_ptr(), torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            return d_o.cpu().numpy()

    def unpackWitnessesHost(self, packed, threads=0):
        """(n, packedWords) u32 -> (n, witnessSize*32) u8, expanded on the HOST (b3w_unpack_host): the same bytes as
        unpackWitnesses(); needs no GPU work."""
        packed = np.ascontiguousarray(packed, np.uint32)
        n = packed.shape[0]
        out = np.empty((n, self.witnessSize * 32), np.uint8)
        _lib.check(self._L.b3w_unpack_host(self._h, packed.ctypes.data, n, out.ctypes.data, threads))
        return out

    def calculateWitnessBatchHybrid(self, rows, out=None, threads=0):
        """b3w_witness_batch_hybrid: every .wtns body in host memory, but only the packed records cross PCIe; the expansion
        runs on host threads.  rows: (n, nInputs) u32."""
        rows = np.ascontiguousarray(rows, np.uint32)
        n = rows.shape[0]
        if out is None:
            out = np.empty((n, self.witnessSize * 32), np.uint8)
        status = np.zeros(n, np.uint8)
        pub = np.zeros((n, self.nPublic), np.uint32)
        _lib.check(self._L.b3w_witness_batch_hybrid(self._h, rows.ctypes.data, n, out.ctypes.data, status.ctypes.data, pub.ctypes.data, threads))
        return {"witness": out, "status": status, "pub": pub}

    def witness_batch_packed_device(self, d_in, n, d_packed, d_status=0, d_pub=0, stream=0):
        _lib.check(self._L.b3w_witness_batch_packed_device(self._h, d_in, n, d_packed, d_status or None, d_pub or None,
                                                           stream or None))

    def unpack_device(self, d_packed, n, d_out, stream=0):
        _lib.check(self._L.b3w_unpack_device(self._h, d_packed, n, d_out, stream or None))

    # ---- NEW: all Nova step witnesses of a file (the batched form of rust_fold's prove_step loop) ----
    def novaChain(self, data, want_witness=False):
        """data: bytes.  Returns dict(n_chunks, total_steps, step_off=u64[n_chunks+1], rows=(steps,32) u32 step inputs,
        status=u8[steps], pub=(steps,15) u32 = z_{i+1}, witness=(steps, witnessSize*32) u8 | None, root=32 bytes)."""
        return _nova_chain(self._L, self._L.b3w_nova_chain, self._h, self.witnessSize, data, want_witness)

    def novaChainDevice(self, data, d_out, d_status=0, d_pub=0, d_rows=0):
        """b3w_nova_chain_device: the step witnesses stay in device memory (caller-supplied device pointers; d_out must hold
        total_steps * witnessSize * 32 bytes).  Returns dict(n_chunks, total_steps, step_off, root)."""
        data = bytes(data)
        nc, ns = C.c_uint64(), C.c_uint64()
        _lib.check(self._L.b3w_nova_chain_size(len(data), C.byref(nc), C.byref(ns)))
        buf = np.frombuffer(data, np.uint8) if data else np.zeros(1, np.uint8)
        step_off = np.zeros(nc.value + 1, np.uint64)
        root = np.zeros(32, np.uint8)
        _lib.check(self._L.b3w_nova_chain_device(self._h, buf.ctypes.data, len(data), d_out, d_status or None, d_pub or None,
                                                 d_rows or None, step_off.ctypes.data, root.ctypes.data))
        return {"n_chunks": nc.value, "total_steps": ns.value, "step_off": step_off, "root": root.tobytes()}

    # ---- NEW: compressible device memory for witness buffers (b3w_device_alloc) ----
    def device_alloc(self, nbytes, compressible=True):
        """-> (device pointer, granted): device memory of this context; `granted` tells whether the driver made it
        compressible (Blackwell compresses such lines between L2 and HBM: witnesses are mostly zero bytes)."""
        p, g = C.c_void_p(), C.c_uint32()
        _lib.check(self._L.b3w_device_alloc(self._h, int(nbytes), _lib.B3W_MEM_COMPRESSIBLE if compressible else 0, C.byref(p), C.byref(g)))
        return p.value, bool(g.value)

    def device_free(self, ptr):
        _lib.check(self._L.b3w_device_free(self._h, ptr))

    # ---- device-pointer plumbing used by bench.py / tests (torch supplies memory and streams) ----
    def witness_batch_device(self, d_in, n, d_out, d_status=0, d_pub=0, stream=0):
        _lib.check(self._L.b3w_witness_batch_device(self._h, d_in, n, d_out, d_status or None, d_pub or None,
                                                    stream or None))

    def witness_batch_device_checked(self, d_in, n, d_out, d_status=0, d_pub=0, d_first_bad=0, stream=0):
        """generation + fused R1CS check (rows evaluated on the shared-memory trace)"""
        _lib.check(self._L.b3w_witness_batch_device_checked(self._h, d_in, n, d_out, d_status or None, d_pub or None,
                                                            d_first_bad or None, stream or None))

    def witness_batch_device_ex(self, d_in, n, d_out, d_status=0, d_pub=0, d_first_bad=0, d_sums=0, d_m_ext=0, check=False, stream=0):
        """b3w_witness_batch_device_ex: everything at once (wide message words, fused check, per-instance checksums)"""
        _lib.check(self._L.b3w_witness_batch_device_ex(self._h, d_in, d_m_ext or None, n, d_out, d_status or None, d_pub or None,
                                                       d_first_bad or None, d_sums or None, 1 if check else 0, stream or None))

    def witness_batch_device_wide(self, d_in, d_m_ext, n, d_out, d_status=0, d_pub=0, d_first_bad=0, stream=0):
        """blake3_compression with wide message words (m = m_ext * 2^32 + row word); d_first_bad != 0 adds the fused check"""
        _lib.check(self._L.b3w_witness_batch_device_wide(self._h, d_in, d_m_ext, n, d_out, d_status or None, d_pub or None,
                                                         d_first_bad or None, stream or None))

    def r1cs_check_device(self, d_wit, n, d_status=0, d_first_bad=0, stream=0):
        """stand-alone R1CS check of witnesses resident in device memory (all four circuits have a built-in system)"""
        _lib.check(self._L.b3w_r1cs_check_device(self._h, d_wit, n, d_status or None, d_first_bad or None, stream or None))

    def r1cs_load(self, r1cs):
        """use the constraint system of an iden3 .r1cs file (bytes or a path) for r1cs_check_device; returns its row count"""
        n = C.c_uint32()
        if isinstance(r1cs, (bytes, bytearray, memoryview, np.ndarray)):
            buf = np.frombuffer(bytes(r1cs), np.uint8)
            _lib.check(self._L.b3w_r1cs_load(self._h, buf.ctypes.data, buf.size, C.byref(n)))
        else:
            _lib.check(self._L.b3w_r1cs_load_file(self._h, str(r1cs).encode(), C.byref(n)))
        return n.value

    def r1cs_program_info(self):
        """-> dict(rows, compiled, xor_runs, tiles): how the context's constraint system was split for the stand-alone check"""
        v = [C.c_uint32() for _ in range(4)]
        _lib.check(self._L.b3w_r1cs_program_info(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("rows", "compiled", "xor_runs", "tiles"), (x.value for x in v)))

    def r1cs_info(self):
        rows, terms = C.c_uint32(), C.c_uint32()
        _lib.check(self._L.b3w_r1cs_info(self.circuit, C.byref(rows), C.byref(terms)))
        return rows.value, terms.value

    def inject_fault(self, trace_word=0xFFFFFFFF, xor_mask=0):
        _lib.check(self._L.b3w_debug_inject_fault(self._h, trace_word, xor_mask))

    def set_launch(self, ctas_per_sm=0, parts=0):
        """tuning hook: cap on resident CTAs per SM, work items per instance (0 = defaults)"""
        _lib.check(self._L.b3w_debug_set_launch(self._h, int(ctas_per_sm), int(parts)))

    def set_store_mode(self, mode=0):
        """tuning hook: 0 = direct 256-bit stores, 1 = shared-memory tiles + TMA bulk stores"""
        _lib.check(self._L.b3w_debug_set_store_mode(self._h, int(mode)))

    def checksum_device(self, d_wit, n, d_sums, stream=0):
        _lib.check(self._L.b3w_checksum_device(self._h, d_wit, n, d_sums, stream or None))

    def calib_fill(self, d_buf, nbytes, stream=0, items=False):
        """pure-store calibration; items=True uses the witness kernels' own store stream (32 KiB dynamic work items)"""
        f = self._L.b3w_calib_fill_bulk if items == "bulk" else self._L.b3w_calib_fill_items if items else self._L.b3w_calib_fill
        _lib.check(f(self._h, d_buf, nbytes, stream or None))


class MultiGpuCalculator:
    """NEW: the batched entry point over several GPUs of one node (b3w_multi_*, include/blake3wit.h): contiguous index
    ranges, one context and host thread per device, no collective.  devices=None takes every visible device."""

    def __init__(self, circuit, devices=None, chunk=0, fused_check=False, compressible_ring=True, reference_siblings=False,
                 byte_check=False):
        L = _lib.lib()
        self._L = L
        cid = circuit if isinstance(circuit, int) else (CIRCUIT_IDS[circuit] if isinstance(circuit, str) else circuit_from_wasm(circuit))
        info = _lib.Info()
        _lib.check(L.b3w_circuit_info(cid, C.byref(info)))
        self.circuit, self.witnessSize, self.nInputs, self.nPublic = cid, info.witness_size, info.n_inputs, info.n_public
        cfg = _lib.Config(cid, -1, chunk, (_lib.B3W_FLAG_FUSED_CHECK if fused_check else 0) |
                          (0 if compressible_ring else _lib.B3W_FLAG_PLAIN_RING) |
                          (_lib.B3W_FLAG_REFERENCE_SIBLINGS if reference_siblings else 0) |
                          (_lib.B3W_FLAG_BYTE_CHECK if byte_check else 0))
        devs = (C.c_int32 * len(devices))(*devices) if devices else None
        h = C.c_void_p()
        _lib.check(L.b3w_multi_create(C.byref(cfg), devs, len(devices) if devices else 0, C.byref(h)))
        self._m = h
        self.nDevices = L.b3w_multi_size(h)

    def close(self):
        if getattr(self, "_m", None):
            self._L.b3w_multi_destroy(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shard(self, n, g):
        """(first, count) of device slot g for a batch of n"""
        first, count = C.c_uint64(), C.c_uint64()
        _lib.check(self._L.b3w_shard_range(n, g, self.nDevices, C.byref(first), C.byref(count)))
        return first.value, count.value

    def calculateWitnessBatch(self, rows, want_witness=True, out=None, sums=False, samples=None, first_bad=False):
        rows = np.ascontiguousarray(rows, np.uint32)
        if rows.ndim != 2 or rows.shape[1] != self.nInputs:
            raise ValueError("expected an (n, %d) uint32 array" % self.nInputs)
        n = rows.shape[0]
        if want_witness and out is None:
            out = _witness_buffer(n, self.witnessSize * 32)
        status = np.zeros(n, np.uint8)
        pub = np.zeros((n, self.nPublic), np.uint32)
        ex, keep = WitnessCalculator._extras(self, n, sums, samples, first_bad)
        _lib.check(self._L.b3w_multi_witness_batch_ex(self._m, rows.ctypes.data, n, out.ctypes.data if want_witness else None,
                                                      status.ctypes.data, pub.ctypes.data, C.byref(ex) if ex is not None else None))
        res = {"witness": out if want_witness else None, "status": status, "pub": pub}
        res.update({k: v for k, v in keep.items() if k != "sample_idx"})
        return res

    def novaChain(self, data, want_witness=False):
        """WitnessCalculator.novaChain over all devices: chunks are sharded (balanced by step count), no collective."""
        return _nova_chain(self._L, self._L.b3w_multi_nova_chain, self._m, self.witnessSize, data, want_witness)

    def witness_batch_host(self, h_in, n, h_out=None, h_status=None, h_pub=None, extras=None):
        """raw host pointers (e.g. from b3w_host_alloc); used by the benches.  extras: a _lib.BatchExtras or None"""
        _lib.check(self._L.b3w_multi_witness_batch_ex(self._m, h_in, n, h_out, h_status, h_pub, C.byref(extras) if extras is not None else None))
