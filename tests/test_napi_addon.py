"""The N-API addon (integration/js/addon/blake3wit_napi.cc) compiled against the stub node_api.h and executed against the
in-process N-API mock of tests/napi_mock (Node is not in this image): the layer a Node host would load between
witness_calculator.js and libblake3wit.so.  CPU tests cover the metadata calls and the argument checks; the GPU tests run
witnesses through it and compare them with the reference's golden vector / fixtures."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

import hot_proofs_blake3_circom_b200 as pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(ROOT, "tests", "napi_mock")
ADDON = os.path.join(ROOT, "integration", "js", "addon", "blake3wit_napi.cc")
K_NULL, K_U32, K_STRING, K_OBJECT, K_ARRAY, K_TYPEDARRAY, K_EXTERNAL, K_ERROR, K_PROMISE = 0, 1, 4, 5, 6, 8, 9, 10, 11
U8, U32, U64 = 1, 6, 10                                   # napi_uint8_array, napi_uint32_array, napi_biguint64_array
P = 21888242871839275222246405745257275088548364400416034343698204186575808495617


@pytest.fixture(scope="module")
def napi(built):
    so = os.path.join(MOCK, "libnapi_addon_test.so")
    srcs = [ADDON, os.path.join(MOCK, "napi_mock.cc"), os.path.join(MOCK, "node_api.h"), os.path.join(ROOT, "include", "blake3wit.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        libdir = os.path.dirname(pkg.lib_path())
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-shared", "-fPIC", "-DNODE_GYP_MODULE_NAME=blake3wit_napi",
                        "-I", MOCK, "-I", os.path.join(ROOT, "include"), ADDON, os.path.join(MOCK, "napi_mock.cc"),
                        "-L", libdir, "-lblake3wit", "-Wl,-rpath," + libdir, "-o", so], check=True)
    L = C.CDLL(so)
    vp = C.c_void_p
    for f in ("mk_exports", "mk_u32", "mk_i32", "mk_bool", "mk_str", "mk_typed", "mk_call", "mk_get", "mk_elem"):
        getattr(L, f).restype = vp
    L.mk_u32.argtypes, L.mk_i32.argtypes, L.mk_bool.argtypes, L.mk_str.argtypes = [C.c_uint32], [C.c_int32], [C.c_int], [C.c_char_p]
    L.mk_typed.argtypes = [C.c_int, vp, C.c_size_t]
    L.mk_call.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.mk_exception.restype = C.c_char_p
    L.mk_module_name.restype = C.c_char_p
    L.mk_kind.argtypes = [vp]
    L.mk_get.argtypes = [vp, C.c_char_p]
    L.mk_elem.argtypes = [vp, C.c_uint32]
    L.mk_as_u32.argtypes, L.mk_as_u32.restype = [vp], C.c_uint32
    L.mk_as_str.argtypes, L.mk_as_str.restype = [vp], C.c_char_p
    L.mk_typed_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(vp)]
    L.mk_promise.argtypes = [vp, C.POINTER(vp)]
    assert L.mk_exports()
    yield Napi(L)
    L.mk_release_all()                                    # runs the addon's finalizers (b3w_destroy of every context)


class Napi:
    def __init__(self, L):
        self.L = L

    def call(self, name, *args):
        """-> result handle; raises RuntimeError with the message the addon threw"""
        argv = (C.c_void_p * max(len(args), 1))(*args)
        r = self.L.mk_call(name.encode(), len(args), argv)
        exc = self.L.mk_exception()
        if exc is not None:
            raise RuntimeError(exc.decode())
        return r

    def typed(self, arr):
        arr = np.ascontiguousarray(arr)
        t = {np.dtype(np.uint8): U8, np.dtype(np.uint32): U32}[arr.dtype]
        return self.L.mk_typed(t, arr.ctypes.data, arr.size)

    def array_of(self, ta):
        t, n, p = C.c_int(), C.c_size_t(), C.c_void_p()
        assert self.L.mk_typed_info(ta, C.byref(t), C.byref(n), C.byref(p)) == 0
        dt = {U8: np.uint8, U32: np.uint32, U64: np.uint64}[t.value]
        if n.value == 0:
            return np.zeros(0, dt)
        ct = {U8: C.c_uint8, U32: C.c_uint32, U64: C.c_uint64}[t.value]
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), (n.value,)).copy()

    def await_(self, promise):
        """-> the resolution handle; raises RuntimeError(message) for a rejection (what `await` does in JS)"""
        v = C.c_void_p()
        st = self.L.mk_promise(promise, C.byref(v))
        assert st in (1, 2), "promise not settled"
        if st == 2:
            assert self.L.mk_kind(v) == K_ERROR
            raise RuntimeError(self.L.mk_as_str(v).decode())
        return v.value

    def get(self, obj, name):
        return self.L.mk_get(obj, name.encode())


def test_addon_registers_its_exports(napi):
    assert napi.L.mk_module_name() == b"blake3wit_napi"
    e = napi.L.mk_exports()
    for name in ("create", "circuitInfo", "inputSignal", "wtnsHeader", "witnessBatch", "witnessOne", "witnessBatchFr", "witnessOneFr"):
        assert napi.get(e, name), name
    with pytest.raises(RuntimeError, match="no such export"):
        napi.call("nope")


def test_circuit_info_and_header_through_the_addon(napi, golden):
    L = napi.L
    info = napi.call("circuitInfo", L.mk_u32(0))
    assert L.mk_kind(info) == K_OBJECT
    assert [L.mk_as_u32(napi.get(info, k)) for k in ("witnessSize", "nInputs", "n32", "nPublic")] == [24093, 28, 8, 16]
    ver = napi.get(info, "version")
    assert [L.mk_as_u32(L.mk_elem(ver, i)) for i in range(3)] == [2, 1, 6]
    prime = napi.array_of(napi.get(info, "prime"))
    assert int.from_bytes(prime.tobytes(), "little") == P          # what witness_calculator.js turns into this.prime
    assert L.mk_as_u32(napi.get(napi.call("circuitInfo", L.mk_u32(2)), "witnessSize")) == 23291
    with pytest.raises(RuntimeError, match="circuit 9 not built"):
        napi.call("circuitInfo", L.mk_u32(9))
    hdr = napi.array_of(napi.call("wtnsHeader", L.mk_u32(0)))
    assert hdr.tobytes() == golden["wtns"].tobytes()[:76]


def test_input_signal_lookup_through_the_addon(napi):
    L = napi.L
    sig = napi.call("inputSignal", L.mk_u32(0), L.mk_str(b"m"))
    assert (L.mk_as_u32(napi.get(sig, "offset")), L.mk_as_u32(napi.get(sig, "size"))) == (8, 16)
    sig = napi.call("inputSignal", L.mk_u32(1), L.mk_str(b"chunk_idx_high"))
    assert (L.mk_as_u32(napi.get(sig, "offset")), L.mk_as_u32(napi.get(sig, "size"))) == (11, 1)
    assert L.mk_kind(napi.call("inputSignal", L.mk_u32(0), L.mk_str(b"nope"))) == K_NULL    # -> signalSize 0 in the JS layer
    with pytest.raises(RuntimeError):
        napi.call("inputSignal", L.mk_str(b"m"), L.mk_u32(0))         # wrong argument types throw, they do not crash


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box WITHOUT a GPU")
def test_create_fails_loudly_without_gpu(napi):
    with pytest.raises(RuntimeError, match="no CUDA device"):
        napi.call("create", napi.L.mk_u32(0), napi.L.mk_i32(-1))


# ---- with a GPU: witnesses through the addon -------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ctx(napi):
    assert torch.cuda.is_available(), "these tests need the B200"
    h = napi.call("create", napi.L.mk_u32(0), napi.L.mk_i32(0))
    assert napi.L.mk_kind(h) == K_EXTERNAL
    return h


@pytest.mark.gpu
def test_witness_one_equals_reference_golden(napi, ctx, golden):
    body = napi.array_of(napi.await_(napi.call("witnessOne", ctx, napi.typed(golden["row"].astype(np.uint32)))))
    assert body.tobytes() == golden["wtns"].tobytes()[76:]
    with pytest.raises(RuntimeError, match="rows must be a Uint32Array"):
        napi.call("witnessOne", ctx, napi.typed(np.zeros(27, np.uint32)))
    with pytest.raises(RuntimeError, match="rows must be a Uint32Array"):
        napi.call("witnessOne", ctx, napi.typed(np.zeros(28, np.uint8)))


@pytest.mark.gpu
def test_witness_batch_through_the_addon(napi, ctx, cases):
    L = napi.L
    rows, want = cases["rows"], cases["witness"]
    n = rows.shape[0]
    res = napi.await_(napi.call("witnessBatch", ctx, napi.typed(rows.reshape(-1)), L.mk_u32(n), L.mk_bool(1)))
    assert napi.array_of(napi.get(res, "witness")).tobytes() == want.tobytes()
    assert (napi.array_of(napi.get(res, "status")) == 0).all()
    pub = napi.array_of(napi.get(res, "pub")).reshape(n, 16)
    assert np.array_equal(pub, want.view(np.uint32).reshape(n, 24093, 8)[:, 1:17, 0])
    res = napi.await_(napi.call("witnessBatch", ctx, napi.typed(rows.reshape(-1)), L.mk_u32(n), L.mk_bool(0)))
    assert L.mk_kind(napi.get(res, "witness")) == K_NULL and np.array_equal(napi.array_of(napi.get(res, "pub")).reshape(n, 16), pub)
    assert napi.get(res, "sums") is None
    # streamed, with the per-instance witness checksums (BigUint64Array): equal to the checksum of the fixture's witnesses
    from conftest import checksum_np
    res = napi.await_(napi.call("witnessBatch", ctx, napi.typed(rows.reshape(-1)), L.mk_u32(n), L.mk_bool(0), L.mk_bool(1)))
    assert L.mk_kind(napi.get(res, "witness")) == K_NULL
    assert np.array_equal(napi.array_of(napi.get(res, "sums")), checksum_np(want, 24093))


@pytest.mark.gpu
def test_field_element_inputs_and_assert_text_through_the_addon(napi, ctx):
    wide = np.load(os.path.join(ROOT, "tests", "golden", "compression_wide_cases.npz"))
    fr, status, valid, want, text = wide["fr"], wide["status"], list(wide["valid"]), wide["witness"], wide["text"]
    for i in (0, 1, 2, 5, 40):
        call = napi.call("witnessOneFr", ctx, napi.typed(fr[i].reshape(-1)))
        if status[i] == 0:
            assert napi.array_of(napi.await_(call)).tobytes() == want[valid.index(i)].tobytes()
        else:
            with pytest.raises(RuntimeError) as e:
                napi.await_(call)
            # witness_calculator.js:21-43,159-162: "Error: " + "Assert Failed.\n" + the printErrorMessage lines
            assert str(e.value) == "Error: Assert Failed.\n" + bytes(text[i]).decode()
    n = len(status)
    res = napi.await_(napi.call("witnessBatchFr", ctx, napi.typed(fr.reshape(-1)), napi.L.mk_u32(n), napi.L.mk_bool(1)))
    assert np.array_equal(napi.array_of(napi.get(res, "status")), status.astype(np.uint8))
    assert np.array_equal(napi.array_of(napi.get(res, "witness")).reshape(n, -1)[valid], want)


@pytest.mark.gpu
def test_nova_field_element_inputs_through_the_addon(napi):
    """a nova context: field-valued n_blocks / depths give the reference's witness, a failing CheckDepth its text"""
    wide = np.load(os.path.join(ROOT, "tests", "golden", "nova_wide_cases.npz"))
    fr, status, valid, want, text = (wide["nova_pasta_o2" + k] for k in ("_fr", "_status", "_valid", "_witness", "_text"))
    valid = list(valid)
    h = napi.call("create", napi.L.mk_u32(2), napi.L.mk_i32(0))
    seen = set()
    for i in range(len(status)):
        if status[i] in seen and i > 8:
            continue
        seen.add(int(status[i]))
        call = napi.call("witnessOneFr", h, napi.typed(fr[i].reshape(-1)))
        if status[i] == 0:
            assert napi.array_of(napi.await_(call)).tobytes() == want[valid.index(i)].tobytes()
        else:
            with pytest.raises(RuntimeError) as e:
                napi.await_(call)
            assert str(e.value) == "Error: Assert Failed.\n" + bytes(text[i]).decode()
    assert seen == {0, 4}


@pytest.mark.gpu
def test_create_with_flags_byte_check(napi, cases):
    """create(circuit, device, flags): a context with B3W_FLAG_BYTE_CHECK (16) checks every witness where it lies in the HBM
    ring -- the fixture's witnesses pass and come back unchanged; unknown flags are refused with the library's text"""
    L = napi.L
    h = napi.call("create", L.mk_u32(0), L.mk_i32(0), L.mk_u32(16))
    rows, want = cases["rows"], cases["witness"]
    n = rows.shape[0]
    res = napi.await_(napi.call("witnessBatch", h, napi.typed(rows.reshape(-1)), L.mk_u32(n), L.mk_bool(1)))
    assert (napi.array_of(napi.get(res, "status")) == 0).all()
    assert napi.array_of(napi.get(res, "witness")).tobytes() == want.tobytes()
    with pytest.raises(RuntimeError, match="unknown flags"):
        napi.call("create", L.mk_u32(0), L.mk_i32(0), L.mk_u32(1 << 20))
