"""The consumer side of the .wtns container (hot_proofs_blake3_circom_b200/wtns.py, check_witness.py): CPU tests of the reader
on the reference's golden file and on damaged images; GPU tests of the check itself (`snarkjs wtns check` / circom_tester
expectPass in the reference's tool chain, test/blake3_hash.test.ts:36,57)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200 import _lib
from hot_proofs_blake3_circom_b200 import inputs as gen
from hot_proofs_blake3_circom_b200.wtns import WtnsError, body_to_ints, parse_wtns

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WS = 24093
CLI = [sys.executable, "-m", "hot_proofs_blake3_circom_b200.check_witness"]


def test_reader_on_the_reference_golden_file(golden, built):
    w = parse_wtns(golden["wtns"])
    wc = pkg.builder("blake3_compression", lazy=True)
    assert (w["version"], w["n8"], w["n_witness"], w["prime"]) == (2, 32, WS, wc.prime)
    assert w["body"].tobytes() == golden["wtns"].tobytes()[76:]
    vals = body_to_ints(w["body"])
    assert len(vals) == WS and vals[0] == 1 and vals[1:17] == [int(x) for x in golden["public"]]     # public.json = main.out
    assert wc.wtnsBody(golden["wtns"].tobytes()).tobytes() == w["body"].tobytes()
    # the reader accepts what the writer's header function produces, for all four circuits
    for name in ("blake3_compression", "blake3_nova", "blake3_nova_pasta", "blake3_nova_o1"):
        c = pkg.builder(name, lazy=True)
        hdr = np.empty(76, np.uint8)
        _lib.check(pkg.lib().b3w_wtns_header(c.circuit, hdr.ctypes.data))
        img = np.concatenate([hdr, np.zeros(c.witnessSize * 32, np.uint8)])
        assert c.wtnsBody(img).size == c.witnessSize * 32 and parse_wtns(img)["prime"] == c.prime


def test_reader_refuses_damaged_and_foreign_images(golden, built):
    img = golden["wtns"].copy()
    wc = pkg.builder("blake3_compression", lazy=True)
    cases = []
    b = img.copy(); b[0] ^= 1; cases.append((b, "magic"))
    b = img.copy(); b[4] = 3; cases.append((b, "version"))
    b = img.copy(); b[8] = 3; cases.append((b, "sections"))
    b = img.copy(); b[24] = 24; cases.append((b, "header section"))
    b = img.copy(); b[64] = 1; cases.append((b, "witness section"))                   # section id 2 -> 1
    b = img.copy(); b[68] ^= 1; cases.append((b, "bytes"))                            # section size
    cases.append((img[:-1], "bytes"))
    cases.append((img[:40], "truncated"))
    cases.append((np.zeros(3, np.uint8), "magic"))
    for buf, word in cases:
        with pytest.raises(WtnsError) as e:
            wc.wtnsBody(buf)
        assert word in str(e.value), (word, str(e.value))
    # a well-formed image of ANOTHER circuit / field
    with pytest.raises(WtnsError) as e:
        pkg.builder("blake3_nova_pasta", lazy=True).wtnsBody(img)
    assert "prime" in str(e.value)
    with pytest.raises(WtnsError) as e:
        pkg.builder("blake3_nova", lazy=True).wtnsBody(img)                           # same prime, other witness size
    assert "witness values" in str(e.value)


def test_cli_usage_and_host_side_refusals(golden, tmp_path, built):
    r = subprocess.run(CLI, cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("Usage:")
    bad = tmp_path / "bad.wtns"
    img = golden["wtns"].copy()
    img[0] ^= 1
    bad.write_bytes(img.tobytes())
    r = subprocess.run(CLI + ["blake3_compression", str(bad)], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout.startswith("INVALID FILE: not a .wtns file")
    good = tmp_path / "good.wtns"
    good.write_bytes(golden["wtns"].tobytes())
    r = subprocess.run(CLI + ["blake3_nova_pasta", str(good)], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 1 and "prime" in r.stdout


def test_check_fails_loudly_without_a_gpu(golden, tmp_path, built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box WITHOUT a GPU")
    good = tmp_path / "good.wtns"
    good.write_bytes(golden["wtns"].tobytes())
    r = subprocess.run(CLI + ["blake3_compression", str(good)], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode not in (0, 1) or "Traceback" in r.stderr                      # an error, not a verdict
    assert "WITNESS IS CORRECT" not in r.stdout


# ---- GPU --------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_golden_wtns_passes_and_a_flipped_bit_fails(golden, built):
    wc = pkg.builder("blake3_compression", device=0)
    ok, bad = wc.checkWTNSBin(golden["wtns"].tobytes())
    assert ok and bad == _lib.B3W_NO_ROW
    assert wc.checkWTNSBin(wc.calculateWTNSBin(dict(zip(("h", "m", "t", "b", "d"), split_row(golden["row"]))), 0)) == (True, _lib.B3W_NO_ROW)
    img = golden["wtns"].copy()
    img[76 + 32 * 1] += 1                                                             # out[0] + 1 (low byte 0x6A: no carry)
    ok, bad = wc.checkWTNSBin(img)
    assert not ok and bad < 24544
    img = golden["wtns"].copy()
    img[76 + 32 * 5:76 + 32 * 6] = 0xFF                                               # slot 5 >= p
    assert wc.checkWTNSBin(img) == (False, 0xFFFFFFFE)
    wc.close()


def split_row(row):
    r = [int(x) for x in row]
    return r[0:8], r[8:24], r[24:26], r[26], r[27]


@pytest.mark.gpu
def test_check_witnesses_of_a_batch(built):
    from oracle import port
    for name, variant, rows_fn in (("blake3_nova_pasta", "nova_pasta_o2", gen.splitmix_nova_inputs),
                                   ("blake3_compression", "compression", gen.splitmix_compression_inputs)):
        wc = pkg.builder(name, device=0)
        rows = rows_fn(40, first=21)
        wit, _, st = port.witness_batch(variant, rows, want="both")                   # witnesses from the ORACLE, checked on the GPU
        keep = st == 0
        wit = wit[keep].copy()
        status, bad = wc.checkWitnesses(wit)
        assert not status.any() and (bad == _lib.B3W_NO_ROW).all()
        wit[3, 32 * 1] += 1                                                            # the first output of instance 3 (low byte 4)
        status, bad = wc.checkWitnesses(wit)
        assert status[3] == _lib.B3W_R1CS_VIOLATION and bad[3] != _lib.B3W_NO_ROW and not np.delete(status, 3).any()
        wc.close()


@pytest.mark.gpu
def test_cli_check_witness(golden, tmp_path, built):
    good = tmp_path / "witness.wtns"
    good.write_bytes(golden["wtns"].tobytes())
    r = subprocess.run(CLI + ["blake3_compression", str(good)], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("WITNESS IS CORRECT (24544 constraints, 24093 values)"), r.stdout + r.stderr
    img = golden["wtns"].copy()
    img[76 + 32 * 1] += 1                                                             # out[0] + 1
    bad = tmp_path / "bad.wtns"
    bad.write_bytes(img.tobytes())
    r = subprocess.run(CLI + ["blake3_compression", str(bad)], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout.startswith("WITNESS CHECK FAILED: constraint "), r.stdout + r.stderr
