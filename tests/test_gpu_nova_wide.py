"""GPU tests of the nova step circuits on their full input domain (any 32 field elements, as the reference's wasm takes
them): tests/golden/nova_wide_cases.npz made with the reference's own witness programs, Oracle B on fresh random inputs.
Bar: bit-exact witnesses, identical status, identical "Assert Failed." text."""
import os
import random

import numpy as np
import pytest
import torch

import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200 import _lib
from hot_proofs_blake3_circom_b200 import inputs as gen
from oracle import port

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NOVA = (("blake3_nova", "nova_bn_o2"), ("blake3_nova_pasta", "nova_pasta_o2"), ("blake3_nova_o1", "nova_bn_o1"))


@pytest.fixture(scope="module")
def nova_wide_cases():
    return np.load(os.path.join(GOLDEN, "nova_wide_cases.npz"))


def as_input(v):
    v = [int(x) for x in v]
    return {"n_blocks": v[0], "block_count": v[1], "h": v[2:10], "chunk_idx_low": v[10], "chunk_idx_high": v[11],
            "leaf_depth": v[12], "total_depth": v[13], "depth": v[14], "m": v[15:31], "b": v[31]}


def random_rows(n, p, seed):
    rnd = random.Random(seed)
    out = []
    for it in range(n):
        leaf = rnd.randrange(1, 65)
        nb = rnd.randrange(1, 17)
        v = [nb, rnd.randrange(nb)] + [rnd.randrange(2**32) for _ in range(8)] + [rnd.randrange(2**32), rnd.randrange(2**32), leaf, leaf,
             rnd.randrange(leaf)] + [rnd.randrange(2**32) for _ in range(16)] + [rnd.randrange(65)]
        X = rnd.randrange(p)
        mode = it % 8
        if mode == 0:
            pass                                                     # a plain u32 step inside the wide batch
        elif mode == 1:
            v[0] = rnd.choice([X, 2**40, p - 3])
            v[1] = rnd.choice([v[1], (v[0] - 1) % p, 0, rnd.randrange(p)])
        elif mode == 2:
            d = rnd.choice([1, 1, 2, 17, 255, 256, 257, 0])
            v[14], v[12], v[13] = X, (X + d) % p, rnd.choice([(X + d) % p, (X + rnd.randrange(2, 66)) % p, rnd.randrange(p)])
        elif mode == 3:
            v[14] = rnd.randrange(max(leaf - 1, 1))                   # mostly parents
            v[10], v[11] = rnd.choice([(2**64 + rnd.randrange(2**32), 0), ((p - 3 * 2**32) % p + rnd.randrange(2**32), 3),
                                       (rnd.randrange(2**32), 2**32 + 1), (rnd.randrange(2**32), 2**33)])
        elif mode == 4:
            v[14] = rnd.randrange(max(leaf - 1, 1))
            v[2 + rnd.randrange(8)] = rnd.choice([2**32 + 5, X, p - 1])
            v[23 + rnd.randrange(8)] = X                              # m[8..15]: ignored on parent steps
        elif mode == 5:
            for j in rnd.sample(range(16), 3):
                v[15 + j] = rnd.choice([2**32 + rnd.randrange(2**30), p - 1 - rnd.randrange(2**30), 2**32, p - 1])
        elif mode == 6:
            v[13] = rnd.choice([X, (v[14] + rnd.randrange(70)) % p])
        else:
            Y = rnd.randrange(p)
            v[14], v[12], v[13] = Y, (Y + rnd.randrange(1, 200)) % p, (Y + rnd.randrange(1, 70)) % p
            v[0] = rnd.randrange(p)
            v[1] = rnd.choice([0, (v[0] - 1) % p, rnd.randrange(p)])
        out.append(v)
    return out


@pytest.mark.parametrize("name,variant", NOVA)
def test_reference_fixture(built, nova_wide_cases, name, variant):
    wc = pkg.builder(name, device=0)
    fr, status, valid, want = (nova_wide_cases[variant + k] for k in ("_fr", "_status", "_valid", "_witness"))
    res = wc.calculateWitnessBatchFr(fr)
    assert np.array_equal(res["status"], status.astype(np.uint8))
    for k, i in enumerate(valid):
        assert np.array_equal(res["witness"][i], want[k]), (variant, i)
    assert (res["pub"][status == 4] == 0).all()
    # pub = the low 32 bits of the 15 outputs = witness slots 1..15
    assert np.array_equal(res["pub"][valid], want.view(np.uint32).reshape(len(valid), wc.witnessSize, 8)[:, 1:16, 0])
    wc.close()


@pytest.mark.parametrize("name,variant", NOVA)
def test_against_oracle_b_every_byte(built, name, variant):
    wc = pkg.builder(name, device=0, chunk=128)                     # several ring chunks
    vals = random_rows(320, wc.prime, 21)
    status, wit = [], []
    for v in vals:
        rc, w = port.witness_fr(variant, [x % wc.prime for x in v])
        status.append(rc), wit.append(w)
    status = np.array(status)
    assert (status == 0).sum() > 120 and (status == 4).sum() > 25
    res = wc.calculateWitnessBatch([as_input(v) for v in vals])
    assert np.array_equal(res["status"], status.astype(np.uint8))
    for i in np.nonzero(status == 0)[0]:
        assert np.array_equal(res["witness"][i], wit[i]), (variant, i, vals[i])
    wc.close()


def test_single_witness_api_and_error_text(built, nova_wide_cases, capsys):
    wc = pkg.builder("blake3_nova", device=0)
    fr, status, valid, want, text = (nova_wide_cases["nova_bn_o2" + k] for k in ("_fr", "_status", "_valid", "_witness", "_text"))
    valid = list(valid)
    n_ok = n_bad = 0
    for i in range(len(status)):
        v = [int.from_bytes(fr[i, k].tobytes(), "little") for k in range(32)]
        if status[i] == 0 and n_ok < 4:
            n_ok += 1
            assert np.array_equal(wc.calculateBinWitness(as_input(v), 0), want[valid.index(i)])
        elif status[i] == 4 and n_bad < 8:
            n_bad += 1
            with pytest.raises(RuntimeError) as e:
                wc.calculateWitness(as_input(v), 0)
            assert str(e.value) == "Error: Assert Failed.\n" + bytes(text[i]).decode()
    assert n_ok == 4 and n_bad == 8
    wc.close()


def test_u32_batches_stay_on_the_hot_kernel_and_agree(built):
    """the same u32 rows through b3w_witness_batch_fr (u32 -> hot kernel) and, with one field-valued instance appended,
    through the general kernel: identical witnesses"""
    wc = pkg.builder("blake3_nova_pasta", device=0)
    rows = gen.splitmix_nova_inputs(96, first=5)
    want = wc.calculateWitnessBatch(rows)
    a = wc.calculateWitnessBatchFr([[int(x) for x in r] for r in rows])
    assert np.array_equal(a["witness"], want["witness"]) and np.array_equal(a["status"], want["status"])
    assert np.array_equal(a["pub"], want["pub"])
    extra = [int(x) for x in rows[0]]
    extra[0] = wc.prime - 3                                          # n_blocks: a field element -> the whole batch goes wide
    b = wc.calculateWitnessBatchFr([[int(x) for x in r] for r in rows] + [extra])
    assert np.array_equal(b["witness"][:96], want["witness"]) and np.array_equal(b["status"][:96], want["status"])
    assert np.array_equal(b["pub"][:96], want["pub"])
    assert b["status"][96] == 0
    wc.close()


def test_fused_check_flag_covers_field_valued_instances(built):
    """round 1 refused the fused-check flag for a batch with a field-valued instance; now the u32 instances get the fused
    check and the field-valued ones the stand-alone evaluator on their finished witnesses (the built-in O2-form system)"""
    wc = pkg.builder("blake3_nova", device=0, fused_check=True)
    rows = gen.splitmix_nova_inputs(40)
    v = [[int(x) for x in r] for r in rows]
    v[1][0] = 2**40                                                  # n_blocks: any field element is fine
    v[2][13] = wc.prime - 5                                          # total_depth
    v[3][12] = 2**40                                                 # leaf_depth - depth out of range: Assert Failed.
    res = wc.calculateWitnessBatchFr(v, sums=True, first_bad=True)
    want = [port.witness_fr("nova_bn_o2", [x % wc.prime for x in r]) for r in v]
    assert list(res["status"]) == [rc for rc, _ in want] and res["status"][3] == 4
    for i, (rc, w) in enumerate(want):
        if rc == 0:
            assert np.array_equal(res["witness"][i], w), i
    assert (res["first_bad"] == _lib.B3W_NO_ROW).all()
    wc.close()
