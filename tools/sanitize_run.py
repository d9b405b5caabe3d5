"""Small end-to-end pass over every kernel for compute-sanitizer (memcheck / racecheck / synccheck); no timing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hot_proofs_blake3_circom_b200 as pkg
from hot_proofs_blake3_circom_b200 import inputs as gen

for name, rows in (("blake3_compression", gen.splitmix_compression_inputs(70)), ("blake3_nova_pasta", gen.splitmix_nova_inputs(70)),
                   ("blake3_nova_o1", gen.splitmix_nova_inputs(40))):
    for fused in (False, True):
        wc = pkg.builder(name, device=0, chunk=32, fused_check=fused)
        res = wc.calculateWitnessBatch(rows)
        assert not (res["status"] & 3).any()
        wc.close()
    wc = pkg.builder(name, device=0)
    pk = wc.calculateWitnessBatchPacked(rows)
    wc.unpackWitnesses(pk["packed"][:8])
    if name != "blake3_compression":
        wc.novaChain(bytes(range(256)) * 13)
    wc.close()
# the wide-domain kernels (blake3_compression with message words outside u32; plain and fused-check contexts)
P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
vals = []
for i, r in enumerate(gen.splitmix_compression_inputs(48)):
    v = [int(x) for x in r]
    if i % 3 == 0:
        v[8 + i % 16] = 2**32 + i
    elif i % 3 == 1:
        v[8 + i % 16] = P - 1 - i
    if i % 11 == 0:
        v[26] = 2**33                      # asserts
    vals.append(v)
for fused in (False, True):
    wc = pkg.builder("blake3_compression", device=0, chunk=32, fused_check=fused)
    res = wc.calculateWitnessBatchFr(vals)
    assert set(np.unique(res["status"])) <= {0, 4}
    wc.close()
# the general nova kernel (field-valued inputs), all three builds
for name in ("blake3_nova", "blake3_nova_pasta", "blake3_nova_o1"):
    wc = pkg.builder(name, device=0, chunk=16)
    vals = [[int(x) for x in r] for r in gen.splitmix_nova_inputs(40)]
    for i, v in enumerate(vals):
        if i % 4 == 0:
            v[0] = wc.prime - 3 - i                     # n_blocks: a field element
        elif i % 4 == 1:
            X = (0x123456789ABCDEF << 120) + i
            v[14], v[12], v[13] = X, X + 1 + i % 200, X + 2 + i % 60
        elif i % 4 == 2:
            v[15 + i % 16] = 2**32 + i
    res = wc.calculateWitnessBatchFr(vals)
    assert set(np.unique(res["status"])) <= {0, 4}
    wc.close()
# round 2: fused checksums + samples, mixed-domain Fr batches (device conversion, wide list, listed stand-alone check), the
# compiled stand-alone checker on all four built-in systems, the TMA store mode, packed / hybrid export, the chain's device form
import torch
for name, rows in (("blake3_compression", gen.splitmix_compression_inputs(70)), ("blake3_nova", gen.splitmix_nova_inputs(70)),
                   ("blake3_nova_pasta", gen.splitmix_nova_inputs(70)), ("blake3_nova_o1", gen.splitmix_nova_inputs(40))):
    wc = pkg.builder(name, device=0, chunk=32, fused_check=True)
    res = wc.calculateWitnessBatch(rows, want_witness=False, sums=True, samples=[0, 33, len(rows) - 1], first_bad=True)
    assert not (res["status"] & 3).any()
    fr = np.zeros((len(rows), rows.shape[1], 32), np.uint8)
    fr[:, :, 0:4] = rows.view(np.uint8).reshape(len(rows), rows.shape[1], 4)
    k = 8 if name == "blake3_compression" else 0
    fr[5, k] = np.frombuffer((wc.prime - 7).to_bytes(32, "little"), np.uint8)       # a field-valued instance inside the batch
    res = wc.calculateWitnessBatchFr(fr, sums=True, first_bad=True)
    assert set(np.unique(res["status"])) <= {0, 4}
    n = len(rows)
    d_in = torch.from_numpy(rows.view(np.int32)).cuda()
    d_out = torch.zeros(n * wc.witnessSize * 32, dtype=torch.uint8, device="cuda")
    d_st = torch.zeros(n, dtype=torch.uint8, device="cuda")
    d_bad = torch.zeros(n, dtype=torch.int32, device="cuda")
    wc.witness_batch_device(d_in.data_ptr(), n, d_out.data_ptr(), d_st.data_ptr(), 0, 0)
    wc.r1cs_check_device(d_out.data_ptr(), n, d_st.data_ptr(), d_bad.data_ptr(), 0)
    d_out.view(n, wc.witnessSize, 32)[:, 100, 0] += 3                               # violated rows: the slow paths of the checker
    d_out.view(n, wc.witnessSize, 32)[0, 200, 31] = 0x10
    wc.r1cs_check_device(d_out.data_ptr(), n, d_st.data_ptr(), d_bad.data_ptr(), 0)
    torch.cuda.synchronize()
    assert int(d_st.min()) == 7
    pk = wc.calculateWitnessBatchPacked(rows)
    wc.unpackWitnessesHost(pk["packed"][:4])
    wc.calculateWitnessBatchHybrid(rows[:40])
    if name == "blake3_compression":
        wc2 = pkg.builder(name, device=0)
        wc2.set_store_mode(1)
        wc2.set_launch(0, 7)
        wc2.witness_batch_device(d_in.data_ptr(), n, d_out.data_ptr(), d_st.data_ptr(), 0, 0)
        torch.cuda.synchronize()
        wc2.close()
    else:
        data = bytes(range(256)) * 13
        ns = wc.novaChain(data)["total_steps"]
        d_o = torch.zeros(ns * wc.witnessSize * 32, dtype=torch.uint8, device="cuda")
        wc.novaChainDevice(data, d_o.data_ptr())
    wc.close()
# round 2, second half: B3W_FLAG_BYTE_CHECK (the stand-alone checker -- rotating-register streaming loop, instances from a
# global counter, side-table places from the slot layout -- chained to the generator on the ring) and the growing ring
for name, rows in (("blake3_compression", gen.splitmix_compression_inputs(150)), ("blake3_nova_o1", gen.splitmix_nova_inputs(90))):
    wc = pkg.builder(name, device=0, byte_check=True)
    for n in (3, len(rows)):
        res = wc.calculateWitnessBatch(rows[:n], want_witness=False, sums=True, first_bad=True)
        assert not (res["status"] & 3).any()
    wc.close()
print("sanitize_run done")
