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
print("sanitize_run done")
