#!/usr/bin/env python3
"""make_golden.py -- regenerates the fixtures in tests/golden/ (run HERE, where /root/reference exists).

compression_golden.npz : the reference's own golden vector build/blake3_compression/testInp/{witness.wtns,
                         public.json} (the whole .wtns image) and the input that produces it
                         (test/witness_gen.test.ts:26,36 = genRandomChunk(new LCG(6429))).
compression_cases.npz  : witnesses of the reference wasm (Oracle A, oracle/_ref) for edge-case and random
                         inputs of blake3_compression.
nova_*_cases.npz       : the same for the three nova witness programs (added with the nova path).
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle.ref_wasm import RefWasm  # noqa: E402
from oracle.blake3_ref import LCG, gen_random_chunk  # noqa: E402
from hot_proofs_blake3_circom_b200.inputs import splitmix_compression_inputs  # noqa: E402


def compression():
    wtns = open(os.path.join(REF, "build/blake3_compression/testInp/witness.wtns"), "rb").read()
    pub = [int(x) for x in json.load(open(os.path.join(REF, "build/blake3_compression/testInp/public.json")))]
    c = gen_random_chunk(LCG(6429))
    row = np.array(c["h"] + c["m"] + c["t"] + [c["b"], c["d"]], np.uint32)
    np.savez_compressed(os.path.join(HERE, "compression_golden.npz"), row=row,
                        wtns=np.frombuffer(wtns, np.uint8), public=np.array(pub, np.uint32),
                        md5=np.frombuffer(hashlib.md5(wtns).hexdigest().encode(), np.uint8))
    print("compression_golden.npz: wtns md5", hashlib.md5(wtns).hexdigest())

    ref = RefWasm("compression")
    rows = [np.zeros(28, np.uint32), np.full(28, 0xFFFFFFFF, np.uint32)]
    r = np.full(28, 0xFFFFFFFF, np.uint32); r[26] = 64; r[27] = 15; rows.append(r)           # max words, legal b/d
    r = np.zeros(28, np.uint32); r[0:8] = row[0:8]; rows.append(r)                           # empty block: b = 0
    r = row.copy(); r[26] = 1; r[8] &= 0xFF; r[9:24] = 0; rows.append(r)                     # ragged: 1 byte
    r = row.copy(); r[27] = 11; r[24] = 0xFFFFFFFF; r[25] = 0xFFFFFFFF; rows.append(r)       # extreme t, d = 1|2|8
    rows += list(splitmix_compression_inputs(10))
    rows = np.stack(rows)
    wit, status, _ = ref.batch_u32(rows, nthreads=8)
    assert (status == 0).all()
    np.savez_compressed(os.path.join(HERE, "compression_cases.npz"), rows=rows, witness=wit)
    print("compression_cases.npz:", rows.shape[0], "cases")


if __name__ == "__main__":
    compression()
    try:
        from make_golden_nova import nova
        nova()
    except ImportError:
        pass
