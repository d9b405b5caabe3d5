"""Nova fixtures: witnesses of the reference's three nova witness programs (Oracle A) for step inputs that cover
leaf / first / last / parent / root, assert failures, and values far outside the honest domain."""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def nova():
    import circuit_model as cm
    from oracle.ref_wasm import RefWasm
    from hot_proofs_blake3_circom_b200.inputs import splitmix_nova_inputs
    rng = random.Random(20261017)
    rows = [cm.random_nova_inputs(rng, edge=e) for e in (1, 2, 3, 4, 5, 6, 4, 0, 0, 0)]
    rows = np.concatenate([np.array(rows, np.uint32), splitmix_nova_inputs(6)])
    for v in ("nova_bn_o2", "nova_pasta_o2", "nova_bn_o1"):
        ref = RefWasm(v)
        wit, status, _ = ref.batch_u32(rows, nthreads=8)
        wit[status != 0] = 0
        np.savez_compressed(os.path.join(HERE, "%s_cases.npz" % v), rows=rows, witness=wit, status=status)
        print("%s_cases.npz: %d cases, status %s" % (v, len(rows), list(status)))


if __name__ == "__main__":
    nova()
