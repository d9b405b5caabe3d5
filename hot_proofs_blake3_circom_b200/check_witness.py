"""check_witness.py -- the consumer-side companion of generate_witness.py: is the witness in a .wtns file valid?

    python -m hot_proofs_blake3_circom_b200.check_witness <file.wasm|circuit-name> <witness.wtns> [<file.r1cs>]

What `snarkjs wtns check <r1cs> <wtns>` answers in the reference's tool chain (snarkjs 0.7.2 under circomkit; the reference's
tests ask the same through circom_tester's expectPass, test/blake3_hash.test.ts:36,57): every constraint of the circuit
evaluated on the file's witness -- here by the stand-alone R1CS checker on the GPU (b3w_r1cs_check_device) against the
built-in constraint system of the circuit, or against the iden3 .r1cs file given as third argument (b3w_r1cs_load_file).
Exit status 0 = the witness satisfies every row; 1 = some row is violated (its number is printed) or the file is not a
.wtns image of this circuit.  There is no CPU path: without the CUDA library or a GPU the command fails.
"""
import os
import sys

from .witness_calculator import builder
from .wtns import WtnsError

USAGE = "Usage: python -m hot_proofs_blake3_circom_b200.check_witness <file.wasm|circuit> <witness.wtns> [<file.r1cs>]"


def main(argv=None):
    argv = sys.argv if argv is None else argv
    if len(argv) not in (3, 4):
        print(USAGE)
        return 0
    if os.path.exists(argv[1]):
        with open(argv[1], "rb") as f:
            code = f.read()
    else:
        code = argv[1]
    with open(argv[2], "rb") as f:
        image = f.read()
    wc = builder(code, lazy=True)                       # the container is checked on the host before a GPU is asked for
    try:
        body = wc.wtnsBody(image)
    except WtnsError as e:
        print("INVALID FILE: %s" % e)
        return 1
    rows = wc.r1cs_load(argv[3]) if len(argv) == 4 else wc.r1cs_program_info()["rows"]
    status, bad = wc.checkWitnesses(body[None, :])
    if status[0] == 0:
        print("WITNESS IS CORRECT (%d constraints, %d values)" % (rows, wc.witnessSize))
        return 0
    if int(bad[0]) == 0xFFFFFFFE:                       # b3w_r1cs_check_device: a slot >= p
        print("WITNESS CHECK FAILED: a value is not a canonical field element")
    else:
        print("WITNESS CHECK FAILED: constraint %d does not hold" % int(bad[0]))
    return 1


if __name__ == "__main__":
    sys.exit(main())
