"""generate_witness.py -- mirror of the reference CLI blake3_nova_js/generate_witness.js:1-20:

    python -m hot_proofs_blake3_circom_b200.generate_witness <file.wasm> <input.json> <output.wtns>

<file.wasm> is only used to identify the circuit (sha256); the witness is computed on the GPU.  Where the
reference's .wasm files are not at hand, a circuit name (blake3_compression, blake3_nova, blake3_nova_pasta,
blake3_nova_o1) is accepted in its place.
"""
import json
import os
import sys

from .witness_calculator import builder


def main(argv=None):
    argv = sys.argv if argv is None else argv
    if len(argv) != 4:
        print("Usage: python -m hot_proofs_blake3_circom_b200.generate_witness <file.wasm> <input.json> <output.wtns>")
        return 0
    with open(argv[2], "r", encoding="utf8") as f:
        inp = json.load(f)
    if os.path.exists(argv[1]):
        with open(argv[1], "rb") as f:
            code = f.read()
    else:
        code = argv[1]
    wc = builder(code)
    buff = wc.calculateWTNSBin(inp, 0)
    with open(argv[3], "wb") as f:
        f.write(buff.tobytes())
    return 0


if __name__ == "__main__":
    sys.exit(main())
