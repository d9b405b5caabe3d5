"""hot_proofs_blake3_circom_b200 -- B200-native batched circom witness generation for the BLAKE3
compression circuit and the blake3_nova / blake3_nova_pasta step circuits.

Only the hot path of banyancomputer/hot-proofs-blake3-circom lives here:
  csrc/                  hand-written sm_100a kernels + the C ABI (libblake3wit.so, include/blake3wit.h)
  witness_calculator.py  host-side mirror of the reference's witness_calculator.js API
  generate_witness.py    mirror of the reference's generate_witness.js CLI
  wtns.py, check_witness.py  the consumer side: .wtns reader and `check_witness` CLI (every constraint on the file's witness, on the GPU)
There is NO CPU fallback: every compute call fails loudly if the CUDA library or a GPU is missing.
"""
from .witness_calculator import builder, WitnessCalculator, MultiGpuCalculator, CIRCUITS, circuit_from_wasm  # noqa: F401
from ._lib import lib, lib_path, B3WError  # noqa: F401

__all__ = ["builder", "WitnessCalculator", "MultiGpuCalculator", "CIRCUITS", "circuit_from_wasm", "lib", "lib_path", "B3WError"]
