"""wtns.py -- reader of the `.wtns` container calculateWTNSBin writes (blake3_nova_js/witness_calculator.js:208-272): the
consumer side of the format (what snarkjs / circom-scotia do with the file the reference's CLI leaves behind).

Layout (all little-endian): "wtns", u32 version = 2, u32 nSections = 2,
  section 1: u32 id = 1, u64 size = 8 + n8, u32 n8 = 32, prime[n8], u32 nWitness        (witness_calculator.js:227-244)
  section 2: u32 id = 2, u64 size = n8 * nWitness, nWitness field elements of n8 bytes    (:246-262)
Host-only (numpy); nothing here touches the GPU.
"""
import numpy as np

HEADER_BYTES = 76


class WtnsError(ValueError):
    pass


def parse_wtns(buf):
    """bytes / np.uint8[] of a .wtns file -> dict(version, n8, prime (int), n_witness, body = np.uint8[n_witness * n8] view).
    Raises WtnsError on anything that is not the two-section image the reference writes."""
    b = np.frombuffer(bytes(buf), np.uint8) if not isinstance(buf, np.ndarray) else np.ascontiguousarray(buf, np.uint8).reshape(-1)
    if b.size < 12 or b[:4].tobytes() != b"wtns":
        raise WtnsError("not a .wtns file (magic)")

    def u32(off):
        if off + 4 > b.size:
            raise WtnsError("truncated .wtns file")
        return int.from_bytes(b[off:off + 4].tobytes(), "little")

    def u64(off):
        if off + 8 > b.size:
            raise WtnsError("truncated .wtns file")
        return int.from_bytes(b[off:off + 8].tobytes(), "little")

    version, n_sections = u32(4), u32(8)
    if version != 2:
        raise WtnsError("unsupported .wtns version %d" % version)
    if n_sections != 2:
        raise WtnsError("expected 2 sections, found %d" % n_sections)
    if u32(12) != 1:
        raise WtnsError("first section is not the header section")
    size1, n8 = u64(16), u32(24)
    if n8 == 0 or n8 % 8 or size1 != 8 + n8:
        raise WtnsError("bad header section (n8 = %d, size = %d)" % (n8, size1))
    if 28 + n8 + 4 > b.size:
        raise WtnsError("truncated .wtns file")
    prime = int.from_bytes(b[28:28 + n8].tobytes(), "little")
    n_witness = u32(28 + n8)
    off2 = 24 + size1
    if u32(off2) != 2:
        raise WtnsError("second section is not the witness section")
    size2 = u64(off2 + 4)
    if size2 != n8 * n_witness:
        raise WtnsError("witness section holds %d bytes, %d x %d expected" % (size2, n_witness, n8))
    body0 = off2 + 12
    if b.size != body0 + size2:
        raise WtnsError("file is %d bytes, %d expected" % (b.size, body0 + size2))
    return {"version": version, "n8": n8, "prime": prime, "n_witness": n_witness, "body": b[body0:]}


def body_to_ints(body, n8=32):
    """np.uint8[n * n8] -> list of n Python ints (what `snarkjs wtns export json` prints, calculateWitness returns)"""
    raw = np.ascontiguousarray(body, np.uint8).tobytes()
    return [int.from_bytes(raw[i:i + n8], "little") for i in range(0, len(raw), n8)]
