#!/usr/bin/env python3
"""gen_tables.py -- OFFLINE generator of the kernel's slot-descriptor tables (needs /root/reference).

For every circuit variant it
  1. runs the reference witness program once (Oracle A build in oracle/_ref) to read the wasm's own
     witness -> signal table ("the table is authoritative"),
  2. evaluates tools/circuit_model.py to learn, for each signal, where its value lives in the kernel's
     compact trace,
  3. validates the model signal-by-signal against the wasm's memory and the expanded witness against the
     wasm's witness on several inputs,
  4. writes run-length-encoded descriptor tables to hot_proofs_blake3_circom_b200/csrc/slot_tables.h
     and the witness->signal tables to oracle/w2s_tables.h (for the C oracle).

Descriptor (u32):  bits 0..15 trace index | bits 16..20 bit index | bits 24..26 kind
   kind 0 BIT   value = (trace[t] >> k) & 1
   kind 1 W32   value = trace[t]
   kind 2 W64   value = trace[t] | trace[t+1] << 32
   kind 3 S64   value = (int64)(trace[t] | trace[t+1] << 32) mod p          (negative -> p - |x|)
   kind 4 INV   value = inverse mod p of that signed value (0 for 0)        (circomlib IsZero.inv)
Run-length record {desc0, count, delta}: slot j of the run has descriptor desc0 + j*delta.

Usage: python tools/gen_tables.py [--check-only] [--trials N]
This script is the only place the product's build touches the reference; its outputs are committed.
"""
import argparse
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import circuit_model as cm  # noqa: E402

KIND = {"B": 0, "W": 1, "Q": 2, "S": 3, "I": 4}
W2S_OFFSET = {"compression": 6244, "nova_bn_o2": 6260, "nova_pasta_o2": 6260, "nova_bn_o1": 6260}


def desc_of(sym):
    k = sym[0]
    if k == "C":
        assert sym[1] in (0, 1), "constant %r cannot be a witness slot" % (sym,)
        return (KIND["W"] << 24) | (cm.TR_ONE if sym[1] else cm.TR_ZERO)
    if k == "B":
        return (KIND["B"] << 24) | (sym[2] << 16) | sym[1]
    return (KIND[k] << 24) | sym[1]


def rle(descs):
    """Greedy run-length encoding with per-run constant delta."""
    out, i, n = [], 0, len(descs)
    while i < n:
        if i + 1 < n:
            delta = descs[i + 1] - descs[i]
            j = i + 1
            while j + 1 < n and descs[j + 1] - descs[j] == delta:
                j += 1
            cnt = j - i + 1
            if cnt >= 3 or (cnt == 2 and delta in (0, 1, 1 << 16)):
                out.append((descs[i], cnt, delta))
                i = j + 1
                continue
        out.append((descs[i], 1, 0))
        i += 1
    return out


def expand(trace, descs, prime):
    """Reference expansion in Python (what the kernel does): -> list of ints."""
    out = []
    for d in descs:
        t, k, kind = d & 0xFFFF, (d >> 16) & 31, d >> 24
        if kind == 0:
            out.append((trace[t] >> k) & 1)
        elif kind == 1:
            out.append(trace[t])
        elif kind == 2:
            out.append(trace[t] | (trace[t + 1] << 32))
        elif kind in (3, 4):
            x = trace[t] | (trace[t + 1] << 32)
            x = x - (1 << 64) if x >> 63 else x
            out.append(x % prime if kind == 3 else (pow(x % prime, -1, prime) if x else 0))
    return out


def model_for(variant, inputs, prime):
    if variant == "compression":
        return cm.CompressionModel(inputs, prime)
    return cm.NovaModel(inputs, prime, o1=(variant == "nova_bn_o1"))


def random_inputs(variant, rng, edge=0):
    if variant == "compression":
        r32 = lambda: rng.getrandbits(32)
        if edge == 1:
            return [0xFFFFFFFF] * 28
        if edge == 2:
            return [0] * 28
        b = 4 * rng.randrange(17)
        m = [r32() if i < b // 4 else 0 for i in range(16)]
        return [r32() for _ in range(8)] + m + [r32(), r32(), b, rng.randrange(16)]
    return cm.random_nova_inputs(rng, edge)


def decode_entry(e, p, rinv):
    """One 40-byte circom Fr element: i32 short | u32 tag (bit31 long, bit30 Montgomery) | 32 B long."""
    tag = int.from_bytes(e[4:8], "little")
    if tag & 0x80000000:
        v = int.from_bytes(e[8:40], "little")
        return v * rinv % p if tag & 0x40000000 else v
    return int.from_bytes(e[0:4], "little", signed=True) % p


def signal_memory_base(ref, n_signals, model):
    """Find the wasm's signal memory (40 B per signal) by locating the values of main's first signals."""
    mem = ref.memory(0, 55 * 65536)
    p = ref.prime
    rinv = pow(1 << 256, -1, p)
    want = [model.b.vals[i].v % p for i in range(0, 24)]
    for base in range(0, len(mem) - 40 * n_signals, 4):
        if all(decode_entry(mem[base + 40 * i: base + 40 * i + 40], p, rinv) == v for i, v in enumerate(want)):
            return base
    raise RuntimeError("signal memory not found")


def decode_signals(ref, base, n):
    mem = ref.memory(base, 40 * n)
    p = ref.prime
    rinv = pow(1 << 256, -1, p)
    return [decode_entry(mem[40 * i: 40 * i + 40], p, rinv) for i in range(n)]


def build_variant(variant, trials, verbose=True):
    from oracle.ref_wasm import RefWasm
    ref = RefWasm(variant)
    ws = ref.witness_size
    w2s = np.frombuffer(ref.memory(W2S_OFFSET[variant], 4 * ws), np.uint32).astype(np.int64)
    rng = random.Random(0xB3B30000 + len(variant))
    descs = None
    base = None
    names = [nm for nm, _ in ref.plan]
    sizes = [sz for _, sz in ref.plan]
    n_checked = 0
    for trial in range(trials):
        inputs = random_inputs(variant, rng, edge=trial if trial < 7 else 0)
        mdl = model_for(variant, inputs, ref.prime)
        d = {}
        pos = 0
        for nm, sz in zip(names, sizes):
            d[nm] = inputs[pos:pos + sz]
            pos += sz
        rc, wit = ref.calculate(d)
        assert (rc == 0) == bool(mdl.ok), "%s trial %d: wasm rc=%d model ok=%s inputs=%r" % (variant, trial, rc, mdl.ok, inputs)
        if rc != 0:
            continue
        nsig = len(mdl.b.names)
        if base is None:
            base = signal_memory_base(ref, nsig, mdl)
        sig = decode_signals(ref, base, nsig)
        bad = [i for i in range(nsig) if sig[i] != mdl.b.vals[i].v % ref.prime]
        assert not bad, "%s trial %d: %d signal mismatches, first %s: wasm %d model %d" % (
            variant, trial, len(bad), mdl.b.names[bad[0]], sig[bad[0]], mdl.b.vals[bad[0]].v)
        dd = [desc_of(mdl.b.vals[s].s) for s in w2s]
        if descs is None:
            descs = dd
        assert dd == descs, "%s: descriptor table depends on the input (trial %d)" % (variant, trial)
        tmax = max(mdl.b.trace) + 1
        trace = [mdl.b.trace.get(i, 0) for i in range(tmax + 8)]
        got = expand(trace, descs, ref.prime)
        want = [int.from_bytes(wit[32 * i:32 * i + 32].tobytes(), "little") for i in range(ws)]
        assert got == want, "%s trial %d: expanded witness differs from the wasm's" % (variant, trial)
        n_checked += 1
    if verbose:
        print("%-14s witness %5d slots, %6d signals, signal memory @%d, %d inputs checked signal-by-signal + witness"
              % (variant, ws, nsig, base, n_checked))
    return dict(variant=variant, ws=ws, prime=ref.prime, descs=descs, w2s=[int(x) for x in w2s],
                n_inputs=ref.input_size, trace_words=tmax)


def emit(results, check_only):
    prod = ["/* GENERATED by tools/gen_tables.py -- do not edit.",
            " * Slot-descriptor tables: where each witness slot's value lives in the kernel's trace.",
            " * Witness order = the reference wasm's own witness->signal table (SURVEY.md 8(a) A7). */",
            "#pragma once", "#include <stdint.h>", "typedef struct { uint32_t desc0, count; int32_t delta; } b3w_seg;", ""]
    orc = ["/* GENERATED by tools/gen_tables.py -- do not edit.  TEST INFRASTRUCTURE (oracle).",
           " * witness slot -> circom signal index, run-length encoded {first signal, count} (consecutive signals). */",
           "#pragma once", "#include <stdint.h>", ""]
    for r in results:
        v = r["variant"]
        segs = rle(r["descs"])
        prod.append("#define B3W_WS_%s %du" % (v.upper(), r["ws"]))
        prod.append("#define B3W_TRACE_WORDS_%s %du" % (v.upper(), r["trace_words"]))
        prod.append("static const b3w_seg B3W_SEGS_%s[%d] = {" % (v.upper(), len(segs)))
        line = "  "
        for s in segs:
            item = "{0x%x,%d,%d}," % s
            if len(line) + len(item) > 118:
                prod.append(line)
                line = "  "
            line += item
        prod.append(line)
        prod.append("};")
        prod.append("")
        runs = []
        w2s = r["w2s"]
        i = 0
        while i < len(w2s):
            j = i
            while j + 1 < len(w2s) and w2s[j + 1] == w2s[j] + 1:
                j += 1
            runs.append((w2s[i], j - i + 1))
            i = j + 1
        orc.append("#define W2S_WS_%s %du" % (v.upper(), r["ws"]))
        orc.append("static const uint32_t W2S_RUNS_%s[%d][2] = {" % (v.upper(), len(runs)))
        line = "  "
        for s in runs:
            item = "{%d,%d}," % s
            if len(line) + len(item) > 118:
                orc.append(line)
                line = "  "
            line += item
        orc.append(line)
        orc.append("};")
        orc.append("")
    nv = ["/* GENERATED by tools/gen_tables.py from tools/circuit_model.py -- do not edit.",
          " * Trace indices (u32 words) of the nova step circuit's values; see circuit_model.py for their meaning. */",
          "#pragma once"]
    for k, v in cm.NV.items():
        nv.append("#define NV_%s %du" % (k, v))
    nv.append("#define NOVA_TRACE_WORDS %du" % cm.NOVA_TRACE_WORDS)
    outs = [(os.path.join(ROOT, "hot_proofs_blake3_circom_b200", "csrc", "nova_trace.h"), "\n".join(nv) + "\n"),
            (os.path.join(ROOT, "hot_proofs_blake3_circom_b200", "csrc", "slot_tables.h"), "\n".join(prod) + "\n"),
            (os.path.join(ROOT, "oracle", "w2s_tables.h"), "\n".join(orc) + "\n")]
    for path, text in outs:
        if check_only:
            same = os.path.exists(path) and open(path).read() == text
            print("%s: %s" % (os.path.relpath(path, ROOT), "up to date" if same else "DIFFERS"))
            if not same:
                sys.exit(1)
        else:
            with open(path, "w") as f:
                f.write(text)
            print("wrote %s (%d bytes)" % (os.path.relpath(path, ROOT), len(text)))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--check-only", action="store_true")
    ap.add_argument("--trials", type=int, default=8)
    ap.add_argument("--variants", default="compression,nova_bn_o2,nova_pasta_o2,nova_bn_o1")
    a = ap.parse_args()
    res = [build_variant(v, a.trials) for v in a.variants.split(",")]
    emit(res, a.check_only)
