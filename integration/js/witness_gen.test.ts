// witness_gen.test.ts -- the reference's test/witness_gen.test.ts rewritten as a GPU-vs-wasm cross-check:
// the same LCG(6429) input goes through (a) the original wasm witness_calculator and (b) the GPU-backed drop-in;
// the two .wtns images must be byte-identical, and equal to the committed golden file.
import { LCG, genRandomChunk } from "./utils";
import { readFileSync } from "fs";
import chai from "chai";

const assert = chai.assert;
const wasmPath = "build/blake3_compression/blake3_compression_js/blake3_compression.wasm";

describe("blake3 compression witness: GPU (libblake3wit) vs wasm", function () {
  this.timeout(10000);

  it("single witness is byte-identical to the wasm calculator and to the golden .wtns", async () => {
    const code = readFileSync(wasmPath);
    const inp = genRandomChunk(new LCG(6429));
    const gpu = await require("../integration/js/witness_calculator.js")(code);
    const ref = await require("../blake3_nova_js/witness_calculator.js")(code);
    let start = Date.now();
    const a = await gpu.calculateWTNSBin(inp, 0);
    console.log("GPU witness generation takes ", Date.now() - start, "ms");
    start = Date.now();
    const b = await ref.calculateWTNSBin(inp, 0);
    console.log("wasm witness generation takes ", Date.now() - start, "ms");
    assert.deepEqual(Buffer.from(a), Buffer.from(b));
    assert.deepEqual(Buffer.from(a), readFileSync("build/blake3_compression/testInp/witness.wtns"));
  });

  it("a batch equals the wasm calculator instance by instance", async () => {
    const code = readFileSync(wasmPath);
    const lcg = new LCG(6429);
    const inputs = Array(64).fill(0).map(() => genRandomChunk(lcg));
    const gpu = await require("../integration/js/witness_calculator.js")(code);
    const ref = await require("../blake3_nova_js/witness_calculator.js")(code);
    const res = await gpu.calculateWitnessBatch(inputs);
    const ws = gpu.witnessSize * 32;
    for (let i = 0; i < inputs.length; i++) {
      const w = await ref.calculateBinWitness(inputs[i], 0);
      assert.deepEqual(Buffer.from(res.witness.subarray(i * ws, (i + 1) * ws)), Buffer.from(w));
      assert.equal(res.status[i], 0);
    }
  });
});
