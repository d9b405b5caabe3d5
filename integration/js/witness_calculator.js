// witness_calculator.js -- drop-in replacement for the reference's witness_calculator.js (same exports, same method
// names and error texts), backed by libblake3wit.so through the N-API addon in addon/blake3wit_napi.cc.
//
// builder(code, options) identifies the circuit from the wasm bytes (sha256 of the reference's committed
// programs).  A wasm it does not know falls through to the ORIGINAL WebAssembly implementation, which the caller
// keeps as ./witness_calculator.wasm.js (the unmodified reference file); so does options.forceWasm.
//
// Input domain: every value the reference takes.  u32 inputs (all that the reference's drivers produce) go to the hot
// kernels as one u32 row; anything else goes to the library as Fr256 (b3w_witness_batch_fr), which answers with the
// reference's witness or the reference's "Assert Failed." text.
//
// NOTE: Node is not available in the build image of this repository, so this file is reviewed, not executed,
// there; hot_proofs_blake3_circom_b200/witness_calculator.py mirrors it line by line and IS exercised by the tests.
const crypto = require("crypto");

const CIRCUITS = {
    "6faf23ddfd697bbb7e8e922577589c2c06486258968a5a14f96fb5a16091b142": 0, // build/blake3_compression/..._js/blake3_compression.wasm
    "020bd11f289864c54c7d02cd05723dcf8323e31fa5c77d8700c618232685978e": 1, // build/blake3_nova_js/blake3_nova.wasm
    "b982f960ebbfcabe957fe13857ea47adfeee30e18fbe05474e9b982eab187f46": 2, // build/blake3_nova_pasta_js/blake3_nova_pasta.wasm
    "8d6317b72eab34d34e12dfd7bd310dce40f4190768669772f992a9510c441fca": 3, // build/blake3_nova{,_pasta}/..._js/*.wasm (O1)
};

module.exports = async function builder(code, options) {
    options = options || {};
    const circuit = CIRCUITS[crypto.createHash("sha256").update(code).digest("hex")];
    if (circuit === undefined || options.forceWasm) {
        return require("./witness_calculator.wasm.js")(code, options);      // the reference path, untouched
    }
    const addon = require("./build/Release/blake3wit_napi.node");
    // options.fusedCheck: the fused on-device R1CS check with every batch (B3W_FLAG_FUSED_CHECK); options.byteCheck: every chunk
    // is re-read from HBM and all rows are evaluated on its bytes before results leave the GPU (B3W_FLAG_BYTE_CHECK)
    const flags = (options.fusedCheck ? 1 : 0) | (options.byteCheck ? 16 : 0);
    return new WitnessCalculator(addon, addon.create(circuit, options.device === undefined ? -1 : options.device, flags), circuit, options);
};

class WitnessCalculator {
    constructor(addon, ctx, circuit, sanityCheck) {
        this.addon = addon;
        this.instance = ctx;                       // the reference keeps the wasm instance here
        this.circuit = circuit;
        const info = addon.circuitInfo(circuit);   // b3w_circuit_info
        this.version = info.version[0];
        this.n32 = info.n32;
        this.prime = BigInt("0x" + Buffer.from(info.prime).reverse().toString("hex"));
        this.witnessSize = info.witnessSize;
        this.nInputs = info.nInputs;
        this.sanityCheck = sanityCheck;
    }

    circom_version() {
        return this.version;
    }

    // witness_calculator.js:131-169 -- same checks, same messages; -> the nInputs values in declaration order (BigInt)
    _values(input) {
        const vals = new Array(this.nInputs).fill(0n);
        let input_counter = 0;
        Object.keys(input).forEach((k) => {
            const fArr = flatArray(input[k]);
            const sig = this.addon.inputSignal(this.circuit, k);          // {offset, size}; null for unknown names
            const signalSize = sig ? sig.size : 0;                        // the wasm's getInputSignalSize returns 0 then
            if (signalSize < 0) throw new Error(`Signal ${k} not found\n`);
            if (fArr.length < signalSize) throw new Error(`Not enough values for input signal ${k}\n`);
            if (fArr.length > signalSize) throw new Error(`Too many values for input signal ${k}\n`);
            for (let i = 0; i < fArr.length; i++) {
                vals[sig.offset + i] = normalize(fArr[i], this.prime);
                input_counter++;
            }
        });
        if (input_counter < this.nInputs) {
            throw new Error(`Not all inputs have been set. Only ${input_counter} out of ${this.nInputs}`);
        }
        return vals;
    }

    _row(vals) {
        const row = new Uint32Array(this.nInputs);
        vals.forEach((v, i) => { row[i] = Number(v); });
        return row;
    }

    _fr(vals) {                                    // nInputs x 32 bytes, little-endian
        const fr = new Uint8Array(vals.length * 32);
        vals.forEach((v, i) => { for (let j = 0; j < 32; j++) { fr[32 * i + j] = Number(v & 0xffn); v >>= 8n; } });
        return fr;
    }

    async _bin(input) {
        const vals = this._values(input);
        const u32 = vals.every((v) => v >> 32n === 0n);
        if (this.circuit !== 0) console.log("D_FLAGS:  0");                 // circuits/blake3_nova.circom:166
        if (!u32) return await this.addon.witnessOneFr(this.instance, this._fr(vals));   // b3w_witness_batch_fr
        return await this.addon.witnessOne(this.instance, this._row(vals)); // napi_async_work around b3w_witness_batch
    }

    async calculateWitness(input, sanityCheck) {
        const buff = await this._bin(input);
        const w = [];
        const b32 = new Uint32Array(buff.buffer, buff.byteOffset, this.witnessSize * this.n32);
        for (let i = 0; i < this.witnessSize; i++) {
            const arr = new Uint32Array(this.n32);
            for (let j = 0; j < this.n32; j++) arr[this.n32 - 1 - j] = b32[i * this.n32 + j];
            w.push(fromArray32(arr));
        }
        return w;
    }

    async calculateBinWitness(input, sanityCheck) {
        return await this._bin(input);
    }

    async calculateWTNSBin(input, sanityCheck) {
        const body = await this._bin(input);
        const hdr = this.addon.wtnsHeader(this.circuit);                   // b3w_wtns_header: the 76 bytes of :214-262
        const buff = new Uint8Array(hdr.length + body.length);
        buff.set(hdr, 0);
        buff.set(body, hdr.length);
        return buff;
    }

    // NEW: inputs = array of input objects, or {rows: Uint32Array(n * nInputs), n}.
    // opts.witness === false keeps the witnesses on the GPU and returns only status + public outputs.
    async calculateWitnessBatch(inputs, opts) {
        opts = opts || {};
        // opts.witness === false: stream only (status + pub come back); opts.sums: per-instance 64-bit witness checksums too
        if (!Array.isArray(inputs)) {
            return await this.addon.witnessBatch(this.instance, inputs.rows, inputs.n, opts.witness !== false, !!opts.sums);   // {witness, status, pub[, sums]}
        }
        const n = inputs.length;
        const vals = inputs.map((inp) => this._values(inp));
        const u32 = vals.every((v) => v.every((x) => x >> 32n === 0n));
        if (!u32) return await this.addon.witnessBatchFr(this.instance, this._fr(vals.flat()), n, opts.witness !== false, !!opts.sums);
        const rows = new Uint32Array(n * this.nInputs);
        vals.forEach((v, i) => rows.set(this._row(v), i * this.nInputs));
        return await this.addon.witnessBatch(this.instance, rows, n, opts.witness !== false, !!opts.sums);
    }
}

function fromArray32(arr) {
    let res = BigInt(0);
    const radix = BigInt(0x100000000);
    for (let i = 0; i < arr.length; i++) res = res * radix + BigInt(arr[i]);
    return res;
}

function flatArray(a) {
    const res = [];
    (function fill(x) { if (Array.isArray(x)) x.forEach(fill); else res.push(x); })(a);
    return res;
}

function normalize(n, prime) {
    let res = BigInt(n) % prime;
    if (res < 0) res += prime;
    return res;
}
