#!/usr/bin/env node
// generate_witness.js -- command line front end of the GPU witness calculator.
//
// Keeps the calling convention of the circom-generated script it replaces (reference: <circuit>_js/generate_witness.js:4-18):
//     node generate_witness.js <file.wasm> <input.json> <output.wtns>
// three positional arguments, the usage line on any other count, the .wtns image of calculateWTNSBin(input, 0) written
// to the third path.  The .wasm file is only read to identify the circuit (witness_calculator.js hashes it); the witness
// comes from libblake3wit.so through the N-API addon.
"use strict";

const fs = require("fs/promises");
const buildCalculator = require("./witness_calculator.js");

const USAGE = "Usage: node generate_witness.js <file.wasm> <input.json> <output.wtns>";

async function main(argv) {
    const positional = argv.slice(2);
    if (positional.length !== 3) {
        console.log(USAGE);
        return 0;
    }
    const [wasmPath, inputPath, wtnsPath] = positional;
    const [program, inputText] = await Promise.all([fs.readFile(wasmPath), fs.readFile(inputPath, "utf8")]);
    const calculator = await buildCalculator(program);
    const wtns = await calculator.calculateWTNSBin(JSON.parse(inputText), 0);
    await fs.writeFile(wtnsPath, wtns);
    return 0;
}

main(process.argv).then(
    (code) => { process.exitCode = code; },
    (err) => { console.error(err && err.message ? err.message : err); process.exitCode = 1; }
);
