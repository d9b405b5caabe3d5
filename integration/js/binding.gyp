{
  # node-gyp build description of the N-API addon.  REPO = the checkout of this repository (libblake3wit.so is built by
  # `python -c "import __graft_entry__ as g; g.build()"` into REPO/hot_proofs_blake3_circom_b200/).
  "variables": { "REPO%": "<!(node -p \"require('path').resolve(__dirname, '..', '..')\")" },
  "targets": [
    {
      "target_name": "blake3wit_napi",
      "sources": [ "addon/blake3wit_napi.cc" ],
      "include_dirs": [ "<(REPO)/include" ],
      "cflags_cc": [ "-std=c++17", "-O2" ],
      "libraries": [
        "-L<(REPO)/hot_proofs_blake3_circom_b200",
        "-lblake3wit",
        "-Wl,-rpath,<(REPO)/hot_proofs_blake3_circom_b200"
      ]
    }
  ]
}
