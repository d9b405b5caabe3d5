/* blake3wit.h -- C ABI of libblake3wit.so: B200-native batched circom witness generation for the
 * BLAKE3 compression circuit and the blake3_nova / blake3_nova_pasta step circuits.
 *
 * This library replaces ONE path of banyancomputer/hot-proofs-blake3-circom: everything that happens
 * between `WitnessCalculator._doCalculateWitness` and the witness read-out, i.e. the circom-generated
 * wasm program and its Fr library (reference: blake3_nova_js/witness_calculator.js:131-272 and
 * build/ ** / *.wasm).  Plain pointers and sizes only; no torch / CUDA types in the signatures (a CUDA
 * stream is passed as an opaque void*).  All functions return 0 on success, a NEGATIVE library error
 * (B3W_ERR_*) or a POSITIVE circom runtime code (the `exceptionHandler` codes of
 * witness_calculator.js:21-39; only 4 = "Assert Failed." can be produced by a witness run).
 * The library never aborts and never falls back to a CPU implementation: without a usable CUDA device
 * every compute entry point fails with B3W_ERR_CUDA.
 */
#ifndef BLAKE3WIT_H
#define BLAKE3WIT_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B3W_VERSION 0x000200

#define B3W_OK 0
#define B3W_ERR_INVALID (-1)     /* bad argument */
#define B3W_ERR_CUDA (-2)        /* CUDA runtime error or no device; text in b3w_last_error() */
#define B3W_ERR_NOMEM (-3)
#define B3W_ERR_DOMAIN (-4)      /* b3w_inputs_from_fr: a value does not fit the u32 row format */
#define B3W_ERR_UNSUPPORTED (-5)
#define B3W_CIRCOM_ASSERT 4      /* "Assert Failed." (witness_calculator.js:29-30) */
#define B3W_R1CS_VIOLATION 7     /* per-instance status of the on-device R1CS check: some row has A.z * B.z != C.z */
#define B3W_NO_ROW 0xFFFFFFFFu   /* "no violated row" in first_bad[] */

#define B3W_FLAG_FUSED_CHECK 1u  /* b3w_config.flags: every batch call also runs the fused R1CS check */
#define B3W_FLAG_COMPRESSIBLE_RING 2u /* accepted for compatibility: the ring IS compressible by default since 0x000200 */
#define B3W_FLAG_PLAIN_RING 4u   /* b3w_config.flags: keep the internal HBM ring of the host-buffer calls in ordinary (cudaMalloc) memory.
                                    Default everywhere (C, Python, N-API): compressible memory when the driver grants it, silently
                                    ordinary memory otherwise */
#define B3W_FLAG_BYTE_CHECK 16u  /* b3w_config.flags: every host-buffer batch call (b3w_witness_batch, _ex, _fr, _fr_ex, _wide; also with out = NULL;
                                    b3w_nova_chain and, where a status array is given, b3w_nova_chain_device; NOT the packed / hybrid
                                    forms, whose witnesses are not expanded on the GPU) re-reads each
                                    chunk's witnesses from the HBM ring right after they were written and evaluates EVERY row of the
                                    constraint system on those bytes (the kernel of b3w_r1cs_check_device): what a consumer does
                                    with the vector it is handed (rust_fold/src/utils.rs:78-85), before the bytes leave the GPU.
                                    status[i] = B3W_R1CS_VIOLATION and first_bad[i] = the row (numbering of b3w_r1cs_check_device) on a
                                    miss; instances with "Assert Failed." keep that status.  Costs one read of the ring per chunk
                                    (about as long as writing it); needs a constraint system (built in for all four circuits). */
#define B3W_FLAG_REFERENCE_SIBLINGS 8u /* b3w_config.flags, b3w_nova_chain*: pick the sibling of every parent step exactly as the
                                    reference does (rust_fold/src/blake3_hash.rs:60-78, by bit of the chunk index) instead of the
                                    BLAKE3 tree's true sibling; see b3w_nova_chain */
#define B3W_MEM_COMPRESSIBLE 1u  /* b3w_device_alloc flags */
#define B3W_MAX_SAMPLES 1024u    /* b3w_batch_extras.n_samples */

/* Circuit variants = the reference's committed witness programs:
 *   COMPRESSION   build/blake3_compression/blake3_compression_js/blake3_compression.wasm (BN254, O1)
 *   NOVA_BN_O2    build/blake3_nova_js/blake3_nova.wasm                                  (BN254, O2)
 *   NOVA_PASTA_O2 build/blake3_nova_pasta_js/blake3_nova_pasta.wasm                      (Pallas scalar, O2)
 *   NOVA_BN_O1    build/blake3_nova/blake3_nova_js/blake3_nova.wasm (== build/blake3_nova_pasta/...) (BN254, O1) */
typedef enum {
  B3W_COMPRESSION = 0,
  B3W_NOVA_BN_O2 = 1,
  B3W_NOVA_PASTA_O2 = 2,
  B3W_NOVA_BN_O1 = 3
} b3w_circuit;

typedef struct {
  uint32_t circuit;        /* b3w_circuit */
  int32_t device;          /* CUDA device ordinal; -1 = current device */
  uint32_t chunk;          /* most instances per internal HBM ring slot for host-buffer batches; 0 = default (4096: two slots of
                              3.2 GB at full size).  The ring is created at the first batch call, sized for that call (a power of
                              two >= 64, <= chunk) and re-created when a larger batch arrives. */
  uint32_t flags;          /* 0 or an OR of B3W_FLAG_* */
} b3w_config;

/* What the WitnessCalculator constructor caches (witness_calculator.js:108-125). */
typedef struct {
  uint32_t witness_size;   /* getWitnessSize(): slots per witness, slot 0 is the constant 1 */
  uint32_t n_inputs;       /* getInputSize(): input values, in circuit declaration order */
  uint32_t n32;            /* getFieldNumLen32() = 8 */
  uint32_t n_public;       /* values copied to `pub`: 16 (compression out[16]) / 15 (nova z_{i+1}) */
  uint32_t version[3];     /* circom version the reference artefact was built with: 2,1,6 */
  uint8_t prime[32];       /* getRawPrime(), little-endian */
} b3w_info;

typedef struct b3w_ctx b3w_ctx;

int b3w_version(void);
const char *b3w_last_error(void);                       /* thread-local text of the last failure */

/* replaces builder() + new WitnessCalculator (witness_calculator.js:1-125): one ctx per GPU.
 * Ownership / threading: a b3w_ctx takes ONE call at a time.  The host-buffer entry points share the context's HBM ring,
 * streams and scratch; each of them holds the context's internal lock for its whole duration, so calls made concurrently
 * from several threads (e.g. overlapping `await`s of one calculator through the N-API addon's worker threads) are
 * serialised, never interleaved -- the reference's wasm calculator runs each call to completion too.  For parallelism use
 * one context per thread, or b3w_multi_*.  Every entry point runs on the context's device and restores the calling
 * thread's current CUDA device before it returns.  The *_device entry points are asynchronous on the caller's stream; at
 * most 32 of their launches may be in flight per context before a new launch waits for the oldest one. */
int b3w_create(const b3w_config *cfg, b3w_ctx **out);
void b3w_destroy(b3w_ctx *ctx);

/* Static circuit metadata; these three need no GPU (circuit = b3w_circuit). */
int b3w_circuit_info(uint32_t circuit, b3w_info *info);

/* the 76-byte .wtns header that calculateWTNSBin writes before the body (witness_calculator.js:214-262) */
int b3w_wtns_header(uint32_t circuit, uint8_t hdr[76]);

/* Name -> (offset, size) of an input signal inside the u32 input row; replaces getInputSignalSize
 * (witness_calculator.js:141).  Returns B3W_ERR_INVALID for an unknown name. */
int b3w_input_signal(uint32_t circuit, const char *name, uint32_t *offset, uint32_t *size);

/* The per-template trace that the reference appends to "Assert Failed.\n" when a constraint of the circuit fails
 * (printErrorMessage, witness_calculator.js:40-43; e.g. "Error in template Blake3NovaTreePath_CheckDepth_5 line: 38\n
 * Error in template Blake3Nova_54 line: 201\n").  in: n_inputs u32 (host).  Returns 0 (buf = "") when the input asserts
 * nowhere, else B3W_CIRCOM_ASSERT with the trace text in buf (truncated to cap).  Host-only, needs no GPU. */
int b3w_assert_trace(uint32_t circuit, const uint32_t *in, char *buf, size_t cap);

/* replaces _doCalculateWitness + calculateBinWitness for ONE input (witness_calculator.js:131-205).
 * in: n_inputs u32 (host).  out: witness_size*32 bytes (host), canonical little-endian Fr256.
 * Returns 0 or B3W_CIRCOM_ASSERT (out is then untouched beyond what was written). */
int b3w_witness_one(b3w_ctx *ctx, const uint32_t *in, uint8_t *out);

/* NEW batched entry point, HOST buffers.  in: n rows of n_inputs u32.  out: n*witness_size*32 bytes or
 * NULL (witnesses are then only streamed through the HBM ring: generate + compact results).
 * status: n bytes (0 ok, 4 assert) or NULL.  pub: n rows of n_public u32 or NULL.
 * Host buffers from b3w_host_alloc() are pinned and copy at full PCIe rate. */
int b3w_witness_batch(b3w_ctx *ctx, const uint32_t *in, uint64_t n, uint8_t *out, uint8_t *status, uint32_t *pub);

/* Optional extra results of a batch call (BASELINE config 5's compact-result contract, SURVEY.md 8(d): per instance the
 * public outputs, a status byte and a 64-bit checksum of the witness bytes, plus full witnesses for a <= 1 024-instance
 * sample).  Every pointer may be NULL.
 *   sums       n u64: the checksum b3w_checksum_device defines, computed by the kernel's expansion warps from the very
 *              values they store (no re-read of HBM); 0 for an instance that asserts (no witness exists).  What a consumer
 *              needs to know that the witness it later reads is the one that was generated and checked
 *              (rust_fold/src/utils.rs:78-85 enforces every row on the vector it was handed).
 *   sample_idx n_samples instance indices in [0, n), any order, n_samples <= B3W_MAX_SAMPLES: their full witnesses are
 *              copied out of the HBM ring into sample_out (n_samples * witness_size * 32 bytes), also when out == NULL.
 *   first_bad  n u32: with the fused check (B3W_FLAG_FUSED_CHECK) the smallest violated row id of the fused system or B3W_NO_ROW;
 *              with B3W_FLAG_BYTE_CHECK the verdict of the stand-alone checker on the stored bytes (its row numbering) instead. */
typedef struct {
  uint64_t *sums;
  const uint64_t *sample_idx;
  uint32_t n_samples;
  uint8_t *sample_out;
  uint32_t *first_bad;
} b3w_batch_extras;
int b3w_witness_batch_ex(b3w_ctx *ctx, const uint32_t *in, uint64_t n, uint8_t *out, uint8_t *status, uint32_t *pub,
                         const b3w_batch_extras *extras);

/* Per-call timing of the LAST host-buffer call of this context (b3w_witness_batch*, b3w_nova_chain, b3w_witness_batch_packed,
 * b3w_witness_batch_hybrid): the reference prints per-stage wall times (test/witness_gen.test.ts:43-50,
 * rust_fold/src/main.rs:167-178).  kernel_ms is measured with CUDA events around every witness-kernel launch of the call;
 * the stages also carry NVTX ranges ("b3w:h2d", "b3w:kernel", "b3w:d2h", "b3w:host_unpack"). */
typedef struct {
  double total_ms;         /* wall clock of the call */
  double kernel_ms;        /* sum over the call's witness-kernel launches (device time); launches on the two ring streams may
                              overlap, so 0 < kernel_ms <= 2 * total_ms */
  double host_ms;          /* host-side work inside the call (Fr conversion set-up, host unpack) */
  uint64_t launches;       /* witness-kernel launches */
  uint64_t instances;
  uint64_t h2d_bytes, d2h_bytes;
} b3w_timing;
int b3w_last_timing(b3w_ctx *ctx, b3w_timing *out);

/* Field-element inputs: n rows of n_inputs Fr256 (32 bytes little-endian each, any 256-bit value: reduced mod p like
 * `normalize`, witness_calculator.js:319-323) -- the form in which rust_fold holds them (`Vec<(String, Vec<F>)>`,
 * rust_fold/src/blake3_circuit.rs:197-289).  b3w_inputs_from_fr converts to the u32 rows of the other entry points
 * (host-only, needs no GPU); a value outside [0, 2^32) after reduction is refused with B3W_ERR_DOMAIN, naming the
 * instance and signal.  b3w_witness_batch_fr takes ANY field elements (see "The FULL input domain" below). */
int b3w_inputs_from_fr(uint32_t circuit, const uint8_t *in_fr, uint64_t n, uint32_t *rows);
int b3w_witness_batch_fr(b3w_ctx *ctx, const uint8_t *in_fr, uint64_t n, uint8_t *out, uint8_t *status, uint32_t *pub);
int b3w_witness_batch_fr_ex(b3w_ctx *ctx, const uint8_t *in_fr, uint64_t n, uint8_t *out, uint8_t *status, uint32_t *pub,
                            const b3w_batch_extras *extras);

/* The FULL input domain.  The reference takes any field element for every input (witness_calculator.js:319-323) and lets
 * the circuit decide.  b3w_witness_batch_fr (above) does the same for all four circuits: every input the reference accepts
 * gives the reference's witness, every other one status 4 ("Assert Failed."), and b3w_assert_trace_fr the text the wasm
 * prints for it.  u32 inputs -- everything the reference's drivers and tests produce -- stay on the hot kernels.
 *
 * blake3_compression: h, t, b, d pass through ToBits(32), so a value >= 2^32 asserts; the message
 * words are not range-checked (circuits/blake3_compression.circom:169-170) and only enter 34-bit sums, so e.g.
 * m[0] = 2^32 or m[0] = p - 1 give VALID witnesses.  A satisfying input has message words m = ext * 2^32 + lo, ext in [-2, 3]:
 *   b3w_inputs_from_fr_wide   Fr256 rows -> u32 rows (lo words) + m_ext (n x 16 int8).  Refuses nothing: an instance that
 *                             cannot satisfy the circuit gets m_ext[i][0] = B3W_EXT_ASSERT and comes back with status 4.
 *                             *n_wide (may be NULL) = instances that are not plain u32 rows.  Host-only, needs no GPU.
 *   b3w_witness_batch_wide / _device_wide   as b3w_witness_batch / b3w_witness_batch_device(_checked) with m_ext; the sums
 *                             that leave [0, 2^34) are detected on the device (status 4, no witness, pub = 0).  With
 *                             d_first_bad != NULL or B3W_FLAG_FUSED_CHECK the fused R1CS check runs as well.
 * nova step circuits: as built they constrain little (leaf_depth - depth in [1, 256], chunk_idx_low + 2^32 chunk_idx_high
 * < 2^65, and the embedded compression's ranges); n_blocks, block_count, total_depth, depth may be ANY field elements.  The
 * Fr256 rows are converted ON THE DEVICE (one warp per instance): u32 instances run on the hot kernel, and only the instances
 * that hold a value outside u32 run on a general kernel that evaluates the nova-level logic on field elements, into the
 * same output buffers (pub then holds the low 32 bits of the outputs).  With B3W_FLAG_FUSED_CHECK the u32 instances get the
 * fused check and the others the stand-alone evaluator on their finished witnesses; sums / samples work for both.
 *   b3w_assert_trace_fr       b3w_assert_trace for ONE Fr256 input row of any circuit: the reference's per-template trace of
 *                             the first failing constraint in the wasm's execution order, e.g. for b = 2^33 "Error in template
 *                             ToBits_3 line: 153\nError in template RotXorWordBits_5 line: 62\nError in template HalfFunG_18 line: 91\n...". */
#define B3W_EXT_ASSERT 127
int b3w_inputs_from_fr_wide(uint32_t circuit, const uint8_t *in_fr, uint64_t n, uint32_t *rows, int8_t *m_ext, uint64_t *n_wide);
int b3w_witness_batch_wide(b3w_ctx *ctx, const uint32_t *in, const int8_t *m_ext, uint64_t n, uint8_t *out, uint8_t *status,
                           uint32_t *pub);
int b3w_witness_batch_device_wide(b3w_ctx *ctx, const uint32_t *d_in, const int8_t *d_m_ext, uint64_t n, uint8_t *d_out,
                                  uint8_t *d_status, uint32_t *d_pub, uint32_t *d_first_bad, void *stream);
int b3w_assert_trace_fr(uint32_t circuit, const uint8_t *in_fr, char *buf, size_t cap);

/* NEW batched entry point, DEVICE buffers (same layouts), asynchronous on `stream` (a cudaStream_t, may
 * be NULL for the default stream).  d_out must hold n*witness_size*32 bytes, 32-byte aligned. */
int b3w_witness_batch_device(b3w_ctx *ctx, const uint32_t *d_in, uint64_t n, uint8_t *d_out, uint8_t *d_status,
                             uint32_t *d_pub, void *stream);

/* As b3w_witness_batch_device, plus the FUSED R1CS satisfiability check: before the witness of an instance is expanded,
 * every row of the circuit's constraint system (re-derived from the circom templates, the .r1cs files being absent from
 * the reference tree) is evaluated on the values the expansion is about to write, straight from shared memory -- nothing
 * is re-read from HBM.  d_status[i] = 0, B3W_CIRCOM_ASSERT or B3W_R1CS_VIOLATION; d_first_bad[i] (may be NULL) = the
 * smallest violated row id or B3W_NO_ROW.  What the reference's tests do with circom_tester's expectPass
 * (test/blake3_hash.test.ts:36) and bellpepper's enforce (rust_fold/src/utils.rs:78-85).
 * WHAT IT VALIDATES: the TRACE the witness is expanded from (the circuit's arithmetic: every sum, carry, rotation, flag).  Rows
 * that hold for any trace content by construction of the expansion (booleanity of an extracted bit, word = sum of its bits)
 * are not evaluated, so a wrong slot descriptor or a lost store is invisible to it; b3w_r1cs_check_device is the check
 * that reads the emitted bytes (B3W_FLAG_BYTE_CHECK chains it to every chunk of the host-buffer batch calls), and
 * b3w_batch_extras.sums ties streamed witnesses to what was stored. */
int b3w_witness_batch_device_checked(b3w_ctx *ctx, const uint32_t *d_in, uint64_t n, uint8_t *d_out, uint8_t *d_status,
                                     uint32_t *d_pub, uint32_t *d_first_bad, void *stream);
/* Everything at once, DEVICE buffers: d_m_ext (compression only, may be NULL) selects the wide-domain kernel; check != 0 or
 * d_first_bad != NULL adds the fused R1CS check; d_sums (n u64, may be NULL) receives the per-instance witness checksums of
 * b3w_batch_extras.sums, computed by the expansion warps from the values they store. */
int b3w_witness_batch_device_ex(b3w_ctx *ctx, const uint32_t *d_in, const int8_t *d_m_ext, uint64_t n, uint8_t *d_out,
                                uint8_t *d_status, uint32_t *d_pub, uint32_t *d_first_bad, uint64_t *d_sums, int check, void *stream);

/* Stand-alone R1CS check of witnesses that are RESIDENT IN DEVICE MEMORY (e.g. from another producer, or read back later):
 * sparse A.z * B.z - C.z over Fr, exact.  One CTA per instance streams the witness from HBM once into a compact
 * shared-memory copy and evaluates a program compiled from the constraint system (booleanity rows as bit masks, XOR rows as
 * bit-field compares, recompositions as runs: hot_proofs_blake3_circom_b200/csrc/kernels_r1cs_fast.cuh).  Built-in systems
 * exist for ALL FOUR circuits: the template-level rows of the O1 builds (blake3_compression 24 544 rows, NOVA_BN_O1 25 064)
 * and the O2-form system of NOVA_BN_O2 / NOVA_PASTA_O2 (23 743 rows over the 23 291 surviving wires; tools/gen_r1cs.py);
 * b3w_r1cs_load replaces them by a file's.  This is the check that LOOKS AT THE EMITTED BYTES -- a wrong slot descriptor or
 * a lost store shows here; the fused check of the *_checked kernels validates the trace the witness is expanded from.
 * d_first_bad[i] = smallest violated row, B3W_NO_ROW, or 0xFFFFFFFE when a slot is not a canonical field element. */
int b3w_r1cs_check_device(b3w_ctx *ctx, const uint8_t *d_wit, uint64_t n, uint8_t *d_status, uint32_t *d_first_bad,
                          void *stream);

/* Load the constraint system that b3w_r1cs_check_device evaluates from an iden3 `.r1cs` file (binary format v1) instead
 * of the built-in template-derived rows: the reference's own build/ *.r1cs where they exist (they are missing from the
 * reference tree, .MISSING_LARGE_BLOBS; rust_fold loads them at blake3_circuit.rs:71-81), or the equivalent systems
 * written by tools/export_r1cs.py -- which is how the two O2 builds get a stand-alone check.  The file's prime and wire
 * count must match the context's circuit.  Rows with coefficients that are not small integers are evaluated in Fr
 * (Montgomery), everything else in exact 128-bit integers with an Fr fallback, so ANY satisfied system is accepted and
 * any violated row reported; d_first_bad then holds the constraint's index in the file.  n_rows (may be NULL)
 * receives the number of constraints. */
int b3w_r1cs_load(b3w_ctx *ctx, const uint8_t *r1cs, size_t len, uint32_t *n_rows);
int b3w_r1cs_load_file(b3w_ctx *ctx, const char *path, uint32_t *n_rows);
/* How the context's system (built-in or loaded) was split: rows in all, rows covered by the compiled program (the others
 * stay with the general row evaluator), XOR runs and row tiles of the program.  Any pointer may be NULL. */
int b3w_r1cs_program_info(b3w_ctx *ctx, uint32_t *n_rows, uint32_t *n_compiled, uint32_t *n_xor_runs, uint32_t *n_tiles);
/* ditto for a circuit's BUILT-IN system, host-only (needs no GPU); n_items = 16-byte items of the row tiles. */
int b3w_r1cs_compile_stats(uint32_t circuit, uint32_t *n_rows, uint32_t *n_compiled, uint32_t *n_xor_runs, uint32_t *n_tiles,
                           uint32_t *n_items);
/* the same and more, as an array (the first n_out of): rows, compiled rows, XOR runs, row tiles, items, VIRTUAL BITS (linear
 * combinations that circom's O2 pass put in place of a bit, evaluated once per witness into a bit slot of their own --
 * 0 for the O1-form systems), tiles summed in plain 64-bit arithmetic, and tiles / items of the program compiled without
 * virtual bits, which a witness whose virtual bits are not bits is evaluated with (0 / 0 when there are none). */
int b3w_r1cs_compile_stats_ex(uint32_t circuit, uint32_t *out, uint32_t n_out);

/* rows / non-zero terms of the circuit's template-level constraint system in O1 form
 * (blake3_compression: 24 544 rows = 23 376 quadratic + 1 168 linear; nova: 25 064) */
int b3w_r1cs_info(uint32_t circuit, uint32_t *n_rows, uint32_t *n_terms);

/* Test hook, host-only (no GPU, no context): compile the constraint system of an `.r1cs` file exactly as b3w_r1cs_load
 * does and copy one table of the resulting program to `out` (n_bytes receives its size; nothing is copied when cap is too
 * small).  plain != 0: the program compiled WITHOUT virtual bits (the kernel's fall-back).  Sections: 0 booleanity mask,
 * 1 booleanity row ids, 2 XOR runs, 3 XOR row ids, 4 row tiles, 5 virtual-bit groups, 6 items, 7 row ids of the tiles,
 * 8 one byte per file row: covered by the program (1) or left to the general evaluator (0).  tests/test_r1cs_program.py
 * evaluates these tables with Python integers against the file's own rows. */
int b3w_debug_r1cs_program(const uint8_t *r1cs, size_t len, const uint8_t prime[32], uint32_t n_wires, int plain, uint32_t section,
                           void *out, size_t cap, size_t *n_bytes);
/* Test hook: the side-table layout of the stand-alone checker for a circuit (host only, no GPU): rank_out[w] (may be NULL;
 * cap words of room) = entries reserved before 32-slot word w by the circuit's slot kinds, *n_words = slot words + n_vtiles + 1,
 * *total = entries in all (even).  The kernel's shared-memory copy of a witness is laid out by this table. */
int b3w_debug_side_layout(uint32_t circuit, uint32_t n_vtiles, uint32_t *rank_out, uint32_t cap, uint32_t *n_words, uint32_t *total);
/* Test hook for the fused check: xor `xor_mask` into trace word `trace_word` of every instance after the trace phase
 * of the *_checked kernels (B3W_NO_ROW disables it). */
int b3w_debug_inject_fault(b3w_ctx *ctx, uint32_t trace_word, uint32_t xor_mask);

/* Tuning hook: cap the persistent grid at `ctas_per_sm` resident CTAs per SM (0 = the per-kernel default) and set the
 * number of work items an instance's witness is split into (0 = default; the items are scheduled dynamically). */
int b3w_debug_set_launch(b3w_ctx *ctx, int ctas_per_sm, uint32_t parts);
/* Tuning hook: how the expansion writes witness bytes.  0 = one 256-bit store per slot straight from registers (default);
 * 1 = expanded tiles staged in shared memory and written by the TMA engine (cp.async.bulk shared -> global), plain
 * u32-input kernels only. */
int b3w_debug_set_store_mode(b3w_ctx *ctx, int mode);

/* Per-instance 64-bit checksum of witnesses resident in device memory (reads them back from HBM):
 *   sum_i = SUM over slots s, limbs j of  (limb64[s][j] + 1) * mix(4*s + j)   (mod 2^64),
 *   mix(x) = (x + 1) * 0x9E3779B97F4A7C15  (mod 2^64).  d_sums: n u64. */
int b3w_checksum_device(b3w_ctx *ctx, const uint8_t *d_wit, uint64_t n, uint64_t *d_sums, void *stream);

/* Pure-store calibration kernel (the write roofline of this GPU): fills `bytes` of device memory with
 * 256-bit streaming stores; asynchronous on `stream`. */
int b3w_calib_fill(b3w_ctx *ctx, uint8_t *d_buf, uint64_t bytes, void *stream);
/* ditto with the store stream of the witness kernels (32 KiB work items from the dynamic counters, 1 KiB per warp store,
 * same grid): the ceiling of the store path for that access pattern, with none of the kernels' work in the way. */
int b3w_calib_fill_items(b3w_ctx *ctx, uint8_t *d_buf, uint64_t bytes, void *stream);
/* ditto through the TMA engine: bulk copies (cp.async.bulk shared -> global) of a constant shared-memory tile, one per
 * item; what "stage tiles in shared memory + bulk stores" could reach at best. */
int b3w_calib_fill_bulk(b3w_ctx *ctx, uint8_t *d_buf, uint64_t bytes, void *stream);

/* Chained-chunk driver (BASELINE config 3): all Nova step witnesses of a file, the batched form of the reference's
 * rust_fold loop (main.rs:71-94,166-171 around blake3_circuit.rs:160-290: update_for_step / format_input / synthesize).
 * For every 1024-byte chunk c the schedule is n_blocks leaf steps + (depth of the chunk in the BLAKE3 tree) parent steps,
 * z0 = [n_blocks, 0, IV, total_depth, leaf_depth-1, chunk_idx_low, chunk_idx_high, leaf_depth]; the chaining between steps and
 * the sibling chaining values (bao slice extraction in rust_fold/src/blake3_hash.rs:17-93) are computed on the device.
 * ctx must be one of the nova circuits.  Buffers are HOST buffers sized from b3w_nova_chain_size():
 *   out      total_steps * witness_size * 32 bytes, or NULL (witnesses only stream through the HBM ring)
 *   status   total_steps bytes or NULL;   pub  total_steps * 15 u32 (z_{i+1} of every step) or NULL
 *   rows     total_steps * 32 u32: the step inputs in circuit declaration order, or NULL
 *   step_off n_chunks + 1 u64: first step of each chunk, or NULL
 *   root     32 bytes: h_out of the last step of chunk 0 (= the BLAKE3 hash of the input for tree shapes on which the
 *            circuit's left/right rule, bit i of chunk_idx, agrees with the BLAKE3 tree: always for 2^k chunks).
 * Non-power-of-two chunk counts.  The circuit orders (h, sibling) by bit `total_depth - 2 - depth` of chunk_idx, which is the
 * real direction only where the BLAKE3 tree is perfect above the chunk.  DEFAULT here: the sibling of every parent step is the
 * TRUE BLAKE3 sibling (what a bao slice proves); chunks on the right spine of an imperfect tree then do not fold to
 * blake3(file) (tests/test_gpu_chain.py records which: e.g. the last chunk of 3, 5, 6, 7).  This deliberately DIFFERS from the
 * reference, whose hash_with_path (rust_fold/src/blake3_hash.rs:60-78) also reads the direction off the chunk index and takes
 * the parent node's other half by THAT direction -- for chunk 4 of 5 the chunk's own chaining value.
 * B3W_FLAG_REFERENCE_SIBLINGS in b3w_config.flags restates that rule literally (pinned against oracle/nova_chain_ref.py's
 * restatement of that file); both rules agree on every 2^k-chunk file, the only shape rust_fold's tests assert on.
 * b3w_nova_chain_device: the same driver with the step witnesses left IN DEVICE MEMORY -- d_out (total_steps * witness_size *
 * 32 bytes, 32-byte aligned, required), d_status / d_pub / d_rows (may be NULL) are caller-supplied DEVICE buffers, all step
 * witnesses come from one launch and only the file crosses PCIe: the hand-off for a prover on the same GPU.  step_off and
 * root are host arrays as above.  Returns after the device work has finished. */
int b3w_nova_chain_size(uint64_t len, uint64_t *n_chunks, uint64_t *total_steps);
int b3w_nova_chain(b3w_ctx *ctx, const uint8_t *data, uint64_t len, uint8_t *out, uint8_t *status, uint32_t *pub,
                   uint32_t *rows, uint64_t *step_off, uint8_t root[32]);
int b3w_nova_chain_device(b3w_ctx *ctx, const uint8_t *data, uint64_t len, uint8_t *d_out, uint8_t *d_status, uint32_t *d_pub,
                          uint32_t *d_rows, uint64_t *step_off, uint8_t root[32]);

/* COMPACT ("packed") witnesses.  Every slot of a witness is a pure function of the instance's trace (the few hundred u32
 * values the circuit really computes: sums with carries, rotated words, step flags) and of the static slot table, so the
 * trace is the witness in compact form: b3w_packed_words() u32 per instance = 3 776 B (blake3_compression) / 5 296 B
 * (nova) instead of 770 976 / 745 312 B.  What the .wtns writer of witness_calculator.js:208-272 would store per
 * instance becomes a lazy export: b3w_unpack_device() expands packed witnesses that are resident in device memory
 * into the .wtns body layout, bit-identical to b3w_witness_batch_device().  An instance whose inputs assert is marked
 * in status[] (and has word 1 of its packed record = 0 instead of 1); unpacking it yields no valid witness.
 * d_packed / packed: n * b3w_packed_words() u32, 16-byte aligned. */
int b3w_packed_words(uint32_t circuit, uint32_t *words);
int b3w_witness_batch_packed_device(b3w_ctx *ctx, const uint32_t *d_in, uint64_t n, uint32_t *d_packed, uint8_t *d_status,
                                    uint32_t *d_pub, void *stream);
int b3w_witness_batch_packed(b3w_ctx *ctx, const uint32_t *in, uint64_t n, uint32_t *packed, uint8_t *status, uint32_t *pub);
int b3w_unpack_device(b3w_ctx *ctx, const uint32_t *d_packed, uint64_t n, uint8_t *d_out, void *stream);
/* The same expansion on the HOST: packed records (host memory) -> .wtns bodies (host memory), bit-identical to
 * b3w_unpack_device; `threads` host threads (0 = all hardware threads), non-temporal stores.  A format conversion of records
 * the GPU computed -- what witness_calculator.js:263-269 does slot by slot -- not a CPU witness path: there is none. */
int b3w_unpack_host(b3w_ctx *ctx, const uint32_t *packed, uint64_t n, uint8_t *out, uint32_t threads);
/* HYBRID export: as b3w_witness_batch (every .wtns body ends up in the caller's host buffer `out`, which need not be
 * pinned), but only the 3.8 / 5.3 KB packed records cross PCIe and the 200x expansion runs on `threads` host threads while
 * the next chunk is computed and copied.  Bound by the host's memory write bandwidth instead of PCIe. */
int b3w_witness_batch_hybrid(b3w_ctx *ctx, const uint32_t *in, uint64_t n, uint8_t *out, uint8_t *status, uint32_t *pub,
                             uint32_t threads);

/* Multi-GPU form of b3w_witness_batch (BASELINE config 5; north_star: "each GPU fills its own slice of the host-pinned
 * output").  Witnesses are independent, so the batch [0, n) is cut into contiguous index ranges, one per device, each
 * driven by its own host thread and context; there is NO collective and no peer traffic.  b3w_shard_range() tells which
 * range shard g of n_shards gets (the first n % n_shards shards hold one extra instance; host-only, needs no GPU).  devices == NULL / n_devices == 0 = every
 * visible device; cfg->device is ignored.  A device may be listed more than once (each entry gets its own context). */
typedef struct b3w_multi b3w_multi;
int b3w_multi_create(const b3w_config *cfg, const int32_t *devices, uint32_t n_devices, b3w_multi **out);
void b3w_multi_destroy(b3w_multi *m);
uint32_t b3w_multi_size(const b3w_multi *m);
int b3w_shard_range(uint64_t n, uint32_t g, uint32_t n_shards, uint64_t *first, uint64_t *count);
int b3w_multi_witness_batch(b3w_multi *m, const uint32_t *in, uint64_t n, uint8_t *out, uint8_t *status, uint32_t *pub);
/* ditto with the extra results of b3w_witness_batch_ex (sums / first_bad are sliced like status; every sample is fetched by
 * the device that owns its instance): BASELINE config 5's streamed form -- out == NULL, fused check, sums, <= 1 024 samples. */
int b3w_multi_witness_batch_ex(b3w_multi *m, const uint32_t *in, uint64_t n, uint8_t *out, uint8_t *status, uint32_t *pub,
                               const b3w_batch_extras *extras);
/* b3w_nova_chain over several GPUs: the unit of sharding is a chunk (its steps chain into each other, chunks do not);
 * every device hashes the whole BLAKE3 tree (cheap) and generates the step witnesses of a contiguous range of chunks,
 * ranges balanced by step count, each writing its slice of the caller's arrays.  Same arguments as b3w_nova_chain. */
int b3w_multi_nova_chain(b3w_multi *m, const uint8_t *data, uint64_t len, uint8_t *out, uint8_t *status, uint32_t *pub,
                         uint32_t *rows, uint64_t *step_off, uint8_t root[32]);

/* Device memory for witness buffers, optionally COMPRESSIBLE (B3W_MEM_COMPRESSIBLE): Blackwell's L2 compresses lines on
 * their way to HBM when the memory was allocated for it, which cudaMalloc never does.  A witness (a bit or a word plus 24+
 * zero bytes per 32-byte slot) is the ideal payload: on B200 the blake3_compression kernel writes 2^16 witnesses into such
 * a buffer in 6.18 ms instead of 6.99 ms (10.6 M witnesses/s), and wide coalesced reads of it run at 8.9 TB/s instead of
 * 6.6 TB/s (narrow reads, 8 bytes per lane, are slower than on ordinary memory).  The bytes read back are identical.
 * *granted (may be NULL) tells whether the driver really made the block compressible; a device without the feature gets
 * ordinary memory.  Blocks belong to the context and are released by b3w_device_free or b3w_destroy. */
int b3w_device_alloc(b3w_ctx *ctx, size_t bytes, uint32_t flags, void **out, uint32_t *granted);
int b3w_device_free(b3w_ctx *ctx, void *p);

/* pinned host memory for batch buffers */
void *b3w_host_alloc(size_t bytes);
/* ditto, placed on the NUMA node the given CUDA device is attached to (-1 = current device): keeps the D2H stream of
 * each GPU of a multi-socket box off the socket interconnect. */
void *b3w_host_alloc_near(size_t bytes, int device);
void b3w_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
