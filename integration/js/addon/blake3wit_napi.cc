// blake3wit_napi.cc -- N-API addon: the thin layer between witness_calculator.js and libblake3wit.so.
// Build (where node-gyp and node_api.h exist; they do not in this repository's build image):
//   node-gyp configure build   with binding.gyp { sources: [addon/blake3wit_napi.cc], include_dirs: [<repo>/include],
//                                                  libraries: [-L<repo>/hot_proofs_blake3_circom_b200 -lblake3wit] }
// Every compute call runs in napi_async_work so that `await` keeps the event loop free, as the async methods of the
// reference's WitnessCalculator promise (witness_calculator.js:171,190,208).
// Thread safety: libuv runs the work items of overlapping awaits (Promise.all over one calculator) on several worker
// threads at once.  A b3w_ctx takes one host-buffer call at a time and SERIALISES concurrent callers itself (an internal
// lock held for the whole call, include/blake3wit.h "Ownership / threading"), so two jobs on one calculator run one
// after the other and never see each other's ring slots -- like the reference's wasm calculator, which runs each call to
// completion on the JS thread.  tests/test_gpu_extras.py::test_one_context_serialises_concurrent_callers pins it.
#include <node_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "blake3wit.h"

#define NAPI_OK(call) do { if ((call) != napi_ok) { napi_throw_error(env, NULL, #call); return NULL; } } while (0)

static napi_value throw_b3w(napi_env env, int rc) {
  napi_throw_error(env, NULL, rc == B3W_CIRCOM_ASSERT ? "Error: Assert Failed.\n" : b3w_last_error());
  return NULL;
}

// create(circuit, device[, flags]) -> external(handle)      flags: an OR of B3W_FLAG_* (fused check 1, byte check 16, ...)
struct handle { b3w_ctx *ctx; uint32_t circuit; };
static void ctx_finalize(napi_env, void *data, void *) { b3w_destroy(((handle *)data)->ctx); free(data); }
static napi_value Create(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value argv[3];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  uint32_t circuit, flags = 0; int32_t device;
  NAPI_OK(napi_get_value_uint32(env, argv[0], &circuit));
  NAPI_OK(napi_get_value_int32(env, argv[1], &device));
  if (argc >= 3) NAPI_OK(napi_get_value_uint32(env, argv[2], &flags));
  // flags 0 = the library's one default everywhere: the HBM ring of the host-buffer calls is compressible memory where the
  // GPU offers it and silently ordinary memory otherwise (B3W_FLAG_PLAIN_RING opts out)
  b3w_config cfg = {circuit, device, 0, flags};
  b3w_ctx *ctx = NULL;
  int rc = b3w_create(&cfg, &ctx);
  if (rc) return throw_b3w(env, rc);
  handle *h = (handle *)malloc(sizeof(handle));
  h->ctx = ctx;
  h->circuit = circuit;
  napi_value ext;
  NAPI_OK(napi_create_external(env, h, ctx_finalize, NULL, &ext));
  return ext;
}

// circuitInfo(circuit) -> {witnessSize, nInputs, n32, nPublic, version:[3], prime:Uint8Array(32)}
static napi_value CircuitInfo(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value argv[1];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  uint32_t circuit;
  NAPI_OK(napi_get_value_uint32(env, argv[0], &circuit));
  b3w_info bi;
  int rc = b3w_circuit_info(circuit, &bi);
  if (rc) return throw_b3w(env, rc);
  napi_value obj, v, arr, ab;
  NAPI_OK(napi_create_object(env, &obj));
  NAPI_OK(napi_create_uint32(env, bi.witness_size, &v)); NAPI_OK(napi_set_named_property(env, obj, "witnessSize", v));
  NAPI_OK(napi_create_uint32(env, bi.n_inputs, &v));     NAPI_OK(napi_set_named_property(env, obj, "nInputs", v));
  NAPI_OK(napi_create_uint32(env, bi.n32, &v));          NAPI_OK(napi_set_named_property(env, obj, "n32", v));
  NAPI_OK(napi_create_uint32(env, bi.n_public, &v));     NAPI_OK(napi_set_named_property(env, obj, "nPublic", v));
  NAPI_OK(napi_create_array_with_length(env, 3, &arr));
  for (uint32_t i = 0; i < 3; i++) { NAPI_OK(napi_create_uint32(env, bi.version[i], &v)); NAPI_OK(napi_set_element(env, arr, i, v)); }
  NAPI_OK(napi_set_named_property(env, obj, "version", arr));
  void *p;
  NAPI_OK(napi_create_arraybuffer(env, 32, &p, &ab));
  memcpy(p, bi.prime, 32);
  NAPI_OK(napi_create_typedarray(env, napi_uint8_array, 32, ab, 0, &v));
  NAPI_OK(napi_set_named_property(env, obj, "prime", v));
  return obj;
}

// inputSignal(circuit, name) -> {offset, size} | null       (b3w_input_signal; replaces getInputSignalSize)
static napi_value InputSignal(napi_env env, napi_callback_info info) {
  size_t argc = 2; napi_value argv[2];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  uint32_t circuit; char name[128]; size_t len;
  NAPI_OK(napi_get_value_uint32(env, argv[0], &circuit));
  NAPI_OK(napi_get_value_string_utf8(env, argv[1], name, sizeof name, &len));
  uint32_t off, size;
  napi_value obj, v;
  if (b3w_input_signal(circuit, name, &off, &size) != B3W_OK) { NAPI_OK(napi_get_null(env, &obj)); return obj; }
  NAPI_OK(napi_create_object(env, &obj));
  NAPI_OK(napi_create_uint32(env, off, &v));  NAPI_OK(napi_set_named_property(env, obj, "offset", v));
  NAPI_OK(napi_create_uint32(env, size, &v)); NAPI_OK(napi_set_named_property(env, obj, "size", v));
  return obj;
}

// wtnsHeader(circuit) -> Uint8Array(76)
static napi_value WtnsHeader(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value argv[1];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  uint32_t circuit;
  NAPI_OK(napi_get_value_uint32(env, argv[0], &circuit));
  void *p; napi_value ab, v;
  NAPI_OK(napi_create_arraybuffer(env, 76, &p, &ab));
  int rc = b3w_wtns_header(circuit, (uint8_t *)p);
  if (rc) return throw_b3w(env, rc);
  NAPI_OK(napi_create_typedarray(env, napi_uint8_array, 76, ab, 0, &v));
  return v;
}

// witnessBatch(ctx, rows: Uint32Array, n, wantWitness[, wantSums]) -> Promise<{witness: Uint8Array|null, status: Uint8Array,
//   pub: Uint32Array[, sums: BigUint64Array]}>; sums = the per-instance 64-bit witness checksums of b3w_batch_extras (computed by
//   the kernel from the values it stores): what lets a caller that streams (wantWitness = false) account for every witness
// witnessOne(ctx, row) -> Promise<Uint8Array>
// witnessBatchFr(ctx, fr: Uint8Array(n * nInputs * 32), n, wantWitness) / witnessOneFr(ctx, fr): the same with inputs as
// little-endian field elements (b3w_witness_batch_fr: blake3_compression takes every value the reference takes)
struct job {
  napi_async_work work; napi_deferred deferred;
  b3w_ctx *ctx; uint32_t circuit; uint32_t *rows; uint8_t *fr; uint64_t n; b3w_info bi;
  uint8_t *out, *status; uint32_t *pub; uint64_t *sums; bool one, out_pinned; int rc; char err[512];
};
static void job_run(napi_env, void *data) {
  job *j = (job *)data;
  b3w_batch_extras ex;
  memset(&ex, 0, sizeof ex);
  ex.sums = j->sums;
  j->rc = j->fr ? b3w_witness_batch_fr_ex(j->ctx, j->fr, j->n, j->out, j->status, j->pub, &ex)
                : b3w_witness_batch_ex(j->ctx, j->rows, j->n, j->out, j->status, j->pub, &ex);
  if (j->rc == B3W_OK && j->one && j->status[0]) j->rc = j->status[0];
  if (j->rc == B3W_CIRCOM_ASSERT) {
    // witness_calculator.js:21-43,159-162: Error("Assert Failed.\n" + the printErrorMessage lines), re-wrapped
    char trace[400];
    if (j->fr) b3w_assert_trace_fr(j->circuit, j->fr, trace, sizeof trace);
    else b3w_assert_trace(j->circuit, j->rows, trace, sizeof trace);
    snprintf(j->err, sizeof j->err, "Error: Assert Failed.\n%s", trace);
  } else if (j->rc) {
    strncpy(j->err, b3w_last_error(), sizeof j->err - 1);
  }
}
static void job_free_buf(napi_env, void *data, void *) { free(data); }
// Witness buffers of large batches are PINNED host memory (b3w_host_alloc): the device-to-host copy of a pageable buffer
// runs at a fraction of the PCIe rate and cannot overlap the next chunk's kernel.  Small ones (a single witness is
// 771 KB) stay on malloc, which is cheaper than pinning.
#define PINNED_FROM ((size_t)32 << 20)
static uint8_t *out_alloc(size_t bytes, bool *pinned) {
  *pinned = bytes >= PINNED_FROM;
  if (*pinned) {
    uint8_t *p = (uint8_t *)b3w_host_alloc(bytes);
    if (p) return p;
    *pinned = false;                                   // pinning can fail (locked-memory limit): fall back to pageable
  }
  return (uint8_t *)malloc(bytes ? bytes : 1);
}
static void job_free_pinned(napi_env, void *data, void *) { b3w_host_free(data); }
static void job_done(napi_env env, napi_status, void *data) {
  job *j = (job *)data;
  napi_value res, v, ab;
  if (j->rc) {
    napi_value msg; napi_create_string_utf8(env, j->err, NAPI_AUTO_LENGTH, &msg);
    napi_create_error(env, NULL, msg, &res);
    napi_reject_deferred(env, j->deferred, res);
    if (j->out_pinned) b3w_host_free(j->out); else free(j->out);
    free(j->status); free(j->pub); free(j->sums);
  } else if (j->one) {
    size_t wb = (size_t)j->bi.witness_size * 32;
    napi_create_external_arraybuffer(env, j->out, wb, j->out_pinned ? job_free_pinned : job_free_buf, NULL, &ab);
    napi_create_typedarray(env, napi_uint8_array, wb, ab, 0, &res);
    napi_resolve_deferred(env, j->deferred, res);
    free(j->status); free(j->pub);
  } else {
    napi_create_object(env, &res);
    if (j->out) {
      size_t wb = (size_t)j->n * j->bi.witness_size * 32;
      napi_create_external_arraybuffer(env, j->out, wb, j->out_pinned ? job_free_pinned : job_free_buf, NULL, &ab);
      napi_create_typedarray(env, napi_uint8_array, wb, ab, 0, &v);
    } else napi_get_null(env, &v);
    napi_set_named_property(env, res, "witness", v);
    napi_create_external_arraybuffer(env, j->status, j->n, job_free_buf, NULL, &ab);
    napi_create_typedarray(env, napi_uint8_array, j->n, ab, 0, &v);
    napi_set_named_property(env, res, "status", v);
    napi_create_external_arraybuffer(env, j->pub, j->n * j->bi.n_public * 4, job_free_buf, NULL, &ab);
    napi_create_typedarray(env, napi_uint32_array, j->n * j->bi.n_public, ab, 0, &v);
    napi_set_named_property(env, res, "pub", v);
    if (j->sums) {
      napi_create_external_arraybuffer(env, j->sums, j->n * 8, job_free_buf, NULL, &ab);
      napi_create_typedarray(env, napi_biguint64_array, j->n, ab, 0, &v);
      napi_set_named_property(env, res, "sums", v);
    }
    napi_resolve_deferred(env, j->deferred, res);
  }
  napi_delete_async_work(env, j->work);
  free(j->rows);
  free(j->fr);
  free(j);
}
static napi_value start_job(napi_env env, napi_callback_info info, bool one, bool fr = false) {
  size_t argc = 5; napi_value argv[5];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  job *j = (job *)calloc(1, sizeof(job));
  handle *h = NULL;
  NAPI_OK(napi_get_value_external(env, argv[0], (void **)&h));
  j->ctx = h->ctx;
  j->circuit = h->circuit;
  b3w_circuit_info(h->circuit, &j->bi);
  napi_typedarray_type tt; size_t len; void *data; napi_value ab; size_t off;
  NAPI_OK(napi_get_typedarray_info(env, argv[1], &tt, &len, &data, &ab, &off));
  bool want = true, want_sums = false;
  if (one) j->n = 1;
  else {
    uint32_t n32; NAPI_OK(napi_get_value_uint32(env, argv[2], &n32)); j->n = n32;
    NAPI_OK(napi_get_value_bool(env, argv[3], &want));
    if (argc >= 5) NAPI_OK(napi_get_value_bool(env, argv[4], &want_sums));
  }
  if (len != (size_t)j->n * j->bi.n_inputs * (fr ? 32 : 1) || tt != (fr ? napi_uint8_array : napi_uint32_array)) {
    free(j);
    napi_throw_error(env, NULL, fr ? "fr must be a Uint8Array of n * nInputs * 32 bytes" : "rows must be a Uint32Array of n * nInputs values");
    return NULL;
  }
  j->one = one;
  if (fr) {
    j->fr = (uint8_t *)malloc(len ? len : 1);
    memcpy(j->fr, data, len);
  } else {
    j->rows = (uint32_t *)malloc(len ? len * 4 : 4);
    memcpy(j->rows, data, len * 4);
  }
  j->status = (uint8_t *)calloc(j->n ? j->n : 1, 1);
  j->pub = (uint32_t *)calloc((j->n ? j->n : 1) * 16, 4);
  j->sums = want_sums ? (uint64_t *)calloc(j->n ? j->n : 1, 8) : NULL;
  j->out = want ? out_alloc((size_t)(j->n ? j->n : 1) * j->bi.witness_size * 32, &j->out_pinned) : NULL;
  if (want && !j->out) {
    free(j->rows); free(j->fr); free(j->status); free(j->pub); free(j->sums); free(j);
    napi_throw_error(env, NULL, "out of host memory for the witness buffer");
    return NULL;
  }
  napi_value promise, name;
  NAPI_OK(napi_create_promise(env, &j->deferred, &promise));
  NAPI_OK(napi_create_string_utf8(env, "b3w_witness_batch", NAPI_AUTO_LENGTH, &name));
  NAPI_OK(napi_create_async_work(env, NULL, name, job_run, job_done, j, &j->work));
  NAPI_OK(napi_queue_async_work(env, j->work));
  return promise;
}
static napi_value WitnessBatch(napi_env env, napi_callback_info info) { return start_job(env, info, false); }
static napi_value WitnessOne(napi_env env, napi_callback_info info) { return start_job(env, info, true); }
static napi_value WitnessBatchFr(napi_env env, napi_callback_info info) { return start_job(env, info, false, true); }
static napi_value WitnessOneFr(napi_env env, napi_callback_info info) { return start_job(env, info, true, true); }

static napi_value Init(napi_env env, napi_value exports) {
  napi_property_descriptor d[] = {
      {"create", 0, Create, 0, 0, 0, napi_default, 0},           {"circuitInfo", 0, CircuitInfo, 0, 0, 0, napi_default, 0},
      {"inputSignal", 0, InputSignal, 0, 0, 0, napi_default, 0}, {"wtnsHeader", 0, WtnsHeader, 0, 0, 0, napi_default, 0},
      {"witnessBatch", 0, WitnessBatch, 0, 0, 0, napi_default, 0}, {"witnessOne", 0, WitnessOne, 0, 0, 0, napi_default, 0},
      {"witnessBatchFr", 0, WitnessBatchFr, 0, 0, 0, napi_default, 0}, {"witnessOneFr", 0, WitnessOneFr, 0, 0, 0, napi_default, 0}};
  napi_define_properties(env, exports, sizeof d / sizeof d[0], d);
  return exports;
}
NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
