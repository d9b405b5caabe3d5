/* wasm_harness.c -- host side of the circom wasm ABI, restated in C.  TEST INFRASTRUCTURE ONLY.
 *
 * Linked with one C file emitted by wasm2c.py from a reference .wasm, this is "Oracle A": the
 * reference's own witness program driven with the protocol of
 *   /root/reference/blake3_nova_js/witness_calculator.js
 *     builder()               :1-106   -> ref_new()            (imports at :20-63)
 *     WitnessCalculator ctor  :108-125 -> ref_version/ref_n32/ref_prime/ref_witness_size
 *     _doCalculateWitness     :131-169 -> ref_calculate()      (init, per value: 8x writeSharedRWMemory
 *                                                               + setInputSignal, completeness check)
 *     calculateBinWitness     :190-205 -> read-out loop in ref_calculate()
 *     fnvHash                 :325-337 -> ref_fnv1a64()
 * It also is the CPU baseline runner (ref_batch(): one instance per thread).
 */
#include "wasm_rt.h"
#include <stdio.h>
#include <pthread.h>
#include <time.h>

void wasm_instantiate(W *w);
uint32_t wx_getVersion(W *w);
uint32_t wx_getMinorVersion(W *w);
uint32_t wx_getPatchVersion(W *w);
uint32_t wx_readSharedRWMemory(W *w, uint32_t j);
void wx_writeSharedRWMemory(W *w, uint32_t j, uint32_t v);
void wx_init(W *w, uint32_t sanity);
void wx_setInputSignal(W *w, uint32_t hmsb, uint32_t hlsb, uint32_t pos);
uint32_t wx_getInputSignalSize(W *w, uint32_t hmsb, uint32_t hlsb);
void wx_getRawPrime(W *w);
uint32_t wx_getFieldNumLen32(W *w);
uint32_t wx_getWitnessSize(W *w);
uint32_t wx_getInputSize(W *w);
void wx_getWitness(W *w, uint32_t i);
uint32_t wx_getMessageChar(W *w);

/* ---- runtime imports (witness_calculator.js:20-63) ---- */
void wasm_trap(W *w, int code) {
  w->err_code = 100 + code;
  if (w->jb) longjmp(*w->jb, 1);
  abort();
}
static void drain_message(W *w, char *dst, size_t cap) {
  size_t n = strlen(dst);
  uint32_t c;
  while ((c = wx_getMessageChar(w)) != 0)
    if (n + 2 < cap) dst[n++] = (char)c;
  dst[n] = 0;
}
void wi_runtime_exceptionHandler(W *w, uint32_t code) {
  w->err_code = (int)code;
  if (w->jb) longjmp(*w->jb, 1);
  abort();
}
void wi_runtime_printErrorMessage(W *w) {
  drain_message(w, w->err_msg, sizeof w->err_msg - 1);
  size_t n = strlen(w->err_msg);
  w->err_msg[n] = '\n';
  w->err_msg[n + 1] = 0;
}
void wi_runtime_writeBufferMessage(W *w) {
  /* js: "\n" flushes, otherwise space-joined */
  char tmp[512] = {0};
  drain_message(w, tmp, sizeof tmp);
  size_t n = strlen(w->log_msg);
  if (strcmp(tmp, "\n") == 0) {
    if (n + 2 < sizeof w->log_msg) { w->log_msg[n] = '\n'; w->log_msg[n + 1] = 0; }
  } else {
    if (n && w->log_msg[n - 1] != '\n' && n + 2 < sizeof w->log_msg) { w->log_msg[n++] = ' '; w->log_msg[n] = 0; }
    strncat(w->log_msg, tmp, sizeof w->log_msg - n - 1);
  }
}
void wi_runtime_showSharedRWMemory(W *w) {
  /* only small values are ever logged by the reference circuits; print the low 64 bits */
  uint64_t v = (uint64_t)wx_readSharedRWMemory(w, 0) | ((uint64_t)wx_readSharedRWMemory(w, 1) << 32);
  size_t n = strlen(w->log_msg);
  if (n && w->log_msg[n - 1] != '\n' && n + 2 < sizeof w->log_msg) { w->log_msg[n++] = ' '; w->log_msg[n] = 0; }
  snprintf(w->log_msg + n, sizeof w->log_msg - n, "%llu", (unsigned long long)v);
}

/* ---- public harness API (ctypes-friendly) ---- */
W *ref_new(void) {
  W *w = (W *)calloc(1, sizeof(W));
  wasm_instantiate(w);
  return w;
}
void ref_free(W *w) {
  if (!w) return;
  free(w->mem);
  free(w);
}
uint32_t ref_version(W *w) { return wx_getVersion(w); }
uint32_t ref_minor_version(W *w) { return wx_getMinorVersion(w); }
uint32_t ref_patch_version(W *w) { return wx_getPatchVersion(w); }
uint32_t ref_n32(W *w) { return wx_getFieldNumLen32(w); }
uint32_t ref_witness_size(W *w) { return wx_getWitnessSize(w); }
uint32_t ref_input_size(W *w) { return wx_getInputSize(w); }
void ref_prime(W *w, uint32_t *limbs8) {
  wx_getRawPrime(w);
  for (uint32_t j = 0; j < 8; j++) limbs8[j] = wx_readSharedRWMemory(w, j);
}
int32_t ref_input_signal_size(W *w, uint32_t hmsb, uint32_t hlsb) { return (int32_t)wx_getInputSignalSize(w, hmsb, hlsb); }
const char *ref_err_msg(W *w) { return w->err_msg; }
const char *ref_log_msg(W *w) { return w->log_msg; }
const uint8_t *ref_memory(W *w) { return w->mem; }

uint64_t ref_fnv1a64(const char *s) {
  uint64_t h = 0xCBF29CE484222325ull;
  for (; *s; s++) { h ^= (uint8_t)*s; h *= 0x100000001B3ull; }
  return h;
}

/* One full witness calculation.
 *   n_vals values; value k belongs to the signal with hash (hmsb[k],hlsb[k]) at position pos[k];
 *   vals[k*8 + j] = limb j (bits 32j..32j+31) of the already-normalised value.
 *   out = witnessSize*32 bytes (canonical little-endian), may be NULL (timing only -> still read out to a scratch).
 * Returns 0, or the exceptionHandler code (1..6), or 100+trap, or -1 ("Not all inputs have been set"). */
int ref_calculate(W *w, uint32_t n_vals, const uint32_t *hmsb, const uint32_t *hlsb, const uint32_t *pos,
                  const uint32_t *vals, uint8_t *out) {
  jmp_buf jb;
  w->err_code = 0;
  w->err_msg[0] = 0;
  w->log_msg[0] = 0;
  w->jb = &jb;
  if (setjmp(jb)) { w->jb = NULL; return w->err_code; }
  wx_init(w, 0);
  uint32_t counter = 0;
  for (uint32_t k = 0; k < n_vals; k++) {
    for (uint32_t j = 0; j < 8; j++) wx_writeSharedRWMemory(w, j, vals[k * 8 + j]);
    wx_setInputSignal(w, hmsb[k], hlsb[k], pos[k]);
    counter++;
  }
  if (counter < wx_getInputSize(w)) { w->jb = NULL; return -1; }
  uint32_t ws = wx_getWitnessSize(w);
  uint32_t scratch[8];
  for (uint32_t i = 0; i < ws; i++) {
    wx_getWitness(w, i);
    uint32_t *dst = out ? (uint32_t *)(out + (size_t)i * 32) : scratch;
    for (uint32_t j = 0; j < 8; j++) dst[j] = wx_readSharedRWMemory(w, j);
  }
  w->jb = NULL;
  return 0;
}

/* Batch over u32 inputs (the honest domain): every instance sets the same n_vals (signal,pos) plan with
 * u32 values in[i*n_vals + k]; instance i's witness goes to out + i*witnessSize*32 (or nowhere if out==NULL).
 * status[i] = return code.  Work is split over nthreads pthreads, one wasm instance each.
 * Returns wall-clock seconds spent (all threads, start to finish). */
typedef struct {
  uint32_t n_vals; const uint32_t *hmsb, *hlsb, *pos, *in; uint8_t *out; int32_t *status;
  uint64_t first, count; uint32_t ws;
} job_t;
static void *batch_worker(void *arg) {
  job_t *j = (job_t *)arg;
  W *w = ref_new();
  uint32_t *vals = (uint32_t *)calloc((size_t)j->n_vals * 8, 4);
  for (uint64_t i = j->first; i < j->first + j->count; i++) {
    for (uint32_t k = 0; k < j->n_vals; k++) vals[k * 8] = j->in[i * j->n_vals + k];
    int rc = ref_calculate(w, j->n_vals, j->hmsb, j->hlsb, j->pos, vals,
                           j->out ? j->out + (size_t)i * j->ws * 32 : NULL);
    if (j->status) j->status[i] = rc;
  }
  free(vals);
  ref_free(w);
  return NULL;
}
double ref_batch(uint32_t n_vals, const uint32_t *hmsb, const uint32_t *hlsb, const uint32_t *pos,
                 const uint32_t *in, uint64_t n, uint8_t *out, int32_t *status, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if ((uint64_t)nthreads > n) nthreads = (int)(n ? n : 1);
  W *probe = ref_new();
  uint32_t ws = wx_getWitnessSize(probe);
  ref_free(probe);
  pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
  job_t *jobs = (job_t *)calloc((size_t)nthreads, sizeof(job_t));
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  uint64_t base = 0;
  for (int t = 0; t < nthreads; t++) {
    uint64_t cnt = n / (uint64_t)nthreads + ((uint64_t)t < n % (uint64_t)nthreads ? 1 : 0);
    jobs[t] = (job_t){n_vals, hmsb, hlsb, pos, in, out, status, base, cnt, ws};
    base += cnt;
    pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(th);
  free(jobs);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
