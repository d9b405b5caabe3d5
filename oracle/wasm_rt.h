/* wasm_rt.h -- minimal runtime for C emitted by oracle/wasm2c.py.  TEST INFRASTRUCTURE ONLY. */
#ifndef WASM_RT_H
#define WASM_RT_H
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <setjmp.h>

#define WASM_PAGE 65536u
#define WASM_MAX_PAGES 128u

typedef struct W {
  uint8_t *mem;
  uint32_t pages;
  /* harness state (wasm_harness.c) */
  jmp_buf *jb;
  int err_code;          /* code passed to runtime.exceptionHandler, or 100+trap */
  char err_msg[4096];    /* text accumulated by runtime.printErrorMessage */
  char log_msg[4096];    /* text accumulated by runtime.writeBufferMessage / showSharedRWMemory */
} W;

void wasm_trap(W *w, int code);   /* provided by the harness; does not return */

static inline void wasm_alloc(W *w, uint32_t pages) {
  w->mem = (uint8_t *)calloc((size_t)WASM_MAX_PAGES, WASM_PAGE);
  w->pages = pages;
}
static inline uint32_t wasm_grow(W *w, uint32_t n) {
  uint32_t old = w->pages;
  if ((uint64_t)old + n > WASM_MAX_PAGES) return 0xFFFFFFFFu;
  w->pages = old + n;
  return old;
}
#define WASM_LDST(T)                                                              \
  static inline T wasm_ld_##T(W *w, uint64_t a) {                                 \
    T v;                                                                          \
    if (a + sizeof(T) > (uint64_t)w->pages * WASM_PAGE) wasm_trap(w, 4);          \
    memcpy(&v, w->mem + a, sizeof(T));                                            \
    return v;                                                                     \
  }                                                                               \
  static inline void wasm_st_##T(W *w, uint64_t a, T v) {                         \
    if (a + sizeof(T) > (uint64_t)w->pages * WASM_PAGE) wasm_trap(w, 4);          \
    memcpy(w->mem + a, &v, sizeof(T));                                            \
  }
WASM_LDST(uint8_t) WASM_LDST(int8_t) WASM_LDST(uint16_t) WASM_LDST(int16_t)
WASM_LDST(uint32_t) WASM_LDST(int32_t) WASM_LDST(uint64_t)

static inline uint32_t wasm_clz32(uint32_t x) { return x ? (uint32_t)__builtin_clz(x) : 32u; }
static inline uint32_t wasm_ctz32(uint32_t x) { return x ? (uint32_t)__builtin_ctz(x) : 32u; }
static inline uint64_t wasm_clz64(uint64_t x) { return x ? (uint64_t)__builtin_clzll(x) : 64u; }
static inline uint64_t wasm_ctz64(uint64_t x) { return x ? (uint64_t)__builtin_ctzll(x) : 64u; }
#endif
