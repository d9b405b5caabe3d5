// napi_mock.cc -- TEST INFRASTRUCTURE: a minimal in-process implementation of the N-API subset declared in the stub
// node_api.h next to this file, plus a small C driver API (mk_*) that tests/test_napi_addon.py calls through ctypes.
// It lets integration/js/addon/blake3wit_napi.cc -- the layer between witness_calculator.js and libblake3wit.so -- be
// compiled and EXECUTED in an image without Node: values are plain heap objects, async work runs synchronously inside
// napi_queue_async_work (execute, then complete), promises record their settlement.
#include "node_api.h"

#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

namespace {
enum kind { K_NULL, K_U32, K_I32, K_BOOL, K_STRING, K_OBJECT, K_ARRAY, K_ARRAYBUFFER, K_TYPEDARRAY, K_EXTERNAL, K_ERROR,
            K_PROMISE, K_FUNCTION };

struct value {
  kind k = K_NULL;
  uint32_t u = 0;
  int32_t i = 0;
  bool b = false;
  std::string s;                                 // string / error message
  std::map<std::string, value *> props;          // object
  std::vector<value *> elems;                    // array
  void *data = nullptr;                          // arraybuffer bytes / external pointer
  size_t len = 0;                                // arraybuffer byte length / typedarray element count
  bool owned = false;                            // arraybuffer allocated by the mock
  napi_finalize fin = nullptr;
  void *fin_hint = nullptr;
  napi_typedarray_type tt = napi_uint8_array;    // typedarray
  value *buffer = nullptr;
  size_t byte_offset = 0;
  int state = 0;                                 // promise: 0 pending, 1 resolved, 2 rejected
  value *settled = nullptr;
  napi_callback fn = nullptr;                    // function
};

struct cbinfo { std::vector<value *> args; };
struct work { napi_async_execute_callback exec; napi_async_complete_callback done; void *data; };

std::vector<value *> g_values;
std::string g_exception;
bool g_has_exception = false;
napi_module *g_module = nullptr;
value *g_exports = nullptr;
napi_env const ENV = (napi_env)0x1;

value *mk(kind k) {
  value *v = new value();
  v->k = k;
  g_values.push_back(v);
  return v;
}
napi_value out(value *v) { return (napi_value)v; }
value *in(napi_value v) { return (value *)v; }
size_t elem_size(napi_typedarray_type t) {
  switch (t) {
    case napi_int8_array: case napi_uint8_array: case napi_uint8_clamped_array: return 1;
    case napi_int16_array: case napi_uint16_array: return 2;
    case napi_int32_array: case napi_uint32_array: case napi_float32_array: return 4;
    default: return 8;
  }
}
}  // namespace

extern "C" {
// ---- the N-API subset ------------------------------------------------------------------------------------------
void napi_module_register(napi_module *mod) { g_module = mod; }

napi_status napi_throw_error(napi_env, const char *, const char *msg) {
  g_exception = msg ? msg : "";
  g_has_exception = true;
  return napi_ok;
}
napi_status napi_get_cb_info(napi_env, napi_callback_info info, size_t *argc, napi_value *argv, napi_value *this_arg, void **data) {
  cbinfo *ci = (cbinfo *)info;
  const size_t cap = argc ? *argc : 0;
  for (size_t i = 0; i < cap; i++) argv[i] = i < ci->args.size() ? out(ci->args[i]) : nullptr;   // missing = undefined
  if (argc) *argc = ci->args.size();
  if (this_arg) *this_arg = nullptr;
  if (data) *data = nullptr;
  return napi_ok;
}
napi_status napi_get_value_uint32(napi_env, napi_value v, uint32_t *r) {
  if (!v || (in(v)->k != K_U32 && in(v)->k != K_I32)) return napi_number_expected;
  *r = in(v)->k == K_U32 ? in(v)->u : (uint32_t)in(v)->i;
  return napi_ok;
}
napi_status napi_get_value_int32(napi_env, napi_value v, int32_t *r) {
  if (!v || (in(v)->k != K_U32 && in(v)->k != K_I32)) return napi_number_expected;
  *r = in(v)->k == K_I32 ? in(v)->i : (int32_t)in(v)->u;
  return napi_ok;
}
napi_status napi_get_value_bool(napi_env, napi_value v, bool *r) {
  if (!v || in(v)->k != K_BOOL) return napi_boolean_expected;
  *r = in(v)->b;
  return napi_ok;
}
napi_status napi_get_value_string_utf8(napi_env, napi_value v, char *buf, size_t bufsize, size_t *result) {
  if (!v || in(v)->k != K_STRING) return napi_string_expected;
  const std::string &s = in(v)->s;
  if (!buf) { if (result) *result = s.size(); return napi_ok; }
  if (bufsize == 0) return napi_invalid_arg;
  const size_t n = s.size() < bufsize - 1 ? s.size() : bufsize - 1;
  memcpy(buf, s.data(), n);
  buf[n] = 0;
  if (result) *result = n;
  return napi_ok;
}
napi_status napi_get_value_external(napi_env, napi_value v, void **r) {
  if (!v || in(v)->k != K_EXTERNAL) return napi_invalid_arg;
  *r = in(v)->data;
  return napi_ok;
}
napi_status napi_create_external(napi_env, void *data, napi_finalize fin, void *hint, napi_value *r) {
  value *v = mk(K_EXTERNAL);
  v->data = data; v->fin = fin; v->fin_hint = hint;
  *r = out(v);
  return napi_ok;
}
napi_status napi_create_object(napi_env, napi_value *r) { *r = out(mk(K_OBJECT)); return napi_ok; }
napi_status napi_create_uint32(napi_env, uint32_t x, napi_value *r) { value *v = mk(K_U32); v->u = x; *r = out(v); return napi_ok; }
napi_status napi_set_named_property(napi_env, napi_value o, const char *name, napi_value val) {
  if (!o || in(o)->k != K_OBJECT) return napi_object_expected;
  in(o)->props[name] = in(val);
  return napi_ok;
}
napi_status napi_create_array_with_length(napi_env, size_t n, napi_value *r) {
  value *v = mk(K_ARRAY);
  v->elems.assign(n, nullptr);
  *r = out(v);
  return napi_ok;
}
napi_status napi_set_element(napi_env, napi_value o, uint32_t i, napi_value val) {
  if (!o || in(o)->k != K_ARRAY) return napi_array_expected;
  if (i >= in(o)->elems.size()) in(o)->elems.resize(i + 1, nullptr);
  in(o)->elems[i] = in(val);
  return napi_ok;
}
napi_status napi_create_arraybuffer(napi_env, size_t n, void **data, napi_value *r) {
  value *v = mk(K_ARRAYBUFFER);
  v->data = calloc(n ? n : 1, 1);
  v->len = n;
  v->owned = true;
  if (data) *data = v->data;
  *r = out(v);
  return napi_ok;
}
napi_status napi_create_external_arraybuffer(napi_env, void *p, size_t n, napi_finalize fin, void *hint, napi_value *r) {
  value *v = mk(K_ARRAYBUFFER);
  v->data = p; v->len = n; v->fin = fin; v->fin_hint = hint;
  *r = out(v);
  return napi_ok;
}
napi_status napi_create_typedarray(napi_env, napi_typedarray_type t, size_t length, napi_value ab, size_t off, napi_value *r) {
  if (!ab || in(ab)->k != K_ARRAYBUFFER) return napi_invalid_arg;
  if (off % elem_size(t) || off + length * elem_size(t) > in(ab)->len) return napi_invalid_arg;   // Node throws a RangeError
  value *v = mk(K_TYPEDARRAY);
  v->tt = t; v->len = length; v->buffer = in(ab); v->byte_offset = off;
  *r = out(v);
  return napi_ok;
}
napi_status napi_get_typedarray_info(napi_env, napi_value ta, napi_typedarray_type *t, size_t *length, void **data, napi_value *ab,
                                     size_t *off) {
  if (!ta || in(ta)->k != K_TYPEDARRAY) return napi_invalid_arg;
  value *v = in(ta);
  if (t) *t = v->tt;
  if (length) *length = v->len;
  if (data) *data = (uint8_t *)v->buffer->data + v->byte_offset;
  if (ab) *ab = out(v->buffer);
  if (off) *off = v->byte_offset;
  return napi_ok;
}
napi_status napi_get_null(napi_env, napi_value *r) { *r = out(mk(K_NULL)); return napi_ok; }
napi_status napi_create_string_utf8(napi_env, const char *s, size_t n, napi_value *r) {
  value *v = mk(K_STRING);
  v->s = n == NAPI_AUTO_LENGTH ? std::string(s) : std::string(s, n);
  *r = out(v);
  return napi_ok;
}
napi_status napi_create_error(napi_env, napi_value, napi_value msg, napi_value *r) {
  if (!msg || in(msg)->k != K_STRING) return napi_string_expected;
  value *v = mk(K_ERROR);
  v->s = in(msg)->s;
  *r = out(v);
  return napi_ok;
}
napi_status napi_create_promise(napi_env, napi_deferred *d, napi_value *p) {
  value *v = mk(K_PROMISE);
  *d = (napi_deferred)v;
  *p = out(v);
  return napi_ok;
}
napi_status napi_resolve_deferred(napi_env, napi_deferred d, napi_value val) {
  value *p = (value *)d;
  if (p->state) return napi_generic_failure;
  p->state = 1; p->settled = in(val);
  return napi_ok;
}
napi_status napi_reject_deferred(napi_env, napi_deferred d, napi_value val) {
  value *p = (value *)d;
  if (p->state) return napi_generic_failure;
  p->state = 2; p->settled = in(val);
  return napi_ok;
}
napi_status napi_create_async_work(napi_env, napi_value, napi_value, napi_async_execute_callback e, napi_async_complete_callback c,
                                   void *data, napi_async_work *r) {
  *r = (napi_async_work) new work{e, c, data};
  return napi_ok;
}
napi_status napi_queue_async_work(napi_env env, napi_async_work w) {
  work *x = (work *)w;
  x->exec(env, x->data);                           // Node runs this on a pool thread and `done` on the main loop
  x->done(env, napi_ok, x->data);
  return napi_ok;
}
napi_status napi_delete_async_work(napi_env, napi_async_work w) { delete (work *)w; return napi_ok; }
napi_status napi_define_properties(napi_env, napi_value o, size_t n, const napi_property_descriptor *d) {
  if (!o || in(o)->k != K_OBJECT) return napi_object_expected;
  for (size_t i = 0; i < n; i++) {
    value *f = mk(K_FUNCTION);
    f->fn = d[i].method;
    in(o)->props[d[i].utf8name] = f;
  }
  return napi_ok;
}

// ---- driver API for the Python test -----------------------------------------------------------------------------
void *mk_exports(void) {
  if (!g_exports && g_module) {
    g_exports = new value();
    g_exports->k = K_OBJECT;
    g_module->nm_register_func(ENV, out(g_exports));
  }
  return g_exports;
}
const char *mk_module_name(void) { return g_module ? g_module->nm_modname : ""; }
void *mk_u32(uint32_t x) { value *v = mk(K_U32); v->u = x; return v; }
void *mk_i32(int32_t x) { value *v = mk(K_I32); v->i = x; return v; }
void *mk_bool(int x) { value *v = mk(K_BOOL); v->b = x != 0; return v; }
void *mk_str(const char *s) { value *v = mk(K_STRING); v->s = s; return v; }
void *mk_typed(int type, const void *data, size_t count) {       // a typed array over a copy of `data`
  napi_value ab, ta;
  void *p;
  const size_t bytes = count * elem_size((napi_typedarray_type)type);
  napi_create_arraybuffer(ENV, bytes, &p, &ab);
  if (bytes) memcpy(p, data, bytes);
  napi_create_typedarray(ENV, (napi_typedarray_type)type, count, ab, 0, &ta);
  return ta;
}
// calls exports[name](args...); returns the result, or NULL when the callback threw (see mk_exception)
void *mk_call(const char *name, int argc, void **argv) {
  value *e = (value *)mk_exports();
  g_has_exception = false;
  g_exception.clear();
  if (!e || !e->props.count(name) || e->props[name]->k != K_FUNCTION) {
    g_exception = std::string("no such export: ") + name;
    g_has_exception = true;
    return nullptr;
  }
  cbinfo ci;
  for (int i = 0; i < argc; i++) ci.args.push_back((value *)argv[i]);
  napi_value r = e->props[name]->fn(ENV, (napi_callback_info)&ci);
  return g_has_exception ? nullptr : (void *)r;
}
const char *mk_exception(void) { return g_has_exception ? g_exception.c_str() : nullptr; }
int mk_kind(void *v) { return v ? ((value *)v)->k : -1; }
void *mk_get(void *o, const char *name) {
  value *v = (value *)o;
  return v && v->k == K_OBJECT && v->props.count(name) ? v->props[name] : nullptr;
}
void *mk_elem(void *a, uint32_t i) {
  value *v = (value *)a;
  return v && v->k == K_ARRAY && i < v->elems.size() ? v->elems[i] : nullptr;
}
uint32_t mk_as_u32(void *v) { return ((value *)v)->u; }
const char *mk_as_str(void *v) { return ((value *)v)->s.c_str(); }       // string or error message
int mk_typed_info(void *ta, int *type, size_t *count, void **data) {
  napi_typedarray_type t;
  if (napi_get_typedarray_info(ENV, (napi_value)ta, &t, count, data, nullptr, nullptr) != napi_ok) return -1;
  *type = (int)t;
  return 0;
}
int mk_promise(void *p, void **settled) {
  value *v = (value *)p;
  if (!v || v->k != K_PROMISE) return -1;
  if (settled) *settled = v->settled;
  return v->state;
}
// drop every value the mock created since the last call (runs the finalizers the addon registered, like a GC would)
void mk_release_all(void) {
  for (value *v : g_values) {
    if (v->fin) v->fin(ENV, v->data, v->fin_hint);
    else if (v->k == K_ARRAYBUFFER && v->owned) free(v->data);
    delete v;
  }
  g_values.clear();
}
}
