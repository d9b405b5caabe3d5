/* node_api.h -- TEST STUB (tests/napi_mock): the subset of Node's stable N-API that integration/js/addon/blake3wit_napi.cc
 * uses, declared with the signatures of Node's own node_api.h / js_native_api.h (N-API version 8).  Node is not part of
 * this repository's build image; this header lets the addon be COMPILED, and napi_mock.cc lets it be EXECUTED against a
 * minimal in-process value model, so that the layer between witness_calculator.js and libblake3wit.so is tested code.
 * It is test infrastructure only: a real build uses the headers that ship with Node. */
#ifndef B3W_TEST_NODE_API_H
#define B3W_TEST_NODE_API_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct napi_env__ *napi_env;
typedef struct napi_value__ *napi_value;
typedef struct napi_deferred__ *napi_deferred;
typedef struct napi_async_work__ *napi_async_work;
typedef struct napi_callback_info__ *napi_callback_info;

typedef enum {
  napi_ok, napi_invalid_arg, napi_object_expected, napi_string_expected, napi_name_expected, napi_function_expected,
  napi_number_expected, napi_boolean_expected, napi_array_expected, napi_generic_failure, napi_pending_exception,
  napi_cancelled, napi_escape_called_twice, napi_handle_scope_mismatch, napi_callback_scope_mismatch,
  napi_queue_full, napi_closing, napi_bigint_expected, napi_date_expected, napi_arraybuffer_expected,
  napi_detachable_arraybuffer_expected, napi_would_deadlock
} napi_status;

typedef enum {
  napi_default = 0, napi_writable = 1 << 0, napi_enumerable = 1 << 1, napi_configurable = 1 << 2, napi_static = 1 << 10
} napi_property_attributes;

typedef enum {
  napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array, napi_int32_array,
  napi_uint32_array, napi_float32_array, napi_float64_array, napi_bigint64_array, napi_biguint64_array
} napi_typedarray_type;

typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void *finalize_data, void *finalize_hint);
typedef void (*napi_async_execute_callback)(napi_env env, void *data);
typedef void (*napi_async_complete_callback)(napi_env env, napi_status status, void *data);

typedef struct {
  const char *utf8name;
  napi_value name;
  napi_callback method;
  napi_callback getter;
  napi_callback setter;
  napi_value value;
  napi_property_attributes attributes;
  void *data;
} napi_property_descriptor;

#define NAPI_AUTO_LENGTH SIZE_MAX

napi_status napi_throw_error(napi_env env, const char *code, const char *msg);
napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t *argc, napi_value *argv, napi_value *this_arg, void **data);
napi_status napi_get_value_uint32(napi_env env, napi_value value, uint32_t *result);
napi_status napi_get_value_int32(napi_env env, napi_value value, int32_t *result);
napi_status napi_get_value_bool(napi_env env, napi_value value, bool *result);
napi_status napi_get_value_string_utf8(napi_env env, napi_value value, char *buf, size_t bufsize, size_t *result);
napi_status napi_get_value_external(napi_env env, napi_value value, void **result);
napi_status napi_create_external(napi_env env, void *data, napi_finalize finalize_cb, void *finalize_hint, napi_value *result);
napi_status napi_create_object(napi_env env, napi_value *result);
napi_status napi_create_uint32(napi_env env, uint32_t value, napi_value *result);
napi_status napi_set_named_property(napi_env env, napi_value object, const char *utf8name, napi_value value);
napi_status napi_create_array_with_length(napi_env env, size_t length, napi_value *result);
napi_status napi_set_element(napi_env env, napi_value object, uint32_t index, napi_value value);
napi_status napi_create_arraybuffer(napi_env env, size_t byte_length, void **data, napi_value *result);
napi_status napi_create_external_arraybuffer(napi_env env, void *external_data, size_t byte_length, napi_finalize finalize_cb,
                                             void *finalize_hint, napi_value *result);
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type type, size_t length, napi_value arraybuffer,
                                   size_t byte_offset, napi_value *result);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type *type, size_t *length,
                                     void **data, napi_value *arraybuffer, size_t *byte_offset);
napi_status napi_get_null(napi_env env, napi_value *result);
napi_status napi_create_string_utf8(napi_env env, const char *str, size_t length, napi_value *result);
napi_status napi_create_error(napi_env env, napi_value code, napi_value msg, napi_value *result);
napi_status napi_create_promise(napi_env env, napi_deferred *deferred, napi_value *promise);
napi_status napi_resolve_deferred(napi_env env, napi_deferred deferred, napi_value resolution);
napi_status napi_reject_deferred(napi_env env, napi_deferred deferred, napi_value rejection);
napi_status napi_create_async_work(napi_env env, napi_value async_resource, napi_value async_resource_name,
                                   napi_async_execute_callback execute, napi_async_complete_callback complete, void *data,
                                   napi_async_work *result);
napi_status napi_queue_async_work(napi_env env, napi_async_work work);
napi_status napi_delete_async_work(napi_env env, napi_async_work work);
napi_status napi_define_properties(napi_env env, napi_value object, size_t property_count,
                                   const napi_property_descriptor *properties);

typedef napi_value (*napi_addon_register_func)(napi_env env, napi_value exports);
typedef struct napi_module {
  int nm_version;
  unsigned int nm_flags;
  const char *nm_filename;
  napi_addon_register_func nm_register_func;
  const char *nm_modname;
  void *nm_priv;
  void *reserved[4];
} napi_module;
void napi_module_register(napi_module *mod);

#ifdef __cplusplus
}
#endif

#define NAPI_MODULE_STR_(x) #x
#define NAPI_MODULE_STR(x) NAPI_MODULE_STR_(x)
#define NAPI_MODULE_X(modname, regfunc, priv, flags)                                                     \
  static napi_module _module = {1, flags, __FILE__, regfunc, NAPI_MODULE_STR(modname), priv, {0}};      \
  static void _register_module(void) __attribute__((constructor));                                       \
  static void _register_module(void) { napi_module_register(&_module); }
#define NAPI_MODULE(modname, regfunc) NAPI_MODULE_X(modname, regfunc, NULL, 0)

#endif
