/*
 * tests/node_api_stub/node_api.h: TEST ONLY.  Declarations of the handful of
 * Node-API (N-API, ABI-stable) calls integration/addon.cc uses, written from
 * the public N-API documentation, so that the addon can be COMPILED in an
 * image without node (tests/test_integration_addon.py).  Nothing here is
 * linked or run; a real build uses node's own <node_api.h>.
 */
#ifndef DNG_TEST_NODE_API_STUB_H
#define DNG_TEST_NODE_API_STUB_H

#include <stddef.h>
#include <stdint.h>

extern "C" {

typedef struct napi_env__ *napi_env;
typedef struct napi_value__ *napi_value;
typedef struct napi_ref__ *napi_ref;
typedef struct napi_callback_info__ *napi_callback_info;
typedef struct napi_async_work__ *napi_async_work;
typedef enum { napi_ok = 0, napi_invalid_arg, napi_generic_failure = 9 } napi_status;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void *data, void *hint);
typedef void (*napi_async_execute_callback)(napi_env env, void *data);
typedef void (*napi_async_complete_callback)(napi_env env, napi_status status,
    void *data);

#define NAPI_AUTO_LENGTH ((size_t)-1)

napi_status napi_get_undefined(napi_env env, napi_value *result);
napi_status napi_create_object(napi_env env, napi_value *result);
napi_status napi_create_array_with_length(napi_env env, size_t length,
    napi_value *result);
napi_status napi_create_double(napi_env env, double value, napi_value *result);
napi_status napi_create_string_utf8(napi_env env, const char *str,
    size_t length, napi_value *result);
napi_status napi_create_error(napi_env env, napi_value code, napi_value msg,
    napi_value *result);
napi_status napi_create_function(napi_env env, const char *utf8name,
    size_t length, napi_callback cb, void *data, napi_value *result);
napi_status napi_set_named_property(napi_env env, napi_value object,
    const char *utf8name, napi_value value);
napi_status napi_set_element(napi_env env, napi_value object, uint32_t index,
    napi_value value);
napi_status napi_get_element(napi_env env, napi_value object, uint32_t index,
    napi_value *result);
napi_status napi_get_array_length(napi_env env, napi_value value,
    uint32_t *result);
napi_status napi_get_value_int32(napi_env env, napi_value value,
    int32_t *result);
napi_status napi_get_value_string_utf8(napi_env env, napi_value value,
    char *buf, size_t bufsize, size_t *result);
napi_status napi_get_buffer_info(napi_env env, napi_value value, void **data,
    size_t *length);
napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo,
    size_t *argc, napi_value *argv, napi_value *this_arg, void **data);
napi_status napi_wrap(napi_env env, napi_value js_object, void *native_object,
    napi_finalize finalize_cb, void *finalize_hint, napi_ref *result);
napi_status napi_unwrap(napi_env env, napi_value js_object, void **result);
napi_status napi_create_reference(napi_env env, napi_value value,
    uint32_t initial_refcount, napi_ref *result);
napi_status napi_delete_reference(napi_env env, napi_ref ref);
napi_status napi_get_reference_value(napi_env env, napi_ref ref,
    napi_value *result);
napi_status napi_call_function(napi_env env, napi_value recv, napi_value func,
    size_t argc, const napi_value *argv, napi_value *result);
napi_status napi_throw_error(napi_env env, const char *code, const char *msg);
napi_status napi_create_async_work(napi_env env, napi_value async_resource,
    napi_value async_resource_name, napi_async_execute_callback execute,
    napi_async_complete_callback complete, void *data,
    napi_async_work *result);
napi_status napi_queue_async_work(napi_env env, napi_async_work work);
napi_status napi_delete_async_work(napi_env env, napi_async_work work);

#define NAPI_MODULE_INIT() \
	extern "C" napi_value napi_register_module_v1(napi_env env, \
	    napi_value exports)

} /* extern "C" */

#endif
