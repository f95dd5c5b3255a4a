// XLA custom-call entry point (legacy "API_VERSION_STATUS_RETURNING" GPU ABI), so that the path can
// be registered with jax.ffi without the XLA FFI headers (jaxlib is not installable in the build
// image).  buffers = { theta [C][D] (in), logp [C] (out), grad [C][D] (out) };
// opaque = bl_xla_opaque (handle pointer of a dataset created through bl_dataset_create in the same
// process, and the batch size).  Runs asynchronously on XLA's stream; errors are reported through
// XlaCustomCallStatusSetFailure when the XLA runtime exports it.
#include <dlfcn.h>

#include <cstring>

#include "engine.cuh"
#include "handle.h"

extern "C" {

typedef void (*xla_set_failure_fn)(void* status, const char* message, size_t message_len);

BL_API void bl_xla_eval(void* stream, void** buffers, const char* opaque, size_t opaque_len, void* status) {
  const char* err = nullptr;
  if (!buffers || !opaque || opaque_len < sizeof(bl_xla_opaque)) {
    err = "biolith_b200: bad custom-call operands";
  } else {
    bl_xla_opaque o;
    memcpy(&o, opaque, sizeof(o));
    bl_dataset* ds = reinterpret_cast<bl_dataset*>((uintptr_t)o.dataset);
    if (!ds || o.n_chains < 1) {
      err = "biolith_b200: bad opaque descriptor";
    } else if (bl_eval(ds, buffers[0], o.n_chains, buffers[1], buffers[2], stream) != BL_OK) {
      err = bl_last_error();
    }
  }
  if (err && status) {
    // error path only: look the callback up every time (the runtime that exports it may be loaded after us)
    xla_set_failure_fn set_failure = (xla_set_failure_fn)dlsym(RTLD_DEFAULT, "XlaCustomCallStatusSetFailure");
    if (set_failure) set_failure(status, err, strlen(err));
  }
}

}  // extern "C"
