/* ggml.h -- shim.  The reference's include/whisper.h includes ggml.h but uses only two typedefs from it
 * (ggml_abort_callback, reference ggml/include/ggml.h:694, used at include/whisper.h:574; ggml_log_callback and
 * enum ggml_log_level, ggml/include/ggml.h:622-629 and 2651, used at include/whisper.h:745).  This library has no
 * ggml backend; the shim supplies exactly those declarations so callers written against the reference compile. */
#ifndef GGML_H
#define GGML_H

#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

enum ggml_log_level {
    GGML_LOG_LEVEL_NONE = 0,
    GGML_LOG_LEVEL_DEBUG = 1,
    GGML_LOG_LEVEL_INFO = 2,
    GGML_LOG_LEVEL_WARN = 3,
    GGML_LOG_LEVEL_ERROR = 4,
    GGML_LOG_LEVEL_CONT = 5,
};

typedef bool (*ggml_abort_callback)(void * data);
typedef void (*ggml_log_callback)(enum ggml_log_level level, const char * text, void * user_data);

/* The reference's CLI, bench and server call this before initialising a context
 * (examples/cli/cli.cpp:929, examples/bench/bench.cpp:168, ggml/include/ggml-backend.h:246).  No-op here. */
#if defined(__GNUC__)
__attribute__((visibility("default")))
#endif
void ggml_backend_load_all(void);

#ifdef __cplusplus
}
#endif

#endif /* GGML_H */
