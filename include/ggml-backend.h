/* ggml-backend.h -- shim: examples/cli/cli.cpp and examples/bench/bench.cpp of the reference include it only for
 * ggml_backend_load_all(), declared in the ggml.h shim. */
#ifndef GGML_BACKEND_H
#define GGML_BACKEND_H
#include "ggml.h"
#endif
