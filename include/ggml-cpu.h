/* ggml-cpu.h -- shim: the reference's include/whisper.h includes it (line 5) but uses nothing from it. */
#ifndef GGML_CPU_H
#define GGML_CPU_H
#include "ggml.h"
#endif
