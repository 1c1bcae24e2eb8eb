// CUDA error latch (no exception may cross the C ABI; see common.cuh).
//
// The latch is per THREAD: every API entry point clears it, runs all of its CUDA work on the calling thread and reports
// `failed` as the API's error value.  Two contexts (one per GPU) driven from two threads therefore never see -- or clear --
// each other's failures; the worker threads of the multi-GPU entry point own one latch each.
#include "common.cuh"

namespace wb {

static thread_local bool t_failed = false;

void cuda_fail(cudaError_t e, const char * expr, const char * file, int line) {
    t_failed = true;
    fprintf(stderr, "whisper_b200: CUDA error %d (%s) at %s:%d: %s\n", (int) e, cudaGetErrorString(e), file, line, expr);
}
bool cuda_failed() { return t_failed; }
void cuda_clear_failure() {
    t_failed = false;
    cudaGetLastError();
}

}  // namespace wb
