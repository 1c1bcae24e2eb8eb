// Process-wide CUDA error latch (no exception may cross the C ABI; see common.cuh).
#include "common.cuh"

#include <atomic>

namespace wb {

static std::atomic<bool> g_failed{false};

void cuda_fail(cudaError_t e, const char * expr, const char * file, int line) {
    g_failed.store(true);
    fprintf(stderr, "whisper_b200: CUDA error %d (%s) at %s:%d: %s\n", (int) e, cudaGetErrorString(e), file, line, expr);
}
bool cuda_failed() { return g_failed.load(); }
void cuda_clear_failure() {
    g_failed.store(false);
    cudaGetLastError();
}

}  // namespace wb
