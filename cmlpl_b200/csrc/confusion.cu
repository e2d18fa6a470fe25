// Integer confusion matrix behind CalAccuracy (tools/hyper_tools.py:208-223):
// cm[label, pred] += 1.  OA = trace/n, PA_i = cm[i,i]/rowsum_i, kappa from row/col sums --
// all float64 host arithmetic on these exact counts (cmlpl_b200/tools/hyper_tools.py).
// Per-CTA shared-memory histogram (32-bit), then one 64-bit global atomic per non-zero bin.
#include "common.cuh"

namespace cmlpl {

__global__ void __launch_bounds__(256)
confusion_kernel(const uint8_t* __restrict__ pred, const int64_t* __restrict__ label, int64_t n, int C,
                 unsigned long long* __restrict__ cm) {
  extern __shared__ unsigned int hist[];  // [C*C]
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t l = label[i];
    const int p = pred[i];
    if (l >= 0 && l < C && p < C) atomicAdd(&hist[int(l) * C + p], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * C; i += blockDim.x)
    if (hist[i]) atomicAdd(&cm[i], (unsigned long long)hist[i]);
}

}  // namespace cmlpl

extern "C" int cmlpl_confusion_i64(const uint8_t* pred, const int64_t* label, int64_t n, int num_classes,
                                   int64_t* cm, cmlpl_stream_t stream) {
  using namespace cmlpl;
  CMLPL_CHECK_ARG(n >= 0 && num_classes > 0 && num_classes <= 64, "confusion: bad dims (n=%lld C=%d)", (long long)n,
                  num_classes);
  if (n == 0) return CMLPL_OK;                             // empty inputs: nothing to count (pointers may be null)
  CMLPL_CHECK_ARG(pred && label && cm, "confusion: null pointer");
  // each CTA sees < 2^32 samples: grid-stride over n with at most 2^31 per CTA by construction
  int64_t grid = (n + 255) / 256;
  const int64_t cap = int64_t(sm_count()) * 8;
  if (grid > cap) grid = cap;
  CMLPL_CHECK_ARG(n / grid < (int64_t(1) << 31), "confusion: n too large for 32-bit per-CTA bins");
  confusion_kernel<<<int(grid), 256, sizeof(unsigned int) * num_classes * num_classes,
                     static_cast<cudaStream_t>(stream)>>>(pred, label, n, num_classes,
                                                          reinterpret_cast<unsigned long long*>(cm));
  CMLPL_CHECK_LAUNCH("confusion");
  return CMLPL_OK;
}
