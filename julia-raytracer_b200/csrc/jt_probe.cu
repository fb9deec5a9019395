// jt_probe.cu -- bandwidth probes for the roofline denominators ("measure an L2 bandwidth peak with a
// micro-benchmark on the box"). Every array the traversal kernel walks is L2-resident for four of the five BASELINE
// configs, so the HBM peak of MEASURED_PEAKS.json is only a yardstick there; bench.py reports the as-implemented
// traffic against the number measured here as well.
#include <cuda_runtime.h>

#include "jt_internal.h"

// Grid-stride 128-bit loads over `n16` uint4 words, `reps` passes; the XOR reduction keeps the loads alive.
__global__ void __launch_bounds__(256) k_probe_read(const uint4* __restrict__ buf, long long n16, int reps,
                                                    unsigned* __restrict__ sink) {
  unsigned acc = 0u;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; r++) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
      uint4 v = __ldcg(buf + i);  // cache-global: L2 hit, no L1 allocation (what node / triangle fetches see on a miss in L1)
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}

extern "C" int jt_probe_read_bandwidth(int device, int64_t bytes, int reps, float* gbs_out) {
  if (!gbs_out || bytes < 4096 || reps < 1) return jt_set_error(JT_ERR_INVALID, "jt_probe_read_bandwidth: bad argument");
  int ndev = jt_device_count();
  if (ndev <= 0) return jt_set_error(JT_ERR_NO_DEVICE, "no CUDA device visible: libjtrace_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return jt_set_error(JT_ERR_INVALID, "device %d out of range", device);
  cudaError_t e = cudaSetDevice(device);
  void* buf = nullptr;
  unsigned* sink = nullptr;
  cudaEvent_t a = nullptr, b = nullptr;
  int sms = 0;
  float best = 0.0f;
  const long long n16 = bytes / 16;
  if (e == cudaSuccess) e = cudaMalloc(&buf, (size_t)n16 * 16);
  if (e == cudaSuccess) e = cudaMalloc((void**)&sink, 4);
  if (e == cudaSuccess) e = cudaMemset(buf, 1, (size_t)n16 * 16);
  if (e == cudaSuccess) e = cudaEventCreate(&a);
  if (e == cudaSuccess) e = cudaEventCreate(&b);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (e == cudaSuccess) {
    const unsigned grid = (unsigned)sms * 8u;  // one resident wave: 8 blocks of 256 threads per SM
    k_probe_read<<<grid, 256>>>((const uint4*)buf, n16, 2, sink);  // warm the L2
    for (int trial = 0; trial < 5 && e == cudaSuccess; trial++) {
      cudaEventRecord(a);
      k_probe_read<<<grid, 256>>>((const uint4*)buf, n16, reps, sink);
      cudaEventRecord(b);
      e = cudaEventSynchronize(b);
      float ms = 0.0f;
      if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, a, b);
      if (e == cudaSuccess && ms > 0.0f) {
        float gbs = (float)((double)n16 * 16.0 * reps / (ms * 1e-3) / 1e9);
        if (gbs > best) best = gbs;
      }
    }
  }
  if (a) cudaEventDestroy(a);
  if (b) cudaEventDestroy(b);
  cudaFree(buf);
  cudaFree(sink);
  if (e != cudaSuccess) return jt_set_error(JT_ERR_CUDA, "jt_probe_read_bandwidth: %s", cudaGetErrorString(e));
  *gbs_out = best;
  return JT_OK;
}
