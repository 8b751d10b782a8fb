// probe.cu -- measurement helper only (bench.py): a dependent-chain-free FFMA loop whose rate is the
// FP32 (non-tensor) pipe peak the LK roofline fraction is quoted against.
#include "common.cuh"

namespace {
__global__ void __launch_bounds__(256) k_probe_fma(int iters, float seed, float *out)
{
    float a0 = seed + threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f,
          a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f, c = 0.001f;
    for (int i = 0; i < iters; i++) {
        a0 = __fmaf_rn(a0, m, c); a1 = __fmaf_rn(a1, m, c); a2 = __fmaf_rn(a2, m, c); a3 = __fmaf_rn(a3, m, c);
        a4 = __fmaf_rn(a4, m, c); a5 = __fmaf_rn(a5, m, c); a6 = __fmaf_rn(a6, m, c); a7 = __fmaf_rn(a7, m, c);
    }
    float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678f) out[0] = s;      // never true; keeps the loop alive
}
}  // namespace

// flops executed = blocks * 256 * iters * 8 * 2
LKF_EXPORT int lkf_probe_fp32_fma(int blocks, int iters, float *out, lkf_stream stream)
{
    if (blocks < 1 || iters < 1 || !out) return LKF_ERR_ARG;
    k_probe_fma<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 1.0f, out);
    LKF_LAUNCH_CHECK(nullptr);
    return LKF_OK;
}
