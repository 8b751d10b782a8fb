// probe.cu -- measurement helpers only (bench.py, tools/): dependent-chain-free instruction loops whose rates are the
// non-tensor pipe peaks the LK roofline fractions are quoted against: FFMA (the FP32 pipe), IMAD and
// IDP.2A (the integer multiply-add datapath the tracker's sampling and mismatch sums run on), SHF/LOP3 (integer ALU).
// 16 independent chains per thread, 8 rounds per loop trip: 128 probe instructions against ~3 of loop overhead.
#include "common.cuh"

namespace {
template <int KIND>
__global__ void __launch_bounds__(256) k_probe(int iters, int seed, float *out)
{
    constexpr int NC = 16;
    float f[NC];
    int v[NC];
#pragma unroll
    for (int i = 0; i < NC; i++) { f[i] = (float)(seed + threadIdx.x + i); v[i] = seed * 7 + threadIdx.x + i; }
    const float m = 0.999f, c = 0.001f;
    const int im = seed | 3, ic = seed + 11;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < NC; i++) {
                if (KIND == 0) f[i] = __fmaf_rn(f[i], m, c);
                else if (KIND == 1) v[i] = v[i] * im + ic;                                   // IMAD
                else if (KIND == 2) asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(v[i]) : "r"(im), "r"(ic));   // IDP.2A
                else if (KIND == 3) v[i] = __funnelshift_r(v[i], ic, 5) ^ im;                // SHF + LOP3 (two ALU instructions)
                else if (KIND == 4) { f[i] = __fmaf_rn(f[i], m, c); v[i] = __funnelshift_r(v[i], ic, 5); } // FFMA + SHF
                else if (KIND == 5) {                                                       // IMAD + SHF on independent chains
                    if (i & 1) v[i] = v[i] * im + ic; else v[i] = __funnelshift_r(v[i], ic, 5);
                } else if (KIND == 6) {                                                     // IMAD + FFMA
                    if (i & 1) v[i] = v[i] * im + ic; else f[i] = __fmaf_rn(f[i], m, c);
                } else {                                                                    // IDP.2A + SHF + FFMA (i % 3)
                    if (i % 3 == 0) asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(v[i]) : "r"(im), "r"(ic));
                    else if (i % 3 == 1) v[i] = __funnelshift_r(v[i], ic, 5);
                    else f[i] = __fmaf_rn(f[i], m, c);
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NC; i++) s += f[i] + (float)v[i];
    if (s == 12345.678f) out[0] = s;      // never true; keeps the loops alive
}
}  // namespace

// kind 0: FFMA, 1: IMAD, 2: IDP.2A, 3: SHF+LOP3 pairs, 4: FFMA+SHF pairs, 5: IMAD/SHF alternating, 6: IMAD/FFMA alternating,
// 7: IDP.2A/SHF/FFMA in turn.  Probe instructions executed per thread: iters * 128 (kinds 3, 4: 256).  Threads: blocks * 256.
LKF_EXPORT int lkf_probe_pipe(int kind, int blocks, int iters, float *out, lkf_stream stream)
{
    if (blocks < 1 || iters < 1 || !out) return LKF_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    switch (kind) {
        case 0: k_probe<0><<<blocks, 256, 0, s>>>(iters, 1, out); break;
        case 1: k_probe<1><<<blocks, 256, 0, s>>>(iters, 1, out); break;
        case 2: k_probe<2><<<blocks, 256, 0, s>>>(iters, 1, out); break;
        case 3: k_probe<3><<<blocks, 256, 0, s>>>(iters, 1, out); break;
        case 4: k_probe<4><<<blocks, 256, 0, s>>>(iters, 1, out); break;
        case 5: k_probe<5><<<blocks, 256, 0, s>>>(iters, 1, out); break;
        case 6: k_probe<6><<<blocks, 256, 0, s>>>(iters, 1, out); break;
        case 7: k_probe<7><<<blocks, 256, 0, s>>>(iters, 1, out); break;
        default: return LKF_ERR_ARG;
    }
    LKF_LAUNCH_CHECK(nullptr);
    return LKF_OK;
}

// FFMA-rate probe: flops executed = blocks * 256 * iters * 128 * 2
LKF_EXPORT int lkf_probe_fp32_fma(int blocks, int iters, float *out, lkf_stream stream)
{
    return lkf_probe_pipe(0, blocks, iters, out, stream);
}
