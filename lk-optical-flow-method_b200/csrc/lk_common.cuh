// lk_common.cuh -- helpers shared by the LK trackers (lk.cu: warp-per-feature and thread-per-feature kernels;
// lk_dense.cu: tiled dense-grid kernel).  Everything here follows SURVEY.md Appendix A.6 / oracle lko_pyrlk().
#pragma once
#include "common.cuh"
#include <limits.h>

namespace {

struct LKLevels {
    const uint8_t *prev[LKF_MAX_LEVELS];
    const uint8_t *next[LKF_MAX_LEVELS];
    long long prp[LKF_MAX_LEVELS], pip[LKF_MAX_LEVELS], nrp[LKF_MAX_LEVELS], nip[LKF_MAX_LEVELS];
    int H[LKF_MAX_LEVELS], W[LKF_MAX_LEVELS];
};

__device__ __forceinline__ int descale(int v, int n) { return (v + (1 << (n - 1))) >> n; }

// exact warp sum of int32 lane partials -> int64 (two REDUX instructions)
__device__ __forceinline__ long long warp_sum(int v)
{
    int hi = __reduce_add_sync(0xffffffffu, v >> 16);
    unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)(v & 0xffff));
    return ((long long)hi << 16) + (long long)lo;
}
__device__ __forceinline__ long long warp_sum(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// cv2: delta.ddot(delta) <= epsilon, the dot product accumulated in double.  A float32 estimate decides all but the
// borderline cases (relative error of the estimate < 3e-7), which take the exact double path.
// (the float bounds are prepared on the host: kernel parameters cost neither registers nor instructions per iteration)
struct LKEps {
    double e2;          // epsilon^2 after cv2's clamping
    float lo, hi;       // (float)e2 * 0.999999f, (float)e2 * 1.000001f
};
static inline LKEps make_eps(double e2)
{
    LKEps r;
    r.e2 = e2;
    const float e = (float)e2;
    r.lo = e * 0.999999f; r.hi = e * 1.000001f;
    return r;
}
__device__ __forceinline__ bool lk_step_below_eps(float dx, float dy, const LKEps &eps)
{
    const float f = dx * dx + dy * dy;
    if (f < eps.lo) return true;
    if (f > eps.hi) return false;
    return __dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= eps.e2;
}
// cv2: std::abs(x) < 0.01 with a float x and the double constant 0.01: the largest float below 0.01 is fl(0.01)
__device__ __forceinline__ bool lk_below_001(float x) { return fabsf(x) <= 0.00999999977648258209228515625f; }

__device__ __forceinline__ void weights(float a, float b, int &w00, int &w01, int &w10, int &w11)
{
    float na = __fsub_rn(1.f, a), nb = __fsub_rn(1.f, b);
    w00 = __float2int_rn(__fmul_rn(__fmul_rn(na, nb), 16384.f));
    w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, nb), 16384.f));
    w10 = __float2int_rn(__fmul_rn(__fmul_rn(na, b), 16384.f));
    w11 = 16384 - w00 - w01 - w10;
}

__device__ __forceinline__ bool out_of_level(float fx, float fy, int ix, int iy, int w, int h, int W, int H)
{
    // non-finite coordinates behave like cv2 on x86 (cvFloor -> INT_MIN -> rejected)
    return !(isfinite(fx) && isfinite(fy)) || ix < -w || ix >= W || iy < -h || iy >= H;
}

// two-way dot product: signed 16-bit halves of `a` (Q14 weights, -1..16384) times unsigned bytes 0,1 (lo) / 2,3 (hi) of `b`
__device__ __forceinline__ int dp2a_lo(int a, unsigned b, int c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi(int a, unsigned b, int c)
{
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <int NB>
__device__ __noinline__ void load_row_bytes_edge(const uint8_t *__restrict__ row, int W, int gx, uint32_t *w)
{   // cold: the row segment touches the border of the level (or the plane is not 4-byte aligned)
#pragma unroll 1
    for (int i = 0; i < (NB + 3) / 4; i++) {
        uint32_t v = 0;
#pragma unroll
        for (int c = 0; c < 4; c++)
            if (4 * i + c < NB) v |= (uint32_t)row[reflect101(gx + 4 * i + c, W)] << (8 * c);
        w[i] = v;
    }
}

// NB consecutive bytes of row gy starting at column gx (REFLECT_101 outside the level), packed from byte 0 of w[0]
template <int NB>
__device__ __forceinline__ void load_row_bytes(const uint8_t *__restrict__ P, long long rp, int W, int H, bool al4, int gx, int gy,
                                               uint32_t (&w)[(NB + 3) / 4])
{
    constexpr int NW = (NB + 3) / 4;
    const uint8_t *row = P + (long long)((unsigned)gy < (unsigned)H ? gy : reflect101(gy, H)) * rp;
    const int ax = (gx >> 2) << 2;
    if (al4 && ax >= 0 && ax + 4 * (NW + 1) <= W) {
        const uint32_t *q = reinterpret_cast<const uint32_t *>(row + ax);
        uint32_t t[NW + 1];
#pragma unroll
        for (int i = 0; i <= NW; i++) t[i] = __ldg(q + i);
        const int sh = (gx & 3) * 8;
#pragma unroll
        for (int i = 0; i < NW; i++) w[i] = __funnelshift_r(t[i], t[i + 1], sh);
    } else {
        uint32_t tmp[NW];
        load_row_bytes_edge<NB>(row, W, gx, tmp);
#pragma unroll
        for (int i = 0; i < NW; i++) w[i] = tmp[i];
    }
}
template <int NW>
__device__ __forceinline__ int byte_of(const uint32_t (&w)[NW], int c) { return (int)((w[c >> 2] >> (8 * (c & 3))) & 255u); }

// bilinear Q14 sums of N pixels from two packed byte rows: v[t] = top(t) w00 + top(t+1) w01 + bot(t) w10 + bot(t+1) w11,
// rounded and shifted to I*32 ((sum + 2^8) >> 9)
template <int N, int NW>
__device__ __forceinline__ void bilinear_packed(const uint32_t (&top)[NW], const uint32_t (&bot)[NW], int first, int wT, int wB,
                                                int (&v)[N])
{
#pragma unroll
    for (int t = 0; t < N; t++) {
        const int c = first + t, i = c >> 2, j = c & 3;
        uint32_t wa, wc;
        if (j & 1) {
            wa = i + 1 < NW ? __funnelshift_r(top[i], top[i + 1], 8) : top[i] >> 8;
            wc = i + 1 < NW ? __funnelshift_r(bot[i], bot[i + 1], 8) : bot[i] >> 8;
        } else { wa = top[i]; wc = bot[i]; }
        int acc = (j & 2) ? dp2a_hi(wT, wa, 256) : dp2a_lo(wT, wa, 256);
        v[t] = ((j & 2) ? dp2a_hi(wB, wc, acc) : dp2a_lo(wB, wc, acc)) >> 9;
    }
}

}  // namespace
