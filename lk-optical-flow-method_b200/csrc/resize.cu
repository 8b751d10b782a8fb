// resize.cu -- "next" row: imutils.resize(frame, width=860) = cv2.resize(..., INTER_AREA) with a
// non-integer scale (VP.py:433,531; 1280x720 -> 860x483).  Same float32 operation order as cv2's ResizeArea
// (oracle lko_resize_area): per source row a horizontal weighted sum accumulated in table order,
// then a vertical weighted accumulation, then round-half-even + saturate.  cv2's multiplies and adds are never
// contracted (the one FMA of the strip kernel restates conversion + multiply exactly, see byte_biased).
// The coefficient tables are built on the host (double arithmetic, exactly cv2's computeResizeAreaTab).
#include "common.cuh"

namespace {

// tab: for destination index d the entries [ofs[d], ofs[d+1]) of (src index, alpha)
__global__ void __launch_bounds__(256)
k_resize_area(const uint8_t *__restrict__ src, long long srp, long long sip, int cn, uint8_t *__restrict__ dst, long long drp,
              long long dip, int Hd, int Wd, const int *__restrict__ xofs, const int *__restrict__ xsi,
              const float *__restrict__ xal, const int *__restrict__ yofs, const int *__restrict__ ysi,
              const float *__restrict__ yal)
{
    const int dx = blockIdx.x * 64 + (threadIdx.x & 63), dy = blockIdx.y * 4 + (threadIdx.x >> 6), b = blockIdx.z;
    if (dx >= Wd || dy >= Hd) return;
    const uint8_t *S = src + (long long)b * sip;
    const int x0 = xofs[dx], x1 = xofs[dx + 1], y0 = yofs[dy], y1 = yofs[dy + 1];
    for (int c = 0; c < cn; c++) {
        float sum = 0.f;
        for (int j = y0; j < y1; j++) {
            const uint8_t *row = S + (long long)ysi[j] * srp;
            float buf = 0.f;
            for (int k = x0; k < x1; k++) buf = __fadd_rn(buf, __fmul_rn((float)row[xsi[k] * cn + c], xal[k]));
            const float t = __fmul_rn(yal[j], buf);
            sum = (j == y0) ? t : __fadd_rn(sum, t);
        }
        float r = rintf(sum);
        dst[(long long)b * dip + (long long)dy * drp + dx * cn + c] = (uint8_t)(r < 0.f ? 0.f : (r > 255.f ? 255.f : r));
    }
}

// Fast form for 3-channel frames whose horizontal footprint is at most 4 source pixels (scale < 3, e.g. 1920 -> 860 and
// 1280 -> 860): one thread per destination pixel, all three channels at once (three independent accumulation chains),
// the pixel's table entries held in registers, exact u8 -> f32 by bit pattern.  Same float32 operation order as above.
__device__ __forceinline__ float u8_to_f32(unsigned v) { return __uint_as_float(0x4B000000u | v) - 8388608.0f; }

__global__ void __launch_bounds__(256)
k_resize_area_bgr4(const uint8_t *__restrict__ src, long long srp, long long sip, uint8_t *__restrict__ dst, long long drp,
                   long long dip, int Hd, int Wd, const int *__restrict__ xofs, const int *__restrict__ xsi,
                   const float *__restrict__ xal, const int *__restrict__ yofs, const int *__restrict__ ysi,
                   const float *__restrict__ yal)
{
    const int dx = blockIdx.x * 64 + (threadIdx.x & 63), dy = blockIdx.y * 4 + (threadIdx.x >> 6), b = blockIdx.z;
    if (dx >= Wd || dy >= Hd) return;
    const uint8_t *S = src + (long long)b * sip;
    const int x0 = __ldg(xofs + dx), nx = __ldg(xofs + dx + 1) - x0, y0 = __ldg(yofs + dy), y1 = __ldg(yofs + dy + 1);
    int xo[4];
    float xa[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const bool on = k < nx;
        xo[k] = on ? 3 * __ldg(xsi + x0 + k) : 0;
        xa[k] = on ? __ldg(xal + x0 + k) : 0.f;
    }
    float sb = 0.f, sg = 0.f, sr = 0.f;
    for (int j = y0; j < y1; j++) {
        const uint8_t *row = S + (long long)__ldg(ysi + j) * srp;
        const float ya = __ldg(yal + j);
        float bb = 0.f, bg = 0.f, br = 0.f;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k < nx) {
                const uint8_t *p = row + xo[k];
                bb = __fadd_rn(bb, __fmul_rn(u8_to_f32(p[0]), xa[k]));
                bg = __fadd_rn(bg, __fmul_rn(u8_to_f32(p[1]), xa[k]));
                br = __fadd_rn(br, __fmul_rn(u8_to_f32(p[2]), xa[k]));
            }
        const float tb = __fmul_rn(ya, bb), tg = __fmul_rn(ya, bg), tr = __fmul_rn(ya, br);
        if (j == y0) { sb = tb; sg = tg; sr = tr; }
        else { sb = __fadd_rn(sb, tb); sg = __fadd_rn(sg, tg); sr = __fadd_rn(sr, tr); }
    }
    uint8_t *o = dst + (long long)b * dip + (long long)dy * drp + 3 * dx;
    const float rb = rintf(sb), rg = rintf(sg), rr = rintf(sr);
    o[0] = (uint8_t)(rb < 0.f ? 0.f : (rb > 255.f ? 255.f : rb));
    o[1] = (uint8_t)(rg < 0.f ? 0.f : (rg > 255.f ? 255.f : rg));
    o[2] = (uint8_t)(rr < 0.f ? 0.f : (rr > 255.f ? 255.f : rr));
}

// Column-strip form of the above (4-byte aligned source rows): a thread owns one destination column and walks down
// RS_ROWS destination rows.  Every source row is loaded and reduced horizontally ONCE (a row that straddles two
// destination rows feeds both, cv2 recomputes it), as four aligned 32-bit words realigned by funnel shifts; a byte becomes
// the float 2^23 + byte by one PRMT and its weighted value by one exact FMA (see byte_biased).  Table entries a pixel
// does not have carry alpha = 0: adding +0 leaves the non-negative partial sums unchanged, so the float32 sequence is
// the one of cv2.
constexpr int RS_ROWS = 16;

// cv2's saturate_cast<uchar>(float): round half to even, clamp to [0, 255] -- one conversion instruction
__device__ __forceinline__ uint8_t sat_u8(float v)
{
    unsigned r;
    asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(v));
    return (uint8_t)r;
}

// byte k of w as the float 2^23 + byte (exact).  The product byte * alpha rounded to float32 -- cv2's
// `S[sx] * alpha` -- is then ONE fused multiply-add: fma(2^23 + byte, alpha, -(2^23 * alpha)); 2^23 * alpha is exact,
// so the single rounding applies to the exact value byte * alpha.
__device__ __forceinline__ float byte_biased(uint32_t w, int k)
{
    const uint32_t sel = k == 0 ? 0x7540u : (k == 1 ? 0x7541u : (k == 2 ? 0x7542u : 0x7543u));
    return __uint_as_float(__byte_perm(w, 0x4B000000u, sel));
}

__global__ void __launch_bounds__(128)
k_resize_area_bgr4_strip(const uint8_t *__restrict__ src, long long srp, long long sip, uint8_t *__restrict__ dst, long long drp,
                         long long dip, int Hd, int Wd, const int *__restrict__ xofs, const int *__restrict__ xsi,
                         const float *__restrict__ xal, const int *__restrict__ yofs, const int *__restrict__ ysi,
                         const float *__restrict__ yal)
{
    const int dx = blockIdx.x * 128 + threadIdx.x, b = blockIdx.z;
    if (dx >= Wd) return;
    const int dy0 = blockIdx.y * RS_ROWS, dy1 = min(Hd, dy0 + RS_ROWS);
    const uint8_t *S = src + (long long)b * sip;
    const int x0 = __ldg(xofs + dx), nx = __ldg(xofs + dx + 1) - x0;
    float xa[4], xc[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { xa[k] = k < nx ? __ldg(xal + x0 + k) : 0.f; xc[k] = -8388608.0f * xa[k]; }
    const int byte0 = 3 * __ldg(xsi + x0), a0 = byte0 & ~3, sh = (byte0 & 3) * 8;
    // the 16 bytes read per row must stay inside the row pitch (threads at the right edge take the byte path)
    const bool words = a0 + 16 <= srp;
    uint8_t *o = dst + (long long)b * dip + (long long)dy0 * drp + 3 * dx;
    float sb = 0.f, sg = 0.f, sr = 0.f;
    int dy = dy0;
    int j = __ldg(yofs + dy0), jn = __ldg(yofs + dy0 + 1);       // jn: one past the last table entry of destination row dy
    bool first = true;                                           // next entry opens a destination row
    const int jend = __ldg(yofs + dy1);
    // area down-scaling uses every source row once, in order, and a source row feeds at most two destination rows:
    // walk the source rows sy0..sy1 with a running pointer and hand each reduced row to its one or two table entries
    const int sy0 = __ldg(ysi + j), sy1 = __ldg(ysi + jend - 1);
    const uint8_t *row = S + (long long)sy0 * srp;
    int sy_next = sy0;
    float ya = __ldg(yal + j);
    for (int sy = sy0; sy <= sy1; sy++, row += srp) {
        float v[12];                                   // 2^23 + byte
        if (words) {
            const uint32_t *q = reinterpret_cast<const uint32_t *>(row + a0);
            const uint32_t q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3);
            const uint32_t w0 = __funnelshift_r(q0, q1, sh), w1 = __funnelshift_r(q1, q2, sh), w2 = __funnelshift_r(q2, q3, sh);
#pragma unroll
            for (int t = 0; t < 4; t++) { v[t] = byte_biased(w0, t); v[4 + t] = byte_biased(w1, t); v[8 + t] = byte_biased(w2, t); }
        } else {
#pragma unroll
            for (int t = 0; t < 12; t++) v[t] = __uint_as_float(0x4B000000u | (t < 3 * nx ? (uint32_t)row[byte0 + t] : 0u));
        }
        // buf = ((s0 a0 + s1 a1) + s2 a2) + s3 a3 per channel (cv2 starts from 0: 0 + s0 a0 = s0 a0)
        float bb = __fmaf_rn(v[0], xa[0], xc[0]), bg = __fmaf_rn(v[1], xa[0], xc[0]), br = __fmaf_rn(v[2], xa[0], xc[0]);
#pragma unroll
        for (int k = 1; k < 4; k++) {
            bb = __fadd_rn(bb, __fmaf_rn(v[3 * k], xa[k], xc[k]));
            bg = __fadd_rn(bg, __fmaf_rn(v[3 * k + 1], xa[k], xc[k]));
            br = __fadd_rn(br, __fmaf_rn(v[3 * k + 2], xa[k], xc[k]));
        }
#pragma unroll
        for (int rep = 0; rep < 2; rep++) {
            if (sy_next != sy) break;
            const float tb = __fmul_rn(ya, bb), tg = __fmul_rn(ya, bg), tr = __fmul_rn(ya, br);
            if (first) { sb = tb; sg = tg; sr = tr; first = false; }
            else { sb = __fadd_rn(sb, tb); sg = __fadd_rn(sg, tg); sr = __fadd_rn(sr, tr); }
            j++;
            if (j == jn) {
                o[0] = sat_u8(sb); o[1] = sat_u8(sg); o[2] = sat_u8(sr);
                o += drp; dy++; first = true;
                if (dy < dy1) jn = __ldg(yofs + dy + 1);
            }
            if (j < jend) { sy_next = __ldg(ysi + j); ya = __ldg(yal + j); } else sy_next = -1;
        }
    }
}

}  // namespace

LKF_EXPORT int lkf_resize_area_u8(const lkf_image *src, const lkf_image *dst, const int32_t *xofs, const int32_t *xsi,
                                  const float *xalpha, const int32_t *yofs, const int32_t *ysi, const float *yalpha,
                                  lkf_stream stream)
{
    if (!src || !dst || !xofs || !xsi || !xalpha || !yofs || !ysi || !yalpha) return LKF_ERR_ARG;
    if (!img_ok(src, src->channels, 1) || !img_ok(dst, dst->channels, 1) || src->channels != dst->channels ||
        src->batch != dst->batch || (src->channels != 1 && src->channels != 3))
        return LKF_ERR_ARG;
    if (dst->width >= src->width || dst->height >= src->height) return LKF_ERR_UNSUPPORTED;      // area DOWN-scaling only
    if (src->width % dst->width == 0 && src->height % dst->height == 0) return LKF_ERR_UNSUPPORTED;  // cv2's integer fast path
    dim3 grid((dst->width + 63) / 64, (dst->height + 3) / 4, dst->batch);
    // horizontal footprint of a destination pixel: at most floor(scale) + 2 table entries
    if (src->channels == 3 && 2LL * src->width < 5LL * dst->width && aligned_to(src, 4) && src->row_pitch >= 16) {
        dim3 g((dst->width + 127) / 128, (dst->height + RS_ROWS - 1) / RS_ROWS, dst->batch);
        k_resize_area_bgr4_strip<<<g, 128, 0, (cudaStream_t)stream>>>((const uint8_t *)src->data, src->row_pitch, src->image_pitch,
                                                                      (uint8_t *)dst->data, dst->row_pitch, dst->image_pitch,
                                                                      dst->height, dst->width, xofs, xsi, xalpha, yofs, ysi, yalpha);
        LKF_LAUNCH_CHECK(nullptr);
        return LKF_OK;
    }
    if (src->channels == 3 && 2LL * src->width < 5LL * dst->width) {      // scale < 2.5 -> at most 4 table entries per pixel
        k_resize_area_bgr4<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint8_t *)src->data, src->row_pitch, src->image_pitch,
                                                                  (uint8_t *)dst->data, dst->row_pitch, dst->image_pitch, dst->height,
                                                                  dst->width, xofs, xsi, xalpha, yofs, ysi, yalpha);
        LKF_LAUNCH_CHECK(nullptr);
        return LKF_OK;
    }
    k_resize_area<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint8_t *)src->data, src->row_pitch, src->image_pitch, src->channels,
                                                         (uint8_t *)dst->data, dst->row_pitch, dst->image_pitch, dst->height,
                                                         dst->width, xofs, xsi, xalpha, yofs, ysi, yalpha);
    LKF_LAUNCH_CHECK(nullptr);
    return LKF_OK;
}
