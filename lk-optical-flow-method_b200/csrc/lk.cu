// lk.cu -- a7: the per-feature iterative pyramidal Lucas-Kanade tracker of cv.calcOpticalFlowPyrLK
// (VP.py:545-546, LKF.py:531-532; algorithm: SURVEY.md Appendix A.6; oracle: lko_pyrlk()).
//
// One warp per feature, all pyramid levels inside one launch.  Per level the warp
//   1. gathers the (w+3)x(h+3) uint8 neighbourhood of the feature in `prev` (REFLECT_101),
//   2. computes the Scharr derivatives on the fly for the (w+1)x(h+1) sample grid (zero outside the
//      level, exactly like cv2's zero-bordered derivative image) -- no derivative image in HBM,
//   3. builds the Q14-bilinear template patch (I*32, Ix, Iy as int16) in shared memory and the 2x2
//      structure tensor with exact integer lane partials + warp REDUX reductions,
//   4. runs the Newton iterations, sampling `next` bilinearly (Q14) and reducing the mismatch vector
//      the same way.  All float32 scalar steps use explicitly rounded intrinsics (no FMA contraction)
//      so they follow cv2's float32 sequence; only the (exact) integer sums replace cv2's float32
//      lane accumulation (difference <= 3e-4 px, see tests).
#include "common.cuh"

namespace {

constexpr int LK_WARPS = 8;

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

// Sum over the window of f(sample of J at the window pixel, template index).  Fast path when the
// (w+1)x(h+1) footprint is inside the level, mirrored indexing otherwise (warp-uniform choice).
template <class ACC, class F>
__device__ __forceinline__ void sweep_J(const uint8_t *__restrict__ J, long long rp, int W, int H, int ix, int iy,
                                        int w, int h, int w00, int w01, int w10, int w11, int lane, F f)
{
    const int area = w * h;
    int x = lane % w, y = lane / w;
    const bool inside = ix >= 0 && iy >= 0 && ix + w < W && iy + h < H;
    if (inside) {
        const uint8_t *base = J + (long long)iy * rp + ix;
        for (int idx = lane; idx < area; idx += 32) {
            const uint8_t *q = base + (long long)y * rp + x;
            int v = descale(q[0] * w00 + q[1] * w01 + q[rp] * w10 + q[rp + 1] * w11, 9);
            f(v, idx);
            x += 32;
            while (x >= w) { x -= w; y++; }
        }
    } else {
        for (int idx = lane; idx < area; idx += 32) {
            int x0 = reflect101(ix + x, W), x1 = reflect101(ix + x + 1, W);
            const uint8_t *r0 = J + (long long)reflect101(iy + y, H) * rp;
            const uint8_t *r1 = J + (long long)reflect101(iy + y + 1, H) * rp;
            int v = descale(r0[x0] * w00 + r0[x1] * w01 + r1[x0] * w10 + r1[x1] * w11, 9);
            f(v, idx);
            x += 32;
            while (x >= w) { x -= w; y++; }
        }
    }
}

template <class ACC>
__global__ void __launch_bounds__(LK_WARPS * 32)
k_lk_track(LKLevels lv, int n_levels, int w, int h, int max_count, double eps2, int flags, double min_eig_thr,
           const float *__restrict__ prev_pts, const int *__restrict__ counts, int max_pts,
           float *__restrict__ next_pts, uint8_t *__restrict__ status, float *__restrict__ err,
           int *__restrict__ iters_out, int per_warp_bytes)
{
    extern __shared__ __align__(16) unsigned char smem_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pt = blockIdx.x * LK_WARPS + warp, b = blockIdx.y;
    const int n = counts ? counts[b] : max_pts;
    if (pt >= n || pt >= max_pts) return;                       // warp-uniform; only __syncwarp below

    const int area = w * h, rw = w + 3, rh = h + 3, dw = w + 1, dh = h + 1;
    unsigned char *sm = smem_all + (size_t)warp * per_warp_bytes;
    short *pI = reinterpret_cast<short *>(sm);                            // area
    short2 *pD = reinterpret_cast<short2 *>(sm + ((area * 2 + 3) & ~3));  // area  (Ix, Iy)
    short2 *der = pD + area;                                              // dw*dh (Scharr at sample grid)
    unsigned char *raw = reinterpret_cast<unsigned char *>(der + dw * dh);  // rw*rh

    const long long pidx = (long long)b * max_pts + pt;
    const float ppx = prev_pts[2 * pidx], ppy = prev_pts[2 * pidx + 1];
    const float hx = __fmul_rn((float)(w - 1), 0.5f), hy = __fmul_rn((float)(h - 1), 0.5f);
    float npx = 0.f, npy = 0.f;
    if (flags & LKF_OPTFLOW_USE_INITIAL_FLOW) { npx = next_pts[2 * pidx]; npy = next_pts[2 * pidx + 1]; }
    bool st = true;
    float ev = 0.f;
    int iters = 0;
    const float FLT_SCALE = 1.f / (1 << 20);

    for (int l = n_levels - 1; l >= 0; --l) {
        const int W = lv.W[l], H = lv.H[l];
        const uint8_t *I = lv.prev[l] + (long long)b * lv.pip[l];
        const uint8_t *J = lv.next[l] + (long long)b * lv.nip[l];
        const long long irp = lv.prp[l], jrp = lv.nrp[l];
        const float sc = 1.f / (float)(1 << l);
        float px = __fmul_rn(ppx, sc), py = __fmul_rn(ppy, sc);
        float nx, ny;
        if (l == n_levels - 1) {
            if (flags & LKF_OPTFLOW_USE_INITIAL_FLOW) { nx = __fmul_rn(npx, sc); ny = __fmul_rn(npy, sc); }
            else { nx = px; ny = py; }
        } else { nx = __fmul_rn(npx, 2.f); ny = __fmul_rn(npy, 2.f); }
        npx = nx; npy = ny;
        px = __fsub_rn(px, hx); py = __fsub_rn(py, hy);
        const int ipx = __float2int_rd(px), ipy = __float2int_rd(py);
        if (out_of_level(px, py, ipx, ipy, w, h, W, H)) {
            if (l == 0) { st = false; ev = 0.f; }
            continue;
        }
        int w00, w01, w10, w11;
        weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy), w00, w01, w10, w11);

        // 1. raw neighbourhood, raw(j,i) = I[reflect(ipy-1+j)][reflect(ipx-1+i)]
        __syncwarp();
        {
            int i = lane % rw, j = lane / rw;
            for (int idx = lane; idx < rw * rh; idx += 32) {
                raw[idx] = I[(long long)reflect101(ipy - 1 + j, H) * irp + reflect101(ipx - 1 + i, W)];
                i += 32;
                while (i >= rw) { i -= rw; j++; }
            }
        }
        __syncwarp();
        // 2. Scharr on the (w+1)x(h+1) sample grid; zero outside the level
        {
            int i = lane % dw, j = lane / dw;
            for (int idx = lane; idx < dw * dh; idx += 32) {
                const unsigned char *r0 = raw + j * rw + i, *r1 = r0 + rw, *r2 = r1 + rw;
                int qx = ipx + i, qy = ipy + j;
                int ix = 0, iy = 0;
                if (qx >= 0 && qx < W && qy >= 0 && qy < H) {
                    ix = 3 * (r0[2] - r0[0]) + 10 * (r1[2] - r1[0]) + 3 * (r2[2] - r2[0]);
                    iy = 3 * (r2[0] - r0[0]) + 10 * (r2[1] - r0[1]) + 3 * (r2[2] - r0[2]);
                }
                der[idx] = make_short2((short)ix, (short)iy);
                i += 32;
                while (i >= dw) { i -= dw; j++; }
            }
        }
        __syncwarp();
        // 3. template patch + structure tensor
        ACC a11 = 0, a12 = 0, a22 = 0;
        {
            int x = lane % w, y = lane / w;
            for (int idx = lane; idx < area; idx += 32) {
                const unsigned char *r0 = raw + (y + 1) * rw + x + 1, *r1 = r0 + rw;
                int iv = descale(r0[0] * w00 + r0[1] * w01 + r1[0] * w10 + r1[1] * w11, 9);
                const short2 d00 = der[y * dw + x], d01 = der[y * dw + x + 1], d10 = der[(y + 1) * dw + x],
                             d11 = der[(y + 1) * dw + x + 1];
                int ixv = descale(d00.x * w00 + d01.x * w01 + d10.x * w10 + d11.x * w11, 14);
                int iyv = descale(d00.y * w00 + d01.y * w01 + d10.y * w10 + d11.y * w11, 14);
                pI[idx] = (short)iv;
                pD[idx] = make_short2((short)ixv, (short)iyv);
                a11 += ixv * ixv; a12 += ixv * iyv; a22 += iyv * iyv;
                x += 32;
                while (x >= w) { x -= w; y++; }
            }
        }
        const float A11 = __fmul_rn((float)warp_sum(a11), FLT_SCALE);
        const float A12 = __fmul_rn((float)warp_sum(a12), FLT_SCALE);
        const float A22 = __fmul_rn((float)warp_sum(a22), FLT_SCALE);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dd = __fsub_rn(A11, A22);
        const float disc = __fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12));
        const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(disc)), (float)(2 * w * h));
        if (flags & LKF_OPTFLOW_LK_GET_MIN_EIGENVALS) ev = minEig;
        if ((double)minEig < min_eig_thr || D < 1.1920929e-07f) {
            if (l == 0) st = false;
            continue;
        }
        D = __fdiv_rn(1.f, D);
        nx = __fsub_rn(nx, hx); ny = __fsub_rn(ny, hy);
        float pdx = 0.f, pdy = 0.f;
        __syncwarp();
        for (int j = 0; j < max_count; j++) {
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if (out_of_level(nx, ny, inx, iny, w, h, W, H)) {
                if (l == 0) st = false;
                break;
            }
            weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), w00, w01, w10, w11);
            ACC b1 = 0, b2 = 0;
            sweep_J<ACC>(J, jrp, W, H, inx, iny, w, h, w00, w01, w10, w11, lane, [&](int v, int idx) {
                int diff = v - pI[idx];
                short2 d = pD[idx];
                b1 += diff * d.x; b2 += diff * d.y;
            });
            iters++;
            const float fb1 = __fmul_rn((float)warp_sum(b1), FLT_SCALE);
            const float fb2 = __fmul_rn((float)warp_sum(b2), FLT_SCALE);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, fb2), __fmul_rn(A22, fb1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, fb1), __fmul_rn(A11, fb2)), D);
            nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
            npx = __fadd_rn(nx, hx); npy = __fadd_rn(ny, hy);
            if (__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= eps2) break;
            if (j > 0 && (double)fabsf(__fadd_rn(dx, pdx)) < 0.01 && (double)fabsf(__fadd_rn(dy, pdy)) < 0.01) {
                npx = __fsub_rn(npx, __fmul_rn(dx, 0.5f));
                npy = __fsub_rn(npy, __fmul_rn(dy, 0.5f));
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (st && l == 0 && !(flags & LKF_OPTFLOW_LK_GET_MIN_EIGENVALS)) {
            const float ex = __fsub_rn(npx, hx), ey = __fsub_rn(npy, hy);
            const int iex = __float2int_rd(ex), iey = __float2int_rd(ey);
            if (out_of_level(ex, ey, iex, iey, w, h, W, H)) { st = false; continue; }
            weights(__fsub_rn(ex, (float)iex), __fsub_rn(ey, (float)iey), w00, w01, w10, w11);
            ACC se = 0;
            sweep_J<ACC>(J, jrp, W, H, iex, iey, w, h, w00, w01, w10, w11, lane, [&](int v, int idx) {
                int diff = v - pI[idx];
                se += diff < 0 ? -diff : diff;
            });
            ev = __fdiv_rn((float)warp_sum(se), (float)(32 * w * h));
        }
    }
    if (lane == 0) {
        next_pts[2 * pidx] = npx;
        next_pts[2 * pidx + 1] = npy;
        status[pidx] = st ? 1 : 0;
        err[pidx] = st ? ev : 0.f;
        if (iters_out) iters_out[pidx] = iters;
    }
}

size_t lk_per_warp_bytes(int w, int h)
{
    size_t area = (size_t)w * h;
    size_t bytes = ((area * 2 + 3) & ~(size_t)3) + area * 4 + (size_t)(w + 1) * (h + 1) * 4 + (size_t)(w + 3) * (h + 3);
    return (bytes + 15) & ~(size_t)15;
}

}  // namespace

LKF_EXPORT int lkf_pyrlk_sparse(lkf_ctx *ctx, const lkf_lk_params *p, int n_levels, const lkf_image *prev_pyr,
                                const lkf_image *next_pyr, const float *prev_pts, const int32_t *pt_counts,
                                int max_pts, float *next_pts, uint8_t *status, float *err, int32_t *iters,
                                lkf_stream stream)
{
    if (!p || !prev_pyr || !next_pyr || n_levels < 1 || n_levels > LKF_MAX_LEVELS) return LKF_ERR_ARG;
    if (p->win_w <= 2 || p->win_h <= 2 || p->max_level < 0) return LKF_ERR_ARG;       // cv2: winSize > 2, maxLevel >= 0
    if (max_pts < 0 || (max_pts > 0 && (!prev_pts || !next_pts || !status || !err))) return LKF_ERR_ARG;
    if (max_pts == 0) return LKF_OK;
    LKLevels lv;
    const int B = prev_pyr[0].batch;
    for (int l = 0; l < n_levels; l++) {
        if (!img_ok(&prev_pyr[l], 1, 1) || !img_ok(&next_pyr[l], 1, 1)) return LKF_ERR_ARG;
        if (prev_pyr[l].batch != B || next_pyr[l].batch != B || prev_pyr[l].height != next_pyr[l].height ||
            prev_pyr[l].width != next_pyr[l].width)
            return LKF_ERR_ARG;
        if (l > 0 && (prev_pyr[l].height != (prev_pyr[l - 1].height + 1) / 2 ||
                      prev_pyr[l].width != (prev_pyr[l - 1].width + 1) / 2))
            return LKF_ERR_ARG;
        lv.prev[l] = (const uint8_t *)prev_pyr[l].data; lv.next[l] = (const uint8_t *)next_pyr[l].data;
        lv.prp[l] = prev_pyr[l].row_pitch; lv.pip[l] = prev_pyr[l].image_pitch;
        lv.nrp[l] = next_pyr[l].row_pitch; lv.nip[l] = next_pyr[l].image_pitch;
        lv.H[l] = prev_pyr[l].height; lv.W[l] = prev_pyr[l].width;
    }
    int max_count = p->max_count;
    double eps = p->epsilon;
    if (!(p->crit_type & LKF_TERM_CRITERIA_COUNT)) max_count = 30;
    else max_count = max_count < 0 ? 0 : (max_count > 100 ? 100 : max_count);
    if (!(p->crit_type & LKF_TERM_CRITERIA_EPS)) eps = 0.01;
    else eps = eps < 0 ? 0 : (eps > 10 ? 10 : eps);
    eps *= eps;
    const size_t pw = lk_per_warp_bytes(p->win_w, p->win_h);
    const size_t smem = pw * LK_WARPS;
    if (smem > 227 * 1024) return LKF_ERR_UNSUPPORTED;
    const int area = p->win_w * p->win_h;
    dim3 grid((max_pts + LK_WARPS - 1) / LK_WARPS, B);
    cudaStream_t s = (cudaStream_t)stream;
    if (area <= 2048) {      // int32 lane partials cannot overflow (<= 64 px/lane * 8160*4080)
        if (smem > 48 * 1024) LKF_CUDA(cudaFuncSetAttribute(k_lk_track<int>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_lk_track<int><<<grid, LK_WARPS * 32, smem, s>>>(lv, n_levels, p->win_w, p->win_h, max_count, eps, p->flags,
                                                          p->min_eig_threshold, prev_pts, pt_counts, max_pts, next_pts,
                                                          status, err, iters, (int)pw);
    } else {
        if (smem > 48 * 1024) LKF_CUDA(cudaFuncSetAttribute(k_lk_track<long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_lk_track<long long><<<grid, LK_WARPS * 32, smem, s>>>(lv, n_levels, p->win_w, p->win_h, max_count, eps, p->flags,
                                                                p->min_eig_threshold, prev_pts, pt_counts, max_pts,
                                                                next_pts, status, err, iters, (int)pw);
    }
    LKF_LAUNCH_CHECK(ctx);
    return LKF_OK;
}
