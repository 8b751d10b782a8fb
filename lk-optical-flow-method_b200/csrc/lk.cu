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
#include "lk_common.cuh"
#include "tma.cuh"
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int LK_WARPS = 8;


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
k_lk_track(LKLevels lv, int n_levels, int w, int h, int max_count, LKEps eps2, int flags, double min_eig_thr,
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
        const float sc = __int_as_float((127 - l) << 23);          // 2^-l, exact
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
            if (lk_step_below_eps(dx, dy, eps2)) break;
            if (j > 0 && lk_below_001(__fadd_rn(dx, pdx)) && lk_below_001(__fadd_rn(dy, pdy))) {
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


// ---------------------------------------------------------------------------------------------
// Specialised tracker for the window sizes the configs name (15x15, 21x21, 31x31): everything is a
// compile-time constant, the template patch lives in REGISTERS, each lane owns horizontal runs of
// RUN window pixels (2*(RUN+1) byte loads per run instead of 4*RUN), and `next` is sampled from a
// per-warp shared-memory tile (window + MARGIN on every side, loaded once per level with aligned
// 32-bit loads and re-centred only if the iterate walks out of it).
// ---------------------------------------------------------------------------------------------
constexpr int LKT_WARPS = 1;     // one feature per CTA: registers and shared memory are released the moment it converges

// TMA = true: tile and template neighbourhood arrive by cp.async.bulk.tensor (3D boxes TWC x TH / RWP x RH of the level
// planes): pitches are multiples of 16 bytes, buffers 128-byte aligned, three mbarriers follow the template patch.
template <int WW, int WH, int RUN, int MARGIN, int DWP_, int TWC_, bool TMA = false>
struct LKT {
    static constexpr int AREA = WW * WH, WW_ = WW, WH_ = WH, RUN_ = RUN;
    static constexpr int RPR = WW / RUN;                 // runs per window row
    static constexpr int NRUNS = RPR * WH;
    static constexpr int RPL = (NRUNS + 31) / 32;        // runs per lane
    static constexpr int RW = WW + 3, RH = WH + 3;
    // x origin of every staged block is a multiple of XAL bytes: 4 for the word loaders; 16 for TMA (the innermost box
    // coordinate of a uint8 tensor must be a multiple of 16 bytes -- measured: any other value raises an illegal-instruction fault)
    static constexpr bool TMA_ = TMA;
    static constexpr int XSH = TMA ? 4 : 2, XAL = 1 << XSH;
    static constexpr int RWW0 = (RW + (XAL - 1) + 3) / 4;             // words per raw row incl. the alignment slack
    static constexpr int RWW = TMA ? ((RWW0 + 3) / 4) * 4 : RWW0;
    static constexpr int RWP = RWW * 4;                               // raw pitch (bytes)
    static constexpr int RAW_LPR = RWW <= 8 ? 8 : 16;                 // lanes per raw row (power of two)
    static constexpr int TILE_WPR = TWC_ / 4;
    static constexpr int TILE_LPR = TILE_WPR <= 8 ? 8 : (TILE_WPR <= 16 ? 16 : 32);
    static constexpr int DW = WW + 1, DH = WH + 1, DWP = DWP_;      // DWP: der pitch (words), chosen bank-conflict-poor
    static constexpr int TWC = TWC_;                                  // tile pitch (bytes), multiple of 4
    static constexpr int TH = WH + 1 + 2 * MARGIN;
    // interior windows: U = Q14-weighted raw sums on the (WW+2)x(WH+2) grid, int32, pitch UP (bank-spread, 16-byte rows)
    static constexpr int USEG = (WW + 2 + 3) / 4, UP = 4 * USEG + ((4 * USEG) % 8 == 0 ? 4 : 0), UH = WH + 2;
    static constexpr int USEG6 = (WW + 2 + 5) / 6;                    // U is produced in runs of 6 values (3 x 64-bit stores)
    static_assert(6 * USEG6 <= UP && UP % 2 == 0, "U pitch");
    static constexpr int DER_ONLY = (DWP * DH * 4 + 15) & ~15, U_BYTES = (UP * UH * 4 + 15) & ~15;
    static constexpr int AL = TMA ? 128 : 16;
    static constexpr int DER_BYTES0 = DER_ONLY > U_BYTES ? DER_ONLY : U_BYTES;
    static constexpr int DER_BYTES = (DER_BYTES0 + AL - 1) / AL * AL, TILE_BYTES = (TWC * TH + AL - 1) / AL * AL;
    static constexpr int RAW_BYTES = (RWP * RH + 16 + AL - 1) / AL * AL;     // + slack: word loaders read up to 2 words past a run
    // template patch in shared memory.  A lane owns NPX = RUN * RPL window pixels (q = k * RUN + t); per lane the patch is a
    // list of 32-bit words holding two int16 each: X[i] = (Ix~(2i), Ix~(2i+1)), Y[i] likewise, then I[i] = (I~(2i), I~(2i+1)).
    // Stored chunk-major -- 16-byte chunk c of lane i at (c * 32 + i) * 16 -- so the 128-bit loads of a sweep are conflict-free.
    static constexpr int NPX = RUN * RPL, NPAIR = (NPX + 1) / 2;
    static constexpr int NW_D = 2 * NPAIR, NW_ALL = 3 * NPAIR;
    static constexpr int NCH_D = (NW_D + 3) / 4, NCH_ALL = (NW_ALL + 3) / 4;
    static constexpr int TPL_BYTES = NCH_ALL * 512;
    static constexpr int BAR_BYTES = TMA ? 128 : 0;                   // mbarriers: tile, raw[0], raw[1]
    static constexpr int PER_WARP = DER_BYTES + TILE_BYTES + 2 * RAW_BYTES + TPL_BYTES + BAR_BYTES;      // raw is double-buffered
    static_assert(!TMA || (TWC % 16 == 0 && RWP % 16 == 0), "TMA box rows are multiples of 16 bytes");
    static_assert(WW % RUN == 0, "RUN must divide the window width");
    static_assert(AREA <= 2048, "int32 lane partials");
    static_assert(TWC % 4 == 0 && TWC >= WW + 1 + 2 * MARGIN + XAL - 1, "tile pitch");
    static_assert(DWP >= DW, "der pitch");
    static_assert(RWW <= 16 && TILE_WPR <= 32, "loader lane mapping");
};


__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
struct SweepB { static constexpr bool value = false; };      // a window sweep that accumulates the mismatch vector
struct SweepErr { static constexpr bool value = true; };     // ... that accumulates sum |J~ - I~|

// Q14 bilinear samples of RUN consecutive pixels whose top-left samples start at byte `col` of the 4-byte aligned
// rows r0 (upper) and r1 (lower): v[t] = (r0[col+t] w00 + r0[col+t+1] w01 + r1[col+t] w10 + r1[col+t+1] w11 + 2^8) >> 9.
// Rows are read as aligned 32-bit words and realigned with funnel shifts; each pixel then costs two DP2A (the byte
// pair (t, t+1) is the low or high half of the aligned word or of the word shifted by one byte).
// DESCALE = false returns the raw Q14 sums (no rounding, no shift).
template <int RUN, bool DESCALE = true>
__device__ __forceinline__ void bilinear_run(const unsigned char *r0, const unsigned char *r1, int col, int wT, int wB, int (&v)[RUN])
{
    constexpr int NA = (RUN + 1 + 3) / 4;
    const uint32_t *q0 = reinterpret_cast<const uint32_t *>(r0) + (col >> 2);
    const uint32_t *q1 = reinterpret_cast<const uint32_t *>(r1) + (col >> 2);
    const int sh = (col & 3) * 8;
    uint32_t a[NA + 1], c[NA + 1];
#pragma unroll
    for (int i = 0; i <= NA; i++) { a[i] = q0[i]; c[i] = q1[i]; }
#pragma unroll
    for (int i = 0; i < NA; i++) { a[i] = __funnelshift_r(a[i], a[i + 1], sh); c[i] = __funnelshift_r(c[i], c[i + 1], sh); }
#pragma unroll
    for (int i = 0; i < NA; i++) {
        const uint32_t a1 = i + 1 < NA ? __funnelshift_r(a[i], a[i + 1], 8) : a[i] >> 8;
        const uint32_t c1 = i + 1 < NA ? __funnelshift_r(c[i], c[i + 1], 8) : c[i] >> 8;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int t = 4 * i + j;
            if (t < RUN) {
                const uint32_t wa = (j & 1) ? a1 : a[i], wc = (j & 1) ? c1 : c[i];
                int acc = (j & 2) ? dp2a_hi(wT, wa, DESCALE ? 256 : 0) : dp2a_lo(wT, wa, DESCALE ? 256 : 0);
                acc = (j & 2) ? dp2a_hi(wB, wc, acc) : dp2a_lo(wB, wc, acc);
                v[t] = DESCALE ? acc >> 9 : acc;
            }
        }
    }
}

// tile origin: window + MARGIN, x aligned down to 4; slid back inside the level when the window itself is inside
// (keeps the aligned 32-bit fast path for features near the border)
template <class T>
__device__ __forceinline__ void place_tile(int ix, int iy, int W, int H, int &tx0, int &ty0)
{
    constexpr int WW1 = T::DW, WH1 = T::DH;       // footprint of the window samples
    tx0 = ((ix - (T::TWC - WW1 - (T::XAL - 1)) / 2) >> T::XSH) << T::XSH;
    ty0 = iy - (T::TH - WH1) / 2;
    if (ix >= 0 && ix + WW1 <= W && W >= T::TWC) {
        const int c = min(max(tx0, 0), ((W - T::TWC) >> T::XSH) << T::XSH);
        if (ix + WW1 <= c + T::TWC) tx0 = c;           // (alignment can cost up to 3 columns at the right edge)
    }
    if (iy >= 0 && iy + WH1 <= H && H >= T::TH) ty0 = min(max(ty0, 0), H - T::TH);
}

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// One TMA box (NROWS rows of 4 NWORDS bytes at column gx0, row gy0 of image b) -> dst, completion on mbarrier `bar`.
// Only boxes inside the level (out-of-bounds elements would be zero-filled, cv2 mirrors) on levels with a tensor map;
// false: the caller copies with the word loaders.  One lane issues; the proxy fence orders the warp's earlier generic
// accesses of dst (made visible to that lane by the preceding __syncwarp) before the asynchronous write.
template <int NROWS, int NWORDS>
__device__ __forceinline__ bool tma_box(const CUtensorMap *map, bool eligible, int W, int H, int gx0, int gy0, int b, void *dst,
                                        uint32_t bar, int lane, bool fence = true)
{
    if (!(eligible && gx0 >= 0 && gy0 >= 0 && gx0 + 4 * NWORDS <= W && gy0 + NROWS <= H)) return false;     // warp-uniform
    if (lane == 0) {
        if (fence) fence_proxy_async();
        mbar_expect_tx(bar, NROWS * NWORDS * 4);
        tma_load_3d(smem_u32(dst), map, gx0, gy0, b, bar);
    }
    return true;
}

// Loads NROWS x NWORDS 32-bit words (rows gy0.., columns from the 4-aligned gx0) of a 4-byte aligned plane into `dst`
// (pitch NWORDS words): LPR lanes per row, 32/LPR rows per pass, all loads issued before the first store (one memory
// round trip).  Rows outside the level are mirrored (REFLECT_101); a word that straddles the left/right border is
// assembled from mirrored bytes -- so windows near the border keep the aligned path.
template <int NROWS, int NWORDS, int LPR>
__device__ __noinline__ void load_words_edge(const uint8_t *__restrict__ P, long long rp, int W, int H, int gx0, int gy0,
                                              uint32_t *dst, int lane)
{   // cold path (window near the border of the level): rolled loops, one word per lane and pass
    constexpr int RPP = 32 / LPR;
    const int k = lane & (LPR - 1), r0 = lane / LPR, gx = gx0 + 4 * k;
    if (k >= NWORDS) return;
    const bool word_ok = gx >= 0 && gx + 3 < W;
    const int x0 = reflect101(gx, W), x1 = reflect101(gx + 1, W), x2 = reflect101(gx + 2, W), x3 = reflect101(gx + 3, W);
#pragma unroll 1
    for (int r = r0; r < NROWS; r += RPP) {
        const uint8_t *row = P + (long long)reflect101(gy0 + r, H) * rp;
        // words inside the row: asynchronous copy like the interior path (the caller commits and waits); only a word
        // that straddles the left / right border is assembled from mirrored bytes
        if (word_ok) cp_async4(dst + r * NWORDS + k, row + gx);
        else dst[r * NWORDS + k] = (uint32_t)row[x0] | ((uint32_t)row[x1] << 8) | ((uint32_t)row[x2] << 16) | ((uint32_t)row[x3] << 24);
    }
}

// ASYNCHRONOUS: the interior path issues cp.async copies (global -> shared, no registers, no stall); the caller commits
// the group and waits (cp_async_wait + __syncwarp) before reading.  The cold border path loads synchronously.
template <int NROWS, int NWORDS, int LPR>
__device__ __forceinline__ void load_words(const uint8_t *__restrict__ P, long long rp, int W, int H, int gx0, int gy0,
                                           uint32_t *dst, int lane)
{
    if (gx0 >= 0 && gy0 >= 0 && gx0 + 4 * NWORDS <= W && gy0 + NROWS <= H) {        // warp-uniform
        constexpr int RPP = 32 / LPR, NP = (NROWS + RPP - 1) / RPP;
        const int k = lane & (LPR - 1), r0 = lane / LPR, gx = gx0 + 4 * k;
        if (k < NWORDS) {
            const uint8_t *src = P + (long long)(gy0 + r0) * rp + gx;
            uint32_t *t = dst + r0 * NWORDS + k;
#pragma unroll
            for (int q = 0; q < NP; q++)
                if (r0 + q * RPP < NROWS) cp_async4(t + q * RPP * NWORDS, src + (long long)q * RPP * rp);
        }
    } else {
        load_words_edge<NROWS, NWORDS, LPR>(P, rp, W, H, gx0, gy0, dst, lane);
    }
}

template <int NROWS, int PITCH>
__device__ __noinline__ void load_bytes_unaligned(const uint8_t *__restrict__ P, long long rp, int W, int H, int gx0, int gy0,
                                                  int ncols, unsigned char *dst, int lane)
{   // cold path: plane not 4-byte aligned.  A lane owns columns (reflect once per column), rows in the inner loop
#pragma unroll 1
    for (int c = lane; c < ncols; c += 32) {
        const uint8_t *col = P + reflect101(gx0 + c, W);
#pragma unroll 1
        for (int r = 0; r < NROWS; r++) dst[r * PITCH + c] = col[(long long)reflect101(gy0 + r, H) * rp];
    }
}

// `next` tile around (tx0, ty0)
template <class T>
__device__ __noinline__ void load_j_tile(const uint8_t *__restrict__ J, long long rp, int W, int H, bool al4,
                                            int tx0, int ty0, unsigned char *tile, int lane)
{
    if (al4) {
        // (TMA form: this is the fallback of a box that is not inside the level -- the rolled border loader does it all)
        if constexpr (T::TMA_) load_words_edge<T::TH, T::TILE_WPR, T::TILE_LPR>(J, rp, W, H, tx0, ty0, reinterpret_cast<uint32_t *>(tile), lane);
        else load_words<T::TH, T::TILE_WPR, T::TILE_LPR>(J, rp, W, H, tx0, ty0, reinterpret_cast<uint32_t *>(tile), lane);
    } else {
        load_bytes_unaligned<T::TH, T::TWC>(J, rp, W, H, tx0, ty0, T::TWC, tile, lane);
    }
}

// template neighbourhood: rows ipy-1 .. ipy+WH+1, columns from the 4-aligned ax <= ipx-1; returns the byte offset of
// column ipx-1 inside a raw row (0..3 on the aligned path, 0 on the unaligned path)
template <class T>
__device__ __noinline__ int load_raw(const uint8_t *__restrict__ I, long long rp, int W, int H, bool al4, int ipx, int ipy,
                                        unsigned char *raw, int lane)
{
    if (al4) {
        const int ax = ((ipx - 1) >> T::XSH) << T::XSH;
        if constexpr (T::TMA_) load_words_edge<T::RH, T::RWW, T::RAW_LPR>(I, rp, W, H, ax, ipy - 1, reinterpret_cast<uint32_t *>(raw), lane);
        else load_words<T::RH, T::RWW, T::RAW_LPR>(I, rp, W, H, ax, ipy - 1, reinterpret_cast<uint32_t *>(raw), lane);
        return (ipx - 1) - ax;
    }
    load_bytes_unaligned<T::RH, T::RWP>(I, rp, W, H, ipx - 1, ipy - 1, T::RW, raw, lane);
    return 0;
}

// Template of a window that touches the border of the level: Scharr on the (WW+1)x(WH+1) sample grid with ZERO outside
// the level, then the Q14 bilinear patch (I*32, Ix, Ix) for every window pixel -> shared memory (oI, oD).
template <class T>
__device__ __noinline__ void template_border(const unsigned char *raw, int ro, short2 *der, unsigned char *tpl, int ipx, int ipy,
                                             int W, int H, int w00, int w01, int w10, int w11, int lane)
{
    // 1. Scharr at the sample grid, four samples of a row per step: the six raw bytes under them come from three aligned
    //    words per row (funnel-shifted), column terms first (V = [3 10 3] down the rows, Dv = bottom - top)
    constexpr int SEG = 4, SPR = (T::DW + SEG - 1) / SEG;
#pragma unroll 1
    for (int sg = lane; sg < T::DH * SPR; sg += 32) {
        const int j = sg / SPR, i0 = (sg - j * SPR) * SEG;
        const int col = ro + i0, sh = (col & 3) * 8;
        int V[6], Dv[6];
        {
            uint32_t lo[3], hi[3];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const uint32_t *q = reinterpret_cast<const uint32_t *>(raw + (j + r) * T::RWP) + (col >> 2);
                const uint32_t a = q[0], b = q[1], c = q[2];
                lo[r] = __funnelshift_r(a, b, sh); hi[r] = __funnelshift_r(b, c, sh);
            }
#pragma unroll
            for (int c = 0; c < 6; c++) {
                const int t = (int)(((c < 4 ? lo[0] : hi[0]) >> (8 * (c & 3))) & 255u);
                const int m = (int)(((c < 4 ? lo[1] : hi[1]) >> (8 * (c & 3))) & 255u);
                const int u = (int)(((c < 4 ? lo[2] : hi[2]) >> (8 * (c & 3))) & 255u);
                V[c] = 3 * (t + u) + 10 * m; Dv[c] = u - t;
            }
        }
        const bool row_ok = ipy + j >= 0 && ipy + j < H;
        short2 *o = der + j * T::DWP + i0;
#pragma unroll
        for (int t = 0; t < SEG; t++) {
            const int qx = ipx + i0 + t;
            const bool ok = row_ok && qx >= 0 && qx < W;
            const int ix = ok ? V[t + 2] - V[t] : 0;
            const int iy = ok ? 3 * (Dv[t] + Dv[t + 2]) + 10 * Dv[t + 1] : 0;
            if (i0 + t < T::DW) o[t] = make_short2((short)ix, (short)iy);
        }
    }
    __syncwarp();
    // 2. the Q14 bilinear patch, run by run like the interior path: I~ by DP2A on the raw bytes, Ix~ / Iy~ from the four
    //    derivative taps; every value is one 16-bit store into the lane's own (zeroed) template words
    constexpr int CH = T::RUN_ <= 8 ? T::RUN_ : 8;
    const int wT = (w00 & 0xffff) | (w01 << 16), wB = (w10 & 0xffff) | (w11 << 16);
    short *mine = reinterpret_cast<short *>(tpl + lane * 16);
    auto put = [&](int word, int half, int v) { mine[(word >> 2) * 256 + (word & 3) * 2 + half] = (short)v; };
#pragma unroll 1
    for (int k = 0; k < T::RPL; k++) {
        const int rr = lane + 32 * k;
        if (rr >= T::NRUNS) break;
        const int y = rr / T::RPR, x = (rr - y * T::RPR) * T::RUN_;
#pragma unroll 1
        for (int c0 = 0; c0 < T::RUN_; c0 += CH) {
            int iv[CH];
            bilinear_run<CH>(raw + (y + 1) * T::RWP, raw + (y + 2) * T::RWP, ro + x + c0 + 1, wT, wB, iv);
            const short2 *d0 = der + y * T::DWP + x + c0, *d1 = d0 + T::DWP;
            short2 p0 = d0[0], p1 = d1[0];
#pragma unroll
            for (int t = 0; t < CH; t++) {
                const short2 n0 = d0[t + 1], n1 = d1[t + 1];
                const int ixv = descale(p0.x * w00 + n0.x * w01 + p1.x * w10 + n1.x * w11, 14);
                const int iyv = descale(p0.y * w00 + n0.y * w01 + p1.y * w10 + n1.y * w11, 14);
                if (c0 + t < T::RUN_) {
                    const int q = k * T::RUN_ + c0 + t;
                    put(q >> 1, q & 1, ixv);
                    put(T::NPAIR + (q >> 1), q & 1, iyv);
                    put(2 * T::NPAIR + (q >> 1), q & 1, iv[t]);
                }
                p0 = n0; p1 = n1;
            }
        }
    }
    __syncwarp();
}

constexpr int LKT_TMA_LEVELS = 4;
struct LKMaps { CUtensorMap prev[LKT_TMA_LEVELS], next[LKT_TMA_LEVELS]; };      // (W, H, B) uint8 planes of the levels
struct LKNoMaps { int unused; };

template <int WW, int WH, int RUN, int MARGIN, int DWP_, int TWC_, int MINB, bool TMA, class MAPS>
__global__ void __launch_bounds__(LKT_WARPS * 32, MINB)
k_lk_track_t(LKLevels lv, int n_levels, int max_count, LKEps eps2, int flags, double min_eig_thr,
             const float *__restrict__ prev_pts, const int *__restrict__ counts, int max_pts,
             float *__restrict__ next_pts, uint8_t *__restrict__ status, float *__restrict__ err,
             int *__restrict__ iters_out, unsigned al4_mask, unsigned ial4_mask, const __grid_constant__ MAPS maps,
             unsigned tma_mask)
{
    using T = LKT<WW, WH, RUN, MARGIN, DWP_, TWC_, TMA>;
    extern __shared__ __align__(128) unsigned char smem_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pt = blockIdx.x * LKT_WARPS + warp, b = blockIdx.y;
    const int n = counts ? counts[b] : max_pts;
    if (pt >= n || pt >= max_pts) return;

    unsigned char *uni = smem_all + (size_t)warp * T::PER_WARP;
    short2 *der = reinterpret_cast<short2 *>(uni);          // Scharr on the sample grid (setup phase)
    unsigned char *tile = uni + T::DER_BYTES;                // `next` tile (prefetched at level entry)
    unsigned char *raw_buf = tile + T::TILE_BYTES;          // raw[l & 1]: template neighbourhood of level l (prefetched one level ahead)
    unsigned char *tpl = raw_buf + 2 * T::RAW_BYTES;        // template patch (layout: LKT)
    // TMA form: mbarriers 0 = tile, 1 / 2 = raw[0] / raw[1]; `pend` bit k: a TMA copy into buffer k is in flight,
    // bit 3 + k: a word-loader (cp.async) copy is; `phase` bit k: parity of barrier k's next completion
    const uint32_t bar0 = TMA ? smem_u32(tpl + T::TPL_BYTES) : 0u;
    unsigned pend = 0, phase = 0;
    if constexpr (TMA) {
        if (lane == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); mbar_init(bar0 + 16, 1); fence_mbar_init(); }
        __syncwarp();
    }
    auto wait_buf = [&](int k) {            // TMA form only
        if (pend & (1u << k)) { mbar_wait(bar0 + 8 * k, (phase >> k) & 1u); phase ^= 1u << k; }
        else if (pend & (8u << k)) cp_async_wait_all();
        pend &= ~(9u << k);
    };

    // this lane's runs: (row, first column) -- compile-time count, runtime validity
    int run_y[T::RPL], run_x[T::RPL], run_off[T::RPL];
    bool run_ok[T::RPL];
#pragma unroll
    for (int k = 0; k < T::RPL; k++) {
        int r = lane + 32 * k;
        run_ok[k] = r < T::NRUNS;
        if (!run_ok[k]) r = 0;
        run_y[k] = r / T::RPR;
        run_x[k] = (r - run_y[k] * T::RPR) * RUN;
        run_off[k] = run_y[k] * T::TWC + run_x[k];
    }
    // half-word masks of the gradient words: zero for runs that do not exist
    uint32_t wmask[T::NPAIR];
#pragma unroll
    for (int i = 0; i < T::NPAIR; i++) {
        const int k0 = (2 * i) / RUN, k1 = (2 * i + 1) / RUN;
        wmask[i] = (run_ok[k0] ? 0xffffu : 0u) | ((2 * i + 1 < T::NPX && run_ok[k1]) ? 0xffff0000u : 0u);
    }

    const long long pidx = (long long)b * max_pts + pt;
    const float ppx = prev_pts[2 * pidx], ppy = prev_pts[2 * pidx + 1];
    const float hx = (WW - 1) * 0.5f, hy = (WH - 1) * 0.5f;
    float npx = 0.f, npy = 0.f;
    if (flags & LKF_OPTFLOW_USE_INITIAL_FLOW) { npx = next_pts[2 * pidx]; npy = next_pts[2 * pidx + 1]; }
    bool st = true;
    float ev = 0.f;
    int iters = 0;
    const float FLT_SCALE = 1.f / (1 << 20);

    // template neighbourhood of level l -> raw[l & 1]; its position depends only on prevPts, so it is fetched one level
    // ahead (cp.async) and never waited for, except at the top level
    auto issue_raw = [&](int l) {
        const float sc = __int_as_float((127 - l) << 23);          // 2^-l, exact
        const float px = __fsub_rn(__fmul_rn(ppx, sc), hx), py = __fsub_rn(__fmul_rn(ppy, sc), hy);
        const int ipx = __float2int_rd(px), ipy = __float2int_rd(py);
        if (!out_of_level(px, py, ipx, ipy, WW, WH, lv.W[l], lv.H[l])) {
            if constexpr (TMA) {
                const int k = 1 + (l & 1);
                if (tma_box<T::RH, T::RWW>(&maps.prev[l], (tma_mask >> l) & 1u, lv.W[l], lv.H[l], ((ipx - 1) >> T::XSH) << T::XSH, ipy - 1, b,
                                           raw_buf + (l & 1) * T::RAW_BYTES, bar0 + 8 * k, lane, !(tma_mask >> 31)))
                    pend |= 1u << k;
                else {
                    load_raw<T>(lv.prev[l] + (long long)b * lv.pip[l], lv.prp[l], lv.W[l], lv.H[l], (ial4_mask >> l) & 1u, ipx, ipy,
                                raw_buf + (l & 1) * T::RAW_BYTES, lane);
                    pend |= 8u << k;
                }
            } else
                load_raw<T>(lv.prev[l] + (long long)b * lv.pip[l], lv.prp[l], lv.W[l], lv.H[l], (ial4_mask >> l) & 1u, ipx, ipy,
                            raw_buf + (l & 1) * T::RAW_BYTES, lane);
        }
        if constexpr (!TMA) cp_async_commit();
    };
    // `next` tile at (tx0, ty0) of level l
    auto issue_tile = [&](int l, const uint8_t *J, long long jrp, int W, int H, bool jal4, int tx0, int ty0) {
        if constexpr (TMA) {
            if (tma_box<T::TH, T::TILE_WPR>(&maps.next[l], (tma_mask >> (8 + l)) & 1u, W, H, tx0, ty0, b, tile, bar0, lane, !(tma_mask >> 31))) pend |= 1u;
            else { load_j_tile<T>(J, jrp, W, H, jal4, tx0, ty0, tile, lane); pend |= 8u; }
        } else {
            load_j_tile<T>(J, jrp, W, H, jal4, tx0, ty0, tile, lane);
            cp_async_commit();
        }
    };
    issue_raw(n_levels - 1);
    bool tile_pending = false;

    for (int l = n_levels - 1; l >= 0; --l) {
        const int W = lv.W[l], H = lv.H[l];
        const uint8_t *J = lv.next[l] + (long long)b * lv.nip[l];
        const long long jrp = lv.nrp[l];
        const bool jal4 = (al4_mask >> l) & 1u, ial4 = (ial4_mask >> l) & 1u;
        unsigned char *raw = raw_buf + (l & 1) * T::RAW_BYTES;
        if (tile_pending) {                                                  // a level left early: its tile copy must land first
            if constexpr (TMA) wait_buf(0); else cp_async_wait<0>();
            tile_pending = false;
        }
        __syncwarp();
        if (l > 0) issue_raw(l - 1); else if constexpr (!TMA) cp_async_commit();
        const float sc = __int_as_float((127 - l) << 23);          // 2^-l, exact
        float px = __fmul_rn(ppx, sc), py = __fmul_rn(ppy, sc);
        float nx, ny;
        if (l == n_levels - 1) {
            if (flags & LKF_OPTFLOW_USE_INITIAL_FLOW) { nx = __fmul_rn(npx, sc); ny = __fmul_rn(npy, sc); }
            else { nx = px; ny = py; }
        } else { nx = __fmul_rn(npx, 2.f); ny = __fmul_rn(npy, 2.f); }
        npx = nx; npy = ny;
        px = __fsub_rn(px, hx); py = __fsub_rn(py, hy);
        const int ipx = __float2int_rd(px), ipy = __float2int_rd(py);
        if (out_of_level(px, py, ipx, ipy, WW, WH, W, H)) {
            if (l == 0) { st = false; ev = 0.f; }
            continue;
        }
        int w00, w01, w10, w11;
        weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy), w00, w01, w10, w11);

        // prefetch the `next` tile around the entry estimate (asynchronous; awaited just before the first sweep)
        int tx0, ty0;
        {
            const float fx = __fsub_rn(nx, hx), fy = __fsub_rn(ny, hy);
            place_tile<T>(__float2int_rd(fx), __float2int_rd(fy), W, H, tx0, ty0);
            issue_tile(l, J, jrp, W, H, jal4, tx0, ty0);
            tile_pending = true;
        }
        // 1. raw (WW+3)x(WH+3) neighbourhood of prev: issued one level ago; the two newest groups (next level's raw,
        //    this level's tile) may still be in flight
        const int ro = ial4 ? (ipx - 1) - (((ipx - 1) >> T::XSH) << T::XSH) : 0;
        if constexpr (TMA) wait_buf(1 + (l & 1)); else cp_async_wait<2>();
        __syncwarp();
        // lane partials of the structure tensor and of c = sum I~ * (Ix~, Iy~): the mismatch vector of an iteration is
        // b = sum J~ * (Ix~, Iy~) - c, so a sweep never touches I~
        int a11 = 0, a12 = 0, a22 = 0, c1 = 0, c2 = 0;
        const bool inside = ipx >= 0 && ipy >= 0 && ipx + WW < W && ipy + WH < H;      // every derivative sample is in the level
        if (inside) {
            // 2a. Bilinear interpolation and the Scharr operator are both linear and cv2 rounds only once, after the
            //     interpolation: Ix~ = (Scharr_x(U) + 2^13) >> 14 with U = sum_k w_k I(. + k) the un-rounded Q14 sums, and
            //     I~ = (U + 2^8) >> 9.  U on the (WW+2)x(WH+2) grid: 2 DP2A per value (4 values per lane step).
            int *U = reinterpret_cast<int *>(der);
            const int wT = (w00 & 0xffff) | (w01 << 16), wB = (w10 & 0xffff) | (w11 << 16);
            for (int sg = lane; sg < T::UH * T::USEG6; sg += 32) {
                const int row = sg / T::USEG6, sx = (sg - row * T::USEG6) * 6;
                int u[6];
                bilinear_run<6, false>(raw + row * T::RWP, raw + (row + 1) * T::RWP, ro + sx, wT, wB, u);
                int2 *o = reinterpret_cast<int2 *>(U + row * T::UP + sx);
                o[0] = make_int2(u[0], u[1]); o[1] = make_int2(u[2], u[3]); o[2] = make_int2(u[4], u[5]);
            }
            __syncwarp();
            // 3a. template from U: window pixel (x, y) is U(x+1, y+1); its 3x3 Scharr support is U(x..x+2, y..y+2).
            //     Packed to int16 pairs (layout: LKT); lanes whose run does not exist store zero gradients (their products
            //     then vanish in every sum, no per-pixel predicate anywhere).
            uint32_t tw[T::NCH_ALL * 4];
#pragma unroll
            for (int i = 0; i < T::NCH_ALL * 4; i++) tw[i] = 0u;
            constexpr int CH = RUN <= 8 ? RUN : 8;
#pragma unroll
            for (int k = 0; k < T::RPL; k++) {
                const int *u0 = U + run_y[k] * T::UP + run_x[k], *u1 = u0 + T::UP, *u2 = u1 + T::UP;
                int s11 = 0, s12 = 0, s22 = 0, t1 = 0, t2 = 0;
#pragma unroll
                for (int c0 = 0; c0 < RUN; c0 += CH) {
                    // column terms first: V = [3 10 3] down the three rows, Dv = bottom - top; then Ix = V(x+2) - V(x) and
                    // Iy = 3 Dv(x) + 10 Dv(x+1) + 3 Dv(x+2)
                    int V[CH + 2], Dv[CH + 2], ctr[CH + 2];
#pragma unroll
                    for (int i = 0; i < CH + 2; i++)
                        if (c0 + i < RUN + 2) {
                            const int a = u0[c0 + i], m = u1[c0 + i], c = u2[c0 + i];
                            V[i] = 3 * (a + c) + 10 * m; Dv[i] = c - a; ctr[i] = m;
                        }
#pragma unroll
                    for (int i = 0; i < CH; i++) {
                        if (c0 + i < RUN) {
                            const int q = k * RUN + c0 + i;
                            const int ixv = (V[i + 2] - V[i] + (1 << 13)) >> 14;
                            const int iyv = (3 * (Dv[i] + Dv[i + 2]) + 10 * Dv[i + 1] + (1 << 13)) >> 14;
                            const int iv = (ctr[i + 1] + (1 << 8)) >> 9;
                            if (q & 1) {
                                tw[q >> 1] |= (uint32_t)ixv << 16;
                                tw[T::NPAIR + (q >> 1)] |= (uint32_t)iyv << 16;
                                tw[2 * T::NPAIR + (q >> 1)] |= (uint32_t)iv << 16;
                            } else {
                                tw[q >> 1] = (uint32_t)ixv & 0xffffu;
                                tw[T::NPAIR + (q >> 1)] = (uint32_t)iyv & 0xffffu;
                                tw[2 * T::NPAIR + (q >> 1)] = (uint32_t)iv;
                            }
                            s11 += ixv * ixv; s12 += ixv * iyv; s22 += iyv * iyv;
                            t1 += iv * ixv; t2 += iv * iyv;
                        }
                    }
                }
                if (run_ok[k]) { a11 += s11; a12 += s12; a22 += s22; c1 += t1; c2 += t2; }
            }
#pragma unroll
            for (int i = 0; i < T::NW_D; i++) tw[i] &= wmask[i < T::NPAIR ? i : i - T::NPAIR];
            uint4 *slot = reinterpret_cast<uint4 *>(tpl) + lane;
#pragma unroll
            for (int c = 0; c < T::NCH_ALL; c++) slot[c * 32] = make_uint4(tw[4 * c], tw[4 * c + 1], tw[4 * c + 2], tw[4 * c + 3]);
        } else {
            // 2b/3b. window touches the border of the level (cold): derivative samples outside the level are ZERO (cv2's
            //        zero-bordered derivative image) while image samples are mirrored, so the linearity shortcut does not
            //        hold; a rolled, non-inlined routine writes the template and the lanes re-read their words.
            for (int i = lane; i < T::TPL_BYTES / 16; i += 32) reinterpret_cast<uint4 *>(tpl)[i] = make_uint4(0u, 0u, 0u, 0u);
            __syncwarp();
            template_border<T>(raw, ro, der, tpl, ipx, ipy, W, H, w00, w01, w10, w11, lane);
            const uint32_t *mine = reinterpret_cast<const uint32_t *>(tpl + lane * 16);
#pragma unroll 1
            for (int i = 0; i < T::NPAIR; i++) {
                const uint32_t wx = mine[(i >> 2) * 128 + (i & 3)], wy = mine[((T::NPAIR + i) >> 2) * 128 + ((T::NPAIR + i) & 3)],
                               wi = mine[((2 * T::NPAIR + i) >> 2) * 128 + ((2 * T::NPAIR + i) & 3)];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int ixv = (int)(short)(wx >> (16 * h)), iyv = (int)(short)(wy >> (16 * h)), iv = (int)((wi >> (16 * h)) & 0xffffu);
                    a11 += ixv * ixv; a12 += ixv * iyv; a22 += iyv * iyv; c1 += iv * ixv; c2 += iv * iyv;
                }
            }
        }
        const long long sc1 = warp_sum(c1), sc2 = warp_sum(c2);
        const float A11 = __fmul_rn((float)warp_sum(a11), FLT_SCALE);
        const float A12 = __fmul_rn((float)warp_sum(a12), FLT_SCALE);
        const float A22 = __fmul_rn((float)warp_sum(a22), FLT_SCALE);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dd = __fsub_rn(A11, A22);
        const float disc = __fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12));
        const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(disc)), (float)(2 * WW * WH));
        if (flags & LKF_OPTFLOW_LK_GET_MIN_EIGENVALS) ev = minEig;
        if ((double)minEig < min_eig_thr || D < 1.1920929e-07f) {
            if (l == 0) st = false;
            continue;
        }
        D = __fdiv_rn(1.f, D);
        nx = __fsub_rn(nx, hx); ny = __fsub_rn(ny, hy);
        float pdx = 0.f, pdy = 0.f;
        if constexpr (TMA) wait_buf(0); else cp_async_wait<0>();          // the `next` tile (and, long since, the next level's raw)
        tile_pending = false;
        __syncwarp();

        // One sweep of the window at integer origin (ix, iy) of `next` (through the smem tile) with the current weights.
        // SweepB: o0..o3 = warp sums of J~ * Ix~ and J~ * Iy~, each split by the byte of J~ (J~ is a 13-bit value: the two
        // bytes of two neighbouring pixels are packed by one PRMT and multiplied with the int16 gradient pairs by DP2A;
        // low-byte and high-byte sums fit one REDUX each, sum = o_lo + 256 * o_hi).  SweepErr: o0 = warp sum of |J~ - I~|.
        auto sweep = [&](int ix, int iy, auto mode, int &o0, int &o1, int &o2, int &o3) {
            constexpr bool ERR = decltype(mode)::value;
            int ox = ix - tx0, oy = iy - ty0;
            if (ox < 0 || oy < 0 || ox + WW + 1 > T::TWC || oy + WH + 1 > T::TH) {
                __syncwarp();
                place_tile<T>(ix, iy, W, H, tx0, ty0);
                issue_tile(l, J, jrp, W, H, jal4, tx0, ty0);
                if constexpr (TMA) wait_buf(0); else cp_async_wait<0>();
                __syncwarp();
                ox = ix - tx0; oy = iy - ty0;
            }
            const int wT = (w00 & 0xffff) | (w01 << 16), wB = (w10 & 0xffff) | (w11 << 16);
            const int toff = oy * T::TWC + ox;
            const uint4 *slot = reinterpret_cast<const uint4 *>(tpl) + lane;
            int v[T::NPX + 1];
            v[T::NPX] = 0;
#pragma unroll
            for (int k = 0; k < T::RPL; k++) {
                int vv[RUN];
                bilinear_run<RUN>(tile, tile + T::TWC, toff + run_off[k], wT, wB, vv);
#pragma unroll
                for (int t = 0; t < RUN; t++) v[k * RUN + t] = vv[t];
            }
            if constexpr (!ERR) {
                uint32_t tw[T::NCH_D * 4];
#pragma unroll
                for (int c = 0; c < T::NCH_D; c++) {
                    const uint4 q = slot[c * 32];
                    tw[4 * c] = q.x; tw[4 * c + 1] = q.y; tw[4 * c + 2] = q.z; tw[4 * c + 3] = q.w;
                }
                int xl = 0, xh = 0, yl = 0, yh = 0;
#pragma unroll
                for (int i = 0; i < T::NPAIR; i++) {
                    const uint32_t jj = prmt((uint32_t)v[2 * i], (uint32_t)v[2 * i + 1], 0x5140u);   // bytes (lo0, lo1, hi0, hi1)
                    xl = dp2a_lo((int)tw[i], jj, xl); xh = dp2a_hi((int)tw[i], jj, xh);
                    yl = dp2a_lo((int)tw[T::NPAIR + i], jj, yl); yh = dp2a_hi((int)tw[T::NPAIR + i], jj, yh);
                }
                o0 = __reduce_add_sync(0xffffffffu, xl); o1 = __reduce_add_sync(0xffffffffu, xh);
                o2 = __reduce_add_sync(0xffffffffu, yl); o3 = __reduce_add_sync(0xffffffffu, yh);
            } else {
                constexpr int C0 = (2 * T::NPAIR) / 4, C1 = (3 * T::NPAIR + 3) / 4;      // chunks holding the I~ words
                uint32_t tw[(C1 - C0) * 4];
#pragma unroll
                for (int c = C0; c < C1; c++) {
                    const uint4 q = slot[c * 32];
                    tw[4 * (c - C0)] = q.x; tw[4 * (c - C0) + 1] = q.y; tw[4 * (c - C0) + 2] = q.z; tw[4 * (c - C0) + 3] = q.w;
                }
                int se = 0;
#pragma unroll
                for (int q = 0; q < T::NPX; q++) {
                    const uint32_t w = tw[2 * T::NPAIR + (q >> 1) - 4 * C0];
                    const int iv = (q & 1) ? (int)(w >> 16) : (int)(w & 0xffffu);
                    se += run_ok[q / RUN] ? abs(v[q] - iv) : 0;
                }
                o0 = __reduce_add_sync(0xffffffffu, se);
            }
        };

        for (int j = 0; j < max_count; j++) {
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if (out_of_level(nx, ny, inx, iny, WW, WH, W, H)) {
                if (l == 0) st = false;
                break;
            }
            weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), w00, w01, w10, w11);
            int s0, s1, s2, s3;
            // (runs that do not exist hold a zero template: their terms vanish)
            sweep(inx, iny, SweepB{}, s0, s1, s2, s3);
            iters++;
            const long long b1 = (long long)s0 + ((long long)s1 << 8) - sc1, b2 = (long long)s2 + ((long long)s3 << 8) - sc2;
            const float fb1 = __fmul_rn((float)b1, FLT_SCALE);
            const float fb2 = __fmul_rn((float)b2, FLT_SCALE);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, fb2), __fmul_rn(A22, fb1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, fb1), __fmul_rn(A11, fb2)), D);
            nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
            npx = __fadd_rn(nx, hx); npy = __fadd_rn(ny, hy);
            if (lk_step_below_eps(dx, dy, eps2)) break;
            if (j > 0 && lk_below_001(__fadd_rn(dx, pdx)) && lk_below_001(__fadd_rn(dy, pdy))) {
                npx = __fsub_rn(npx, __fmul_rn(dx, 0.5f));
                npy = __fsub_rn(npy, __fmul_rn(dy, 0.5f));
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (st && l == 0 && !(flags & LKF_OPTFLOW_LK_GET_MIN_EIGENVALS)) {
            const float ex = __fsub_rn(npx, hx), ey = __fsub_rn(npy, hy);
            const int iex = __float2int_rd(ex), iey = __float2int_rd(ey);
            if (out_of_level(ex, ey, iex, iey, WW, WH, W, H)) { st = false; continue; }
            weights(__fsub_rn(ex, (float)iex), __fsub_rn(ey, (float)iey), w00, w01, w10, w11);
            int se, u1, u2, u3;
            sweep(iex, iey, SweepErr{}, se, u1, u2, u3);
            ev = __fdiv_rn((float)se, (float)(32 * WW * WH));
        }
    }
    if constexpr (TMA) {          // no bulk copy may still be writing this CTA's shared memory when it exits
#pragma unroll
        for (int k = 0; k < 3; k++)
            if (pend & (1u << k)) wait_buf(k);
    }
    if (lane == 0) {
        next_pts[2 * pidx] = npx;
        next_pts[2 * pidx + 1] = npy;
        status[pidx] = st ? 1 : 0;
        err[pidx] = st ? ev : 0.f;
        if (iters_out) iters_out[pidx] = iters;
    }
}

// (W, H, B) uint8 tensor map of one pyramid level with box (bw, bh, 1); false if the level cannot be described
static bool lk_level_map(const lkf_image *im, int bw, int bh, CUtensorMap *map)
{
    lkf_encode_tiled_fn enc = encode_tiled();
    if (!enc || im->width < bw || im->height < bh || !aligned_to(im, 16)) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)im->width, (cuuint64_t)im->height, (cuuint64_t)im->batch};
    const cuuint64_t gstr[2] = {(cuuint64_t)im->row_pitch, (cuuint64_t)(im->batch > 1 ? im->image_pitch : im->row_pitch * im->height)};
    if (gstr[1] % 16) return false;
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, im->data, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// TWC_TMA: tile pitch of the TMA form (0: the window has none).  The TMA form (LKF_LK_TMA=1; OFF by default) needs level 0 of
// both pyramids describable by a tensor map (16-byte aligned planes and pitches).  Measured on the C2 stage (128 000 features,
// 21x21, profiles/r02_lk_tma.md): 1000 us against 741 us for the cp.async form, bit-identical outputs.  The bulk copies
// themselves are cheap; the kernel grows from 2520 to 3184 SASS instructions (40 -> 51 KB: 16-byte aligned box origins need
// 48-byte pitches and wider realignment, barrier bookkeeping, both loader families) and the warps' `no_instruction` stall
// share -- already the third largest of the cp.async form -- quadruples: the kernel is instruction-cache bound first.
template <int WW, int WH, int RUN, int MARGIN, int DWP_, int TWC_, int MINB, int TWC_TMA>
int launch_lk_t(const LKLevels &lv, int n_levels, int max_count, double eps, const lkf_lk_params *p, const float *prev_pts,
                const int32_t *pt_counts, int max_pts, float *next_pts, uint8_t *status, float *err, int32_t *iters,
                unsigned al4_mask, unsigned ial4_mask, int B, cudaStream_t s, const lkf_image *prev_pyr, const lkf_image *next_pyr)
{
    dim3 grid((max_pts + LKT_WARPS - 1) / LKT_WARPS, B);
    if constexpr (TWC_TMA > 0) {
        static const int use_tma = getenv("LKF_LK_TMA") ? atoi(getenv("LKF_LK_TMA")) : 0;
        if (use_tma && n_levels <= LKT_TMA_LEVELS) {
            using T = LKT<WW, WH, RUN, MARGIN, DWP_, TWC_TMA, true>;
            LKMaps maps;
            unsigned tma_mask = 0;
            for (int l = 0; l < n_levels; l++) {
                if (lk_level_map(&prev_pyr[l], T::RWP, T::RH, &maps.prev[l])) tma_mask |= 1u << l;
                else memset(&maps.prev[l], 0, sizeof(CUtensorMap));
                if (lk_level_map(&next_pyr[l], T::TWC, T::TH, &maps.next[l])) tma_mask |= 1u << (8 + l);
                else memset(&maps.next[l], 0, sizeof(CUtensorMap));
            }
            for (int l = n_levels; l < LKT_TMA_LEVELS; l++) { memset(&maps.prev[l], 0, sizeof(CUtensorMap)); memset(&maps.next[l], 0, sizeof(CUtensorMap)); }
            static const int dbg = getenv("LKF_LK_TMA_DBG") ? atoi(getenv("LKF_LK_TMA_DBG")) : 0;      // 1: no raw boxes, 2: no tile boxes
            const unsigned full_mask = tma_mask;
            if (dbg & 1) tma_mask &= ~0xffu;
            if (dbg & 2) tma_mask &= ~0xff00u;
            if (dbg & 4) tma_mask |= 0x80000000u;      // experiment: no proxy fence before the bulk copies
            if ((full_mask & 1u) && (full_mask & 0x100u)) {
                constexpr size_t smem = (size_t)T::PER_WARP * LKT_WARPS;
                auto kern = k_lk_track_t<WW, WH, RUN, MARGIN, DWP_, TWC_TMA, MINB, true, LKMaps>;
                LKF_FUNC_ATTRS({
                    LKF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    LKF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
                });
                kern<<<grid, LKT_WARPS * 32, smem, s>>>(lv, n_levels, max_count, make_eps(eps), p->flags, p->min_eig_threshold, prev_pts, pt_counts,
                                                        max_pts, next_pts, status, err, iters, al4_mask, ial4_mask, maps, tma_mask);
                return LKF_OK;
            }
        }
    }
    using T = LKT<WW, WH, RUN, MARGIN, DWP_, TWC_>;
    constexpr size_t smem = (size_t)T::PER_WARP * LKT_WARPS;
    auto kern = k_lk_track_t<WW, WH, RUN, MARGIN, DWP_, TWC_, MINB, false, LKNoMaps>;
    LKF_FUNC_ATTRS({
        LKF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LKF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    });
    kern<<<grid, LKT_WARPS * 32, smem, s>>>(lv, n_levels, max_count, make_eps(eps), p->flags, p->min_eig_threshold, prev_pts, pt_counts, max_pts,
                                            next_pts, status, err, iters, al4_mask, ial4_mask, LKNoMaps{0}, 0u);
    return LKF_OK;
}

// ---------------------------------------------------------------------------------------------
// Small windows (3x3, 5x5, 7x7; config C4: dense per-pixel flow with a 5x5 window): ONE THREAD per feature.
// A 25-pixel window cannot occupy a warp, and a dense grid of features makes neighbouring threads read
// overlapping neighbourhoods, so every load is served by L1.  The template lives in registers (packed 16-bit),
// all window sums are exact int32 (AREA <= 49), everything is unrolled at compile time.  Same arithmetic and
// control flow as the warp kernels: results are bit-identical to them.
// ---------------------------------------------------------------------------------------------

template <int WW, int WH>
__global__ void __launch_bounds__(128)
k_lk_track_small(LKLevels lv, int n_levels, int max_count, LKEps eps2, int flags, double min_eig_thr,
                 const float *__restrict__ prev_pts, const int *__restrict__ counts, int max_pts,
                 float *__restrict__ next_pts, uint8_t *__restrict__ status, float *__restrict__ err,
                 int *__restrict__ iters_out, unsigned al4_mask, unsigned ial4_mask, const int *__restrict__ skip_flag)
{
    constexpr int AREA = WW * WH, RW = WW + 3, NWR = (RW + 3) / 4, NWJ = (WW + 1 + 3) / 4;
    static_assert(AREA <= 49, "exact int32 window sums");
    if (skip_flag && *skip_flag) return;                  // prevPts is the raster grid: k_lk_dense does this call
    const int pt = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    const int n = counts ? counts[b] : max_pts;
    if (pt >= n || pt >= max_pts) return;
    const long long pidx = (long long)b * max_pts + pt;
    const float ppx = prev_pts[2 * pidx], ppy = prev_pts[2 * pidx + 1];
    const float hx = (WW - 1) * 0.5f, hy = (WH - 1) * 0.5f;
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
        const bool jal4 = (al4_mask >> l) & 1u, ial4 = (ial4_mask >> l) & 1u;
        const float sc = __int_as_float((127 - l) << 23);          // 2^-l, exact
        float px = __fmul_rn(ppx, sc), py = __fmul_rn(ppy, sc);
        float nx, ny;
        if (l == n_levels - 1) {
            if (flags & LKF_OPTFLOW_USE_INITIAL_FLOW) { nx = __fmul_rn(npx, sc); ny = __fmul_rn(npy, sc); }
            else { nx = px; ny = py; }
        } else { nx = __fmul_rn(npx, 2.f); ny = __fmul_rn(npy, 2.f); }
        npx = nx; npy = ny;
        px = __fsub_rn(px, hx); py = __fsub_rn(py, hy);
        const int ipx = __float2int_rd(px), ipy = __float2int_rd(py);
        if (out_of_level(px, py, ipx, ipy, WW, WH, W, H)) {
            if (l == 0) { st = false; ev = 0.f; }
            continue;
        }
        int w00, w01, w10, w11;
        weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy), w00, w01, w10, w11);
        int wT = (w00 & 0xffff) | (w01 << 16), wB = (w10 & 0xffff) | (w11 << 16);

        // template: raw rows r = 0..WH+2 (image rows ipy-1+r, columns ipx-1 ..) stream through registers; the horizontal
        // Scharr terms d (difference) and s ([3 10 3] smoothing) of the last three rows give derivative row j = r-2
        // (zero outside the level), derivative rows y, y+1 and raw rows y+1, y+2 give template row y.
        unsigned tI[(AREA + 1) / 2], tD[AREA];
        int a11 = 0, a12 = 0, a22 = 0;
        {
            uint32_t raw[3][NWR];
            int d[3][WW + 1], s[3][WW + 1];
            int dIx[2][WW + 1], dIy[2][WW + 1];
#pragma unroll
            for (int r = 0; r < WH + 3; r++) {
                load_row_bytes<RW>(I, irp, W, H, ial4, ipx - 1, ipy - 1 + r, raw[r % 3]);
#pragma unroll
                for (int i = 0; i <= WW; i++) {
                    const int a = byte_of<NWR>(raw[r % 3], i), m = byte_of<NWR>(raw[r % 3], i + 1), c = byte_of<NWR>(raw[r % 3], i + 2);
                    d[r % 3][i] = c - a;
                    s[r % 3][i] = 3 * (a + c) + 10 * m;
                }
                if (r >= 2) {
                    const int j = r - 2;
                    const bool row_ok = ipy + j >= 0 && ipy + j < H;
#pragma unroll
                    for (int i = 0; i <= WW; i++) {
                        const bool ok = row_ok && ipx + i >= 0 && ipx + i < W;
                        dIx[j & 1][i] = ok ? 3 * (d[j % 3][i] + d[(j + 2) % 3][i]) + 10 * d[(j + 1) % 3][i] : 0;
                        dIy[j & 1][i] = ok ? s[(j + 2) % 3][i] - s[j % 3][i] : 0;
                    }
                    if (j >= 1) {
                        const int y = j - 1;
                        int iv[WW];
                        bilinear_packed<WW, NWR>(raw[(y + 1) % 3], raw[(y + 2) % 3], 1, wT, wB, iv);
#pragma unroll
                        for (int x = 0; x < WW; x++) {
                            const int ivv = iv[x];
                            const int ixv = descale(dIx[y & 1][x] * w00 + dIx[y & 1][x + 1] * w01 + dIx[(y + 1) & 1][x] * w10 +
                                                    dIx[(y + 1) & 1][x + 1] * w11, 14);
                            const int iyv = descale(dIy[y & 1][x] * w00 + dIy[y & 1][x + 1] * w01 + dIy[(y + 1) & 1][x] * w10 +
                                                    dIy[(y + 1) & 1][x + 1] * w11, 14);
                            const int idx = y * WW + x;
                            if (idx & 1) tI[idx / 2] |= (unsigned)ivv << 16; else tI[idx / 2] = (unsigned)ivv;
                            tD[idx] = ((unsigned)ixv & 0xffffu) | ((unsigned)iyv << 16);
                            a11 += ixv * ixv; a12 += ixv * iyv; a22 += iyv * iyv;
                        }
                    }
                }
            }
        }
        const float A11 = __fmul_rn((float)a11, FLT_SCALE), A12 = __fmul_rn((float)a12, FLT_SCALE), A22 = __fmul_rn((float)a22, FLT_SCALE);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dd = __fsub_rn(A11, A22);
        const float disc = __fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12));
        const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(disc)), (float)(2 * WW * WH));
        if (flags & LKF_OPTFLOW_LK_GET_MIN_EIGENVALS) ev = minEig;
        if ((double)minEig < min_eig_thr || D < 1.1920929e-07f) {
            if (l == 0) st = false;
            continue;
        }
        D = __fdiv_rn(1.f, D);
        nx = __fsub_rn(nx, hx); ny = __fsub_rn(ny, hy);
        float pdx = 0.f, pdy = 0.f;

        // J - I~ over the window at integer origin (ix, iy) with the current weights; f(idx, diff)
        // the (WW+1)x(WH+1) bytes of `next` under the window stay in registers while the integer origin does not move
        // (after the first step of a level most iterations only change the sub-pixel weights)
        uint32_t jrows[WH + 1][NWJ];
        int jx = INT_MIN, jy = INT_MIN;
        auto sweep = [&](int ix, int iy, auto f) {
            if (ix != jx || iy != jy) {
#pragma unroll
                for (int y = 0; y <= WH; y++) load_row_bytes<WW + 1>(J, jrp, W, H, jal4, ix, iy + y, jrows[y]);
                jx = ix; jy = iy;
            }
#pragma unroll
            for (int y = 0; y < WH; y++) {
                int v[WW];
                bilinear_packed<WW, NWJ>(jrows[y], jrows[y + 1], 0, wT, wB, v);
#pragma unroll
                for (int x = 0; x < WW; x++) {
                    const int idx = y * WW + x;
                    const int iv = (idx & 1) ? (int)(tI[idx / 2] >> 16) : (int)(tI[idx / 2] & 0xffffu);
                    f(idx, v[x] - iv);
                }
            }
        };

        for (int j = 0; j < max_count; j++) {
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if (out_of_level(nx, ny, inx, iny, WW, WH, W, H)) {
                if (l == 0) st = false;
                break;
            }
            weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), w00, w01, w10, w11);
            wT = (w00 & 0xffff) | (w01 << 16); wB = (w10 & 0xffff) | (w11 << 16);
            int b1 = 0, b2 = 0;
            sweep(inx, iny, [&](int idx, int diff) { b1 += diff * (int)(short)(tD[idx] & 0xffffu); b2 += diff * ((int)tD[idx] >> 16); });
            iters++;
            const float fb1 = __fmul_rn((float)b1, FLT_SCALE), fb2 = __fmul_rn((float)b2, FLT_SCALE);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, fb2), __fmul_rn(A22, fb1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, fb1), __fmul_rn(A11, fb2)), D);
            nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
            npx = __fadd_rn(nx, hx); npy = __fadd_rn(ny, hy);
            if (lk_step_below_eps(dx, dy, eps2)) break;
            if (j > 0 && lk_below_001(__fadd_rn(dx, pdx)) && lk_below_001(__fadd_rn(dy, pdy))) {
                npx = __fsub_rn(npx, __fmul_rn(dx, 0.5f));
                npy = __fsub_rn(npy, __fmul_rn(dy, 0.5f));
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (st && l == 0 && !(flags & LKF_OPTFLOW_LK_GET_MIN_EIGENVALS)) {
            const float ex = __fsub_rn(npx, hx), ey = __fsub_rn(npy, hy);
            const int iex = __float2int_rd(ex), iey = __float2int_rd(ey);
            if (out_of_level(ex, ey, iex, iey, WW, WH, W, H)) { st = false; continue; }
            weights(__fsub_rn(ex, (float)iex), __fsub_rn(ey, (float)iey), w00, w01, w10, w11);
            wT = (w00 & 0xffff) | (w01 << 16); wB = (w10 & 0xffff) | (w11 << 16);
            int se = 0;
            sweep(iex, iey, [&](int, int diff) { se += abs(diff); });
            ev = __fdiv_rn((float)se, (float)(32 * WW * WH));
        }
    }
    next_pts[2 * pidx] = npx;
    next_pts[2 * pidx + 1] = npy;
    status[pidx] = st ? 1 : 0;
    err[pidx] = st ? ev : 0.f;
    if (iters_out) iters_out[pidx] = iters;
}

template <int WW, int WH>
int launch_lk_small(const LKLevels &lv, int n_levels, int max_count, double eps, const lkf_lk_params *p, const float *prev_pts,
                    const int32_t *pt_counts, int max_pts, float *next_pts, uint8_t *status, float *err, int32_t *iters,
                    unsigned al4_mask, unsigned ial4_mask, int B, cudaStream_t s, const int *skip_flag)
{
    dim3 grid((max_pts + 127) / 128, B);
    k_lk_track_small<WW, WH><<<grid, 128, 0, s>>>(lv, n_levels, max_count, make_eps(eps), p->flags, p->min_eig_threshold, prev_pts, pt_counts,
                                                  max_pts, next_pts, status, err, iters, al4_mask, ial4_mask, skip_flag);
    return LKF_OK;
}

size_t lk_per_warp_bytes(int w, int h)
{
    size_t area = (size_t)w * h;
    size_t bytes = ((area * 2 + 3) & ~(size_t)3) + area * 4 + (size_t)(w + 1) * (h + 1) * 4 + (size_t)(w + 3) * (h + 3);
    return (bytes + 15) & ~(size_t)15;
}

}  // namespace

// lk_dense.cu
bool lk_dense_supported(const lkf_lk_params *p, int n_levels);
int lk_dense_grid_check(const float *prev_pts, int B, int W, int H, int *flag, cudaStream_t s);
int lk_dense_launch(const lkf_lk_params *p, int n_levels, const lkf_image *prev_pyr, const lkf_image *next_pyr, int max_count, double eps,
                    float *next_pts, uint8_t *status, float *err, int32_t *iters, const int *grid_flag, cudaStream_t s);

static void lk_resolve_criteria(const lkf_lk_params *p, int *max_count, double *eps2)
{   // cv2: missing COUNT -> 30 iterations, missing EPS -> 0.01; clamps as in lkpyramid.cpp; epsilon is squared
    int mc = p->max_count;
    double eps = p->epsilon;
    if (!(p->crit_type & LKF_TERM_CRITERIA_COUNT)) mc = 30;
    else mc = mc < 0 ? 0 : (mc > 100 ? 100 : mc);
    if (!(p->crit_type & LKF_TERM_CRITERIA_EPS)) eps = 0.01;
    else eps = eps < 0 ? 0 : (eps > 10 ? 10 : eps);
    *max_count = mc; *eps2 = eps * eps;
}

static int lk_check_pyramids(const lkf_image *prev_pyr, const lkf_image *next_pyr, int n_levels)
{
    const int B = prev_pyr[0].batch;
    for (int l = 0; l < n_levels; l++) {
        if (!img_ok(&prev_pyr[l], 1, 1) || !img_ok(&next_pyr[l], 1, 1)) return LKF_ERR_ARG;
        if (prev_pyr[l].batch != B || next_pyr[l].batch != B || prev_pyr[l].height != next_pyr[l].height ||
            prev_pyr[l].width != next_pyr[l].width)
            return LKF_ERR_ARG;
        if (l > 0 && (prev_pyr[l].height != (prev_pyr[l - 1].height + 1) / 2 ||
                      prev_pyr[l].width != (prev_pyr[l - 1].width + 1) / 2))
            return LKF_ERR_ARG;
    }
    return LKF_OK;
}

// a7 for prevPts = every pixel (x, y) of level 0 in raster order (config C4); outputs have H*W entries per image
LKF_EXPORT int lkf_pyrlk_dense(lkf_ctx *ctx, const lkf_lk_params *p, int n_levels, const lkf_image *prev_pyr,
                               const lkf_image *next_pyr, float *next_pts, uint8_t *status, float *err, int32_t *iters,
                               lkf_stream stream)
{
    if (!p || !prev_pyr || !next_pyr || n_levels < 1 || n_levels > LKF_MAX_LEVELS) return LKF_ERR_ARG;
    if (p->win_w <= 2 || p->win_h <= 2 || p->max_level < 0 || !next_pts || !status || !err) return LKF_ERR_ARG;
    int rc = lk_check_pyramids(prev_pyr, next_pyr, n_levels);
    if (rc != LKF_OK) return rc;
    if (!lk_dense_supported(p, n_levels)) return LKF_ERR_UNSUPPORTED;
    int max_count; double eps;
    lk_resolve_criteria(p, &max_count, &eps);
    rc = lk_dense_launch(p, n_levels, prev_pyr, next_pyr, max_count, eps, next_pts, status, err, iters, nullptr, (cudaStream_t)stream);
    if (rc != LKF_OK) return rc;
    LKF_LAUNCH_CHECK(ctx);
    return LKF_OK;
}

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
    {
        const int rc0 = lk_check_pyramids(prev_pyr, next_pyr, n_levels);
        if (rc0 != LKF_OK) return rc0;
    }
    for (int l = 0; l < n_levels; l++) {
        lv.prev[l] = (const uint8_t *)prev_pyr[l].data; lv.next[l] = (const uint8_t *)next_pyr[l].data;
        lv.prp[l] = prev_pyr[l].row_pitch; lv.pip[l] = prev_pyr[l].image_pitch;
        lv.nrp[l] = next_pyr[l].row_pitch; lv.nip[l] = next_pyr[l].image_pitch;
        lv.H[l] = prev_pyr[l].height; lv.W[l] = prev_pyr[l].width;
    }
    int max_count;
    double eps;
    lk_resolve_criteria(p, &max_count, &eps);
    const size_t pw = lk_per_warp_bytes(p->win_w, p->win_h);
    const size_t smem = pw * LK_WARPS;
    if (smem > 227 * 1024) return LKF_ERR_UNSUPPORTED;
    const int area = p->win_w * p->win_h;
    dim3 grid((max_pts + LK_WARPS - 1) / LK_WARPS, B);
    cudaStream_t s = (cudaStream_t)stream;
    unsigned al4_mask = 0, ial4_mask = 0;
    for (int l = 0; l < n_levels; l++) {
        if (aligned_to(&next_pyr[l], 4)) al4_mask |= 1u << l;
        if (aligned_to(&prev_pyr[l], 4)) ial4_mask |= 1u << l;
    }
    // One point per pixel of level 0 and a small window: prevPts may be the raster grid (dense flow, config C4).  A device
    // flag decides without a host round trip: k_lk_dense runs when the points ARE the grid, k_lk_track_small otherwise
    // (the other one returns at once), so the call stays capturable.
    const int *grid_flag = nullptr;
    if (ctx && !pt_counts && (long long)max_pts == (long long)lv.W[0] * lv.H[0] && lk_dense_supported(p, n_levels)) {
        void *fl;
        int rcf = lkf_scratch(ctx, 5, 256, &fl, s);
        if (rcf != LKF_OK) return rcf;
        if ((rcf = lk_dense_grid_check(prev_pts, B, lv.W[0], lv.H[0], (int *)fl, s)) != LKF_OK) return rcf;
        if ((rcf = lk_dense_launch(p, n_levels, prev_pyr, next_pyr, max_count, eps, next_pts, status, err, iters, (const int *)fl, s)) != LKF_OK)
            return rcf;
        lkf_count_launch(ctx); lkf_count_launch(ctx);
        grid_flag = (const int *)fl;
    }
    int rc = 1;
    if (p->win_w == 21 && p->win_h == 21)
        rc = launch_lk_t<21, 21, 7, 1, 22, 28, 28, 48>(lv, n_levels, max_count, eps, p, prev_pts, pt_counts, max_pts, next_pts, status, err, iters, al4_mask, ial4_mask, B, s, prev_pyr, next_pyr);
    else if (p->win_w == 15 && p->win_h == 15)
        rc = launch_lk_t<15, 15, 5, 1, 17, 24, 24, 48>(lv, n_levels, max_count, eps, p, prev_pts, pt_counts, max_pts, next_pts, status, err, iters, al4_mask, ial4_mask, B, s, prev_pyr, next_pyr);
    else if (p->win_w == 31 && p->win_h == 31)
        rc = launch_lk_t<31, 31, 31, 1, 33, 44, 16, 0>(lv, n_levels, max_count, eps, p, prev_pts, pt_counts, max_pts, next_pts, status, err, iters, al4_mask, ial4_mask, B, s, prev_pyr, next_pyr);
    else if (p->win_w == 5 && p->win_h == 5)
        rc = launch_lk_small<5, 5>(lv, n_levels, max_count, eps, p, prev_pts, pt_counts, max_pts, next_pts, status, err, iters, al4_mask, ial4_mask, B, s, grid_flag);
    else if (p->win_w == 3 && p->win_h == 3)
        rc = launch_lk_small<3, 3>(lv, n_levels, max_count, eps, p, prev_pts, pt_counts, max_pts, next_pts, status, err, iters, al4_mask, ial4_mask, B, s, grid_flag);
    else if (p->win_w == 7 && p->win_h == 7)
        rc = launch_lk_small<7, 7>(lv, n_levels, max_count, eps, p, prev_pts, pt_counts, max_pts, next_pts, status, err, iters, al4_mask, ial4_mask, B, s, grid_flag);
    if (rc < 0) return rc;
    if (rc == LKF_OK) { LKF_LAUNCH_CHECK(ctx); return LKF_OK; }
    if (area <= 2048) {      // int32 lane partials cannot overflow (<= 64 px/lane * 8160*4080)
        LKF_FUNC_ATTRS(LKF_CUDA(cudaFuncSetAttribute(k_lk_track<int>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)));
        k_lk_track<int><<<grid, LK_WARPS * 32, smem, s>>>(lv, n_levels, p->win_w, p->win_h, max_count, make_eps(eps), p->flags,
                                                          p->min_eig_threshold, prev_pts, pt_counts, max_pts, next_pts,
                                                          status, err, iters, (int)pw);
    } else {
        LKF_FUNC_ATTRS(LKF_CUDA(cudaFuncSetAttribute(k_lk_track<long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)));
        k_lk_track<long long><<<grid, LK_WARPS * 32, smem, s>>>(lv, n_levels, p->win_w, p->win_h, max_count, make_eps(eps), p->flags,
                                                                p->min_eig_threshold, prev_pts, pt_counts, max_pts,
                                                                next_pts, status, err, iters, (int)pw);
    }
    LKF_LAUNCH_CHECK(ctx);
    return LKF_OK;
}
